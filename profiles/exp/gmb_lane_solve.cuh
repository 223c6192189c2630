// profiles/exp/gmb_lane_solve.cuh -- MEASURED NEGATIVE RESULT (round 2), kept out of the library.
//
// Speculative fused PLU solve with ONE THREAD PER SYSTEM for n = 5..8: rows never move, finished rows become NaN, the
// pivot row is gathered with one select chain per column into a static register row of U, systems with ties / zero
// columns / non-finite values / operands outside the fast division window are handed to an exact list kernel.
// Bit-identical to the exact kernels on hostile 2^18-system batches (it was wired into gmb_plu_solve_* behind
// gmb_set_option("lanesolve", 1) and tested there), but SLOWER than k_sub_plu, which exchanges rows with two selects
// per element and candidate row (profiles/r2_lane_kernels.md; fraction of the HBM peak, SoA container, float):
//     n           5      6      7      8
//     k_sub_plu   0.79   0.70   0.77   0.49 (row kernel)
//     this        0.61   0.39   0.39   0.31
// ncu (n = 7, SoA): the same 34 M warp instructions per 2^20 systems as k_sub_plu (35.7 M) -- the gather saves selects,
// updating all n rows in every step spends them again on FMAs -- but issue slots 40 % busy against 67 %: the gather
// is a chain of n-1 DEPENDENT selects per column where the exchange is n-1-i independent swaps, and 106..168 registers
// leave 3 CTAs per SM.  To build: include after gmb_lane.cuh (uses lane_load / lane_store / LaneWin / lane_fma2 there).
#pragma once
#include "gmb_lane.cuh"

namespace gmb {

// k_lane_solve -- speculative fused PLU solve (linear_solve, LUDecomp.h:388-399; decomp_plu :188-250 + the permutation
//   :35 + backsolve_lu :332-366 in one pass), m = 1..4 right-hand sides riding along as extra columns.  Same idea as
//   gmb_rowfast.cuh, without the lanes: ROWS NEVER MOVE.  Per elimination step i
//     * pivot search (:205-213) = max |a(.,i)| over the N register rows (FMNMX); a row that has already served as a
//       pivot row is all NaN from column i on, and fmax / fmin ignore NaN, so finished rows drop out for free;
//     * the pivot row is the row whose |a(.,i)| equals the maximum; it is gathered with one select chain per column
//       into U(i, i..) -- a STATIC register row, which is exactly where the reference's exchange would have put it;
//     * multipliers (:234): one reciprocal per pivot, three FMA-pipe operations per quotient (RcpOf, gmb_math.cuh:
//       the in-line part of IEEE division, same bits); the pivot row's own multiplier is replaced by NaN, so the
//       update below finishes the row;
//     * rank-1 update (:241) and the riders' forward sweep (:350) as plain FMAs on all N rows, two columns per
//       instruction (fma.rn.f32x2).
//   The reference's exchange costs 2 selects per element and candidate row on the half-rate ALU pipe (k_sub_plu, n = 8:
//   582 FSEL + 230 compare / logic of 2064 instructions per system, against 451 FFMA); the gather costs one.
//   After the last step the single unfinished row is row n-1 of U; the backward sweep (:356-365) is statically indexed.
//   A system LEAVES the fast path -- nothing is written for it, its index goes to a list that g_solve_list
//   (gmb_generic.cu) solves with the reference's exact rules -- when a column maximum is zero or not unique (ties need
//   the reference's position rule, a zero column its degenerate-column handling: caught by the window test / by the
//   count of unfinished rows at the end), when a pivot or a dividend is outside the window in which the
//   three-operation quotient is the IEEE quotient (this includes every zero multiplier, so the reference's `b != 0`
//   test, :235, is always true on the fast path), or when a NaN / Inf meets a comparison.  Results are bit-identical
//   to the exact kernels for every input; only the time differs (tests force both).
//

// ================================================================== speculative fused PLU solve
template <typename T, int N, int M, bool SOA, bool RM, int MINB>
__global__ void __launch_bounds__(kLaneThreads, MINB)
k_lane_solve(const T* __restrict__ A, const T* Bm, T* X, uint8_t* __restrict__ ok_out, int skip, int64_t B,
             int32_t* __restrict__ redo_list, int* __restrict__ redo_count) {
    constexpr int NC = N + M;                                  // augmented row: n columns + the riders
    constexpr int NCX = (NC + 1) / 2 * 2;                      // ... padded to whole column pairs
    using LW = LaneWin<T>;
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * kLaneThreads + threadIdx.x;
    if (s - lane >= B) return;                                 // warp-uniform: the staged ingest is cooperative
    const bool valid = s < B;
    const bool full = (s - lane) + 32 <= B;

    // ---- ingest ---------------------------------------------------------------------------------------------
    unsigned char* sm = nullptr;
    if constexpr (!SOA) sm = lane_stage_buffer<T, (N > M ? N * N : N * M)>();
    T a[N][NCX];
    {
        T qa[N * N];
        lane_load<T, N * N, SOA>(A, s, lane, full, valid, qa, sm);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = qa[RM ? N * r + c : N * c + r];
        T qb[N * M];
        lane_load<T, N * M, SOA>(Bm, s, lane, full, valid, qb, sm);
#pragma unroll
        for (int r = 0; r < N; ++r) {
#pragma unroll
            for (int j = 0; j < M; ++j) a[r][N + j] = qb[RM ? M * r + j : N * j + r];
            if (NCX > NC) a[r][NC] = (T)0;
        }
    }
    bool bad = false;
    T u[N][NCX];                                               // U and the forward-substituted riders, BY POSITION

    // ---- elimination ------------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        // pivot search (:205-213): finished rows are NaN and ignored by fmax / fmin
        T m = abs_(a[0][i]), mn = m;
#pragma unroll
        for (int k = 1; k < N; ++k) {
            const T av = abs_(a[k][i]);
            m = lane_max(m, av);
            mn = lane_min(mn, av);
        }
        // no unfinished row at all (an earlier tie finished two rows at once) leaves m = NaN: caught here
        bad = bad || !(m >= LW::lo() && m < LW::hi()) || (mn < LW::lo());
        // gather the pivot row into U(i, i..) and the riders: row by row, so that only ONE "is the pivot row" predicate is
        // live at a time (eight of them at once spill to a register and cost a LOP3 + ISETP per select); column i first
        // -- the reciprocal of the pivot starts while the rest of the row is still being gathered
        {
            T v = a[0][i];
#pragma unroll
            for (int k = 1; k < N; ++k) v = (abs_(a[k][i]) == m) ? a[k][i] : v;
            u[i][i] = v;
        }
        const RcpOf<T> rc(u[i][i]);
#pragma unroll
        for (int c = i + 1; c < NC; ++c) u[i][c] = a[0][c];
#pragma unroll
        for (int k = 1; k < N; ++k) {
            const bool isk = abs_(a[k][i]) == m;
#pragma unroll
            for (int c = i + 1; c < NC; ++c) u[i][c] = isk ? a[k][c] : u[i][c];
        }
        if (NCX > NC) u[i][NC] = (T)0;
        // multipliers (:234); the pivot row's own becomes NaN
        T nb[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            bool unused;
            const T q = rc.fast(a[k][i], unused);
            nb[k] = (abs_(a[k][i]) == m) ? LW::nan() : -q;
        }
        // rank-1 update (:241) and the riders (:350), two columns at a time; for even i the first pair also rewrites
        // the dead column i (2 * j0 >= i: every U entry used below has been gathered)
        const int j0 = (i + 1) / 2;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int j = j0; j < NCX / 2; ++j) lane_fma2(a[k][2 * j], a[k][2 * j + 1], nb[k], u[i][2 * j], u[i][2 * j + 1]);
        }
    }
    // ---- the one unfinished row is row n-1 -------------------------------------------------------------------------
    {
        int live = 0;
        bool is[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            is[k] = a[k][N - 1] == a[k][N - 1];
            live += is[k] ? 1 : 0;
        }
        bad = bad || (live != 1);
#pragma unroll
        for (int c = N - 1; c < NC; ++c) {
            T v = a[0][c];
#pragma unroll
            for (int k = 1; k < N; ++k) v = is[k] ? a[k][c] : v;
            u[N - 1][c] = v;
        }
    }
    if (bad) {
        if (valid) redo_list[atomicAdd(redo_count, 1)] = (int32_t)s;
    }

    // ---- backward sweep (:356-365), statically indexed -----------------------------------------------------------
    T x[N][M];
#pragma unroll
    for (int r = N - 1; r >= 0; --r) {
        const RcpOf<T> rc(u[r][r]);
#pragma unroll
        for (int j = 0; j < M; ++j) {
            T v = u[r][N + j];
#pragma unroll
            for (int c = N - 1; c > r; --c) v = fma_rn(-x[c][j], u[r][c], v);
            x[r][j] = (r >= skip) ? quot(v, rc) : u[r][N + j];  // rows < skip stay forward-substituted (:356)
        }
    }

    // ---- egress (nothing for a system that left the fast path: the exact kernel writes x and ok for it) -------------
    // the staged store is cooperative: every lane takes part, a lane whose system left the fast path re-stores what the
    // destination must keep -- which is not known here (x may alias b), so warps with such a lane store per thread
    const bool any_bad = __any_sync(0xffffffffu, bad);
    T qx[N * M];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int j = 0; j < M; ++j) qx[RM ? M * r + j : N * j + r] = x[r][j];
    lane_store<T, N * M, SOA>(X, s, lane, full && !any_bad, valid && !bad, qx, sm);
    if (valid && !bad && ok_out) ok_out[s] = 1;                // a degenerate column never stays on the fast path
}

template <typename T> int generic_solve_list(int, int, const T*, const T*, T*, uint8_t*, int, int, int, const int32_t*, const int*,
                                             int64_t, cudaStream_t);

// registers: a[N][N+M] + U; three CTAs of 128 threads per SM (168 registers) for the larger blocks, four below
template <typename T, int N, int M> struct LaneCfg {
    static constexpr int words = (N * (N + M) + N * (N + 1) / 2 + N * M) * (int)sizeof(T) / 4;
    static constexpr int minb = words <= 80 ? 4 : (words <= 110 ? 3 : 2);
};

inline void lane_keep_pool_memory() {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done.fetch_or(bit, std::memory_order_relaxed);
}

template <typename T, int NMAX>
int lane_solve_impl(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, int skip, int layout, int rm,
                    cudaStream_t st) {
    if (B >= (int64_t)1 << 31) return GMB_ERR_UNSUPPORTED;     // 32-bit list entries
    if (n > NMAX) return GMB_ERR_UNSUPPORTED;
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8>(n, [&](auto N) {
        if constexpr (N.value <= NMAX) {
            dispatch_int<1, 2, 3, 4>(m, [&](auto M) {
                if (B == 0) {
                    rc = GMB_OK;
                    return;
                }
                lane_keep_pool_memory();
                int32_t* scratch = nullptr;
                if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)(B + 4) * sizeof(int32_t), st) != cudaSuccess) {
                    cudaGetLastError();
                    rc = GMB_ERR_ALLOC;
                    return;
                }
                int* count = reinterpret_cast<int*>(scratch);
                int32_t* list = scratch + 4;
                cudaMemsetAsync(count, 0, sizeof(int), st);
                dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                    dispatch_bool(rm != 0, [&](auto RM) {
                        constexpr int MINB = LaneCfg<T, N.value, M.value>::minb;
                        const unsigned grid = (unsigned)ceil_div(B, kLaneThreads);
                        k_lane_solve<T, N.value, M.value, SOA.value, RM.value, MINB><<<grid, kLaneThreads, 0, st>>>(A, Bm, X, ok, skip, B, list, count);
                        g_launch_count.fetch_add(1, std::memory_order_relaxed);
                    });
                });
                // exact re-solve of the systems handed back (almost always none)
                rc = generic_solve_list<T>(n, m, A, Bm, X, ok, skip, layout, rm, list, count, B, st);
                cudaFreeAsync(scratch, st);
            });
        }
    });
    return rc;
}


}  // namespace gmb
