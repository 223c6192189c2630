// gmb_lane.cuh -- ONE THREAD PER SYSTEM for n = 5..8 (float) / 5..6 (double): the whole augmented matrix of a system
// lives in one thread's registers, so there are no shuffles and no replicated bookkeeping; the only cooperative step is
// the ingest / egress of AoS records (a warp moves its 32 records as ONE contiguous range with coalesced 128-bit
// accesses through a warp-private shared-memory transpose, aos_load_staged / aos_store_staged in gmb_io.cuh; the SoA
// container needs none: lane l reads word l of every element plane, 128 contiguous bytes per warp instruction).
//
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
// k_lane_chol -- cholesky (Cholesky.h:37-58) and the fused Cholesky solve: the row kernel's arithmetic (one reciprocal
//   per diagonal entry, shared by its column and the two substitution sweeps) with the staged ingest, so that record
//   sizes that are not multiples of 16 bytes (n = 5, 6, 7) no longer load scalars at a stride of one record.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

constexpr int kLaneThreads = 128;          // 4 warps: stage_buffer() holds four warp-private buffers

template <typename T> struct LaneWin;
template <> struct LaneWin<float> {
    __device__ static __forceinline__ float lo() { return 6.938893903907228e-18f; }        // 2^-57
    __device__ static __forceinline__ float hi() { return 1.4411518807585587e17f; }        // 2^57
    __device__ static __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
};
template <> struct LaneWin<double> {
    __device__ static __forceinline__ double lo() { return 3.872591914849318e-121; }       // 2^-400
    __device__ static __forceinline__ double hi() { return 2.582249878086909e120; }        // 2^400
    __device__ static __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
};

__device__ __forceinline__ float  lane_max(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double lane_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  lane_min(float a, float b)   { return fminf(a, b); }
__device__ __forceinline__ double lane_min(double a, double b) { return fmin(a, b); }

// d0 = fma(s, b0, d0); d1 = fma(s, b1, d1): one FFMA2 for float (sm_100a; same IEEE result per element), two DFMAs
__device__ __forceinline__ void lane_fma2(float& d0, float& d1, float s, float b0, float b1) {
    asm("{\n\t.reg .b64 bb, dd, ss;\n\tmov.b64 bb, {%2, %3};\n\tmov.b64 dd, {%0, %1};\n\tmov.b64 ss, {%4, %4};\n\t"
        "fma.rn.f32x2 dd, ss, bb, dd;\n\tmov.b64 {%0, %1}, dd;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(b0), "f"(b1), "f"(s));
}
__device__ __forceinline__ void lane_fma2(double& d0, double& d1, double s, double b0, double b1) {
    d0 = __fma_rn(s, b0, d0);
    d1 = __fma_rn(s, b1, d1);
}

// ---- one record of E elements per thread -----------------------------------------------------------------------
// `s` = this thread's system, `lane` = its lane, `full` = all 32 systems of the warp exist (warp-uniform).
// ONE warp-private staging buffer per kernel, shared by all of its staged loads and stores (they are sequential, each
// ends with a __syncwarp): EMAX = the largest record the kernel moves.
template <typename T, int EMAX>
__device__ __forceinline__ unsigned char* lane_stage_buffer() {
    __shared__ __align__(16) unsigned char buf[kLaneThreads / 32][32 * EMAX * (int)sizeof(T)];
    return buf[threadIdx.x >> 5];
}

template <typename T, int E, bool SOA>
__device__ __forceinline__ void lane_load(const T* __restrict__ base, int64_t s, int lane, bool full, bool valid, T (&q)[E],
                                          unsigned char* sm) {
    if constexpr (SOA) {
        constexpr int W = soa_w<T>();
        const T* p = base + (s / W) * (int64_t)(E * W) + (s % W);
#pragma unroll
        for (int e = 0; e < E; ++e) q[e] = valid ? __ldg(p + e * W) : (T)0;
    } else {
        if (StageCfg<E * (int)sizeof(T)>::use && full) {
            aos_load_staged<T, E>(base, s - lane, lane, q, sm);
        } else if (valid) {
            aos_load<T, E>(base, s, q);
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) q[e] = (T)0;
        }
    }
}
template <typename T, int E, bool SOA>
__device__ __forceinline__ void lane_store(T* __restrict__ base, int64_t s, int lane, bool full, bool valid, const T (&q)[E],
                                           unsigned char* sm) {
    if constexpr (SOA) {
        constexpr int W = soa_w<T>();
        T* p = base + (s / W) * (int64_t)(E * W) + (s % W);
        if (valid) {
#pragma unroll
            for (int e = 0; e < E; ++e) p[e * W] = q[e];
        }
    } else {
        if (StageCfg<E * (int)sizeof(T)>::use && full) {
            aos_store_staged<T, E>(base, s - lane, lane, q, sm);
        } else if (valid) {
            aos_store<T, E>(base, s, q);
        }
    }
}

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

// ================================================================== Cholesky (+ fused solve)
// Row sweep of Cholesky.h:40-56 in registers; one RcpOf per diagonal entry serves the quotients of its column and the two
// substitution sweeps (same bits as div_rn, gmb_math.cuh).  Only the lower triangle of the input is used; the strict upper
// triangle of the factor is zero (:52).
template <typename T, int N, int M, bool SOA, bool RM, bool SOLVE, int MINB>
__global__ void __launch_bounds__(kLaneThreads, MINB)
k_lane_chol(const T* A, const T* Bm, T* X, T* L_out, uint8_t* __restrict__ ok_out, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * kLaneThreads + threadIdx.x;
    if (s - lane >= B) return;
    const bool valid = s < B;
    const bool full = (s - lane) + 32 <= B;

    unsigned char* sm = nullptr;
    if constexpr (!SOA) sm = lane_stage_buffer<T, (N > M ? N * N : N * M)>();
    T a[N][N];
    if constexpr (SOA) {
        // the container gives element planes: the strict upper triangle is never read (Cholesky.h:42-44)
        constexpr int W = soa_w<T>();
        const T* p = A + (s / W) * (int64_t)(N * N * W) + (s % W);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = (c <= r && valid) ? __ldg(p + (RM ? N * r + c : N * c + r) * W) : (T)0;
    } else {
        T qa[N * N];
        lane_load<T, N * N, false>(A, s, lane, full, valid, qa, sm);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = qa[RM ? N * r + c : N * c + r];
    }
    T x[N][M];
    if constexpr (SOLVE) {
        T qb[N * M];
        lane_load<T, N * M, SOA>(Bm, s, lane, full, valid, qb, sm);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) x[r][j] = qb[RM ? M * r + j : N * j + r];
    }

    bool ok = true;
    T ry[N];                                                   // reciprocal part of RcpOf per diagonal entry
    bool rs[N];
#pragma unroll
    for (int r = 0; r < N; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            T sum = a[r][c];
#pragma unroll
            for (int k = 0; k < c; ++k) sum = fma_rn(-a[r][k], a[c][k], sum);
            if (r == c) {
                ok = ok && (sum > (T)0);
                sum = sqrt_rn(sum);
                const RcpOf<T> rc(sum);
                ry[r] = rc.y;
                rs[r] = rc.safe;
            } else {
                sum = quot(sum, RcpOf<T>(a[c][c], ry[c], rs[c]));
                a[c][r] = (T)0;
            }
            a[r][c] = sum;
        }
    }
    if constexpr (SOLVE) {
        // L y = b forward, L^T x = y backward (cholesky_solve_regs, gmb_math.cuh)
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const RcpOf<T> rc(a[r][r], ry[r], rs[r]);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                T v = x[r][j];
#pragma unroll
                for (int c = 0; c < r; ++c) v = fma_rn(-x[c][j], a[r][c], v);
                x[r][j] = quot(v, rc);
            }
        }
#pragma unroll
        for (int r = N - 1; r >= 0; --r) {
            const RcpOf<T> rc(a[r][r], ry[r], rs[r]);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                T v = x[r][j];
#pragma unroll
                for (int c = N - 1; c > r; --c) v = fma_rn(-x[c][j], a[c][r], v);
                x[r][j] = quot(v, rc);
            }
        }
        T qx[N * M];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) qx[RM ? M * r + j : N * j + r] = x[r][j];
        lane_store<T, N * M, SOA>(X, s, lane, full, valid, qx, sm);
    }
    if (L_out != nullptr) {
        T ql[N * N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) ql[RM ? N * r + c : N * c + r] = a[r][c];
        lane_store<T, N * N, SOA>(L_out, s, lane, full, valid, ql, sm);
    }
    if (valid && ok_out) ok_out[s] = ok ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------- launchers
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

template <typename T, int NMAX>
int lane_chol_impl(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm, bool solve,
                   cudaStream_t st) {
    if (n > NMAX || (solve && (m < 1 || m > 4))) return GMB_ERR_UNSUPPORTED;
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8>(n, [&](auto N) {
        if constexpr (N.value <= NMAX) {
            dispatch_int<1, 2, 3, 4>(solve ? m : 1, [&](auto M) {
                const unsigned grid = (unsigned)ceil_div(B, kLaneThreads);
                if (grid == 0) {
                    rc = GMB_OK;
                    return;
                }
                dispatch_bool(solve, [&](auto SOLVE) {
                    if constexpr (!SOLVE.value && M.value != 1) return;
                    else
                    dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                        dispatch_bool(rm != 0, [&](auto RM) {
                            constexpr int MINB = (N.value * N.value * (int)sizeof(T) / 4 <= 50) ? 4 : 3;
                            k_lane_chol<T, N.value, M.value, SOA.value, RM.value, SOLVE.value, MINB><<<grid, kLaneThreads, 0, st>>>(A, Bm, X, L_out, ok, B);
                        });
                    });
                });
                rc = after_launch();
            });
        }
    });
    return rc;
}

}  // namespace gmb
