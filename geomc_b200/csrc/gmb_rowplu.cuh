// gmb_rowplu.cuh -- row-distributed cooperative fused PLU solve (linear_solve, LUDecomp.h:388) for
// the upper end of the fixed-size range (n = 9..16, one right-hand side).
//
// The column-distributed kernel (gmb_subwarp.cuh) keeps whole columns in a lane, so a row exchange
// costs two selects per (candidate row, local column) -- for n = 16 that alone is ~1200 selects per
// lane, and the pivot search and the multipliers are computed redundantly by every lane of the
// group.  Here the ROWS are distributed: row p of [A | b] lives in lane p % G, register slot p / G
// (RL = ceil(n / G) slots of n + 1 registers).  Consequences, per elimination step i:
//   * pivot search: each lane scans its own slots (position order = slot order, strict `>`), then a
//     log2(G)-round shuffle reduction with "larger value, then lower position" -- exactly the
//     reference's first strict maximum (LUDecomp.h:205-213); NaN keys are mapped so that a NaN
//     diagonal is never displaced and a NaN candidate never wins;
//   * row exchange i <-> pvt (whole rows, L part and rider included, :222): the two rows travel
//     through a 2-row shared-memory bounce buffer private to the group.  Row i sits in a static
//     (lane, slot); row pvt is addressed with lane/slot predicates on 128-bit shared accesses, so no
//     register is ever indexed dynamically and there are no select chains;
//   * every lane reads the pivot row from the bounce buffer (broadcast), computes the multipliers of
//     ITS OWN rows only (one IEEE division per live row, no redundancy) and updates them:
//     a(r,c) = fma(-b, a(i,c), a(r,c)), skipped when b == 0 (:235-241); the rider column is updated
//     unconditionally (= backsolve_lu's forward sweep, :345-353);
//   * slots whose rows are all above the diagonal are skipped statically.
// Two exchange forms (template flag SHFLX; the dispatcher picks per (T, n), measured):
//   * bounce buffer (above): LSU-bound -- ncu on 16x16 float: 65 % LSU wavefront utilisation, top stall MIO throttle,
//     because the dynamic slot is reached with predicated 128-bit shared accesses over every candidate slot;
//   * register-only: each lane collapses its candidate slots to one row with a select chain on (ps == k), a shuffle
//     from the pivot's lane broadcasts it, row i is broadcast the same way and dropped into slot ps with predicated
//     moves.  ALU work instead of LSU wavefronts; 16x16 float 1.23e9 -> 1.62e9 systems/s with 4 lanes per system.
// WANT_LU = false (fused solve without an LU output): the L part of a row is dead, so only columns >= i travel.
// Backward sweep: column-oriented (x(c) /= u(c,c); x(r) = fma(-x(c), u(r,c), x(r)) for r < c), rows
// are in position order so everything is statically indexed; one shuffle per column.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

#ifndef GMB_ROW_THREADS
#define GMB_ROW_THREADS 128
#endif
constexpr int kRowThreads = GMB_ROW_THREADS;
#ifndef GMB_ROW_REG_BUDGET
#define GMB_ROW_REG_BUDGET 140     // register words for the row block (picks G); measured: 140 beats 72 by 12-40%
#endif
#ifndef GMB_ROW_MIN_BLOCKS
#define GMB_ROW_MIN_BLOCKS 2
#endif

constexpr int row_rl(int n, int g) { return (n + g - 1) / g; }
constexpr int row_group(int n, int words) {
    for (int g = 1; g <= 16; g *= 2)
        if (row_rl(n, g) * (n + 1) * words <= GMB_ROW_REG_BUDGET) return g;
    return 16;
}

#ifndef GMB_ROW_SRCP_FLOAT_MIN_N
#define GMB_ROW_SRCP_FLOAT_MIN_N 13
#endif

template <typename T> struct InfOf;
template <> struct InfOf<float>  { __device__ static __forceinline__ float get()  { return __int_as_float(0x7f800000); } };
template <> struct InfOf<double> { __device__ static __forceinline__ double get() { return __longlong_as_double(0x7ff0000000000000LL); } };

// One system's solve / decomposition by the G lanes that own it (body of k_row_solve and k_row_solve_list).
template <typename T, int N, bool SOA, bool RM, bool DECOMP, int G, bool WANT_LU, bool SHFLX, int THREADS>
__device__ __forceinline__ void
row_solve_one(const T* __restrict__ A, const T* Bm, T* X, T* __restrict__ LU, uint8_t* __restrict__ ok_out,
              uint8_t* __restrict__ perm_out, uint8_t* __restrict__ parity_out, uint8_t* __restrict__ degen_out, int skip,
              int64_t B, int64_t s, const bool valid, const int64_t s_warp0) {
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int RL = row_rl(N, G);
    constexpr int NC = N + 1;                                  // augmented row: n columns + the rider
    constexpr int V = soa_v<T>();                              // elements per 128-bit access
    constexpr int NCP = (NC + V - 1) / V * V;                  // padded row length in the bounce buffer
    constexpr int NCH = NCP / V;                               // 16-byte chunks per row
    constexpr int GS_RAW = 2 * NCP * WORDS;                    // words per group (two rows)
    constexpr int GS = ((GS_RAW / 4) % 2 == 1) ? GS_RAW : GS_RAW + 4;   // odd number of 16-byte units: conflict-free across groups
    constexpr int ST = SOA ? soa_w<T>() : 1;
    using VT = typename VecOf<T>::type;

    __shared__ __align__(16) uint32_t bounce[SHFLX ? 4 : (THREADS / G) * GS];   // only the bounce-buffer exchange uses it

    const int lane_g = (G == 1) ? 0 : (int)(threadIdx.x % G);
    const int group = (int)(threadIdx.x / G);
    T* buf0 = reinterpret_cast<T*>(bounce + (SHFLX ? 0 : group * GS));       // row i
    T* buf1 = buf0 + NCP;                                      // row pvt

    constexpr int W = soa_w<T>();
    const int64_t baseA = SOA ? (s / W) * (int64_t)(N * N) * W + (s % W) : s * (int64_t)(N * N);
    const int64_t baseX = SOA ? (s / W) * (int64_t)N * W + (s % W) : s * (int64_t)N;

    // ---- load: slot k holds row p = k*G + lane_g ----------------------------------------------
    T a[RL][NC];
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int p = k * G + lane_g;
        const bool live = p < N;
        const int pr = live ? p : 0;
        if constexpr (!SOA && RM && (N * (int)sizeof(T)) % 16 == 0) {
            ldg_row<T, N>(A + baseA + (int64_t)pr * N, a[k], (reinterpret_cast<uintptr_t>(A) & 31) == 0);
        } else {
#pragma unroll
            for (int c = 0; c < N; ++c) a[k][c] = A[baseA + (int64_t)(RM ? N * pr + c : N * c + pr) * ST];
        }
        if (DECOMP) a[k][N] = (T)pr;                           // the row's source index rides along
        else a[k][N] = Bm[baseX + (int64_t)pr * ST];           // m = 1: same address in both layouts
    }

    int degenerate = 0;
    bool parity = false;

    // ---- elimination ----------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        const int oi_lane = i % G;                             // static
        const int oi_slot = i / G;                             // static

        // pivot search in column i over positions >= i
        T best_v = (T)-2;
        int best_p = 1 << 20;
        // position tests against the static step: slot k > oi_slot lies wholly below row i, slot oi_slot is split by
        // the lane; a slot is "live" statically unless it is the ragged last one (fewer compares on the ALU pipe)
        auto row_live = [&](int k) { return ((k + 1) * G <= N) || (k * G + lane_g < N); };
        auto at_or_below = [&](int k) { return (k > oi_slot || lane_g >= oi_lane) && row_live(k); };
        auto is_below = [&](int k) { return (k > oi_slot || lane_g > oi_lane) && row_live(k); };
#pragma unroll
        for (int k = oi_slot; k < RL; ++k) {
            const int p = k * G + lane_g;
            const bool cand = at_or_below(k);
            T v = abs_(a[k][i]);
            // NaN: `v > best_v` is false, so a NaN below the diagonal never wins; a NaN ON the diagonal keeps the pivot
            // (:205-213 start from |m(i,i)| and nothing compares greater than NaN): it becomes +Inf here
            if (k == oi_slot) v = ((lane_g == oi_lane) && (v != v)) ? InfOf<T>::get() : v;
            const bool take = cand && (v > best_v);
            best_v = take ? v : best_v;
            best_p = take ? p : best_p;
        }
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            const T ov = __shfl_xor_sync(0xffffffffu, best_v, off, G);
            const int op = __shfl_xor_sync(0xffffffffu, best_p, off, G);
            const bool take = (ov > best_v) || (ov == best_v && op < best_p);
            best_v = take ? ov : best_v;
            best_p = take ? op : best_p;
        }
        const int pvt = best_p;
        const bool dead = (best_v == (T)0);                    // :215
        degenerate += dead ? 1 : 0;
        const bool do_swap = pvt != i;
        parity = parity != do_swap;
        const int pl = pvt % G, ps = pvt / G;
        // without an LU output the L part of a row (columns < i) is dead: only the 16-byte chunks from column i on travel
        const int Q0 = WANT_LU ? 0 : i / V;

        T piv[NC];
        if constexpr (SHFLX) {
            // Register-only exchange.  The pivot row sits in a slot that is only known at run time: each lane first
            // collapses its candidate slots to ONE row with a select chain on (ps == k), then a shuffle from the
            // pivot's lane broadcasts it; row i (static lane and slot) is broadcast the same way and the pivot's
            // former owner drops it into slot ps with predicated moves.  No shared memory, no 128-bit register
            // quads to keep aligned; costs ALU work instead of LSU wavefronts (the bounce-buffer form is LSU-bound).
            const int c0 = WANT_LU ? 0 : i;
            const int gbase = (int)(threadIdx.x & 31) & ~(G - 1);
            const bool mine_is_pvt = do_swap && (lane_g == pl);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c >= c0) {
                    T v = a[oi_slot][c];
#pragma unroll
                    for (int k = oi_slot + 1; k < RL; ++k) v = (ps == k) ? a[k][c] : v;
                    const T pc = __shfl_sync(0xffffffffu, v, gbase + pl);          // row pvt (= row i when !do_swap: ps, pl static then)
                    const T ri = __shfl_sync(0xffffffffu, a[oi_slot][c], gbase + oi_lane);   // row i
                    piv[c] = pc;
#pragma unroll
                    for (int k = oi_slot; k < RL; ++k) {
                        const bool to_pvt = mine_is_pvt && (ps == k);
                        const bool to_i = do_swap && (lane_g == oi_lane) && (k == oi_slot);
                        a[k][c] = to_pvt ? ri : (to_i ? pc : a[k][c]);
                    }
                }
            }
        } else {
        // rows i and pvt -> bounce buffer (row i always: it is also the broadcast copy of the pivot row)
        if (lane_g == oi_lane) {
#pragma unroll
            for (int q = 0; q < NCH; ++q) if (q >= Q0) {
                T t[V];
#pragma unroll
                for (int v = 0; v < V; ++v) t[v] = (q * V + v < NC) ? a[oi_slot][q * V + v] : (T)0;
                reinterpret_cast<VT*>(buf0)[q] = pack(t);
            }
        }
#pragma unroll
        for (int k = oi_slot; k < RL; ++k) {
            if (do_swap && lane_g == pl && ps == k) {
#pragma unroll
                for (int q = 0; q < NCH; ++q) if (q >= Q0) {
                    T t[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) t[v] = (q * V + v < NC) ? a[k][q * V + v] : (T)0;
                    reinterpret_cast<VT*>(buf1)[q] = pack(t);
                }
            }
        }
        __syncwarp();
        // pivot row (after the exchange) for everybody; the two owners also take their new rows
        const T* prow = do_swap ? buf1 : buf0;
#pragma unroll
        for (int q = 0; q < NCH; ++q) if (q >= Q0) {
            T t[V];
            unpack(reinterpret_cast<const VT*>(prow)[q], t);
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (q * V + v < NC) piv[q * V + v] = t[v];
        }
        if (do_swap && lane_g == oi_lane) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c >= Q0 * V) a[oi_slot][c] = piv[c];
        }
#pragma unroll
        for (int k = oi_slot; k < RL; ++k) {
            if (do_swap && lane_g == pl && ps == k) {
#pragma unroll
                for (int q = 0; q < NCH; ++q) if (q >= Q0) {
                    T t[V];
                    unpack(reinterpret_cast<const VT*>(buf0)[q], t);
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (q * V + v < NC) a[k][q * V + v] = t[v];
                }
            }
        }
        __syncwarp();

        }
        // multipliers and rank-1 update of this lane's own rows below the pivot
        // multipliers: one reciprocal per pivot, three operations per quotient (RcpOf, gmb_math.cuh); a lane whose
        // operands leave the fast domain redoes its quotients with div_rn -- same bits either way.  Measured
        // (profiles/r1l_shared_reciprocal.md): double +6..21 %, float n >= 13 +4..7 %, float n <= 12 -3..-6 % (the
        // window tests cost what the shorter chains save), so small float sizes keep the plain division.
        constexpr bool SRCP = sizeof(T) == 8 ? (G >= 2 && !(DECOMP && N == 12)) : N >= GMB_ROW_SRCP_FLOAT_MIN_N;   // measured exceptions: one-lane double, 12 x 12 double decomp
        const T pv = piv[i];
        T bq[RL];
        if constexpr (SRCP) {
            const RcpOf<T> rc(pv);
            bool all_ok = true;
#pragma unroll
            for (int k = oi_slot; k < RL; ++k) {
                bool ok;
                bq[k] = rc.fast(a[k][i], ok);
                all_ok = all_ok && (ok || !is_below(k));
            }
            if (!all_ok) {
#pragma unroll
                for (int k = oi_slot; k < RL; ++k) bq[k] = div_rn(a[k][i], pv);
            }
        }
#pragma unroll
        for (int k = oi_slot; k < RL; ++k) {
            const bool below = is_below(k);
            const T b = SRCP ? bq[k] : div_rn(a[k][i], pv);    // :234
            const bool upd = below && !dead && (b != (T)0);    // :235
            if (below && !dead) a[k][i] = b;                   // :245
#pragma unroll
            for (int c = i + 1; c < N; ++c) {
                const T v = fma_rn(-b, piv[c], a[k][c]);       // :241
                a[k][c] = upd ? v : a[k][c];
            }
            if (!DECOMP) {                                     // rider: forward sweep, never skipped (:350)
                const T v = fma_rn(-piv[N], b, a[k][N]);
                a[k][N] = below ? v : a[k][N];
            }
        }
    }

    const bool good = degenerate == 0;

    // ---- optional: the packed LU the reference leaves in `a` --------------------------------------
    // AoS: per-lane row stores would write partial sectors at a stride of one system; the warp's 32/G
    // systems go through a staging buffer instead and leave with coalesced 128-bit stores (gmb_io.cuh)
    extern __shared__ __align__(16) unsigned char gmb_dyn_smem[];
    constexpr int NSYS = 32 / G;
    constexpr int SA = stage_stride<N * N>();
    if (WANT_LU && !SOA && LU != nullptr && s_warp0 + NSYS <= B) {
        T* smw = reinterpret_cast<T*>(gmb_dyn_smem) + (threadIdx.x >> 5) * (NSYS * SA);
        T* smd = smw + ((threadIdx.x & 31) / G) * SA;
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int p = k * G + lane_g;
            if (p < N) {
#pragma unroll
                for (int c = 0; c < N; ++c) smd[RM ? N * p + c : N * c + p] = a[k][c];
            }
        }
        warp_stage_out<T, N * N, NSYS>(LU + s_warp0 * (int64_t)(N * N), smw, threadIdx.x & 31);
    } else if (WANT_LU && LU != nullptr && valid) {
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int p = k * G + lane_g;
            if (p < N) {
#pragma unroll
                for (int c = 0; c < N; ++c) LU[baseA + (int64_t)(RM ? N * p + c : N * c + p) * ST] = a[k][c];
            }
        }
    }

    if constexpr (DECOMP) {
        if (valid) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int p = k * G + lane_g;
                if (p < N && perm_out) perm_out[s * N + p] = (uint8_t)(int)a[k][N];
            }
            if (lane_g == 0) {
                if (parity_out) parity_out[s] = parity ? 1 : 0;
                if (degen_out) degen_out[s] = (uint8_t)degenerate;
            }
        }
    } else {
    // ---- backward sweep, column-oriented (:356-365) -----------------------------------------------
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        const int oc_lane = c % G, oc_slot = c / G;            // static
        const bool on = c >= skip;
        const T q = div_rn(a[oc_slot][N], a[oc_slot][c]);
        T xc = (G == 1) ? q : __shfl_sync(0xffffffffu, q, oc_lane, G);
        if (on && lane_g == oc_lane) a[oc_slot][N] = q;
#pragma unroll
        for (int k = 0; k <= oc_slot; ++k) {
            const int p = k * G + lane_g;
            const bool row_on = on && (p < c) && (p >= skip);
            const T v = fma_rn(-xc, a[k][c], a[k][N]);
            a[k][N] = row_on ? v : a[k][N];
        }
    }

    // ---- store ------------------------------------------------------------------------------------
    if (valid) {
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int p = k * G + lane_g;
            if (p < N) {
                const int64_t off = baseX + (int64_t)p * ST;
                X[off] = good ? a[k][N] : Bm[off];             // x untouched on failure (:391-393)
            }
        }
        if (lane_g == 0 && ok_out) ok_out[s] = good ? 1 : 0;
    }
    }   // !DECOMP
}

// DECOMP = false: fused linear_solve (rider = right-hand side).
// DECOMP = true : decomp_plu in place (LUDecomp.h:188): the rider carries the row's ORIGINAL index,
//                 so after the exchanges slot p holds reorder[p]; no rider arithmetic, no backward sweep.
//                 (A is then read and written through `LU`; perm / parity / degenerate counts are stored.)
template <typename T, int N, bool SOA, bool RM, bool DECOMP, int G = row_group(N, (int)sizeof(T) / 4), int MINB = GMB_ROW_MIN_BLOCKS,
          bool WANT_LU = true, bool SHFLX = false, int THREADS = kRowThreads>
__global__ void __launch_bounds__(THREADS, MINB)
k_row_solve(const T* __restrict__ A, const T* Bm, T* X, T* __restrict__ LU, uint8_t* __restrict__ ok_out,
            uint8_t* __restrict__ perm_out, uint8_t* __restrict__ parity_out, uint8_t* __restrict__ degen_out, int skip,
            int64_t B) {
    int64_t s = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;                                     // keep all lanes for the shuffles / syncwarp
    const int64_t s_warp0 = ((int64_t)blockIdx.x * THREADS + (threadIdx.x & ~31u)) / G;
    row_solve_one<T, N, SOA, RM, DECOMP, G, WANT_LU, SHFLX, THREADS>(A, Bm, X, LU, ok_out, perm_out, parity_out, degen_out, skip, B, s,
                                                                    valid, s_warp0);
}

// The same solve for the systems named in list[0 .. *count): the exact re-solve of the systems that the speculative
// kernels (gmb_rowfast.cuh) could not finish on their fast path.  A small persistent grid walks the list.
template <typename T, int N, bool SOA, bool RM, int G, int MINB, bool SHFLX, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB)
k_row_solve_list(const T* __restrict__ A, const T* Bm, T* X, uint8_t* __restrict__ ok_out, int skip, int64_t B,
                 const int32_t* __restrict__ list, const int* __restrict__ count) {
    const int64_t cnt = *count;
    const int64_t stride = (int64_t)gridDim.x * THREADS / G;
    for (int64_t j = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / G;; j += stride) {
        const bool valid = j < cnt;
        if (!__any_sync(0xffffffffu, valid)) break;
        const int64_t s = valid ? (int64_t)list[j] : (B - 1);
        row_solve_one<T, N, SOA, RM, false, G, false, SHFLX, THREADS>(A, Bm, X, nullptr, ok_out, nullptr, nullptr, nullptr, skip, B,
                                                                     s, valid, B);
    }
}

// dynamic shared memory of the AoS output staging: one buffer of 32/G systems per warp
template <typename T, int N>
constexpr size_t row_stage_bytes() {
    return (size_t)(kRowThreads / 32) * (32 / row_group(N, (int)sizeof(T) / 4)) * stage_stride<N * N>() * sizeof(T);
}

// Exchange variant per (T, n) for the call WITHOUT an LU output (measured, profiles/r1h_row_exchange.md; 2^20 systems):
// float: shuffle exchange with 2 lanes per system up to n = 14, with 4 lanes and 256-thread CTAs for n = 15, 16;
// double: shuffle exchange with 2 lanes for n = 10..14 (256-thread CTAs, one per SM's register file, from n = 13),
// bounce buffer (trailing chunks only) for n = 9, 15, 16.  Below 9 the column-distributed kernel is the default; the row
// kernel takes 8 x 8 (2 lanes: float 0.68 -> 0.39 ms, double 1.38 -> 0.92 ms per 2^22 systems) and double 7 x 7 (one lane).
template <typename T, int N> struct RowNoLu {
    static constexpr int WORDS = (int)sizeof(T) / 4;
    static constexpr bool shfl = WORDS == 1 ? true : (N <= 8 || (N >= 10 && N <= 14));
    static constexpr int G = !shfl ? row_group(N, WORDS) : (N <= 7 ? 1 : ((WORDS == 1 && N >= 15) ? 4 : 2));
    static constexpr int threads = (shfl && ((WORDS == 1 && G == 4) || (WORDS == 2 && N >= 13))) ? 256 : kRowThreads;
    static constexpr int minb = (WORDS == 2 && threads == 256) ? 1 : GMB_ROW_MIN_BLOCKS;
};

template <typename T>
int rowplu_solve_impl(int n, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* LU, int skip, int layout, int rm,
                      cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        if (B == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            dispatch_bool(rm != 0, [&](auto RM) {
                if (LU == nullptr) {
                    using P = RowNoLu<T, N.value>;
                    auto kern = k_row_solve<T, N.value, SOA.value, RM.value, false, P::G, P::minb, false, P::shfl, P::threads>;
                    const unsigned grid = (unsigned)ceil_div(B * P::G, P::threads);
                    kern<<<grid, P::threads, 0, st>>>(A, Bm, X, nullptr, ok, nullptr, nullptr, nullptr, skip, B);
                } else {
                    constexpr int G = row_group(N.value, (int)sizeof(T) / 4);
                    const unsigned grid = (unsigned)ceil_div(B * G, kRowThreads);
                    auto kern = k_row_solve<T, N.value, SOA.value, RM.value, false>;
                    const size_t smem = SOA.value ? 0 : row_stage_bytes<T, N.value>();
                    if (smem > 0) {   // static + dynamic may exceed the 48 KB default
                        static std::atomic<uint64_t> raised{0};   // per instantiation, one bit per device
                        raise_dyn_smem(kern, smem, raised);
                    }
                    kern<<<grid, kRowThreads, smem, st>>>(A, Bm, X, LU, ok, nullptr, nullptr, nullptr, skip, B);
                }
            });
        });
        rc = after_launch();
    });
    return rc;
}

template <typename T>
int rowplu_decomp_impl(int n, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout, int rm,
                       cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        constexpr int G = row_group(N.value, (int)sizeof(T) / 4);
        const unsigned grid = (unsigned)ceil_div(B * G, kRowThreads);
        if (grid == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            dispatch_bool(rm != 0, [&](auto RM) {
                auto kern = k_row_solve<T, N.value, SOA.value, RM.value, true>;
                const size_t smem = SOA.value ? 0 : row_stage_bytes<T, N.value>();
                if (smem > 0) {   // static + dynamic may exceed the 48 KB default
                    static std::atomic<uint64_t> raised{0};   // per instantiation, one bit per device
                    raise_dyn_smem(kern, smem, raised);
                }
                kern<<<grid, kRowThreads, smem, st>>>(A, nullptr, nullptr, A, nullptr, perm, parity, degen, 0, B);
            });
        });
        rc = after_launch();
    });
    return rc;
}

}  // namespace gmb
