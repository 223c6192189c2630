// gmb_subwarp.cuh -- sub-warp cooperative PLU kernels for n = 5..16 (instantiated per element type
// and mode in gmb_subwarp_{solve,decomp,inv}_{f32,f64}.cu so that the translation units build in parallel).
//
// A group of G lanes (G = 1, 2, 4, 8 -- the smallest that keeps the working set in registers)
// owns one system.  The augmented matrix [A | riders] (N rows, N + XC columns) is split into
// blocks of LC = ceil((N + XC) / G) consecutive columns; lane l of the group holds all N rows of
// its LC columns in registers.  Per elimination step i (fully unrolled, every register index
// static):
//   * the lane owning column i searches the pivot (first strict maximum of |a(p,i)|, p >= i,
//     LUDecomp.h:205-213) and broadcasts it with ONE shuffle as a one-hot row mask (+ the
//     `biggest == 0` flag) -- a mask, not an index, so the exchange below stays in registers;
//   * every lane exchanges rows i and pvt in its own columns (whole rows, L part included, :222);
//   * the owner computes the multipliers b_r = a(r,i) / a(i,i) (IEEE division, :234) and
//     broadcasts them (N-1-i shuffles); every lane applies the rank-1 update
//     a(r,c) = fma(-b_r, a(i,c), a(r,c)) to its own columns c > i, skipped where b_r == 0 (:235-241).
// Riders (right-hand sides for linear_solve, the permuted identity for invNxN) are extra
// columns: exchanged in lock-step (= destructive_apply_permutation, :35) and eliminated with the
// same multipliers without the b == 0 skip (= backsolve_lu's forward sweep, :345-353).
// The backward sweep is column-oriented: for c = N-1..0, x(c) /= u(c,c), then x(r) =
// fma(-x(c), u(r,c), x(r)) for r < c -- for every row the terms arrive in the reference's
// order (c descending, :356-365), so the result is bit-identical.
//
// No tensor cores: the systems are independent, 16x16 at most, and the pivot search is
// sequential in i; this is FFMA/DFMA + IEEE division + SHFL.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

constexpr int kSubThreads = 128;
#ifndef GMB_SUB_MIN_BLOCKS
#define GMB_SUB_MIN_BLOCKS 3      // cap at 128 registers: occupancy matters more than a few spills here (ncu: 12% -> 25%)
#endif
// Register words for the local block [A | riders] of one lane (picks G, the lanes per system).  Measured in round 2
// (profiles/r2p_budget112.md): 132 words for the kernels that carry riders puts float n <= 11 (m = 1), float n <= 10 (m = 3) and
// double n = 7 (m = 1) on ONE lane per system -- with the staged I/O and without the exchange of dead multiplier columns that
// beats two lanes and the row kernel (double 7x7 fused solve 0.43 / 0.59 -> 0.78 / 0.67 of the HBM peak, float 10x10 0.39 / 0.52
// -> 0.56 / 0.57, float 11x11 0.35 / 0.43 -> 0.54 / 0.49, float 10x10 with three right-hand sides 0.24 / 0.33 -> 0.56 / 0.55); the
// rider-free kernels (decomposition, deferred-rider inverse, rectangular mode) take 121 = float n <= 11 (decomp_plu 11x11 from
// AoS records 0.54 -> 0.74; 130 words spilled: double 8x8).
#ifndef GMB_SUB_REG_BUDGET
#define GMB_SUB_REG_BUDGET 132
#endif
#ifndef GMB_SUB_REG_BUDGET_XC0
#define GMB_SUB_REG_BUDGET_XC0 121
#endif
constexpr int kRegBudgetWords = GMB_SUB_REG_BUDGET;

constexpr int MODE_SOLVE = 0;   // A, Bm -> X, ok (, LU)
constexpr int MODE_DECOMP = 1;  // A in place, perm, parity, degenerate
constexpr int MODE_INV = 2;     // A -> Ainv, ok
constexpr int MODE_INVD = 3;    // A -> Ainv, ok: one lane per system, riders deferred (see k_sub_plu)
constexpr int MODE_RECT = 4;    // nb x N bases (nb < N rows, runtime) -> null basis / orthogonal vector, ok (Orthogonal.h:73-178)

constexpr int sub_lc(int n, int xc, int g) { return (n + xc + g - 1) / g; }
// does ANY lane of the group hold a rider column / a column beyond `i` at local index j?  (static)
constexpr bool sub_any_rider(int n, int xc, int g, int j) {
    const int lc = sub_lc(n, xc, g);
    for (int l = 0; l < g; ++l)
        if (l * lc + j >= n && l * lc + j < n + xc) return true;
    return false;
}
// which directions of an AoS batch go through the warp's staging buffer (see k_sub_plu)
template <typename T>
constexpr bool sub_stage_in(int n, int mode, bool soa) {
    return !soa && sizeof(T) == 4 && mode != MODE_INV && (n == 8 || n == 9);
}
constexpr bool sub_stage_out(int mode, bool soa) { return !soa && mode != MODE_SOLVE; }
// "Wide" one-lane kernels (round 2, profiles/r2t_wide_lane.md): a local block [A | riders] of up to GMB_SUB_WIDE_WORDS register
// words that exceeds the budgets above still runs on ONE lane per system, under the 255-register ceiling (two 128-thread
// CTAs per SM) instead of two lanes with a shuffle per exchanged / broadcast element: double n = 8, 9, 10, float n = 12, 13, 14 and
// the multi-right-hand-side solves just above the budget (double 7x7 .. 9x9 and float 11x11 .. 13x13 with three of them).
#ifndef GMB_SUB_WIDE_WORDS
#define GMB_SUB_WIDE_WORDS 220            // 0 switches the wide kernels off; 220 = double 10x10 with one rider (spills ~300 bytes, still 1.6x the row kernel)
#endif
constexpr bool sub_wide(int n, int xc, int words) {
    const int block = words * n * (n + xc);
    return block > (xc == 0 ? GMB_SUB_REG_BUDGET_XC0 : GMB_SUB_REG_BUDGET) && block <= GMB_SUB_WIDE_WORDS;
}
constexpr int sub_group(int n, int xc, int words) {
    if (sub_wide(n, xc, words)) return 1;
    for (int g = 1; g <= 16; g *= 2)
        if (n * sub_lc(n, xc, g) * words <= (xc == 0 ? GMB_SUB_REG_BUDGET_XC0 : kRegBudgetWords)) return g;
    return 16;
}
// Per-mode / per-layout hook.  Until the SoA plane segments were staged through shared memory (sub_soa_stage) the row-pivot
// decomposition of an SoA batch kept two lanes per system for float n = 12, 14 and double n = 10 (0.64 / 0.54 / 0.59 of the HBM
// peak against 0.57 / 0.51 / 0.46 on one wide lane with direct plane loads, profiles/r2t_wide_lane.md); with the staging the
// one-lane form wins there too (run r2as) and the exception is gone.
constexpr int sub_group_for(int n, int xc, int words, int mode, bool soa, bool colpiv) {
    (void)mode; (void)soa; (void)colpiv;
    return sub_group(n, xc, words);
}
constexpr int sub_min_blocks(int n, int xc, int words, int mode, bool soa, bool colpiv) {
    return (sub_wide(n, xc, words) && sub_group_for(n, xc, words, mode, soa, colpiv) == 1) ? 2 : GMB_SUB_MIN_BLOCKS;
}

template <int G>
__device__ __forceinline__ unsigned shfl_u(unsigned v, int src_in_group) {
    if constexpr (G == 1) return v;
    else return __shfl_sync(0xffffffffu, v, src_in_group, G);
}
template <int G>
__device__ __forceinline__ float shfl_t(float v, int src_in_group) {
    if constexpr (G == 1) return v;
    else return __shfl_sync(0xffffffffu, v, src_in_group, G);
}
template <int G>
__device__ __forceinline__ double shfl_t(double v, int src_in_group) {
    if constexpr (G == 1) return v;
    else return __shfl_sync(0xffffffffu, v, src_in_group, G);
}

// element e of system s lives at p[base(s) + e * STRIDE]; STRIDE is a compile-time constant so every
// element offset folds into the load/store instruction's immediate
template <typename T, bool SOA>
struct BatchAddr {
    static constexpr int STRIDE = SOA ? soa_w<T>() : 1;
    __device__ static __forceinline__ int64_t base(int64_t s, int E) {
        constexpr int W = soa_w<T>();
        return SOA ? (s / W) * (int64_t)E * W + (s % W) : s * (int64_t)E;
    }
};

// One lane per system (G == 1) on an AoS batch: the warp's 32 records are one contiguous range, moved with coalesced
// 128-bit accesses through a warp-private, swizzled shared-memory transpose (aos_load_staged / aos_store_staged,
// gmb_io.cuh) instead of 32 lanes walking 32 records at a stride of one record.  `sm` = the warp's staging buffer.
template <typename T, int E>
__device__ __forceinline__ void sub_lane_load(const T* __restrict__ base, int64_t s, int lane, bool full, bool valid, T (&q)[E],
                                              unsigned char* sm) {
    if (StageCfg<E * (int)sizeof(T)>::use && full) {
        aos_load_staged<T, E>(base, s - lane, lane, q, sm);
    } else if (valid) {
        aos_load<T, E>(base, s, q);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) q[e] = (T)0;
    }
}
template <typename T, int E>
__device__ __forceinline__ void sub_lane_store(T* __restrict__ base, int64_t s, int lane, bool full, bool valid, const T (&q)[E],
                                               unsigned char* sm) {
    if (StageCfg<E * (int)sizeof(T)>::use && full) {
        aos_store_staged<T, E>(base, s - lane, lane, q, sm);
    } else if (valid) {
        aos_store<T, E>(base, s, q);
    }
}
// elements per system of a G == 1 AoS kernel's staging buffer: the matrix and the rider block side by side (both are in
// flight at once), never less than n*n + 1 (the deferred-rider inverse uses the buffer as an odd-stride scratch)
// SoA batches of the one-lane kernels that are staged through shared memory (see the load section of k_sub_plu)
constexpr bool sub_soa_stage(int n, int words, int mode, int g, bool soa) {
    if (!soa || g != 1) return false;
    if (mode == 1 /* MODE_DECOMP */) return n * n * words * 4 >= 400 && !(words == 2 && n == 8);
    // the deferred-rider inverse: the two sizes where the staging measured a gain (run r2ap: double 7x7 0.65 -> 0.76, float 10x10 0.48 -> 0.53)
    if (mode == 3 /* MODE_INVD */) return (words == 2 && n == 7) || (words == 1 && n == 10);
    return false;
}
constexpr int sub_lane_emax(int n, int xc, int mode) { return n * n + (mode == 0 /* MODE_SOLVE */ ? n * xc : 0); }

// MODE_RECT (XC = 0, G = 1): orthogonal (Orthogonal.h:73-98) and nullspace (:136-178).  Both run the RECTANGULAR row-pivot
// decomposition of the nb x N matrix whose rows are the given vectors (nb = `skip`, a run-time value, nb < N) and then a
// small back-substitution.  The matrix is held as N x N with rows >= nb zero: a zero row never wins the strict pivot
// comparison (:205-213) and its multipliers are 0 / pivot, so rows < nb see exactly the rectangular algorithm provided
// the elimination stops after min(nb, N) - 1 = nb - 1 steps, as the reference's does (:203).
// MODE_INVD (XC = 0, G = 1): invNxN (MatrixInv.h:219-231) without carrying the identity through the elimination.  The
// reference decomposes, sets out(i, p[i]) = 1 and runs backsolve_lu on the n columns of `out`.  Column p[i] of that
// right-hand side is e_i, so its forward sweep (LUDecomp.h:345-353) is y(r) = 0 for r < i, y(i) = 1 and the reference's
// chain from c = i on for r > i -- the terms with c < i multiply an exact zero: with finite multipliers they leave the
// accumulator at +0, bit for bit.  The kernel therefore factors [A] alone (half the columns to exchange, the
// decomposition runs at the HBM roofline for these sizes), then produces the inverse ONE COLUMN AT A TIME in i-space
// -- forward chain on the statically known non-zero part, dense backward sweep (:356-365), one reciprocal per
// diagonal entry shared by all columns -- and stores column i at its dynamic position p[i] through a per-thread,
// bank-conflict-free shared-memory scratch (a register file cannot be indexed at run time; memory can).  A system
// whose multipliers are not all finite (0 * inf = NaN in the skipped terms) takes the dense forward sweep instead,
// in the same kernel: same bits as the reference for every input.
// COLPIV = true: decomp_lup (LUDecomp.h:102-165, column pivoting) on the TRANSPOSED view of the matrix -- the launcher
// flips A_RM, so "row p" below is column p of M: the search then runs along row i of M (:124-131), the exchange swaps
// columns of M (:140-145), and only the elimination differs from decomp_plu: the multipliers b_r = M(r,i) / M(i,i)
// (:152) sit in ROW i of the view (one per column, computed by the lane that owns the column, pivot broadcast with one
// shuffle) and the unscaled M(i,c) (:157) is COLUMN i of the view, broadcast from its owner with n-1-i shuffles --
// the same number of shuffles as decomp_plu, with the roles of the two factors exchanged.
// NO_L = true (one lane per system, fused solve, no LU output asked for): nobody reads the multipliers afterwards, so they
// are neither stored nor exchanged -- only columns >= i of a row travel (a quarter of the exchange's selects for n = 7).
template <typename T, int N, int XC, int MODE, bool SOA, bool A_RM, bool X_RM, bool COLPIV = false, bool NO_L = false>
__global__ void __launch_bounds__(kSubThreads, sub_min_blocks(N, XC, (int)sizeof(T) / 4, MODE, SOA, COLPIV))
k_sub_plu(const T* __restrict__ A, const T* Bm, T* X, T* LU, uint8_t* __restrict__ ok_out, uint8_t* __restrict__ perm_out,
          uint8_t* __restrict__ parity_out, uint8_t* __restrict__ degen_out, int skip, int64_t B, int pf_dist) {
    constexpr int ST = BatchAddr<T, SOA>::STRIDE;
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int G = sub_group_for(N, XC, WORDS, MODE, SOA, COLPIV);
    constexpr int NC = N + XC;
    constexpr int LC = sub_lc(N, XC, G);
    const int lane_g = (G == 1) ? 0 : (int)(threadIdx.x % G);
    const int c0 = lane_g * LC;                               // first augmented column of this lane
    int64_t s = ((int64_t)blockIdx.x * kSubThreads + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;                                    // keep every lane alive for the shuffles

    // One lane per system: the 128 systems of a CTA are ONE contiguous range in either layout (AoS: 128 records; SoA: whole
    // tiles, 128 % W == 0).  The CTA that will run `pf_dist` CTAs later (about one wave: the SM count times the resident CTAs)
    // gets its range pulled into L2 now, so that its loads -- issued by the few warps the register footprint leaves -- meet an
    // L2 latency instead of a DRAM one.  No extra traffic: every byte is still fetched from DRAM once.
    if constexpr (G == 1 && MODE != MODE_RECT) {
        const int64_t s_pf = ((int64_t)blockIdx.x + pf_dist) * kSubThreads;
        if (pf_dist > 0 && s_pf + kSubThreads <= B) {
            constexpr unsigned CHUNK = 8192;
            constexpr unsigned BYTES_A = kSubThreads * N * N * (unsigned)sizeof(T);
            const char* pa = reinterpret_cast<const char*>(A + s_pf * (int64_t)(N * N));
            for (unsigned o = threadIdx.x * CHUNK; o < BYTES_A; o += kSubThreads * CHUNK)
                l2_prefetch_bulk(pa + o, (BYTES_A - o < CHUNK) ? BYTES_A - o : CHUNK);
            if constexpr (MODE == MODE_SOLVE) {
                constexpr unsigned BYTES_B = kSubThreads * N * XC * (unsigned)sizeof(T);
                const char* pb = reinterpret_cast<const char*>(Bm + s_pf * (int64_t)(N * XC));
                for (unsigned o = threadIdx.x * CHUNK; o < BYTES_B; o += kSubThreads * CHUNK)
                    l2_prefetch_bulk(pb + o, (BYTES_B - o < CHUNK) ? BYTES_B - o : CHUNK);
            }
        }
    }
    const int64_t baseA = BatchAddr<T, SOA>::base(s, N * N);
    const int64_t baseX = BatchAddr<T, SOA>::base(s, N * (XC > 0 ? XC : 1));
    const T* Ap = A + baseA + (int64_t)c0 * (A_RM ? ST : N * ST);          // this lane's first column of A

    // AoS batches: the warp's NSYS systems are staged through shared memory (gmb_io.cuh, warp_stage_in)
    extern __shared__ __align__(16) unsigned char gmb_dyn_smem[];
    constexpr int NSYS = 32 / G;
    constexpr int SA = stage_stride<N * N>();
    T* smw = reinterpret_cast<T*>(gmb_dyn_smem) + (threadIdx.x >> 5) * (NSYS * SA);
    const int64_t s_warp0 = ((int64_t)blockIdx.x * kSubThreads + (threadIdx.x & ~31u)) / G;
    // Measured (profiles/r1e_*): per-lane strided LOADS go through L1 and every fetched line is consumed,
    // so staging them only pays where the record stride is a power of two (f32 N=8) or nearly so;
    // per-lane strided STORES write partial sectors, so matrix outputs are always staged.
    constexpr bool LANE_IO = (G == 1) && !SOA;                             // one lane per system: staged record I/O
    constexpr bool STAGE_IN = !LANE_IO && sub_stage_in<T>(N, MODE, SOA), STAGE_OUT = !LANE_IO && sub_stage_out(MODE, SOA);
    const bool stage = (STAGE_IN || STAGE_OUT) && (s_warp0 + NSYS <= B);   // warp-uniform
    const int64_t s_raw = (int64_t)blockIdx.x * kSubThreads + threadIdx.x; // LANE_IO: this thread's system before clamping
    const bool lane_full = LANE_IO && (s_warp0 + 32 <= B);                 // warp-uniform
    unsigned char* lane_sm = gmb_dyn_smem + (size_t)(threadIdx.x >> 5) * (32 * (sub_lane_emax(N, XC, MODE) + 1) * sizeof(T));
    const T* sms = smw + ((threadIdx.x & 31) / G) * SA + (A_RM ? c0 : N * c0);   // this lane's system, first local column
    if (STAGE_IN && stage) warp_stage_in<T, N * N, NSYS>(A + s_warp0 * (int64_t)(N * N), smw, threadIdx.x & 31);

    // ---- load the local block -------------------------------------------------------------
    T a[N][LC];
    if constexpr (MODE == MODE_RECT) {
        static_assert(G == 1 && XC == 0, "rectangular mode: one lane per system");
        const int nb = skip;
        const int64_t baseV = BatchAddr<T, SOA>::base(s, N * nb);       // nb vectors of N: row-major nb x N
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = (r < nb) ? A[baseV + (int64_t)(N * r + c) * ST] : (T)0;
    } else if constexpr (LANE_IO) {
        T qa[N * N];
        T qb[MODE == MODE_SOLVE ? N * XC : 1];
        constexpr bool STAGE_B = MODE == MODE_SOLVE && StageCfg<N * XC * (int)sizeof(T)>::use;
        const int lane = threadIdx.x & 31;
        if (lane_full) {
            // matrix and right-hand sides: every global -> shared copy is issued (cp.async) before anybody waits
            aos_stage_issue<T, N * N>(A, s_warp0, lane, lane_sm);
            if constexpr (STAGE_B) aos_stage_issue<T, N * XC>(Bm, s_warp0, lane, lane_sm + 32 * N * N * sizeof(T));
            if constexpr (MODE == MODE_SOLVE && !STAGE_B) aos_load<T, N * XC>(Bm, s_raw, qb);     // records under 32 bytes: direct
            cp_async_wait_all_();
            __syncwarp();
            aos_stage_read<T, N * N>(lane, qa, lane_sm);
            if constexpr (STAGE_B) aos_stage_read<T, N * XC>(lane, qb, lane_sm + 32 * N * N * sizeof(T));
            __syncwarp();                                      // the buffer is reused by the staged stores
        } else {
            sub_lane_load<T, N * N>(A, s_raw, lane, false, valid, qa, lane_sm);
            if constexpr (MODE == MODE_SOLVE) sub_lane_load<T, N * XC>(Bm, s_raw, lane, false, valid, qb, lane_sm);
        }
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = qa[A_RM ? N * r + c : N * c + r];
        if constexpr (MODE == MODE_SOLVE) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < XC; ++j) a[r][N + j] = qb[X_RM ? XC * r + j : N * j + r];
        }
        if constexpr (MODE == MODE_INV) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < XC; ++j) a[r][N + j] = (r == j) ? (T)1 : (T)0;
        }
    } else {
        // SoA, one lane per system, decompositions of matrices of 400 bytes and more (sub_soa_stage): the warp's 32-system segment
        // of every element plane is copied global -> shared with cp.async (soa_stage_issue, gmb_io.cuh) and read back as
        // [plane][lane] -- no registers held and no load / store queue slots taken while the data travel (ncu: LG throttle was
        // the top stall of the direct plane loads).  Measured for every mode (profiles/r2ap_soa_staging.md): it wins for the
        // decompositions from float n = 10 / double n = 9 on (float 11x11 0.63 -> 0.81 of the HBM peak, double 9x9 0.63 -> 0.79) and
        // loses 0.03 .. 0.08 for the ALU-bound solves and inverses of the small sizes (extra shared-memory reads), which keep
        // their direct plane loads.
        bool soa_staged = false;
        if constexpr (sub_soa_stage(N, WORDS, MODE, G, SOA)) {
            if (s_warp0 + 32 <= B) {                               // warp-uniform; the 32 systems sit inside one tile (W % 32 == 0)
                const int lane = threadIdx.x & 31;
                unsigned char* smb = lane_sm + 32 * N * N * sizeof(T);
                soa_stage_issue<T, N * N>(A, s_warp0, lane, lane_sm);
                if constexpr (MODE == MODE_SOLVE) soa_stage_issue<T, N * XC>(Bm, s_warp0, lane, smb);
                cp_async_wait_all_();
                __syncwarp();
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) a[r][c] = soa_stage_elem<T>(lane, A_RM ? N * r + c : N * c + r, lane_sm);
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int j = 0; j < XC; ++j) {
                        if (MODE == MODE_SOLVE) a[r][N + j] = soa_stage_elem<T>(lane, X_RM ? XC * r + j : N * j + r, smb);
                        if (MODE == MODE_INV) a[r][N + j] = (r == j) ? (T)1 : (T)0;
                    }
                __syncwarp();                                      // the buffer is reused (deferred-rider scratch)
                soa_staged = true;
            }
        }
        if (!soa_staged) {
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int j = 0; j < LC; ++j) {
            const int c = c0 + j;
            T v = (T)0;
            if (c < N) {
                v = (STAGE_IN && stage) ? sms[A_RM ? N * r + j : N * j + r] : Ap[(A_RM ? N * r + j : N * j + r) * ST];
            } else if (c < NC) {
                if (MODE == MODE_SOLVE) v = Bm[baseX + (int64_t)(X_RM ? XC * r + (c - N) : N * (c - N) + r) * ST];
                if (MODE == MODE_INV) v = (r == c - N) ? (T)1 : (T)0;     // identity; rows exchanged below
            }
            a[r][j] = v;
        }
        }
    }

    uint8_t perm[N];
    if (MODE == MODE_DECOMP || MODE == MODE_INVD) {
#pragma unroll
        for (int r = 0; r < N; ++r) perm[r] = (uint8_t)r;
    }
    int degenerate = 0;
    bool parity = false;

    // ---- elimination ----------------------------------------------------------------------
    static_assert(!NO_L || (MODE == MODE_SOLVE && G == 1), "NO_L: fused solve, one lane per system");
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        if (MODE == MODE_RECT && i >= skip - 1) break;       // rectangular: min(rows, cols) - 1 steps (warp-uniform)
        const int owner = i / LC;          // static
        const int k = i % LC;              // static local column of the pivot column in the owner lane
        // pivot search on local column k (meaningful in the owner lane only)
        T biggest = abs_(a[i][k]);
        bool gt[N];
#pragma unroll
        for (int p = i + 1; p < N; ++p) {
            const T v = abs_(a[p][k]);
            gt[p] = v > biggest;
            biggest = gt[p] ? v : biggest;
        }
        // one-hot "row p is the pivot row" flags: the LAST strict improvement of the scan wins (= first strict maximum)
        bool swp[N];
        bool later = false;
#pragma unroll
        for (int p = N - 1; p > i; --p) {
            swp[p] = gt[p] && !later;
            later = later || gt[p];
        }
        bool dead = biggest == (T)0;
        if constexpr (G > 1) {
            // the owner's flags travel as ONE word (a mask, not an index, so the exchange below stays in registers)
            unsigned info = 0;
#pragma unroll
            for (int p = N - 1; p > i; --p) info |= swp[p] ? (1u << p) : 0u;
            info |= dead ? 0x80000000u : 0u;
            info = shfl_u<G>(info, owner);
            dead = (info & 0x80000000u) != 0;
            later = (info & 0x7fffffffu) != 0;
#pragma unroll
            for (int p = i + 1; p < N; ++p) swp[p] = ((info >> p) & 1u) != 0;
        }
        degenerate += dead ? 1 : 0;
        parity = parity != later;
        // exchange rows i <-> pvt in the local columns (NO_L: columns < i hold dead multipliers and stay where they are)
        const int JDEAD = NO_L ? i : 0;                        // first local column that must travel (static after unrolling)
#pragma unroll
        for (int p = i + 1; p < N; ++p) {
            const bool sw = swp[p];
#pragma unroll
            for (int j = JDEAD; j < LC; ++j) {
                const T t = a[i][j];
                a[i][j] = sw ? a[p][j] : t;
                a[p][j] = sw ? t : a[p][j];
            }
            if (MODE == MODE_DECOMP || MODE == MODE_INVD) {
                const uint8_t t = perm[i];
                perm[i] = sw ? perm[p] : t;
                perm[p] = sw ? t : perm[p];
            }
        }
        // multipliers from the owner lane, then the rank-1 update of the local columns
        const T piv = a[i][k];
        const bool mine = (G == 1) || (lane_g == owner);
        if constexpr (COLPIV) {
            const T pv = shfl_t<G>(piv, owner);
            const RcpOf<T> rc(pv);
            T bq[LC];
            bool all_ok = true;
#pragma unroll
            for (int j = 0; j < LC; ++j) {
                bq[j] = (T)0;
                if ((G - 1) * LC + j > i) {                       // static: some lane's column at j is right of i
                    bool ok;
                    bq[j] = rc.fast(a[i][j], ok);
                    all_ok = all_ok && (ok || !(c0 + j > i && c0 + j < N));
                }
            }
            if (!all_ok) {
#pragma unroll
                for (int j = 0; j < LC; ++j)
                    if ((G - 1) * LC + j > i) bq[j] = div_rn(a[i][j], pv);
            }
            bool on[LC];
#pragma unroll
            for (int j = 0; j < LC; ++j) {
                on[j] = !dead && (c0 + j > i) && (c0 + j < N);
                if ((G - 1) * LC + j > i) a[i][j] = on[j] ? bq[j] : a[i][j];       // M(r,i) = b (:163)
                on[j] = on[j] && (bq[j] != (T)0);                                    // `if (b != 0)` (:153)
            }
#pragma unroll
            for (int r = i + 1; r < N; ++r) {
                const T t = shfl_t<G>(a[r][k], owner);           // M(i, r): row i of M is not touched by this step
#pragma unroll
                for (int j = 0; j < LC; ++j) {
                    if ((G - 1) * LC + j > i) {
                        if (on[j]) a[r][j] = fma_rn(-bq[j], t, a[r][j]);
                    }
                }
            }
        } else {
        // one reciprocal per pivot, three operations per quotient (RcpOf, gmb_math.cuh); the owner lane redoes its
        // quotients with div_rn when an operand leaves the fast domain -- same bits either way
        T bq[N];
        {
            const RcpOf<T> rc(piv);
            bool all_ok = true;
#pragma unroll
            for (int r = i + 1; r < N; ++r) {
                bool ok;
                bq[r] = rc.fast(a[r][k], ok);
                all_ok = all_ok && ok;
            }
            if (mine && !all_ok) {
#pragma unroll
                for (int r = i + 1; r < N; ++r) bq[r] = div_rn(a[r][k], piv);
            }
        }
#pragma unroll
        for (int r = i + 1; r < N; ++r) {
            T b = shfl_t<G>(bq[r], owner);
            const bool upd = !dead && (b != (T)0);
            if (mine && !NO_L) a[r][k] = dead ? a[r][k] : b;    // L(r,i)
#pragma unroll
            for (int j = 0; j < LC; ++j) {
                if ((G - 1) * LC + j > i) {                       // static: some lane's column at j is right of i
                    const int c = c0 + j;
                    const bool is_rider = XC > 0 && c >= N;
                    const bool go = (c > i) && (is_rider || upd);
                    if (go) a[r][j] = fma_rn(-b, a[i][j], a[r][j]);     // predicated FMA, no select
                }
            }
        }
        }
    }

    const bool good = degenerate == 0;

    if constexpr (MODE == MODE_RECT) {
        const int nb = skip, nn = N - nb;
        const int64_t baseO = BatchAddr<T, SOA>::base(s, N * nn);
        if (valid && ok_out) ok_out[s] = (nn == 1 || good) ? 1 : 0;     // N - n == 1 returns true whatever happened (:138-142)
        if (nn == 1) {
            // orthogonal (:73-98): o = 0 when the decomposition reports a degenerate column; else o(N-1) = 1 and, for
            // r = N-2 .. 0, o(r) = (0 - sum_{c = N-1 .. r+1} o(c) m(r,c)) / m(r,r)
            T o[N];
#pragma unroll
            for (int e = 0; e < N; ++e) o[e] = (T)0;
            if (good) {
                o[N - 1] = (T)1;
#pragma unroll
                for (int r = N - 2; r >= 0; --r) {
                    T acc = (T)0;
#pragma unroll
                    for (int c = N - 1; c > r; --c) acc = fma_rn(-o[c], a[r][c], acc);
                    o[r] = div_rn(acc, a[r][r]);
                }
            }
            if (valid) {
#pragma unroll
                for (int e = 0; e < N; ++e) X[baseO + (int64_t)e * ST] = o[e];
            }
            return;
        }
        // nullspace (:144-176): null vector b has 1 at row nb + b, x(r) = (-m(r, nb + b) - sum_{c = r+1 .. nb-1} m(r,c) x(c)) / m(r,r)
        // for r = nb-1 .. 0, zeros elsewhere; all-zero vectors when the decomposition reports a degenerate column
#pragma unroll
        for (int b = 0; b < N - 1; ++b) {
            if (b >= nn) break;
            T x[N];
#pragma unroll
            for (int r = N - 1; r >= 0; --r) {
                x[r] = (T)0;
                if (r < nb && r <= N - 3) {
                    T t = (T)0;
#pragma unroll
                    for (int c = 1; c < N; ++c) t = (c == nb + b) ? a[r][c] : t;       // m(r, nb + b): run-time column
                    T xi = -t;
#pragma unroll
                    for (int c = r + 1; c <= N - 3; ++c)
                        if (c < nb) xi = fma_rn(-a[r][c], x[c], xi);
                    x[r] = div_rn(xi, a[r][r]);
                }
            }
            if (valid) {
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    const T v = !good ? (T)0 : (e < nb ? x[e] : (e == nb + b ? (T)1 : (T)0));
                    X[baseO + (int64_t)(N * b + e) * ST] = v;
                }
            }
        }
        return;
    }

    if constexpr (MODE == MODE_INVD) {
        static_assert(G == 1 && XC == 0, "deferred-rider inverse: one lane per system");
        const int lane = threadIdx.x & 31;
        // per-thread scratch for the dynamically placed columns: SoA [element][lane] (bank = lane: conflict-free for any
        // index), AoS [record][element] at an odd stride (the layout warp_stage_out flushes with coalesced stores)
        T* scr = reinterpret_cast<T*>(lane_sm);
        constexpr int SA2 = stage_stride<N * N>();
        auto slot = [&](int e) -> T& { return SOA ? scr[e * 32 + lane] : scr[lane * SA2 + e]; };
        // are all multipliers finite?  (0 * l accumulates to 0 unless some l is Inf / NaN; FMA pipe, no compares)
        T fin = (T)0;
#pragma unroll
        for (int r = 1; r < N; ++r)
#pragma unroll
            for (int c = 0; c < r; ++c) fin = fma_rn(a[r][c], (T)0, fin);
        const bool sparse_ok = fin == (T)0;
        // one reciprocal per diagonal entry of U, shared by the n columns (RcpOf, gmb_math.cuh)
        T ry[N];
        bool rs[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const RcpOf<T> rc(a[r][r]);
            ry[r] = rc.y;
            rs[r] = rc.safe;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            T y[N];
            if (sparse_ok) {
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    if (r < i) y[r] = (T)0;
                    else if (r == i) y[r] = (T)1;
                    else {
                        T v = (T)0;
#pragma unroll
                        for (int c = i; c < r; ++c) v = fma_rn(-y[c], a[r][c], v);
                        y[r] = v;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    T v = (r == i) ? (T)1 : (T)0;
#pragma unroll
                    for (int c = 0; c < r; ++c) v = fma_rn(-y[c], a[r][c], v);
                    y[r] = v;
                }
            }
            // backward sweep (:356-365): x(r) = (y(r) - sum_{c = n-1 .. r+1} u(r,c) x(c)) / u(r,r)
            T x[N];
            bool all_ok = true;
#pragma unroll
            for (int r = N - 1; r >= 0; --r) {
                T v = y[r];
#pragma unroll
                for (int c = N - 1; c > r; --c) v = fma_rn(-x[c], a[r][c], v);
                y[r] = v;                                       // the dividend, kept for the exact redo
                bool okq;
                x[r] = RcpOf<T>(a[r][r], ry[r], rs[r]).fast(v, okq);
                all_ok = all_ok && okq;
            }
            if (!all_ok) {                                      // some operand outside the fast division window: IEEE divisions
                // redo the whole column with div_rn (the quotients feed the rows above them)
                T yy[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    T v = (r == i) ? (T)1 : (T)0;
#pragma unroll
                    for (int c = 0; c < r; ++c) v = fma_rn(-yy[c], a[r][c], v);
                    yy[r] = v;
                }
#pragma unroll
                for (int r = N - 1; r >= 0; --r) {
                    T v = yy[r];
#pragma unroll
                    for (int c = N - 1; c > r; --c) v = fma_rn(-x[c], a[r][c], v);
                    x[r] = div_rn(v, a[r][r]);
                }
            }
            const int pc = perm[i];                             // out(r, p[i]) = x(r): row-major flat index n*r + p[i] (:192)
#pragma unroll
            for (int r = 0; r < N; ++r) slot(N * r + pc) = x[r];
        }
        // ---- egress
        if constexpr (SOA) {
            if (valid && good) {
                T* dst = X + baseA;
#pragma unroll
                for (int e = 0; e < N * N; ++e) dst[e * ST] = slot(e);
            }
        } else {
            const bool staged = lane_full && __all_sync(0xffffffffu, good);
            if (staged) {
                warp_stage_out<T, N * N, 32>(X + s_warp0 * (int64_t)(N * N), scr, lane);
            } else if (valid && good) {
                T* dst = X + baseA;
#pragma unroll
                for (int e = 0; e < N * N; ++e) dst[e] = slot(e);
            }
        }
        if (valid && ok_out) ok_out[s] = good ? 1 : 0;
        return;
    }

    // ---- outputs that need no backward sweep ------------------------------------------------
    if (MODE == MODE_DECOMP || (MODE == MODE_SOLVE && LU != nullptr)) {
        T* dstb = (MODE == MODE_DECOMP) ? const_cast<T*>(A) : LU;
        T* dst = dstb + baseA + (int64_t)c0 * (A_RM ? ST : N * ST);
        if constexpr (LANE_IO) {
            T ql[N * N];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) ql[A_RM ? N * r + c : N * c + r] = a[r][c];
            sub_lane_store<T, N * N>(dstb, s_raw, threadIdx.x & 31, lane_full, valid, ql, lane_sm);
        } else if (STAGE_OUT && stage) {
            __syncwarp();                                              // every lane is done reading the staged input
            T* smd = smw + ((threadIdx.x & 31) / G) * SA + (A_RM ? c0 : N * c0);
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < LC; ++j)
                    if (c0 + j < N) smd[A_RM ? N * r + j : N * j + r] = a[r][j];
            warp_stage_out<T, N * N, NSYS>(dstb + s_warp0 * (int64_t)(N * N), smw, threadIdx.x & 31);
        } else if (valid) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < LC; ++j) {
                    const int c = c0 + j;
                    if (c < N) dst[(A_RM ? N * r + j : N * j + r) * ST] = a[r][j];
                }
        }
    }
    if (MODE == MODE_DECOMP) {
        if (valid && lane_g == 0) {
            if (perm_out) {
#pragma unroll
                for (int r = 0; r < N; ++r) perm_out[s * N + r] = perm[r];
            }
            if (parity_out) parity_out[s] = parity ? 1 : 0;
            if (degen_out) degen_out[s] = (uint8_t)degenerate;
        }
        return;
    }

    // ---- backward sweep on the riders (column-oriented) -------------------------------------
    if (XC > 0) {
#pragma unroll
        for (int c = N - 1; c >= 0; --c) {
            const int owner = c / LC;
            const int k = c % LC;
            const bool solve_row = c >= skip;
            // u(r, c) for r <= c lives in the owner lane's local column k
            const T ucc = shfl_t<G>(a[c][k], owner);
#pragma unroll
            for (int j = 0; j < LC; ++j) {
                if (sub_any_rider(N, XC, G, j)) {                 // static
                    const bool is_rider = (c0 + j >= N);
                    const T q = div_rn(a[c][j], ucc);
                    a[c][j] = (is_rider && solve_row) ? q : a[c][j];
                }
            }
#pragma unroll
            for (int r = c - 1; r >= 0; --r) {
                const T urc = shfl_t<G>(a[r][k], owner);
                const bool row_on = r >= skip;
#pragma unroll
                for (int j = 0; j < LC; ++j) {
                    if (sub_any_rider(N, XC, G, j)) {             // static
                        const bool is_rider = (c0 + j >= N);
                        const T v = fma_rn(-a[c][j], urc, a[r][j]);
                        a[r][j] = (is_rider && row_on) ? v : a[r][j];
                    }
                }
            }
        }
    }

    // ---- store riders -----------------------------------------------------------------------
    // inverse, AoS: when every system of the warp is regular the result goes back through the staging
    // buffer (coalesced); a warp containing a singular system stores per lane so that the singular
    // system's destination stays untouched
    bool inv_staged = false;
    if (MODE == MODE_INV && STAGE_OUT && stage) {
        inv_staged = __all_sync(0xffffffffu, good);
        if (inv_staged) {
            __syncwarp();
            T* smd = smw + ((threadIdx.x & 31) / G) * SA;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < LC; ++j) {
                    const int c = c0 + j;
                    if (c >= N && c < NC) smd[N * r + (c - N)] = a[r][j];
                }
            warp_stage_out<T, N * N, NSYS>(X + s_warp0 * (int64_t)(N * N), smw, threadIdx.x & 31);
        }
    }
    if constexpr (LANE_IO && (MODE == MODE_SOLVE || MODE == MODE_INV)) {
        // a warp whose systems are all regular stores through the staging buffer; one with a failed system stores per
        // thread below, so that the failed system's destination keeps what it must keep (x = b / `out` untouched)
        if (__all_sync(0xffffffffu, good || !valid)) {
            T qx[N * XC];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < XC; ++j) qx[(MODE == MODE_INV || X_RM) ? XC * r + j : N * j + r] = a[r][N + j];
            sub_lane_store<T, N * XC>(X, s_raw, threadIdx.x & 31, lane_full, valid, qx, lane_sm);
            if (valid && ok_out) ok_out[s] = 1;
            return;
        }
    }
    if (valid) {
        if (MODE == MODE_SOLVE) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < LC; ++j) {
                    const int c = c0 + j;
                    if (c >= N && c < NC) {
                        const int64_t off = baseX + (int64_t)(X_RM ? XC * r + (c - N) : N * (c - N) + r) * ST;
                        X[off] = good ? a[r][j] : Bm[off];              // x untouched on failure (:391-393)
                    }
                }
        }
        if (MODE == MODE_INV && good && !(stage && inv_staged)) {        // out untouched on failure
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < LC; ++j) {
                    const int c = c0 + j;
                    if (c >= N && c < NC) X[baseA + (int64_t)(N * r + (c - N)) * ST] = a[r][j];   // invNxN writes `out` row-major (:192)
                }
        }
        if (lane_g == 0 && ok_out) ok_out[s] = good ? 1 : 0;
    }
}

// ---------------------------------------------------------------- launchers
template <typename T, int N, int XC, int MODE, bool COLPIV = false>
static int launch_sub(const T* A, const T* Bm, T* X, T* LU, uint8_t* ok, uint8_t* perm, uint8_t* parity, uint8_t* degen,
                      int skip, int64_t B, int layout, int a_rm, int x_rm, cudaStream_t st) {
    if (B == 0) return GMB_OK;
    dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
        constexpr int G = sub_group_for(N, XC, (int)sizeof(T) / 4, MODE, SOA.value, COLPIV);
        const unsigned grid = (unsigned)ceil_div(B * G, kSubThreads);
        dispatch_bool(a_rm != 0, [&](auto ARM) {
            // riders: linear_solve uses the matrix' own layout flag; invNxN always writes row-major
            constexpr bool XRM = (MODE == MODE_INV || MODE == MODE_INVD) ? true : ARM.value;
            (void)x_rm;
            constexpr bool CAN_NO_L = (G == 1) && (MODE == MODE_SOLVE) && !COLPIV;
            auto kern = (CAN_NO_L && LU == nullptr) ? k_sub_plu<T, N, XC, MODE, SOA.value, ARM.value, XRM, COLPIV, CAN_NO_L>
                                                    : k_sub_plu<T, N, XC, MODE, SOA.value, ARM.value, XRM, COLPIV, false>;
            // AoS: one staging buffer of 32/G systems per warp (dynamic shared memory)
            const size_t smem = (G == 1 && (!SOA.value || MODE == MODE_INVD || sub_soa_stage(N, (int)sizeof(T) / 4, MODE, G, true)))
                                    ? (size_t)(kSubThreads / 32) * 32 * (sub_lane_emax(N, XC, MODE) + 1) * sizeof(T)
                                : !(sub_stage_in<T>(N, MODE, SOA.value) || sub_stage_out(MODE, SOA.value))
                                    ? 0
                                    : (size_t)(kSubThreads / 32) * (32 / G) * stage_stride<N * N>() * sizeof(T);
            if (smem > 48 * 1024) {
                static std::atomic<uint64_t> raised[2];   // per instantiation (with / without the L part), one bit per device
                raise_dyn_smem(kern, smem, raised[(CAN_NO_L && LU == nullptr) ? 1 : 0]);
            }
            constexpr int MINB = sub_min_blocks(N, XC, (int)sizeof(T) / 4, MODE, SOA.value, COLPIV);
            const int pf = g_sub_prefetch.load(std::memory_order_relaxed);
            // Measured (profiles/r2ab_l2_prefetch.md): large gains in the SoA container up to 400-byte matrices and for the wide
            // kernels up to 576 bytes (double 8x8 fused solve 0.58 -> 0.75 of the HBM peak, decomposition 0.73 -> 0.92, float 9x9
            // decomposition 0.74 -> 1.02), small ones for AoS records up to 200 bytes; beyond that the prefetch bursts cost more
            // than the latency they hide (float 11x11 decomposition 0.77 -> 0.70): off there unless the option forces a distance.
            constexpr int REC = N * N * (int)sizeof(T);
            constexpr bool PF_ON = G == 1 && (SOA.value ? (REC <= 400 || (MINB == 2 && REC <= 576)) : REC <= 200);
            const int pf_dist = (G != 1 || pf == 0) ? 0 : pf > 0 ? pf : PF_ON ? sm_count() * MINB : 0;
            kern<<<grid, kSubThreads, smem, st>>>(A, Bm, X, LU, ok, perm, parity, degen, skip, B, pf_dist);
        });
    });
    return after_launch();
}

template <typename T>
int subwarp_plu_solve_impl(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* LU, int skip, int layout,
                           int rm, cudaStream_t st) {
    if (m != 1) return GMB_ERR_UNSUPPORTED;
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        rc = launch_sub<T, N.value, 1, MODE_SOLVE>(A, Bm, X, LU, ok, nullptr, nullptr, nullptr, skip, B, layout, rm, rm, st);
    });
    return rc;
}

// m = 2..4 right-hand sides: the same kernel with XC = m rider columns (instantiated in their own translation units)
template <typename T, int XC>
int subwarp_plu_solve_m_impl(int n, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* LU, int skip, int layout, int rm,
                             cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        rc = launch_sub<T, N.value, XC, MODE_SOLVE>(A, Bm, X, LU, ok, nullptr, nullptr, nullptr, skip, B, layout, rm, rm, st);
    });
    return rc;
}

template <typename T>
int subwarp_decomp_impl(int n, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout, int rm,
                        cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        rc = launch_sub<T, N.value, 0, MODE_DECOMP>(A, nullptr, nullptr, nullptr, nullptr, perm, parity, degen, 0, B, layout,
                                                    rm, rm, st);
    });
    return rc;
}

// decomp_lup for n >= 5: the column-pivot variant runs on the transposed view (a_rm flipped)
template <typename T>
int subwarp_lup_impl(int n, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout, int rm,
                     cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        rc = launch_sub<T, N.value, 0, MODE_DECOMP, true>(A, nullptr, nullptr, nullptr, nullptr, perm, parity, degen, 0, B, layout,
                                                          rm ? 0 : 1, rm ? 0 : 1, st);
    });
    return rc;
}

// inv for n >= 5 (MatrixInv.h:243-249 + 219-231): the working copy has the DESTINATION's layout and
// is then treated as row-major, i.e. the kernel reads A with a_rm = (src layout == dst layout)
// and writes the result flat row-major into the destination buffer.
// orthogonal / nullspace (MODE_RECT) for the sizes whose matrix fits one lane (float N <= 10, double N <= 7): nb vectors of N
template <typename T>
int subwarp_rect_impl(int N_, int nb, int64_t B, const T* bases, T* null_basis, uint8_t* ok, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    if (nb < 1 || nb >= N_) return rc;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14>(N_, [&](auto N) {
        if constexpr (sub_group(N.value, 0, (int)sizeof(T) / 4) == 1) {
            rc = launch_sub<T, N.value, 0, MODE_RECT>(bases, nullptr, null_basis, nullptr, ok, nullptr, nullptr, nullptr, nb, B, layout,
                                                      1, 1, st);
        }
    });
    return rc;
}

// deferred-rider inverse (MODE_INVD): the sizes whose matrix alone fits one lane (float n <= 10, double n <= 7)
template <typename T>
int subwarp_invd_impl(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm,
                      cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    const int a_rm = ((src_rm != 0) == (dst_rm != 0)) ? 1 : 0;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14>(n, [&](auto N) {
        if constexpr (sub_group(N.value, 0, (int)sizeof(T) / 4) == 1) {
            rc = launch_sub<T, N.value, 0, MODE_INVD>(A, nullptr, Ainv, nullptr, ok, nullptr, nullptr, nullptr, 0, B, layout,
                                                      a_rm, 1, st);
        }
    });
    return rc;
}

template <typename T>
int subwarp_inv_impl(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm,
                     cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    const int a_rm = ((src_rm != 0) == (dst_rm != 0)) ? 1 : 0;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        rc = launch_sub<T, N.value, N.value, MODE_INV>(A, nullptr, Ainv, nullptr, ok, nullptr, nullptr, nullptr, 0, B,
                                                       layout, a_rm, 1, st);
    });
    return rc;
}

}  // namespace gmb
