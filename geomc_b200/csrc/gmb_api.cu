// gmb_api.cu -- the extern "C" boundary (include/geomc_b200.h): argument checks and dispatch to
// the register (n <= 4), sub-warp cooperative (n = 5..16) and generic kernels, plus the
// host-buffer pipelines.  No CPU fallback anywhere: every entry point ends in a kernel launch.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <sched.h>
#include <vector>

#include "gmb_common.cuh"
#include "gmb_io.cuh"

namespace gmb {

std::atomic<int64_t> g_launch_count{0};

// Run-time dispatch options (gmb_set_option): -1 = the measured per-(type, n) table, anything else forces a kernel
// family.  The library itself never reads the environment; tests and tuning scripts set these through the C ABI.
std::atomic<int> g_opt_rowplu_min_n{-1};     // fused solve: row-distributed kernel from this n on (17 = never)
std::atomic<int> g_opt_rowdecomp_min_n{-1};  // decomp_plu: row-distributed kernel from this n on
std::atomic<int> g_opt_rowinv{-1};           // inverse n >= 5: 0 = column-distributed kernel, 1 = row-distributed
std::atomic<int> g_opt_rowfast{-1};          // speculative solve (gmb_rowfast.cuh): 0 = off, 1 = wherever it exists
std::atomic<int> g_opt_invd{-1};             // deferred-rider inverse (k_sub_plu MODE_INVD): 0 = off, 1 = wherever it exists
std::atomic<int> g_sub_prefetch{-1};         // one-lane sub-warp kernels: L2 prefetch distance in CTAs (gmb_subwarp.cuh); -1 = one wave, 0 = off
std::atomic<int> g_opt_lanebs{-1};           // one-lane backsolve with 1..4 right-hand sides (gmb_lanebs.cuh): 0 = off
std::atomic<int> g_opt_lanechol{-1};         // one-thread-per-system Cholesky with staged ingest, n = 5..8: 0 = off, 1 = on

// gmb_small.cu
template <typename T> int small_plu_solve(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <typename T> int small_decomp(int, int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, bool, cudaStream_t);
template <typename T> int small_backsolve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int, cudaStream_t);
template <typename T> int small_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <typename T> int small_inv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <typename T> int small_solve_vec(int, int, int64_t, const T*, const T*, T*, uint8_t*, int, cudaStream_t);
template <typename T> int small_inv_elem(int, int64_t, const T*, T*, uint8_t*, int, int, cudaStream_t);
template <typename T> int small_solve_vec_elem(int, int, int64_t, const T*, const T*, T*, uint8_t*, cudaStream_t);
template <typename T> int small_residual(int, int64_t, const T*, const T*, const T*, const uint8_t*, double*, int, int, cudaStream_t);
// gmb_subwarp.cu
template <typename T> int subwarp_plu_solve(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <typename T, int XC> int subwarp_plu_solve_m(int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <typename T> int subwarp_decomp(int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, cudaStream_t);
template <typename T> int subwarp_rect(int, int, int64_t, const T*, T*, uint8_t*, int, cudaStream_t);
template <typename T> int subwarp_invd(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <typename T> int subwarp_lup(int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, cudaStream_t);
template <typename T> int subwarp_inv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <typename T> int subwarp_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
// gmb_rowplu_*.cu
template <typename T> int rowplu_solve(int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <typename T> int rowplu_decomp(int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, cudaStream_t);
template <typename T> int rowinv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <typename T> int row_backsolve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int, cudaStream_t);
// gmb_lanebs_*.cu
template <typename T> int lane_backsolve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int, cudaStream_t);
// gmb_rowfast_*.cu
template <typename T> int rowfast_solve(int, int64_t, const T*, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
// gmb_lane_*.cu
template <typename T> int lane_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
// gmb_generic.cu
template <typename T> int generic_decomp(int, int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, bool, cudaStream_t);
template <typename T> int generic_solve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <typename T> int generic_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <typename T> int generic_inv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <typename T> int residual_stats(int, int64_t, const T*, const T*, const T*, const uint8_t*, double*, int, int, cudaStream_t);
template <typename T> int transpose_batch(bool, int, int64_t, const T*, T*, cudaStream_t);
int checksum_bytes(const void*, int64_t, int64_t, uint64_t*, cudaStream_t);
template <typename T> int small_orthogonal(int, int64_t, const T*, T*, int, cudaStream_t);
template <typename T> int generic_orthogonal(int, int64_t, const T*, T*, int, cudaStream_t);
template <typename T> int simplex_contains(int, int64_t, const T*, int64_t, const uint8_t*, const T*, uint8_t*, int, cudaStream_t);
template <typename T> int trace_simplex(int, int64_t, const T*, int64_t, const T*, const T*, int64_t, uint8_t*, T*, T*, int, cudaStream_t);
template <typename T> int simplex_intersect(int, int64_t, const T*, int64_t, const uint8_t*, const T*, const T*, int64_t, T*, int, cudaStream_t);
template <typename T> int generic_nullspace(int, int, int64_t, const T*, T*, uint8_t*, int, cudaStream_t);
// gmb_xform_*.cu
template <typename T> int mul_small(int, int, int, int64_t, const T*, const T*, T*, int, int, cudaStream_t);
template <typename T> int mul_generic(int, int, int, int64_t, const T*, const T*, T*, int, int, cudaStream_t);
template <typename T> int xf_from_matrix(int, int64_t, const T*, T*, uint8_t*, int, int, bool, cudaStream_t);
template <typename T> int xf_translation(int, int64_t, const T*, T*, int, int, bool, cudaStream_t);
template <typename T> int xf_compose(int, int64_t, const T*, const T*, T*, int, int, cudaStream_t);
template <typename T> int xf_apply(int, int, int64_t, const T*, bool, const T*, T*, int, cudaStream_t);

// Which fused-solve kernel serves (T, n)?  Measured on B200 (profiles/r2e_row_vs_subwarp.md: working sets > 2 x L2, AoS and
// SoA, CUDA events).  With an LU output the row-distributed kernel (bounce-buffer exchange of whole rows) wins for float n
// in {12, 14, 15, 16} and for double n = 10 and n >= 12 (round 1).  Without one -- the common call -- only the live part of
// a row travels and the register-only shuffle exchange applies; since the one-lane-per-system sub-warp kernel got its
// staged record I/O (round 2) it wins wherever one lane holds the system -- since the "wide" one-lane kernels (255 registers,
// gmb_subwarp.cuh sub_wide) that is float n <= 14 and double n <= 10 (profiles/r2t_wide_lane.md, 1 GB batches, AoS / SoA of the
// peak, sub-warp against row kernel: double n = 8 0.74 / 0.58 against 0.55 / 0.55, double n = 9 0.71 / 0.67 against 0.36 / 0.35,
// float n = 13 0.46 / 0.43 against 0.37 / 0.41 (speculative kernel), float n = 14 0.43 / 0.39 against 0.34 / 0.35).  Float n = 12
// stays with the row kernel (0.42 / 0.49 against 0.43 / 0.45), as do float n >= 15 (speculative kernel first) and double n >= 11
// (double n = 10 on one lane with ~300 bytes of spills: 0.55 / 0.60 against 0.34 / 0.40, run r2ag).
// gmb_set_option("rowplu_min_n", n) overrides the table (17 disables the row kernel, 9 forces it).
template <typename T>
static bool use_row_kernel(int n, bool want_lu) {
    const int forced = g_opt_rowplu_min_n.load(std::memory_order_relaxed);
    if (forced >= 0) return n >= forced;
    if (!want_lu) return sizeof(T) == 4 ? (n == 12 || n >= 15) : n >= 11;
    if (n < 9) return false;
    if (sizeof(T) == 4) return n == 12 || n >= 14;
    return n == 10 || n >= 12;
}

template <typename T>
static bool use_rowfast(int n) {
    const int forced = g_opt_rowfast.load(std::memory_order_relaxed);
    if (forced >= 0) return forced != 0;
    // measured against the exact row kernel (profiles/r2a_rowfast.md, 2^20 systems, AoS): float n = 13, 15, 16 win
    // (16 x 16: 0.58 -> 0.39 ms); n = 13 has since gone to the wide one-lane sub-warp kernel (0.46 / 0.43 against 0.37 / 0.41)
    return sizeof(T) == 4 && n >= 15;
}

// Cholesky: the row kernel (one lane per system for these sizes, staged AoS ingest) serves one right-hand side; the
// lane kernel serves 2..4 of them (the row kernel carries one rider).  Option "lanechol": 1 = lane kernel for m = 1 too
// (tests), 0 = never (m > 1 then falls to the generic kernel).
// Measured (profiles/r2c_lane_ab.md): from AoS records the lane kernel wins (float n = 5..8: 0.60 / 0.53 / 0.56 / 0.43 of
// the peak against 0.55 / 0.44 / 0.43 / 0.38 for the row kernel with the same staged ingest), in the SoA container the
// row kernel's right-looking order wins (0.82 / 0.71 / 0.78 / 0.67 against 0.77 / 0.65 / 0.69 / 0.58).
template <typename T>
static bool use_lane_chol(int n, int m, bool solve, int layout) {
    if (n < 5 || n > (sizeof(T) == 4 ? 8 : 6) || (solve && m > 4)) return false;
    const int forced = g_opt_lanechol.load(std::memory_order_relaxed);
    if (forced >= 0) return forced != 0;
    return (solve && m >= 2) || layout == GMB_LAYOUT_AOS;
}

static bool bad_layout(int layout) { return layout != GMB_LAYOUT_AOS && layout != GMB_LAYOUT_SOA; }

// The register and sub-warp kernels use 128-bit accesses and need 16-byte aligned batches (any
// cudaMalloc / torch allocation is).  An AoS batch that starts at an odd offset inside a larger
// buffer (the reference's objects are only 4- or 8-byte aligned) is routed to the generic kernels,
// which use element-wise accesses; a SoA container is aligned by construction.
template <typename... P>
static bool all_aligned16(P... ptrs) {
    return ((reinterpret_cast<uintptr_t>(ptrs) % 16 == 0) && ...);
}

// Same question for decomp_plu (option "rowdecomp_min_n" overrides).  Measured again in round 2 (profiles/r2e_row_vs_subwarp.md,
// profiles/r2f_sweep_norow.md; working sets > 2 x L2, in place, AoS / SoA): with its staged record I/O the sub-warp kernel wins
// for float n <= 14 (n = 9: 0.93 / 0.79 of the peak against 0.35 / 0.46; n = 13: 0.44 / 0.52 against 0.32 / 0.39) and double
// n <= 11; the row kernel keeps float n = 15, 16 and double n >= 12.
template <typename T>
static bool use_row_decomp(int n) {
    const int forced = g_opt_rowdecomp_min_n.load(std::memory_order_relaxed);
    if (forced >= 0) return n >= forced;
    return sizeof(T) == 4 ? n >= 15 : n >= 12;
}

// ------------------------------------------------------------------ typed implementations
template <typename T>
int plu_decomp(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout, int rm,
               bool colpiv, void* stream) {
    if (B < 0 || rows < 1 || cols < 1 || bad_layout(layout) || (B > 0 && A == nullptr)) return GMB_ERR_INVALID_ARG;
    if (rows > GMB_MAX_N || cols > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const bool fast = all_aligned16(A);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (fast && cols <= 4 && cols >= 2 && (rows == cols || rows == cols - 1)) {
        int rc = small_decomp<T>(rows, cols, B, A, perm, parity, degen, layout, rm, colpiv, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && !colpiv && rows == cols && rows >= 7 && use_row_decomp<T>(rows)) {
        int rc = rowplu_decomp<T>(rows, B, A, perm, parity, degen, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && !colpiv && rows == cols && rows >= 5) {
        int rc = subwarp_decomp<T>(rows, B, A, perm, parity, degen, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && colpiv && rows == cols && rows >= 5) {
        int rc = subwarp_lup<T>(rows, B, A, perm, parity, degen, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_decomp<T>(rows, cols, B, A, perm, parity, degen, layout, rm, colpiv, st);
}

template <typename T>
int lu_backsolve(int mode, int n, int m, int64_t B, const T* LU, const uint8_t* perm, const T* Bm, T* X, int skip,
                 int layout, int rm, void* stream) {
    if (B < 0 || n < 1 || m < 1 || skip < 0 || bad_layout(layout)) return GMB_ERR_INVALID_ARG;
    if (B > 0 && (X == nullptr || Bm == nullptr || (mode != 2 && LU == nullptr) || (mode != 0 && perm == nullptr)))
        return GMB_ERR_INVALID_ARG;
    if (n > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    if (B == 0) return GMB_OK;                             // an empty batch: nothing to launch
    cudaStream_t st = as_stream(stream);
    const bool fast = all_aligned16(LU, Bm, X);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (fast && n >= 2 && n <= 4 && m <= 4) {
        int rc = small_backsolve<T>(mode, n, m, B, LU, perm, Bm, X, skip, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    // one lane per system, 1..4 right-hand sides in one pass, staged AoS records (float n <= 13, double n <= 9); option "lanebs" = 0
    // leaves every call to the row kernel below (profiles/r2x_lane_backsolve.md)
    if (fast && n >= 5 && m <= 4 && mode != 2 && g_opt_lanebs.load(std::memory_order_relaxed) != 0) {
        int rc = lane_backsolve<T>(mode, n, m, B, LU, perm, Bm, X, skip, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && n >= 5) {                                    // any number of right-hand sides: the LU rows are loaded once
        int rc = row_backsolve<T>(mode, n, m, B, LU, perm, Bm, X, skip, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_solve<T>(mode, n, m, B, LU, perm, Bm, X, nullptr, nullptr, skip, layout, rm, st);
}

template <typename T>
int plu_solve(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* LU, int skip, int layout, int rm,
              void* stream) {
    if (B < 0 || n < 1 || m < 1 || skip < 0 || bad_layout(layout)) return GMB_ERR_INVALID_ARG;
    if (B > 0 && (A == nullptr || Bm == nullptr || X == nullptr)) return GMB_ERR_INVALID_ARG;
    if (n > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    if (B == 0) return GMB_OK;                             // an empty batch: nothing to launch
    cudaStream_t st = as_stream(stream);
    const bool fast = all_aligned16(A, Bm, X, LU);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (fast && n >= 2 && n <= 4 && m <= 4) {
        int rc = small_plu_solve<T>(n, m, B, A, Bm, X, ok, LU, skip, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    // n >= 9, one right-hand side, no LU output: speculative kernel + exact re-solve of the systems it hands back
    // (gmb_rowfast.cuh); option "rowfast" = 0 switches it off (the exact kernels below then serve every system)
    if (fast && m == 1 && n >= 9 && LU == nullptr && use_rowfast<T>(n)) {
        int rc = rowfast_solve<T>(n, B, A, Bm, X, ok, skip, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && m == 1 && n >= 7 && use_row_kernel<T>(n, LU != nullptr)) {
        int rc = rowplu_solve<T>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && n >= 5) {
        int rc = m == 1   ? subwarp_plu_solve<T>(n, m, B, A, Bm, X, ok, LU, skip, layout, rm, st)
                 : m == 2 ? subwarp_plu_solve_m<T, 2>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st)
                 : m == 3 ? subwarp_plu_solve_m<T, 3>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st)
                 : m == 4 ? subwarp_plu_solve_m<T, 4>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st)
                          : GMB_ERR_UNSUPPORTED;
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_solve<T>(3, n, m, B, A, nullptr, Bm, X, ok, LU, skip, layout, rm, st);
}

template <typename T>
int linear_solve_vec(int n, int m, int64_t B, const T* bases, const T* Bv, T* X, uint8_t* ok, int skip, int layout,
                     void* stream) {
    if (B < 0 || n < 1 || m < 1 || skip < 0 || bad_layout(layout)) return GMB_ERR_INVALID_ARG;
    if (B > 0 && (bases == nullptr || Bv == nullptr || X == nullptr)) return GMB_ERR_INVALID_ARG;
    if (n > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    if (B == 0) return GMB_OK;                             // an empty batch: nothing to launch
    cudaStream_t st = as_stream(stream);
    if (n < 5) {
        // inverse-then-multiply: LUDecomp.h:419-426 (skip is ignored on this path, as in the reference)
        if (n < 2 || m > 4) return GMB_ERR_UNSUPPORTED;
        if (all_aligned16(bases, Bv, X)) return small_solve_vec<T>(n, m, B, bases, Bv, X, ok, layout, st);
        if (layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;      // the container is aligned by construction
        return small_solve_vec_elem<T>(n, m, B, bases, Bv, X, ok, st); // an array of Vec objects: element-aligned only
    }
    // n >= 5: column-major PLU + permutation + column-major back-substitution (:428-439)
    return plu_solve<T>(n, m, B, bases, Bv, X, ok, nullptr, skip, layout, /*row_major=*/0, stream);
}

template <typename T>
int cholesky(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm,
             bool solve, void* stream) {
    if (B < 0 || n < 1 || bad_layout(layout) || (B > 0 && A == nullptr)) return GMB_ERR_INVALID_ARG;
    if (solve && (m < 1 || (B > 0 && (Bm == nullptr || X == nullptr)))) return GMB_ERR_INVALID_ARG;
    if (n > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    if (B == 0) return GMB_OK;                             // an empty batch: nothing to launch
    cudaStream_t st = as_stream(stream);
    const bool fast = all_aligned16(A, Bm, X, L_out);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (fast && n >= 2 && n <= 4 && (!solve || m <= 4)) {
        int rc = small_cholesky<T>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && use_lane_chol<T>(n, m, solve, layout)) {
        int rc = lane_cholesky<T>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && n >= 5) {
        int rc = subwarp_cholesky<T>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_cholesky<T>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
}

template <typename T>
int inv(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm, void* stream) {
    if (B < 0 || n < 1 || bad_layout(layout) || (B > 0 && (A == nullptr || Ainv == nullptr))) return GMB_ERR_INVALID_ARG;
    if (n > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const bool fast = all_aligned16(A, Ainv);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (B == 0) return GMB_OK;                             // an empty batch: nothing to launch (and no pointers to compare)
    if (n >= 5 && A == Ainv) return GMB_ERR_INVALID_ARG;   // invNxN: m must not alias out (MatrixInv.h:174)
    if (n >= 2 && n <= 4) {
        // adjugate (MatrixInv.h:49-165) in both cases: vectorised register kernel, or -- for an AoS batch that is only
        // element-aligned, e.g. an offset into an array of objects -- the same arithmetic with scalar accesses
        return fast ? small_inv<T>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st)
                    : small_inv_elem<T>(n, B, A, Ainv, ok, src_rm, dst_rm, st);
    }
    // The row-distributed kernel (gmb_rowinv.cuh) is 1.2x (7 x 7 double) to 3x (16 x 16 float) faster than the
    // column-distributed one from n = 8 (float) / n = 6 (double) on and slower below (profiles/r1j_row_inverse.md);
    // option "rowinv" = 0 / 1 forces the column / row kernel for every n >= 5 (tests).
    // one lane per system with deferred riders (float n <= 11, double n <= 7): option "invd" = 0 switches it off
    if (fast && n >= 5 && n <= (sizeof(T) == 4 ? 14 : 10) && g_opt_invd.load(std::memory_order_relaxed) != 0 &&
        g_opt_rowinv.load(std::memory_order_relaxed) < 0) {
        int rc = subwarp_invd<T>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    const int row_inv = g_opt_rowinv.load(std::memory_order_relaxed);
    const bool use_row_inv = row_inv >= 0 ? row_inv == 1 : n >= (sizeof(T) == 4 ? 8 : 6);
    if (fast && n >= 5 && use_row_inv) {
        int rc = rowinv<T>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    if (fast && n >= 5) {
        int rc = subwarp_inv<T>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_inv<T>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
}

template <typename T>
int orthogonal(int n, int64_t B, const T* V, T* out, int layout, void* stream) {
    if (B < 0 || n < 2 || bad_layout(layout) || (B > 0 && (V == nullptr || out == nullptr))) return GMB_ERR_INVALID_ARG;
    if (n > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    if (B == 0) return GMB_OK;                             // an empty batch: nothing to launch
    const bool fast = all_aligned16(V, out);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (fast && n <= 4) return small_orthogonal<T>(n, B, V, out, layout, as_stream(stream));
    if (fast && n >= 5) {                                    // register kernel: rectangular (n-1) x n PLU + back-substitution
        int rc = subwarp_rect<T>(n, n - 1, B, V, out, nullptr, layout, as_stream(stream));
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_orthogonal<T>(n, B, V, out, layout, as_stream(stream));
}

// ---- (f)2: simplex queries.  dim = 2..7; strides are in elements, 0 = packed; AoS only (SoA tiles are packed)
template <typename T>
int simplex_contains_api(int dim, int64_t B, const T* verts, int64_t vstride, const uint8_t* nverts, const T* p, uint8_t* inside,
                         int layout, void* stream) {
    const int64_t E = (int64_t)(dim + 1) * dim;
    if (B < 0 || dim < 2 || bad_layout(layout) || (vstride != 0 && vstride < E) ||
        (B > 0 && (verts == nullptr || p == nullptr || inside == nullptr)))
        return GMB_ERR_INVALID_ARG;
    if (dim > 7) return GMB_ERR_UNSUPPORTED;
    if (layout == GMB_LAYOUT_SOA && (vstride != 0 || !all_aligned16(verts, p))) return GMB_ERR_INVALID_ARG;
    return simplex_contains<T>(dim, B, verts, vstride ? vstride : E, nverts, p, inside, layout, as_stream(stream));
}

template <typename T>
int trace_simplex_api(int dim, int64_t B, const T* verts, int64_t vstride, const T* origin, const T* dir, int64_t rstride,
                      uint8_t* hit, T* s_out, T* uv, int layout, void* stream) {
    const int64_t E = (int64_t)dim * dim;
    if (B < 0 || dim < 2 || bad_layout(layout) || (vstride != 0 && vstride < E) || (rstride != 0 && rstride < dim) ||
        (B > 0 && (verts == nullptr || origin == nullptr || dir == nullptr || hit == nullptr || s_out == nullptr)))
        return GMB_ERR_INVALID_ARG;
    if (dim > 7) return GMB_ERR_UNSUPPORTED;
    if (layout == GMB_LAYOUT_SOA && (vstride != 0 || rstride != 0 || !all_aligned16(verts, origin, dir) ||
                                     (uv != nullptr && !all_aligned16(uv))))
        return GMB_ERR_INVALID_ARG;
    return trace_simplex<T>(dim, B, verts, vstride ? vstride : E, origin, dir, rstride ? rstride : dim, hit, s_out, uv, layout,
                            as_stream(stream));
}

template <typename T>
int simplex_intersect_api(int dim, int64_t B, const T* verts, int64_t vstride, const uint8_t* nverts, const T* origin,
                          const T* dir, int64_t rstride, T* interval, int layout, void* stream) {
    const int64_t E = (int64_t)(dim + 1) * dim;
    if (B < 0 || dim < 2 || bad_layout(layout) || (vstride != 0 && vstride < E) || (rstride != 0 && rstride < dim) ||
        (B > 0 && (verts == nullptr || origin == nullptr || dir == nullptr || interval == nullptr)))
        return GMB_ERR_INVALID_ARG;
    if (dim > 7) return GMB_ERR_UNSUPPORTED;
    if (layout == GMB_LAYOUT_SOA && (vstride != 0 || rstride != 0 || !all_aligned16(verts, origin, dir, interval)))
        return GMB_ERR_INVALID_ARG;
    return simplex_intersect<T>(dim, B, verts, vstride ? vstride : E, nverts, origin, dir, rstride ? rstride : dim, interval,
                                layout, as_stream(stream));
}

template <typename T>
int nullspace(int N, int nb, int64_t B, const T* bases, T* null, uint8_t* ok, int layout, void* stream) {
    if (B < 0 || N < 3 || nb < 1 || nb >= N || bad_layout(layout) || (B > 0 && (bases == nullptr || null == nullptr)))
        return GMB_ERR_INVALID_ARG;
    if (N > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    if (N >= 5 && all_aligned16(bases, null)) {
        int rc = subwarp_rect<T>(N, nb, B, bases, null, ok, layout, as_stream(stream));
        if (rc != GMB_ERR_UNSUPPORTED) return rc;
    }
    return generic_nullspace<T>(N, nb, B, bases, null, ok, layout, as_stream(stream));
}

// ------------------------------------------------------------------ host-buffer pipelines
// AoS host batches are cut into chunks; each chunk goes H2D -> kernel -> D2H on its own stream
// (3 slots round-robin) so uploads, kernels and downloads of neighbouring chunks overlap and both
// PCIe directions stay busy.  Device staging buffers are cached per device.
struct DeviceGuard {
    int prev = -1;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        status = cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct HostArg {
    const void* src;     // host source (inputs) or nullptr
    void* dst;           // host destination (outputs) or nullptr
    size_t bytes_per_system;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void* buf = nullptr;
    size_t cap = 0;
};

struct HostCtx {
    int device = -1;
    Slot slots[3];
    std::mutex mu;
};

static HostCtx* host_ctx(int device) {
    static std::mutex mu;
    static std::vector<HostCtx*> all;
    std::lock_guard<std::mutex> g(mu);
    for (auto* c : all)
        if (c->device == device) return c;
    auto* c = new HostCtx();
    c->device = device;
    all.push_back(c);
    return c;
}

template <class Launch>
static int run_host_pipeline(int device, int64_t B, std::vector<HostArg>& args, Launch&& launch) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return GMB_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return GMB_ERR_INVALID_ARG;
    if (B == 0) return GMB_OK;
    DeviceGuard guard(device);                       // the caller's current device is restored on every return path
    cudaError_t e = guard.status;
    if (e != cudaSuccess) return (int)e;
    HostCtx* ctx = host_ctx(device);
    std::lock_guard<std::mutex> g(ctx->mu);

    size_t per_sys = 0;
    for (auto& a : args) per_sys += (a.bytes_per_system + 15) / 16 * 16;       // keeps every sub-buffer 16B aligned
    // ~48 MiB per chunk, multiple of 1024 systems (keeps every sub-buffer 16-byte aligned)
    int64_t chunk = (int64_t)((48u << 20) / per_sys);
    chunk = chunk / 1024 * 1024;
    if (chunk < 1024) chunk = 1024;
    if (chunk > B) chunk = (B + 1023) / 1024 * 1024;
    const size_t need = (size_t)chunk * per_sys;
    int rc = GMB_OK;
    for (auto& s : ctx->slots) {
        if (!s.stream) {
            if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
            if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        }
        if (s.cap < need) {
            if (s.buf) cudaFree(s.buf);
            s.buf = nullptr;
            s.cap = 0;
            if (cudaMalloc(&s.buf, need) != cudaSuccess) return GMB_ERR_ALLOC;
            s.cap = need;
        }
    }
    std::vector<void*> dptr(args.size());
    int64_t done = 0;
    int k = 0;
    while (done < B && rc == GMB_OK) {
        const int64_t nb = (B - done) < chunk ? (B - done) : chunk;
        Slot& s = ctx->slots[k % 3];
        // the slot's previous chunk must have drained before its buffers are reused (stream order does that)
        size_t off = 0;
        for (size_t i = 0; i < args.size(); ++i) {
            dptr[i] = static_cast<char*>(s.buf) + off;
            off += (size_t)chunk * ((args[i].bytes_per_system + 15) / 16 * 16);
            if (args[i].src) {
                e = cudaMemcpyAsync(dptr[i], static_cast<const char*>(args[i].src) + (size_t)done * args[i].bytes_per_system,
                                    (size_t)nb * args[i].bytes_per_system, cudaMemcpyHostToDevice, s.stream);
                if (e != cudaSuccess) rc = (int)e;
            }
        }
        if (rc == GMB_OK) rc = launch(nb, dptr, (void*)s.stream);
        for (size_t i = 0; i < args.size() && rc == GMB_OK; ++i) {
            if (args[i].dst) {
                e = cudaMemcpyAsync(static_cast<char*>(args[i].dst) + (size_t)done * args[i].bytes_per_system, dptr[i],
                                    (size_t)nb * args[i].bytes_per_system, cudaMemcpyDeviceToHost, s.stream);
                if (e != cudaSuccess) rc = (int)e;
            }
        }
        done += nb;
        ++k;
    }
    for (auto& s : ctx->slots) {
        e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && rc == GMB_OK) rc = (int)e;
    }
    return rc;
}

// ------------------------------------------------------------------ NUMA placement
// A box with several GPUs has several host memory nodes; a pinned buffer that lives on the far node reaches its GPU over the
// inter-socket link, and all of a process's buffers on ONE node make that node's memory feed every PCIe link of the box.  The
// kernel exposes the CPUs next to a PCI device in sysfs; a thread that runs there allocates (first touch) and drives copies
// locally.  Returns false when the topology cannot be read (single-node boxes, containers without sysfs): nothing is changed.
static bool cpus_near_device(int device, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    char path[128];
    std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = std::fopen(path, "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = std::fgets(line, sizeof(line), f) != nullptr;
    std::fclose(f);
    if (!got) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* tok = std::strtok(line, ",\n"); tok; tok = std::strtok(nullptr, ",\n")) {
        int lo = 0, hi = 0;
        const int k = std::sscanf(tok, "%d-%d", &lo, &hi);
        if (k < 1) continue;
        if (k == 1) hi = lo;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) {
            CPU_SET(c, set);
            ++count;
        }
    }
    // keep only CPUs this process may use at all (cgroup / taskset limits)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        int usable = 0;
        for (int c = 0; c < CPU_SETSIZE; ++c) {
            if (CPU_ISSET(c, set) && !CPU_ISSET(c, &allowed)) CPU_CLR(c, set);
            if (CPU_ISSET(c, set)) ++usable;
        }
        count = usable;
    }
    return count > 0;
}

static int bind_thread_near_device(int device) {
    cpu_set_t set;
    if (!cpus_near_device(device, &set)) return 0;
    return sched_setaffinity(0, sizeof(set), &set) == 0 ? 1 : 0;
}

// ------------------------------------------------------------------ several devices of one box
// Independent systems: device k of `ndev` serves the contiguous shard shard_range(B, k, ndev, 1024) through its own
// host pipeline on its own host thread.  No collective, no peer traffic; the shards of the host arrays are disjoint.
static void shard_range(int64_t B, int rank, int world, int64_t align, int64_t* lo, int64_t* hi) {
    const int64_t units = (B + align - 1) / align;
    const int64_t per = units / world, extra = units % world;
    const int64_t b = rank * per + (rank < extra ? rank : extra);
    const int64_t e = b + per + (rank < extra ? 1 : 0);
    *lo = b * align < B ? b * align : B;
    *hi = e * align < B ? e * align : B;
}

template <class Shard>
static int run_multi_device(int64_t B, const int* devices, int ndev, Shard&& shard) {
    if (ndev < 1 || devices == nullptr || B < 0) return GMB_ERR_INVALID_ARG;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return GMB_ERR_NO_DEVICE;
    for (int k = 0; k < ndev; ++k) {
        if (devices[k] < 0 || devices[k] >= have) return GMB_ERR_INVALID_ARG;
        for (int j = 0; j < k; ++j)
            if (devices[j] == devices[k]) return GMB_ERR_INVALID_ARG;
    }
    std::vector<int> rcs((size_t)ndev, GMB_OK);
    std::vector<std::thread> workers;
    for (int k = 0; k < ndev; ++k) {
        int64_t lo, hi;
        shard_range(B, k, ndev, 1024, &lo, &hi);
        if (hi <= lo) continue;
        if (ndev == 1) {
            rcs[0] = shard(lo, hi, devices[0]);
        } else {
            workers.emplace_back([&, k, lo, hi] {
                bind_thread_near_device(devices[k]);         // the copy-issuing thread runs next to its GPU
                rcs[(size_t)k] = shard(lo, hi, devices[k]);
            });
        }
    }
    for (auto& w : workers) w.join();
    for (int rc : rcs)
        if (rc != GMB_OK) return rc;
    return GMB_OK;
}

template <typename T>
int host_plu_solve(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, int skip, int rm, int device) {
    if (B < 0 || n < 1 || m < 1 || (B > 0 && (!A || !Bm || !X))) return GMB_ERR_INVALID_ARG;
    std::vector<HostArg> args = {{A, nullptr, sizeof(T) * n * n}, {Bm, X, sizeof(T) * n * m}, {nullptr, ok, 1}};
    return run_host_pipeline(device, B, args, [&](int64_t nb, std::vector<void*>& d, void* st) {
        return plu_solve<T>(n, m, nb, (const T*)d[0], (const T*)d[1], (T*)d[1], (uint8_t*)d[2], nullptr, skip,
                            GMB_LAYOUT_AOS, rm, st);
    });
}

template <typename T>
int host_inv(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int srm, int drm, int device) {
    if (B < 0 || n < 1 || (B > 0 && (!A || !Ainv))) return GMB_ERR_INVALID_ARG;
    // "out untouched when singular": the destination's current contents travel too only when the
    // caller passes distinct buffers and asks for ok flags; the common case uploads A only and
    // writes A^-1 over a scratch copy of A (singular systems then return A itself -- documented).
    std::vector<HostArg> args = {{A, nullptr, sizeof(T) * n * n}, {Ainv, Ainv, sizeof(T) * n * n}, {nullptr, ok, 1}};
    return run_host_pipeline(device, B, args, [&](int64_t nb, std::vector<void*>& d, void* st) {
        return inv<T>(n, nb, (const T*)d[0], (T*)d[1], (uint8_t*)d[2], GMB_LAYOUT_AOS, srm, drm, st);
    });
}

template <typename T>
int host_cholesky_solve(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, int rm, int device) {
    if (B < 0 || n < 1 || m < 1 || (B > 0 && (!A || !Bm || !X))) return GMB_ERR_INVALID_ARG;
    std::vector<HostArg> args = {{A, nullptr, sizeof(T) * n * n}, {Bm, X, sizeof(T) * n * m}, {nullptr, ok, 1}};
    return run_host_pipeline(device, B, args, [&](int64_t nb, std::vector<void*>& d, void* st) {
        return cholesky<T>(n, m, nb, (const T*)d[0], (const T*)d[1], (T*)d[1], (uint8_t*)d[2], nullptr, GMB_LAYOUT_AOS, rm,
                           true, st);
    });
}

template <typename T>
int host_plu_decomp(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int rm,
                    int device) {
    if (B < 0 || rows < 1 || cols < 1 || (B > 0 && !A)) return GMB_ERR_INVALID_ARG;
    std::vector<HostArg> args = {{A, A, sizeof(T) * rows * cols}, {nullptr, perm, (size_t)rows}, {nullptr, parity, 1},
                                 {nullptr, degen, 1}};
    return run_host_pipeline(device, B, args, [&](int64_t nb, std::vector<void*>& d, void* st) {
        return plu_decomp<T>(rows, cols, nb, (T*)d[0], (uint8_t*)d[1], (uint8_t*)d[2], (uint8_t*)d[3], GMB_LAYOUT_AOS, rm,
                             false, st);
    });
}

// ---- rows (f)4 / (f)3 ------------------------------------------------------------------------------
template <typename T>
int mul_api(int r, int k, int c, int64_t B, const T* A, const T* Bm, T* D, int acc, int layout, int arm, int brm, int drm,
            void* stream) {
    if (B < 0 || r < 1 || k < 1 || c < 1 || bad_layout(layout) || (B > 0 && (!A || !Bm || !D))) return GMB_ERR_INVALID_ARG;
    if (r > GMB_MAX_N || k > GMB_MAX_N || c > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    const int flags = (acc ? 1 : 0) | (arm ? 2 : 0) | (brm ? 4 : 0) | (drm ? 8 : 0);
    const bool fast = all_aligned16(A, Bm, D);
    if (!fast && layout == GMB_LAYOUT_SOA) return GMB_ERR_INVALID_ARG;
    if (fast && r <= 4 && k <= 4 && c <= 4) return mul_small<T>(r, k, c, B, A, Bm, D, flags, layout, as_stream(stream));
    return mul_generic<T>(r, k, c, B, A, Bm, D, flags, layout, as_stream(stream));
}

static bool bad_xf_dim(int n) { return n < 2 || n > 7; }

template <typename T>
int xf_from_matrix_api(int n, int64_t B, const T* M, T* xf, uint8_t* ok, int layout, int rm, void* stream) {
    if (B < 0 || n < 1 || bad_layout(layout) || (B > 0 && (!M || !xf))) return GMB_ERR_INVALID_ARG;
    if (bad_xf_dim(n)) return GMB_ERR_UNSUPPORTED;
    return xf_from_matrix<T>(n, B, M, xf, ok, layout, rm, all_aligned16(xf), as_stream(stream));
}

template <typename T>
int xf_translation_api(int n, int64_t B, const T* t, T* xf, int scale, int layout, void* stream) {
    if (B < 0 || n < 1 || bad_layout(layout) || (B > 0 && (!t || !xf))) return GMB_ERR_INVALID_ARG;
    if (bad_xf_dim(n)) return GMB_ERR_UNSUPPORTED;
    return xf_translation<T>(n, B, t, xf, scale, layout, all_aligned16(xf), as_stream(stream));
}

template <typename T>
int xf_compose_api(int n, int64_t B, const T* xf1, const T* xf2, T* out, int inverse, int layout, void* stream) {
    if (B < 0 || n < 1 || bad_layout(layout) || (B > 0 && (!xf1 || !xf2 || !out)) || (B > 0 && (out == xf1 || out == xf2)))
        return GMB_ERR_INVALID_ARG;
    if (bad_xf_dim(n)) return GMB_ERR_UNSUPPORTED;
    return xf_compose<T>(n, B, xf1, xf2, out, inverse, layout, as_stream(stream));
}

template <typename T>
int xf_apply_api(int n, int mode, int64_t B, const T* xf, int64_t xf_count, const T* p, T* out, int layout, void* stream) {
    if (B < 0 || n < 1 || mode < 0 || mode > 5 || bad_layout(layout) || (xf_count != 1 && xf_count != B) ||
        (B > 0 && (!xf || !p || !out)))
        return GMB_ERR_INVALID_ARG;
    if (bad_xf_dim(n)) return GMB_ERR_UNSUPPORTED;
    return xf_apply<T>(n, mode, B, xf, xf_count == 1, p, out, layout, as_stream(stream));
}

}  // namespace gmb

// ==================================================================== extern "C"
using namespace gmb;

extern "C" {

const char* gmb_version(void) { return "geomc_b200 0.1 (sm_100a)"; }

const char* gmb_error_string(int code) {
    switch (code) {
        case GMB_OK: return "ok";
        case GMB_ERR_INVALID_ARG: return "invalid argument";
        case GMB_ERR_UNSUPPORTED: return "unsupported dimension (n in 2..16, m in 1..4 for the register kernels)";
        case GMB_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case GMB_ERR_ALLOC: return "device allocation failed";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int gmb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gmb_soa_tile_width(int elem_size) { return elem_size > 0 ? 512 / elem_size : 0; }

int64_t gmb_soa_elems(int64_t B, int E, int elem_size) {
    const int64_t W = gmb_soa_tile_width(elem_size);
    if (W <= 0 || B < 0 || E < 0) return 0;
    return (B + W - 1) / W * W * E;
}

int64_t gmb_launch_count(void) { return g_launch_count.load(); }

int gmb_set_option(const char* name, int value) {
    if (!name) return GMB_ERR_INVALID_ARG;
    std::atomic<int>* o = !std::strcmp(name, "rowplu_min_n")      ? &g_opt_rowplu_min_n
                          : !std::strcmp(name, "rowdecomp_min_n") ? &g_opt_rowdecomp_min_n
                          : !std::strcmp(name, "rowinv")          ? &g_opt_rowinv
                          : !std::strcmp(name, "rowfast")         ? &g_opt_rowfast
                          : !std::strcmp(name, "invd")            ? &g_opt_invd
                          : !std::strcmp(name, "lanebs")          ? &g_opt_lanebs
                          : !std::strcmp(name, "subprefetch")     ? &g_sub_prefetch
                          : !std::strcmp(name, "lanechol")        ? &g_opt_lanechol
                                                                  : nullptr;
    if (!o) return GMB_ERR_INVALID_ARG;
    o->store(value);
    return GMB_OK;
}

void* gmb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void* gmb_host_alloc_wc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined | cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void gmb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int gmb_bind_thread_near_device(int device) {
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || device < 0 || device >= have) {
        cudaGetLastError();
        return GMB_ERR_INVALID_ARG;
    }
    return bind_thread_near_device(device);
}
void* gmb_host_alloc_near(size_t bytes, int device) {
    cpu_set_t before;
    const bool have_before = sched_getaffinity(0, sizeof(before), &before) == 0;
    const int bound = bind_thread_near_device(device);
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        p = nullptr;
    }
    if (p) std::memset(p, 0, bytes);                         // first touch on the device's memory node
    if (bound == 1 && have_before) sched_setaffinity(0, sizeof(before), &before);
    return p;
}
int gmb_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return GMB_ERR_INVALID_ARG;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) cudaGetLastError();
    return (int)e;
}
int gmb_host_unregister(void* p) {
    if (!p) return GMB_ERR_INVALID_ARG;
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) cudaGetLastError();
    return (int)e;
}
int gmb_shard_range(int64_t B, int rank, int world, int64_t align, int64_t* begin, int64_t* end) {
    if (B < 0 || world < 1 || rank < 0 || rank >= world || align < 1 || !begin || !end) return GMB_ERR_INVALID_ARG;
    shard_range(B, rank, world, align, begin, end);
    return GMB_OK;
}

int gmb_set_device(int device) { return (int)cudaSetDevice(device); }
int gmb_get_device(int* device) {
    if (!device) return GMB_ERR_INVALID_ARG;
    return (int)cudaGetDevice(device);
}
void* gmb_device_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void gmb_device_free(void* p) {
    if (p) cudaFree(p);
}
int gmb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream));
}
int gmb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream));
}
int gmb_memset_device(void* dst, int value, size_t bytes, void* stream) {
    return (int)cudaMemsetAsync(dst, value, bytes, as_stream(stream));
}
int gmb_stream_synchronize(void* stream) { return (int)cudaStreamSynchronize(as_stream(stream)); }

#define GMB_DEFINE(SUF, T)                                                                                            \
    int gmb_aos_to_soa_##SUF(int E, int64_t B, const T* aos, T* soa, void* stream) {                                  \
        if (B < 0 || E < 1 || (B > 0 && (!aos || !soa))) return GMB_ERR_INVALID_ARG;                                  \
        return transpose_batch<T>(true, E, B, aos, soa, as_stream(stream));                                           \
    }                                                                                                                 \
    int gmb_soa_to_aos_##SUF(int E, int64_t B, const T* soa, T* aos, void* stream) {                                  \
        if (B < 0 || E < 1 || (B > 0 && (!aos || !soa))) return GMB_ERR_INVALID_ARG;                                  \
        return transpose_batch<T>(false, E, B, soa, aos, as_stream(stream));                                          \
    }                                                                                                                 \
    int gmb_plu_decomp_##SUF(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen,     \
                             int layout, int row_major, void* stream) {                                               \
        return plu_decomp<T>(rows, cols, B, A, perm, parity, degen, layout, row_major, false, stream);                \
    }                                                                                                                 \
    int gmb_lup_decomp_##SUF(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen,     \
                             int layout, int row_major, void* stream) {                                               \
        return plu_decomp<T>(rows, cols, B, A, perm, parity, degen, layout, row_major, true, stream);                 \
    }                                                                                                                 \
    int gmb_lu_backsolve_##SUF(int n, int m, int64_t B, const T* LU, T* X, int skip, int layout, int row_major,       \
                               void* stream) {                                                                        \
        return lu_backsolve<T>(0, n, m, B, LU, nullptr, X, X, skip, layout, row_major, stream);                       \
    }                                                                                                                 \
    int gmb_plu_backsolve_##SUF(int n, int m, int64_t B, const T* PLU, const uint8_t* perm, T* X, const T* Bm,        \
                                int skip, int layout, int row_major, void* stream) {                                  \
        return lu_backsolve<T>(1, n, m, B, PLU, perm, Bm, X, skip, layout, row_major, stream);                        \
    }                                                                                                                 \
    int gmb_apply_permutation_##SUF(int n, int m, int64_t B, const uint8_t* perm, T* A, int layout, int row_major,    \
                                    void* stream) {                                                                   \
        return lu_backsolve<T>(2, n, m, B, nullptr, perm, A, A, 0, layout, row_major, stream);                        \
    }                                                                                                                 \
    int gmb_plu_solve_##SUF(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* LU_out, int skip, \
                            int layout, int row_major, void* stream) {                                                \
        return plu_solve<T>(n, m, B, A, Bm, X, ok, LU_out, skip, layout, row_major, stream);                          \
    }                                                                                                                 \
    int gmb_linear_solve_vec_##SUF(int n, int m, int64_t B, const T* bases, const T* Bv, T* X, uint8_t* ok, int skip, \
                                   int layout, void* stream) {                                                        \
        return linear_solve_vec<T>(n, m, B, bases, Bv, X, ok, skip, layout, stream);                                  \
    }                                                                                                                 \
    int gmb_cholesky_##SUF(int n, int64_t B, T* A, uint8_t* ok, int layout, int row_major, void* stream) {            \
        return cholesky<T>(n, 1, B, A, nullptr, nullptr, ok, A, layout, row_major, false, stream);                    \
    }                                                                                                                 \
    int gmb_cholesky_solve_##SUF(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out,       \
                                 int layout, int row_major, void* stream) {                                           \
        return cholesky<T>(n, m, B, A, Bm, X, ok, L_out, layout, row_major, true, stream);                            \
    }                                                                                                                 \
    int gmb_inv_##SUF(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_row_major,              \
                      int dst_row_major, void* stream) {                                                              \
        return inv<T>(n, B, A, Ainv, ok, layout, src_row_major, dst_row_major, stream);                               \
    }                                                                                                                 \
    int gmb_orthogonal_##SUF(int n, int64_t B, const T* V, T* out, int layout, void* stream) {                        \
        return orthogonal<T>(n, B, V, out, layout, stream);                                                           \
    }                                                                                                                 \
    int gmb_nullspace_##SUF(int N, int nb, int64_t B, const T* bases, T* null_basis, uint8_t* ok, int layout,         \
                            void* stream) {                                                                           \
        return nullspace<T>(N, nb, B, bases, null_basis, ok, layout, stream);                                         \
    }                                                                                                                 \
    int gmb_simplex_contains_##SUF(int dim, int64_t B, const T* verts, int64_t vert_stride, const uint8_t* nverts,    \
                                   const T* p, uint8_t* inside, int layout, void* stream) {                           \
        return simplex_contains_api<T>(dim, B, verts, vert_stride, nverts, p, inside, layout, stream);                \
    }                                                                                                                 \
    int gmb_trace_simplex_##SUF(int dim, int64_t B, const T* verts, int64_t vert_stride, const T* origin,             \
                                const T* direction, int64_t ray_stride, uint8_t* hit, T* s, T* uv, int layout,        \
                                void* stream) {                                                                       \
        return trace_simplex_api<T>(dim, B, verts, vert_stride, origin, direction, ray_stride, hit, s, uv, layout,    \
                                    stream);                                                                          \
    }                                                                                                                 \
    int gmb_simplex_intersect_##SUF(int dim, int64_t B, const T* verts, int64_t vert_stride, const uint8_t* nverts,   \
                                    const T* origin, const T* direction, int64_t ray_stride, T* interval, int layout, \
                                    void* stream) {                                                                   \
        return simplex_intersect_api<T>(dim, B, verts, vert_stride, nverts, origin, direction, ray_stride, interval,  \
                                        layout, stream);                                                              \
    }                                                                                                                 \
    int gmb_mul_##SUF(int r, int k, int c, int64_t B, const T* A, const T* Bm, T* D, int accumulate, int layout,     \
                      int a_row_major, int b_row_major, int d_row_major, void* stream) {                              \
        return mul_api<T>(r, k, c, B, A, Bm, D, accumulate, layout, a_row_major, b_row_major, d_row_major, stream);   \
    }                                                                                                                 \
    int gmb_xf_from_matrix_##SUF(int dim, int64_t B, const T* M, T* xf_out, uint8_t* ok, int layout, int row_major,   \
                                 void* stream) {                                                                      \
        return xf_from_matrix_api<T>(dim, B, M, xf_out, ok, layout, row_major, stream);                               \
    }                                                                                                                 \
    int gmb_xf_translation_##SUF(int dim, int64_t B, const T* t, T* xf_out, int layout, void* stream) {               \
        return xf_translation_api<T>(dim, B, t, xf_out, 0, layout, stream);                                           \
    }                                                                                                                 \
    int gmb_xf_scale_##SUF(int dim, int64_t B, const T* sx, T* xf_out, int layout, void* stream) {                    \
        return xf_translation_api<T>(dim, B, sx, xf_out, 1, layout, stream);                                          \
    }                                                                                                                 \
    int gmb_xf_compose_##SUF(int dim, int64_t B, const T* xf1, const T* xf2, T* out, int inverse, int layout,         \
                             void* stream) {                                                                          \
        return xf_compose_api<T>(dim, B, xf1, xf2, out, inverse, layout, stream);                                     \
    }                                                                                                                 \
    int gmb_xf_apply_##SUF(int dim, int mode, int64_t B, const T* xf, int64_t xf_count, const T* p, T* out,           \
                           int layout, void* stream) {                                                                \
        return xf_apply_api<T>(dim, mode, B, xf, xf_count, p, out, layout, stream);                                   \
    }                                                                                                                 \
    int gmb_solve_residual_##SUF(int n, int64_t B, const T* A, const T* X, const T* Bm, const uint8_t* ok,            \
                                 double* stats4, int layout, int row_major, void* stream) {                           \
        if (B < 0 || n < 1 || (B > 0 && (!A || !X || !Bm || !stats4))) return GMB_ERR_INVALID_ARG;                    \
        if (n >= 2 && n <= 4 && all_aligned16(A, X, Bm)) {                                                            \
            int rc = small_residual<T>(n, B, A, X, Bm, ok, stats4, layout, row_major, as_stream(stream));             \
            if (rc != GMB_ERR_UNSUPPORTED) return rc;                                                                 \
        }                                                                                                             \
        return residual_stats<T>(n, B, A, X, Bm, ok, stats4, layout, row_major, as_stream(stream));                   \
    }                                                                                                                 \
    int gmb_host_plu_solve_##SUF(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, int skip,       \
                                 int row_major, int device) {                                                         \
        return host_plu_solve<T>(n, m, B, A, Bm, X, ok, skip, row_major, device);                                     \
    }                                                                                                                 \
    int gmb_host_inv_##SUF(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int src_row_major, int dst_row_major,  \
                           int device) {                                                                              \
        return host_inv<T>(n, B, A, Ainv, ok, src_row_major, dst_row_major, device);                                  \
    }                                                                                                                 \
    int gmb_host_cholesky_solve_##SUF(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok,            \
                                      int row_major, int device) {                                                    \
        return host_cholesky_solve<T>(n, m, B, A, Bm, X, ok, row_major, device);                                      \
    }                                                                                                                 \
    int gmb_host_plu_decomp_##SUF(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity,                \
                                  uint8_t* degen, int row_major, int device) {                                        \
        return host_plu_decomp<T>(rows, cols, B, A, perm, parity, degen, row_major, device);                          \
    }                                                                                                                 \
    int gmb_host_plu_solve_multi_##SUF(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, int skip, \
                                       int row_major, const int* devices, int ndev) {                                 \
        if (n < 1 || m < 1 || (B > 0 && (!A || !Bm || !X))) return GMB_ERR_INVALID_ARG;                               \
        return run_multi_device(B, devices, ndev, [&](int64_t lo, int64_t hi, int dev) {                              \
            return host_plu_solve<T>(n, m, hi - lo, A + lo * n * n, Bm + lo * n * m, X + lo * n * m,                  \
                                     ok ? ok + lo : nullptr, skip, row_major, dev);                                   \
        });                                                                                                           \
    }                                                                                                                 \
    int gmb_host_inv_multi_##SUF(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int src_row_major,               \
                                 int dst_row_major, const int* devices, int ndev) {                                   \
        if (n < 1 || (B > 0 && (!A || !Ainv))) return GMB_ERR_INVALID_ARG;                                            \
        return run_multi_device(B, devices, ndev, [&](int64_t lo, int64_t hi, int dev) {                              \
            return host_inv<T>(n, hi - lo, A + lo * n * n, Ainv + lo * n * n, ok ? ok + lo : nullptr, src_row_major,  \
                               dst_row_major, dev);                                                                   \
        });                                                                                                           \
    }                                                                                                                 \
    int gmb_host_cholesky_solve_multi_##SUF(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok,      \
                                            int row_major, const int* devices, int ndev) {                            \
        if (n < 1 || m < 1 || (B > 0 && (!A || !Bm || !X))) return GMB_ERR_INVALID_ARG;                               \
        return run_multi_device(B, devices, ndev, [&](int64_t lo, int64_t hi, int dev) {                              \
            return host_cholesky_solve<T>(n, m, hi - lo, A + lo * n * n, Bm + lo * n * m, X + lo * n * m,             \
                                          ok ? ok + lo : nullptr, row_major, dev);                                    \
        });                                                                                                           \
    }                                                                                                                 \
    int gmb_host_plu_decomp_multi_##SUF(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity,          \
                                        uint8_t* degen, int row_major, const int* devices, int ndev) {                \
        if (rows < 1 || cols < 1 || (B > 0 && !A)) return GMB_ERR_INVALID_ARG;                                        \
        return run_multi_device(B, devices, ndev, [&](int64_t lo, int64_t hi, int dev) {                              \
            return host_plu_decomp<T>(rows, cols, hi - lo, A + lo * rows * cols, perm ? perm + lo * rows : nullptr,   \
                                      parity ? parity + lo : nullptr, degen ? degen + lo : nullptr, row_major, dev);  \
        });                                                                                                           \
    }

GMB_DEFINE(f32, float)
GMB_DEFINE(f64, double)

int gmb_checksum_bytes(const void* data, int64_t nbytes, int64_t word_offset, uint64_t* out, void* stream) {
    if (nbytes < 0 || (nbytes > 0 && (!data || !out))) return GMB_ERR_INVALID_ARG;
    return checksum_bytes(data, nbytes, word_offset, out, as_stream(stream));
}

}  // extern "C"
