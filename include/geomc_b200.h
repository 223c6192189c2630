/* geomc_b200.h -- C ABI of the B200-native batched small-matrix factor/solve engine.
 *
 * Drop-in boundary for ONE hot path of trbabb/geomc: the geomc/linalg factor / solve / inverse
 * functions, applied to a BATCH of B independent systems per call instead of one object per
 * call.  The reference has no FFI layer of its own (it is a header-only C++ template library),
 * so each entry point below names the reference template it replaces (file:line under
 * geomc/linalg/ of the reference) and keeps that function's argument meaning, return meaning
 * and error behaviour per system.  include/geomc_b200/batch.hpp wraps these in C++ templates
 * with the reference's names (geom::batch::linear_solve, ...); INTEGRATION.md shows the binding
 * a geomc maintainer would add.
 *
 * Conventions
 *   - T in {float (suffix _f32), double (suffix _f64)}.  n = matrix dimension, m = number of
 *     right-hand-side columns, B = number of systems.  2 <= n <= 16 (register/sub-warp kernels).
 *   - Every pointer in the gmb_* (non gmb_host_*) functions is a DEVICE pointer; `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous.
 *   - `layout`: GMB_LAYOUT_AOS = the reference's own per-object storage, system s / element e at
 *     p[s*E + e] (SimpleMatrix.h:145, VecBase.h:112; records must be packed: E = n*n or n*m).
 *     GMB_LAYOUT_SOA = the SoA-interleaved batch container: tiles of W = 512/sizeof(T) systems
 *     (128 float / 64 double); p[(s/W * E + e) * W + s%W]; allocate gmb_soa_elems(B, E, sizeof(T))
 *     elements (whole tiles).  gmb_aos_to_soa_* / gmb_soa_to_aos_* convert.
 *   - `row_major`: how the E = n*n (or n*m) elements of ONE system map to (r, c), exactly as the
 *     reference's RowMajor template flag (MatrixGlue.h:28-58): 1 -> e = cols*r + c, 0 -> e = rows*c + r.
 *   - Per-system results mirror the reference's return values: `ok` (uint8, 1 = true), `parity`
 *     (uint8), `degenerate` (uint8 count), `perm` (uint8[B][n], AoS in every layout: the reference's
 *     index_t reorder[] narrowed; row i of P*M is row perm[i] of M, LUDecomp.h:182-183).  Any of
 *     these output pointers may be NULL.
 *   - Alignment: AoS batches that start on a 16-byte boundary take the vectorised kernels (32-byte alignment -- any
 *     cudaMalloc'ed buffer -- additionally enables 256-bit loads); other AoS batches are served by element-wise
 *     kernels with identical results; a SoA container must be 16-byte aligned.
 *   - Return value: 0 = success; < 0 = GMB_ERR_*; > 0 = a cudaError_t from the launch.
 *     There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef GEOMC_B200_H
#define GEOMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMB_LAYOUT_AOS 0
#define GMB_LAYOUT_SOA 1

#define GMB_OK 0
#define GMB_ERR_INVALID_ARG (-1)
#define GMB_ERR_UNSUPPORTED (-2)   /* n or m outside the compiled template range */
#define GMB_ERR_NO_DEVICE (-3)
#define GMB_ERR_ALLOC (-4)

#define GMB_MAX_N 16
#define GMB_MAX_M 4                /* right-hand-side columns per call (n for the inverse kernels) */

/* ---- library / container helpers --------------------------------------------------------- */
const char* gmb_version(void);
const char* gmb_error_string(int code);
int gmb_device_count(void);
/* Systems per SoA tile for an element size (128 for 4-byte, 64 for 8-byte elements). */
int gmb_soa_tile_width(int elem_size);
/* Elements to allocate for a SoA batch of B systems with E elements each (whole tiles). */
int64_t gmb_soa_elems(int64_t B, int E, int elem_size);

/* AoS <-> SoA-interleaved conversion of a batch of B records of E elements (coalesced both
 * sides through a shared-memory transpose).  Replaces: per-object storage, SimpleMatrix.h:145. */
int gmb_aos_to_soa_f32(int E, int64_t B, const float* aos, float* soa, void* stream);
int gmb_aos_to_soa_f64(int E, int64_t B, const double* aos, double* soa, void* stream);
int gmb_soa_to_aos_f32(int E, int64_t B, const float* soa, float* aos, void* stream);
int gmb_soa_to_aos_f64(int E, int64_t B, const double* soa, double* aos, void* stream);

/* ---- a1: decomp_plu  (LUDecomp.h:188-250) -------------------------------------------------
 * In-place Doolittle LU with partial ROW pivoting of a rows x cols matrix: LU = P M.
 * rows == cols, or rows == cols - 1 (the rectangular case Orthogonal.h:83,161 uses).
 * perm: uint8[B][rows]; parity: 1 if an odd number of row exchanges; degenerate: number of
 * zero-pivot columns (the reference's return value). */
int gmb_plu_decomp_f32(int rows, int cols, int64_t B, float* A_inout, uint8_t* perm, uint8_t* parity,
                       uint8_t* degenerate, int layout, int row_major, void* stream);
int gmb_plu_decomp_f64(int rows, int cols, int64_t B, double* A_inout, uint8_t* perm, uint8_t* parity,
                       uint8_t* degenerate, int layout, int row_major, void* stream);

/* ---- a2: decomp_lup  (LUDecomp.h:102-165) -- COLUMN pivoting, LU = M P; perm: uint8[B][cols]. */
int gmb_lup_decomp_f32(int rows, int cols, int64_t B, float* A_inout, uint8_t* perm, uint8_t* parity,
                       uint8_t* degenerate, int layout, int row_major, void* stream);
int gmb_lup_decomp_f64(int rows, int cols, int64_t B, double* A_inout, uint8_t* perm, uint8_t* parity,
                       uint8_t* degenerate, int layout, int row_major, void* stream);

/* ---- a3: backsolve_lu  (LUDecomp.h:332-366) -----------------------------------------------
 * Solves L U X = X_in in place for the n x m matrix X (already row-permuted).  Rows < skip are
 * left as forward-substituted values ("nonsense" by the reference's contract). */
int gmb_lu_backsolve_f32(int n, int m, int64_t B, const float* LU, float* X_inout, int skip, int layout,
                         int row_major, void* stream);
int gmb_lu_backsolve_f64(int n, int m, int64_t B, const double* LU, double* X_inout, int skip, int layout,
                         int row_major, void* stream);

/* ---- a4: backsolve_plu  (LUDecomp.h:271-312) -- Y(r,:) = Bm(perm[r],:), then as a3. */
int gmb_plu_backsolve_f32(int n, int m, int64_t B, const float* PLU, const uint8_t* perm, float* X_out,
                          const float* Bm, int skip, int layout, int row_major, void* stream);
int gmb_plu_backsolve_f64(int n, int m, int64_t B, const double* PLU, const uint8_t* perm, double* X_out,
                          const double* Bm, int skip, int layout, int row_major, void* stream);

/* ---- a5: detail::destructive_apply_permutation  (LUDecomp.h:35-64) -------------------------
 * A <- P A for the n x m matrix A (row i <- row perm[i]).  perm is NOT clobbered here. */
int gmb_apply_permutation_f32(int n, int m, int64_t B, const uint8_t* perm, float* A_inout, int layout,
                              int row_major, void* stream);
int gmb_apply_permutation_f64(int n, int m, int64_t B, const uint8_t* perm, double* A_inout, int layout,
                              int row_major, void* stream);

/* ---- a6: linear_solve, pointer form  (LUDecomp.h:388-399) -- THE METRIC KERNEL -------------
 * Fused decomp_plu + permutation + backsolve_lu: solves A X = Bm.  ok[s] = 0 iff the reference
 * returns false (degenerate_ct > 0); then X[s] = Bm[s] ("x untouched").  A is not modified
 * (the reference destroys `a`; pass LU_out != NULL to receive what it would hold: the packed
 * LU factors, only partially reduced for degenerate systems).  X may alias Bm. */
int gmb_plu_solve_f32(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok,
                      float* LU_out, int skip, int layout, int row_major, void* stream);
int gmb_plu_solve_f64(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok,
                      double* LU_out, int skip, int layout, int row_major, void* stream);

/* ---- a7: linear_solve, Vec form  (LUDecomp.h:417-441) --------------------------------------
 * bases = n column vectors (column-major n x n), x = m vectors of n.  n < 5: inverse-then-
 * multiply (inv + mul, :419-426; bases unchanged).  n >= 5: column-major PLU + backsolve
 * (bases would hold the LU in the reference; here bases is const and LU is discarded). */
int gmb_linear_solve_vec_f32(int n, int m, int64_t B, const float* bases, const float* Bv, float* X,
                             uint8_t* ok, int skip, int layout, void* stream);
int gmb_linear_solve_vec_f64(int n, int m, int64_t B, const double* bases, const double* Bv, double* X,
                             uint8_t* ok, int skip, int layout, void* stream);

/* ---- a8/a9: cholesky  (Cholesky.h:37-58 / 83-88) -------------------------------------------
 * In-place lower Cholesky factor (strict upper triangle zeroed); ok = 0 if some pivot <= 0
 * (the factorization still runs to the end, like the reference). */
int gmb_cholesky_f32(int n, int64_t B, float* A_inout, uint8_t* ok, int layout, int row_major, void* stream);
int gmb_cholesky_f64(int n, int64_t B, double* A_inout, uint8_t* ok, int layout, int row_major, void* stream);

/* ---- NEW (no reference counterpart): fused Cholesky factor + solve A X = Bm ---------------
 * L as a8, then L y = b, L^T x = y (one fma per term, one IEEE division per row).  X is written
 * for every system (NaNs where ok = 0).  L_out may be NULL. */
int gmb_cholesky_solve_f32(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok,
                           float* L_out, int layout, int row_major, void* stream);
int gmb_cholesky_solve_f64(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok,
                           double* L_out, int layout, int row_major, void* stream);

/* ---- a10/a12/a13: inv  (Matrix.h:717-735 -> MatrixInv.h:49-165, 219-231, 241-337) ---------
 * n <= 4: adjugate with the reference's diff_of_products call tree; n >= 5: row-major PLU in the
 * destination's layout + back-substitution against the permuted identity.  ok = 0 -> Ainv[s] is
 * left untouched.  Ainv may alias A for n <= 4 only (as in the reference). */
int gmb_inv_f32(int n, int64_t B, const float* A, float* Ainv, uint8_t* ok, int layout, int src_row_major,
                int dst_row_major, void* stream);
int gmb_inv_f64(int n, int64_t B, const double* A, double* Ainv, uint8_t* ok, int layout, int src_row_major,
                int dst_row_major, void* stream);

/* ---- "next" row (f)1: orthogonal / nullspace  (Orthogonal.h:73-110 / 136-178) ---------------
 * gmb_orthogonal: V holds n-1 vectors of n per system (row-major (n-1) x n); out receives the vector
 * orthogonal to all of them (n = 2: left_perpendicular, n = 3: cross product, n >= 4: rectangular
 * decomp_plu + back-substitution from o[n-1] = 1; the zero vector if a degenerate column is found).
 * gmb_nullspace: bases holds nb vectors of N (1 <= nb < N), null_basis receives N - nb vectors of N
 * whose dot products with the bases vanish; ok = 0 (and a zero-filled null basis) when the bases are
 * linearly dependent.  nb == N - 1 delegates to orthogonal, as the reference does. */
int gmb_orthogonal_f32(int n, int64_t B, const float* V, float* out, int layout, void* stream);
int gmb_orthogonal_f64(int n, int64_t B, const double* V, double* out, int layout, void* stream);
int gmb_nullspace_f32(int N, int nb, int64_t B, const float* bases, float* null_basis, uint8_t* ok, int layout,
                      void* stream);
int gmb_nullspace_f64(int N, int nb, int64_t B, const double* bases, double* null_basis, uint8_t* ok, int layout,
                      void* stream);

/* ---- "next" row (f)2: batched simplex queries, the in-library consumers of linear_solve --------
 * (shape/Simplex.h).  dim = N = 2..7.  Vertices are vertex-major, as in Simplex::pts (Simplex.h:92).
 * AoS strides are in elements, 0 = packed: `vert_stride` between consecutive simplexes (an array of
 * geom::Simplex<T,N> objects passes sizeof(Simplex<T,N>)/sizeof(T)), `ray_stride` between consecutive
 * origins / directions (an array of geom::Ray<T,N> passes 2N with direction = origin + N).  With
 * GMB_LAYOUT_SOA every array is a packed interleaved container and the strides must be 0.
 * nverts (nullable, one byte per simplex) = Simplex::n; NULL means every simplex has N + 1 vertices.
 *
 * gmb_simplex_contains: Simplex<T,N>::contains(p), Simplex.h:450-475.  verts = N+1 vertices of N;
 *   inside[s] = 1 iff p_s is on or inside simplex s (0 for n < N+1 and for degenerate simplexes).
 * gmb_trace_simplex: trace_simplex<T,N>(verts, ray, uv, s), Simplex.h:647-675.  verts = N vertices
 *   of N (a triangle in 3D).  hit[s]; s[s] and uv[s] (N-1 surface coordinates, nullable) are written
 *   only where hit.  As in the reference, uv == NULL solves with skip = N-1.  The reference stores
 *   x[N] -- one past the end of its Vec<T,N>, undefined -- as the ray parameter (:675); this library
 *   stores the coefficient it means, x[N-1].
 * gmb_simplex_intersect: Simplex<T,N>::intersect(Ray), Simplex.h:251-296.  verts = N+1 vertices of N;
 *   interval[s] = {lo, hi} of the Rect<T,1> the reference returns: the parameter range inside a full
 *   simplex; (s, s) for n == N; (inf, -inf) = Rect::empty or (max, lowest) = Rect() when there is no hit. */
int gmb_simplex_contains_f32(int dim, int64_t B, const float* verts, int64_t vert_stride, const uint8_t* nverts,
                             const float* p, uint8_t* inside, int layout, void* stream);
int gmb_simplex_contains_f64(int dim, int64_t B, const double* verts, int64_t vert_stride, const uint8_t* nverts,
                             const double* p, uint8_t* inside, int layout, void* stream);
int gmb_trace_simplex_f32(int dim, int64_t B, const float* verts, int64_t vert_stride, const float* origin,
                          const float* direction, int64_t ray_stride, uint8_t* hit, float* s, float* uv, int layout,
                          void* stream);
int gmb_trace_simplex_f64(int dim, int64_t B, const double* verts, int64_t vert_stride, const double* origin,
                          const double* direction, int64_t ray_stride, uint8_t* hit, double* s, double* uv, int layout,
                          void* stream);
int gmb_simplex_intersect_f32(int dim, int64_t B, const float* verts, int64_t vert_stride, const uint8_t* nverts,
                              const float* origin, const float* direction, int64_t ray_stride, float* interval,
                              int layout, void* stream);
int gmb_simplex_intersect_f64(int dim, int64_t B, const double* verts, int64_t vert_stride, const uint8_t* nverts,
                              const double* origin, const double* direction, int64_t ray_stride, double* interval,
                              int layout, void* stream);

/* ---- "next" row (f)4: batched mul / mul_acc  (Matrix.h:386-421 / 433-456 -> MatrixMult.h:54-80) --------
 * D (r x c) = A (r x k) * Bm (k x c) per system; accumulate != 0 -> D += A * Bm (mul_acc).  Each D(i,j) is the
 * running sum `sum = 0 (or D(i,j)); sum = fma(A(i,n), Bm(n,j), sum)` for n ascending, as the reference's build
 * evaluates `sum += a(r,n) * b(n,c)`.  c == 1 is matrix * Vec (MatrixMult.h:181-190), r == 1 is Vec * matrix
 * (:239-248).  Each operand has its own row_major flag (MatrixGlue.h:28-58).  1 <= r, k, c <= 16.
 * D may alias A or Bm (the reference multiplies into a temporary then, Matrix.h:409-417). */
int gmb_mul_f32(int r, int k, int c, int64_t B, const float* A, const float* Bm, float* D, int accumulate, int layout,
                int a_row_major, int b_row_major, int d_row_major, void* stream);
int gmb_mul_f64(int r, int k, int c, int64_t B, const double* A, const double* Bm, double* D, int accumulate, int layout,
                int a_row_major, int b_row_major, int d_row_major, void* stream);

/* ---- "next" row (f)3: AffineTransform<T,N>  (linalg/AffineTransform.h), dim = N = 2..7 -------------------
 * A transform record is PACKED: K*K elements of `mat` then K*K of `inv`, K = N + 1, both row-major
 * (AffineTransform.h:91-93).  AoS = consecutive records of 2*K*K elements (note that the reference's objects
 * can carry alignment padding -- sizeof(AffineTransform<float,2>) is 80, not 72 -- include/geomc_b200/batch.hpp
 * packs and unpacks); SoA = an interleaved container with E = 2*K*K.
 *
 * gmb_xf_from_matrix: transformation(mat), :862-884.  M = N x N per system (row_major flag as elsewhere).
 *   mat = M embedded in the identity, inv = inv(M) embedded likewise.  The reference ignores inv()'s result
 *   (:873), so a singular M yields inv = identity; ok[s] (nullable) reports inv()'s return value.
 * gmb_xf_translation / gmb_xf_scale: translation(tx) :814-821 (inv = -tx), scale(sx) :838-845 (inv = 1 / sx[i]).
 * gmb_xf_compose: inverse == 0: xf1 * xf2 (:142-147 -> apply :300-305: mat = xf1.mat * xf2.mat,
 *   inv = xf2.inv * xf1.inv); inverse != 0: xf1 / xf2 (:154-159 -> apply_inverse :312-317: mat = xf2.inv * xf1.mat,
 *   inv = xf1.inv * xf2.mat).  `out` must not alias an input.
 * gmb_xf_apply: p and out hold B vectors of N.  mode: 0 apply (:224), 1 apply_inverse (:262), 2 apply_direction
 *   (:233), 3 apply_inverse_direction (:271), 4 apply_normal (:248), 5 apply_inverse_normal (:284).
 *   xf_count == B: one transform per point (layout as the other arguments); xf_count == 1: ONE packed record
 *   (2*K*K contiguous elements, whatever `layout` says -- also when B == 1) applied to the whole point array. */
#define GMB_XF_APPLY 0
#define GMB_XF_APPLY_INVERSE 1
#define GMB_XF_APPLY_DIRECTION 2
#define GMB_XF_APPLY_INVERSE_DIRECTION 3
#define GMB_XF_APPLY_NORMAL 4
#define GMB_XF_APPLY_INVERSE_NORMAL 5
int gmb_xf_from_matrix_f32(int dim, int64_t B, const float* M, float* xf_out, uint8_t* ok, int layout, int row_major,
                           void* stream);
int gmb_xf_from_matrix_f64(int dim, int64_t B, const double* M, double* xf_out, uint8_t* ok, int layout, int row_major,
                           void* stream);
int gmb_xf_translation_f32(int dim, int64_t B, const float* t, float* xf_out, int layout, void* stream);
int gmb_xf_translation_f64(int dim, int64_t B, const double* t, double* xf_out, int layout, void* stream);
int gmb_xf_scale_f32(int dim, int64_t B, const float* sx, float* xf_out, int layout, void* stream);
int gmb_xf_scale_f64(int dim, int64_t B, const double* sx, double* xf_out, int layout, void* stream);
int gmb_xf_compose_f32(int dim, int64_t B, const float* xf1, const float* xf2, float* out, int inverse, int layout,
                       void* stream);
int gmb_xf_compose_f64(int dim, int64_t B, const double* xf1, const double* xf2, double* out, int inverse, int layout,
                       void* stream);
int gmb_xf_apply_f32(int dim, int mode, int64_t B, const float* xf, int64_t xf_count, const float* p, float* out,
                     int layout, void* stream);
int gmb_xf_apply_f64(int dim, int mode, int64_t B, const double* xf, int64_t xf_count, const double* p, double* out,
                     int layout, void* stream);

/* ---- batch statistics (what the reference's tests assert, reduced on the device) ------------
 * stats[0] = sum over systems of max_i |(A x - b)_i| (only systems with ok != 0),
 * stats[1] = max of the same, stats[2] = number of systems with ok == 0,
 * stats[3] = number of systems whose residual is not finite.   (double[4], device memory;
 * accumulated into -- zero it first.)  m = 1. */
int gmb_solve_residual_f32(int n, int64_t B, const float* A, const float* X, const float* Bm, const uint8_t* ok,
                           double* stats4, int layout, int row_major, void* stream);
int gmb_solve_residual_f64(int n, int64_t B, const double* A, const double* X, const double* Bm, const uint8_t* ok,
                           double* stats4, int layout, int row_major, void* stream);
/* Additive 64-bit checksum of a byte buffer: sum (mod 2^64) over 8-byte words w of
 * mix64(word + K * (word_offset + w + 1)).  Shards of one buffer hashed with their global word
 * offsets add up to the checksum of the whole ("checksum of checksums" across GPUs; a shard must
 * start on an 8-byte boundary of the whole).  *out (device) is accumulated into. */
int gmb_checksum_bytes(const void* data, int64_t nbytes, int64_t word_offset, uint64_t* out, void* stream);

/* ---- host-buffer entry points (the call a host application makes) ---------------------------
 * All pointers are HOST pointers to AoS batches (the reference's own arrays of objects).  The
 * call copies host->device, runs the kernel and copies the results back, pipelined in chunks
 * over several streams on `device`; it returns when the results are in the host buffers.
 * Pinned buffers (gmb_host_alloc, or cudaHostRegister'ed memory) give full PCIe bandwidth. */
void* gmb_host_alloc(size_t bytes);
/* Write-combined pinned memory for INPUT batches the host only writes (fills) and the GPU only reads: the DMA engine
 * does not have to snoop the CPU caches.  Reading it back on the CPU is slow; free with gmb_host_free. */
void* gmb_host_alloc_wc(size_t bytes);
/* NUMA placement on boxes with several host memory nodes.  gmb_bind_thread_near_device pins the CALLING thread to the CPUs the
 * kernel lists as local to `device` (sysfs local_cpulist of its PCI function): memory it allocates and touches afterwards
 * lives on that node, and the copies it issues start there.  Returns 1 when bound, 0 when the topology is not readable
 * (nothing changed), < 0 for a bad device.  gmb_host_alloc_near = bind, pinned allocation, first touch, restore the
 * caller's affinity.  The worker threads of the gmb_host_*_multi_* entry points bind themselves the same way. */
int gmb_bind_thread_near_device(int device);
void* gmb_host_alloc_near(size_t bytes, int device);
void gmb_host_free(void* p);
/* Pin / unpin memory the application already owns (an array of SimpleMatrix objects, a std::vector's storage) so that
 * the host entry points reach full PCIe bandwidth on it without a copy into a pinned staging buffer.  Returns a
 * cudaError_t (> 0) when the range cannot be registered; the host entry points still work on pageable memory. */
int gmb_host_register(void* p, size_t bytes);
int gmb_host_unregister(void* p);
int gmb_host_plu_solve_f32(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok,
                           int skip, int row_major, int device);
int gmb_host_plu_solve_f64(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok,
                           int skip, int row_major, int device);
int gmb_host_inv_f32(int n, int64_t B, const float* A, float* Ainv, uint8_t* ok, int src_row_major,
                     int dst_row_major, int device);
int gmb_host_inv_f64(int n, int64_t B, const double* A, double* Ainv, uint8_t* ok, int src_row_major,
                     int dst_row_major, int device);
int gmb_host_cholesky_solve_f64(int n, int m, int64_t B, const double* A, const double* Bm, double* X,
                                uint8_t* ok, int row_major, int device);
int gmb_host_cholesky_solve_f32(int n, int m, int64_t B, const float* A, const float* Bm, float* X,
                                uint8_t* ok, int row_major, int device);
int gmb_host_plu_decomp_f64(int rows, int cols, int64_t B, double* A_inout, uint8_t* perm, uint8_t* parity,
                            uint8_t* degenerate, int row_major, int device);
int gmb_host_plu_decomp_f32(int rows, int cols, int64_t B, float* A_inout, uint8_t* perm, uint8_t* parity,
                            uint8_t* degenerate, int row_major, int device);
/* ---- several GPUs of one box -------------------------------------------------------------------
 * The systems of a batch are independent, so a batch is sharded CONTIGUOUSLY over devices[0..ndev): device k serves
 * systems [begin, end) = gmb_shard_range(B, k, ndev, 1024) of every host array through its own host pipeline, driven
 * by its own host thread.  There is no collective and no peer traffic on the data path; the call returns when every
 * shard's results are in the host buffers.  ndev = 1 is the single-device call.  Devices must be distinct. */
int gmb_shard_range(int64_t B, int rank, int world, int64_t align, int64_t* begin, int64_t* end);
int gmb_host_plu_solve_multi_f32(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok,
                                 int skip, int row_major, const int* devices, int ndev);
int gmb_host_plu_solve_multi_f64(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok,
                                 int skip, int row_major, const int* devices, int ndev);
int gmb_host_inv_multi_f32(int n, int64_t B, const float* A, float* Ainv, uint8_t* ok, int src_row_major,
                           int dst_row_major, const int* devices, int ndev);
int gmb_host_inv_multi_f64(int n, int64_t B, const double* A, double* Ainv, uint8_t* ok, int src_row_major,
                           int dst_row_major, const int* devices, int ndev);
int gmb_host_cholesky_solve_multi_f32(int n, int m, int64_t B, const float* A, const float* Bm, float* X,
                                      uint8_t* ok, int row_major, const int* devices, int ndev);
int gmb_host_cholesky_solve_multi_f64(int n, int m, int64_t B, const double* A, const double* Bm, double* X,
                                      uint8_t* ok, int row_major, const int* devices, int ndev);
int gmb_host_plu_decomp_multi_f32(int rows, int cols, int64_t B, float* A_inout, uint8_t* perm, uint8_t* parity,
                                  uint8_t* degenerate, int row_major, const int* devices, int ndev);
int gmb_host_plu_decomp_multi_f64(int rows, int cols, int64_t B, double* A_inout, uint8_t* perm, uint8_t* parity,
                                  uint8_t* degenerate, int row_major, const int* devices, int ndev);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t gmb_launch_count(void);
/* Dispatch options for tests and tuning (the library never reads the environment).  value = -1 restores the measured
 * per-(type, n) table.  "rowplu_min_n" / "rowdecomp_min_n": smallest n served by the row-distributed fused solve /
 * decomposition (17 = never); "rowinv": 0 / 1 = column- / row-distributed inverse for every n >= 5; "rowfast": 0 / 1 =
 * never / always use the speculative fused solve (n >= 9, one right-hand side, no LU output) with its exact re-solve.
 * "invd": 0 = no deferred-rider inverse; "lanechol": 0 / 1 = never / always the one-thread Cholesky for n = 5..8; "lanebs": 0 = the
 * backsolves never take the one-lane kernel with 1..4 right-hand sides (float n <= 14, double n <= 9); "subprefetch": L2 prefetch
 * distance of the one-lane sub-warp kernels in CTAs (0 = off, > 0 = that distance for every size, -1 = the measured rule).
 * Every choice computes the same bits; only the time differs.  Returns GMB_ERR_INVALID_ARG for an unknown name. */
int gmb_set_option(const char* name, int value);

/* ---- device memory and streams for hosts that do not link the CUDA runtime themselves --------
 * (include/geomc_b200/batch.hpp builds with a plain C++ compiler on top of these). */
int gmb_set_device(int device);
int gmb_get_device(int* device);
void* gmb_device_alloc(size_t bytes);                 /* NULL on failure */
void gmb_device_free(void* p);
int gmb_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes, void* stream);
int gmb_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes, void* stream);
int gmb_memset_device(void* dst_device, int value, size_t bytes, void* stream);
int gmb_stream_synchronize(void* stream);             /* NULL = the default stream */

#ifdef __cplusplus
}
#endif
#endif /* GEOMC_B200_H */
