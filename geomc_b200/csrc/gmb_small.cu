// gmb_small.cu -- N <= 4 kernels: one system per thread slot, everything register-resident.
//
// SoA-interleaved batches: a warp owns one tile (W = 32*V systems); lane l holds the V = 16/sizeof(T)
// consecutive systems l*V.. and moves every element plane with one 128-bit access.
// AoS batches (the reference's per-object storage): one system per thread; records of 32 bytes or
// more are written (and, for the inverse, read) through a warp-private shared-memory transpose so
// that every global access is a fully coalesced 128-bit access (gmb_io.cuh, aos_*_staged).
//
// No tensor cores: these are independent tiny systems (<= 0.6 flop/byte), bound by HBM; the
// kernels are plain FFMA/DFMA + IEEE division on the CUDA cores.
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"
#include "gmb_work.cuh"

namespace gmb {

// ============================================================ a6: fused linear_solve
template <typename T, int N, int M, bool RM, bool SOA>
__global__ void __launch_bounds__(kThreads)
k_plu_solve(const T* __restrict__ A, const T* Bm, T* X, uint8_t* __restrict__ ok, T* __restrict__ LU, int skip,
            int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qa[V][N * N], qb[V][N * M];
    load_elems<T, N * N, SOA>(A, unit, lane, B, qa);
    load_elems<T, N * M, SOA>(Bm, unit, lane, B, qb);
    uint8_t okv[V][1];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T a[N][N], x[N][M];
        unflatten<T, N, N, RM>(qa[v], a);
        unflatten<T, N, M, RM>(qb[v], x);
        uint8_t perm[N];
        bool parity;
        const int degenerate = plu_rows<T, N, N, M, false>(a, x, perm, parity);
        backsolve_lu_regs<T, N, M, /*FORWARD=*/false>(a, x, skip);
        const bool good = degenerate == 0;                      // LUDecomp.h:391-393
        okv[v][0] = good ? 1 : 0;
        T qx[N * M];
        flatten<T, N, M, RM>(x, qx);
#pragma unroll
        for (int e = 0; e < N * M; ++e) qb[v][e] = good ? qx[e] : qb[v][e];   // x untouched on failure
        if (LU != nullptr) flatten<T, N, N, RM>(a, qa[v]);
    }
    store_elems<T, N * M, SOA>(X, unit, lane, B, qb);
    if (LU != nullptr) store_elems<T, N * N, SOA>(LU, unit, lane, B, qa);
    store_bytes<V, 1>(ok, s0, B, okv);
}

// ============================================================ a1 / a2: decomp_plu / decomp_lup
template <typename T, int R, int C, bool RM, bool SOA, bool COLPIV>
__global__ void __launch_bounds__(kThreads)
k_decomp(T* __restrict__ A, uint8_t* __restrict__ perm_out, uint8_t* __restrict__ parity_out,
         uint8_t* __restrict__ degen_out, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    constexpr int NP = COLPIV ? C : R;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qa[V][R * C];
    load_elems<T, R * C, SOA>(A, unit, lane, B, qa);
    uint8_t pv[V][NP], par[V][1], deg[V][1];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T a[R][C];
        unflatten<T, R, C, RM>(qa[v], a);
        bool parity;
        int d;
        if constexpr (COLPIV) {
            d = lup_cols<T, R, C>(a, pv[v], parity);
        } else {
            T dummy[R][1];
            d = plu_rows<T, R, C, 0, true>(a, dummy, pv[v], parity);
        }
        par[v][0] = parity ? 1 : 0;
        deg[v][0] = (uint8_t)d;
        flatten<T, R, C, RM>(a, qa[v]);
    }
    store_elems<T, R * C, SOA>(A, unit, lane, B, qa);
    store_bytes<V, NP>(perm_out, s0, B, pv);
    store_bytes<V, 1>(parity_out, s0, B, par);
    store_bytes<V, 1>(degen_out, s0, B, deg);
}

// ============================================================ a3 / a4 / a5
// MODE 0: backsolve_lu (X in place); MODE 1: backsolve_plu (gather Bm rows through perm);
// MODE 2: apply permutation only.
template <typename T, int N, int M, bool RM, bool SOA, int MODE>
__global__ void __launch_bounds__(kThreads)
k_backsolve(const T* __restrict__ LU, const uint8_t* __restrict__ perm, const T* Bm, T* X, int skip, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qa[V][N * N], qb[V][N * M];
    if (MODE != 2) load_elems<T, N * N, SOA>(LU, unit, lane, B, qa);
    load_elems<T, N * M, SOA>(Bm, unit, lane, B, qb);
    uint8_t pv[V][N];
    if (MODE != 0) load_bytes<V, N>(perm, s0, B, pv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T a[N][N], x[N][M];
        unflatten<T, N, M, RM>(qb[v], x);
        if (MODE != 0) {
            // y[r] = b[perm[r]]  (LUDecomp.h:294 / :35-64): dynamic source row -> select chain
            T y[N][M];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    T t = x[0][i];
#pragma unroll
                    for (int k = 1; k < N; ++k) t = (pv[v][r] == k) ? x[k][i] : t;
                    y[r][i] = t;
                }
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int i = 0; i < M; ++i) x[r][i] = y[r][i];
        }
        if (MODE != 2) {
            unflatten<T, N, N, RM>(qa[v], a);
            backsolve_lu_regs<T, N, M, /*FORWARD=*/true>(a, x, skip);
        }
        flatten<T, N, M, RM>(x, qb[v]);
    }
    store_elems<T, N * M, SOA>(X, unit, lane, B, qb);
}

// ============================================================ a8: cholesky, and fused cholesky + solve
template <typename T, int N, int M, bool RM, bool SOA, bool SOLVE>
__global__ void __launch_bounds__(kThreads)
k_cholesky(const T* A, const T* __restrict__ Bm, T* X, uint8_t* __restrict__ ok, T* L_out, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qa[V][N * N], qb[V][SOLVE ? N * M : 1];
    load_elems<T, N * N, SOA>(A, unit, lane, B, qa);
    if constexpr (SOLVE) load_elems<T, N * M, SOA>(Bm, unit, lane, B, qb);
    uint8_t okv[V][1];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T a[N][N];
        unflatten<T, N, N, RM>(qa[v], a);
        okv[v][0] = cholesky_regs<T, N>(a) ? 1 : 0;
        if constexpr (SOLVE) {
            T x[N][M];
            unflatten<T, N, M, RM>(qb[v], x);
            cholesky_solve_regs<T, N, M>(a, x);
            flatten<T, N, M, RM>(x, qb[v]);
        }
        flatten<T, N, N, RM>(a, qa[v]);
    }
    if constexpr (SOLVE) store_elems<T, N * M, SOA>(X, unit, lane, B, qb);
    if (L_out != nullptr) store_elems<T, N * N, SOA>(L_out, unit, lane, B, qa);
    store_bytes<V, 1>(ok, s0, B, okv);
}

// ============================================================ a10 / a13: inv, N <= 4 (adjugate)
// The flat-array routine is layout-agnostic (inverse of the transpose is the transpose of the
// inverse); when source and destination layouts differ the result is transposed
// (MatrixInv.h:290-293, 311-314, 330-334).
template <typename T, int N, bool SOA, bool TRANSPOSE_OUT>
__global__ void __launch_bounds__(kThreads)
k_inv_small(const T* A, T* Ainv, uint8_t* __restrict__ ok, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qa[V][N * N], qo[V][N * N];
    load_elems<T, N * N, SOA, true>(A, unit, lane, B, qa);
    uint8_t okv[V][1];
    bool all_ok = true;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T out[N * N];
#pragma unroll
        for (int e = 0; e < N * N; ++e) out[e] = qa[v][e];       // placeholder when singular (see below)
        const bool good = inv_flat<T>(out, qa[v]);
        okv[v][0] = good ? 1 : 0;
        all_ok = all_ok && good;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) qo[v][N * r + c] = out[TRANSPOSE_OUT ? N * c + r : N * r + c];
    }
    if (__any_sync(0xffffffffu, !all_ok)) {                  // warp-uniform: the staged AoS load is cooperative
        // "out is unchanged" for singular systems (MatrixInv.h:45-46): re-read the destination.
        T qd[V][N * N];
        load_elems<T, N * N, SOA>(Ainv, unit, lane, B, qd);
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int e = 0; e < N * N; ++e) qo[v][e] = okv[v][0] ? qo[v][e] : qd[v][e];
    }
    store_elems<T, N * N, SOA>(Ainv, unit, lane, B, qo);
    store_bytes<V, 1>(ok, s0, B, okv);
}

// ============================================================ a7: linear_solve(Vec), N < 5
// m_inv = inv(bases) into a ROW_MAJOR matrix from a COL_MAJOR source (adjugate on the flat array,
// then transpose), X <- m_inv * X with `sum = 0; sum = fma(a, b, sum)` n ascending
// (LUDecomp.h:419-426, MatrixMult.h:54-64).  x holds m vectors of N (column-major N x m).
template <typename T, int N, int M, bool SOA>
__global__ void __launch_bounds__(kThreads)
k_solve_vec_small(const T* __restrict__ bases, const T* Bv, T* X, uint8_t* __restrict__ ok, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qa[V][N * N], qb[V][N * M];
    load_elems<T, N * N, SOA>(bases, unit, lane, B, qa);
    load_elems<T, N * M, SOA>(Bv, unit, lane, B, qb);
    uint8_t okv[V][1];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T inv[N * N];
#pragma unroll
        for (int e = 0; e < N * N; ++e) inv[e] = (T)0;
        const bool good = inv_flat<T>(inv, qa[v]);
        okv[v][0] = good ? 1 : 0;
        // transposed flat result: minv(r,k) row-major = inv[N*k + r]
        T out[N * M];
#pragma unroll
        for (int c = 0; c < M; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) {
                T sum = (T)0;
#pragma unroll
                for (int k = 0; k < N; ++k) sum = fma_rn(inv[N * k + r], qb[v][N * c + k], sum);
                out[N * c + r] = sum;
            }
#pragma unroll
        for (int e = 0; e < N * M; ++e) qb[v][e] = good ? out[e] : qb[v][e];
    }
    store_elems<T, N * M, SOA>(X, unit, lane, B, qb);
    store_bytes<V, 1>(ok, s0, B, okv);
}

// ============================================================ a10 / a7 for AoS batches that are only element-aligned
// An array of the reference's objects is 4- or 8-byte aligned, not 16: these two kernels run the SAME arithmetic
// (inv_flat, the fma product) with scalar loads and stores, one system per thread.  Slower than the vectorised kernels
// (every 32-byte sector is touched by several instructions) but bit-identical, so the alignment of a batch never
// changes its results.
template <typename T, int N, bool TRANSPOSE_OUT>
__global__ void __launch_bounds__(kThreads)
k_inv_small_elem(const T* A, T* Ainv, uint8_t* __restrict__ ok, int64_t B) {
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (s >= B) return;
    T qa[N * N], out[N * N];
#pragma unroll
    for (int e = 0; e < N * N; ++e) qa[e] = A[s * (N * N) + e];
#pragma unroll
    for (int e = 0; e < N * N; ++e) out[e] = qa[e];
    const bool good = inv_flat<T>(out, qa);
    if (ok != nullptr) ok[s] = good ? 1 : 0;
    if (!good) return;                                        // "out is unchanged" (MatrixInv.h:45-46)
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) Ainv[s * (N * N) + N * r + c] = out[TRANSPOSE_OUT ? N * c + r : N * r + c];
}

template <typename T, int N, int M>
__global__ void __launch_bounds__(kThreads)
k_solve_vec_small_elem(const T* __restrict__ bases, const T* Bv, T* X, uint8_t* __restrict__ ok, int64_t B) {
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (s >= B) return;
    T qa[N * N], qb[N * M], inv[N * N];
#pragma unroll
    for (int e = 0; e < N * N; ++e) qa[e] = bases[s * (N * N) + e];
#pragma unroll
    for (int e = 0; e < N * M; ++e) qb[e] = Bv[s * (N * M) + e];
#pragma unroll
    for (int e = 0; e < N * N; ++e) inv[e] = (T)0;
    const bool good = inv_flat<T>(inv, qa);
    if (ok != nullptr) ok[s] = good ? 1 : 0;
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) {
            T sum = (T)0;
#pragma unroll
            for (int k = 0; k < N; ++k) sum = fma_rn(inv[N * k + r], qb[N * c + k], sum);
            X[s * (N * M) + N * c + r] = good ? sum : qb[N * c + r];
        }
}

// ============================================================ (f)1: orthogonal, N <= 4
// Orthogonal.h:73-110: N = 2 left_perpendicular, N = 3 cross product (diff_of_products), N = 4
// rectangular decomp_plu of the 3 x 4 matrix + back-substitution from o[3] = 1; the zero vector when
// the decomposition reports a degenerate column.
template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kThreads)
k_orthogonal_small(const T* __restrict__ Vin, T* __restrict__ Out, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    T qv[V][(N - 1) * N], qo[V][N];
    load_elems<T, (N - 1) * N, SOA>(Vin, unit, lane, B, qv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if constexpr (N == 2) {
            qo[v][0] = -qv[v][1];
            qo[v][1] = qv[v][0];
        } else if constexpr (N == 3) {
            const T ax = qv[v][0], ay = qv[v][1], az = qv[v][2], bx = qv[v][3], by = qv[v][4], bz = qv[v][5];
            qo[v][0] = dop(ay, bz, az, by);                  // Vec3.h:207-219
            qo[v][1] = dop(az, bx, ax, bz);
            qo[v][2] = dop(ax, by, ay, bx);
        } else {
            T a[N - 1][N], dummy[N - 1][1];
            unflatten<T, N - 1, N, true>(qv[v], a);
            uint8_t perm[N - 1];
            bool parity;
            const int degenerate = plu_rows<T, N - 1, N, 0, false>(a, dummy, perm, parity);
            T o[N];
            o[N - 1] = (T)1;
#pragma unroll
            for (int r = N - 2; r >= 0; --r) {
                T acc = (T)0;
#pragma unroll
                for (int c = N - 1; c > r; --c) acc = fma_rn(-o[c], a[r][c], acc);     // :93
                o[r] = div_rn(acc, a[r][r]);
            }
#pragma unroll
            for (int e = 0; e < N; ++e) qo[v][e] = degenerate > 0 ? (T)0 : o[e];
        }
    }
    store_elems<T, N, SOA>(Out, unit, lane, B, qo);
}

// ============================================================ batch statistics of a solve (test / verify path)
// max_r |sum_c A(r,c) x(c) - b(r)| per system in double, then {sum, max, #ok == 0, #non-finite} over the batch.  Same
// tile access as the solve it verifies (128-bit plane loads in the container): reads the same 96 bytes per 4 x 4 float
// system and should run at the same HBM-bound rate (the element-wise generic kernel ran at 0.14 of the peak).
template <typename T, int N, bool RM, bool SOA>
__global__ void __launch_bounds__(kThreads)
k_residual_small(const T* __restrict__ A, const T* __restrict__ X, const T* __restrict__ Bm, const uint8_t* __restrict__ ok,
                 double* __restrict__ stats, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    double sum = 0.0, mx = 0.0, failed = 0.0, nonfinite = 0.0;
    if (Wk::locate(B, unit, lane, s0)) {
        T qa[V][N * N], qx[V][N], qb[V][N];
        load_elems<T, N * N, SOA>(A, unit, lane, B, qa);
        load_elems<T, N, SOA>(X, unit, lane, B, qx);
        load_elems<T, N, SOA>(Bm, unit, lane, B, qb);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (s0 + v >= B) continue;
            if (ok != nullptr && ok[s0 + v] == 0) {
                failed += 1.0;
                continue;
            }
            double res = 0.0;
            bool fin = true;
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double acc = -(double)qb[v][r];
#pragma unroll
                for (int c = 0; c < N; ++c) acc = fma((double)qa[v][RM ? N * r + c : N * c + r], (double)qx[v][c], acc);
                res = fmax(res, fabs(acc));
                fin = fin && isfinite(acc);
            }
            if (fin) {
                sum += res;
                mx = fmax(mx, res);
            } else {
                nonfinite += 1.0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        failed += __shfl_xor_sync(0xffffffffu, failed, o);
        nonfinite += __shfl_xor_sync(0xffffffffu, nonfinite, o);
    }
    __shared__ double part[kThreads / 32][4];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        part[w][0] = sum; part[w][1] = mx; part[w][2] = failed; part[w][3] = nonfinite;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < kThreads / 32; ++i) {
            sum += part[i][0]; mx = fmax(mx, part[i][1]); failed += part[i][2]; nonfinite += part[i][3];
        }
        if (sum != 0.0) atomicAdd(&stats[0], sum);
        if (mx != 0.0) atomicMax(reinterpret_cast<unsigned long long*>(&stats[1]), (unsigned long long)__double_as_longlong(mx));
        if (failed != 0.0) atomicAdd(&stats[2], failed);
        if (nonfinite != 0.0) atomicAdd(&stats[3], nonfinite);
    }
}

template <typename T>
int small_residual(int n, int64_t B, const T* A, const T* X, const T* Bm, const uint8_t* ok, double* stats, int layout,
                   int rm, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            dispatch_bool(rm != 0, [&](auto RM) {
                const int64_t grid = Work<T, SOA.value>::grid(B);
                if (grid > 0) k_residual_small<T, N.value, RM.value, SOA.value><<<(unsigned)grid, kThreads, 0, st>>>(A, X, Bm, ok, stats, B);
                rc = grid > 0 ? after_launch() : GMB_OK;
            });
        });
    });
    return rc;
}

template <typename T>
int small_orthogonal(int n, int64_t B, const T* V, T* out, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            if (Work<T, SOA.value>::grid(B) > 0)
                k_orthogonal_small<T, N.value, SOA.value><<<(unsigned)Work<T, SOA.value>::grid(B), kThreads, 0, st>>>(V, out, B);
            rc = after_launch();
        });
    });
    return rc;
}

// ============================================================ launchers (called from gmb_api.cu)
#define GMB_LAUNCH(kernel, grid, stream, ...)                                   \
    do {                                                                        \
        if ((grid) > 0) kernel<<<(unsigned)(grid), kThreads, 0, stream>>>(__VA_ARGS__); \
        rc = after_launch();                                                    \
    } while (0)

template <typename T>
int small_plu_solve(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* LU, int skip, int layout,
                    int row_major, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_int<1, 2, 3, 4>(m, [&](auto M) {
            dispatch_bool(row_major != 0, [&](auto RM) {
                dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                    GMB_LAUNCH((k_plu_solve<T, N.value, M.value, RM.value, SOA.value>), (Work<T, SOA.value>::grid(B)), st,
                               A, Bm, X, ok, LU, skip, B);
                });
            });
        });
    });
    return rc;
}

template <typename T>
int small_decomp(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout,
                 int row_major, bool colpiv, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(cols, [&](auto C) {
        dispatch_bool(rows == cols, [&](auto SQ) {
            constexpr int R = SQ.value ? C.value : C.value - 1;
            if (rows != R) return;
            if constexpr (R >= 1) {
                dispatch_bool(row_major != 0, [&](auto RM) {
                    dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                        dispatch_bool(colpiv, [&](auto CP) {
                            GMB_LAUNCH((k_decomp<T, R, C.value, RM.value, SOA.value, CP.value>),
                                       (Work<T, SOA.value>::grid(B)), st, A, perm, parity, degen, B);
                        });
                    });
                });
            }
        });
    });
    return rc;
}

template <typename T>
int small_backsolve(int mode, int n, int m, int64_t B, const T* LU, const uint8_t* perm, const T* Bm, T* X, int skip,
                    int layout, int row_major, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_int<1, 2, 3, 4>(m, [&](auto M) {
            dispatch_bool(row_major != 0, [&](auto RM) {
                dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                    dispatch_int<0, 1, 2>(mode, [&](auto MODE) {
                        GMB_LAUNCH((k_backsolve<T, N.value, M.value, RM.value, SOA.value, MODE.value>),
                                   (Work<T, SOA.value>::grid(B)), st, LU, perm, Bm, X, skip, B);
                    });
                });
            });
        });
    });
    return rc;
}

template <typename T>
int small_cholesky(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout,
                   int row_major, bool solve, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_bool(row_major != 0, [&](auto RM) {
            dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                if (solve) {
                    dispatch_int<1, 2, 3, 4>(m, [&](auto M) {
                        GMB_LAUNCH((k_cholesky<T, N.value, M.value, RM.value, SOA.value, true>),
                                   (Work<T, SOA.value>::grid(B)), st, A, Bm, X, ok, L_out, B);
                    });
                } else {
                    GMB_LAUNCH((k_cholesky<T, N.value, 1, RM.value, SOA.value, false>), (Work<T, SOA.value>::grid(B)),
                               st, A, Bm, X, ok, L_out, B);
                }
            });
        });
    });
    return rc;
}

template <typename T>
int small_inv(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            dispatch_bool((src_rm != 0) != (dst_rm != 0), [&](auto TR) {
                GMB_LAUNCH((k_inv_small<T, N.value, SOA.value, TR.value>), (Work<T, SOA.value>::grid(B)), st, A, Ainv,
                           ok, B);
            });
        });
    });
    return rc;
}

template <typename T>
int small_inv_elem(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int src_rm, int dst_rm, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_bool((src_rm != 0) != (dst_rm != 0), [&](auto TR) {
            GMB_LAUNCH((k_inv_small_elem<T, N.value, TR.value>), (ceil_div(B, kThreads)), st, A, Ainv, ok, B);
        });
    });
    return rc;
}

template <typename T>
int small_solve_vec_elem(int n, int m, int64_t B, const T* bases, const T* Bv, T* X, uint8_t* ok, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_int<1, 2, 3, 4>(m, [&](auto M) {
            GMB_LAUNCH((k_solve_vec_small_elem<T, N.value, M.value>), (ceil_div(B, kThreads)), st, bases, Bv, X, ok, B);
        });
    });
    return rc;
}

template <typename T>
int small_solve_vec(int n, int m, int64_t B, const T* bases, const T* Bv, T* X, uint8_t* ok, int layout,
                    cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4>(n, [&](auto N) {
        dispatch_int<1, 2, 3, 4>(m, [&](auto M) {
            dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                GMB_LAUNCH((k_solve_vec_small<T, N.value, M.value, SOA.value>), (Work<T, SOA.value>::grid(B)), st, bases,
                           Bv, X, ok, B);
            });
        });
    });
    return rc;
}

#define INSTANTIATE(T)                                                                                                 \
    template int small_plu_solve<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int,            \
                                    cudaStream_t);                                                                     \
    template int small_decomp<T>(int, int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, bool, cudaStream_t);   \
    template int small_backsolve<T>(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int,     \
                                    cudaStream_t);                                                                     \
    template int small_cholesky<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool,            \
                                   cudaStream_t);                                                                      \
    template int small_inv<T>(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);                      \
    template int small_solve_vec<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, int, cudaStream_t);          \
    template int small_inv_elem<T>(int, int64_t, const T*, T*, uint8_t*, int, int, cudaStream_t);                      \
    template int small_solve_vec_elem<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, cudaStream_t);           \
    template int small_orthogonal<T>(int, int64_t, const T*, T*, int, cudaStream_t);                                   \
    template int small_residual<T>(int, int64_t, const T*, const T*, const T*, const uint8_t*, double*, int, int,      \
                                   cudaStream_t);

INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace gmb
