// gmb_simplex.cu -- SURVEY.md 8(f)2: the in-library consumers of linear_solve, batched and fused.
//
//   Simplex<T,N>::contains(p)        shape/Simplex.h:450-475   -> k_simplex_contains
//   trace_simplex<T,N>(verts, ray)   shape/Simplex.h:647-675   -> k_trace_simplex
//   Simplex<T,N>::intersect(ray)     shape/Simplex.h:251-296   -> k_simplex_intersect
//
// One thread per simplex: matrix assembly, the reference's linear_solve (pointer form, column-major,
// skip = 1 for contains; Vec form for the ray queries: inverse-then-multiply below 5 unknowns,
// LUDecomp.h:419-426, PLU above) and the barycentric / interval test all happen in registers, so a
// query costs its vertices + point/ray in and one byte / one interval out.  N = 2..7 (the
// dimensions regression/simplex.cpp exercises), K = N + 1 <= 8 unknowns.
//
// Layouts: AoS = vertex-major records (Simplex::pts, :92) `vert_stride` elements apart, which lets
// an array of Simplex<T,N> objects be passed as is; SoA = the package's interleaved tiles
// (gmb_io.cuh), element planes read with coalesced accesses.
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"
#include "gmb_work.cuh"

#include <limits>

namespace gmb {

constexpr int kSimplexThreads = 128;

// Vec-form linear_solve, LUDecomp.h:417-441, on K basis vectors `bases[K*i + d]` (= column-major
// K x K) and M right-hand sides x[j][d].  K < 5: inverse, then X <- inv * X (sum from 0, k ascending,
// MatrixMult.h:54-64); skip is ignored on that path, as in the reference.  K >= 5: column-major
// decomp_plu + permutation + backsolve_lu with `skip`.  x is left untouched on failure.
template <typename T, int K, int M>
__device__ __forceinline__ bool solve_vec_regs(const T (&bases)[K * K], T (&x)[M][K], int skip) {
    if constexpr (K < 5) {
        T inv[K * K];
#pragma unroll
        for (int e = 0; e < K * K; ++e) inv[e] = (T)0;
        if (!inv_flat<T>(inv, bases)) return false;          // transposed flat result: minv(r,k) = inv[K*k + r]
#pragma unroll
        for (int j = 0; j < M; ++j) {
            T out[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                T sum = (T)0;
#pragma unroll
                for (int k = 0; k < K; ++k) sum = fma_rn(inv[K * k + r], x[j][k], sum);
                out[r] = sum;
            }
#pragma unroll
            for (int r = 0; r < K; ++r) x[j][r] = out[r];
        }
        return true;
    } else {
        T a[K][K], xr[K][M];
#pragma unroll
        for (int r = 0; r < K; ++r) {
#pragma unroll
            for (int c = 0; c < K; ++c) a[r][c] = bases[K * c + r];
#pragma unroll
            for (int j = 0; j < M; ++j) xr[r][j] = x[j][r];
        }
        uint8_t perm[K];
        bool parity;
        if (plu_rows<T, K, K, M, false>(a, xr, perm, parity) > 0) return false;
        backsolve_lu_regs<T, K, M, /*FORWARD=*/false>(a, xr, skip);
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) x[j][r] = xr[r][j];
        return true;
    }
}

// ------------------------------------------------------------------------------ contains
template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kSimplexThreads)
k_simplex_contains(const T* __restrict__ verts, int64_t vstride, const uint8_t* __restrict__ nverts, const T* __restrict__ P,
                   uint8_t* __restrict__ inside, int64_t B) {
    constexpr int K = N + 1;
    const int64_t s = (int64_t)blockIdx.x * kSimplexThreads + threadIdx.x;
    if (s >= B) return;
    T v[K * N], p[N];
    rec_load<T, K * N, SOA>(verts, s, vstride, v);
    rec_load<T, N, SOA>(P, s, N, p);
    // column i of the K x K system is (1, pts[i]); the right-hand side is (1, p)   (:455-465)
    T a[K][K], x[K][1];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        a[0][i] = (T)1;
#pragma unroll
        for (int d = 0; d < N; ++d) a[1 + d][i] = v[i * N + d];
    }
    x[0][0] = (T)1;
#pragma unroll
    for (int d = 0; d < N; ++d) x[1 + d][0] = p[d];
    uint8_t perm[K];
    bool parity;
    const int degenerate = plu_rows<T, K, K, 1, false>(a, x, perm, parity);
    backsolve_lu_regs<T, K, 1, /*FORWARD=*/false>(a, x, /*skip=*/1);          // the weight sum is known to be 1 (:467-469)
    bool in = degenerate == 0;
    if (nverts != nullptr && nverts[s] < K) in = false;                       // not a volume (:451)
    T sum = (T)0;
#pragma unroll
    for (int i = 1; i < K; ++i) {                                             // :471-474
        sum = add_rn(sum, x[i][0]);
        if (x[i][0] < (T)0 || sum > (T)1) in = false;
    }
    inside[s] = in ? 1 : 0;
}

// ------------------------------------------------------------------------------ trace_simplex
// returns hit; x holds the solution (uv = x[0..N-2], ray parameter = x[N-1])
template <typename T, int N>
__device__ __forceinline__ bool trace_regs(const T (&v)[N * N], const T (&o)[N], const T (&d)[N], bool want_uv, T (&x)[1][N]) {
    T bases[N * N];
#pragma unroll
    for (int i = 1; i < N; ++i)
#pragma unroll
        for (int k = 0; k < N; ++k) bases[(i - 1) * N + k] = add_rn(v[i * N + k], -v[k]);     // verts[i] - verts[0] (:657-659)
#pragma unroll
    for (int k = 0; k < N; ++k) {
        bases[(N - 1) * N + k] = -d[k];                                                      // :660
        x[0][k] = add_rn(o[k], -v[k]);                                                       // :663
    }
    bool hit = solve_vec_regs<T, N, 1>(bases, x, want_uv ? 0 : N - 1);                       // :664-665
    T sum = (T)0;
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {                                                        // :668-672
        if (x[0][i] < (T)0) hit = false;
        sum = add_rn(sum, x[0][i]);
    }
    if (sum > (T)1) hit = false;
    return hit;
}

template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kSimplexThreads)
k_trace_simplex(const T* __restrict__ verts, int64_t vstride, const T* __restrict__ origin, const T* __restrict__ dir,
                int64_t rstride, uint8_t* __restrict__ hit_out, T* __restrict__ s_out, T* __restrict__ uv_out, int64_t B) {
    const int64_t s = (int64_t)blockIdx.x * kSimplexThreads + threadIdx.x;
    if (s >= B) return;
    T v[N * N], o[N], d[N], x[1][N];
    rec_load<T, N * N, SOA>(verts, s, vstride, v);
    rec_load<T, N, SOA>(origin, s, rstride, o);
    rec_load<T, N, SOA>(dir, s, rstride, d);
    const bool hit = trace_regs<T, N>(v, o, d, uv_out != nullptr, x);
    hit_out[s] = hit ? 1 : 0;
    if (hit) {                                                                               // outputs untouched on a miss
        // the reference reads x[N] here (:675), one past the end of its Vec<T,N>; the ray parameter is x[N-1]
        s_out[s] = x[0][N - 1];
        if (uv_out != nullptr) {
            T uv[N - 1];
#pragma unroll
            for (int i = 0; i < N - 1; ++i) uv[i] = x[0][i];
            rec_store<T, N - 1, SOA>(uv_out, s, uv);
        }
    }
}

// ------------------------------------------------------------------------------ intersect(Ray)
template <typename T> __device__ __forceinline__ T std_min(T a, T b) { return (b < a) ? b : a; }   // std::min
template <typename T> __device__ __forceinline__ T std_max(T a, T b) { return (a < b) ? b : a; }   // std::max

template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kSimplexThreads)
k_simplex_intersect(const T* __restrict__ verts, int64_t vstride, const uint8_t* __restrict__ nverts,
                    const T* __restrict__ origin, const T* __restrict__ dir, int64_t rstride, T* __restrict__ interval,
                    int64_t B) {
    constexpr int K = N + 1;
    const int64_t s = (int64_t)blockIdx.x * kSimplexThreads + threadIdx.x;
    if (s >= B) return;
    T v[K * N], o[N], d[N];
    rec_load<T, K * N, SOA>(verts, s, vstride, v);
    rec_load<T, N, SOA>(origin, s, rstride, o);
    rec_load<T, N, SOA>(dir, s, rstride, d);
    const int nv = nverts ? (int)nverts[s] : K;
    const T inf = std::numeric_limits<T>::infinity();
    // no hit: the default Rect() = (max, lowest), Rect.h:88-90 (:294-295)
    T lohi[2] = {std::numeric_limits<T>::max(), -std::numeric_limits<T>::max()};
    if (nv == K) {
        // m[i] = (pts[i], 1); B = {(origin, 1), (direction, 0)}   (:258-263)
        T m[K * K], x[2][K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int k = 0; k < N; ++k) m[i * K + k] = v[i * N + k];
            m[i * K + N] = (T)1;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            x[0][k] = o[k];
            x[1][k] = d[k];
        }
        x[0][N] = (T)1;
        x[1][N] = (T)0;
        if (solve_vec_regs<T, K, 2>(m, x, 0)) {                                              // :266
            T lo = -inf, hi = inf;                                                           // Rect::full, Rect.h:1200-1203
            bool open = true, empty = false;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                open = open && (hi > lo);                                                    // loop condition (:269)
                const T p_i = x[0][i], k_i = x[1][i];
                if (open && !empty) {
                    if (k_i != (T)0) {
                        const T s0 = div_rn(-p_i, k_i);
                        const T s1 = div_rn(add_rn((T)1, -p_i), k_i);
                        lo = std_max(lo, std_min(s0, s1));                                   // &= from_corners (Rect.h:168-176, 350-354)
                        hi = std_min(hi, std_max(s0, s1));
                    } else if (p_i < (T)0 || p_i > (T)1) {
                        empty = true;                                                        // Rect::empty (:279)
                    }
                }
            }
            lohi[0] = empty ? inf : lo;
            lohi[1] = empty ? -inf : hi;
        }
    } else if (nv == N) {
        T vt[N * N], x[1][N];
#pragma unroll
        for (int e = 0; e < N * N; ++e) vt[e] = v[e];
        if (trace_regs<T, N>(vt, o, d, /*want_uv=*/false, x)) {                              // :288-292
            lohi[0] = x[0][N - 1];
            lohi[1] = x[0][N - 1];
        }
    }
    rec_store<T, 2, SOA>(interval, s, lohi);
}

// ------------------------------------------------------------------------------ launchers
template <typename T>
int simplex_contains(int dim, int64_t B, const T* verts, int64_t vstride, const uint8_t* nverts, const T* p, uint8_t* inside,
                     int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(dim, [&](auto N) {
        const unsigned grid = (unsigned)ceil_div(B, kSimplexThreads);
        if (grid == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            k_simplex_contains<T, N.value, SOA.value><<<grid, kSimplexThreads, 0, st>>>(verts, vstride, nverts, p, inside, B);
        });
        rc = after_launch();
    });
    return rc;
}

template <typename T>
int trace_simplex(int dim, int64_t B, const T* verts, int64_t vstride, const T* origin, const T* dir, int64_t rstride,
                  uint8_t* hit, T* s_out, T* uv, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(dim, [&](auto N) {
        const unsigned grid = (unsigned)ceil_div(B, kSimplexThreads);
        if (grid == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            k_trace_simplex<T, N.value, SOA.value><<<grid, kSimplexThreads, 0, st>>>(verts, vstride, origin, dir, rstride, hit,
                                                                                 s_out, uv, B);
        });
        rc = after_launch();
    });
    return rc;
}

template <typename T>
int simplex_intersect(int dim, int64_t B, const T* verts, int64_t vstride, const uint8_t* nverts, const T* origin,
                      const T* dir, int64_t rstride, T* interval, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(dim, [&](auto N) {
        const unsigned grid = (unsigned)ceil_div(B, kSimplexThreads);
        if (grid == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            k_simplex_intersect<T, N.value, SOA.value><<<grid, kSimplexThreads, 0, st>>>(verts, vstride, nverts, origin, dir,
                                                                                     rstride, interval, B);
        });
        rc = after_launch();
    });
    return rc;
}

#define INSTANTIATE(T)                                                                                                  \
    template int simplex_contains<T>(int, int64_t, const T*, int64_t, const uint8_t*, const T*, uint8_t*, int,          \
                                     cudaStream_t);                                                                     \
    template int trace_simplex<T>(int, int64_t, const T*, int64_t, const T*, const T*, int64_t, uint8_t*, T*, T*, int,  \
                                  cudaStream_t);                                                                        \
    template int simplex_intersect<T>(int, int64_t, const T*, int64_t, const uint8_t*, const T*, const T*, int64_t, T*, \
                                      int, cudaStream_t);

INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace gmb
