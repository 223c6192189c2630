// gmb_xform.cuh -- SURVEY.md 8(f)3 and 8(f)4: the steps either side of the inverse, batched.
//
//   (f)4  mul / mul_acc                linalg/Matrix.h:386-421 / 433-456 -> _ImplMtxMul::mul / mul_acc,
//         matrix * Vec, Vec * matrix   linalg/mtxdetail/MatrixMult.h:54-80, 181-190, 239-248
//   (f)3  AffineTransform<T,N>         linalg/AffineTransform.h: transformation(mat) :862-884, translation :814-821,
//                                      scale :838-845, `*` and `/` :142-159 (-> apply / apply_inverse :300-317),
//                                      apply / apply_inverse / apply_direction / apply_inverse_direction /
//                                      apply_normal / apply_inverse_normal on points :224-293
//
// Arithmetic: every product term is ONE fused multiply-add into the running sum, which starts at 0 (mul, the
// apply family) or at the destination element (mul_acc), terms in ascending inner index -- what the reference's
// build emits for `sum += a(r,n) * b(n,c)` (see oracle/Makefile on contraction).
//
// Included once per element type by gmb_xform_f32.cu / gmb_xform_f64.cu (GMB_XFORM_T).
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"
#include "gmb_work.cuh"

namespace gmb {

constexpr int kMulAcc = 1, kMulARm = 2, kMulBRm = 4, kMulDRm = 8;

// ================================================================================ (f)4 mul, register shapes
// D (R x C) = A (R x K) * B (K x C), every dimension <= 4: one system per thread slot (V slots per thread in
// SoA), operands and result in registers, all loops unrolled.  The three layout flags are runtime values: each
// logical element is picked from its two possible flat positions with one select.
// D may alias A or B (the reference multiplies into a temporary then, Matrix.h:409-417): a thread loads all of
// its systems before it stores any.
template <typename T, int R, int K, int C, bool SOA>
__global__ void __launch_bounds__(kThreads)
k_mul_small(const T* A, const T* Bm, T* D, int flags, int64_t B) {
    using Wk = Work<T, SOA>;
    constexpr int V = Wk::V;
    int64_t unit, s0;
    int lane;
    if (!Wk::locate(B, unit, lane, s0)) return;
    const bool acc = flags & kMulAcc, arm = flags & kMulARm, brm = flags & kMulBRm, drm = flags & kMulDRm;
    T qa[V][R * K], qb[V][K * C], qd[V][R * C];
    load_elems<T, R * K, SOA>(A, unit, lane, B, qa);
    load_elems<T, K * C, SOA>(Bm, unit, lane, B, qb);
    if (acc) {                                                    // warp-uniform
        load_elems<T, R * C, SOA>(D, unit, lane, B, qd);
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int e = 0; e < R * C; ++e) qd[v][e] = (T)0;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T a[R][K], b[K][C], d[R][C];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < K; ++k) a[r][k] = (R == 1 || K == 1 || arm) ? qa[v][K * r + k] : qa[v][R * k + r];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) b[k][c] = (K == 1 || C == 1 || brm) ? qb[v][C * k + c] : qb[v][K * c + k];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                T sum = (R == 1 || C == 1 || drm) ? qd[v][C * r + c] : qd[v][R * c + r];     // 0, or d(r,c) for mul_acc
#pragma unroll
                for (int k = 0; k < K; ++k) sum = fma_rn(a[r][k], b[k][c], sum);             // MatrixMult.h:60
                d[r][c] = sum;
            }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // flat position e holds d(r,c) with (r,c) = row-major split of e, or the column-major one
                const int e = C * r + c;
                qd[v][e] = (R == 1 || C == 1 || drm) ? d[r][c] : d[e % R][e / R];
            }
    }
    store_elems<T, R * C, SOA>(D, unit, lane, B, qd);
}

// ================================================================================ (f)4 mul, any shape up to 16
// One thread per (system, row of D): the row of A sits in registers, B is read through L1 (every thread of a
// warp works on the same row index of consecutive systems, so SoA planes are read coalesced).  D must not
// alias A or B on this path.
template <typename T>
struct ElemAddr {
    int layout;
    int E;
    __device__ __forceinline__ int64_t operator()(int64_t s, int e) const {
        if (layout == GMB_LAYOUT_SOA) {
            constexpr int W = soa_w<T>();
            return ((s / W) * E + e) * W + (s % W);
        }
        return s * E + e;
    }
};

template <typename T>
__global__ void __launch_bounds__(128)
g_mul(int R, int K, int C, int64_t B, const T* __restrict__ A, const T* __restrict__ Bm, T* __restrict__ D, int flags,
      int layout) {
    const int64_t s = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= B) return;
    const bool acc = flags & kMulAcc, arm = flags & kMulARm, brm = flags & kMulBRm, drm = flags & kMulDRm;
    const ElemAddr<T> ata{layout, R * K}, atb{layout, K * C}, atd{layout, R * C};
    T arow[GMB_MAX_N];
#pragma unroll
    for (int k = 0; k < GMB_MAX_N; ++k) arow[k] = k < K ? __ldg(A + ata(s, arm ? K * r + k : R * k + r)) : (T)0;
    for (int c = 0; c < C; ++c) {
        const int64_t di = atd(s, drm ? C * r + c : R * c + r);
        T sum = acc ? D[di] : (T)0;
#pragma unroll
        for (int k = 0; k < GMB_MAX_N; ++k)
            if (k < K) sum = fma_rn(arow[k], __ldg(Bm + atb(s, brm ? C * k + c : K * c + k)), sum);
        D[di] = sum;
    }
}

// ================================================================================ (f)3 AffineTransform
// A transform record = K*K elements of `mat` then K*K of `inv`, K = N + 1, both row-major (AffineTransform.h:91-93;
// SimpleMatrix's default layout).  One transform per thread in both layouts.
constexpr int kXfThreads = 128;

template <typename T, int N>
struct Xf {
    static constexpr int K = N + 1;
    static constexpr int KK = K * K;
    static constexpr int E = 2 * KK;
};

// store a whole record: AoS records of 32..256 bytes go through the warp-staged transpose (gmb_io.cuh) when the
// warp is complete and the buffer is 16-byte aligned; everything else element by element
template <typename T, int E, bool SOA>
__device__ __forceinline__ void xf_store_record(T* base, int64_t s, int lane, int64_t B, bool aligned, const T (&q)[E]) {
    constexpr int Rb = E * (int)sizeof(T);
    if constexpr (!SOA && Rb >= 32 && Rb <= 256) {
        const int64_t s_warp0 = s - lane;
        if (aligned && s_warp0 + 32 <= B) {                        // warp-uniform
            aos_store_staged<T, E>(base, s_warp0, lane, q);
            return;
        }
    }
    if (s < B) rec_store<T, E, SOA>(base, s, q);
}

// ---- transformation(mat), AffineTransform.h:862-884 ------------------------------------------------
// m_inv starts as the identity and inv()'s result is ignored (:873): a singular matrix yields inv = identity.
// N <= 4: adjugate on the flat array, transposed when the source is column-major (MatrixInv.h:290-293 etc.;
// m_inv is row-major).  N >= 5: the logical matrix is copied to a row-major temporary, then invNxN
// (MatrixInv.h:243-249, 219-231): decomp_plu, fail on a degenerate column, back-substitution of the permuted
// identity -- here the identity rides through the elimination (gmb_math.cuh, plu_rows).
template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kXfThreads)
k_xf_from_matrix(const T* __restrict__ M, T* __restrict__ xf_out, uint8_t* __restrict__ ok, int src_rm, int aligned, int64_t B) {
    using X = Xf<T, N>;
    constexpr int K = X::K;
    const int64_t s = (int64_t)blockIdx.x * kXfThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (s - lane >= B) return;                                     // warp-uniform: the staged store is cooperative
    const bool valid = s < B;
    T q[N * N];
    if (valid) rec_load<T, N * N, SOA>(M, s, N * N, q);
    else {
#pragma unroll
        for (int e = 0; e < N * N; ++e) q[e] = (T)0;
    }
    T minv[N][N];
    bool good;
    if constexpr (N <= 4) {
        T out[N * N];
#pragma unroll
        for (int e = 0; e < N * N; ++e) out[e] = (e / N == e % N) ? (T)1 : (T)0;
        good = inv_flat<T>(out, q);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) minv[r][c] = src_rm ? out[N * r + c] : out[N * c + r];
    } else {
        T a[N][N], x[N][N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                a[r][c] = src_rm ? q[N * r + c] : q[N * c + r];
                x[r][c] = (r == c) ? (T)1 : (T)0;
            }
        uint8_t perm[N];
        bool parity;
        good = plu_rows<T, N, N, N, false>(a, x, perm, parity) == 0;
        backsolve_lu_regs<T, N, N, /*FORWARD=*/false>(a, x, 0);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) minv[r][c] = good ? x[r][c] : ((r == c) ? (T)1 : (T)0);
    }
    T rec[X::E];
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const bool inner = r < N && c < N;
            const T id = (r == c) ? (T)1 : (T)0;
            rec[K * r + c] = inner ? (src_rm ? q[N * (r % N) + (c % N)] : q[N * (c % N) + (r % N)]) : id;
            rec[X::KK + K * r + c] = inner ? minv[r % N][c % N] : id;
        }
    xf_store_record<T, X::E, SOA>(xf_out, s, lane, B, aligned != 0, rec);
    if (ok != nullptr && valid) ok[s] = good ? 1 : 0;
}

// ---- translation(tx) :814-821 / scale(sx) :838-845 ---------------------------------------------------
template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kXfThreads)
k_xf_translation(const T* __restrict__ t, T* __restrict__ xf_out, int scale, int aligned, int64_t B) {
    using X = Xf<T, N>;
    constexpr int K = X::K;
    const int64_t s = (int64_t)blockIdx.x * kXfThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (s - lane >= B) return;
    T v[N];
    if (s < B) rec_load<T, N, SOA>(t, s, N, v);
    else {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = (T)1;
    }
    T rec[X::E];
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const T id = (r == c) ? (T)1 : (T)0;
            T m = id, mi = id;
            if (r < N && c == N) {                                 // translation column
                m = scale ? id : v[r % N];
                mi = scale ? id : -v[r % N];
            }
            if (r < N && c == r) {                                 // scale diagonal
                m = scale ? v[r % N] : id;
                mi = scale ? div_rn((T)1, v[r % N]) : id;
            }
            rec[K * r + c] = m;
            rec[X::KK + K * r + c] = mi;
        }
    xf_store_record<T, X::E, SOA>(xf_out, s, lane, B, aligned != 0, rec);
}

// ---- xf1 * xf2 and xf1 / xf2, AffineTransform.h:142-159 -> :300-317 -----------------------------------
//   *:  mat = xf1.mat * xf2.mat     inv = xf2.inv * xf1.inv
//   /:  mat = xf2.inv * xf1.mat     inv = xf1.inv * xf2.mat
// Each product keeps the right operand (K*K) in registers and streams the left operand and the result row by
// row.  `out` must not alias an input.
template <typename T, int N, bool SOA>
__device__ __forceinline__ void xf_product(const T* __restrict__ L, int loff, const T* __restrict__ Rt, int roff,
                                           T* __restrict__ out, int ooff, int64_t s) {
    using X = Xf<T, N>;
    constexpr int K = X::K;
    T y[X::KK];
    rec_load_part<T, X::E, X::KK, SOA>(Rt, s, X::E, roff, y);
#pragma unroll
    for (int r = 0; r < K; ++r) {
        T xr[K], o[K];
        rec_load_part<T, X::E, K, SOA>(L, s, X::E, loff + K * r, xr);
#pragma unroll
        for (int c = 0; c < K; ++c) {
            T sum = (T)0;
#pragma unroll
            for (int k = 0; k < K; ++k) sum = fma_rn(xr[k], y[K * k + c], sum);
            o[c] = sum;
        }
        rec_store_part<T, X::E, K, SOA>(out, s, ooff + K * r, o);
    }
}

template <typename T, int N, bool SOA>
__global__ void __launch_bounds__(kXfThreads)
k_xf_compose(const T* __restrict__ xf1, const T* __restrict__ xf2, T* __restrict__ out, int inverse, int64_t B) {
    using X = Xf<T, N>;
    const int64_t s = (int64_t)blockIdx.x * kXfThreads + threadIdx.x;
    if (s >= B) return;
    if (inverse) {
        xf_product<T, N, SOA>(xf2, X::KK, xf1, 0, out, 0, s);
        xf_product<T, N, SOA>(xf1, X::KK, xf2, 0, out, X::KK, s);
    } else {
        xf_product<T, N, SOA>(xf1, 0, xf2, 0, out, 0, s);
        xf_product<T, N, SOA>(xf2, X::KK, xf1, X::KK, out, X::KK, s);
    }
}

// ---- the apply family on point arrays, AffineTransform.h:224-293 ---------------------------------------
//   0 apply                   (mat * (p,1)) truncated   sum from 0, c = 0..N (the last term is mat(r,N) * 1)
//   1 apply_inverse           (inv * (p,1)) truncated
//   2 apply_direction         o[r] += mat(r,c) * v[c]   c = 0..N-1
//   3 apply_inverse_direction o[r] += inv(r,c) * v[c]
//   4 apply_normal            o[r] += inv(c,r) * n[c]
//   5 apply_inverse_normal    o[r] += mat(c,r) * n[c]
// Only the first N rows of the chosen matrix are read (row N of an affine matrix never reaches the result).
// BCAST: ONE packed record serves every point (the common "transform a point array" call): each thread reads
// it once and walks ITER points.
template <typename T, int N, bool SOA, bool BCAST>
__global__ void __launch_bounds__(kXfThreads)
k_xf_apply(const T* __restrict__ xf, const T* __restrict__ P, T* __restrict__ out, int mode, int64_t B) {
    using X = Xf<T, N>;
    constexpr int K = X::K;
    constexpr int ITER = BCAST ? 8 : 1;
    const bool use_inv = mode == 1 || mode == 3 || mode == 4;
    const bool hom = mode <= 1;
    const bool transposed = mode >= 4;
    const int off = use_inv ? X::KK : 0;
    const int64_t s_first = ((int64_t)blockIdx.x * ITER) * kXfThreads + threadIdx.x;
    if (s_first >= B) return;
    T m[N * K];
    if constexpr (BCAST) rec_load_part<T, X::E, N * K, false>(xf, 0, X::E, off, m);
    else rec_load_part<T, X::E, N * K, SOA>(xf, s_first, X::E, off, m);
    // all of this thread's points are requested before any is used (ITER independent loads in flight)
    T p[ITER][N];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int64_t s = s_first + (int64_t)it * kXfThreads;
        if (s < B) rec_load<T, N, SOA>(P, s, N, p[it]);
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int64_t s = s_first + (int64_t)it * kXfThreads;
        if (s >= B) break;
        T o[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            T x = (T)0;
#pragma unroll
            for (int c = 0; c < N; ++c) x = fma_rn(transposed ? m[K * c + r] : m[K * r + c], p[it][c], x);
            const T xh = fma_rn(m[K * r + N], (T)1, x);
            o[r] = hom ? xh : x;
        }
        rec_store<T, N, SOA>(out, s, o);
    }
}

// ================================================================================ launchers
#define GMB_XF_LAUNCH(kernel, grid, threads, stream, ...)                                    \
    do {                                                                                     \
        if ((grid) > 0) kernel<<<(unsigned)(grid), threads, 0, stream>>>(__VA_ARGS__);       \
        rc = after_launch();                                                                 \
    } while (0)

template <typename T>
int mul_small(int r, int k, int c, int64_t B, const T* A, const T* Bm, T* D, int flags, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<1, 2, 3, 4>(r, [&](auto R) {
        dispatch_int<1, 2, 3, 4>(k, [&](auto K) {
            dispatch_int<1, 2, 3, 4>(c, [&](auto C) {
                dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                    GMB_XF_LAUNCH((k_mul_small<T, R.value, K.value, C.value, SOA.value>), (Work<T, SOA.value>::grid(B)),
                                  kThreads, st, A, Bm, D, flags, B);
                });
            });
        });
    });
    return rc;
}

template <typename T>
int mul_generic(int r, int k, int c, int64_t B, const T* A, const T* Bm, T* D, int flags, int layout, cudaStream_t st) {
    int rc = GMB_OK;
    if (B > 0) {
        dim3 grid((unsigned)ceil_div(B, 128), (unsigned)r);
        g_mul<T><<<grid, 128, 0, st>>>(r, k, c, B, A, Bm, D, flags, layout);
    }
    rc = after_launch();
    return rc;
}

template <typename T>
int xf_from_matrix(int n, int64_t B, const T* M, T* xf_out, uint8_t* ok, int layout, int src_rm, bool aligned,
                   cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            GMB_XF_LAUNCH((k_xf_from_matrix<T, N.value, SOA.value>), ceil_div(B, kXfThreads), kXfThreads, st, M, xf_out, ok,
                          src_rm, (int)aligned, B);
        });
    });
    return rc;
}

template <typename T>
int xf_translation(int n, int64_t B, const T* t, T* xf_out, int scale, int layout, bool aligned, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            GMB_XF_LAUNCH((k_xf_translation<T, N.value, SOA.value>), ceil_div(B, kXfThreads), kXfThreads, st, t, xf_out,
                          scale, (int)aligned, B);
        });
    });
    return rc;
}

template <typename T>
int xf_compose(int n, int64_t B, const T* xf1, const T* xf2, T* out, int inverse, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            GMB_XF_LAUNCH((k_xf_compose<T, N.value, SOA.value>), ceil_div(B, kXfThreads), kXfThreads, st, xf1, xf2, out,
                          inverse, B);
        });
    });
    return rc;
}

template <typename T>
int xf_apply(int n, int mode, int64_t B, const T* xf, bool bcast, const T* p, T* out, int layout, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<2, 3, 4, 5, 6, 7>(n, [&](auto N) {
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            dispatch_bool(bcast, [&](auto BC) {
                const int64_t per_block = (int64_t)kXfThreads * (BC.value ? 8 : 1);
                GMB_XF_LAUNCH((k_xf_apply<T, N.value, SOA.value, BC.value>), ceil_div(B, per_block), kXfThreads, st, xf, p,
                              out, mode, B);
            });
        });
    });
    return rc;
}

#define GMB_XFORM_INSTANTIATE(T)                                                                                       \
    template int mul_small<T>(int, int, int, int64_t, const T*, const T*, T*, int, int, cudaStream_t);                 \
    template int mul_generic<T>(int, int, int, int64_t, const T*, const T*, T*, int, int, cudaStream_t);               \
    template int xf_from_matrix<T>(int, int64_t, const T*, T*, uint8_t*, int, int, bool, cudaStream_t);                \
    template int xf_translation<T>(int, int64_t, const T*, T*, int, int, bool, cudaStream_t);                          \
    template int xf_compose<T>(int, int64_t, const T*, const T*, T*, int, int, cudaStream_t);                          \
    template int xf_apply<T>(int, int, int64_t, const T*, bool, const T*, T*, int, cudaStream_t);

}  // namespace gmb
