// oracle/_ref driver, second translation unit -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// the steps either side of the inverse.  Compiles the UNMODIFIED reference
// headers where they lie under /root/reference and loops the reference's own functions over an AoS
// batch (see ref_driver.cpp for the rules that apply to this library).
//
//   ref_mul_*            geom::mul(Md*, const Ma&, const Mb&) / geom::mul_acc     linalg/Matrix.h:386 / :433
//                        (-> detail::_ImplMtxMul<Ma,Mb>::mul / mul_acc, mtxdetail/MatrixMult.h:54-80)
//   ref_mul_mv_*         SimpleMatrix * Vec  (operator* -> _ImplMtxMul<Mat,Vec>::mul, MatrixMult.h:181-190)
//   ref_mul_vm_*         Vec * SimpleMatrix  (MatrixMult.h:239-248)
//   ref_xf_from_matrix_* geom::transformation(mat)                                linalg/AffineTransform.h:862-884
//   ref_xf_translation_* geom::translation(Vec)                                   AffineTransform.h:814-821
//   ref_xf_scale_*       geom::scale(Vec)                                         AffineTransform.h:838-845
//   ref_xf_compose_*     xf1 * xf2 / xf1 / xf2                                    AffineTransform.h:142-159, 300-317
//   ref_xf_apply_*       apply / apply_inverse / apply_direction / apply_inverse_direction / apply_normal /
//                        apply_inverse_normal                                     AffineTransform.h:224-293
//
// A transform record is packed: (N+1)^2 elements of `mat` then (N+1)^2 of `inv`, both row-major (the
// SimpleMatrix default layout) -- sizeof(AffineTransform<float,2>) is 80, not 72, because SimpleMatrix
// is 8-byte aligned, so records are copied through local objects rather than reinterpreted.
#include <geomc/linalg/Matrix.h>
#include <geomc/linalg/Vec.h>
#include <geomc/linalg/AffineTransform.h>

#include <cstdint>
#include <cstring>

namespace {

using geom::COL_MAJOR;
using geom::MatrixLayout;
using geom::ROW_MAJOR;

// ------------------------------------------------------------------------------------ mul
template <typename T, index_t R, index_t K, index_t C, MatrixLayout LA, MatrixLayout LB, MatrixLayout LD>
void mul_batch(int64_t B, const T* a, const T* b, T* d, int acc) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::SimpleMatrix<T, R, K, LA> ma;
        geom::SimpleMatrix<T, K, C, LB> mb;
        geom::SimpleMatrix<T, R, C, LD> md;
        std::memcpy(ma.data_begin(), a + s * R * K, sizeof(T) * R * K);
        std::memcpy(mb.data_begin(), b + s * K * C, sizeof(T) * K * C);
        std::memcpy(md.data_begin(), d + s * R * C, sizeof(T) * R * C);
        if (acc) geom::mul_acc(&md, ma, mb);
        else geom::mul(&md, ma, mb);
        std::memcpy(d + s * R * C, md.data_begin(), sizeof(T) * R * C);
    }
}

template <typename T, index_t R, index_t K, index_t C>
int mul_layouts(int64_t B, const T* a, const T* b, T* d, int acc, int arm, int brm, int drm) {
    const int key = (arm ? 4 : 0) | (brm ? 2 : 0) | (drm ? 1 : 0);
    switch (key) {
        case 7: mul_batch<T, R, K, C, ROW_MAJOR, ROW_MAJOR, ROW_MAJOR>(B, a, b, d, acc); return 0;
        case 6: mul_batch<T, R, K, C, ROW_MAJOR, ROW_MAJOR, COL_MAJOR>(B, a, b, d, acc); return 0;
        case 5: mul_batch<T, R, K, C, ROW_MAJOR, COL_MAJOR, ROW_MAJOR>(B, a, b, d, acc); return 0;
        case 4: mul_batch<T, R, K, C, ROW_MAJOR, COL_MAJOR, COL_MAJOR>(B, a, b, d, acc); return 0;
        case 3: mul_batch<T, R, K, C, COL_MAJOR, ROW_MAJOR, ROW_MAJOR>(B, a, b, d, acc); return 0;
        case 2: mul_batch<T, R, K, C, COL_MAJOR, ROW_MAJOR, COL_MAJOR>(B, a, b, d, acc); return 0;
        case 1: mul_batch<T, R, K, C, COL_MAJOR, COL_MAJOR, ROW_MAJOR>(B, a, b, d, acc); return 0;
        default: mul_batch<T, R, K, C, COL_MAJOR, COL_MAJOR, COL_MAJOR>(B, a, b, d, acc); return 0;
    }
}

// shapes with every layout combination: squares 2..6 and 8, and the rectangular 2/3/4 mixes used by the tests
// shapes with row-major operands only (and the all-column-major combination): the larger squares
template <typename T, index_t R, index_t K, index_t C>
int mul_rowcol_only(int64_t B, const T* a, const T* b, T* d, int acc, int arm, int brm, int drm) {
    if (arm && brm && drm) { mul_batch<T, R, K, C, ROW_MAJOR, ROW_MAJOR, ROW_MAJOR>(B, a, b, d, acc); return 0; }
    if (!arm && !brm && !drm) { mul_batch<T, R, K, C, COL_MAJOR, COL_MAJOR, COL_MAJOR>(B, a, b, d, acc); return 0; }
    return -2;
}

template <typename T>
int mul_dispatch(int r, int k, int c, int64_t B, const T* a, const T* b, T* d, int acc, int arm, int brm, int drm) {
    const int key = r * 10000 + k * 100 + c;
    switch (key) {
#define FULL(R, K, C) case R * 10000 + K * 100 + C: return mul_layouts<T, R, K, C>(B, a, b, d, acc, arm, brm, drm);
#define SOME(R, K, C) case R * 10000 + K * 100 + C: return mul_rowcol_only<T, R, K, C>(B, a, b, d, acc, arm, brm, drm);
        FULL(2, 2, 2) FULL(3, 3, 3) FULL(4, 4, 4) FULL(5, 5, 5) FULL(8, 8, 8)
        FULL(2, 3, 4) FULL(4, 3, 2) FULL(3, 4, 2) FULL(2, 4, 3) FULL(4, 2, 3) FULL(3, 2, 4) FULL(3, 5, 2) FULL(5, 2, 7)
        SOME(6, 6, 6) SOME(7, 7, 7) SOME(9, 9, 9) SOME(10, 10, 10) SOME(11, 11, 11) SOME(12, 12, 12) SOME(13, 13, 13)
        SOME(14, 14, 14) SOME(15, 15, 15) SOME(16, 16, 16) SOME(16, 3, 9) SOME(7, 16, 5)
#undef FULL
#undef SOME
        default: return -1;
    }
}

// matrix * Vec and Vec * matrix (the operator* forms AffineTransform::apply uses, AffineTransform.h:226)
template <typename T, index_t N, MatrixLayout L>
void mul_mv_batch(int64_t B, const T* a, const T* v, T* out) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::SimpleMatrix<T, N, N, L> ma;
        geom::Vec<T, N> x, y;
        std::memcpy(ma.data_begin(), a + s * N * N, sizeof(T) * N * N);
        std::memcpy(x.begin(), v + s * N, sizeof(T) * N);
        y = ma * x;
        std::memcpy(out + s * N, y.begin(), sizeof(T) * N);
    }
}
template <typename T, index_t N, MatrixLayout L>
void mul_vm_batch(int64_t B, const T* v, const T* a, T* out) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::SimpleMatrix<T, N, N, L> ma;
        geom::Vec<T, N> x, y;
        std::memcpy(ma.data_begin(), a + s * N * N, sizeof(T) * N * N);
        std::memcpy(x.begin(), v + s * N, sizeof(T) * N);
        y = x * ma;
        std::memcpy(out + s * N, y.begin(), sizeof(T) * N);
    }
}

#define FOR_N_2_8(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#define FOR_N_2_7(X) X(2) X(3) X(4) X(5) X(6) X(7)

template <typename T>
int mul_vec_dispatch(int n, int vec_first, int64_t B, const T* a, const T* v, T* out, int rm) {
    switch (n) {
#define CASE(N)                                                                                 \
    case N:                                                                                     \
        if (vec_first) { if (rm) mul_vm_batch<T, N, ROW_MAJOR>(B, v, a, out); else mul_vm_batch<T, N, COL_MAJOR>(B, v, a, out); } \
        else           { if (rm) mul_mv_batch<T, N, ROW_MAJOR>(B, a, v, out); else mul_mv_batch<T, N, COL_MAJOR>(B, a, v, out); } \
        return 0;
        FOR_N_2_8(CASE)
#undef CASE
        default: return -1;
    }
}

// ------------------------------------------------------------------------------------ AffineTransform
template <typename T, index_t N>
void xf_get(geom::AffineTransform<T, N>& xf, const T* rec) {
    constexpr index_t KK = (N + 1) * (N + 1);
    std::memcpy(xf.mat.data_begin(), rec, sizeof(T) * KK);
    std::memcpy(xf.inv.data_begin(), rec + KK, sizeof(T) * KK);
}
template <typename T, index_t N>
void xf_put(T* rec, const geom::AffineTransform<T, N>& xf) {
    constexpr index_t KK = (N + 1) * (N + 1);
    std::memcpy(rec, xf.mat.data_begin(), sizeof(T) * KK);
    std::memcpy(rec + KK, xf.inv.data_begin(), sizeof(T) * KK);
}

template <typename T, index_t N, MatrixLayout L>
void xf_from_matrix_batch(int64_t B, const T* m, T* xf_out) {
    constexpr index_t E = 2 * (N + 1) * (N + 1);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::SimpleMatrix<T, N, N, L> mx;
        std::memcpy(mx.data_begin(), m + s * N * N, sizeof(T) * N * N);
        geom::AffineTransform<T, N> xf = geom::transformation(mx);
        xf_put<T, N>(xf_out + s * E, xf);
    }
}

template <typename T, index_t N>
void xf_translation_batch(int64_t B, const T* t, T* xf_out, int scale) {
    constexpr index_t E = 2 * (N + 1) * (N + 1);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::Vec<T, N> v;
        std::memcpy(v.begin(), t + s * N, sizeof(T) * N);
        geom::AffineTransform<T, N> xf = scale ? geom::scale(v) : geom::translation(v);
        xf_put<T, N>(xf_out + s * E, xf);
    }
}

template <typename T, index_t N>
void xf_compose_batch(int64_t B, const T* xf1, const T* xf2, T* out, int inverse) {
    constexpr index_t E = 2 * (N + 1) * (N + 1);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::AffineTransform<T, N> a, b;
        xf_get<T, N>(a, xf1 + s * E);
        xf_get<T, N>(b, xf2 + s * E);
        geom::AffineTransform<T, N> r = inverse ? (a / b) : (a * b);
        xf_put<T, N>(out + s * E, r);
    }
}

// mode: 0 apply, 1 apply_inverse, 2 apply_direction, 3 apply_inverse_direction, 4 apply_normal, 5 apply_inverse_normal
template <typename T, index_t N>
void xf_apply_batch(int mode, int64_t B, const T* xf, int64_t xf_count, const T* p, T* out) {
    constexpr index_t E = 2 * (N + 1) * (N + 1);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::AffineTransform<T, N> a;
        xf_get<T, N>(a, xf + (xf_count == 1 ? 0 : s) * E);
        geom::Vec<T, N> v, o;
        std::memcpy(v.begin(), p + s * N, sizeof(T) * N);
        switch (mode) {
            case 0: o = a.apply(v); break;
            case 1: o = a.apply_inverse(v); break;
            case 2: o = a.apply_direction(v); break;
            case 3: o = a.apply_inverse_direction(v); break;
            case 4: o = a.apply_normal(v); break;
            default: o = a.apply_inverse_normal(v); break;
        }
        std::memcpy(out + s * N, o.begin(), sizeof(T) * N);
    }
}

template <typename T>
int xf_from_matrix_dispatch(int n, int64_t B, const T* m, T* xf_out, int rm) {
    switch (n) {
#define CASE(N) case N: if (rm) xf_from_matrix_batch<T, N, ROW_MAJOR>(B, m, xf_out); else xf_from_matrix_batch<T, N, COL_MAJOR>(B, m, xf_out); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}
template <typename T>
int xf_translation_dispatch(int n, int64_t B, const T* t, T* xf_out, int scale) {
    switch (n) {
#define CASE(N) case N: xf_translation_batch<T, N>(B, t, xf_out, scale); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}
template <typename T>
int xf_compose_dispatch(int n, int64_t B, const T* xf1, const T* xf2, T* out, int inverse) {
    switch (n) {
#define CASE(N) case N: xf_compose_batch<T, N>(B, xf1, xf2, out, inverse); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}
template <typename T>
int xf_apply_dispatch(int n, int mode, int64_t B, const T* xf, int64_t xf_count, const T* p, T* out) {
    switch (n) {
#define CASE(N) case N: xf_apply_batch<T, N>(mode, B, xf, xf_count, p, out); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}

}  // namespace

#define DEFINE_FOR_TYPE(SUF, T)                                                                                      \
    extern "C" int ref_mul_##SUF(int r, int k, int c, int64_t B, const T* a, const T* b, T* d, int acc, int a_rm,    \
                                 int b_rm, int d_rm) {                                                               \
        return mul_dispatch<T>(r, k, c, B, a, b, d, acc, a_rm, b_rm, d_rm);                                          \
    }                                                                                                                \
    extern "C" int ref_mul_vec_##SUF(int n, int vec_first, int64_t B, const T* a, const T* v, T* out, int rm) {      \
        return mul_vec_dispatch<T>(n, vec_first, B, a, v, out, rm);                                                  \
    }                                                                                                                \
    extern "C" int ref_xf_from_matrix_##SUF(int n, int64_t B, const T* m, T* xf_out, int rm) {                       \
        return xf_from_matrix_dispatch<T>(n, B, m, xf_out, rm);                                                      \
    }                                                                                                                \
    extern "C" int ref_xf_translation_##SUF(int n, int64_t B, const T* t, T* xf_out, int scale) {                    \
        return xf_translation_dispatch<T>(n, B, t, xf_out, scale);                                                   \
    }                                                                                                                \
    extern "C" int ref_xf_compose_##SUF(int n, int64_t B, const T* xf1, const T* xf2, T* out, int inverse) {         \
        return xf_compose_dispatch<T>(n, B, xf1, xf2, out, inverse);                                                 \
    }                                                                                                                \
    extern "C" int ref_xf_apply_##SUF(int n, int mode, int64_t B, const T* xf, int64_t xf_count, const T* p,         \
                                      T* out) {                                                                      \
        return xf_apply_dispatch<T>(n, mode, B, xf, xf_count, p, out);                                               \
    }

DEFINE_FOR_TYPE(f32, float)
DEFINE_FOR_TYPE(f64, double)
