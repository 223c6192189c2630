// oracle/_ref driver -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the UNMODIFIED reference headers where they lie under /root/reference (geomc/linalg/*.h)
// into oracle/_ref/libgeomc_ref.so and exposes the hot-path functions as a C ABI that loops over
// an AoS batch (one reference object per system).  Used by tests/ as the parity checker, by
// tests/golden/make_golden.py to generate the committed fixtures, and by bench.py's
// `cpu_baseline` / `--impl reference` legs (kind = "reference").  Nothing under geomc_b200/
// may link or load this.
//
// Each entry point calls exactly one reference function per system:
//   ref_decomp_plu_*      geom::decomp_plu<T,RowMajor>          LUDecomp.h:188
//   ref_decomp_lup_*      geom::decomp_lup<T,RowMajor>          LUDecomp.h:102
//   ref_backsolve_lu_*    geom::backsolve_lu<T,RowMajor>        LUDecomp.h:332
//   ref_backsolve_plu_*   geom::backsolve_plu<T,RowMajor>       LUDecomp.h:271
//   ref_apply_perm_*      geom::detail::destructive_apply_permutation  LUDecomp.h:35
//   ref_linear_solve_*    geom::linear_solve<T,RowMajor>(T*,n,m,T*,skip)   LUDecomp.h:388
//   ref_linear_solve_vec_*  geom::linear_solve<T,N>(Vec<T,N>[N],m,Vec<T,N>*,skip)  LUDecomp.h:417
//   ref_cholesky_*        geom::cholesky<T,RowMajor>(T*,n)      Cholesky.h:37
//   ref_cholesky_mtx_*    geom::cholesky(SimpleMatrix*)         Cholesky.h:83
//   ref_inv_*             geom::inv(Md*, const Mx&)             Matrix.h:717 (static N = 2..16)
//   ref_inv_raw_*         geom::inv2x2/inv3x3/inv4x4 / invNxN(T*,T*,n)   MatrixInv.h:49/75/117/186
//   ref_orthogonal_*      geom::orthogonal<T,N>(const Vec<T,N>[N-1])    Orthogonal.h:73 (+ N=2,3 overloads :101,:107)
//   ref_nullspace_*       geom::nullspace<T,N>(bases, n, null_basis)    Orthogonal.h:136
//   ref_simplex_contains_*   geom::Simplex<T,N>::contains(p)             shape/Simplex.h:450
//   ref_trace_simplex_*      geom::trace_simplex<T,N>(verts, ray, uv, s) shape/Simplex.h:647
//   ref_simplex_intersect_*  geom::Simplex<T,N>::intersect(ray)          shape/Simplex.h:251
//
// Build: see oracle/Makefile (g++ -std=c++23 -O3 -mavx2 -mfma -ffp-contract=fast
//        --param=avoid-fma-max-bits=0 -fno-tree-loop-vectorize -fopenmp; the Makefile explains each flag).
#include <geomc/linalg/Matrix.h>
#include <geomc/linalg/LUDecomp.h>
#include <geomc/linalg/Cholesky.h>
#include <geomc/linalg/Orthogonal.h>
#include <geomc/linalg/Vec.h>
#include <geomc/linalg/Ray.h>
#include <geomc/shape/Simplex.h>

#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// index_t is the reference's global typedef (geomc_defs.h:85): std::ptrdiff_t
static_assert(sizeof(index_t) == sizeof(int64_t), "index_t must be 64-bit");

namespace {

template <typename T, bool RM>
void plu_batch(int rows, int cols, int64_t B, T* m, int64_t* reorder, uint8_t* parity, int64_t* degen) {
    const int64_t sz = (int64_t)rows * cols;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        bool par = false;
        index_t* p = reorder ? reinterpret_cast<index_t*>(reorder + s * rows) : nullptr;
        index_t d = geom::decomp_plu<T, RM>(m + s * sz, rows, cols, p, &par);
        if (parity) parity[s] = par ? 1 : 0;
        if (degen) degen[s] = d;
    }
}

template <typename T, bool RM>
void lup_batch(int rows, int cols, int64_t B, T* m, int64_t* reorder, uint8_t* parity, int64_t* degen) {
    const int64_t sz = (int64_t)rows * cols;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        bool par = false;
        index_t* p = reorder ? reinterpret_cast<index_t*>(reorder + s * cols) : nullptr;
        index_t d = geom::decomp_lup<T, RM>(m + s * sz, rows, cols, p, &par);
        if (parity) parity[s] = par ? 1 : 0;
        if (degen) degen[s] = d;
    }
}

template <typename T, bool RM>
void backsolve_lu_batch(int n, int m, int64_t B, const T* lu, T* x, int skip) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::backsolve_lu<T, RM>(lu + s * n * n, n, m, x + s * n * m, skip);
    }
}

template <typename T, bool RM>
void backsolve_plu_batch(int n, int m, int64_t B, const T* plu, const int64_t* p, T* x, const T* b, int skip) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::backsolve_plu<T, RM>(plu + s * n * n, reinterpret_cast<const index_t*>(p + s * n), n, m,
                                   x + s * n * m, b + s * n * m, skip);
    }
}

template <typename T, bool RM>
void apply_perm_batch(int n, int m, int64_t B, int64_t* p, T* a) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::detail::destructive_apply_permutation<T, RM>(reinterpret_cast<index_t*>(p + s * n), a + s * n * m, n, m);
    }
}

template <typename T, bool RM>
void linear_solve_batch(int n, int m, int64_t B, T* a, T* x, int skip, uint8_t* ok) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        bool r = geom::linear_solve<T, RM>(a + s * n * n, n, m, x + s * n * m, skip);
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, index_t N>
void linear_solve_vec_batch(int m, int64_t B, T* bases, T* x, int skip, uint8_t* ok) {
    static_assert(sizeof(geom::Vec<T, N>) == sizeof(T) * N);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        auto* bs = reinterpret_cast<geom::Vec<T, N>*>(bases + s * N * N);
        auto* xs = reinterpret_cast<geom::Vec<T, N>*>(x + s * N * m);
        bool r = geom::linear_solve<T, N>(bs, m, xs, skip);
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, bool RM>
void cholesky_batch(int n, int64_t B, T* a, uint8_t* ok) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        bool r = geom::cholesky<T, RM>(a + s * n * n, n);
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, index_t N>
void cholesky_mtx_batch(int64_t B, T* a, uint8_t* ok) {
    // sizeof(SimpleMatrix<float,3,3>) is 40, not 36 (alignof 8): the packed batch is copied
    // through a local object rather than reinterpreted.
    using Mx = geom::SimpleMatrix<T, N, N>;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        Mx mx;
        std::memcpy(mx.data_begin(), a + s * N * N, sizeof(T) * N * N);
        bool r = geom::cholesky(&mx);
        std::memcpy(a + s * N * N, mx.data_begin(), sizeof(T) * N * N);
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, index_t N, geom::MatrixLayout LS, geom::MatrixLayout LD>
void inv_mtx_batch(int64_t B, const T* src, T* dst, uint8_t* ok) {
    using Ms = geom::SimpleMatrix<T, N, N, LS>;
    using Md = geom::SimpleMatrix<T, N, N, LD>;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        Ms a;
        Md d;
        std::memcpy(a.data_begin(), src + s * N * N, sizeof(T) * N * N);
        std::memcpy(d.data_begin(), dst + s * N * N, sizeof(T) * N * N);  // "untouched on failure" is observable
        bool r = geom::inv(&d, a);
        std::memcpy(dst + s * N * N, d.data_begin(), sizeof(T) * N * N);
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, index_t N>
int inv_mtx_dispatch(int64_t B, const T* src, T* dst, uint8_t* ok, int srm, int drm) {
    using geom::ROW_MAJOR;
    using geom::COL_MAJOR;
    if (srm && drm)        inv_mtx_batch<T, N, ROW_MAJOR, ROW_MAJOR>(B, src, dst, ok);
    else if (srm && !drm)  inv_mtx_batch<T, N, ROW_MAJOR, COL_MAJOR>(B, src, dst, ok);
    else if (!srm && drm)  inv_mtx_batch<T, N, COL_MAJOR, ROW_MAJOR>(B, src, dst, ok);
    else                   inv_mtx_batch<T, N, COL_MAJOR, COL_MAJOR>(B, src, dst, ok);
    return 0;
}

template <typename T>
void inv_raw_batch(int n, int64_t B, T* src_destroyed, T* dst, uint8_t* ok) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        T* a = src_destroyed + s * n * n;
        T* d = dst + s * n * n;
        bool r;
        switch (n) {
            case 2: r = geom::inv2x2(d, a); break;
            case 3: r = geom::inv3x3(d, a); break;
            case 4: r = geom::inv4x4(d, a); break;
            default: r = geom::invNxN(d, a, (index_t)n); break;
        }
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, index_t N>
void orthogonal_batch(int64_t B, const T* v, T* out) {
    static_assert(sizeof(geom::Vec<T, N>) == sizeof(T) * N);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        const auto* vs = reinterpret_cast<const geom::Vec<T, N>*>(v + s * N * (N - 1));
        geom::Vec<T, N> o = geom::orthogonal(vs);
        std::memcpy(out + s * N, o.begin(), sizeof(T) * N);
    }
}

template <typename T, index_t N>
void nullspace_batch(int n, int64_t B, const T* bases, T* null_basis, uint8_t* ok) {
    const int64_t nn = N - n;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        const auto* bs = reinterpret_cast<const geom::Vec<T, N>*>(bases + s * N * n);
        auto* ns = reinterpret_cast<geom::Vec<T, N>*>(null_basis + s * N * (nn > 0 ? nn : 0));
        bool r = geom::nullspace<T, N>(bs, n, ns);
        if (ok) ok[s] = r ? 1 : 0;
    }
}

template <typename T, index_t N>
void simplex_contains_batch(int64_t B, const T* verts, const uint8_t* nverts, const T* p, uint8_t* inside) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::Simplex<T, N> sx;
        std::memcpy(sx.pts[0].begin(), verts + s * (N + 1) * N, sizeof(T) * (N + 1) * N);
        sx.n = nverts ? nverts[s] : N + 1;
        geom::Vec<T, N> pt;
        std::memcpy(pt.begin(), p + s * N, sizeof(T) * N);
        inside[s] = sx.contains(pt) ? 1 : 0;
    }
}

// `s_out` receives what the reference writes: x[N], one past the end of its Vec<T,N> (Simplex.h:669,
// undefined behaviour) -- callers compare only `hit` and `uv`.
template <typename T, index_t N>
void trace_simplex_batch(int64_t B, const T* verts, const T* origin, const T* dir, int want_uv, uint8_t* hit, T* s_out,
                         T* uv) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::Vec<T, N> vs[N];
        std::memcpy(vs[0].begin(), verts + s * N * N, sizeof(T) * N * N);
        geom::Vec<T, N> o, d;
        std::memcpy(o.begin(), origin + s * N, sizeof(T) * N);
        std::memcpy(d.begin(), dir + s * N, sizeof(T) * N);
        geom::Ray<T, N> ray(o, d);
        geom::Vec<T, N - 1> uvv;
        T sv = 0;
        const bool h = geom::trace_simplex<T, N>(vs, ray, want_uv ? &uvv : nullptr, &sv);
        hit[s] = h ? 1 : 0;
        if (h) {
            s_out[s] = sv;
            if (want_uv) std::memcpy(uv + s * (N - 1), reinterpret_cast<const T*>(&uvv), sizeof(T) * (N - 1));
        }
    }
}

template <typename T, index_t N>
void simplex_intersect_batch(int64_t B, const T* verts, const uint8_t* nverts, const T* origin, const T* dir, T* lohi) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < B; ++s) {
        geom::Simplex<T, N> sx;
        std::memcpy(sx.pts[0].begin(), verts + s * (N + 1) * N, sizeof(T) * (N + 1) * N);
        sx.n = nverts ? nverts[s] : N + 1;
        geom::Vec<T, N> o, d;
        std::memcpy(o.begin(), origin + s * N, sizeof(T) * N);
        std::memcpy(d.begin(), dir + s * N, sizeof(T) * N);
        geom::Rect<T, 1> r = sx.intersect(geom::Ray<T, N>(o, d));
        lohi[2 * s] = r.lo;
        lohi[2 * s + 1] = r.hi;
    }
}

#define FOR_N_2_7(X) X(2) X(3) X(4) X(5) X(6) X(7)

template <typename T>
int simplex_contains_dispatch(int n, int64_t B, const T* verts, const uint8_t* nverts, const T* p, uint8_t* inside) {
    switch (n) {
#define CASE(N) case N: simplex_contains_batch<T, N>(B, verts, nverts, p, inside); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}
template <typename T>
int trace_simplex_dispatch(int n, int64_t B, const T* verts, const T* origin, const T* dir, int want_uv, uint8_t* hit,
                           T* s_out, T* uv) {
    switch (n) {
#define CASE(N) case N: trace_simplex_batch<T, N>(B, verts, origin, dir, want_uv, hit, s_out, uv); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}
template <typename T>
int simplex_intersect_dispatch(int n, int64_t B, const T* verts, const uint8_t* nverts, const T* origin, const T* dir,
                               T* lohi) {
    switch (n) {
#define CASE(N) case N: simplex_intersect_batch<T, N>(B, verts, nverts, origin, dir, lohi); return 0;
        FOR_N_2_7(CASE)
#undef CASE
        default: return -1;
    }
}

#define FOR_N_2_16(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

template <typename T>
int linear_solve_vec_dispatch(int n, int m, int64_t B, T* bases, T* x, int skip, uint8_t* ok) {
    switch (n) {
#define CASE(N) case N: linear_solve_vec_batch<T, N>(m, B, bases, x, skip, ok); return 0;
        FOR_N_2_16(CASE)
#undef CASE
        default: return -1;
    }
}

template <typename T>
int orthogonal_dispatch(int n, int64_t B, const T* v, T* out) {
    switch (n) {
#define CASE(N) case N: orthogonal_batch<T, N>(B, v, out); return 0;
        FOR_N_2_16(CASE)
#undef CASE
        default: return -1;
    }
}

template <typename T>
int nullspace_dispatch(int N_, int n, int64_t B, const T* bases, T* null_basis, uint8_t* ok) {
    switch (N_) {
#define CASE(N) case N: nullspace_batch<T, N>(n, B, bases, null_basis, ok); return 0;
        CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
        default: return -1;
    }
}

template <typename T>
int cholesky_mtx_dispatch(int n, int64_t B, T* a, uint8_t* ok) {
    switch (n) {
#define CASE(N) case N: cholesky_mtx_batch<T, N>(B, a, ok); return 0;
        FOR_N_2_16(CASE)
#undef CASE
        default: return -1;
    }
}

template <typename T>
int inv_dispatch(int n, int64_t B, const T* src, T* dst, uint8_t* ok, int srm, int drm) {
    switch (n) {
#define CASE(N) case N: return inv_mtx_dispatch<T, N>(B, src, dst, ok, srm, drm);
        FOR_N_2_16(CASE)
#undef CASE
        default: return -1;
    }
}

}  // namespace

#define DEFINE_FOR_TYPE(SUF, T)                                                                                   \
    extern "C" int ref_decomp_plu_##SUF(int rows, int cols, int64_t B, T* m, int64_t* reorder, uint8_t* parity,   \
                                        int64_t* degen, int row_major) {                                          \
        if (row_major) plu_batch<T, true>(rows, cols, B, m, reorder, parity, degen);                              \
        else           plu_batch<T, false>(rows, cols, B, m, reorder, parity, degen);                             \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_decomp_lup_##SUF(int rows, int cols, int64_t B, T* m, int64_t* reorder, uint8_t* parity,   \
                                        int64_t* degen, int row_major) {                                          \
        if (row_major) lup_batch<T, true>(rows, cols, B, m, reorder, parity, degen);                              \
        else           lup_batch<T, false>(rows, cols, B, m, reorder, parity, degen);                             \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_backsolve_lu_##SUF(int n, int m, int64_t B, const T* lu, T* x, int skip, int row_major) {  \
        if (row_major) backsolve_lu_batch<T, true>(n, m, B, lu, x, skip);                                         \
        else           backsolve_lu_batch<T, false>(n, m, B, lu, x, skip);                                        \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_backsolve_plu_##SUF(int n, int m, int64_t B, const T* plu, const int64_t* p, T* x,         \
                                           const T* b, int skip, int row_major) {                                 \
        if (row_major) backsolve_plu_batch<T, true>(n, m, B, plu, p, x, b, skip);                                 \
        else           backsolve_plu_batch<T, false>(n, m, B, plu, p, x, b, skip);                                \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_apply_perm_##SUF(int n, int m, int64_t B, int64_t* p, T* a, int row_major) {               \
        if (row_major) apply_perm_batch<T, true>(n, m, B, p, a);                                                  \
        else           apply_perm_batch<T, false>(n, m, B, p, a);                                                 \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_linear_solve_##SUF(int n, int m, int64_t B, T* a, T* x, int skip, uint8_t* ok,             \
                                          int row_major) {                                                        \
        if (row_major) linear_solve_batch<T, true>(n, m, B, a, x, skip, ok);                                      \
        else           linear_solve_batch<T, false>(n, m, B, a, x, skip, ok);                                     \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_linear_solve_vec_##SUF(int n, int m, int64_t B, T* bases, T* x, int skip, uint8_t* ok) {   \
        return linear_solve_vec_dispatch<T>(n, m, B, bases, x, skip, ok);                                         \
    }                                                                                                             \
    extern "C" int ref_cholesky_##SUF(int n, int64_t B, T* a, uint8_t* ok, int row_major) {                       \
        if (row_major) cholesky_batch<T, true>(n, B, a, ok);                                                      \
        else           cholesky_batch<T, false>(n, B, a, ok);                                                     \
        return 0;                                                                                                 \
    }                                                                                                             \
    extern "C" int ref_cholesky_mtx_##SUF(int n, int64_t B, T* a, uint8_t* ok) {                                  \
        return cholesky_mtx_dispatch<T>(n, B, a, ok);                                                             \
    }                                                                                                             \
    extern "C" int ref_inv_##SUF(int n, int64_t B, const T* src, T* dst, uint8_t* ok, int src_row_major,          \
                                 int dst_row_major) {                                                             \
        return inv_dispatch<T>(n, B, src, dst, ok, src_row_major, dst_row_major);                                 \
    }                                                                                                             \
    extern "C" int ref_orthogonal_##SUF(int n, int64_t B, const T* v, T* out) {                                   \
        return orthogonal_dispatch<T>(n, B, v, out);                                                              \
    }                                                                                                             \
    extern "C" int ref_nullspace_##SUF(int N_, int n, int64_t B, const T* bases, T* null_basis, uint8_t* ok) {    \
        return nullspace_dispatch<T>(N_, n, B, bases, null_basis, ok);                                            \
    }                                                                                                             \
    extern "C" int ref_simplex_contains_##SUF(int n, int64_t B, const T* verts, const uint8_t* nverts, const T* p,   \
                                              uint8_t* inside) {                                                  \
        return simplex_contains_dispatch<T>(n, B, verts, nverts, p, inside);                                      \
    }                                                                                                             \
    extern "C" int ref_trace_simplex_##SUF(int n, int64_t B, const T* verts, const T* origin, const T* dir,          \
                                           int want_uv, uint8_t* hit, T* s_out, T* uv) {                          \
        return trace_simplex_dispatch<T>(n, B, verts, origin, dir, want_uv, hit, s_out, uv);                      \
    }                                                                                                             \
    extern "C" int ref_simplex_intersect_##SUF(int n, int64_t B, const T* verts, const uint8_t* nverts,              \
                                               const T* origin, const T* dir, T* lohi) {                          \
        return simplex_intersect_dispatch<T>(n, B, verts, nverts, origin, dir, lohi);                             \
    }                                                                                                             \
    extern "C" int ref_inv_raw_##SUF(int n, int64_t B, T* src_destroyed, T* dst, uint8_t* ok) {                   \
        inv_raw_batch<T>(n, B, src_destroyed, dst, ok);                                                           \
        return 0;                                                                                                 \
    }

DEFINE_FOR_TYPE(f32, float)
DEFINE_FOR_TYPE(f64, double)

extern "C" int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

extern "C" void ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

extern "C" const char* ref_build_info() {
    return "geomc reference headers (/root/reference), g++ " __VERSION__
           " -std=c++23 -O3 -mavx2 -mfma -ffp-contract=fast --param=avoid-fma-max-bits=0 -fno-tree-loop-vectorize"
#ifdef _OPENMP
           " -fopenmp"
#endif
        ;
}
