// Exercises include/geomc_b200/batch.hpp the way a geomc user would after switching:
// the reference's call shapes (linear_solve(T*,n,m,T*), inv(&into, src), cholesky(&m),
// linear_solve(Vec[N],...), decomp_plu(...)), one call per batch.
//
//   g++ -std=c++17 -O2 -I include tests/cpp/test_batch_hpp.cpp -L geomc_b200 -lgeomc_b200 -Wl,-rpath,$PWD/geomc_b200
//   add -DWITH_GEOMC -std=c++23 -include oracle/shim/prelude.h -I oracle/shim -I /root/reference
//   to run the object forms on the reference's real SimpleMatrix / Vec types.
// `--compile-only` semantics: the not-gpu suite only compiles this file; the gpu suite runs it.
#ifdef WITH_GEOMC
#include <geomc/linalg/Matrix.h>
#include <geomc/linalg/Vec.h>
#include <geomc/linalg/Ray.h>
#include <geomc/shape/Simplex.h>
#include <geomc/linalg/AffineTransform.h>
#endif
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <geomc_b200/batch.hpp>

namespace gb = geom::batch;

#ifndef WITH_GEOMC
// Minimal stand-ins with the members batch.hpp's object forms rely on (same names as the reference).
namespace geom {
enum MatrixLayout { ROW_MAJOR, COL_MAJOR };
template <typename T, long M, long N, MatrixLayout L = ROW_MAJOR>
struct SimpleMatrix {
    typedef T elem_t;
    static constexpr long ROWDIM = M, COLDIM = N;
    static constexpr MatrixLayout Layout = L;
    T data[M * N];
    SimpleMatrix() { for (long i = 0; i < M * N; ++i) data[i] = (i / N == i % N) ? 1 : 0; }
    T* data_begin() { return data; }
    const T* data_begin() const { return data; }
    T& operator()(long r, long c) { return data[L == ROW_MAJOR ? N * r + c : M * c + r]; }
    T operator()(long r, long c) const { return data[L == ROW_MAJOR ? N * r + c : M * c + r]; }
};
template <typename T, long N>
struct Vec {
    T v[N];
    T* begin() { return v; }
    const T* begin() const { return v; }
    T& operator[](long i) { return v[i]; }
    T operator[](long i) const { return v[i]; }
};
template <typename T, long N>
struct Ray {
    Vec<T, N> origin, direction;
};
template <typename T, long N>
struct Simplex {
    Vec<T, N> pts[N + 1];
    long n;
};
template <typename T, long N>
struct Rect;
template <typename T>
struct Rect<T, 1> {
    T lo, hi;
};
template <typename T, long N>
struct AffineTransform {
    SimpleMatrix<T, N + 1, N + 1> mat, inv;
};
}  // namespace geom
#endif

static int failures = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { ++failures; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

template <typename T>
static void test_pointer_linear_solve(long n, long B) {
    std::mt19937_64 rng(17 + n);
    std::normal_distribution<double> g(0, 1);
    std::vector<T> a(B * n * n), b(B * n), x;
    for (auto& v : a) v = (T)g(rng);
    for (auto& v : b) v = (T)g(rng);
    for (long c = 0; c < n; ++c) a[7 * n * n + c * n] = 0;       // system 7: zero first column -> false
    x = b;
    std::vector<char> okc(B);
    bool* ok = reinterpret_cast<bool*>(okc.data());
    long nfalse = gb::linear_solve<T, true>(a.data(), n, 1, x.data(), B, ok);
    EXPECT(nfalse == 1 && !ok[7], "expected exactly system 7 to be degenerate (got %ld)", nfalse);
    for (long r = 0; r < n; ++r) EXPECT(x[7 * n + r] == b[7 * n + r], "x must be untouched on failure");
    double worst = 0;
    for (long s = 0; s < B; ++s) {
        if (!ok[s]) continue;
        double xm = 1;
        for (long r = 0; r < n; ++r) xm = std::fmax(xm, std::fabs((double)x[s * n + r]));
        for (long r = 0; r < n; ++r) {
            double acc = -(double)b[s * n + r];
            for (long c = 0; c < n; ++c) acc += (double)a[s * n * n + r * n + c] * (double)x[s * n + c];
            worst = std::fmax(worst, std::fabs(acc) / xm);
        }
    }
    EXPECT(worst < (sizeof(T) == 4 ? 2e-2 : 1e-8), "residual %g too large (n=%ld)", worst, n);
}

template <typename T, long N, geom::MatrixLayout LS, geom::MatrixLayout LD>
static void test_object_inv(long B) {
    std::mt19937_64 rng(5 + N);
    std::normal_distribution<double> g(0, 1);
    std::vector<geom::SimpleMatrix<T, N, N, LS>> src(B);
    std::vector<geom::SimpleMatrix<T, N, N, LD>> dst(B);
    for (auto& m : src) for (long r = 0; r < N; ++r) for (long c = 0; c < N; ++c) m(r, c) = (T)g(rng);
    for (long r = 0; r < N; ++r) for (long c = 0; c < N; ++c) src[3](r, c) = 0;      // singular
    std::vector<char> okc(B);
    bool* ok = reinterpret_cast<bool*>(okc.data());
    long nfalse = gb::inv(dst.data(), src.data(), B, ok);
    EXPECT(nfalse == 1 && !ok[3], "expected exactly system 3 to be singular");
    for (long r = 0; r < N; ++r) for (long c = 0; c < N; ++c)
        EXPECT(dst[3](r, c) == (r == c ? 1 : 0), "destination must keep its (identity) contents on failure");
    double worst = 0;
    for (long s = 0; s < B; ++s) {
        if (!ok[s]) continue;
        for (long r = 0; r < N; ++r) for (long c = 0; c < N; ++c) {
            double acc = 0;
            for (long k = 0; k < N; ++k) acc += (double)src[s](r, k) * (double)dst[s](k, c);
            worst = std::fmax(worst, std::fabs(acc - (r == c ? 1.0 : 0.0)));
        }
    }
    EXPECT(worst < (sizeof(T) == 4 ? 0.5 : 1e-6), "M * M^-1 deviates from I by %g (N=%ld)", worst, N);
}

template <typename T, long N>
static void test_vec_linear_solve(long B) {
    std::mt19937_64 rng(9 + N);
    std::normal_distribution<double> g(0, 1);
    std::vector<geom::Vec<T, N>> bases(B * N), x(B), b;
    for (auto& v : bases) for (long i = 0; i < N; ++i) v[i] = (T)g(rng);
    for (auto& v : x) for (long i = 0; i < N; ++i) v[i] = (T)g(rng);
    b = x;
    long nfalse = gb::linear_solve_vec(bases.data(), x.data(), N, B);
    EXPECT(nfalse == 0, "random bases should be regular");
    double worst = 0;
    for (long s = 0; s < B; ++s)
        for (long r = 0; r < N; ++r) {                  // sum_i x[i] * basis_i == b   (linear_solve.cpp:46-52)
            double acc = -(double)b[s][r];
            for (long i = 0; i < N; ++i) acc += (double)x[s][i] * (double)bases[s * N + i][r];
            worst = std::fmax(worst, std::fabs(acc));
        }
    EXPECT(worst < 1e-5, "Vec-form residual %g (N=%ld)", worst, N);
}

static void test_decomp_and_cholesky(long B) {
    const long n = 6;
    std::mt19937_64 rng(3);
    std::normal_distribution<double> g(0, 1);
    std::vector<double> m(B * n * n), lu;
    for (auto& v : m) v = g(rng);
    lu = m;
    std::vector<gb::index_t> p(B * n), deg(B);
    std::vector<char> par(B);
    gb::decomp_plu<double, true>(lu.data(), n, n, B, p.data(), reinterpret_cast<bool*>(par.data()), deg.data());
    double worst = 0;
    for (long s = 0; s < B; ++s)
        for (long r = 0; r < n; ++r) for (long c = 0; c < n; ++c) {       // L U == P M   (linear_solve.cpp:86-99)
            double acc = 0;
            for (long k = 0; k <= std::min(r, c); ++k) acc += (k == r ? 1.0 : lu[s * 36 + r * n + k]) * lu[s * 36 + k * n + c];
            worst = std::fmax(worst, std::fabs(acc - m[s * 36 + p[s * n + r] * n + c]));
        }
    EXPECT(worst < 1e-5, "L U != P M (%g)", worst);
    // SPD 3x3 double: cholesky object form + fused cholesky_solve host form
    std::vector<geom::SimpleMatrix<double, 3, 3>> S(B);
    std::vector<double> flat(B * 9), rhs(B * 3), x;
    for (long s = 0; s < B; ++s) {
        double R[9];
        for (auto& v : R) v = g(rng);
        for (long r = 0; r < 3; ++r) for (long c = 0; c < 3; ++c) {
            double acc = r == c ? 1 : 0;
            for (long k = 0; k < 3; ++k) acc += R[r * 3 + k] * R[c * 3 + k];
            S[s](r, c) = acc;
            flat[s * 9 + r * 3 + c] = acc;
        }
        for (long r = 0; r < 3; ++r) rhs[s * 3 + r] = g(rng);
    }
    x = rhs;
    const long not_spd = gb::cholesky_solve<double, true>(flat.data(), 3, 1, x.data(), B);
    EXPECT(not_spd == 0, "SPD systems must factor");
    EXPECT(gb::cholesky(S.data(), B) == 0, "SPD matrices must factor");
    worst = 0;
    for (long s = 0; s < B; ++s)
        for (long r = 0; r < 3; ++r) {
            double acc = -rhs[s * 3 + r];
            for (long c = 0; c < 3; ++c) acc += flat[s * 9 + r * 3 + c] * x[s * 3 + c];
            worst = std::fmax(worst, std::fabs(acc));
            for (long c = 0; c < 3; ++c) {
                double ll = 0;
                for (long k = 0; k < 3; ++k) ll += S[s](r, k) * S[s](c, k);
                worst = std::fmax(worst, std::fabs(ll - flat[s * 9 + r * 3 + c]));
            }
        }
    EXPECT(worst < 1e-9, "cholesky / cholesky_solve residual %g", worst);
}

template <typename T, long N>
static void test_orthogonal_nullspace(long B) {
    std::mt19937_64 rng(21 + N);
    std::normal_distribution<double> g(0, 1);
    std::vector<geom::Vec<T, N>> v(B * (N - 1)), o(B);
    for (auto& x : v) for (long i = 0; i < N; ++i) x[i] = (T)g(rng);
    gb::orthogonal(v.data(), o.data(), N, B);
    double worst = 0;
    for (long s = 0; s < B; ++s)
        for (long k = 0; k < N - 1; ++k) {
            double d = 0, nv = 0, no = 0;
            for (long i = 0; i < N; ++i) { d += (double)v[s * (N - 1) + k][i] * (double)o[s][i]; nv += (double)v[s * (N - 1) + k][i] * v[s * (N - 1) + k][i]; no += (double)o[s][i] * o[s][i]; }
            worst = std::fmax(worst, std::fabs(d) / std::sqrt(nv * no + 1e-300));
        }
    EXPECT(worst < (sizeof(T) == 4 ? 1e-3 : 1e-9), "orthogonal: cosine %g (N=%ld)", worst, N);
    if (N >= 4) {
        const long nb = 2;
        std::vector<geom::Vec<T, N>> bs(B * nb), ns(B * (N - nb));
        for (auto& x : bs) for (long i = 0; i < N; ++i) x[i] = (T)g(rng);
        long bad = gb::nullspace(bs.data(), nb, ns.data(), N, B);
        EXPECT(bad == 0, "random bases are independent");
        worst = 0;
        for (long s = 0; s < B; ++s)
            for (long j = 0; j < nb; ++j) for (long k = 0; k < N - nb; ++k) {   // linear_solve.cpp:141-147
                double d = 0;
                for (long i = 0; i < N; ++i) d += (double)bs[s * nb + j][i] * (double)ns[s * (N - nb) + k][i];
                worst = std::fmax(worst, std::fabs(d));
            }
        EXPECT(worst < (sizeof(T) == 4 ? 1e-2 : 1e-5), "nullspace: dot %g (N=%ld)", worst, N);
    }
}

// Simplex::contains / Simplex::intersect / trace_simplex on arrays of the reference's own objects
template <typename T, long N>
static void test_simplex_queries(long B) {
    std::mt19937_64 rng(77 + N);
    std::normal_distribution<double> g(0, 1);
    std::uniform_real_distribution<double> u(0.05, 1);
    std::vector<geom::Simplex<T, N>> sx(B);
    std::vector<geom::Vec<T, N>> pin(B), pout(B), tri(B * N), target(B);
    std::vector<geom::Ray<T, N>> rays(B), trays(B);
    for (long s = 0; s < B; ++s) {
        sx[s].n = N + 1;
        double w[N + 1], wsum = 0;
        for (long k = 0; k <= N; ++k) { w[k] = u(rng); wsum += w[k]; }
        for (long k = 0; k <= N; ++k) for (long i = 0; i < N; ++i) sx[s].pts[k][i] = (T)g(rng);
        for (long i = 0; i < N; ++i) {
            double c = 0;
            for (long k = 0; k <= N; ++k) c += w[k] / wsum * (double)sx[s].pts[k][i];
            pin[s][i] = (T)c;
            pout[s][i] = (T)(c + 50);
            rays[s].origin[i] = (T)(4 * g(rng));
        }
        for (long i = 0; i < N; ++i) rays[s].direction[i] = pin[s][i] - rays[s].origin[i];
        // a facet (N vertices) and a ray aimed at a point on it
        double wf[N], wfs = 0;
        for (long k = 0; k < N; ++k) { wf[k] = u(rng); wfs += wf[k]; }
        for (long k = 0; k < N; ++k) for (long i = 0; i < N; ++i) tri[s * N + k][i] = (T)g(rng);
        for (long i = 0; i < N; ++i) {
            double c = 0;
            for (long k = 0; k < N; ++k) c += wf[k] / wfs * (double)tri[s * N + k][i];
            target[s][i] = (T)c;
            trays[s].origin[i] = (T)(4 * g(rng));
        }
        for (long i = 0; i < N; ++i) trays[s].direction[i] = target[s][i] - trays[s].origin[i];
    }
    sx[5].n = N;                                                   // not a volume: contains nothing (Simplex.h:451)
    std::vector<char> in(B), out(B), hit(B);
    long n_in = gb::contains(sx.data(), pin.data(), B, reinterpret_cast<bool*>(in.data()));
    long n_out = gb::contains(sx.data(), pout.data(), B, reinterpret_cast<bool*>(out.data()));
    EXPECT(n_in == B - 1 && !in[5], "convex combinations must be inside (%ld of %ld, N=%ld)", n_in, B, N);
    EXPECT(n_out == 0, "far points must be outside (%ld, N=%ld)", n_out, N);
    sx[5].n = N + 1;
    std::vector<geom::Rect<T, 1>> iv(B);
    long nonempty = gb::intersect(sx.data(), rays.data(), B, iv.data());
    long covers = 0;
    const double tol = sizeof(T) == 4 ? 1e-2 : 1e-8;
    for (long s = 0; s < B; ++s) covers += iv[s].lo <= 1 + tol && iv[s].hi >= 1 - tol;
    EXPECT(nonempty == B && covers == B, "rays through interior points must cross the simplex around t = 1 (%ld, %ld of %ld, N=%ld)",
           nonempty, covers, B, N);
    std::vector<geom::Vec<T, N - 1>> uv(B);
    std::vector<T> t(B, (T)-1);
    long hits = gb::trace_simplex(tri.data(), trays.data(), uv.data(), t.data(), B, reinterpret_cast<bool*>(hit.data()));
    double worst = 0;
    for (long s = 0; s < B; ++s) worst = std::fmax(worst, std::fabs((double)t[s] - 1));
    EXPECT(hits == B && worst < tol, "rays aimed at a facet point must hit at t = 1 (%ld of %ld, |t-1| %g, N=%ld)", hits, B, worst, N);
}

// transformation / translation / compose / apply on arrays of AffineTransform objects, and mul on the
// matrices inside them (rows (f)3 and (f)4)
template <typename T, long N>
static void test_affine(long B) {
    std::mt19937_64 rng(99 + N);
    std::normal_distribution<double> g(0, 1);
    const double tol = sizeof(T) == 4 ? 2e-3 : 1e-10;
    std::vector<geom::SimpleMatrix<T, N, N>> M(B);
    std::vector<geom::Vec<T, N>> tx(B), p(B), q(B), back(B), q1(B);
    for (long s = 0; s < B; ++s)
        for (long r = 0; r < N; ++r) {
            for (long c = 0; c < N; ++c) M[s](r, c) = (T)(g(rng) + (r == c ? 4 : 0));
            tx[s][r] = (T)g(rng);
            p[s][r] = (T)g(rng);
        }
    for (long r = 0; r < N; ++r) for (long c = 0; c < N; ++c) M[2](r, c) = 0;          // singular
    std::vector<geom::AffineTransform<T, N>> xm(B), xt(B), xf(B);
    long singular = gb::transformation(xm.data(), M.data(), B);
    EXPECT(singular == 1, "exactly one singular matrix expected (%ld)", singular);
    for (long r = 0; r <= N; ++r) for (long c = 0; c <= N; ++c)
        EXPECT(xm[2].inv(r, c) == (r == c ? 1 : 0), "a singular matrix leaves inv = identity (AffineTransform.h:873)");
    gb::translation(xt.data(), tx.data(), B);
    gb::compose(xf.data(), xt.data(), xm.data(), B);                 // translation(t) * transformation(M): p -> M p + t
    gb::apply(xf.data(), p.data(), q.data(), B);
    gb::apply(xf.data(), q.data(), back.data(), B, GMB_XF_APPLY_INVERSE);
    gb::apply(xf[7], p.data(), q1.data(), B);
    double worst = 0, worst_back = 0, worst1 = 0;
    for (long s = 0; s < B; ++s)
        for (long r = 0; r < N; ++r) {
            double acc = (double)tx[s][r], acc1 = (double)tx[7][r];
            for (long c = 0; c < N; ++c) {
                acc += (double)M[s](r, c) * (double)p[s][c];
                acc1 += (double)M[7](r, c) * (double)p[s][c];
            }
            worst = std::fmax(worst, std::fabs(acc - (double)q[s][r]));
            worst1 = std::fmax(worst1, std::fabs(acc1 - (double)q1[s][r]));
            if (s != 2) worst_back = std::fmax(worst_back, std::fabs((double)back[s][r] - (double)p[s][r]));
        }
    EXPECT(worst < tol * 10 && worst1 < tol * 10, "apply must compute M p + t (%g, %g, N=%ld)", worst, worst1, N);
    EXPECT(worst_back < tol * 100, "apply_inverse must undo apply (%g, N=%ld)", worst_back, N);
    // mat * inv = identity, computed with the batched mul on the objects' own matrices; then mul_acc adds it again
    std::vector<geom::SimpleMatrix<T, N + 1, N + 1>> a(B), b(B), prod(B);
    for (long s = 0; s < B; ++s) { a[s] = xf[s].mat; b[s] = xf[s].inv; }
    gb::mul(prod.data(), a.data(), b.data(), B);
    double worst_id = 0;
    for (long s = 0; s < B; ++s) if (s != 2)
        for (long r = 0; r <= N; ++r) for (long c = 0; c <= N; ++c)
            worst_id = std::fmax(worst_id, std::fabs((double)prod[s](r, c) - (r == c ? 1.0 : 0.0)));
    EXPECT(worst_id < tol * 100, "mat * inv must be the identity (%g, N=%ld)", worst_id, N);
    gb::mul_acc(prod.data(), a.data(), b.data(), B);
    EXPECT(std::fabs((double)prod[0](0, 0) - 2.0) < tol * 100 && std::fabs((double)prod[0](0, 1)) < tol * 100, "mul_acc accumulates");
}

// a crash prints where it happened (build with -g -rdynamic for symbol names)
static void on_crash(int sig) {
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "fatal signal; backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

#define STAGE(call) do { std::printf("  %s\n", #call); call; } while (0)

int main() {
    std::setvbuf(stdout, nullptr, _IONBF, 0);              // a crash must not swallow the progress lines
    {   // handlers on an alternate stack, so that a stack overflow is reported too
        static char alt[1 << 16];
        stack_t ss{};
        ss.ss_sp = alt;
        ss.ss_size = sizeof(alt);
        sigaltstack(&ss, nullptr);
        struct sigaction sa{};
        sa.sa_handler = on_crash;
        sa.sa_flags = SA_ONSTACK;
        sigaction(SIGSEGV, &sa, nullptr);
        sigaction(SIGBUS, &sa, nullptr);
        sigaction(SIGABRT, &sa, nullptr);
    }
    if (gmb_device_count() == 0) {
        // no CPU fallback: the call must fail loudly
        float a[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, x[4] = {1, 2, 3, 4};
        try {
            gb::linear_solve<float, true>(a, 4, 1, x, 1);
            std::printf("FAIL: computed without a GPU\n");
            return 1;
        } catch (const gb::Error& e) {
            std::printf("no GPU: %s (expected)\n", e.what());
            return e.code == GMB_ERR_NO_DEVICE ? 0 : 1;
        }
    }
    for (long n : {2L, 3L, 4L, 5L, 8L, 16L}) {
        STAGE((test_pointer_linear_solve<float>(n, 3001)));
        STAGE((test_pointer_linear_solve<double>(n, 1777)));
    }
    STAGE((test_object_inv<float, 4, geom::ROW_MAJOR, geom::ROW_MAJOR>(2049)));
    STAGE((test_object_inv<double, 4, geom::ROW_MAJOR, geom::COL_MAJOR>(1000)));
    STAGE((test_object_inv<float, 3, geom::COL_MAJOR, geom::ROW_MAJOR>(1000)));     // 40-byte objects: packed path
    STAGE((test_object_inv<double, 6, geom::ROW_MAJOR, geom::ROW_MAJOR>(500)));
    STAGE((test_vec_linear_solve<double, 3>(1000)));
    STAGE((test_vec_linear_solve<double, 4>(1000)));
    STAGE((test_vec_linear_solve<double, 7>(1000)));
    STAGE((test_decomp_and_cholesky(999)));
    STAGE((test_orthogonal_nullspace<double, 3>(500)));
    STAGE((test_orthogonal_nullspace<double, 4>(500)));
    STAGE((test_orthogonal_nullspace<float, 4>(500)));
    STAGE((test_orthogonal_nullspace<double, 7>(300)));
    STAGE((test_simplex_queries<float, 3>(2000)));
    STAGE((test_simplex_queries<double, 2>(1000)));
    STAGE((test_simplex_queries<double, 3>(1000)));
    STAGE((test_simplex_queries<double, 5>(500)));
    STAGE((test_affine<float, 3>(3000)));
    STAGE((test_affine<double, 2>(1000)));
    STAGE((test_affine<double, 3>(1000)));
    STAGE((test_affine<double, 6>(400)));
    std::printf(failures ? "FAILED (%d)\n" : "batch.hpp OK (%d failures)\n", failures);
    return failures ? 1 : 0;
}
