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
// A CTA of 128 threads takes S consecutive systems: A, B (and D for mul_acc) are staged into shared memory with
// coalesced loads, each in its own flat order but with the inner dimension padded to 17 words (flat (q, m) at
// q * 17 + m, zero filled): consecutive threads write consecutive banks whatever the layout flag, and the strided
// reads below are conflict-free.  Each thread then owns one 4 x 4 tile of one system's D: per k it reads four a's
// and four b's and issues 16 FMAs, k ascending (the reference's term order, MatrixMult.h:58-61).  Rows / columns
// past the end read the zero padding and only feed tile entries that are never stored.  The tiles go back through shared memory in D's flat order and leave with coalesced stores.
// Every input of a system is staged before any of its outputs is written and CTAs own disjoint systems, so D may
// alias A or B here too.
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

constexpr int kMulThreads = 128;
constexpr int kMulPad = 17;                                       // staged operands: flat (q, m) -> q * 17 + m: odd stride, no bank conflicts

// n / d for n < 2^20, d <= 4096 with one multiply-high (the staging loops split flat indices by run-time sizes)
struct FastDiv {
    uint32_t d, m;
    FastDiv() = default;
    explicit FastDiv(int div) : d((uint32_t)div), m(div > 1 ? (uint32_t)((1ull << 32) / (uint32_t)div + 1) : 0) {}
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d > 1 ? __umulhi(n, m) : n; }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};
struct MulShape {
    int R, K, C, S;
    FastDiv ea, a2, eb, b2, ed, s, tps, tc, sv;  // R*K, (K|R), K*C, (C|K), R*C, S, threads per system, tiles per row, S / V
};

template <typename T>
__global__ void __launch_bounds__(kMulThreads)
k_mul_tiled(MulShape sh, int64_t B, const T* A, const T* Bm, T* D, int flags, int layout) {
    extern __shared__ __align__(16) unsigned char gmb_mul_smem[];
    const int R = sh.R, K = sh.K, C = sh.C, S = sh.S;
    const bool acc = flags & kMulAcc, arm = flags & kMulARm, brm = flags & kMulBRm, drm = flags & kMulDRm;
    const int per_sys = 2 * 16 * kMulPad + R * C;                     // staged elements per system: A and B blocks of 16 x 17, then D
    T* sm = reinterpret_cast<T*>(gmb_mul_smem);
    const int64_t s0 = (int64_t)blockIdx.x * S;
    const int ns = (int)min((int64_t)S, B - s0);
    const ElemAddr<T> ata{layout, R * K}, atb{layout, K * C}, atd{layout, R * C};
    const bool soa = layout == GMB_LAYOUT_SOA;

    // ---- stage: consecutive threads take consecutive addresses (AoS) / consecutive systems of one plane (SoA).
    // flat index -> (system in CTA, element): AoS idx = sl * E + e, SoA idx = e * S + sl.  Loads are issued in
    // batches of U before any is stored (a plain load -> store loop would serialise on the global-memory latency).
    // The padding is left uninitialised: whatever it holds only reaches tile entries that are never stored.
    constexpr int U = 8;
    auto stage = [&](const T* g, const ElemAddr<T>& at, int E, const FastDiv& fe, const FastDiv& f2, int dst_off, bool pad) {
        for (int base = threadIdx.x; base < S * E; base += kMulThreads * U) {
            T v[U];
            uint32_t sl[U], e[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t idx = (uint32_t)(base + u * kMulThreads);
                if (soa) sh.s.divmod(idx, e[u], sl[u]);
                else fe.divmod(idx, sl[u], e[u]);
                const bool ok = idx < (uint32_t)(S * E) && (int)sl[u] < ns;
                v[u] = ok ? g[at(s0 + (ok ? sl[u] : 0), (int)(ok ? e[u] : 0))] : (T)0;
                if (!ok) sl[u] = 0xffffffffu;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (sl[u] != 0xffffffffu) {
                    uint32_t q = 0, m = e[u];
                    if (pad) f2.divmod(e[u], q, m);
                    sm[sl[u] * per_sys + dst_off + (pad ? q * kMulPad + m : e[u])] = v[u];
                }
            }
        }
    };
    // AoS fast path: the CTA's S records are ONE contiguous range; when the inner dimension is a multiple of the
    // vector width a 128-bit load never straddles a row, so one index split serves V elements.
    constexpr int V = soa_v<T>();
    using VT = typename VecOf<T>::type;
    auto stage_vec = [&](const T* g, int E, int inner, const FastDiv& fe, const FastDiv& f2, int dst_off, bool pad) {
        const VT* gv = reinterpret_cast<const VT*>(g + s0 * (int64_t)E);
        const int nvec = ns * E / V;
        for (int base = threadIdx.x; base < nvec; base += kMulThreads * U) {
            VT v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kMulThreads;
                if (i < nvec) v[u] = __ldg(gv + i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kMulThreads;
                if (i < nvec) {
                    uint32_t sl, e, q = 0, m;
                    fe.divmod((uint32_t)(i * V), sl, e);
                    m = e;
                    if (pad) f2.divmod(e, q, m);
                    T t[V];
                    unpack(v[u], t);
                    T* dst = sm + sl * per_sys + dst_off + (pad ? q * kMulPad + m : e);
#pragma unroll
                    for (int w = 0; w < V; ++w) dst[w] = t[w];
                }
            }
        }
    };
    // SoA fast path: a 128-bit load covers V consecutive systems of one element plane (the CTA's systems sit inside
    // one tile and start on a multiple of V); one index split serves V systems.
    auto stage_vec_soa = [&](const T* g, int E, const FastDiv& f2, int dst_off, bool pad) {
        constexpr int W = soa_w<T>();
        const T* tile = g + (s0 / W) * (int64_t)(E * W) + (s0 % W);
        const int svec = S / V, nvec = E * svec;
        for (int base = threadIdx.x; base < nvec; base += kMulThreads * U) {
            VT v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kMulThreads;
                if (i < nvec) {
                    uint32_t e, j;
                    sh.sv.divmod((uint32_t)i, e, j);
                    v[u] = __ldg(reinterpret_cast<const VT*>(tile + (int64_t)e * W) + j);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kMulThreads;
                if (i < nvec) {
                    uint32_t e, j, q = 0, m;
                    sh.sv.divmod((uint32_t)i, e, j);
                    m = e;
                    if (pad) f2.divmod(e, q, m);
                    T t[V];
                    unpack(v[u], t);
                    T* dst = sm + (j * V) * per_sys + dst_off + (pad ? q * kMulPad + m : e);
#pragma unroll
                    for (int w = 0; w < V; ++w) dst[w * per_sys] = t[w];
                }
            }
        }
    };
    const bool soa_vec = soa && ns == S && S % V == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(Bm) & 15) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0;
    auto vec_ok = [&](const T* g, int E, int inner) {
        return !soa && inner % V == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && ((int64_t)S * E) % V == 0;
    };
    if (soa_vec) stage_vec_soa(A, R * K, sh.a2, 0, true);
    else if (vec_ok(A, R * K, arm ? K : R)) stage_vec(A, R * K, arm ? K : R, sh.ea, sh.a2, 0, true);
    else stage(A, ata, R * K, sh.ea, sh.a2, 0, true);
    if (soa_vec) stage_vec_soa(Bm, K * C, sh.b2, 16 * kMulPad, true);
    else if (vec_ok(Bm, K * C, brm ? C : K)) stage_vec(Bm, K * C, brm ? C : K, sh.eb, sh.b2, 16 * kMulPad, true);
    else stage(Bm, atb, K * C, sh.eb, sh.b2, 16 * kMulPad, true);
    if (acc) {
        if (soa_vec) stage_vec_soa(D, R * C, sh.ed, 2 * 16 * kMulPad, false);
        else stage(D, atd, R * C, sh.ed, sh.ed, 2 * 16 * kMulPad, false);
    }
    __syncthreads();

    // ---- one 4 x 4 tile per thread
    uint32_t sl, tile, tr, tc;
    sh.tps.divmod(threadIdx.x, sl, tile);
    sh.tc.divmod(tile, tr, tc);
    const int r0 = (int)tr * 4, c0 = (int)tc * 4;
    const bool active = (int)sl < ns;
    T d[4][4];
    if (active) {
        const T* sa = sm + sl * per_sys;
        const T* sb = sa + 16 * kMulPad;
        const T* sd = sb + 16 * kMulPad;
        // a(r,k) and b(k,c) in the staged flat order: row-major (q, m) = (row, column), column-major the other way round
        const int a_r = arm ? kMulPad : 1, a_k = arm ? 1 : kMulPad, b_k = brm ? kMulPad : 1, b_c = brm ? 1 : kMulPad;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + i, c = c0 + j;
                d[i][j] = (acc && r < R && c < C) ? sd[drm ? C * r + c : R * c + r] : (T)0;
            }
        for (int k = 0; k < K; ++k) {
            T a4[4], b4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a4[i] = sa[(r0 + i) * a_r + k * a_k];
                b4[i] = sb[k * b_k + (c0 + i) * b_c];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) d[i][j] = fma_rn(a4[i], b4[j], d[i][j]);       // MatrixMult.h:60
        }
    }
    __syncthreads();                                                  // everyone is done reading D's staged copy
    if (active) {
        T* sd = sm + sl * per_sys + 2 * 16 * kMulPad;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + i, c = c0 + j;
                if (r < R && c < C) sd[drm ? C * r + c : R * c + r] = d[i][j];
            }
    }
    __syncthreads();
    if (!soa && (R * C) % V == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 && ((int64_t)S * R * C) % V == 0 && per_sys % V == 0) {
        VT* gv = reinterpret_cast<VT*>(D + s0 * (int64_t)(R * C));
        const int nvec = ns * R * C / V;
        for (int i = threadIdx.x; i < nvec; i += kMulThreads) {
            uint32_t sl2, e;
            sh.ed.divmod((uint32_t)(i * V), sl2, e);
            gv[i] = *reinterpret_cast<const VT*>(sm + sl2 * per_sys + 2 * 16 * kMulPad + e);
        }
        return;
    }
    if (soa_vec) {
        constexpr int W = soa_w<T>();
        T* tile = D + (s0 / W) * (int64_t)(R * C * W) + (s0 % W);
        const int svec = S / V, nvec = R * C * svec;
        for (int i = threadIdx.x; i < nvec; i += kMulThreads) {
            uint32_t e, j;
            sh.sv.divmod((uint32_t)i, e, j);
            T t[V];
#pragma unroll
            for (int w = 0; w < V; ++w) t[w] = sm[(j * V + w) * per_sys + 2 * 16 * kMulPad + e];
            reinterpret_cast<VT*>(tile + (int64_t)e * W)[j] = pack(t);
        }
        return;
    }
    for (int idx = threadIdx.x; idx < S * R * C; idx += kMulThreads) {
        uint32_t sl2, e;
        if (soa) sh.s.divmod((uint32_t)idx, e, sl2);
        else sh.ed.divmod((uint32_t)idx, sl2, e);
        if ((int)sl2 < ns) D[atd(s0 + sl2, (int)e)] = sm[sl2 * per_sys + 2 * 16 * kMulPad + e];
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
        const int tiles_c = (c + 3) / 4, tps = ((r + 3) / 4) * tiles_c;
        const size_t per_sys = (size_t)(2 * 16 * kMulPad + r * c);
        int S = kMulThreads / tps;
        const int cap = (int)((48 * 1024) / (per_sys * sizeof(T)));
        if (S > cap) S = cap;
        if (layout == GMB_LAYOUT_SOA) {                    // keep a CTA inside one tile so that a plane segment is contiguous
            const int W = soa_w<T>();
            while (W % S != 0) --S;
        }
        MulShape sh;
        sh.R = r; sh.K = k; sh.C = c; sh.S = S;
        sh.ea = FastDiv(r * k); sh.a2 = FastDiv((flags & kMulARm) ? k : r);
        sh.eb = FastDiv(k * c); sh.b2 = FastDiv((flags & kMulBRm) ? c : k);
        sh.ed = FastDiv(r * c); sh.s = FastDiv(S); sh.tps = FastDiv(tps); sh.tc = FastDiv(tiles_c); sh.sv = FastDiv(S >= soa_v<T>() ? S / soa_v<T>() : 1);
        const unsigned grid = (unsigned)ceil_div(B, S);
        k_mul_tiled<T><<<grid, kMulThreads, (size_t)S * per_sys * sizeof(T), st>>>(sh, B, A, Bm, D, flags, layout);
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
