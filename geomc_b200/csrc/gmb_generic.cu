// gmb_generic.cu -- runtime-n kernels (2 <= n <= 16, any m that fits), one system per thread with
// the working matrix in thread-local memory.  Correctness path for every (n, m, layout) the
// specialised kernels (gmb_small.cu: n <= 4 in registers; gmb_subwarp.cu: n = 5..16 cooperative)
// do not cover: rectangular decompositions, m > 4, column-pivot LU for n >= 5, ...
// Same arithmetic, same operation order as gmb_math.cuh; each routine cites the reference.
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

constexpr int kGenThreads = 64;
constexpr int kMaxN = GMB_MAX_N;

// element e of system s for a batch with E elements per system
template <typename T>
struct Addr {
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

__device__ __forceinline__ int midx(bool rm, int rows, int cols, int r, int c) { return rm ? cols * r + c : rows * c + r; }

// decomp_plu / decomp_lup on a local row-major working copy a[rows*cols].  LUDecomp.h:188-250 / 102-165.
template <typename T, bool COLPIV>
__device__ int lu_local(T* a, int rows, int cols, uint8_t* perm, bool& parity) {
    const int n = rows < cols ? rows : cols;
    const int np = COLPIV ? cols : rows;
    int degenerate = 0;
    parity = false;
    for (int i = 0; i < np; ++i) perm[i] = (uint8_t)i;
    for (int i = 0; i < n - 1; ++i) {
        T biggest = abs_(a[cols * i + i]);
        int pvt = i;
        for (int p = i + 1; p < np; ++p) {
            T v = abs_(COLPIV ? a[cols * i + p] : a[cols * p + i]);
            if (v > biggest) {
                pvt = p;
                biggest = v;
            }
        }
        if (biggest == (T)0) {
            degenerate += 1;
            continue;
        } else if (pvt != i) {
            if (COLPIV) {
                for (int r = 0; r < rows; ++r) {
                    T t = a[cols * r + i];
                    a[cols * r + i] = a[cols * r + pvt];
                    a[cols * r + pvt] = t;
                }
            } else {
                for (int c = 0; c < cols; ++c) {
                    T t = a[cols * i + c];
                    a[cols * i + c] = a[cols * pvt + c];
                    a[cols * pvt + c] = t;
                }
            }
            uint8_t t = perm[i];
            perm[i] = perm[pvt];
            perm[pvt] = t;
            parity = !parity;
        }
        const T piv = a[cols * i + i];
        for (int r = i + 1; r < rows; ++r) {
            const T b = div_rn(a[cols * r + i], piv);
            if (b != (T)0) {
                for (int c = i + 1; c < cols; ++c) a[cols * r + c] = fma_rn(-b, a[cols * i + c], a[cols * r + c]);
            }
            a[cols * r + i] = b;
        }
    }
    return degenerate;
}

// backsolve_lu on local row-major copies: lu[n*n], x[n*m].  LUDecomp.h:332-366.
template <typename T>
__device__ void backsolve_local(const T* lu, int n, int m, T* x, int skip) {
    for (int r = 1; r < n; ++r)
        for (int i = 0; i < m; ++i) {
            T v = x[m * r + i];
            for (int c = 0; c < r; ++c) v = fma_rn(-x[m * c + i], lu[n * r + c], v);
            x[m * r + i] = v;
        }
    for (int r = n - 1; r >= skip; --r) {
        const T k = lu[n * r + r];
        for (int i = 0; i < m; ++i) {
            T v = x[m * r + i];
            for (int c = n - 1; c > r; --c) v = fma_rn(-x[m * c + i], lu[n * r + c], v);
            x[m * r + i] = div_rn(v, k);
        }
    }
}

template <typename T>
__device__ void load_local(const T* g, Addr<T> at, int64_t s, bool rm, int rows, int cols, T* a) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) a[cols * r + c] = g[at(s, midx(rm, rows, cols, r, c))];
}
template <typename T>
__device__ void store_local(T* g, Addr<T> at, int64_t s, bool rm, int rows, int cols, const T* a) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) g[at(s, midx(rm, rows, cols, r, c))] = a[cols * r + c];
}

// ---------------------------------------------------------------- kernels
template <typename T, bool COLPIV>
__global__ void __launch_bounds__(kGenThreads)
g_decomp(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout, int rm) {
    const int64_t s = (int64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (s >= B) return;
    T a[kMaxN * kMaxN];
    uint8_t p[kMaxN];
    Addr<T> at{layout, rows * cols};
    load_local(A, at, s, rm, rows, cols, a);
    bool par;
    const int d = lu_local<T, COLPIV>(a, rows, cols, p, par);
    store_local(A, at, s, rm, rows, cols, a);
    const int np = COLPIV ? cols : rows;
    if (perm) for (int i = 0; i < np; ++i) perm[s * np + i] = p[i];
    if (parity) parity[s] = par ? 1 : 0;
    if (degen) degen[s] = (uint8_t)d;
}

// MODE 0 backsolve_lu, 1 backsolve_plu, 2 apply permutation, 3 fused linear_solve
template <typename T, int MODE>
__global__ void __launch_bounds__(kGenThreads)
g_solve(int n, int m, int64_t B, const T* A, const uint8_t* perm, const T* Bm, T* X, uint8_t* ok, T* LU_out, int skip,
        int layout, int rm) {
    const int64_t s = (int64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (s >= B) return;
    T a[kMaxN * kMaxN];
    T x[kMaxN * GMB_MAX_N];
    uint8_t p[kMaxN];
    Addr<T> atA{layout, n * n}, atX{layout, n * m};
    if (MODE != 2) load_local(A, atA, s, rm, n, n, a);
    if (MODE == 0) {
        load_local(Bm, atX, s, rm, n, m, x);
    } else if (MODE == 1 || MODE == 2) {
        for (int r = 0; r < n; ++r) {
            const int src = perm[s * n + r];
            for (int i = 0; i < m; ++i) x[m * r + i] = Bm[atX(s, midx(rm, n, m, src, i))];
        }
    } else {
        bool par;
        const int d = lu_local<T, false>(a, n, n, p, par);
        if (LU_out) store_local(LU_out, atA, s, rm, n, n, a);
        if (ok) ok[s] = d == 0 ? 1 : 0;
        if (d > 0) {                                    // LUDecomp.h:391-393: x untouched
            if (X != Bm)
                for (int e = 0; e < n * m; ++e) X[atX(s, e)] = Bm[atX(s, e)];
            return;
        }
        for (int r = 0; r < n; ++r)
            for (int i = 0; i < m; ++i) x[m * r + i] = Bm[atX(s, midx(rm, n, m, p[r], i))];
    }
    if (MODE != 2) backsolve_local(a, n, m, x, skip);
    store_local(X, atX, s, rm, n, m, x);
}

template <typename T, bool SOLVE>
__global__ void __launch_bounds__(kGenThreads)
g_cholesky(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm) {
    const int64_t s = (int64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (s >= B) return;
    T a[kMaxN * kMaxN];
    Addr<T> atA{layout, n * n}, atX{layout, n * m};
    load_local(A, atA, s, rm, n, n, a);
    bool good = true;
    for (int r = 0; r < n; ++r)                         // Cholesky.h:40-56
        for (int c = 0; c <= r; ++c) {
            T v = a[n * r + c];
            for (int k = 0; k < c; ++k) v = fma_rn(-a[n * r + k], a[n * c + k], v);
            if (r == c) {
                good = good && (v > (T)0);
                v = sqrt_rn(v);
            } else {
                v = div_rn(v, a[n * c + c]);
                a[n * c + r] = (T)0;
            }
            a[n * r + c] = v;
        }
    if (ok) ok[s] = good ? 1 : 0;
    if (L_out) store_local(L_out, atA, s, rm, n, n, a);
    if (SOLVE) {
        T x[kMaxN * GMB_MAX_N];
        load_local(Bm, atX, s, rm, n, m, x);
        for (int r = 0; r < n; ++r) {
            const T k = a[n * r + r];
            for (int i = 0; i < m; ++i) {
                T v = x[m * r + i];
                for (int c = 0; c < r; ++c) v = fma_rn(-x[m * c + i], a[n * r + c], v);
                x[m * r + i] = div_rn(v, k);
            }
        }
        for (int r = n - 1; r >= 0; --r) {
            const T k = a[n * r + r];
            for (int i = 0; i < m; ++i) {
                T v = x[m * r + i];
                for (int c = n - 1; c > r; --c) v = fma_rn(-x[m * c + i], a[n * c + r], v);
                x[m * r + i] = div_rn(v, k);
            }
        }
        store_local(X, atX, s, rm, n, m, x);
    }
}

// inv for n >= 5: MatrixInv.h:243-249 + 219-231.  The working copy has the DESTINATION's layout
// and is then treated as row-major by invNxN; out(i, p[i]) = 1; backsolve with m = n.
template <typename T>
__global__ void __launch_bounds__(kGenThreads)
g_inv(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm) {
    const int64_t s = (int64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (s >= B) return;
    T a[kMaxN * kMaxN];
    T o[kMaxN * kMaxN];
    uint8_t p[kMaxN];
    Addr<T> at{layout, n * n};
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) a[midx(dst_rm, n, n, r, c)] = A[at(s, midx(src_rm, n, n, r, c))];
    bool par;
    const int d = lu_local<T, false>(a, n, n, p, par);
    if (ok) ok[s] = d == 0 ? 1 : 0;
    if (d > 0) return;                                  // out untouched
    for (int e = 0; e < n * n; ++e) o[e] = (T)0;
    for (int i = 0; i < n; ++i) o[n * i + p[i]] = (T)1;
    backsolve_local(a, n, n, o, 0);
    for (int e = 0; e < n * n; ++e) Ainv[at(s, e)] = o[e];
}

// linear_solve(Vec) for n >= 5 (LUDecomp.h:428-439) is g_solve<T,3> with rm = 0: the local
// routines index logical (r, c) and load_local / store_local map the column-major storage.

// orthogonal (Orthogonal.h:73-98, general N; N = 2, 3 follow the specialisations :101-110) and
// nullspace (Orthogonal.h:136-178).  Vectors are rows of a row-major matrix; nb = number of bases.
template <typename T>
__device__ void orthogonal_local(const T* g, Addr<T> atV, int64_t s, int n, T* o) {
    if (n == 2) {
        o[0] = -g[atV(s, 1)];
        o[1] = g[atV(s, 0)];
        return;
    }
    if (n == 3) {
        const T ax = g[atV(s, 0)], ay = g[atV(s, 1)], az = g[atV(s, 2)];
        const T bx = g[atV(s, 3)], by = g[atV(s, 4)], bz = g[atV(s, 5)];
        o[0] = dop(ay, bz, az, by);
        o[1] = dop(az, bx, ax, bz);
        o[2] = dop(ax, by, ay, bx);
        return;
    }
    T a[kMaxN * kMaxN];
    uint8_t p[kMaxN];
    for (int e = 0; e < n * (n - 1); ++e) a[e] = g[atV(s, e)];
    for (int e = 0; e < n; ++e) o[e] = (T)0;
    bool par;
    if (lu_local<T, false>(a, n - 1, n, p, par) > 0) return;
    o[n - 1] = (T)1;
    for (int r = n - 2; r >= 0; --r) {
        T acc = o[r];
        for (int c = n - 1; c > r; --c) acc = fma_rn(-o[c], a[n * r + c], acc);
        o[r] = div_rn(acc, a[n * r + r]);
    }
}

template <typename T>
__global__ void __launch_bounds__(kGenThreads)
g_orthogonal(int n, int64_t B, const T* V, T* out, int layout) {
    const int64_t s = (int64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (s >= B) return;
    T o[kMaxN];
    orthogonal_local<T>(V, Addr<T>{layout, n * (n - 1)}, s, n, o);
    Addr<T> atO{layout, n};
    for (int e = 0; e < n; ++e) out[atO(s, e)] = o[e];
}

template <typename T>
__global__ void __launch_bounds__(kGenThreads)
g_nullspace(int N, int nb, int64_t B, const T* bases, T* null, uint8_t* ok, int layout) {
    const int64_t s = (int64_t)blockIdx.x * kGenThreads + threadIdx.x;
    if (s >= B) return;
    const int nn = N - nb;
    Addr<T> atB{layout, N * nb}, atN{layout, N * nn};
    if (nn == 1) {                                          // :138-142
        T o[kMaxN];
        orthogonal_local<T>(bases, atB, s, N, o);
        for (int e = 0; e < N; ++e) null[atN(s, e)] = o[e];
        if (ok) ok[s] = 1;
        return;
    }
    T a[kMaxN * kMaxN];
    uint8_t p[kMaxN];
    for (int e = 0; e < N * nb; ++e) a[e] = bases[atB(s, e)];
    bool par;
    const int d = lu_local<T, false>(a, nb, N, p, par);
    if (ok) ok[s] = d == 0 ? 1 : 0;
    if (d > 0) {                                            // :161: the zero-filled null basis is returned
        for (int e = 0; e < N * nn; ++e) null[atN(s, e)] = (T)0;
        return;
    }
    T x[kMaxN];
    for (int b = 0; b < nn; ++b) {
        for (int e = 0; e < N; ++e) x[e] = (T)0;
        x[b + nb] = (T)1;
        for (int r = nb - 1; r >= 0; --r) {
            T xi = -a[N * r + (nb + b)];
            for (int c = r + 1; c < nb; ++c) xi = fma_rn(-a[N * r + c], x[c], xi);   // :171
            x[r] = div_rn(xi, a[N * r + r]);
        }
        for (int e = 0; e < N; ++e) null[atN(s, N * b + e)] = x[e];
    }
}

// ---------------------------------------------------------------- residual statistics
template <typename T>
__global__ void __launch_bounds__(128)
g_residual(int n, int64_t B, const T* A, const T* X, const T* Bm, const uint8_t* ok, double* stats, int layout, int rm) {
    const int64_t s = (int64_t)blockIdx.x * 128 + threadIdx.x;
    double res = 0.0;
    double failed = 0.0, nonfinite = 0.0;
    bool counted = false;
    if (s < B) {
        if (ok != nullptr && ok[s] == 0) {
            failed = 1.0;
        } else {
            Addr<T> atA{layout, n * n}, atX{layout, n};
            for (int r = 0; r < n; ++r) {
                double acc = -(double)Bm[atX(s, r)];
                for (int c = 0; c < n; ++c) acc = fma((double)A[atA(s, midx(rm, n, n, r, c))], (double)X[atX(s, c)], acc);
                res = fmax(res, fabs(acc));
                if (!isfinite(acc)) nonfinite = 1.0;
            }
            counted = nonfinite == 0.0;
        }
    }
    double sum = counted ? res : 0.0;
    double mx = counted ? res : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        failed += __shfl_xor_sync(0xffffffffu, failed, o);
        nonfinite += __shfl_xor_sync(0xffffffffu, nonfinite, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sum != 0.0) atomicAdd(&stats[0], sum);
        if (mx != 0.0) atomicMax(reinterpret_cast<unsigned long long*>(&stats[1]), (unsigned long long)__double_as_longlong(mx));
        if (failed != 0.0) atomicAdd(&stats[2], failed);
        if (nonfinite != 0.0) atomicAdd(&stats[3], nonfinite);
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
g_checksum(const uint8_t* data, int64_t nbytes, int64_t word_offset, unsigned long long* out) {
    const int64_t nwords = (nbytes + 7) / 8;
    uint64_t acc = 0;
    for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * 256) {
        uint64_t v = 0;
        const int64_t b0 = w * 8;
        if (b0 + 8 <= nbytes && (reinterpret_cast<uintptr_t>(data) & 7) == 0) {
            v = *reinterpret_cast<const uint64_t*>(data + b0);
        } else {
            for (int k = 0; k < 8; ++k)
                if (b0 + k < nbytes) v |= (uint64_t)data[b0 + k] << (8 * k);
        }
        acc += mix64(v + 0x9e3779b97f4a7c15ull * (uint64_t)(word_offset + w + 1));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out, (unsigned long long)acc);
}

// ---------------------------------------------------------------- AoS <-> SoA through shared memory
// One CTA per tile of W systems.  AoS side: the tile is W*E contiguous elements, moved with
// coalesced 16-byte accesses.  SoA side: one 16-byte access per lane per element plane.
// Shared rows are padded to an odd element count so both access patterns are conflict-free.
template <typename T, bool TO_SOA>
__global__ void __launch_bounds__(128)
g_transpose(int E, int64_t B, const T* __restrict__ src, T* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sm = reinterpret_cast<T*>(smem_raw);
    constexpr int V = soa_v<T>();
    constexpr int W = soa_w<T>();
    using VT = typename VecOf<T>::type;
    const int EP = E | 1;
    const int64_t tile = blockIdx.x;
    const int64_t s_base = tile * W;
    const int64_t nsys = (B - s_base) < W ? (B - s_base) : W;          // systems valid in this tile
    const int64_t aos_elems = nsys * E;
    const int tot = W * E;
    // the AoS side may be an array of objects that is only element-aligned: scalar accesses there (warp-uniform test)
    const bool al16 = (reinterpret_cast<uintptr_t>(TO_SOA ? src + s_base * E : dst + s_base * E) & 15) == 0;
    if (TO_SOA) {
        const T* a = src + s_base * E;
        for (int g = threadIdx.x * V; g < tot; g += 128 * V) {
            T t[V];
            if (al16 && g + V <= aos_elems) {
                unpack(ldg_stream(reinterpret_cast<const VT*>(a + g)), t);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) t[v] = (g + v < aos_elems) ? a[g + v] : (T)0;
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int sys = (g + v) / E, e = (g + v) - sys * E;
                sm[sys * EP + e] = t[v];
            }
        }
        __syncthreads();
        T* d = dst + tile * (int64_t)tot;
        for (int i = threadIdx.x; i < E * 32; i += 128) {               // (plane e, lane l)
            const int e = i >> 5, l = i & 31;
            T t[V];
#pragma unroll
            for (int v = 0; v < V; ++v) t[v] = sm[(l * V + v) * EP + e];
            reinterpret_cast<VT*>(d + e * W)[l] = pack(t);
        }
    } else {
        const T* a = src + tile * (int64_t)tot;
        for (int i = threadIdx.x; i < E * 32; i += 128) {
            const int e = i >> 5, l = i & 31;
            T t[V];
            unpack(ldg_stream(reinterpret_cast<const VT*>(a + e * W) + l), t);
#pragma unroll
            for (int v = 0; v < V; ++v) sm[(l * V + v) * EP + e] = t[v];
        }
        __syncthreads();
        T* d = dst + s_base * E;
        for (int g = threadIdx.x * V; g < tot; g += 128 * V) {
            T t[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int sys = (g + v) / E, e = (g + v) - sys * E;
                t[v] = sm[sys * EP + e];
            }
            if (al16 && g + V <= aos_elems) {
                *reinterpret_cast<VT*>(d + g) = pack(t);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (g + v < aos_elems) d[g + v] = t[v];
            }
        }
    }
}

// ---------------------------------------------------------------- launchers
template <typename T>
int generic_decomp(int rows, int cols, int64_t B, T* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout,
                   int rm, bool colpiv, cudaStream_t st) {
    if (rows < 1 || cols < 1 || rows > kMaxN || cols > kMaxN) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, kGenThreads);
    if (grid == 0) return GMB_OK;
    if (colpiv) g_decomp<T, true><<<grid, kGenThreads, 0, st>>>(rows, cols, B, A, perm, parity, degen, layout, rm);
    else g_decomp<T, false><<<grid, kGenThreads, 0, st>>>(rows, cols, B, A, perm, parity, degen, layout, rm);
    return after_launch();
}

template <typename T>
int generic_solve(int mode, int n, int m, int64_t B, const T* A, const uint8_t* perm, const T* Bm, T* X, uint8_t* ok,
                  T* LU_out, int skip, int layout, int rm, cudaStream_t st) {
    if (n < 1 || n > kMaxN || m < 1 || m > GMB_MAX_N) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, kGenThreads);
    if (grid == 0) return GMB_OK;
    switch (mode) {
        case 0: g_solve<T, 0><<<grid, kGenThreads, 0, st>>>(n, m, B, A, perm, Bm, X, ok, LU_out, skip, layout, rm); break;
        case 1: g_solve<T, 1><<<grid, kGenThreads, 0, st>>>(n, m, B, A, perm, Bm, X, ok, LU_out, skip, layout, rm); break;
        case 2: g_solve<T, 2><<<grid, kGenThreads, 0, st>>>(n, m, B, A, perm, Bm, X, ok, LU_out, skip, layout, rm); break;
        default: g_solve<T, 3><<<grid, kGenThreads, 0, st>>>(n, m, B, A, perm, Bm, X, ok, LU_out, skip, layout, rm); break;
    }
    return after_launch();
}

template <typename T>
int generic_cholesky(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm,
                     bool solve, cudaStream_t st) {
    if (n < 1 || n > kMaxN || (solve && (m < 1 || m > GMB_MAX_N))) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, kGenThreads);
    if (grid == 0) return GMB_OK;
    if (solve) g_cholesky<T, true><<<grid, kGenThreads, 0, st>>>(n, m, B, A, Bm, X, ok, L_out, layout, rm);
    else g_cholesky<T, false><<<grid, kGenThreads, 0, st>>>(n, m, B, A, Bm, X, ok, L_out, layout, rm);
    return after_launch();
}

template <typename T>
int generic_inv(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm, cudaStream_t st) {
    if (n < 1 || n > kMaxN) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, kGenThreads);
    if (grid == 0) return GMB_OK;
    g_inv<T><<<grid, kGenThreads, 0, st>>>(n, B, A, Ainv, ok, layout, src_rm, dst_rm);
    return after_launch();
}

template <typename T>
int generic_orthogonal(int n, int64_t B, const T* V, T* out, int layout, cudaStream_t st) {
    if (n < 2 || n > kMaxN) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, kGenThreads);
    if (grid == 0) return GMB_OK;
    g_orthogonal<T><<<grid, kGenThreads, 0, st>>>(n, B, V, out, layout);
    return after_launch();
}

template <typename T>
int generic_nullspace(int N, int nb, int64_t B, const T* bases, T* null, uint8_t* ok, int layout, cudaStream_t st) {
    if (N < 3 || N > kMaxN || nb < 1 || nb >= N) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, kGenThreads);
    if (grid == 0) return GMB_OK;
    g_nullspace<T><<<grid, kGenThreads, 0, st>>>(N, nb, B, bases, null, ok, layout);
    return after_launch();
}

template <typename T>
int residual_stats(int n, int64_t B, const T* A, const T* X, const T* Bm, const uint8_t* ok, double* stats, int layout,
                   int rm, cudaStream_t st) {
    if (n < 1 || n > kMaxN) return GMB_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(B, 128);
    if (grid == 0) return GMB_OK;
    g_residual<T><<<grid, 128, 0, st>>>(n, B, A, X, Bm, ok, stats, layout, rm);
    return after_launch();
}

int checksum_bytes(const void* data, int64_t nbytes, int64_t word_offset, uint64_t* out, cudaStream_t st) {
    if (nbytes <= 0) return GMB_OK;
    const int64_t nwords = (nbytes + 7) / 8;
    int64_t grid = ceil_div(nwords, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    g_checksum<<<(unsigned)grid, 256, 0, st>>>(static_cast<const uint8_t*>(data), nbytes, word_offset,
                                                reinterpret_cast<unsigned long long*>(out));
    return after_launch();
}

template <typename T>
int transpose_batch(bool to_soa, int E, int64_t B, const T* src, T* dst, cudaStream_t st) {
    if (E < 1 || E > 1024) return GMB_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(to_soa ? dst : src) % 16 != 0) return GMB_ERR_INVALID_ARG;   // the container is aligned by construction
    constexpr int W = soa_w<T>();
    const unsigned grid = (unsigned)ceil_div(B, W);
    if (grid == 0) return GMB_OK;
    const size_t smem = (size_t)W * (E | 1) * sizeof(T);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(g_transpose<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(g_transpose<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (to_soa) g_transpose<T, true><<<grid, 128, smem, st>>>(E, B, src, dst);
    else g_transpose<T, false><<<grid, 128, smem, st>>>(E, B, src, dst);
    return after_launch();
}

#define INSTANTIATE(T)                                                                                                  \
    template int generic_decomp<T>(int, int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, bool, cudaStream_t);  \
    template int generic_solve<T>(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, uint8_t*, T*, int,    \
                                  int, int, cudaStream_t);                                                              \
    template int generic_cholesky<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool,           \
                                     cudaStream_t);                                                                     \
    template int generic_inv<T>(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);                     \
    template int residual_stats<T>(int, int64_t, const T*, const T*, const T*, const uint8_t*, double*, int, int,       \
                                   cudaStream_t);                                                                       \
    template int transpose_batch<T>(bool, int, int64_t, const T*, T*, cudaStream_t);                                     \
    template int generic_orthogonal<T>(int, int64_t, const T*, T*, int, cudaStream_t);                                  \
    template int generic_nullspace<T>(int, int, int64_t, const T*, T*, uint8_t*, int, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace gmb
