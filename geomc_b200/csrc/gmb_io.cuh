// gmb_io.cuh -- how a thread reaches its systems in HBM.
//
// Two batch layouts (include/geomc_b200.h):
//
//   GMB_LAYOUT_AOS   the reference's per-object storage: system s, element e at  p[s*E + e]
//                    (SimpleMatrix.h:145: contiguous T[M*N] per object; Vec: T[N], VecBase.h:112).
//
//   GMB_LAYOUT_SOA   the SoA-interleaved batch container: the batch is cut into tiles of
//                    W = 32 * (16 / sizeof(T)) systems (128 float / 64 double); inside a tile,
//                    element e of all W systems is contiguous:   p[(tile*E + e)*W + (s % W)].
//                    A warp reads one element plane of one tile with ONE 128-bit load per lane
//                    (lane l owns the V = 16/sizeof(T) consecutive systems l*V .. l*V+V-1):
//                    512 contiguous bytes per warp instruction, every sector fully used.
//                    Buffers are allocated in whole tiles (gmb_soa_elems()).
//
// Loads go through the non-coherent path with L1::no_allocate (streamed once); stores are plain
// 128-bit STG.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gmb {

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };

template <typename T> constexpr int soa_v() { return 16 / (int)sizeof(T); }
template <typename T> constexpr int soa_w() { return 32 * soa_v<T>(); }

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
// 256-bit global loads (sm_100a, LDG.E.256): one whole 32-byte sector per lane and instruction.  A lane that walks its
// own AoS row with 128-bit loads touches every sector twice, half of it each time, and with L1::no_allocate the second
// half comes from L2 again.
__device__ __forceinline__ void ldg_stream256(const float* p, float (&o)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "l"(p));
}
__device__ __forceinline__ void ldg_stream256(const double* p, double (&o)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}
__device__ __forceinline__ void unpack(const float4& v, float (&o)[4])  { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4  pack(const float (&o)[4])  { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 pack(const double (&o)[2]) { return make_double2(o[0], o[1]); }

// ---- SoA tile access: q[v][e] for the V systems owned by this lane -------------------------
// One row of N elements of a row-major AoS record into registers (dst[0..N-1], static indices after unrolling):
// 256-bit loads when the row is a multiple of 32 bytes and the batch is 32-byte aligned (`al32`, warp-uniform),
// 128-bit loads otherwise.  N * sizeof(T) must be a multiple of 16.
template <typename T, int N>
__device__ __forceinline__ void ldg_row(const T* __restrict__ p, T* dst, bool al32) {
    constexpr int V = 16 / (int)sizeof(T);
    static_assert((N * (int)sizeof(T)) % 16 == 0, "row must be a multiple of 16 bytes");
    using VT = typename VecOf<T>::type;
    if constexpr ((N * (int)sizeof(T)) % 32 == 0) {
        if (al32) {
#pragma unroll
            for (int q = 0; q < N / (2 * V); ++q) {
                T t[2 * V];
                ldg_stream256(p + q * 2 * V, t);
#pragma unroll
                for (int v = 0; v < 2 * V; ++v) dst[q * 2 * V + v] = t[v];
            }
            return;
        }
    }
    const VT* src = reinterpret_cast<const VT*>(p);
#pragma unroll
    for (int q = 0; q < N / V; ++q) {
        T t[V];
        unpack(ldg_stream(src + q), t);
#pragma unroll
        for (int v = 0; v < V; ++v) dst[q * V + v] = t[v];
    }
}

template <typename T, int E>
__device__ __forceinline__ void soa_load(const T* __restrict__ base, int64_t tile, int lane, T (&q)[soa_v<T>()][E]) {
    constexpr int V = soa_v<T>();
    constexpr int W = soa_w<T>();
    using VT = typename VecOf<T>::type;
    const VT* p = reinterpret_cast<const VT*>(base + tile * (int64_t)(E * W)) + lane;
    VT tmp[E];
#pragma unroll
    for (int e = 0; e < E; ++e) tmp[e] = ldg_stream(p + e * 32);      // all loads in flight first
#pragma unroll
    for (int e = 0; e < E; ++e) {
        T t[V];
        unpack(tmp[e], t);
#pragma unroll
        for (int v = 0; v < V; ++v) q[v][e] = t[v];
    }
}

template <typename T, int E>
__device__ __forceinline__ void soa_store(T* __restrict__ base, int64_t tile, int lane, const T (&q)[soa_v<T>()][E]) {
    constexpr int V = soa_v<T>();
    constexpr int W = soa_w<T>();
    using VT = typename VecOf<T>::type;
    VT* p = reinterpret_cast<VT*>(base + tile * (int64_t)(E * W)) + lane;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        T t[V];
#pragma unroll
        for (int v = 0; v < V; ++v) t[v] = q[v][e];
        p[e * 32] = pack(t);
    }
}

// ---- AoS access: one system per thread, widest vector the record size and alignment allow ---
template <typename T, int E>
__device__ __forceinline__ void aos_load(const T* __restrict__ base, int64_t s, T (&q)[E]) {
    const T* p = base + s * (int64_t)E;
    constexpr int BYTES = E * (int)sizeof(T);
    if constexpr (BYTES % 32 == 0) {
        // 256-bit loads (one whole sector per lane and instruction) when the batch is 32-byte aligned
        if ((reinterpret_cast<uintptr_t>(base) & 31) == 0) {
            constexpr int V2 = 32 / (int)sizeof(T);
            T tmp2[E / V2][V2];
#pragma unroll
            for (int i = 0; i < E / V2; ++i) ldg_stream256(p + i * V2, tmp2[i]);
#pragma unroll
            for (int i = 0; i < E / V2; ++i)
#pragma unroll
                for (int v = 0; v < V2; ++v) q[i * V2 + v] = tmp2[i][v];
            return;
        }
    }
    if constexpr (BYTES % 16 == 0) {
        using VT = typename VecOf<T>::type;
        constexpr int V = soa_v<T>();
        const VT* pv = reinterpret_cast<const VT*>(p);
        VT tmp[E / V];
#pragma unroll
        for (int i = 0; i < E / V; ++i) tmp[i] = ldg_stream(pv + i);
#pragma unroll
        for (int i = 0; i < E / V; ++i) {
            T t[V];
            unpack(tmp[i], t);
#pragma unroll
            for (int v = 0; v < V; ++v) q[i * V + v] = t[v];
        }
    } else if constexpr (sizeof(T) == 4 && BYTES % 8 == 0) {
        const float2* pv = reinterpret_cast<const float2*>(p);
#pragma unroll
        for (int i = 0; i < E / 2; ++i) {
            float2 t = __ldg(pv + i);
            q[2 * i] = t.x;
            q[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) q[e] = __ldg(p + e);
    }
}

template <typename T, int E>
__device__ __forceinline__ void aos_store(T* __restrict__ base, int64_t s, const T (&q)[E]) {
    T* p = base + s * (int64_t)E;
    constexpr int BYTES = E * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
        using VT = typename VecOf<T>::type;
        constexpr int V = soa_v<T>();
        VT* pv = reinterpret_cast<VT*>(p);
#pragma unroll
        for (int i = 0; i < E / V; ++i) {
            T t[V];
#pragma unroll
            for (int v = 0; v < V; ++v) t[v] = q[i * V + v];
            pv[i] = pack(t);
        }
    } else if constexpr (sizeof(T) == 4 && BYTES % 8 == 0) {
        float2* pv = reinterpret_cast<float2*>(p);
#pragma unroll
        for (int i = 0; i < E / 2; ++i) pv[i] = make_float2(q[2 * i], q[2 * i + 1]);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) p[e] = q[e];
    }
}

// ---- AoS access staged through shared memory (warp-private transpose) -----------------------
// A warp owns 32 consecutive records of R = E*sizeof(T) bytes: one contiguous, 16-byte aligned
// range of 32*R bytes.  Instead of every thread fetching its own record (R-byte stride between
// lanes: half-used sectors, 16-32 L1 wavefronts per instruction), the warp moves the range with
// fully coalesced 128-bit accesses (512 contiguous bytes per instruction) through a warp-private
// shared-memory buffer, and each lane then reads/writes its record there.  For power-of-two
// records of 32 / 64 / 128 bytes the 16-byte chunks are XOR-swizzled so that both sides are
// bank-conflict free; other record sizes (36, 48, 72, 96 bytes) are conflict free (or 2-way at
// worst) in linear order.  Records under 32 bytes are already coalesced per thread and are not staged.
template <int R>
struct StageCfg {
    static constexpr bool vec16 = (R % 16 == 0);
    static constexpr int CH = R / 16;                        // 16-byte chunks per record (if vec16)
    static constexpr bool swz = vec16 && (CH == 2 || CH == 4 || CH == 8 || CH == 16 || CH == 32);
    static constexpr bool use = R >= 32;
    static constexpr int warp_bytes = 32 * R;
    __device__ static __forceinline__ int chunk(int t, int j) {           // position of chunk j of record t
        if constexpr (swz && CH <= 8) return t * CH + (j ^ ((t / (8 / CH)) & (CH - 1)));
        else if constexpr (swz) return t * CH + (j ^ (t & 7));            // records of 256 / 512 bytes: rotate inside groups of 8 chunks
        else return t * CH + j;
    }
};

template <typename T, int E>
__device__ __forceinline__ unsigned char* stage_buffer() {
    constexpr int R = E * (int)sizeof(T);
    __shared__ __align__(16) unsigned char buf[4][StageCfg<R>::warp_bytes];
    return buf[(threadIdx.x >> 5) & 3];
}

// all 32 lanes of the warp call this; record t = lane; `full` = all 32 records exist
template <typename T, int E>
__device__ __forceinline__ void aos_load_staged(const T* __restrict__ base, int64_t s_warp0, int lane, T (&q)[E],
                                                unsigned char* sm = stage_buffer<T, E>()) {
    constexpr int R = E * (int)sizeof(T);
    using Cfg = StageCfg<R>;
    const uint4* g = reinterpret_cast<const uint4*>(base + s_warp0 * (int64_t)E);
    constexpr int NCH = 2 * R;                               // 32 * R / 16
#pragma unroll
    for (int c0 = 0; c0 < NCH; c0 += 32) {
        const int c = c0 + lane;
        if (NCH % 32 == 0 || c < NCH) {
            uint4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g + c));
            int pos = c;
            if constexpr (Cfg::swz) pos = Cfg::chunk(c / Cfg::CH, c % Cfg::CH);
            *reinterpret_cast<uint4*>(sm + pos * 16) = v;
        }
    }
    __syncwarp();
    if constexpr (Cfg::vec16) {
        using VT = typename VecOf<T>::type;
        constexpr int V = soa_v<T>();
#pragma unroll
        for (int j = 0; j < Cfg::CH; ++j) {
            T t[V];
            unpack(*reinterpret_cast<const VT*>(sm + Cfg::chunk(lane, j) * 16), t);
#pragma unroll
            for (int v = 0; v < V; ++v) q[j * V + v] = t[v];
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) q[e] = *reinterpret_cast<const T*>(sm + lane * R + e * (int)sizeof(T));
    }
    __syncwarp();                                            // the buffer may be reused right away
}

// The same ingest in two halves, with the global -> shared copy done by cp.async (LDGSTS: no register staging, and the copies
// of SEVERAL record arrays -- matrix and right-hand sides -- are all in flight before anybody waits: one DRAM latency instead of
// one per array).  aos_stage_issue for every array, then cp_async_wait_all_() + __syncwarp(), then aos_stage_read for each.
template <typename T, int E>
__device__ __forceinline__ void aos_stage_issue(const T* __restrict__ base, int64_t s_warp0, int lane, unsigned char* sm) {
    constexpr int R = E * (int)sizeof(T);
    using Cfg = StageCfg<R>;
    const uint4* g = reinterpret_cast<const uint4*>(base + s_warp0 * (int64_t)E);
    constexpr int NCH = 2 * R;                               // 32 * R / 16
#pragma unroll
    for (int c0 = 0; c0 < NCH; c0 += 32) {
        const int c = c0 + lane;
        if (NCH % 32 == 0 || c < NCH) {
            int pos = c;
            if constexpr (Cfg::swz) pos = Cfg::chunk(c / Cfg::CH, c % Cfg::CH);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + pos * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + c) : "memory");
        }
    }
}
__device__ __forceinline__ void cp_async_wait_all_() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// The SoA counterpart: the warp's 32 consecutive systems (s_warp0 a multiple of 32, all inside one tile of W systems) are a
// 32-element segment of each of the E element planes.  The segments are copied global -> shared with cp.async as whole 16-byte
// chunks (a plane segment is 128 / 256 contiguous bytes), plane e landing at sm + e * 32 * sizeof(T): element e of the lane's
// system is then sm[e * 32 + lane] -- bank = lane, conflict-free for static and run-time e alike.  No registers are held while
// the data travel, and the copies of several arrays are in flight together (same protocol as aos_stage_issue).
template <typename T, int E>
__device__ __forceinline__ void soa_stage_issue(const T* __restrict__ base, int64_t s_warp0, int lane, unsigned char* sm) {
    constexpr int W = soa_w<T>();
    constexpr int V = soa_v<T>();                             // elements per 16-byte chunk
    constexpr int CPP = 32 / V;                               // chunks per plane segment
    constexpr int NCH = E * CPP;
    const T* g = base + (s_warp0 / W) * (int64_t)E * W + (s_warp0 % W);
#pragma unroll
    for (int c0 = 0; c0 < NCH; c0 += 32) {
        const int c = c0 + lane;
        if (NCH % 32 == 0 || c < NCH) {
            const int e = c / CPP, w = c % CPP;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + c * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + (int64_t)e * W + w * V) : "memory");
        }
    }
}
template <typename T>
__device__ __forceinline__ T soa_stage_elem(int lane, int e, const unsigned char* sm) {
    return reinterpret_cast<const T*>(sm)[e * 32 + lane];
}
template <typename T, int E>
__device__ __forceinline__ void aos_stage_read(int lane, T (&q)[E], const unsigned char* sm) {
    constexpr int R = E * (int)sizeof(T);
    using Cfg = StageCfg<R>;
    if constexpr (Cfg::vec16) {
        using VT = typename VecOf<T>::type;
        constexpr int V = soa_v<T>();
#pragma unroll
        for (int j = 0; j < Cfg::CH; ++j) {
            T t[V];
            unpack(*reinterpret_cast<const VT*>(sm + Cfg::chunk(lane, j) * 16), t);
#pragma unroll
            for (int v = 0; v < V; ++v) q[j * V + v] = t[v];
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) q[e] = *reinterpret_cast<const T*>(sm + lane * R + e * (int)sizeof(T));
    }
}

template <typename T, int E>
__device__ __forceinline__ void aos_store_staged(T* __restrict__ base, int64_t s_warp0, int lane, const T (&q)[E],
                                                 unsigned char* sm = stage_buffer<T, E>()) {
    constexpr int R = E * (int)sizeof(T);
    using Cfg = StageCfg<R>;
    if constexpr (Cfg::vec16) {
        using VT = typename VecOf<T>::type;
        constexpr int V = soa_v<T>();
#pragma unroll
        for (int j = 0; j < Cfg::CH; ++j) {
            T t[V];
#pragma unroll
            for (int v = 0; v < V; ++v) t[v] = q[j * V + v];
            *reinterpret_cast<VT*>(sm + Cfg::chunk(lane, j) * 16) = pack(t);
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) *reinterpret_cast<T*>(sm + lane * R + e * (int)sizeof(T)) = q[e];
    }
    __syncwarp();
    uint4* g = reinterpret_cast<uint4*>(base + s_warp0 * (int64_t)E);
    constexpr int NCH = 2 * R;
#pragma unroll
    for (int c0 = 0; c0 < NCH; c0 += 32) {
        const int c = c0 + lane;
        if (NCH % 32 == 0 || c < NCH) {
            int pos = c;
            if constexpr (Cfg::swz) pos = Cfg::chunk(c / Cfg::CH, c % Cfg::CH);
            g[c] = *reinterpret_cast<const uint4*>(sm + pos * 16);
        }
    }
    __syncwarp();
}

// ---- warp-cooperative staging of whole systems for the sub-warp / row-distributed kernels ----------
// A warp that works on NSYS = 32 / G consecutive AoS systems of E elements owns one contiguous,
// 16-byte aligned range of NSYS * E elements.  Per-lane element loads from it would touch up to 32
// sectors per instruction (stride E); instead the range is copied with coalesced 128-bit accesses into
// shared memory, system q at q * (E | 1) elements (odd stride: lanes reading the same element of
// different systems hit different banks), and the lanes pick their elements there.  The same buffer
// serves the way back for in-place / matrix outputs.
template <int E> constexpr int stage_stride() { return E | 1; }

template <typename T, int E, int NSYS>
__device__ __forceinline__ void warp_stage_in(const T* __restrict__ g, T* sm, int lane) {
    constexpr int S = stage_stride<E>();
    constexpr int V = soa_v<T>();
    constexpr int TOT = NSYS * E;
    static_assert(TOT % V == 0, "a warp's range must be a whole number of 16-byte chunks");
    using VT = typename VecOf<T>::type;
    const VT* gv = reinterpret_cast<const VT*>(g);
    for (int i = lane; i < TOT / V; i += 32) {
        T t[V];
        unpack(ldg_stream(gv + i), t);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = i * V + v, q = idx / E;
            sm[q * S + (idx - q * E)] = t[v];
        }
    }
    __syncwarp();
}

template <typename T, int E, int NSYS>
__device__ __forceinline__ void warp_stage_out(T* __restrict__ g, const T* sm, int lane) {
    constexpr int S = stage_stride<E>();
    constexpr int V = soa_v<T>();
    constexpr int TOT = NSYS * E;
    using VT = typename VecOf<T>::type;
    VT* gv = reinterpret_cast<VT*>(g);
    __syncwarp();
    for (int i = lane; i < TOT / V; i += 32) {
        T t[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = i * V + v, q = idx / E;
            t[v] = sm[q * S + (idx - q * E)];
        }
        gv[i] = pack(t);
    }
    __syncwarp();
}

}  // namespace gmb
