// gmb_work.cuh -- work decomposition and register <-> memory helpers of the one-thread-per-system
// kernels (gmb_small.cu, gmb_simplex.cu).
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"

namespace gmb {

constexpr int kThreads = 128;

template <typename T, bool SOA>
struct Work {
    static constexpr int V = SOA ? soa_v<T>() : 1;
    static constexpr int W = soa_w<T>();
    // unit = tile index (SOA) or system index (AOS); s0 = first system of this thread
    __device__ static __forceinline__ bool locate(int64_t B, int64_t& unit, int& lane, int64_t& s0) {
        if (SOA) {
            lane = threadIdx.x & 31;
            unit = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
            s0 = unit * W + lane * V;
            return unit * W < B;
        } else {
            // warp-uniform exit: a warp stays whole while any of its 32 systems exists (the staged
            // AoS transposes are warp-cooperative); lanes past the end compute on zeros and store nothing
            lane = threadIdx.x & 31;
            unit = (int64_t)blockIdx.x * kThreads + threadIdx.x;
            s0 = unit;
            return unit - lane < B;
        }
    }
    static int64_t grid(int64_t B) {
        return SOA ? ceil_div(ceil_div(B, W), kThreads / 32) : ceil_div(B, kThreads);
    }
};

// B = batch size (AoS only: guards the ragged last warp)
// Measured on B200 (profiles/): staging pays for STORES of >= 32-byte records (a per-thread store
// leaves every 128-byte line partially written by four separate instructions: 4x4 float inverse
// 0.80 -> 0.97 of the HBM peak), while per-thread LOADS are served well by L1 sector reuse and the
// extra shared-memory round trip costs more than it saves (fused solve 0.94 -> 0.78 when staged).
// The inverse kernel (64 bytes in, 64 bytes out, no other traffic) is the exception: staging both
// sides gives 0.97 against 0.93 with staged stores only, so it passes STAGE_LOAD = true.
template <typename T, int E, bool SOA, bool STAGE_LOAD = false>
__device__ __forceinline__ void load_elems(const T* base, int64_t unit, int lane, int64_t B, T (&q)[Work<T, SOA>::V][E]) {
    if constexpr (SOA) {
        soa_load<T, E>(base, unit, lane, q);
    } else {
        const int64_t s_warp0 = unit - lane;
        if (STAGE_LOAD && StageCfg<E * (int)sizeof(T)>::use && s_warp0 + 32 <= B) {      // warp-uniform
            aos_load_staged<T, E>(base, s_warp0, lane, q[0]);
        } else if (unit < B) {
            aos_load<T, E>(base, unit, q[0]);
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) q[0][e] = (T)0;
        }
    }
}
template <typename T, int E, bool SOA>
__device__ __forceinline__ void store_elems(T* base, int64_t unit, int lane, int64_t B, const T (&q)[Work<T, SOA>::V][E]) {
    if constexpr (SOA) {
        soa_store<T, E>(base, unit, lane, q);
    } else {
        const int64_t s_warp0 = unit - lane;
        if (StageCfg<E * (int)sizeof(T)>::use && s_warp0 + 32 <= B) {
            aos_store_staged<T, E>(base, s_warp0, lane, q[0]);
        } else if (unit < B) {
            aos_store<T, E>(base, unit, q[0]);
        }
    }
}

// K bytes per system, V systems per thread, AoS byte array [B][K]
template <int V, int K>
__device__ __forceinline__ void store_bytes(uint8_t* p, int64_t s0, int64_t B, const uint8_t (&b)[V][K]) {
    if (p == nullptr) return;
    constexpr int TOT = V * K;
    uint8_t* dst = p + s0 * K;
    if constexpr (TOT % 4 == 0) {
        if (s0 + V <= B && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
            uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
            for (int w = 0; w < TOT / 4; ++w) {
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = w * 4 + k;
                    x |= (uint32_t)b[i / K][i % K] << (8 * k);
                }
                d32[w] = x;
            }
            return;
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (s0 + v < B) {
#pragma unroll
            for (int k = 0; k < K; ++k) dst[v * K + k] = b[v][k];
        }
}

template <int V, int K>
__device__ __forceinline__ void load_bytes(const uint8_t* p, int64_t s0, int64_t B, uint8_t (&b)[V][K]) {
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k < K; ++k) b[v][k] = (s0 + v < B) ? p[(s0 + v) * K + k] : (uint8_t)k;
}

// ---- one record per thread, element by element (the kernels that keep ONE system per thread in both
// layouts: gmb_simplex.cu, gmb_xform.cuh).  AoS: `stride` elements between records; SoA: tiles of W
// records, consecutive threads read consecutive 4/8-byte words of one element plane (coalesced).
// rec_*_part moves CNT elements starting at element `off` of records that hold ETOT elements.
// AoS chunks whose byte size is a multiple of 16 move with 128-bit accesses when their address is 16-byte
// aligned (checked per thread; warp-uniform whenever the record size is a multiple of 16 too): a warp-wide
// scalar access to records >= 128 bytes apart costs one L1 wavefront per lane, so this is a 4x (float) / 2x
// (double) cut in LSU work -- measured: per-point AffineTransform::apply on AoS records was wavefront-bound.
template <typename T, int ETOT, int CNT, bool SOA>
__device__ __forceinline__ void rec_load_part(const T* __restrict__ base, int64_t s, int64_t stride, int off, T (&q)[CNT]) {
    if constexpr (SOA) {
        constexpr int W = soa_w<T>();
        const T* p = base + (s / W) * (int64_t)(ETOT * W) + (int64_t)off * W + (s % W);
#pragma unroll
        for (int e = 0; e < CNT; ++e) q[e] = __ldg(p + e * W);
    } else {
        const T* p = base + s * stride + off;
        constexpr int V = soa_v<T>();
        if constexpr (CNT % V == 0) {
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                using VT = typename VecOf<T>::type;
                const VT* pv = reinterpret_cast<const VT*>(p);
                VT tmp[CNT / V];
#pragma unroll
                for (int i = 0; i < CNT / V; ++i) tmp[i] = __ldg(pv + i);
#pragma unroll
                for (int i = 0; i < CNT / V; ++i) {
                    T t[V];
                    unpack(tmp[i], t);
#pragma unroll
                    for (int v = 0; v < V; ++v) q[i * V + v] = t[v];
                }
                return;
            }
        }
#pragma unroll
        for (int e = 0; e < CNT; ++e) q[e] = __ldg(p + e);
    }
}
template <typename T, int ETOT, int CNT, bool SOA>
__device__ __forceinline__ void rec_store_part(T* __restrict__ base, int64_t s, int off, const T (&q)[CNT]) {
    if constexpr (SOA) {
        constexpr int W = soa_w<T>();
        T* p = base + (s / W) * (int64_t)(ETOT * W) + (int64_t)off * W + (s % W);
#pragma unroll
        for (int e = 0; e < CNT; ++e) p[e * W] = q[e];
    } else {
        T* p = base + s * (int64_t)ETOT + off;
        constexpr int V = soa_v<T>();
        if constexpr (CNT % V == 0) {
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                using VT = typename VecOf<T>::type;
                VT* pv = reinterpret_cast<VT*>(p);
#pragma unroll
                for (int i = 0; i < CNT / V; ++i) {
                    T t[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) t[v] = q[i * V + v];
                    pv[i] = pack(t);
                }
                return;
            }
        }
#pragma unroll
        for (int e = 0; e < CNT; ++e) p[e] = q[e];
    }
}
template <typename T, int E, bool SOA>
__device__ __forceinline__ void rec_load(const T* __restrict__ base, int64_t s, int64_t stride, T (&q)[E]) {
    rec_load_part<T, E, E, SOA>(base, s, stride, 0, q);
}
template <typename T, int E, bool SOA>
__device__ __forceinline__ void rec_store(T* __restrict__ base, int64_t s, const T (&q)[E]) {
    rec_store_part<T, E, E, SOA>(base, s, 0, q);
}

template <typename T, int R, int C, bool RM>
__device__ __forceinline__ void unflatten(const T (&q)[R * C], T (&a)[R][C]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) a[r][c] = q[RM ? C * r + c : R * c + r];
}
template <typename T, int R, int C, bool RM>
__device__ __forceinline__ void flatten(const T (&a)[R][C], T (&q)[R * C]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) q[RM ? C * r + c : R * c + r] = a[r][c];
}

}  // namespace gmb
