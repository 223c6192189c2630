// gmb_common.cuh -- host-side helpers shared by the launchers.
#pragma once
#include <atomic>
#include <cstdint>
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>

#include "../../include/geomc_b200.h"

namespace gmb {

extern std::atomic<int64_t> g_launch_count;

inline int after_launch() {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GMB_OK : (int)e;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Call f(std::integral_constant<int, V>{}) for the V in Vs... equal to v; false if none matches.
template <int... Vs, class F>
inline bool dispatch_int(int v, F&& f) {
    bool hit = false;
    ((v == Vs ? (f(std::integral_constant<int, Vs>{}), hit = true) : false), ...);
    return hit;
}

template <class F>
inline void dispatch_bool(bool b, F&& f) {
    if (b) f(std::true_type{});
    else f(std::false_type{});
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute applies to the CURRENT device only: remember, per kernel instantiation, which devices have
// had their dynamic shared-memory limit raised (one bit per device ordinal), so that a process driving several GPUs
// through the C ABI gets the opt-in on every one of them and the common case costs one cudaGetDevice.
template <class K>
inline void raise_dyn_smem(K kern, size_t bytes, std::atomic<uint64_t>& done) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_relaxed) & bit)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        done.fetch_or(bit, std::memory_order_relaxed);
    }
}

// cp.async.bulk.prefetch.L2: `bytes` (a multiple of 16) starting at the 16-byte aligned `p` are brought into L2; no destination
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// option "subprefetch": -1 = one wave ahead (SM count x resident CTAs), 0 = off, > 0 = that many CTAs ahead
extern std::atomic<int> g_sub_prefetch;
inline int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int v = cached[dev & 63].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

}  // namespace gmb
