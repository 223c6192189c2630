// Experiment harness for the speculative fused PLU solve (gmb_rowfast.cuh) against the exact row kernel, n = 9..16:
// same random batch (plus singular, tie-heavy, NaN / Inf, huge and tiny systems sprinkled in), bit-for-bit comparison,
// CUDA-event timings, number of systems that left the fast path.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
//        -I geomc_b200/csrc profiles/exp/rowfast_bench.cu -o profiles/exp/rowfast_bench
//   rowfast_bench [log2 B = 20] [aos|soa] [f32|f64] [nmin nmax]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gmb_rowfast.cuh"

namespace gmb { std::atomic<int64_t> g_launch_count{0}; }
using namespace gmb;

template <typename T>
__global__ void fill(T* p, int64_t n, uint32_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    int v = (int)(z & 0xFFFFF) - (1 << 19);
    p[i] = (T)v * (T)(1.0 / 131072.0);             // exact grid in [-4, 4)
}

// every 997th system gets a special structure (AoS element order; applied before the optional SoA transpose)
template <typename T>
__global__ void spice(T* A, T* b, int n, int64_t B) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 997 + 5;
    if (s >= B) return;
    T* a = A + s * n * n;
    const int kind = (int)((s / 997) % 10);
    switch (kind) {
        case 0: for (int e = 0; e < n * n; ++e) a[e] = 0; break;                                   // all zero
        case 1: for (int e = 0; e < n * n; ++e) a[e] = (e % 3 == 0) ? (T)1 : (T)-1; break;          // ties everywhere
        case 2: for (int r = 0; r < n; ++r) a[r * n] = 0; break;                                    // zero first column
        case 3: for (int c = 0; c < n; ++c) a[n + c] = a[c]; break;                                 // duplicate row
        case 4: a[3 * n + 2] = (T)NAN; break;
        case 5: a[0] = (T)INFINITY; break;
        case 6: for (int e = 0; e < n * n; ++e) a[e] *= (T)1e30; break;                             // large (float: > 2^57)
        case 7: for (int e = 0; e < n * n; ++e) a[e] *= (T)1e-30; break;                            // tiny
        case 8: for (int e = 0; e < n * n; ++e) a[e] = (T)(int)(a[e]); break;                       // small integers: ties + zeros
        case 9: a[n * n - 1] = (T)NAN; b[s * n + 1] = 0; break;
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <class F>
static float time_ms(F&& f, int reps = 5) {
    f();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

template <typename T> int transpose_batch_local(bool to_soa, int E, int64_t B, const T* src, T* dst);
template <typename T>
__global__ void k_tr(bool to_soa, int E, int64_t B, const T* src, T* dst) {
    constexpr int W = soa_w<T>();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * E) return;
    int64_t s = i / E; int e = (int)(i % E);
    int64_t j = (s / W) * (int64_t)E * W + (int64_t)e * W + (s % W);
    if (to_soa) dst[j] = src[i]; else dst[i] = src[j];
}

template <typename T, int N, bool SOA>
static void run_n(int64_t B) {
    const int64_t W = soa_w<T>();
    const int64_t Bp = (B + W - 1) / W * W;
    T *A, *b, *x0, *x, *tmp;
    uint8_t *ok0, *ok;
    CK(cudaMalloc(&A, Bp * N * N * sizeof(T))); CK(cudaMalloc(&b, Bp * N * sizeof(T)));
    CK(cudaMalloc(&x0, Bp * N * sizeof(T))); CK(cudaMalloc(&x, Bp * N * sizeof(T)));
    CK(cudaMalloc(&tmp, Bp * N * N * sizeof(T)));
    CK(cudaMalloc(&ok0, B)); CK(cudaMalloc(&ok, B));
    CK(cudaMemset(A, 0, Bp * N * N * sizeof(T))); CK(cudaMemset(b, 0, Bp * N * sizeof(T)));
    fill<T><<<(unsigned)((B * N * N + 255) / 256), 256>>>(A, B * N * N, 1);
    fill<T><<<(unsigned)((B * N + 255) / 256), 256>>>(b, B * N, 2);
    spice<T><<<(unsigned)((B / 997 + 256) / 256), 256>>>(A, b, N, B);
    if (SOA) {
        k_tr<T><<<(unsigned)((B * N * N + 255) / 256), 256>>>(true, N * N, B, A, tmp);
        CK(cudaMemcpy(A, tmp, Bp * N * N * sizeof(T), cudaMemcpyDeviceToDevice));
        k_tr<T><<<(unsigned)((B * N + 255) / 256), 256>>>(true, N, B, b, tmp);
        CK(cudaMemcpy(b, tmp, Bp * N * sizeof(T), cudaMemcpyDeviceToDevice));
    }
    CK(cudaDeviceSynchronize());
    std::vector<T> hx0(Bp * N), hx(Bp * N);
    std::vector<uint8_t> hok0(B), hok(B);
    const double bytes = (double)B * (N * N + 2 * N) * sizeof(T);

    using P = RowNoLu<T, N>;
    {
        auto kern = k_row_solve<T, N, SOA, true, false, P::G, P::minb, false, P::shfl, P::threads>;
        const unsigned grid = (unsigned)((B * P::G + P::threads - 1) / P::threads);
        CK(cudaMemset(x0, 0, Bp * N * sizeof(T))); CK(cudaMemset(ok0, 7, B));
        float ms = time_ms([&] { kern<<<grid, P::threads>>>(A, b, x0, nullptr, ok0, nullptr, nullptr, nullptr, 0, B); });
        CK(cudaGetLastError());
        CK(cudaMemcpy(hx0.data(), x0, Bp * N * sizeof(T), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hok0.data(), ok0, B, cudaMemcpyDeviceToHost));
        printf("n=%2d %s %s exact  %8.3f ms  %9.4g sys/s  of 6548.8: %.3f\n", N, sizeof(T) == 4 ? "f32" : "f64", SOA ? "soa" : "aos", ms,
               B / (ms * 1e-3), bytes / ms * 1e-9 / 6.5488);
    }
    {
        CK(cudaMemset(x, 0, Bp * N * sizeof(T))); CK(cudaMemset(ok, 7, B));
        float ms = time_ms([&] {
            int rc = rowfast_solve_impl<T>(N, B, A, b, x, ok, 0, SOA ? GMB_LAYOUT_SOA : GMB_LAYOUT_AOS, 1, 0);
            if (rc != 0) { printf("rowfast rc=%d\n", rc); exit(1); }
        });
        CK(cudaGetLastError());
        CK(cudaMemcpy(hx.data(), x, Bp * N * sizeof(T), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hok.data(), ok, B, cudaMemcpyDeviceToHost));
        int64_t bad = 0, nok = 0;
        for (int64_t i = 0; i < B; ++i) nok += hok0[i] == 1;
        for (int64_t i = 0; i < B; ++i)
            if (hok[i] != hok0[i]) { if (bad < 5) printf("  ok mismatch at %lld: %d vs %d\n", (long long)i, hok[i], hok0[i]); ++bad; }
        if (memcmp(hx.data(), hx0.data(), Bp * N * sizeof(T)) != 0) {
            for (int64_t i = 0; i < Bp * N; ++i) {
                T u = hx[i], v = hx0[i];
                if (memcmp(&u, &v, sizeof(T)) != 0 && !(u != u && v != v)) { if (bad < 5) printf("  x mismatch at elem %lld: %g vs %g\n", (long long)i, (double)u, (double)v); ++bad; }
            }
        }
        printf("n=%2d %s %s fast   %8.3f ms  %9.4g sys/s  of 6548.8: %.3f  (ok=1: %lld)  mismatches: %lld\n",
               N, sizeof(T) == 4 ? "f32" : "f64", SOA ? "soa" : "aos", ms, B / (ms * 1e-3), bytes / ms * 1e-9 / 6.5488, (long long)nok, (long long)bad);
        // the speculative kernel alone, launch variants (threads per CTA, CTAs per SM, select mode)
        int32_t* scratch; CK(cudaMalloc(&scratch, (B + 4) * 4)); CK(cudaMemset(scratch, 0, 16));
        auto variant = [&](auto kern, int g, int threads, size_t smem, const char* name) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned grid = (unsigned)((B * g + threads - 1) / threads);
            float ms_fast = time_ms([&] { CK(cudaMemsetAsync(scratch, 0, 4)); kern<<<grid, threads, smem>>>(A, b, x, ok, 0, B, scratch + 4, scratch); });
            CK(cudaGetLastError());
            int cnt = 0; CK(cudaMemcpy(&cnt, scratch, 4, cudaMemcpyDeviceToHost));
            printf("      speculative kernel alone %-28s %8.3f ms  of 6548.8: %.3f  handed back %d\n", name, ms_fast, bytes / ms_fast * 1e-9 / 6.5488, cnt);
            fflush(stdout);
        };
#define VAR(G_, TH, MB, SI) variant(k_fast_solve<T, N, SOA, true, G_, TH, MB, true, SI>, G_, TH, fast_smem_bytes<T, N, G_, TH>(), "G=" #G_ " thr=" #TH " minb=" #MB " selimad=" #SI)
        if (sizeof(T) == 4) {
            VAR(4, 256, 2, 0);
            VAR(4, 128, 4, 0);
            VAR(4, 64, 8, 0);
            VAR(4, 64, 8, 1);
            VAR(4, 64, 8, 2);
            VAR(4, 64, 8, 3);
            VAR(4, 64, 8, 4);
            VAR(4, 32, 16, 2);
            VAR(4, 128, 4, 2);
        } else {
            VAR(4, 256, 1, 0);
            VAR(4, 128, 2, 0);
            VAR(4, 64, 4, 0);
            VAR(4, 128, 3, 0);
        }
        CK(cudaFree(scratch));
    }
    cudaFree(A); cudaFree(b); cudaFree(x0); cudaFree(x); cudaFree(tmp); cudaFree(ok0); cudaFree(ok);
}

template <typename T, bool SOA>
static void run_all(int64_t B, int nmin, int nmax) {
    for (int n = nmin; n <= nmax; ++n)
        dispatch_int<9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) { run_n<T, N.value, SOA>(B); });
}

int main(int argc, char** argv) {
    const int64_t B = (int64_t)1 << (argc > 1 ? atoi(argv[1]) : 20);
    const bool soa = argc > 2 && !strcmp(argv[2], "soa");
    const bool f64 = argc > 3 && !strcmp(argv[3], "f64");
    const int nmin = argc > 5 ? atoi(argv[4]) : 16, nmax = argc > 5 ? atoi(argv[5]) : 16;
#ifndef ROWFAST_ONLY_F32
    if (f64) { if (soa) run_all<double, true>(B, nmin, nmax); else run_all<double, false>(B, nmin, nmax); return 0; }
#endif
    if (soa) run_all<float, true>(B, nmin, nmax); else run_all<float, false>(B, nmin, nmax);
    return 0;
}
