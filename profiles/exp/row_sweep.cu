// Sweep of the row-distributed fused PLU solve variants over n = 9..16, float and double (AoS, row-major,
// 2^20 systems, no LU output): bounce-buffer exchange (G from the register budget) vs. register-only shuffle
// exchange with G = 2 / 4 / 8 lanes per system.  Every variant is checked bit for bit against the first.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
//        -I geomc_b200/csrc profiles/exp/row_sweep.cu -o profiles/exp/row_sweep
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gmb_rowplu.cuh"

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
    p[i] = (T)v * (T)(1.0 / 131072.0);
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <class F>
static float time_ms(F&& f, int reps = 4) {
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

template <typename T, int N, int G, int MINB, bool SHX, int THREADS = 128>
static float run(const T* A, const T* b, T* x, uint8_t* ok, int64_t B) {
    auto kern = k_row_solve<T, N, false, true, false, G, MINB, false, SHX, THREADS>;
    const unsigned grid = (unsigned)((B * G + THREADS - 1) / THREADS);
    float ms = time_ms([&] { kern<<<grid, THREADS>>>(A, b, x, nullptr, ok, nullptr, nullptr, nullptr, 0, B); });
    CK(cudaGetLastError());
    return ms;
}

template <typename T, int N>
static void sweep(int64_t B) {
    T *A, *b, *x0, *x;
    uint8_t *ok0, *ok;
    CK(cudaMalloc(&A, B * N * N * sizeof(T))); CK(cudaMalloc(&b, B * N * sizeof(T)));
    CK(cudaMalloc(&x0, B * N * sizeof(T))); CK(cudaMalloc(&x, B * N * sizeof(T)));
    CK(cudaMalloc(&ok0, B)); CK(cudaMalloc(&ok, B));
    fill<T><<<(unsigned)((B * N * N + 255) / 256), 256>>>(A, B * N * N, 1 + N);
    fill<T><<<(unsigned)((B * N + 255) / 256), 256>>>(b, B * N, 2 + N);
    CK(cudaMemset(A + 7 * N * N, 0, N * N * sizeof(T)));
    std::vector<T> hx0(B * N), hx(B * N);
    std::vector<uint8_t> hok0(B), hok(B);
    auto check = [&]() {
        CK(cudaMemcpy(hx.data(), x, B * N * sizeof(T), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hok.data(), ok, B, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int64_t i = 0; i < B; ++i) {
            if (hok[i] != hok0[i]) { ++bad; continue; }
            for (int k = 0; k < N; ++k) {
                T u = hx[i * N + k], v = hx0[i * N + k];
                if (memcmp(&u, &v, sizeof(T)) != 0 && !(u != u && v != v)) { ++bad; break; }
            }
        }
        return bad;
    };
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int G0 = row_group(N, WORDS);
    float t0 = run<T, N, G0, 2, false>(A, b, x0, ok0, B);
    CK(cudaMemcpy(hx0.data(), x0, B * N * sizeof(T), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hok0.data(), ok0, B, cudaMemcpyDeviceToHost));
    float t2 = run<T, N, 2, 2, true>(A, b, x, ok, B);   int b2 = check();
    float t4 = run<T, N, 4, 3, true>(A, b, x, ok, B);   int b4 = check();
    float t8 = run<T, N, 2, 1, true, 256>(A, b, x, ok, B);   int b8 = check();
    float t9 = run<T, N, 4, 2, true, 256>(A, b, x, ok, B);   int b9 = check();
    float t1 = 0;
    if constexpr (N <= 8) { t1 = run<T, N, 1, 2, true>(A, b, x, ok, B); b9 += check(); }
    printf("| %s | %2d | %d | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f | %d |\n", sizeof(T) == 4 ? "float" : "double", N, G0, t0, t2, t4, t8, t9, t1, b2 + b4 + b8 + b9);
    fflush(stdout);
    cudaFree(A); cudaFree(b); cudaFree(x0); cudaFree(x); cudaFree(ok0); cudaFree(ok);
}

int main(int argc, char** argv) {
    const int64_t B = argc > 1 ? atoll(argv[1]) : (1 << 20);
    printf("| type | n | bounce G | bounce ms | shuffle G=2 ms | shuffle G=4 ms | shuffle G=2, 256 thr | shuffle G=4, 256 thr | shuffle G=1 | mismatches |\n|---|---|---|---|---|---|---|---|---|---|\n");
#define BOTH(N) sweep<float, N>(B); sweep<double, N>(B);
    if (argc > 2) { BOTH(5) BOTH(6) BOTH(7) BOTH(8) return 0; }
    BOTH(9) BOTH(10) BOTH(11) BOTH(12) BOTH(13) BOTH(14) BOTH(15) BOTH(16)
    return 0;
}
