// Sweep of the row-distributed decomp_plu variants (in place, SoA-interleaved container, 2^21 systems):
// bounce-buffer exchange (production) vs. register-only shuffle exchange with 2 / 4 lanes per system.
// Factors, permutation, parity and degenerate counts of every variant are compared bit for bit with the first.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
//        -I geomc_b200/csrc profiles/exp/decomp_sweep.cu -o profiles/exp/decomp_sweep
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

template <typename T, int N, int G, int MINB, bool SHX>
static float run(const T* A0, T* A, uint8_t* perm, uint8_t* par, uint8_t* deg, int64_t B) {
    auto kern = k_row_solve<T, N, true, true, true, G, MINB, true, SHX>;
    const unsigned grid = (unsigned)((B * G + kRowThreads - 1) / kRowThreads);
    const size_t bytes = (size_t)B * N * N * sizeof(T);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e0);
        kern<<<grid, kRowThreads>>>(A, nullptr, nullptr, A, nullptr, perm, par, deg, 0, B);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

template <typename T, int N>
static void sweep(int64_t B) {
    T *A0, *A;
    uint8_t *perm, *par, *deg;
    const size_t n = (size_t)B * N * N;
    CK(cudaMalloc(&A0, n * sizeof(T))); CK(cudaMalloc(&A, n * sizeof(T)));
    CK(cudaMalloc(&perm, B * N)); CK(cudaMalloc(&par, B)); CK(cudaMalloc(&deg, B));
    fill<T><<<(unsigned)((n + 255) / 256), 256>>>(A0, n, 3 + N);
    CK(cudaMemset(A0, 0, 4096));                         // a few zero entries -> some degenerate columns in tile 0
    std::vector<T> h0(n), h(n);
    std::vector<uint8_t> p0(B * N), p(B * N), q0(B), q(B), d0(B), d(B);
    auto fetch = [&](std::vector<T>& ha, std::vector<uint8_t>& hp, std::vector<uint8_t>& hq, std::vector<uint8_t>& hd) {
        CK(cudaMemcpy(ha.data(), A, n * sizeof(T), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hp.data(), perm, B * N, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hq.data(), par, B, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hd.data(), deg, B, cudaMemcpyDeviceToHost));
    };
    auto same = [&]() {
        fetch(h, p, q, d);
        if (p != p0 || q != q0 || d != d0) return 1;
        for (size_t i = 0; i < n; ++i) {
            T u = h[i], v = h0[i];
            if (memcmp(&u, &v, sizeof(T)) != 0 && !(u != u && v != v)) return 1;
        }
        return 0;
    };
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int G0 = row_group(N, WORDS);
    float t0 = run<T, N, G0, 2, false>(A0, A, perm, par, deg, B);
    fetch(h0, p0, q0, d0);
    float t2 = run<T, N, 2, 2, true>(A0, A, perm, par, deg, B);   int b2 = same();
    float t4 = run<T, N, 4, 3, true>(A0, A, perm, par, deg, B);   int b4 = same();
    printf("| %s | %2d | %d | %.3f | %.3f | %.3f | %d |\n", sizeof(T) == 4 ? "float" : "double", N, G0, t0, t2, t4, b2 + b4);
    fflush(stdout);
    cudaFree(A0); cudaFree(A); cudaFree(perm); cudaFree(par); cudaFree(deg);
}

int main(int argc, char** argv) {
    const int64_t B = argc > 1 ? atoll(argv[1]) : (1 << 21);
    printf("| type | n | bounce G | bounce ms | shuffle G=2 ms | shuffle G=4 ms | mismatches |\n|---|---|---|---|---|---|---|\n");
#define BOTH(N) sweep<float, N>(B); sweep<double, N>(B);
    BOTH(7) BOTH(8) BOTH(9) BOTH(10) BOTH(12) BOTH(14) BOTH(16)
    return 0;
}
