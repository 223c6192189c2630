// Experiment harness for the shared-memory resident fused PLU solve (gmb_tile.cuh) against the production register
// kernel (gmb_rowplu.cuh), float, AoS, row-major: times both on the same random batch and compares bit for bit.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
//        -I geomc_b200/csrc -I profiles/exp profiles/exp/tile_bench.cu -o profiles/exp/tile_bench
//   profiles/exp/tile_bench [systems] [n|0 = all]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gmb_rowplu.cuh"
#include "gmb_tile.cuh"
#include "gmb_tile4.cuh"

namespace gmb { std::atomic<int64_t> g_launch_count{0}; }
using namespace gmb;

__global__ void fill(float* p, int64_t n, uint32_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    int v = (int)(z & 0xFFFFF) - (1 << 19);
    p[i] = (float)v * (1.0f / 131072.0f);          // exact grid in [-4, 4)
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

template <int N, int NT, int MINB, int MODE = 0>
static float run_tile(const float* A, const float* b, float* x, uint8_t* ok, int64_t B) {
    auto kern = k_tile_solve<float, N, false, NT, MINB, MODE>;
    constexpr size_t smem = tile_smem_bytes<float, N, NT>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((B + NT - 1) / NT);
    float ms = time_ms([&] { kern<<<grid, NT, smem>>>(A, b, x, ok, 0, B); });
    CK(cudaGetLastError());
    return ms;
}

template <int NT, int MINB>
static float run_tile4(const float* A, const float* b, float* x, uint8_t* ok, int64_t B) {
    auto kern = k_tile4_solve<false, NT, MINB>;
    constexpr size_t smem = tile4_smem_bytes<NT>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((B + NT / 4 - 1) / (NT / 4));
    float ms = time_ms([&] { kern<<<grid, NT, smem>>>(A, b, x, ok, 0, B); });
    CK(cudaGetLastError());
    return ms;
}

template <int N>
static void bench(int64_t B) {
    float *A, *b, *x0, *x;
    uint8_t *ok0, *ok;
    CK(cudaMalloc(&A, B * N * N * 4)); CK(cudaMalloc(&b, B * N * 4)); CK(cudaMalloc(&x0, B * N * 4)); CK(cudaMalloc(&x, B * N * 4));
    CK(cudaMalloc(&ok0, B)); CK(cudaMalloc(&ok, B));
    fill<<<(unsigned)((B * N * N + 255) / 256), 256>>>(A, B * N * N, 1);
    fill<<<(unsigned)((B * N + 255) / 256), 256>>>(b, B * N, 2);
    // singular, tie-heavy and zero-multiplier systems
    CK(cudaMemset(A + 7 * N * N, 0, N * N * 4));                       // all zero
    CK(cudaMemset(A + 9 * N * N + N, 0, N * 4));                       // a zero row
    {
        std::vector<float> h(N * N);
        for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) h[r * N + c] = (float)(((r * 7 + c * 3) % 5) - 2);   // many ties / zeros
        for (int r = 0; r < N; ++r) h[r * N + r] += 3.0f;
        CK(cudaMemcpy(A + 11 * N * N, h.data(), N * N * 4, cudaMemcpyHostToDevice));
        for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) h[r * N + c] = (r == c || r == c + 1 || c == r + 2) ? 1.0f + (r % 3) : 0.0f;   // banded: zero multipliers
        CK(cudaMemcpy(A + 13 * N * N, h.data(), N * N * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> hx0(B * N), hx(B * N);
    std::vector<uint8_t> hok0(B), hok(B);
    const double bytes = (double)B * (N * N + 2 * N) * 4;
    auto report = [&](const char* name, float ms, bool check) {
        int bad = -1;
        if (check) {
            CK(cudaMemcpy(hx.data(), x, B * N * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hok.data(), ok, B, cudaMemcpyDeviceToHost));
            bad = 0;
            int shown = 0;
            for (int64_t i = 0; i < B; ++i) {
                bool diff = hok[i] != hok0[i];
                if (!diff && memcmp(&hx[i * N], &hx0[i * N], N * 4) != 0) {
                    for (int k = 0; k < N; ++k) {
                        float u = hx[i * N + k], v = hx0[i * N + k];
                        if (memcmp(&u, &v, 4) != 0 && !(u != u && v != v)) diff = true;
                    }
                }
                if (diff) {
                    ++bad;
                    if (shown++ < 3) printf("   system %lld: ok %d vs %d, x0 %a vs %a\n", (long long)i, hok[i], hok0[i], hx[i * N], hx0[i * N]);
                }
            }
        }
        printf("n=%2d %-34s %8.3f ms  %6.3g sys/s  of 6548.8 GB/s: %.3f  mismatches: %d\n", N, name, ms, B / (ms * 1e-3),
               bytes / ms * 1e-9 / 6.5488 * 1e-3, bad);
        fflush(stdout);
    };
    {
        using P = RowNoLu<float, N>;
        auto kern = k_row_solve<float, N, false, true, false, P::G, P::minb, false, P::shfl, P::threads>;
        const unsigned grid = (unsigned)((B * P::G + P::threads - 1) / P::threads);
        float ms = time_ms([&] { kern<<<grid, P::threads>>>(A, b, x0, nullptr, ok0, nullptr, nullptr, nullptr, 0, B); });
        CK(cudaGetLastError());
        CK(cudaMemcpy(hx0.data(), x0, B * N * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hok0.data(), ok0, B, cudaMemcpyDeviceToHost));
        report("register kernel (production)", ms, false);
    }
#define TILE(NT_, MB_)                                                                \
    {                                                                                 \
        CK(cudaMemset(x, 0, B * N * 4)); CK(cudaMemset(ok, 7, B));                    \
        float ms = run_tile<N, NT_, MB_>(A, b, x, ok, B);                             \
        report("tile NT=" #NT_ " minb=" #MB_, ms, true);                              \
    }
#define TILE4(NT_, MB_)                                                               \
    if constexpr (N == 16) {                                                          \
        CK(cudaMemset(x, 0, B * N * 4)); CK(cudaMemset(ok, 7, B));                    \
        float ms = run_tile4<NT_, MB_>(A, b, x, ok, B);                               \
        report("tile4 NT=" #NT_ " minb=" #MB_, ms, true);                             \
    }
    TILE4(128, 1)
    TILE4(128, 6)
    TILE4(256, 3)
    TILE4(64, 1)
    TILE(32, 1)
    { float ms = run_tile<N, 32, 1, 1>(A, b, x, ok, B); report("tile NT=32 ingest only", ms, false); }
    { float ms = run_tile<N, 32, 1, 2>(A, b, x, ok, B); report("tile NT=32 no ingest", ms, false); }
    TILE(64, 1)
    TILE(96, 1)
    if constexpr (tile_smem_bytes<float, N, 192>() <= 227 * 1024) TILE(192, 1)
    CK(cudaFree(A)); CK(cudaFree(b)); CK(cudaFree(x0)); CK(cudaFree(x)); CK(cudaFree(ok0)); CK(cudaFree(ok));
}

int main(int argc, char** argv) {
    const int64_t B = argc > 1 ? atoll(argv[1]) : (1 << 20);
    const int only = argc > 2 ? atoi(argv[2]) : 16;
    if (only == 0 || only == 16) bench<16>(B);
    if (only == 0 || only == 8) bench<8>(B);
    return 0;
}
