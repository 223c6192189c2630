// Experiment harness for the 16x16 float fused PLU solve (config C5): times kernel variants on the same
// random batch and checks every variant bit for bit against the production kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
//        -I geomc_b200/csrc profiles/exp/row16_bench.cu -o profiles/exp/row16_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gmb_rowplu.cuh"
#ifdef WITH_LANEROW
#include "gmb_lanerow.cuh"   // profiles/exp/ (kept as a measured negative result)
#endif

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

int main(int argc, char** argv) {
    constexpr int N = 16;
    const int64_t B = argc > 1 ? atoll(argv[1]) : (1 << 20);
    const char* only = argc > 2 ? argv[2] : nullptr;
    float *A, *b, *x0, *x;
    uint8_t *ok0, *ok;
    CK(cudaMalloc(&A, B * N * N * 4)); CK(cudaMalloc(&b, B * N * 4)); CK(cudaMalloc(&x0, B * N * 4)); CK(cudaMalloc(&x, B * N * 4));
    CK(cudaMalloc(&ok0, B)); CK(cudaMalloc(&ok, B));
    fill<<<(unsigned)((B * N * N + 255) / 256), 256>>>(A, B * N * N, 1);
    fill<<<(unsigned)((B * N + 255) / 256), 256>>>(b, B * N, 2);
    // a few singular / tie-heavy systems
    CK(cudaMemset(A + 7 * N * N, 0, N * N * 4));
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
            for (int64_t i = 0; i < B; ++i) {
                if (hok[i] != hok0[i]) { ++bad; continue; }
                if (memcmp(&hx[i * N], &hx0[i * N], N * 4) != 0) {
                    bool nan_only = true;
                    for (int k = 0; k < N; ++k) {
                        float u = hx[i * N + k], v = hx0[i * N + k];
                        if (memcmp(&u, &v, 4) != 0 && !(u != u && v != v)) nan_only = false;
                    }
                    if (!nan_only) ++bad;
                }
            }
        }
        printf("%-44s %8.3f ms  %6.3g sys/s  %5.2f TB/s  of 6548.8: %.3f  mismatches: %d\n", name, ms, B / (ms * 1e-3), bytes / ms * 1e-9,
               bytes / ms * 1e-9 / 6.5488, bad);
        fflush(stdout);
    };
    auto want = [&](const char* name) { return !only || strstr(name, only); };

    {   // production kernel = reference result
        constexpr int G = row_group(N, 1);
        auto kern = k_row_solve<float, N, false, true, false>;
        const unsigned grid = (unsigned)((B * G + kRowThreads - 1) / kRowThreads);
        float ms = time_ms([&] { kern<<<grid, kRowThreads>>>(A, b, x0, nullptr, ok0, nullptr, nullptr, nullptr, 0, B); });
        CK(cudaGetLastError());
        CK(cudaMemcpy(hx0.data(), x0, B * N * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hok0.data(), ok0, B, cudaMemcpyDeviceToHost));
        report("row G=2 minb=2 (production)", ms, false);
    }
#define ROWVAR(G_, MB_)                                                                                          \
    if (want("rowvar")) {                                                                                        \
        auto kern = k_row_solve<float, N, false, true, false, G_, MB_>;                                          \
        const unsigned grid = (unsigned)((B * G_ + kRowThreads - 1) / kRowThreads);                              \
        CK(cudaMemset(x, 0, B * N * 4));                                                                         \
        float ms = time_ms([&] { kern<<<grid, kRowThreads>>>(A, b, x, nullptr, ok, nullptr, nullptr, nullptr, 0, B); }); \
        CK(cudaGetLastError());                                                                                  \
        report("rowvar G=" #G_ " minb=" #MB_, ms, true);                                                         \
    }
#define NOLU(MB_)                                                                                                \
    if (want("nolu")) {                                                                                          \
        auto kern = k_row_solve<float, N, false, true, false, 2, MB_, false>;                                    \
        const unsigned grid = (unsigned)((B * 2 + kRowThreads - 1) / kRowThreads);                               \
        CK(cudaMemset(x, 0, B * N * 4));                                                                         \
        float ms = time_ms([&] { kern<<<grid, kRowThreads>>>(A, b, x, nullptr, ok, nullptr, nullptr, nullptr, 0, B); }); \
        CK(cudaGetLastError());                                                                                  \
        report("row G=2 no-LU minb=" #MB_, ms, true);                                                            \
    }
    NOLU(1)
    NOLU(2)
#define SHX(G_, MB_, LU_)                                                                                        \
    if (want("shx")) {                                                                                           \
        auto kern = k_row_solve<float, N, false, true, false, G_, MB_, LU_, true>;                               \
        const unsigned grid = (unsigned)((B * G_ + kRowThreads - 1) / kRowThreads);                              \
        CK(cudaMemset(x, 0, B * N * 4));                                                                         \
        float ms = time_ms([&] { kern<<<grid, kRowThreads>>>(A, b, x, nullptr, ok, nullptr, nullptr, nullptr, 0, B); }); \
        CK(cudaGetLastError());                                                                                  \
        report("row shuffle-exchange G=" #G_ " minb=" #MB_ " lu=" #LU_, ms, true);                               \
    }
#define SHXT(G_, MB_, TH_)                                                                                       \
    if (want("shx")) {                                                                                           \
        auto kern = k_row_solve<float, N, false, true, false, G_, MB_, false, true, TH_>;                        \
        const unsigned grid = (unsigned)((B * G_ + TH_ - 1) / TH_);                                              \
        CK(cudaMemset(x, 0, B * N * 4));                                                                         \
        float ms = time_ms([&] { kern<<<grid, TH_>>>(A, b, x, nullptr, ok, nullptr, nullptr, nullptr, 0, B); }); \
        CK(cudaGetLastError());                                                                                  \
        report("row shuffle G=" #G_ " minb=" #MB_ " threads=" #TH_, ms, true);                                   \
    }
    SHXT(4, 2, 256)
    SHXT(4, 1, 256)
    SHXT(4, 3, 128)
    SHXT(4, 4, 128)
    SHXT(4, 1, 384)
    SHXT(4, 1, 512)
    SHXT(4, 5, 96)
    SHXT(4, 8, 64)
    SHXT(2, 1, 256)
    SHXT(2, 2, 256)
    SHX(2, 1, false)
    SHX(2, 2, false)
    SHX(2, 2, true)
    SHX(4, 1, false)
    SHX(4, 2, false)
    SHX(4, 3, false)
    SHX(4, 4, false)
    SHX(8, 4, false)
    SHX(8, 6, false)
#if defined(WITH_LANEROW)
#define LANEVAR(G_, MB_)                                                                                         \
    if (want("lane")) {                                                                                          \
        auto kern = k_lane_solve<float, N, G_, false, true, MB_>;                                                \
        const unsigned grid = (unsigned)((B * G_ + kLaneThreads - 1) / kLaneThreads);                            \
        CK(cudaMemset(x, 0, B * N * 4));                                                                         \
        float ms = time_ms([&] { kern<<<grid, kLaneThreads>>>(A, b, x, nullptr, ok, 0, B); });                   \
        CK(cudaGetLastError());                                                                                  \
        report("lanerow G=" #G_ " minb=" #MB_, ms, true);                                                        \
    }
    LANEVAR(16, 8)
    LANEVAR(8, 4)
    LANEVAR(8, 6)
    LANEVAR(8, 8)
    LANEVAR(4, 4)
#endif
    return 0;
}
