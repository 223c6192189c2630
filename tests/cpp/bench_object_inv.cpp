// End-to-end time of the OBJECT form of config C3: geom::batch::inv on arrays of SimpleMatrix<float,4,4> objects
// (Matrix.h:717 applied to every element), beside the raw-pointer host entry point gmb_host_inv_f32 on the same data.
// Prints one JSON line.   bench_object_inv [log2_B = 22] [device = 0] [reps = 3]
//
//   g++ -std=c++17 -O2 -I include tests/cpp/bench_object_inv.cpp -L geomc_b200 -lgeomc_b200 -Wl,-rpath,$PWD/geomc_b200
//   add -DWITH_GEOMC -std=c++23 -include oracle/shim/prelude.h -I oracle/shim -I /root/reference for the reference's
//   real SimpleMatrix (oracle/Makefile builds that variant into oracle/_ref/, from where it travels to the GPU box).
#ifdef WITH_GEOMC
#include <geomc/linalg/Matrix.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <geomc_b200/batch.hpp>

namespace gb = geom::batch;

#ifndef WITH_GEOMC
namespace geom {
enum MatrixLayout { ROW_MAJOR, COL_MAJOR };
template <typename T, long M, long N, MatrixLayout L = ROW_MAJOR>
struct SimpleMatrix {
    typedef T elem_t;
    static constexpr long ROWDIM = M, COLDIM = N;
    static constexpr MatrixLayout Layout = L;
    T data[M * N];
    T* data_begin() { return data; }
    const T* data_begin() const { return data; }
};
}  // namespace geom
#endif

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    const int log2b = argc > 1 ? std::atoi(argv[1]) : 22;
    const int device = argc > 2 ? std::atoi(argv[2]) : 0;
    const int reps = argc > 3 ? std::atoi(argv[3]) : 3;
    const long B = 1L << log2b;
    using Mx = geom::SimpleMatrix<float, 4, 4>;
    if (gmb_device_count() == 0) {
        std::printf("{\"error\": \"no CUDA device (no CPU fallback)\"}\n");
        return 0;
    }
    static_assert(sizeof(bool) == 1, "");
    // --- raw arm: pinned flat arrays straight into the C ABI
    float* ra = static_cast<float*>(gmb_host_alloc((size_t)B * 64));
    float* ro = static_cast<float*>(gmb_host_alloc((size_t)B * 64));
    std::uint8_t* rok = static_cast<std::uint8_t*>(gmb_host_alloc((size_t)B));
    if (!ra || !ro || !rok) {
        std::printf("{\"error\": \"pinned allocation failed\"}\n");
        return 1;
    }
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    for (long s = 0; s < B; ++s) {
        float* m = ra + s * 16;
        for (int e = 0; e < 16; ++e) m[e] = u(rng) * 0.25f;
        for (int d = 0; d < 4; ++d) m[5 * d] += 1.f + (float)(s & 7) * 0.125f;
        if ((s & 0xfffff) == 77)
            for (int e = 0; e < 16; ++e) m[e] = 0.f;      // a singular record now and then
    }
    std::memset(ro, 0, (size_t)B * 64);
    auto time_of = [&](auto&& fn) {
        fn();                                              // warm-up: staging buffers, first-touch
        double best = 1e30;
        for (int r = 0; r < reps; ++r) {
            const double t0 = now();
            fn();
            const double dt = now() - t0;
            best = dt < best ? dt : best;
        }
        return best;
    };
    int rc = 0;
    const double t_raw = time_of([&] { rc |= gmb_host_inv_f32(4, B, ra, ro, rok, 1, 1, device); });
    if (rc != 0) {
        std::printf("{\"error\": \"gmb_host_inv_f32: %s\"}\n", gmb_error_string(rc));
        return 1;
    }
    // --- object arm: arrays of SimpleMatrix objects owned by the application (std::vector storage)
    std::vector<Mx> src((size_t)B), dst((size_t)B);
    std::vector<char> okc((size_t)B);
    for (long s = 0; s < B; ++s) {
        std::memcpy(src[(size_t)s].data_begin(), ra + s * 16, 64);
        std::memset(dst[(size_t)s].data_begin(), 0, 64);
    }
    bool* ok = reinterpret_cast<bool*>(okc.data());
    long nfalse = 0;
    const double t_pageable = time_of([&] { nfalse = gb::inv(dst.data(), src.data(), B, ok, device); });
    double t_pinned;
    bool pinned;
    {
        gb::PinnedScope p1(src.data(), (size_t)B * sizeof(Mx)), p2(dst.data(), (size_t)B * sizeof(Mx)), p3(ok, (size_t)B);
        pinned = p1.pinned() && p2.pinned() && p3.pinned();
        t_pinned = time_of([&] { nfalse = gb::inv(dst.data(), src.data(), B, ok, device); });
    }
    bool identical = true;
    for (long s = 0; s < B && identical; ++s)
        identical = std::memcmp(dst[(size_t)s].data_begin(), ro + s * 16, 64) == 0 && (ok[s] ? 1 : 0) == rok[s];
    std::printf("{\"systems\": %ld, \"object_type\": \"%s\", \"sizeof_object\": %zu, \"raw_pointer_systems_per_s\": %.6g, "
                "\"object_pinned_systems_per_s\": %.6g, \"object_pageable_systems_per_s\": %.6g, \"object_over_raw\": %.4f, "
                "\"registered\": %s, \"singular\": %ld, \"identical_to_raw\": %s, \"h2d_bytes\": %ld, \"d2h_bytes\": %ld}\n",
                B,
#ifdef WITH_GEOMC
                "geom::SimpleMatrix<float,4,4> (the reference's own type)",
#else
                "stand-in SimpleMatrix<float,4,4>",
#endif
                sizeof(Mx), B / t_raw, B / t_pinned, B / t_pageable, t_raw / t_pinned, pinned ? "true" : "false", nfalse,
                identical ? "true" : "false", B * 128, B * 65);
    gmb_host_free(ra);
    gmb_host_free(ro);
    gmb_host_free(rok);
    return identical ? 0 : 1;
}
