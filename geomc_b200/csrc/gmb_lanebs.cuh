// gmb_lanebs.cuh -- backsolve_lu / backsolve_plu (LUDecomp.h:332-366 / 271-312) with ONE lane per system and a compile-time
// number of right-hand sides (m = 1..4), for the sizes whose LU factors and right-hand sides fit one lane's registers
// (float n <= 14, double n <= 9 with m = 1; float n <= 13, double n <= 9 with m = 3).
//
// k_row_backsolve (gmb_rowchol.cuh) takes m at run time: it walks the columns of X one after the other, fetching each one with
// dynamically addressed scalar loads (the gather through `reorder`, :294) between the sweeps and writing it back with
// per-lane strided scalar stores -- half-used sectors on an AoS batch (run r2w, before this kernel: 0.45 .. 0.55 of the HBM
// peak for float n = 5 .. 8 with three right-hand sides, 0.72 from the SoA container).  Here
//   * AoS records travel like the sub-warp kernels' (gmb_io.cuh): the warp's 32 LU records AND its 32 right-hand-side records
//     are copied global -> shared with cp.async, all in flight together, as whole 16-byte chunks of one contiguous range;
//     the gather b(reorder[r]) is a dynamically indexed read of the lane's record in shared memory; X leaves through the
//     same buffer with coalesced 128-bit stores;
//   * SoA batches travel the same way (soa_stage_issue: the warp's 32-system segment of every element plane, whole 16-byte chunks,
//     plane e at [e][lane] in shared memory -- conflict-free for the gather's run-time e); X leaves with direct plane stores.
//     Direct plane loads into registers measured 0.02 .. 0.2 of the peak lower (run r2am: float 13x13 0.69 -> 0.89);
//   * all m columns ride through one forward and one backward sweep: a(r, c) is used m times per load, the m division
//     chains of the backward sweep overlap.
// Per element the operations and their order are those of backsolve_lu (c ascending in the forward sweep, c descending in the
// backward one, division by u(c,c), rows < skip untouched): bit-identical to k_row_backsolve, tests/test_gpu_parity.py.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

#ifndef GMB_LANEBS_WORDS
#define GMB_LANEBS_WORDS 216          // register words for [LU | X] of one lane
#endif
constexpr int kLaneBsThreads = 128;
constexpr bool lanebs_fits(int n, int m, int words) { return words * n * (n + m) <= GMB_LANEBS_WORDS; }

// element e (run-time index) of the lane's record in a staging buffer filled by aos_stage_issue<T, E>
template <typename T, int E>
__device__ __forceinline__ T aos_stage_elem(int lane, int e, const unsigned char* sm) {
    constexpr int R = E * (int)sizeof(T);
    using Cfg = StageCfg<R>;
    if constexpr (Cfg::vec16) {
        constexpr int V = soa_v<T>();
        return *reinterpret_cast<const T*>(sm + Cfg::chunk(lane, e / V) * 16 + (e % V) * (int)sizeof(T));
    } else {
        return *reinterpret_cast<const T*>(sm + lane * R + e * (int)sizeof(T));
    }
}

template <typename T, int N, int M, bool PLU, bool SOA, bool RM>
__global__ void __launch_bounds__(kLaneBsThreads, ((int)sizeof(T) / 4 * N * (N + M) > 120) ? 2 : 3)
k_lane_backsolve(const T* __restrict__ LUm, const uint8_t* __restrict__ perm, const T* Bm, T* X, int skip, int64_t B) {
    constexpr int ST = SOA ? soa_w<T>() : 1;
    constexpr int W = soa_w<T>();
    constexpr int EA = N * N, EX = N * M;
    extern __shared__ __align__(16) unsigned char gmb_dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int64_t s_raw = (int64_t)blockIdx.x * kLaneBsThreads + threadIdx.x;
    const bool valid = s_raw < B;
    const int64_t s = valid ? s_raw : B - 1;
    const int64_t s_warp0 = s_raw - lane;
    const bool full = s_warp0 + 32 <= B;                      // warp-uniform: all 32 systems exist (SoA: inside one tile, W % 32 == 0)
    unsigned char* smA = gmb_dyn_smem + (size_t)(threadIdx.x >> 5) * (32 * (EA + EX) * sizeof(T));
    unsigned char* smX = smA + 32 * EA * sizeof(T);
    const int64_t baseA = SOA ? (s / W) * (int64_t)EA * W + (s % W) : s * (int64_t)EA;
    const int64_t baseX = SOA ? (s / W) * (int64_t)EX * W + (s % W) : s * (int64_t)EX;

    if (full) {
        if constexpr (SOA) {
            soa_stage_issue<T, EA>(LUm, s_warp0, lane, smA);
            soa_stage_issue<T, EX>(Bm, s_warp0, lane, smX);
        } else {
            aos_stage_issue<T, EA>(LUm, s_warp0, lane, smA);
            aos_stage_issue<T, EX>(Bm, s_warp0, lane, smX);
        }
    }
    int src[N];
#pragma unroll
    // y(r) = b(reorder[r]) (:294).  The index addresses shared memory: a corrupt `reorder` entry (>= n, never produced by decomp_plu)
    // is clamped so that it cannot leave the warp's staging buffer
    for (int r = 0; r < N; ++r) src[r] = PLU ? min((int)perm[s * N + r], N - 1) : r;

    T a[N][N], x[N][M];
    if (full) {
        cp_async_wait_all_();
        __syncwarp();
        if constexpr (SOA) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) a[r][c] = soa_stage_elem<T>(lane, RM ? N * r + c : N * c + r, smA);
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < M; ++j) x[r][j] = soa_stage_elem<T>(lane, RM ? M * src[r] + j : N * j + src[r], smX);
        } else {
            T qa[EA];
            aos_stage_read<T, EA>(lane, qa, smA);
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) a[r][c] = qa[RM ? N * r + c : N * c + r];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < M; ++j) x[r][j] = aos_stage_elem<T, EX>(lane, RM ? M * src[r] + j : N * j + src[r], smX);
        }
        __syncwarp();                                          // the buffer is reused by the staged store
    } else {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = LUm[baseA + (int64_t)(RM ? N * r + c : N * c + r) * ST];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) x[r][j] = Bm[baseX + (int64_t)(RM ? M * src[r] + j : N * j + src[r]) * ST];
    }

    // forward: L y = b, unit diagonal (:345-353)
#pragma unroll
    for (int c = 0; c < N - 1; ++c)
#pragma unroll
        for (int r = c + 1; r < N; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) x[r][j] = fma_rn(-x[c][j], a[r][c], x[r][j]);
    // backward: U x = y (:356-365); rows below `skip` keep y
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        const bool on = c >= skip;
        T q[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            q[j] = div_rn(x[c][j], a[c][c]);
            x[c][j] = on ? q[j] : x[c][j];
        }
#pragma unroll
        for (int r = 0; r < c; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const T v = fma_rn(-q[j], a[r][c], x[r][j]);
                if (on && r >= skip) x[r][j] = v;
            }
    }

    T qx[EX];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int j = 0; j < M; ++j) qx[RM ? M * r + j : N * j + r] = x[r][j];
    if (full && !SOA) {
        aos_store_staged<T, EX>(X, s_warp0, lane, qx, smX);
    } else if (valid) {
#pragma unroll
        for (int e = 0; e < EX; ++e) X[baseX + (int64_t)e * ST] = qx[e];
    }
}

template <typename T, int M>
int lanebs_launch(int mode, int n, int64_t B, const T* LU, const uint8_t* perm, const T* Bm, T* X, int skip, int layout, int rm,
                  cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14>(n, [&](auto N) {
        if constexpr (lanebs_fits(N.value, M, (int)sizeof(T) / 4)) {
            const unsigned grid = (unsigned)ceil_div(B, kLaneBsThreads);
            if (grid == 0) {
                rc = GMB_OK;
                return;
            }
            dispatch_bool(mode == 1, [&](auto PLU) {
                dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                    dispatch_bool(rm != 0, [&](auto RM) {
                        auto kern = k_lane_backsolve<T, N.value, M, PLU.value, SOA.value, RM.value>;
                        const size_t smem = (size_t)(kLaneBsThreads / 32) * 32 * (N.value * (N.value + M)) * sizeof(T);
                        if (smem > 48 * 1024) {
                            static std::atomic<uint64_t> raised;          // per instantiation, one bit per device
                            raise_dyn_smem(kern, smem, raised);
                        }
                        kern<<<grid, kLaneBsThreads, smem, st>>>(LU, perm, Bm, X, skip, B);
                    });
                });
            });
            rc = after_launch();
        }
    });
    return rc;
}

template <typename T>
int lanebs_impl(int mode, int n, int m, int64_t B, const T* LU, const uint8_t* perm, const T* Bm, T* X, int skip, int layout, int rm,
                cudaStream_t st) {
    if (mode != 0 && mode != 1) return GMB_ERR_UNSUPPORTED;
    if (mode == 1 && X == Bm) return GMB_ERR_UNSUPPORTED;        // the gather of backsolve_plu needs distinct buffers (as in the reference)
    switch (m) {
        case 1: return lanebs_launch<T, 1>(mode, n, B, LU, perm, Bm, X, skip, layout, rm, st);
        case 2: return lanebs_launch<T, 2>(mode, n, B, LU, perm, Bm, X, skip, layout, rm, st);
        case 3: return lanebs_launch<T, 3>(mode, n, B, LU, perm, Bm, X, skip, layout, rm, st);
        case 4: return lanebs_launch<T, 4>(mode, n, B, LU, perm, Bm, X, skip, layout, rm, st);
        default: return GMB_ERR_UNSUPPORTED;
    }
}

}  // namespace gmb
