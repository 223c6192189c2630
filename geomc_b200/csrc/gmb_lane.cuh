// gmb_lane.cuh -- ONE THREAD PER SYSTEM Cholesky for n = 5..8 (float) / 5..6 (double): the whole matrix and up to four
// right-hand sides live in one thread's registers, so there are no shuffles; the only cooperative step is
// the ingest / egress of AoS records (a warp moves its 32 records as ONE contiguous range with coalesced 128-bit
// accesses through a warp-private shared-memory transpose, aos_load_staged / aos_store_staged in gmb_io.cuh; the SoA
// container needs none: lane l reads word l of every element plane, 128 contiguous bytes per warp instruction).
//
// k_lane_chol -- cholesky (Cholesky.h:37-58) and the fused Cholesky solve: the row kernel's arithmetic (one reciprocal
//   per diagonal entry, shared by its column and the two substitution sweeps) with the staged ingest, so that record
//   sizes that are not multiples of 16 bytes (n = 5, 6, 7) no longer load scalars at a stride of one record.
// (A speculative one-thread-per-system PLU solve with the same ingest was measured and lost to k_sub_plu everywhere:
// profiles/exp/gmb_lane_solve.cuh, profiles/r2_lane_kernels.md.)
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

constexpr int kLaneThreads = 128;          // 4 warps: stage_buffer() holds four warp-private buffers

template <typename T> struct LaneWin;
template <> struct LaneWin<float> {
    __device__ static __forceinline__ float lo() { return 6.938893903907228e-18f; }        // 2^-57
    __device__ static __forceinline__ float hi() { return 1.4411518807585587e17f; }        // 2^57
    __device__ static __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
};
template <> struct LaneWin<double> {
    __device__ static __forceinline__ double lo() { return 3.872591914849318e-121; }       // 2^-400
    __device__ static __forceinline__ double hi() { return 2.582249878086909e120; }        // 2^400
    __device__ static __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
};

__device__ __forceinline__ float  lane_max(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double lane_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  lane_min(float a, float b)   { return fminf(a, b); }
__device__ __forceinline__ double lane_min(double a, double b) { return fmin(a, b); }

// d0 = fma(s, b0, d0); d1 = fma(s, b1, d1): one FFMA2 for float (sm_100a; same IEEE result per element), two DFMAs
__device__ __forceinline__ void lane_fma2(float& d0, float& d1, float s, float b0, float b1) {
    asm("{\n\t.reg .b64 bb, dd, ss;\n\tmov.b64 bb, {%2, %3};\n\tmov.b64 dd, {%0, %1};\n\tmov.b64 ss, {%4, %4};\n\t"
        "fma.rn.f32x2 dd, ss, bb, dd;\n\tmov.b64 {%0, %1}, dd;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(b0), "f"(b1), "f"(s));
}
__device__ __forceinline__ void lane_fma2(double& d0, double& d1, double s, double b0, double b1) {
    d0 = __fma_rn(s, b0, d0);
    d1 = __fma_rn(s, b1, d1);
}

// ---- one record of E elements per thread -----------------------------------------------------------------------
// `s` = this thread's system, `lane` = its lane, `full` = all 32 systems of the warp exist (warp-uniform).
// ONE warp-private staging buffer per kernel, shared by all of its staged loads and stores (they are sequential, each
// ends with a __syncwarp): EMAX = the largest record the kernel moves.
template <typename T, int EMAX>
__device__ __forceinline__ unsigned char* lane_stage_buffer() {
    __shared__ __align__(16) unsigned char buf[kLaneThreads / 32][32 * EMAX * (int)sizeof(T)];
    return buf[threadIdx.x >> 5];
}

template <typename T, int E, bool SOA>
__device__ __forceinline__ void lane_load(const T* __restrict__ base, int64_t s, int lane, bool full, bool valid, T (&q)[E],
                                          unsigned char* sm) {
    if constexpr (SOA) {
        constexpr int W = soa_w<T>();
        const T* p = base + (s / W) * (int64_t)(E * W) + (s % W);
#pragma unroll
        for (int e = 0; e < E; ++e) q[e] = valid ? __ldg(p + e * W) : (T)0;
    } else {
        if (StageCfg<E * (int)sizeof(T)>::use && full) {
            aos_load_staged<T, E>(base, s - lane, lane, q, sm);
        } else if (valid) {
            aos_load<T, E>(base, s, q);
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) q[e] = (T)0;
        }
    }
}
template <typename T, int E, bool SOA>
__device__ __forceinline__ void lane_store(T* __restrict__ base, int64_t s, int lane, bool full, bool valid, const T (&q)[E],
                                           unsigned char* sm) {
    if constexpr (SOA) {
        constexpr int W = soa_w<T>();
        T* p = base + (s / W) * (int64_t)(E * W) + (s % W);
        if (valid) {
#pragma unroll
            for (int e = 0; e < E; ++e) p[e * W] = q[e];
        }
    } else {
        if (StageCfg<E * (int)sizeof(T)>::use && full) {
            aos_store_staged<T, E>(base, s - lane, lane, q, sm);
        } else if (valid) {
            aos_store<T, E>(base, s, q);
        }
    }
}

// ================================================================== Cholesky (+ fused solve)
// Row sweep of Cholesky.h:40-56 in registers; one RcpOf per diagonal entry serves the quotients of its column and the two
// substitution sweeps (same bits as div_rn, gmb_math.cuh).  Only the lower triangle of the input is used; the strict upper
// triangle of the factor is zero (:52).
template <typename T, int N, int M, bool SOA, bool RM, bool SOLVE, int MINB>
__global__ void __launch_bounds__(kLaneThreads, MINB)
k_lane_chol(const T* A, const T* Bm, T* X, T* L_out, uint8_t* __restrict__ ok_out, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * kLaneThreads + threadIdx.x;
    if (s - lane >= B) return;
    const bool valid = s < B;
    const bool full = (s - lane) + 32 <= B;

    // AoS: matrix and right-hand sides enter through ONE staging buffer side by side when that fits the static shared-memory
    // limit (both cp.async copies in flight before anybody waits), else one after the other
    constexpr bool STAGE_B = SOLVE && StageCfg<N * M * (int)sizeof(T)>::use;
    constexpr bool OVERLAP = !SOA && (size_t)(N * N + (STAGE_B ? N * M : 0)) * 32 * sizeof(T) * (kLaneThreads / 32) <= 48 * 1024;
    constexpr int EBUF = OVERLAP ? N * N + (STAGE_B ? N * M : 0) : (N > M ? N * N : N * M);
    unsigned char* sm = nullptr;
    if constexpr (!SOA) sm = lane_stage_buffer<T, EBUF>();
    T a[N][N];
    T x[N][M];
    if constexpr (SOA) {
        // the container gives element planes: the strict upper triangle is never read (Cholesky.h:42-44)
        constexpr int W = soa_w<T>();
        const T* p = A + (s / W) * (int64_t)(N * N * W) + (s % W);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = (c <= r && valid) ? __ldg(p + (RM ? N * r + c : N * c + r) * W) : (T)0;
        if constexpr (SOLVE) {
            T qb[N * M];
            lane_load<T, N * M, SOA>(Bm, s, lane, full, valid, qb, sm);
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < M; ++j) x[r][j] = qb[RM ? M * r + j : N * j + r];
        }
    } else {
        T qa[N * N];
        T qb[SOLVE ? N * M : 1];
        if (OVERLAP && full) {
            aos_stage_issue<T, N * N>(A, s - lane, lane, sm);
            if constexpr (STAGE_B) aos_stage_issue<T, N * M>(Bm, s - lane, lane, sm + 32 * N * N * sizeof(T));
            if constexpr (SOLVE && !STAGE_B) aos_load<T, N * M>(Bm, s, qb);
            cp_async_wait_all_();
            __syncwarp();
            aos_stage_read<T, N * N>(lane, qa, sm);
            if constexpr (STAGE_B) aos_stage_read<T, N * M>(lane, qb, sm + 32 * N * N * sizeof(T));
            __syncwarp();
        } else {
            lane_load<T, N * N, false>(A, s, lane, full, valid, qa, sm);
            if constexpr (SOLVE) lane_load<T, N * M, false>(Bm, s, lane, full, valid, qb, sm);
        }
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = qa[RM ? N * r + c : N * c + r];
        if constexpr (SOLVE) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int j = 0; j < M; ++j) x[r][j] = qb[RM ? M * r + j : N * j + r];
        }
    }

    bool ok = true;
    T ry[N];                                                   // reciprocal part of RcpOf per diagonal entry
    bool rs[N];
#pragma unroll
    for (int r = 0; r < N; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            T sum = a[r][c];
#pragma unroll
            for (int k = 0; k < c; ++k) sum = fma_rn(-a[r][k], a[c][k], sum);
            if (r == c) {
                ok = ok && (sum > (T)0);
                sum = sqrt_rn(sum);
                const RcpOf<T> rc(sum);
                ry[r] = rc.y;
                rs[r] = rc.safe;
            } else {
                sum = quot(sum, RcpOf<T>(a[c][c], ry[c], rs[c]));
                a[c][r] = (T)0;
            }
            a[r][c] = sum;
        }
    }
    if constexpr (SOLVE) {
        // L y = b forward, L^T x = y backward (cholesky_solve_regs, gmb_math.cuh)
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const RcpOf<T> rc(a[r][r], ry[r], rs[r]);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                T v = x[r][j];
#pragma unroll
                for (int c = 0; c < r; ++c) v = fma_rn(-x[c][j], a[r][c], v);
                x[r][j] = quot(v, rc);
            }
        }
#pragma unroll
        for (int r = N - 1; r >= 0; --r) {
            const RcpOf<T> rc(a[r][r], ry[r], rs[r]);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                T v = x[r][j];
#pragma unroll
                for (int c = N - 1; c > r; --c) v = fma_rn(-x[c][j], a[c][r], v);
                x[r][j] = quot(v, rc);
            }
        }
        T qx[N * M];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int j = 0; j < M; ++j) qx[RM ? M * r + j : N * j + r] = x[r][j];
        lane_store<T, N * M, SOA>(X, s, lane, full, valid, qx, sm);
    }
    if (L_out != nullptr) {
        T ql[N * N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) ql[RM ? N * r + c : N * c + r] = a[r][c];
        lane_store<T, N * N, SOA>(L_out, s, lane, full, valid, ql, sm);
    }
    if (valid && ok_out) ok_out[s] = ok ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------- launchers
template <typename T, int NMAX>
int lane_chol_impl(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm, bool solve,
                   cudaStream_t st) {
    if (n > NMAX || (solve && (m < 1 || m > 4))) return GMB_ERR_UNSUPPORTED;
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8>(n, [&](auto N) {
        if constexpr (N.value <= NMAX) {
            dispatch_int<1, 2, 3, 4>(solve ? m : 1, [&](auto M) {
                const unsigned grid = (unsigned)ceil_div(B, kLaneThreads);
                if (grid == 0) {
                    rc = GMB_OK;
                    return;
                }
                dispatch_bool(solve, [&](auto SOLVE) {
                    if constexpr (!SOLVE.value && M.value != 1) return;
                    else
                    dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                        dispatch_bool(rm != 0, [&](auto RM) {
                            constexpr int MINB = (N.value * N.value * (int)sizeof(T) / 4 <= 50) ? 4 : 3;
                            k_lane_chol<T, N.value, M.value, SOA.value, RM.value, SOLVE.value, MINB><<<grid, kLaneThreads, 0, st>>>(A, Bm, X, L_out, ok, B);
                        });
                    });
                });
                rc = after_launch();
            });
        }
    });
    return rc;
}

}  // namespace gmb
