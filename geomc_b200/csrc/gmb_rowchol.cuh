// gmb_rowchol.cuh -- row-distributed cooperative Cholesky (and fused Cholesky + solve), n = 5..16.
//
// cholesky<T,RowMajor> (Cholesky.h:37-58) sweeps rows: s = A(r,c) - sum_{k<c} L(r,k) L(c,k), k
// ascending; diagonal: ok &= s > 0, L(r,r) = sqrt(s); off-diagonal: L(r,c) = s / L(c,c); the strict
// upper triangle is zeroed and only the lower triangle of the input is read.  The right-looking order
// used here applies, for k = 0..n-1, a(r,c) = fma(-L(r,k), L(c,k), a(r,c)) to every r >= c > k: each
// element receives the reference's terms in the reference's order (k ascending) -> bit-identical.
//
// Rows are distributed like in gmb_rowplu.cuh: row p of [A | b] lives in lane p % G, slot p / G, so
// every lane divides and updates only its own rows (no redundant divisions); there is no pivoting,
// hence no row exchange and no shared memory.  Per step k: the owner of row k broadcasts
// d = sqrt(a(k,k)); every lane forms L(r,k) = a(r,k) / d for its rows; column k of L is then
// broadcast with n-1-k shuffles (static source lane and register) and each lane updates the part of
// its rows inside the lower triangle.
//
// Fused solve (framework-defined, same as the n <= 4 kernel and oracle.c): forward L y = b during
// the factorization (y(k) = b(k) / L(k,k); b(r) = fma(-y(k), L(r,k), b(r))), backward L^T x = y column
// by column: x(c) = y(c) / L(c,c), then y(r) = fma(-x(c), L(c,r), y(r)) for r < c, where L(c,r) comes
// from the owner of row c by shuffle and is consumed by the (static) owner of row r.
#pragma once
#include "gmb_rowplu.cuh"

namespace gmb {

// Lanes per system.  The PLU row kernels' rule (row_group: whole rows of [A | b] within 140 words) is far too cautious here: only
// the lower triangle is ever held, n (n + 1) / 2 + n words.  In the SoA container, where a lane reads its planes coalesced whatever
// their number, ONE lane per system therefore serves every float size and double n <= 11 (no shuffles at all; round 2, run r2ad:
// double 8x8 Cholesky + solve 0.47 -> 0.92 of the HBM peak, float 12x12 0.47 -> 0.84: profiles/r2ad_lane_cholesky.md).  AoS batches take one lane where the warp's staging buffer
// (whole records, n * n + n elements per lane) still leaves two CTAs per SM: float n <= 14, double n <= 9.
#ifndef GMB_CHOL_LANE_WORDS
#define GMB_CHOL_LANE_WORDS 210
#endif
#ifndef GMB_CHOL_LANE_AOS_BYTES
#define GMB_CHOL_LANE_AOS_BYTES 840       // AoS: bytes of [A | b] per system that the warp's staging buffer may take (two CTAs per SM)
#endif
// SoA batches whose lower-triangle planes are staged through shared memory (see the load section of k_row_chol).  Measured for
// every size (run r2bb, fused Cholesky + solve, fraction of the HBM peak, direct plane loads -> staged): float n = 12..16 0.85 / 0.73 /
// 0.70 / 0.54 / 0.53 -> 0.88 / 0.76 / 0.75 / 0.57 / 0.57, double n = 10..12 0.71 / 0.74 / 0.74 -> 0.76 / 0.80 / 0.78; the smaller sizes, at or
// near the roofline with direct loads, lose 0.03 .. 0.13 to the extra shared-memory round trip and keep them.
constexpr bool chol_soa_stage(int n, int words, bool soa, int g) {
    return soa && g == 1 && n >= (words == 1 ? 12 : 10);
}
constexpr int chol_group(int n, int words, bool soa) {
    if (words * (n * (n + 1) / 2 + 2 * n) <= GMB_CHOL_LANE_WORDS && (soa || (n * n + n) * words * 4 <= GMB_CHOL_LANE_AOS_BYTES)) return 1;
    return row_group(n, words);
}

template <typename T, int N, bool SOLVE, bool SOA, bool RM>
__global__ void __launch_bounds__(kRowThreads, GMB_ROW_MIN_BLOCKS)
k_row_chol(const T* A, const T* Bm, T* X, T* L_out, uint8_t* __restrict__ ok_out, int64_t B) {
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int G = chol_group(N, WORDS, SOA);
    constexpr int RL = row_rl(N, G);
    constexpr int ST = SOA ? soa_w<T>() : 1;
    constexpr int W = soa_w<T>();
    const int lane_g = (G == 1) ? 0 : (int)(threadIdx.x % G);
    int64_t s = ((int64_t)blockIdx.x * kRowThreads + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;
    const int64_t baseA = SOA ? (s / W) * (int64_t)(N * N) * W + (s % W) : s * (int64_t)(N * N);
    const int64_t baseX = SOA ? (s / W) * (int64_t)N * W + (s % W) : s * (int64_t)N;

    // ---- load the lower triangle of this lane's rows (and b) ------------------------------------
    T a[RL][N];
    T y[RL];
    // One lane per system on an AoS batch: the warp's 32 records travel as one contiguous range through a warp-private
    // shared-memory transpose (coalesced 128-bit accesses on both sides, gmb_io.cuh) -- records whose rows are not
    // multiples of 16 bytes (n = 5, 6, 7, 9, ...) used to be read with scalar loads at a stride of one record.
    constexpr bool LANE_IO = (G == 1) && !SOA;
    extern __shared__ __align__(16) unsigned char gmb_dyn_smem[];
    unsigned char* lane_sm = gmb_dyn_smem + (size_t)(threadIdx.x >> 5) * (32 * (N * N + N) * sizeof(T));
    const int64_t s_raw = (int64_t)blockIdx.x * kRowThreads + threadIdx.x;
    const bool lane_full = LANE_IO && ((s_raw - (threadIdx.x & 31)) + 32 <= B);
    if constexpr (LANE_IO) {
        T qa[N * N];
        T qb[N];
        constexpr bool STAGE_B = SOLVE && StageCfg<N * (int)sizeof(T)>::use;
        if (lane_full) {
            // matrix and right-hand side: both cp.async copies are in flight before anybody waits
            aos_stage_issue<T, N * N>(A, s_raw - (threadIdx.x & 31), threadIdx.x & 31, lane_sm);
            if constexpr (STAGE_B) aos_stage_issue<T, N>(Bm, s_raw - (threadIdx.x & 31), threadIdx.x & 31, lane_sm + 32 * N * N * sizeof(T));
            if constexpr (SOLVE && !STAGE_B) aos_load<T, N>(Bm, s, qb);
            cp_async_wait_all_();
            __syncwarp();
            aos_stage_read<T, N * N>(threadIdx.x & 31, qa, lane_sm);
            if constexpr (STAGE_B) aos_stage_read<T, N>(threadIdx.x & 31, qb, lane_sm + 32 * N * N * sizeof(T));
            __syncwarp();
        } else {
            aos_load<T, N * N>(A, s, qa);
            if (SOLVE) aos_load<T, N>(Bm, s, qb);
        }
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) a[r][c] = (c <= r) ? qa[RM ? N * r + c : N * c + r] : (T)0;
#pragma unroll
        for (int r = 0; r < N; ++r) y[r] = SOLVE ? qb[r] : (T)0;
    } else {
    // SoA, one lane per system, the larger sizes (chol_soa_stage): the warp's 32-system segments of the LOWER-TRIANGLE planes (a
    // contiguous run of planes per row / column) and of b go global -> shared with cp.async, packed [plane][lane]; no registers
    // held and no load-queue slots taken by up to 136 + 16 direct plane loads per thread while the data travel
    bool soa_staged = false;
    if constexpr (chol_soa_stage(N, WORDS, SOA, G)) {
        const int lane = threadIdx.x & 31;
        const int64_t s_w0 = s_raw - lane;
        if (s_w0 + 32 <= B) {                                  // warp-uniform; the 32 systems sit inside one tile (W % 32 == 0)
            constexpr int V = soa_v<T>();
            constexpr int CPP = 32 / V;                        // 16-byte chunks per plane segment
            constexpr int TRI = N * (N + 1) / 2;
            T* sm = reinterpret_cast<T*>(gmb_dyn_smem) + (size_t)(threadIdx.x >> 5) * (32 * (TRI + N));
            const T* gA = A + (s_w0 / W) * (int64_t)(N * N) * W + (s_w0 % W);
            const T* gB = Bm + (s_w0 / W) * (int64_t)N * W + (s_w0 % W);
#pragma unroll
            for (int k = 0; k <= N; ++k) {
                if (k == N && !SOLVE) break;
                // run k < N: row k, columns 0..k (row-major) or column k, rows k..N-1 (column-major); run N: the N planes of b
                const int first = k == N ? 0 : (RM ? N * k : N * k + k);
                const int cnt = k == N ? N : (RM ? k + 1 : N - k);
                const int off = k == N ? TRI : (RM ? k * (k + 1) / 2 : k * N - k * (k - 1) / 2);
                const T* g = k == N ? gB : gA;
#pragma unroll
                for (int c0 = 0; c0 < cnt * CPP; c0 += 32) {
                    const int c = c0 + lane;
                    if (c < cnt * CPP) {
                        const int e = c / CPP, w = c % CPP;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + (off + e) * 32 + w * V);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + (int64_t)(first + e) * W + w * V)
                                     : "memory");
                    }
                }
            }
            cp_async_wait_all_();
            __syncwarp();
#pragma unroll
            for (int r = 0; r < N; ++r) {
#pragma unroll
                for (int c = 0; c < N; ++c)
                    a[r][c] = (c <= r) ? sm[(RM ? r * (r + 1) / 2 + c : c * N - c * (c - 1) / 2 + (r - c)) * 32 + lane] : (T)0;
                y[r] = SOLVE ? sm[(TRI + r) * 32 + lane] : (T)0;
            }
            soa_staged = true;
        }
    }
    if (!soa_staged) {
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int p = k * G + lane_g;
        const int pr = p < N ? p : 0;
        if constexpr (!SOA && RM && (N * (int)sizeof(T)) % 16 == 0) {
            // whole row with 128- / 256-bit loads (per-lane scalar loads at a stride of one record use 1/8 of every
            // sector they touch); the upper triangle is never looked at (Cholesky.h:40-56 reads the lower one only)
            ldg_row<T, N>(A + baseA + (int64_t)pr * N, a[k], (reinterpret_cast<uintptr_t>(A) & 31) == 0);
#pragma unroll
            for (int c = 0; c < N; ++c) a[k][c] = (c <= pr) ? a[k][c] : (T)0;
        } else {
#pragma unroll
            for (int c = 0; c < N; ++c) {
                a[k][c] = (T)0;
                if (c <= k * G + (G - 1)) {                    // static: some lane's row in this slot reaches column c
                    if (c <= pr) a[k][c] = A[baseA + (int64_t)(RM ? N * pr + c : N * c + pr) * ST];
                }
            }
        }
        y[k] = SOLVE ? Bm[baseX + (int64_t)pr * ST] : (T)0;
    }
    }
    }

    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int ok_lane = k % G, ok_slot = k / G;            // static owner of row k
        // diagonal (Cholesky.h:46-48)
        const T sdiag = (G == 1) ? a[ok_slot][k] : __shfl_sync(0xffffffffu, a[ok_slot][k], ok_lane, G);
        ok = ok && (sdiag > (T)0);
        const T d = sqrt_rn(sdiag);
        if (lane_g == ok_lane) a[ok_slot][k] = d;
        // column k of L for this lane's rows below k (Cholesky.h:50)
        // one reciprocal of L(k,k) for the whole column and the rider (RcpOf, gmb_math.cuh): same bits as div_rn
        const RcpOf<T> rcd(d);
        T vq[RL];
        T ykq = (T)0;
        {
            bool all_ok = true;
#pragma unroll
            for (int sl = 0; sl < RL; ++sl) {
                if (sl >= ok_slot) {                           // static
                    const int p = sl * G + lane_g;
                    bool okq;
                    vq[sl] = rcd.fast(a[sl][k], okq);
                    all_ok = all_ok && (okq || !(p > k && p < N));
                }
            }
            if (SOLVE) {
                bool okq;
                ykq = rcd.fast(y[ok_slot], okq);
                all_ok = all_ok && (okq || lane_g != ok_lane);
            }
            if (!all_ok) {
#pragma unroll
                for (int sl = 0; sl < RL; ++sl)
                    if (sl >= ok_slot) vq[sl] = div_rn(a[sl][k], d);
                if (SOLVE) ykq = div_rn(y[ok_slot], d);
            }
        }
#pragma unroll
        for (int sl = 0; sl < RL; ++sl) {
            if (sl >= ok_slot) {                               // static
                const int p = sl * G + lane_g;
                const T v = vq[sl];
                if (p > k && p < N) a[sl][k] = v;
            }
        }
        if (SOLVE) {
            // forward substitution: y(k) = b(k) / L(k,k), then b(r) -= y(k) L(r,k) for r > k
            const T yk = (G == 1) ? ykq : __shfl_sync(0xffffffffu, ykq, ok_lane, G);
            if (lane_g == ok_lane) y[ok_slot] = yk;
#pragma unroll
            for (int sl = 0; sl < RL; ++sl) {
                if (sl >= ok_slot) {                           // static
                    const int p = sl * G + lane_g;
                    const T v = fma_rn(-yk, a[sl][k], y[sl]);
                    if (p > k && p < N) y[sl] = v;
                }
            }
        }
        // trailing update: a(r,c) -= L(r,k) L(c,k) for k < c <= r
#pragma unroll
        for (int c = 0; c < N; ++c) {
            if (c > k) {                                       // static
                const T lck = (G == 1) ? a[c / G][k] : __shfl_sync(0xffffffffu, a[c / G][k], c % G, G);   // L(c,k), static source
#pragma unroll
                for (int sl = 0; sl < RL; ++sl) {
                    if (sl >= c / G) {                         // static: lower slots hold rows < c
                        const int p = sl * G + lane_g;
                        const T v = fma_rn(-a[sl][k], lck, a[sl][c]);
                        if (p >= c && p < N) a[sl][c] = v;
                    }
                }
            }
        }
    }

    // ---- L with the strict upper triangle zeroed (Cholesky.h:52) ---------------------------------
    if constexpr (LANE_IO) {
        if (L_out != nullptr) {
            T ql[N * N];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) ql[RM ? N * r + c : N * c + r] = (c <= r) ? a[r][c] : (T)0;
            if (StageCfg<N * N * (int)sizeof(T)>::use && lane_full) {
                aos_store_staged<T, N * N>(L_out, s_raw - (threadIdx.x & 31), threadIdx.x & 31, ql, lane_sm);
            } else if (valid) {
                aos_store<T, N * N>(L_out, s, ql);
            }
        }
    } else
    if (L_out != nullptr && valid) {
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int p = k * G + lane_g;
            if (p < N) {
#pragma unroll
                for (int c = 0; c < N; ++c)
                    L_out[baseA + (int64_t)(RM ? N * p + c : N * c + p) * ST] = (c <= p) ? a[k][c] : (T)0;
            }
        }
    }

    // ---- backward sweep L^T x = y -----------------------------------------------------------------
    if (SOLVE) {
#pragma unroll
        for (int c = N - 1; c >= 0; --c) {
            const int oc_lane = c % G, oc_slot = c / G;        // static
            const T q = div_rn(y[oc_slot], a[oc_slot][c]);
            const T xc = (G == 1) ? q : __shfl_sync(0xffffffffu, q, oc_lane, G);
            if (lane_g == oc_lane) y[oc_slot] = xc;
#pragma unroll
            for (int r = N - 1; r >= 0; --r) {
                if (r < c) {                                   // static
                    // L(c, r) sits in the owner of row c; the (static) owner of row r consumes it
                    const T lcr = (G == 1) ? a[oc_slot][r] : __shfl_sync(0xffffffffu, a[oc_slot][r], oc_lane, G);
                    const T v = fma_rn(-xc, lcr, y[r / G]);
                    if (lane_g == r % G) y[r / G] = v;
                }
            }
        }
        if constexpr (LANE_IO) {
            T qx[N];
#pragma unroll
            for (int r = 0; r < N; ++r) qx[r] = y[r];
            if (StageCfg<N * (int)sizeof(T)>::use && lane_full) {
                aos_store_staged<T, N>(X, s_raw - (threadIdx.x & 31), threadIdx.x & 31, qx, lane_sm);
            } else if (valid) {
                aos_store<T, N>(X, s, qx);
            }
        } else
        if (valid) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int p = k * G + lane_g;
                if (p < N) X[baseX + (int64_t)p * ST] = y[k];
            }
        }
    }
    if (valid && lane_g == 0 && ok_out) ok_out[s] = ok ? 1 : 0;
}

template <typename T>
int rowchol_impl(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm, bool solve,
                 cudaStream_t st) {
    if (solve && m != 1) return GMB_ERR_UNSUPPORTED;
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        if (B == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_bool(solve, [&](auto SOLVE) {
            dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                constexpr int G = chol_group(N.value, (int)sizeof(T) / 4, SOA.value);
                const unsigned grid = (unsigned)ceil_div(B * G, kRowThreads);
                dispatch_bool(rm != 0, [&](auto RM) {
                    auto kern = k_row_chol<T, N.value, SOLVE.value, SOA.value, RM.value>;
                    // one lane per system on an AoS batch: one staging buffer of 32 records per warp
                    const size_t smem = (G == 1 && !SOA.value) ? (size_t)(kRowThreads / 32) * 32 * (N.value * N.value + N.value) * sizeof(T)
                                        : chol_soa_stage(N.value, (int)sizeof(T) / 4, SOA.value, G)
                                            ? (size_t)(kRowThreads / 32) * 32 * (N.value * (N.value + 1) / 2 + N.value) * sizeof(T)
                                            : 0;
                    if (smem > 48 * 1024) {
                        static std::atomic<uint64_t> raised{0};
                        raise_dyn_smem(kern, smem, raised);
                    }
                    // (an L2 prefetch of the next wave's lower-triangle planes, 512 bytes per plane, was measured in run r2af and
                    // lost everywhere -- double 8x8 0.91 -> 0.74, float 12x12 0.84 -> 0.65 of the HBM peak: profiles/r2ad_lane_cholesky.md)
                    kern<<<grid, kRowThreads, smem, st>>>(A, Bm, X, L_out, ok, B);
                });
            });
        });
        rc = after_launch();
    });
    return rc;
}

}  // namespace gmb

namespace gmb {

// ---------------------------------------------------------------------------------------------------
// backsolve_lu / backsolve_plu (LUDecomp.h:332-366 / 271-312) for n = 5..16, one right-hand side, rows
// distributed as above.  PLU = true gathers y(r) = b(perm[r]) first (:294).  Forward sweep column by
// column (c ascending: x(r) = fma(-x(c), lu(r,c), x(r)) for r > c -- for every row the terms arrive in
// the reference's order), backward sweep as in k_row_solve (c descending, then divide by u(c,c)).
template <typename T, int N, int MODE, bool SOA, bool RM>
__global__ void __launch_bounds__(kRowThreads, GMB_ROW_MIN_BLOCKS)
k_row_backsolve(const T* __restrict__ LUm, const uint8_t* __restrict__ perm, const T* Bm, T* X, int skip, int m, int64_t B) {
    // MODE 0 = backsolve_lu, 1 = backsolve_plu (gathers b(perm[r]) first), 2 = destructive_apply_permutation (gather only).
    // m right-hand sides: the LU rows are loaded once, the columns of X go through the sweeps one after the other; element
    // (r, j) of the n x m block sits at RM ? m*r + j : n*j + r (MatrixGlue.h:28-58).
    constexpr bool PLU = MODE != 0;
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int G = row_group(N, WORDS);
    constexpr int RL = row_rl(N, G);
    constexpr int ST = SOA ? soa_w<T>() : 1;
    constexpr int W = soa_w<T>();
    const int lane_g = (G == 1) ? 0 : (int)(threadIdx.x % G);
    int64_t s = ((int64_t)blockIdx.x * kRowThreads + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;
    const int64_t baseA = SOA ? (s / W) * (int64_t)(N * N) * W + (s % W) : s * (int64_t)(N * N);
    const int64_t baseX = SOA ? (s / W) * (int64_t)(N * m) * W + (s % W) : s * (int64_t)(N * m);

    T a[RL][N];
    int src[RL];
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int p = k * G + lane_g;
        const int pr = p < N ? p : 0;
        if constexpr (MODE != 2) {
            if constexpr (!SOA && RM && (N * (int)sizeof(T)) % 16 == 0) {
                ldg_row<T, N>(LUm + baseA + (int64_t)pr * N, a[k], (reinterpret_cast<uintptr_t>(LUm) & 31) == 0);   // whole row, 128 / 256 bit
            } else {
#pragma unroll
                for (int c = 0; c < N; ++c) a[k][c] = LUm[baseA + (int64_t)(RM ? N * pr + c : N * c + pr) * ST];
            }
        }
        src[k] = PLU ? (int)perm[s * N + pr] : pr;
    }
    for (int j = 0; j < m; ++j) {
        T x[RL];
#pragma unroll
        for (int k = 0; k < RL; ++k) x[k] = Bm[baseX + (int64_t)(RM ? m * src[k] + j : N * j + src[k]) * ST];
        if constexpr (MODE == 2) __syncwarp();                 // in place: every source row is read before any is written
        if constexpr (MODE != 2) {
            // forward: L y = b (unit diagonal)
#pragma unroll
            for (int c = 0; c < N - 1; ++c) {
                const T xc = (G == 1) ? x[c / G] : __shfl_sync(0xffffffffu, x[c / G], c % G, G);
#pragma unroll
                for (int sl = 0; sl < RL; ++sl) {
                    if (sl >= c / G) {                                 // static
                        const int p = sl * G + lane_g;
                        const T v = fma_rn(-xc, a[sl][c], x[sl]);
                        if (p > c && p < N) x[sl] = v;
                    }
                }
            }
            // backward: U x = y
#pragma unroll
            for (int c = N - 1; c >= 0; --c) {
                const int oc_lane = c % G, oc_slot = c / G;            // static
                const bool on = c >= skip;
                const T q = div_rn(x[oc_slot], a[oc_slot][c]);
                const T xc = (G == 1) ? q : __shfl_sync(0xffffffffu, q, oc_lane, G);
                if (on && lane_g == oc_lane) x[oc_slot] = q;
#pragma unroll
                for (int sl = 0; sl < RL; ++sl) {
                    if (sl <= oc_slot) {                               // static
                        const int p = sl * G + lane_g;
                        const T v = fma_rn(-xc, a[sl][c], x[sl]);
                        if (on && p < c && p >= skip) x[sl] = v;
                    }
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int p = k * G + lane_g;
                if (p < N) X[baseX + (int64_t)(RM ? m * p + j : N * j + p) * ST] = x[k];
            }
        }
    }
}

template <typename T>
int rowbacksolve_impl(int mode, int n, int m, int64_t B, const T* LU, const uint8_t* perm, const T* Bm, T* X, int skip, int layout,
                      int rm, cudaStream_t st) {
    if (mode < 0 || mode > 2 || m < 1) return GMB_ERR_UNSUPPORTED;
    if (mode == 1 && X == Bm) return GMB_ERR_UNSUPPORTED;        // the gather of backsolve_plu needs distinct buffers (as in the reference)
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        constexpr int G = row_group(N.value, (int)sizeof(T) / 4);
        const unsigned grid = (unsigned)ceil_div(B * G, kRowThreads);
        if (grid == 0) {
            rc = GMB_OK;
            return;
        }
        dispatch_int<0, 1, 2>(mode, [&](auto MODE) {
            dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                dispatch_bool(rm != 0, [&](auto RM) {
                    k_row_backsolve<T, N.value, MODE.value, SOA.value, RM.value><<<grid, kRowThreads, 0, st>>>(LU, perm, Bm, X, skip, m, B);
                });
            });
        });
        rc = after_launch();
    });
    return rc;
}

}  // namespace gmb
