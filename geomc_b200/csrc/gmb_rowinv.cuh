// gmb_rowinv.cuh -- row-distributed matrix inverse for n = 5..16: invNxN (MatrixInv.h:219-231) = decomp_plu,
// fail on a degenerate column, back-substitution of the permuted identity with n right-hand sides.
//
// Same decomposition as the fused solve in gmb_rowplu.cuh with the register-only exchange: row p of [M | I] lives
// in lane p % G, slot p / G; the identity rides through the elimination (exchanged in lock-step = "out(i, p[i]) = 1",
// MatrixInv.h:226-228; eliminated with the same multipliers, never skipped = backsolve_lu's forward sweep,
// LUDecomp.h:345-353); the pivot row -- matrix part from column i on, and all n riders -- is collapsed from the
// lane's candidate slots with a select chain, broadcast with one shuffle per column from the pivot's lane, and row i
// is dropped into the pivot's slot with predicated moves.  Backward sweep column by column over all n riders
// (LUDecomp.h:356-365: subtract, then divide by lu(r,r)).
//
// The arithmetic is row-major on the flat buffer IN THE DESTINATION'S LAYOUT (Matrix.h:717 -> MatrixInv.h:243-249
// copies the source into a temporary with the destination's layout first): the kernel loads M' = src when source
// and destination layouts agree and M' = src^T when they differ, and writes M'^-1 flat row-major.
// ok = 0 leaves the destination untouched.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"
#include "gmb_rowplu.cuh"

namespace gmb {

template <typename T, int N, bool SOA, int G, int MINB, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB)
k_row_inv(const T* __restrict__ A, T* __restrict__ Ainv, uint8_t* __restrict__ ok_out, int same_layout, int64_t B) {
    constexpr int RL = row_rl(N, G);
    constexpr int NC = 2 * N;                                  // [M | riders]
    constexpr int V = soa_v<T>();
    constexpr int ST = SOA ? soa_w<T>() : 1;
    constexpr int W = soa_w<T>();
    using VT = typename VecOf<T>::type;

    const int lane_g = (G == 1) ? 0 : (int)(threadIdx.x % G);
    const int gbase = (int)(threadIdx.x & 31) & ~(G - 1);
    int64_t s = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;                                     // keep all lanes for the shuffles
    const int64_t baseA = SOA ? (s / W) * (int64_t)(N * N) * W + (s % W) : s * (int64_t)(N * N);

    // ---- load: slot k holds row p = k*G + lane_g of M' and row p of the identity ---------------------------
    T a[RL][NC];
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int p = k * G + lane_g;
        const bool live = p < N;
        const int pr = live ? p : 0;
        if (!SOA && same_layout && (N * (int)sizeof(T)) % 16 == 0) {
            if constexpr ((N * (int)sizeof(T)) % 16 == 0)
                ldg_row<T, N>(A + baseA + (int64_t)pr * N, a[k], N <= 8 && (reinterpret_cast<uintptr_t>(A) & 31) == 0);   // 256-bit loads: 8 x 8 +6 %, 16 x 16 -15 % (measured)
        } else {
#pragma unroll
            for (int c = 0; c < N; ++c) a[k][c] = A[baseA + (int64_t)(same_layout ? N * pr + c : N * c + pr) * ST];
        }
#pragma unroll
        for (int j = 0; j < N; ++j) a[k][N + j] = (live && j == p) ? (T)1 : (T)0;
    }

    int degenerate = 0;

    // ---- elimination ----------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        const int oi_lane = i % G, oi_slot = i / G;            // static

        // pivot search in column i over positions >= i (LUDecomp.h:205-213)
        T best_v = (T)-2;
        int best_p = 1 << 20;
        auto row_live = [&](int k) { return ((k + 1) * G <= N) || (k * G + lane_g < N); };   // static unless the slot is ragged
        auto at_or_below = [&](int k) { return (k > oi_slot || lane_g >= oi_lane) && row_live(k); };
        auto is_below = [&](int k) { return (k > oi_slot || lane_g > oi_lane) && row_live(k); };
#pragma unroll
        for (int k = oi_slot; k < RL; ++k) {
            const int p = k * G + lane_g;
            const bool cand = at_or_below(k);
            T v = abs_(a[k][i]);
            // a NaN below the diagonal never wins (`v > best_v` is false); a NaN ON the diagonal keeps the pivot
            if (k == oi_slot) v = ((lane_g == oi_lane) && (v != v)) ? InfOf<T>::get() : v;
            const bool take = cand && (v > best_v);
            best_v = take ? v : best_v;
            best_p = take ? p : best_p;
        }
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            const T ov = __shfl_xor_sync(0xffffffffu, best_v, off, G);
            const int op = __shfl_xor_sync(0xffffffffu, best_p, off, G);
            const bool take = (ov > best_v) || (ov == best_v && op < best_p);
            best_v = take ? ov : best_v;
            best_p = take ? op : best_p;
        }
        const int pvt = best_p;
        const bool dead = (best_v == (T)0);                    // :215
        degenerate += dead ? 1 : 0;
        const bool do_swap = pvt != i;
        const int pl = pvt % G, ps = pvt / G;
        const bool mine_is_pvt = do_swap && (lane_g == pl);

        // exchange rows i <-> pvt in registers and broadcast the pivot row (columns i.. and every rider)
        T piv[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c >= i) {
                T v = a[oi_slot][c];
#pragma unroll
                for (int k = oi_slot + 1; k < RL; ++k) v = (ps == k) ? a[k][c] : v;
                const T pc = __shfl_sync(0xffffffffu, v, gbase + pl);
                const T ri = __shfl_sync(0xffffffffu, a[oi_slot][c], gbase + oi_lane);
                piv[c] = pc;
#pragma unroll
                for (int k = oi_slot; k < RL; ++k) {
                    const bool to_pvt = mine_is_pvt && (ps == k);
                    const bool to_i = do_swap && (lane_g == oi_lane) && (k == oi_slot);
                    a[k][c] = to_pvt ? ri : (to_i ? pc : a[k][c]);
                }
            }
        }

        // multipliers and updates of this lane's own rows below the pivot (:232-246); riders never skipped
        const T pv = piv[i];
        T bq[RL];
        // one reciprocal per pivot (RcpOf, gmb_math.cuh) when a lane has three or more rows to divide; with fewer the
        // replicated reciprocal costs what it saves (double n >= 13, eight lanes: -6 .. -17 % measured)
        constexpr bool SRCP = RL >= 3;
        if constexpr (SRCP) {
            const RcpOf<T> rc(pv);
            bool all_ok = true;
#pragma unroll
            for (int k = oi_slot; k < RL; ++k) {
                bool ok;
                bq[k] = rc.fast(a[k][i], ok);
                all_ok = all_ok && (ok || !is_below(k));
            }
            if (!all_ok) {
#pragma unroll
                for (int k = oi_slot; k < RL; ++k) bq[k] = div_rn(a[k][i], pv);
            }
        }
#pragma unroll
        for (int k = oi_slot; k < RL; ++k) {
            const bool below = is_below(k) && !dead;
            const T b = SRCP ? bq[k] : div_rn(a[k][i], pv);
            const bool upd = below && (b != (T)0);
#pragma unroll
            for (int c = i + 1; c < N; ++c) {
                const T v = fma_rn(-b, piv[c], a[k][c]);
                a[k][c] = upd ? v : a[k][c];
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T v = fma_rn(-piv[N + j], b, a[k][N + j]);
                a[k][N + j] = below ? v : a[k][N + j];
            }
        }
    }
    const bool good = degenerate == 0;

    // ---- backward sweep, column-oriented, all n riders (:356-365) ---------------------------------------------
    // Row c's n quotients x(c,j) / u(c,c) all sit in one lane; computed there, every lane of the group would issue
    // all n IEEE divisions (n^2 per system and lane: the dominant cost, above all in double).  Instead the numerators
    // are dealt out -- rider j goes to lane j % G -- each lane divides its share, and the quotients are broadcast
    // back: 2n + 1 shuffles and n / G divisions per column instead of n shuffles and n divisions.
    constexpr int JL = (N + G - 1) / G;
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        const int oc_lane = c % G, oc_slot = c / G;            // static
        if constexpr (G == 1) {
            const RcpOf<T> rc(a[oc_slot][c]);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T q = quot(a[oc_slot][N + j], rc);
                a[oc_slot][N + j] = q;
#pragma unroll
                for (int k = 0; k <= oc_slot; ++k) {
                    const T v = fma_rn(-q, a[k][c], a[k][N + j]);
                    a[k][N + j] = (k < c) ? v : a[k][N + j];
                }
            }
        } else {
            const T ucc = __shfl_sync(0xffffffffu, a[oc_slot][c], gbase + oc_lane);
            T num[JL], quo[JL];
#pragma unroll
            for (int jj = 0; jj < JL; ++jj) num[jj] = (T)1;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T t = __shfl_sync(0xffffffffu, a[oc_slot][N + j], gbase + oc_lane);
                num[j / G] = (lane_g == j % G) ? t : num[j / G];
            }
            if constexpr (JL < 3) {
#pragma unroll
                for (int jj = 0; jj < JL; ++jj) quo[jj] = div_rn(num[jj], ucc);
            } else {
                const RcpOf<T> rc(ucc);
                bool all_ok = true;
#pragma unroll
                for (int jj = 0; jj < JL; ++jj) {
                    bool ok;
                    quo[jj] = rc.fast(num[jj], ok);
                    all_ok = all_ok && ok;
                }
                if (!all_ok) {
#pragma unroll
                    for (int jj = 0; jj < JL; ++jj) quo[jj] = div_rn(num[jj], ucc);
                }
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T xc = __shfl_sync(0xffffffffu, quo[j / G], gbase + j % G);
                if (lane_g == oc_lane) a[oc_slot][N + j] = xc;
#pragma unroll
                for (int k = 0; k <= oc_slot; ++k) {
                    const int p = k * G + lane_g;
                    const T v = fma_rn(-xc, a[k][c], a[k][N + j]);
                    a[k][N + j] = (p < c) ? v : a[k][N + j];
                }
            }
        }
    }

    // ---- store M'^-1 flat row-major; nothing is written for a degenerate system -------------------------------
    if (valid && good) {
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int p = k * G + lane_g;
            if (p < N) {
                if constexpr (!SOA && (N * (int)sizeof(T)) % 16 == 0) {
                    VT* dst = reinterpret_cast<VT*>(Ainv + baseA + (int64_t)p * N);
#pragma unroll
                    for (int q = 0; q < N / V; ++q) {
                        T t[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) t[v] = a[k][N + q * V + v];
                        dst[q] = pack(t);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < N; ++j) Ainv[baseA + (int64_t)(N * p + j) * ST] = a[k][N + j];
                }
            }
        }
    }
    if (valid && lane_g == 0 && ok_out) ok_out[s] = good ? 1 : 0;
}

// Lanes per system, measured (profiles/r1j_row_inverse.md, 2^19 systems; odd sizes take the neighbouring even size's
// choice among the groups whose row block [M | I] stays within 200 register words).
template <typename T, int N> struct RowInvCfg {
    static constexpr int gf[17] = {0, 0, 0, 0, 0, 2, 2, 2, 2, 2, 2, 2, 4, 4, 4, 4, 4};
    static constexpr int gd[17] = {0, 0, 0, 0, 0, 2, 2, 2, 2, 2, 2, 4, 4, 8, 8, 8, 8};
    static constexpr int G = sizeof(T) == 4 ? gf[N] : gd[N];
};

template <typename T>
int rowinv_impl(int n, int64_t B, const T* A, T* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm, cudaStream_t st) {
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        if (B == 0) {
            rc = GMB_OK;
            return;
        }
        constexpr int G = RowInvCfg<T, N.value>::G;
        constexpr int THREADS = 128;
        const unsigned grid = (unsigned)ceil_div(B * G, THREADS);
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            k_row_inv<T, N.value, SOA.value, G, 2, THREADS><<<grid, THREADS, 0, st>>>(A, Ainv, ok, (src_rm != 0) == (dst_rm != 0) ? 1 : 0, B);
        });
        rc = after_launch();
    });
    return rc;
}

}  // namespace gmb
