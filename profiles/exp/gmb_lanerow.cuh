// gmb_lanerow.cuh -- fused PLU solve (linear_solve, LUDecomp.h:388-399) for n = 9..16 with the rows spread over
// the LANES of a group and never moved: "implicit pivoting".
//
// Why: a register file cannot be indexed dynamically, and partial pivoting needs exactly that -- "row pvt" is
// only known at run time.  The row-distributed kernel (gmb_rowplu.cuh, few lanes, many rows per lane) pays for
// every exchange with predicated shared-memory traffic over all candidate slots (ncu: 10 800 thread instructions
// per 16x16 system, LSU/MIO-bound at 35 % issue utilisation); here the only dynamic index is a LANE index, which a
// shuffle resolves for free:
//   * a group of G lanes owns one system, R = 16 / G rows per lane; a row stays in the (lane, slot) it was loaded
//     into and carries its current POSITION p (the index the reference's physical row exchanges would have given
//     it; -1 for padding slots);
//   * pivot search (LUDecomp.h:205-213): every lane reduces its own slots to (key, position), key = the bits of
//     |a(p,i)| (monotonic for non-NaN floats; NaN -> 0 so it never wins, except on the diagonal row where it maps
//     to the maximum so that it is never displaced); two `redux.sync` instructions give the group maximum and,
//     among its holders, the lowest position -- the reference's first strict maximum -- together with the
//     winner's lane and slot;
//   * "exchange" (:222-228) = two position updates: the pivot row takes position i, the row that held position i
//     takes the pivot's old position;
//   * the pivot row is broadcast with one shuffle per column from the winner's lane; every lane divides and
//     updates its own live rows: b = a(r,i) / a(i,i), a(r,c) = fma(-b, a(i,c), a(r,c)) unless b == 0 (:232-246),
//     the right-hand side rides along unconditionally (= backsolve_lu's forward sweep, :345-353);
//   * zero pivot (:215-219): count it, touch nothing.
// After the elimination the rows go through shared memory once, addressed by position (a dynamic MEMORY address
// is free), and come back in position order for the statically indexed backward sweep (:356-365) and the stores.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"

namespace gmb {

constexpr int kLaneThreads = 128;

template <typename T> struct KeyOf;
template <> struct KeyOf<float> {
    using type = uint32_t;
    __device__ static __forceinline__ uint32_t abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
    static constexpr uint32_t inf = 0x7f800000u, top = 0xffffffffu;
};

// T = float only for now (the key reduction uses the 32-bit redux.sync).
template <typename T, int N, int G, bool SOA, bool RM, int MINB>
__global__ void __launch_bounds__(kLaneThreads, MINB)
k_lane_solve(const T* __restrict__ A, const T* Bm, T* X, T* __restrict__ LU, uint8_t* __restrict__ ok_out, int skip, int64_t B) {
    static_assert(G == 4 || G == 8 || G == 16, "group size");
    constexpr int R = 16 / G;                                    // row slots per lane (N <= 16)
    constexpr int V = soa_v<T>();
    constexpr int ST = SOA ? soa_w<T>() : 1;
    constexpr int W = soa_w<T>();
    constexpr int SYS_PER_WARP = 32 / G;
    constexpr int ROWSTRIDE = 20;                                // 16 + rider, padded to a multiple of 4: 80-byte rows
    using VT = typename VecOf<T>::type;

    __shared__ __align__(16) T stage[(kLaneThreads / 32) * SYS_PER_WARP * 16 * ROWSTRIDE];

    const int lane = (int)(threadIdx.x & 31);
    const int lg = lane & (G - 1);                               // lane within the group
    int64_t s = ((int64_t)blockIdx.x * kLaneThreads + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;                                       // keep whole groups for the collectives
    const int64_t baseA = SOA ? (s / W) * (int64_t)(N * N) * W + (s % W) : s * (int64_t)(N * N);
    const int64_t baseX = SOA ? (s / W) * (int64_t)N * W + (s % W) : s * (int64_t)N;

    // ---- load: slot k of lane lg holds row p = lg * R + k (position = row index) -----------------------------
    T a[R][N], x[R];
    int pos[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int p = lg * R + k;
        const bool live = p < N;
        const int pr = live ? p : 0;
        pos[k] = live ? p : -1;
        if constexpr (!SOA && RM && (N * (int)sizeof(T)) % 16 == 0) {
            const VT* src = reinterpret_cast<const VT*>(A + baseA + (int64_t)pr * N);
#pragma unroll
            for (int q = 0; q < N / V; ++q) {
                T t[V];
                unpack(ldg_stream(src + q), t);
#pragma unroll
                for (int v = 0; v < V; ++v) a[k][q * V + v] = t[v];
            }
        } else {
#pragma unroll
            for (int c = 0; c < N; ++c) a[k][c] = A[baseA + (int64_t)(RM ? N * pr + c : N * c + pr) * ST];
        }
        x[k] = Bm[baseX + (int64_t)pr * ST];
    }

    int degenerate = 0;

    // ---- elimination ------------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        // this lane's best candidate: larger key, then lower position
        uint32_t bkey = 0;
        int bpos = 1 << 20, bslot = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const T v = a[k][i];
            uint32_t key = KeyOf<T>::abs_bits(v);
            if (key > KeyOf<T>::inf) key = (pos[k] == i) ? KeyOf<T>::top : 0u;     // NaN
            key = (pos[k] >= i) ? key : 0u;                                        // finished rows and padding
            const bool take = (key > bkey) || (key == bkey && pos[k] >= i && pos[k] < bpos);
            bkey = take ? key : bkey;
            bpos = take ? pos[k] : bpos;
            bslot = take ? k : bslot;
        }
        // group maximum of the keys, then the lowest (position, slot, lane) among its holders: XOR butterflies stay
        // inside the aligned group of G lanes (redux.sync with per-group masks takes a slow serialised path)
        uint32_t m = bkey;
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
        const bool dead = (m == 0u);                                               // biggest == 0 (:215)
        uint32_t win = (bkey == m) ? (((uint32_t)bpos << 8) | ((uint32_t)bslot << 5) | (uint32_t)lane) : 0xffffffffu;
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) win = min(win, __shfl_xor_sync(0xffffffffu, win, off));
        const int wpos = (int)(win >> 8), ps = (int)((win >> 5) & 7u), pl = (int)(win & 31u);
        degenerate += dead ? 1 : 0;

        // pivot row -> everybody (columns i.. and the rider)
        T piv[N], pivx;
#pragma unroll
        for (int c = i; c < N; ++c) {
            T v = a[0][c];
#pragma unroll
            for (int k = 1; k < R; ++k) v = (ps == k) ? a[k][c] : v;
            piv[c] = __shfl_sync(0xffffffffu, v, pl);
        }
        {
            T v = x[0];
#pragma unroll
            for (int k = 1; k < R; ++k) v = (ps == k) ? x[k] : v;
            pivx = __shfl_sync(0xffffffffu, v, pl);
        }

        const T pv = piv[i];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            // the exchange, as bookkeeping: pivot row -> position i, the row at position i -> the pivot's position
            const bool is_piv = (lane == pl) && (ps == k);
            const int np = is_piv ? i : ((pos[k] == i) ? wpos : pos[k]);
            pos[k] = dead ? pos[k] : np;
            const bool elim = !dead && (pos[k] > i);
            const T b = div_rn(a[k][i], pv);                                       // :234
            const bool upd = elim && (b != (T)0);                                  // :235
#pragma unroll
            for (int c = i + 1; c < N; ++c) {
                const T v = fma_rn(-b, piv[c], a[k][c]);                           // :241
                a[k][c] = upd ? v : a[k][c];
            }
            a[k][i] = elim ? b : a[k][i];                                          // :245
            const T xv = fma_rn(-pivx, b, x[k]);                                   // forward sweep, never skipped (:350)
            x[k] = elim ? xv : x[k];
        }
    }
    const bool good = degenerate == 0;

    // ---- rows into position order (through shared memory, addressed by position) -----------------------------
    T* sys = stage + ((threadIdx.x >> 5) * SYS_PER_WARP + (lane / G)) * (16 * ROWSTRIDE);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (pos[k] >= 0) {
            T* row = sys + pos[k] * ROWSTRIDE;
#pragma unroll
            for (int q = 0; q < N / V; ++q) {
                T t[V];
#pragma unroll
                for (int v = 0; v < V; ++v) t[v] = a[k][q * V + v];
                reinterpret_cast<VT*>(row)[q] = pack(t);
            }
#pragma unroll
            for (int c = N / V * V; c < N; ++c) row[c] = a[k][c];
            row[16] = x[k];
        }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int p = lg * R + k;
        const T* row = sys + (p < N ? p : 0) * ROWSTRIDE;
#pragma unroll
        for (int q = 0; q < N / V; ++q) {
            T t[V];
            unpack(reinterpret_cast<const VT*>(row)[q], t);
#pragma unroll
            for (int v = 0; v < V; ++v) a[k][q * V + v] = t[v];
        }
#pragma unroll
        for (int c = N / V * V; c < N; ++c) a[k][c] = row[c];
        x[k] = row[16];
    }

    // ---- optional: the packed LU the reference leaves in `a` -------------------------------------------------
    if (LU != nullptr && valid) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int p = lg * R + k;
            if (p < N) {
                if constexpr (!SOA && RM && (N * (int)sizeof(T)) % 16 == 0) {
                    VT* dst = reinterpret_cast<VT*>(LU + baseA + (int64_t)p * N);
#pragma unroll
                    for (int q = 0; q < N / V; ++q) {
                        T t[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) t[v] = a[k][q * V + v];
                        dst[q] = pack(t);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < N; ++c) LU[baseA + (int64_t)(RM ? N * p + c : N * c + p) * ST] = a[k][c];
                }
            }
        }
    }

    // ---- backward sweep, column-oriented (:356-365); row p now sits in lane p / R, slot p % R ------------------
    const int gbase = lane & ~(G - 1);
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        constexpr int dummy = 0;
        (void)dummy;
        const int oc_lane = c / R, oc_slot = c % R;              // static
        const bool on = c >= skip;
        const T q = div_rn(x[oc_slot], a[oc_slot][c]);
        const T xc = __shfl_sync(0xffffffffu, q, gbase + oc_lane);
        if (on && lg == oc_lane) x[oc_slot] = q;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int p = lg * R + k;
            const bool row_on = on && (p < c) && (p >= skip);
            const T v = fma_rn(-xc, a[k][c], x[k]);
            x[k] = row_on ? v : x[k];
        }
    }

    // ---- store -------------------------------------------------------------------------------------------------
    if (valid) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int p = lg * R + k;
            if (p < N) {
                const int64_t off = baseX + (int64_t)p * ST;
                X[off] = good ? x[k] : Bm[off];                  // x untouched on failure (:391-393)
            }
        }
        if (lg == 0 && ok_out) ok_out[s] = good ? 1 : 0;
    }
}

}  // namespace gmb
