// gmb_tile4.cuh -- shared-memory resident fused PLU solve for 16 x 16 float systems (config C5), four lanes per system:
// linear_solve<T,RowMajor>(a, n, 1, x, skip), LUDecomp.h:388-399 (= decomp_plu :188-250, permutation of x :35-64,
// backsolve_lu :332-366).  One right-hand side, no LU output, row-major.
//
// Why.  The register kernels (gmb_rowplu.cuh) cannot index a register file with the run-time pivot row and spend 60 %
// of their instructions on select chains and shuffles (0.31 of the HBM peak, issue-bound).  Shared memory can be
// indexed at run time, so here the system lives in shared memory and rows are NEVER moved: a position -> physical-row
// table (16 nibbles in a register pair) is exchanged instead.  One thread per system (gmb_tile.cuh, kept as the
// measured first step) is bit-exact but leaves 6 warps per SM -- neither the HBM latency of the ingest nor the
// dependent chains of the elimination are hidden.  With FOUR lanes per system the same 208 resident systems per SM
// give 24 warps.
//
// Layout: 16-byte chunk q = 4 * row + (column / 4) of CTA-local system s at float4 index q * NSYS + s.  Lane h of a
// system owns the columns c = h (mod 4), i.e. word h of every chunk; lanes are arranged h = lane / 8, s = lane % 8 so
// that a quarter-warp is eight different systems:
//   * a scalar access to "my word of chunk (p, k)" hits bank 4 s + h for ANY run-time row p: conflict-free;
//   * a 128-bit access to any chunk per lane is conflict-free (eight consecutive 16-byte slots per quarter-warp);
//   * the AoS ingest is cp.async (16 bytes, global -> shared without registers), 64 contiguous bytes per system and
//     instruction, asynchronous and conflict-free: the shared-memory transpose the SoA container would have done on
//     the host side happens on the way in.
// The rider (right-hand side) has its own array sy[p * NSYS + s].
//
// Elimination, blocked by four columns (= one chunk) with the reference's per-element FMA order:
//   panel  : the rows below the pivot are dealt to the four lanes; a lane reads its row's panel chunk (128 bit), divides
//            (reciprocal shared per pivot, see tile_rcp), updates the <= 3 columns to the right, writes the chunk back and
//            tracks the next column's pivot candidate; a two-step shuffle butterfly combines the candidates with the
//            reference's tie rule (largest |a|, lowest position among equals);
//   trailing: every lane updates ITS columns of every row below the block with the block's four steps in one pass
//            (multipliers = one broadcast 128-bit load, the four pivot rows' own columns in registers).
// Bit-exact with the register kernels and the oracle.
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"
#include "gmb_rowplu.cuh"                                    // InfOf
#include "gmb_tile.cuh"                                      // tile_rcp / tile_div_fast / tile_div_window

namespace gmb {

template <int NT> constexpr size_t tile4_smem_bytes() { return (size_t)(NT / 4) * (16 * 16 + 16) * 4; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

template <bool SOA, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_tile4_solve(const float* __restrict__ A, const float* Bm, float* X, uint8_t* __restrict__ ok_out, int skip, int64_t B) {
    using T = float;
    constexpr int N = 16, NN = 256;
    constexpr int NSYS = NT / 4;                               // systems per CTA
    static_assert(NT % 32 == 0 && NSYS % 8 == 0, "");
    constexpr int W = soa_w<T>();
    constexpr int CS = NSYS * 4;                               // words between consecutive chunks of one system
    constexpr int RS = 4 * CS;                                 // words between consecutive rows
    extern __shared__ __align__(16) unsigned char gmb_dyn_smem[];
    T* sm = reinterpret_cast<T*>(gmb_dyn_smem);
    T* sy = sm + (size_t)NN * NSYS;                            // rider: sy[p * NSYS + s]

    const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
    const int h = lane >> 3, sw = lane & 7;
    const int sl = warp * 8 + sw;                              // CTA-local system
    const int64_t s_w0 = (int64_t)blockIdx.x * NSYS + warp * 8;   // the warp's first system
    const int64_t s = s_w0 + sw;
    const bool valid = s < B;
    const int64_t sc = valid ? s : B - 1;
    const int64_t baseX = SOA ? (sc / W) * (int64_t)N * W + (sc % W) : sc * (int64_t)N;

    T* const myq = sm + sl * 4;                                // chunk (p, k) of my system: myq + p * RS + k * CS
    T* const myw = myq + h;                                    // my word of that chunk
    T* const myy = sy + sl;                                    // rider of row p: myy[p * NSYS]

    // ---- ingest -------------------------------------------------------------------------------------------------
    if (!SOA) {
        if (valid) {
            const T* src = A + s * (int64_t)NN + 4 * h;
#pragma unroll
            for (int i = 0; i < 16; ++i) cp_async16(myq + (4 * i + h) * CS, src + 16 * i);   // chunk q = 4 i + h
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float4 bq = *reinterpret_cast<const float4*>(Bm + baseX + 4 * h);
        myy[(4 * h + 0) * NSYS] = bq.x;
        myy[(4 * h + 1) * NSYS] = bq.y;
        myy[(4 * h + 2) * NSYS] = bq.z;
        myy[(4 * h + 3) * NSYS] = bq.w;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        const int64_t baseA = (sc / W) * (int64_t)NN * W + (sc % W);
#pragma unroll 16
        for (int i = 0; i < 64; ++i) {                         // word w = 4 i + h of my system
            const int w = 4 * i + h;
            myq[(w >> 2) * CS + (w & 3)] = A[baseA + (int64_t)w * W];
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) myy[(4 * h + v) * NSYS] = Bm[baseX + (int64_t)(4 * h + v) * W];
    }
    __syncwarp();

#define GMB_Q(p, k) (*reinterpret_cast<float4*>(myq + (p) * RS + (k) * CS))
#define GMB_W(p, k) myw[(p) * RS + (k) * CS]
#define GMB_Y(p) myy[(p) * NSYS]

    // position -> physical row, one nibble each; identical in the four lanes of a system
    uint64_t perm = 0xfedcba9876543210ull;
    auto perm_at = [&](int r) -> int { return (int)((uint32_t)(perm >> (4 * r)) & 15u); };
    auto comp = [](const float4& q, int i) -> T { return i == 0 ? q.x : i == 1 ? q.y : i == 2 ? q.z : q.w; };

    // combine the four lanes' pivot candidates: largest value, lowest position among equals (:205-213)
    auto combine = [&](T& bv, int& bp) {
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            const T ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int op = __shfl_xor_sync(0xffffffffu, bp, off);
            const bool take = (ov > bv) || (ov == bv && op < bp);
            bv = take ? ov : bv;
            bp = take ? op : bp;
        }
    };

    bool bad = false;

    // pivot search of step 0 (the later ones ride on the updates); lane h looks at rows h, h + 4, h + 8, h + 12
    T best = (T)-2;
    int br = 1 << 20;
    {
        T c0[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) c0[i] = abs_(GMB_Q(4 * i + h, 0).x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            T v = c0[i];
            if (v != v) v = (4 * i + h == 0) ? InfOf<T>::get() : (T)-1;   // NaN: the diagonal keeps the pivot, others never win
            if (v > best) {
                best = v;
                br = 4 * i + h;
            }
        }
        combine(best, br);
    }

#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {                           // block of columns 4 kb .. 4 kb + 3 = chunk kb of every row
        const int j0 = 4 * kb;
        const int KB = (kb == 3) ? 3 : 4;                      // elimination steps in this block (step 15 does not exist, :203)

        // ---- panel ------------------------------------------------------------------------------------------------
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            if (jj < KB) {
                const int j = j0 + jj;
                bad = bad || (best == (T)0);                   // :215 -- linear_solve then fails, x stays b
                {
                    const int pj = perm_at(j), pb = perm_at(br);
                    const uint64_t d = (uint64_t)(pj ^ pb);
                    perm ^= (d << (4 * j)) ^ (d << (4 * br));  // br == j: no change
                }
                const int pp = perm_at(j);
                const float4 pq = GMB_Q(pp, kb);               // the pivot row's panel chunk (same address in the four lanes)
                const TileRcp rc = tile_rcp(comp(pq, jj));

                const int KR = (N - 1 - j + 3) / 4;            // rows per lane (the last one may be missing)
                int p[4];
                float4 a[4];
                bool have[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < KR) {
                        const int r = j + 1 + h + 4 * k;
                        have[k] = r < N;
                        p[k] = perm_at(have[k] ? r : j);       // a harmless row for the missing one (never stored)
                        a[k] = GMB_Q(p[k], kb);
                    }
                }
                bool fast = rc.safe;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < KR) fast = fast && (tile_div_window(comp(a[k], jj)) || !have[k]);
                T qv[4];
                if (fast) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < KR) qv[k] = tile_div_fast(comp(a[k], jj), rc);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < KR) qv[k] = div_rn(comp(a[k], jj), rc.b);
                }
                T nbest = (T)-2;
                int nbr = 1 << 20;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < KR) {
                        const int r = j + 1 + h + 4 * k;
                        const T b = qv[k];                     // :234
                        const bool upd = b != (T)0;            // :235
                        T e[4] = {a[k].x, a[k].y, a[k].z, a[k].w};
                        e[jj] = b;                             // :245
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c > jj) {
                                const T v2 = fma_rn(-b, comp(pq, c), e[c]);   // :241
                                e[c] = upd ? v2 : e[c];
                            }
                        }
                        if (jj < 3) {                          // candidate of the next step's search
                            T v = abs_(e[jj + 1]);
                            if (v != v) v = (r == j + 1) ? InfOf<T>::get() : (T)-1;
                            if (have[k] && v > nbest) {
                                nbest = v;
                                nbr = r;
                            }
                        }
                        if (have[k]) GMB_Q(p[k], kb) = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
                if (jj < 3) {
                    combine(nbest, nbr);
                    best = nbest;
                    br = nbr;
                }
                __syncwarp();
            }
        }

        // ---- the block's pivot rows on my trailing columns (chunks kb+1 .. 3) and the rider: steps in order ----------
        const int NK = 3 - kb;                                 // trailing chunks
        T u[4][4];                                             // u[a][i]: pivot row a, my column of chunk kb + 1 + i; [.][3] = rider
        {
            int pa[4];
            float4 lq[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (a < KB) {
                    pa[a] = perm_at(j0 + a);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < NK) u[a][i] = GMB_W(pa[a], kb + 1 + i);
                    u[a][3] = GMB_Y(pa[a]);
                    if (a > 0) lq[a] = GMB_Q(pa[a], kb);
                }
            }
            __syncwarp();
#pragma unroll
            for (int a = 1; a < 4; ++a) {
                if (a < KB) {
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) {
                        if (b2 < a) {
                            const T l = comp(lq[a], b2);
                            const bool upd = l != (T)0;
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                if (i < NK) {
                                    const T v2 = fma_rn(-l, u[b2][i], u[a][i]);
                                    u[a][i] = upd ? v2 : u[a][i];
                                }
                            }
                            u[a][3] = fma_rn(-u[b2][3], l, u[a][3]);   // rider: never skipped (:350)
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < NK) GMB_W(pa[a], kb + 1 + i) = u[a][i];
                    if (h == 0) GMB_Y(pa[a]) = u[a][3];
                }
            }
        }

        // ---- rows below the block: the block's steps in one pass over my columns; the search of step j0 + 4 rides along
        {
            T nbest = (T)-2;
            int nbr = 1 << 20;
            const int R0 = j0 + KB;
#pragma unroll
            for (int rb = R0; rb < N; rb += 4) {               // batches of four rows: loads, arithmetic, stores
                {
                    int p[4];
                    float4 lq[4];
                    T v[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (rb + k < N) {
                            p[k] = perm_at(rb + k);
                            lq[k] = GMB_Q(p[k], kb);
#pragma unroll
                            for (int i = 0; i < 3; ++i)
                                if (i < NK) v[k][i] = GMB_W(p[k], kb + 1 + i);
                            v[k][3] = GMB_Y(p[k]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (rb + k < N) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q < KB) {
                                    const T l = comp(lq[k], q);
                                    const bool upd = l != (T)0;
#pragma unroll
                                    for (int i = 0; i < 3; ++i) {
                                        if (i < NK) {
                                            const T v2 = fma_rn(-l, u[q][i], v[k][i]);
                                            v[k][i] = upd ? v2 : v[k][i];
                                        }
                                    }
                                    v[k][3] = fma_rn(-u[q][3], l, v[k][3]);
                                }
                            }
                            if (NK > 0) {                      // column j0 + 4 is lane 0's column of chunk kb + 1
                                T w = abs_(v[k][0]);
                                if (w != w) w = (rb + k == j0 + 4) ? InfOf<T>::get() : (T)-1;
                                if (w > nbest) {
                                    nbest = w;
                                    nbr = rb + k;
                                }
                            }
                        }
                    }
                    __syncwarp();                              // every lane has read the riders before lane 0 rewrites them
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (rb + k < N) {
#pragma unroll
                            for (int i = 0; i < 3; ++i)
                                if (i < NK) GMB_W(p[k], kb + 1 + i) = v[k][i];
                            if (h == 0) GMB_Y(p[k]) = v[k][3];
                        }
                    }
                }
            }
            if (KB == 4) {                                     // the last panel step had no column to its right
                best = __shfl_sync(0xffffffffu, nbest, sw);    // lane h = 0 of my system holds column j0 + 4
                br = __shfl_sync(0xffffffffu, nbr, sw);
            }
            __syncwarp();
        }
    }

    // ---- backward sweep, column-oriented, descending columns (:356-365); lane h keeps positions h, h+4, h+8, h+12 --
    int pr[4];
    T y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pr[i] = perm_at(4 * i + h);
        y[i] = GMB_Y(pr[i]);
    }
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        const int hc = c & 3, ic = c >> 2;                     // the owner lane and slot of position c
        const bool on = c >= skip;
        T ucol[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i <= ic) ucol[i] = myq[pr[i] * RS + (c >> 2) * CS + (c & 3)];   // u(position 4 i + h, c)
        const T qmine = div_rn(y[ic], ucol[ic]);               // meaningful in the owner lane
        const T q = __shfl_sync(0xffffffffu, qmine, hc * 8 + sw);
        if (h == hc) y[ic] = on ? q : y[ic];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i <= ic) {
                const int pos = 4 * i + h;
                const bool row_on = on && (pos < c) && (pos >= skip);
                const T v = fma_rn(-q, ucol[i], y[i]);
                y[i] = row_on ? v : y[i];
            }
        }
    }

    // ---- store: x, or b again when a pivot column was zero (x untouched, :391-393) ---------------------------------
    const bool good = !bad;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) myy[(4 * i + h) * NSYS] = y[i];   // x by position
    __syncwarp();
    if (valid) {
        if (!SOA) {
            if (good) {
                float4 o;
                o.x = myy[(4 * h + 0) * NSYS];
                o.y = myy[(4 * h + 1) * NSYS];
                o.z = myy[(4 * h + 2) * NSYS];
                o.w = myy[(4 * h + 3) * NSYS];
                *reinterpret_cast<float4*>(X + baseX + 4 * h) = o;
            } else if (X != Bm) {
                *reinterpret_cast<float4*>(X + baseX + 4 * h) = *reinterpret_cast<const float4*>(Bm + baseX + 4 * h);
            }
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int64_t off = baseX + (int64_t)(4 * h + v) * W;
                if (good) X[off] = myy[(4 * h + v) * NSYS];
                else if (X != Bm) X[off] = Bm[off];
            }
        }
        if (h == 0 && ok_out) ok_out[s] = good ? 1 : 0;
    }
#undef GMB_Q
#undef GMB_W
#undef GMB_Y
}

}  // namespace gmb
