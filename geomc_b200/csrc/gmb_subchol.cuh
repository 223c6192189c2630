// gmb_subchol.cuh -- sub-warp cooperative Cholesky (and fused Cholesky + solve) for n = 5..16.
//
// cholesky<T,RowMajor> (Cholesky.h:37-58) sweeps rows: s = A(r,c) - sum_{k<c} L(r,k) L(c,k), k
// ascending; diagonal: ok &= s > 0, L(r,r) = sqrt(s); off-diagonal: L(r,c) = s / L(c,c); the strict
// upper triangle is zeroed.  Only the lower triangle of the input is read.  The right-looking order
// used here applies, for k = 0..n-1, a(r,c) = fma(-L(r,k), L(c,k), a(r,c)) to every r >= c > k: each
// element receives exactly the reference's terms in the reference's order (k ascending), so the
// factor is bit-identical.  There is no pivoting, hence no row exchanges: the column-distributed
// layout of gmb_subwarp.cuh fits without its select chains.  A group of G lanes owns one system;
// lane l holds all n rows of LC consecutive columns of [A | b]; per step the owner lane of column k
// takes the square root and divides its column, the multipliers are broadcast with shuffles, and
// every lane updates its own columns.
//
// The fused solve (no reference counterpart; same definition as the n <= 4 kernel and oracle.c):
// forward L y = b during the factorization (y(k) = b(k) / L(k,k), then b(r) = fma(-y(k), L(r,k), b(r))),
// backward L^T x = y column by column (x(c) = y(c) / L(c,c); y(r) = fma(-x(c), L(c,r), y(r)), r < c).
#pragma once
#include "gmb_subwarp.cuh"

namespace gmb {

template <typename T, int N, int XC, bool SOA, bool RM>
__global__ void __launch_bounds__(kSubThreads, GMB_SUB_MIN_BLOCKS)
k_sub_chol(const T* __restrict__ A, const T* Bm, T* X, T* L_out, uint8_t* __restrict__ ok_out, int64_t B) {
    constexpr int WORDS = (int)sizeof(T) / 4;
    constexpr int G = sub_group(N, XC, WORDS);
    constexpr int NC = N + XC;
    constexpr int LC = sub_lc(N, XC, G);
    constexpr int ST = BatchAddr<T, SOA>::STRIDE;
    const int lane_g = (G == 1) ? 0 : (int)(threadIdx.x % G);
    const int c0 = lane_g * LC;
    int64_t s = ((int64_t)blockIdx.x * kSubThreads + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;

    const int64_t baseA = BatchAddr<T, SOA>::base(s, N * N);
    const int64_t baseX = BatchAddr<T, SOA>::base(s, N * (XC > 0 ? XC : 1));

    // ---- load the lower triangle of the local columns (and the rider) ------------------------
    T a[N][LC];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int j = 0; j < LC; ++j) {
            const int c = c0 + j;
            T v = (T)0;
            if (c < N) {
                if (r >= c) v = A[baseA + (int64_t)(RM ? N * r + c : N * c + r) * ST];
            } else if (c < NC) {
                v = Bm[baseX + (int64_t)r * ST];               // XC == 1: b(r)
            }
            a[r][j] = v;
        }

    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int owner = k / LC;      // static
        const int kk = k % LC;         // static
        // diagonal (Cholesky.h:46-48)
        const T sdiag = shfl_t<G>(a[k][kk], owner);
        ok = ok && (sdiag > (T)0);
        const T d = sqrt_rn(sdiag);
        if (G == 1 || lane_g == owner) a[k][kk] = d;
        // column k of L (Cholesky.h:50), broadcast to the group
        T lk[N];
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            T v = div_rn(a[r][kk], d);
            v = shfl_t<G>(v, owner);
            lk[r] = v;
            if (G == 1 || lane_g == owner) a[r][kk] = v;
        }
        // rank-1 update of the trailing lower triangle in this lane's columns
#pragma unroll
        for (int j = 0; j < LC; ++j) {
            if ((G - 1) * LC + j > k) {                        // static: some lane has a column right of k here
                const int c = c0 + j;
                const bool in_a = (c > k) && (c < N);
                // L(c, k) for this lane's column c = lane_g * LC + j: one of G statically indexed
                // entries of the broadcast column, picked by the lane index (G - 1 selects)
                T lkc = (T)0;
                if (j > k) lkc = lk[j];
#pragma unroll
                for (int l = 1; l < G; ++l) {
                    const int cl = l * LC + j;                 // static
                    if (cl > k && cl < N) lkc = (lane_g == l) ? lk[cl] : lkc;
                }
#pragma unroll
                for (int r = k + 1; r < N; ++r) {              // rows r >= c only (lower triangle)
                    const T v = fma_rn(-lk[r], lkc, a[r][j]);
                    a[r][j] = (in_a && r >= c) ? v : a[r][j];
                }
            }
        }
        if (XC > 0) {
            // forward substitution on the rider column (global column N: static owner / local index)
            constexpr int xo = N / LC, xj = N % LC;
            const T yk = div_rn(a[k][xj], d);
            if (G == 1 || lane_g == xo) {
                a[k][xj] = yk;
#pragma unroll
                for (int r = k + 1; r < N; ++r) a[r][xj] = fma_rn(-yk, lk[r], a[r][xj]);
            }
        }
    }

    // ---- L (strict upper triangle zeroed, Cholesky.h:52) ----------------------------------------
    if (L_out != nullptr && valid) {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int j = 0; j < LC; ++j) {
                const int c = c0 + j;
                if (c < N) L_out[baseA + (int64_t)(RM ? N * r + c : N * c + r) * ST] = (r >= c) ? a[r][j] : (T)0;
            }
    }

    // ---- backward sweep L^T x = y ---------------------------------------------------------------
    if (XC > 0) {
        constexpr int xo = N / LC, xj = N % LC;
#pragma unroll
        for (int c = N - 1; c >= 0; --c) {
            const T lcc = shfl_t<G>(a[c][c % LC], c / LC);
            const T xc = div_rn(a[c][xj], lcc);                // meaningful in the rider lane
            if (G == 1 || lane_g == xo) a[c][xj] = xc;
#pragma unroll
            for (int r = c - 1; r >= 0; --r) {
                const T lcr = shfl_t<G>(a[c][r % LC], r / LC);  // L(c, r): row c of the factor, column r's owner
                if (G == 1 || lane_g == xo) a[r][xj] = fma_rn(-xc, lcr, a[r][xj]);
            }
        }
        if (valid && (G == 1 || lane_g == xo)) {
#pragma unroll
            for (int r = 0; r < N; ++r) X[baseX + (int64_t)r * ST] = a[r][xj];
        }
    }
    if (valid && lane_g == 0 && ok_out) ok_out[s] = ok ? 1 : 0;
}

template <typename T>
int subwarp_cholesky_impl(int n, int m, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, T* L_out, int layout, int rm,
                          bool solve, cudaStream_t st) {
    if (solve && m != 1) return GMB_ERR_UNSUPPORTED;
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        dispatch_bool(solve, [&](auto SOLVE) {
            constexpr int XC = SOLVE.value ? 1 : 0;
            constexpr int G = sub_group(N.value, XC, (int)sizeof(T) / 4);
            const unsigned grid = (unsigned)ceil_div(B * G, kSubThreads);
            if (grid == 0) {
                rc = GMB_OK;
                return;
            }
            dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
                dispatch_bool(rm != 0, [&](auto RM) {
                    k_sub_chol<T, N.value, XC, SOA.value, RM.value><<<grid, kSubThreads, 0, st>>>(A, Bm, X, L_out, ok, B);
                });
            });
            rc = after_launch();
        });
    });
    return rc;
}

}  // namespace gmb
