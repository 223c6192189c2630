// gmb_rowfast.cuh -- speculative fused PLU solve (linear_solve, LUDecomp.h:388-399) for n = 9..16, one right-hand side.
//
// The exact row kernel (gmb_rowplu.cuh) spends more than half of its instructions moving rows: the pivot row sits in
// a register slot that is only known at run time, so every exchange is a chain of selects on the half-rate ALU pipe
// (ncu, 16x16 float: 1691 of 3177 warp instructions; ALU pipe 76 % busy, FMA pipe 17 %).  This kernel removes the
// exchange altogether and keeps only what a *generic* system needs; everything else is detected and handed, system by
// system, to the exact kernel (k_row_solve_list).  Same layout: G lanes per system, row p in lane p % G, slot p / G.
//
// Fast path, per elimination step i (LUDecomp.h:203-246):
//   * rows NEVER move.  A row that has served as pivot row is finished; its registers are turned into NaN (its own
//     multiplier is replaced by NaN, the update then poisons the rest of the row), so the NaN-ignoring min / max below
//     skip it for free -- no "is this row still live" bookkeeping, no positions;
//   * pivot search (:205-213) = max |a(.,i)| over all slots (FMNMX) + a log2(G) shuffle reduction; the winner is the
//     row whose |a(.,i)| equals the maximum.  When the maximum is unique this IS the reference's pivot row, whatever
//     positions the reference's exchanges would have produced;
//   * the winner stores its row (columns >= i and the rider) into row i of the system's U store in shared memory
//     (predicated 128-bit stores, the dynamic index is a memory address); one __syncwarp; every lane reads the pivot
//     row back with broadcast 128-bit loads.  The U store ends up holding U and the forward-substituted rider BY
//     POSITION -- exactly what the reference leaves after its row exchanges;
//   * multipliers (:234): one reciprocal per pivot, three FMA-pipe operations per quotient (RcpOf, gmb_math.cuh: the
//     in-line part of IEEE division, same bits);
//   * rank-1 update (:241) and rider update (:350) as plain FMAs, no selects.
// After the last step the one unfinished row becomes row n-1; the rows come back from the U store in position order
// and the backward sweep (:356-365) is statically indexed, with full IEEE divisions.
//
// A system leaves the fast path (its index is appended to a list; nothing is written for it) when
//   * the maximum of a column is not unique or is zero (two rows finish in one step, or none: caught when a later step
//     finds no live row, or by the count of live rows at the end) -- ties need the reference's position rule, a zero
//     column its degenerate-column handling;
//   * a pivot or a dividend is outside the window in which the three-operation quotient is the IEEE quotient
//     (|pivot| in [2^-57, 2^57), |dividend| >= 2^-57; double: 2^-400 .. 2^400), which includes every zero multiplier
//     (the reference's `b != 0` test, :235, is therefore always true on the fast path);
//   * a NaN or an Inf meets a comparison.  No separate test is needed: an Inf in a pivot column fails the window test;
//     a NaN in a pivot column makes its row look finished, the row can then never serve as a pivot row, and the count
//     of live rows at the end is wrong (or a later step finds no live row); NaN / Inf in the last column or in the
//     right-hand side meet no comparison in the reference either and simply flow into x through the same operations.
// k_row_solve_list then solves the listed systems with the reference's exact rules.  Results are bit-identical to the
// exact kernel for every input; only the time differs (tests/test_gpu_parity.py forces both).
#pragma once
#include "gmb_common.cuh"
#include "gmb_io.cuh"
#include "gmb_math.cuh"
#include "gmb_rowplu.cuh"

namespace gmb {

template <typename T> struct FastWin;
template <> struct FastWin<float> {
    __device__ static __forceinline__ float lo() { return 6.938893903907228e-18f; }        // 2^-57
    __device__ static __forceinline__ float hi() { return 1.4411518807585587e17f; }        // 2^57
    __device__ static __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
};
template <> struct FastWin<double> {
    __device__ static __forceinline__ double lo() { return 3.872591914849318e-121; }       // 2^-400
    __device__ static __forceinline__ double hi() { return 2.582249878086909e120; }        // 2^400
    __device__ static __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
};

__device__ __forceinline__ float  fmax_(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double fmax_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  fmin_(float a, float b)   { return fminf(a, b); }
__device__ __forceinline__ double fmin_(double a, double b) { return fmin(a, b); }

// Two rank-1 update elements per instruction: d = fma(s, b, d) on a pair of adjacent columns with the multiplier
// broadcast (FFMA2 with a .F32 broadcast operand on sm_100a; same IEEE result per element as two scalar FMAs, half the
// issue slots -- the kernel is issue-bound, the FMA pipe a quarter busy).  double: two DFMAs.
__device__ __forceinline__ void fma_pair(float& d0, float& d1, float s, float b0, float b1) {
    asm("{\n\t.reg .b64 bb, dd, ss;\n\tmov.b64 bb, {%2, %3};\n\tmov.b64 dd, {%0, %1};\n\tmov.b64 ss, {%4, %4};\n\t"
        "fma.rn.f32x2 dd, ss, bb, dd;\n\tmov.b64 {%0, %1}, dd;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(b0), "f"(b1), "f"(s));
}
__device__ __forceinline__ void fma_pair(double& d0, double& d1, double s, double b0, double b1) {
    d0 = __fma_rn(s, b0, d0);
    d1 = __fma_rn(s, b1, d1);
}

// shared-memory elements per system: n matrix rows of NCP (= n padded to whole 16-byte chunks) and the n riders, plus
// one chunk when that makes the number of chunks odd (the groups of a warp then start in different bank quads:
// conflict-free 128-bit accesses)
template <typename T, int N>
constexpr int fast_sys_elems() {
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int CH = N * ((N + V - 1) / V) + (N + V - 1) / V;
    return (CH + (CH % 2 == 0 ? 1 : 0)) * V;
}
template <typename T, int N, int G, int THREADS>
constexpr size_t fast_smem_bytes() { return (size_t)(THREADS / G) * fast_sys_elems<T, N>() * sizeof(T); }

template <typename T, int N, bool SOA, bool RM, int G, int THREADS, int MINB, bool SKIP0 = true, int SELIMAD = 0>
__global__ void __launch_bounds__(THREADS, MINB)
k_fast_solve(const T* __restrict__ A, const T* Bm, T* X, uint8_t* __restrict__ ok_out, int skip, int64_t B,
             int32_t* __restrict__ redo_list, int* __restrict__ redo_count) {
    constexpr int RL = row_rl(N, G);
    constexpr int NC = N + 1;                                  // augmented row: n columns + the rider (index N)
    constexpr int NCX = (NC + 1) / 2 * 2;                      // ... padded to whole column pairs (fma_pair)
    constexpr int V = soa_v<T>();                              // elements per 128-bit access
    constexpr int NCH = (N + V - 1) / V;                       // 16-byte chunks per stored matrix row
    constexpr int NCP = NCH * V;
    constexpr int SYS = fast_sys_elems<T, N>();
    constexpr int ST = SOA ? soa_w<T>() : 1;
    constexpr int W = soa_w<T>();
    using VT = typename VecOf<T>::type;
    using FW = FastWin<T>;

    extern __shared__ __align__(16) unsigned char gmb_dyn_smem[];
    const int lane_g = (int)(threadIdx.x % G);
    T* usys = reinterpret_cast<T*>(gmb_dyn_smem) + (size_t)(threadIdx.x / G) * SYS;     // this system's U store: n rows ...
    T* ysys = usys + N * NCP;                                                           // ... and the n riders, by position

    int64_t s = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / G;
    const bool valid = s < B;
    if (!valid) s = B - 1;                                     // keep all lanes for the shuffles / syncwarp
    const int64_t baseA = SOA ? (s / W) * (int64_t)(N * N) * W + (s % W) : s * (int64_t)(N * N);
    const int64_t baseX = SOA ? (s / W) * (int64_t)N * W + (s % W) : s * (int64_t)N;

    // ---- load: slot k holds row p = k*G + lane_g; rows past n are born finished (NaN) -----------------------
    T a[RL][NCX];
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int p = k * G + lane_g;
        const bool live = ((k + 1) * G <= N) || (p < N);
        const int pr = live ? p : 0;
        if (NCX > NC) a[k][NC] = (T)0;                         // pad column of the last pair
        if constexpr (!SOA && RM && (N * (int)sizeof(T)) % 16 == 0) {
            ldg_row<T, N>(A + baseA + (int64_t)pr * N, a[k], (reinterpret_cast<uintptr_t>(A) & 31) == 0);
        } else {
#pragma unroll
            for (int c = 0; c < N; ++c) a[k][c] = A[baseA + (int64_t)(RM ? N * pr + c : N * c + pr) * ST];
        }
        a[k][N] = Bm[baseX + (int64_t)pr * ST];                // m = 1: same address in both layouts
        if ((k + 1) * G > N) {
#pragma unroll
            for (int c = 0; c < NC; ++c) a[k][c] = live ? a[k][c] : FW::nan();
        }
    }
    bool bad = false;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));

    // State carried from one step to the next (software pipeline: the search and the lane-local row selection for
    // column i+1 are issued inside step i, so that their latencies overlap the rest of step i's rank-1 update):
    T m, lm, mn;                   // column i: group maximum, this lane's maximum and minimum of |a(.,i)| over its slots
    bool take[RL];                 // lane-local selection chain: slot k replaces the running best
    T cand[NC];                    // this lane's candidate pivot row = its slot with the largest |a(.,i)| (columns >= i)

    // lane-local part of the pivot search (:205-213) in column `col`, then the log2(G) shuffle reduction of the maximum;
    // finished rows are NaN and ignored by fmax / fmin; `av == nv` is false for a NaN av and true for the first live one
    auto search = [&](int col) {
        T v = abs_(a[0][col]);
        mn = v;
#pragma unroll
        for (int k = 1; k < RL; ++k) {
            const T av = abs_(a[k][col]);
            const T nv = fmax_(v, av);
            take[k] = av == nv;
            mn = fmin_(mn, av);
            v = nv;
        }
        lm = v;
        m = v;
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) m = fmax_(m, __shfl_xor_sync(0xffffffffu, m, off, G));
    };
    // The lane collapses its slots to its candidate row BEFORE the group maximum is known (the selects then run under
    // the shuffles' latency); only the lane whose maximum is the group's publishes.  One predicated 128-bit store per
    // chunk: a predicated store per slot and chunk would cost the same 4 cycles of the SM-wide register-to-MIO path
    // whether its lanes are active or not (ncu: 61 % busy, top stall short scoreboard / MIO throttle).
    auto preselect = [&](int col) {
        // one-hot integer masks of the chosen slot, for the columns that select on the FMA pipe (SELIMAD != 0): an integer
        // multiply-add of the value's bits with a 0/1 mask is exact for every bit pattern (NaN, -0) and leaves the
        // half-rate ALU pipe -- the busiest one -- to the plain select chains of the other columns
        int oh[RL];
        if constexpr (SELIMAD != 0 && sizeof(T) == 4 && RL > 1) {
            bool later = false;
#pragma unroll
            for (int k = RL - 1; k >= 1; --k) {
                oh[k] = (take[k] && !later) ? 1 : 0;
                later = later || take[k];
            }
            oh[0] = later ? 0 : 1;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c < col) continue;
            if constexpr (SELIMAD != 0 && sizeof(T) == 4 && RL > 1) {
                if (c % SELIMAD == SELIMAD - 1) {
                    int acc;                                   // inline PTX: the compiler would turn x * (p ? 1 : 0) back into a select
                    asm("mul.lo.s32 %0, %1, %2;" : "=r"(acc) : "r"(__float_as_int((float)a[0][c])), "r"(oh[0]));
#pragma unroll
                    for (int k = 1; k < RL; ++k)
                        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(__float_as_int((float)a[k][c])), "r"(oh[k]));
                    cand[c] = (T)__int_as_float(acc);
                    continue;
                }
            }
            T v = a[0][c];
#pragma unroll
            for (int k = 1; k < RL; ++k) v = take[k] ? a[k][c] : v;
            cand[c] = v;
        }
    };
    auto publish = [&](int i, bool has) {                      // row i of the U store <- cand (columns >= i, whole chunks)
        if (has) {
            T* urow = usys + i * NCP;
#pragma unroll
            for (int q = 0; q < NCH; ++q) if (q >= i / V) {
                T t[V];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int c = q * V + v < N ? q * V + v : N - 1;
                    t[v] = c >= i ? cand[c] : a[0][c];         // columns < i are dead: any register
                }
                reinterpret_cast<VT*>(urow)[q] = pack(t);
            }
            ysys[i] = cand[N];
        }
    };

    search(0);
    preselect(0);

    // ---- elimination -----------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        // no live row at all (an earlier tie finished two rows at once) leaves m = NaN: caught by the window test
        bad = bad || !(m >= FW::lo() && m < FW::hi()) || (mn < FW::lo());
        const bool has = lm == m;                              // this lane holds the pivot row
        publish(i, has);

        // multipliers (:234) while the row travels: one reciprocal per pivot, three operations per quotient.  |pivot| =
        // m is known already, its sign comes from a vote; the reciprocal's refinement is odd in the divisor (y(-b) =
        // -y(b), bit for bit).  The winner's own multiplier becomes NaN: the update below then finishes the row.
        const bool neg = (__ballot_sync(0xffffffffu, has && cand[i] < (T)0) & gmask) != 0;
        RcpOf<T> rc(m);
        rc.b = neg ? -m : m;
        rc.y = neg ? -rc.y : rc.y;
        T nb[RL];
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            bool ok_unused;
            const T q = rc.fast(a[k][i], ok_unused);
            nb[k] = (abs_(a[k][i]) == m) ? FW::nan() : -q;
        }
        __syncwarp();
        T piv[NCX];
        if (NCX > NC) piv[NC] = (T)0;
        {
            const T* urow = usys + i * NCP;
#pragma unroll
            for (int q = 0; q < NCH; ++q) if (q >= (i + 1) / V) {
                T t[V];
                unpack(reinterpret_cast<const VT*>(urow)[q], t);
#pragma unroll
                for (int v = 0; v < V; ++v) if (q * V + v < N) piv[q * V + v] = t[v];
            }
            piv[N] = ysys[i];
        }
        // rank-1 update (:241) and the rider (:350), two columns at a time (for even i the first pair also rewrites the
        // dead column i).  The pair that holds column i+1 goes first: the next search starts from it.
        const int j0 = (i + 1) / 2;                             // (i even: columns i and i+1 share a chunk, piv[i] is loaded)
#pragma unroll
        for (int k = 0; k < RL; ++k) fma_pair(a[k][2 * j0], a[k][2 * j0 + 1], nb[k], piv[2 * j0], piv[2 * j0 + 1]);
        if (i + 1 < N - 1) search(i + 1);
#pragma unroll
        for (int k = 0; k < RL; ++k) {
#pragma unroll
            for (int j = j0 + 1; j < NCX / 2; ++j) fma_pair(a[k][2 * j], a[k][2 * j + 1], nb[k], piv[2 * j], piv[2 * j + 1]);
        }
        if (i + 1 < N - 1) preselect(i + 1);
    }

    // ---- the one unfinished row is row n-1 ---------------------------------------------------------------------
    {
        int live_rows = 0;
        bool has = false;
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const bool lv = a[k][N - 1] == a[k][N - 1];
            live_rows += lv ? 1 : 0;
            has = has || lv;
            if (k > 0) take[k] = lv;
        }
        preselect(N - 1);
        publish(N - 1, has);
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) live_rows += __shfl_xor_sync(0xffffffffu, live_rows, off, G);
        bad = bad || (live_rows != 1);
    }
    // any lane of the group saw a reason to leave the fast path -> the whole system is redone by the exact kernel
    const bool redo = (__ballot_sync(0xffffffffu, bad) & gmask) != 0;
    if (redo) {
        if (valid && lane_g == 0) redo_list[atomicAdd(redo_count, 1)] = (int32_t)s;
    }
    __syncwarp();

    // ---- backward sweep (:356-365), row by row from the U store, replicated in every lane of the group ------------------
    // x(r) = (y(r) - sum_{c = n-1 .. r+1} u(r,c) x(c)) / u(r,r).  Splitting the 120 FMAs of a 16 x 16 system over the
    // group's lanes costs a shuffle, a division in one lane and a predicate per column -- more instructions than doing
    // all of it in every lane, where every index is static and x ends up whole in each lane (one 128-bit store of the
    // lane's chunk).  The reciprocals of the diagonal do not depend on the sweep: computed up front, off the chain; a
    // quotient outside the fast window (rare) sends the lane through the sweep again with full IEEE divisions.
    T x[N];
    T yv[NCP];
    bool wide = true;
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
        T t[V];
        unpack(reinterpret_cast<const VT*>(ysys)[q], t);
#pragma unroll
        for (int v = 0; v < V; ++v) yv[q * V + v] = t[v];
    }
    // row r of U (columns >= r, whole chunks) and the reciprocal of its diagonal entry: fetched one row ahead of the
    // sweep (the compiler-level fences keep the loads from all being hoisted to the top -- 136 live registers)
    auto fetch = [&](int r, T (&urow)[NCP], T& rcy, bool& rcsafe) {
#pragma unroll
        for (int q = 0; q < NCH; ++q) if (q >= r / V) {
            T t[V];
            unpack(reinterpret_cast<const VT*>(usys + r * NCP)[q], t);
#pragma unroll
            for (int v = 0; v < V; ++v) urow[q * V + v] = t[v];
        }
        const RcpOf<T> rc(urow[r]);
        rcy = rc.y;
        rcsafe = rc.safe;
    };
    auto sweep = [&](const bool exact) {
        T rowbuf[2][NCP];
        T rcy[2];
        bool rcs[2];
        fetch(N - 1, rowbuf[(N - 1) & 1], rcy[(N - 1) & 1], rcs[(N - 1) & 1]);
#pragma unroll
        for (int r = N - 1; r >= 0; --r) {
            if (r > 0) fetch(r - 1, rowbuf[(r - 1) & 1], rcy[(r - 1) & 1], rcs[(r - 1) & 1]);
            asm volatile("" ::: "memory");
            const T (&urow)[NCP] = rowbuf[r & 1];
            T v = yv[r];
#pragma unroll
            for (int c = N - 1; c > r; --c) v = fma_rn(-x[c], urow[c], v);
            T q;
            if (exact) {
                q = div_rn(v, urow[r]);
            } else {
                const RcpOf<T> rc(urow[r], rcy[r & 1], rcs[r & 1]);
                bool ok;
                q = rc.fast(v, ok);
                wide = wide && ok;
            }
            x[r] = (SKIP0 || r >= skip) ? q : yv[r];           // rows < skip stay forward-substituted (:356)
        }
    };
    sweep(false);
    if (!wide) sweep(true);

    // ---- store (nothing for a system that left the fast path: the exact kernel writes x and ok for it) -------------
    if (valid && !redo) {
        if constexpr (!SOA && N % V == 0 && (N / V) % G == 0) {
            // lane j of the group stores chunks j, j + G, ...: one coalesced 128-bit store per lane and chunk
#pragma unroll
            for (int qq = 0; qq < N / V / G; ++qq) {
                T t[V];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    t[v] = x[(qq * G) * V + v];
#pragma unroll
                    for (int j = 1; j < G; ++j) t[v] = (lane_g == j) ? x[(qq * G + j) * V + v] : t[v];
                }
                reinterpret_cast<VT*>(X + baseX)[qq * G + lane_g] = pack(t);
            }
        } else {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int p = k * G + lane_g;
                T t = x[k * G < N ? k * G : 0];
#pragma unroll
                for (int j = 1; j < G; ++j) if (k * G + j < N) t = (lane_g == j) ? x[k * G + j] : t;
                if (((k + 1) * G <= N) || (p < N)) X[baseX + (int64_t)p * ST] = t;
            }
        }
        if (lane_g == 0 && ok_out) ok_out[s] = 1;              // a degenerate column never stays on the fast path
    }
}

// Launch configuration per (T, n): lanes per system, threads per CTA, CTAs per SM.  Measured (profiles/r2a_rowfast.md):
// four lanes per system for every n (two lanes need ~190 registers, eight spend twice the per-lane bookkeeping per
// system); SMALL CTAs -- 64 threads, eight per SM for float -- because the warps of one CTA run in lock-step phases
// (select burst on the ALU pipe, FMA burst, shuffle wait) and more independent CTAs per scheduler interleave better:
// 16 x 16 float 0.410 ms per 2^20 systems with 256-thread CTAs, 0.371 ms with 64-thread CTAs, same occupancy.
template <typename T, int N> struct FastCfg {
    static constexpr int WORDS = (int)sizeof(T) / 4;
    static constexpr int G = 4;
    static constexpr int threads = 64;
    static constexpr int minb = WORDS == 1 ? 8 : 4;
};

// The default memory pool gives freed blocks back to the driver at the next synchronisation unless told otherwise;
// the scratch below is allocated on every call, so the pool is asked (once per device) to keep what it has.
inline void keep_pool_memory() {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done.fetch_or(bit, std::memory_order_relaxed);
}

// Scratch for the redo list lives in stream-ordered memory: concurrent calls on different streams never share it.
template <typename T>
int rowfast_solve_impl(int n, int64_t B, const T* A, const T* Bm, T* X, uint8_t* ok, int skip, int layout, int rm,
                       cudaStream_t st) {
    if (B >= (int64_t)1 << 31) return GMB_ERR_UNSUPPORTED;     // 32-bit list entries; the exact kernel serves larger batches
    int rc = GMB_ERR_UNSUPPORTED;
    dispatch_int<9, 10, 11, 12, 13, 14, 15, 16>(n, [&](auto N) {
        if (B == 0) {
            rc = GMB_OK;
            return;
        }
        keep_pool_memory();
        int32_t* scratch = nullptr;
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)(B + 4) * sizeof(int32_t), st);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rc = GMB_ERR_ALLOC;
            return;
        }
        int* count = reinterpret_cast<int*>(scratch);
        int32_t* list = scratch + 4;
        cudaMemsetAsync(count, 0, sizeof(int), st);
        int sms = 148;
        {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        dispatch_bool(layout == GMB_LAYOUT_SOA, [&](auto SOA) {
            dispatch_bool(rm != 0, [&](auto RM) {
                using C = FastCfg<T, N.value>;
                auto kern = skip == 0 ? k_fast_solve<T, N.value, SOA.value, RM.value, C::G, C::threads, C::minb, true>
                                      : k_fast_solve<T, N.value, SOA.value, RM.value, C::G, C::threads, C::minb, false>;
                constexpr size_t smem = fast_smem_bytes<T, N.value, C::G, C::threads>();
                static std::atomic<uint64_t> raised[2];        // per instantiation, one bit per device
                raise_dyn_smem(kern, smem, raised[skip == 0]);
                const unsigned grid = (unsigned)ceil_div(B * C::G, C::threads);
                kern<<<grid, C::threads, smem, st>>>(A, Bm, X, ok, skip, B, list, count);
                g_launch_count.fetch_add(1, std::memory_order_relaxed);
                // exact re-solve of the listed systems: a small persistent grid (the list is almost always empty)
                using P = RowNoLu<T, N.value>;
                auto fix = k_row_solve_list<T, N.value, SOA.value, RM.value, P::G, P::minb, P::shfl, P::threads>;
                const int64_t want = ceil_div(B * P::G, P::threads);
                const unsigned fgrid = (unsigned)(want < 2 * sms ? want : 2 * sms);
                fix<<<fgrid, P::threads, 0, st>>>(A, Bm, X, ok, skip, B, list, count);
            });
        });
        rc = after_launch();
        cudaFreeAsync(scratch, st);
    });
    return rc;
}

}  // namespace gmb
