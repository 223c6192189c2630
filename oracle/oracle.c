/* oracle.c -- CPU oracle for the batched small-matrix factor/solve hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  A plain-C restatement of the reference's algorithm
 * (trbabb/geomc, geomc/linalg/{LUDecomp.h,Cholesky.h,mtxdetail/MatrixInv.h,mtxdetail/MatrixDet.h},
 * geomc/function/functiondetail/UtilDetail.h; and the linear_solve consumers of shape/Simplex.h:
 * Simplex::contains, Simplex::intersect(Ray), trace_simplex) looping over an AoS batch, one system
 * at a time.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker.  The product (geomc_b200/) never links, loads or calls it
 * and fails loudly when its CUDA library is missing.
 *
 * Pinning: the reference ships no golden vectors.  This file is pinned bit-for-bit against the
 * reference itself (oracle/_ref/libgeomc_ref.so, compiled from the unmodified headers under
 * /root/reference by oracle/Makefile) in tests/test_oracle_vs_ref.py, and against fixtures that
 * _ref generated (tests/golden/, made by tests/golden/make_golden.py) in tests/test_oracle_golden.py.
 * One function has no reference counterpart and is therefore "parity unpinned":
 * orc_cholesky_solve_* (the reference has a Cholesky factorization but no Cholesky solve).
 *
 * Each function in oracle_impl.inc cites the reference file:line it follows.
 * Build: gcc -O2 -mfma -ffp-contract=off -shared -fPIC oracle.c -o liboracle.so -lm   (oracle/Makefile)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define ORC_MAX_N 32
#define ORC_MAX_M 32

#define T float
#define FN(x) x##_f32
#define FMA fmaf
#define SQRT sqrtf
#define FABS fabsf
#define TMAXV FLT_MAX
#include "oracle_impl.inc"
#undef T
#undef FN
#undef FMA
#undef SQRT
#undef FABS
#undef TMAXV

#define T double
#define FN(x) x##_f64
#define FMA fma
#define SQRT sqrt
#define FABS fabs
#define TMAXV DBL_MAX
#include "oracle_impl.inc"
#undef T
#undef FN
#undef FMA
#undef SQRT
#undef FABS
#undef TMAXV

const char* orc_build_info(void) {
    return "oracle.c (C port of geomc linalg hot path), gcc " __VERSION__ " -O2 -mfma -ffp-contract=off";
}
