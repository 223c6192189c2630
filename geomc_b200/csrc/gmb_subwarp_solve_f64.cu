// explicit instantiation unit: sub-warp fused PLU solve, double
#include "gmb_subwarp.cuh"
namespace gmb {
template <typename T> int subwarp_plu_solve(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <> int subwarp_plu_solve<double>(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok, double* LU,
                                       int skip, int layout, int rm, cudaStream_t st) {
    return subwarp_plu_solve_impl<double>(n, m, B, A, Bm, X, ok, LU, skip, layout, rm, st);
}
}  // namespace gmb
