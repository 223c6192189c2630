// explicit instantiation unit: sub-warp fused PLU solve, float
#include "gmb_subwarp.cuh"
namespace gmb {
template <typename T> int subwarp_plu_solve(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <> int subwarp_plu_solve<float>(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok, float* LU,
                                       int skip, int layout, int rm, cudaStream_t st) {
    return subwarp_plu_solve_impl<float>(n, m, B, A, Bm, X, ok, LU, skip, layout, rm, st);
}
}  // namespace gmb
