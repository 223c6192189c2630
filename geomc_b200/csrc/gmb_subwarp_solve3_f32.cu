// explicit instantiation unit: sub-warp fused PLU solve with 3 right-hand sides (n = 5..16), float
#include "gmb_subwarp.cuh"
namespace gmb {
template <typename T, int XC> int subwarp_plu_solve_m(int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <> int subwarp_plu_solve_m<float, 3>(int n, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok, float* LU, int skip,
                                            int layout, int rm, cudaStream_t st) {
    return subwarp_plu_solve_m_impl<float, 3>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st);
}
}  // namespace gmb
