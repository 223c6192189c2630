// explicit instantiation unit: sub-warp fused PLU solve with 3 right-hand sides (n = 5..16), double
#include "gmb_subwarp.cuh"
namespace gmb {
template <typename T, int XC> int subwarp_plu_solve_m(int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <> int subwarp_plu_solve_m<double, 3>(int n, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok, double* LU, int skip,
                                            int layout, int rm, cudaStream_t st) {
    return subwarp_plu_solve_m_impl<double, 3>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st);
}
}  // namespace gmb
