// explicit instantiation unit: one-lane backsolve_lu / backsolve_plu with 1..4 right-hand sides, double
#include "gmb_lanebs.cuh"
namespace gmb {
template <typename T> int lane_backsolve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int, cudaStream_t);
template <> int lane_backsolve<double>(int mode, int n, int m, int64_t B, const double* LU, const uint8_t* perm, const double* Bm, double* X, int skip,
                                    int layout, int rm, cudaStream_t st) {
    return lanebs_impl<double>(mode, n, m, B, LU, perm, Bm, X, skip, layout, rm, st);
}
}  // namespace gmb
