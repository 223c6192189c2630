// explicit instantiation unit: one-thread-per-system kernels (n = 5..6), double
#include "gmb_lane.cuh"
namespace gmb {
template <typename T> int lane_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <> int lane_cholesky<double>(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok, double* L_out,
                                      int layout, int rm, bool solve, cudaStream_t st) {
    return lane_chol_impl<double, 6>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
}
}  // namespace gmb
