// explicit instantiation unit: row-distributed fused PLU solve, double
#include "gmb_rowplu.cuh"
namespace gmb {
template <typename T> int rowplu_solve(int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t);
template <> int rowplu_solve<double>(int n, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok, double* LU, int skip,
                                  int layout, int rm, cudaStream_t st) {
    return rowplu_solve_impl<double>(n, B, A, Bm, X, ok, LU, skip, layout, rm, st);
}
}  // namespace gmb
