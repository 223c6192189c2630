// explicit instantiation unit: row-distributed Cholesky / fused Cholesky solve (n = 5..16), double
#include "gmb_rowchol.cuh"
namespace gmb {
template <typename T> int subwarp_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <> int subwarp_cholesky<double>(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok, double* L_out,
                                      int layout, int rm, bool solve, cudaStream_t st) {
    return rowchol_impl<double>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
}
template <typename T> int row_backsolve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int, cudaStream_t);
template <> int row_backsolve<double>(int mode, int n, int m, int64_t B, const double* LU, const uint8_t* perm, const double* Bm, double* X, int skip,
                                   int layout, int rm, cudaStream_t st) {
    return rowbacksolve_impl<double>(mode, n, m, B, LU, perm, Bm, X, skip, layout, rm, st);
}
}  // namespace gmb
