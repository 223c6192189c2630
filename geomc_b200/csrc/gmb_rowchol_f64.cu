// explicit instantiation unit: row-distributed Cholesky / fused Cholesky solve (n = 5..16), double
#include "gmb_rowchol.cuh"
namespace gmb {
template <typename T> int subwarp_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <> int subwarp_cholesky<double>(int n, int m, int64_t B, const double* A, const double* Bm, double* X, uint8_t* ok, double* L_out,
                                      int layout, int rm, bool solve, cudaStream_t st) {
    return rowchol_impl<double>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
}
}  // namespace gmb
