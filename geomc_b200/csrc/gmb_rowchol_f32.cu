// explicit instantiation unit: row-distributed Cholesky / fused Cholesky solve (n = 5..16), float
#include "gmb_rowchol.cuh"
namespace gmb {
template <typename T> int subwarp_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <> int subwarp_cholesky<float>(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok, float* L_out,
                                      int layout, int rm, bool solve, cudaStream_t st) {
    return rowchol_impl<float>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
}
template <typename T> int row_backsolve(int, int, int, int64_t, const T*, const uint8_t*, const T*, T*, int, int, int, cudaStream_t);
template <> int row_backsolve<float>(int mode, int n, int m, int64_t B, const float* LU, const uint8_t* perm, const float* Bm, float* X, int skip,
                                   int layout, int rm, cudaStream_t st) {
    return rowbacksolve_impl<float>(mode, n, m, B, LU, perm, Bm, X, skip, layout, rm, st);
}
}  // namespace gmb
