// explicit instantiation unit: row-distributed Cholesky / fused Cholesky solve (n = 5..16), float
#include "gmb_rowchol.cuh"
namespace gmb {
template <typename T> int subwarp_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t);
template <> int subwarp_cholesky<float>(int n, int m, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok, float* L_out,
                                      int layout, int rm, bool solve, cudaStream_t st) {
    return rowchol_impl<float>(n, m, B, A, Bm, X, ok, L_out, layout, rm, solve, st);
}
}  // namespace gmb
