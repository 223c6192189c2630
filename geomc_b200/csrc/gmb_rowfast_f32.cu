// explicit instantiation unit: speculative row-distributed fused PLU solve, float
#include "gmb_rowfast.cuh"
namespace gmb {
template <typename T> int rowfast_solve(int, int64_t, const T*, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <> int rowfast_solve<float>(int n, int64_t B, const float* A, const float* Bm, float* X, uint8_t* ok, int skip, int layout, int rm,
                                  cudaStream_t st) {
    return rowfast_solve_impl<float>(n, B, A, Bm, X, ok, skip, layout, rm, st);
}
}  // namespace gmb
