// explicit instantiation unit: sub-warp column-pivot LU (decomp_lup, n = 5..16), float
#include "gmb_subwarp.cuh"
namespace gmb {
template <typename T> int subwarp_lup(int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, cudaStream_t);
template <> int subwarp_lup<float>(int n, int64_t B, float* A, uint8_t* perm, uint8_t* parity, uint8_t* degen, int layout, int rm,
                                cudaStream_t st) {
    return subwarp_lup_impl<float>(n, B, A, perm, parity, degen, layout, rm, st);
}
template <typename T> int subwarp_rect(int, int, int64_t, const T*, T*, uint8_t*, int, cudaStream_t);
template <> int subwarp_rect<float>(int N, int nb, int64_t B, const float* bases, float* null_basis, uint8_t* ok, int layout, cudaStream_t st) {
    return subwarp_rect_impl<float>(N, nb, B, bases, null_basis, ok, layout, st);
}
}  // namespace gmb
