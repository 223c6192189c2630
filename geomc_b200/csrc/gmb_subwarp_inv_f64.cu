// explicit instantiation unit: sub-warp inverse (invNxN), double
#include "gmb_subwarp.cuh"
namespace gmb {
template <typename T> int subwarp_inv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <> int subwarp_inv<double>(int n, int64_t B, const double* A, double* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm,
                                 cudaStream_t st) {
    return subwarp_inv_impl<double>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
}
template <typename T> int subwarp_invd(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <> int subwarp_invd<double>(int n, int64_t B, const double* A, double* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm,
                                 cudaStream_t st) {
    return subwarp_invd_impl<double>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
}
}  // namespace gmb
