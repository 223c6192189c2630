// gmb_subwarp.cu -- sub-warp cooperative kernels for n = 5..16 (placeholder: falls through to the
// generic one-thread-per-system kernels until the cooperative kernels land).
#include "gmb_common.cuh"

namespace gmb {

template <typename T>
int subwarp_plu_solve(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int, cudaStream_t) {
    return GMB_ERR_UNSUPPORTED;
}
template <typename T>
int subwarp_decomp(int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, cudaStream_t) {
    return GMB_ERR_UNSUPPORTED;
}
template <typename T>
int subwarp_inv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t) {
    return GMB_ERR_UNSUPPORTED;
}
template <typename T>
int subwarp_cholesky(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool, cudaStream_t) {
    return GMB_ERR_UNSUPPORTED;
}

#define INSTANTIATE(T)                                                                                              \
    template int subwarp_plu_solve<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, int,       \
                                      cudaStream_t);                                                                \
    template int subwarp_decomp<T>(int, int64_t, T*, uint8_t*, uint8_t*, uint8_t*, int, int, cudaStream_t);         \
    template int subwarp_inv<T>(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);                 \
    template int subwarp_cholesky<T>(int, int, int64_t, const T*, const T*, T*, uint8_t*, T*, int, int, bool,       \
                                     cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace gmb
