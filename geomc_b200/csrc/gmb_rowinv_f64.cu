// explicit instantiation unit: row-distributed matrix inverse, double
#include "gmb_rowinv.cuh"
namespace gmb {
template <typename T> int rowinv(int, int64_t, const T*, T*, uint8_t*, int, int, int, cudaStream_t);
template <> int rowinv<double>(int n, int64_t B, const double* A, double* Ainv, uint8_t* ok, int layout, int src_rm, int dst_rm,
                           cudaStream_t st) {
    return rowinv_impl<double>(n, B, A, Ainv, ok, layout, src_rm, dst_rm, st);
}
}  // namespace gmb
