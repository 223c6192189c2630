// float instantiations of the (f)3 / (f)4 kernels (gmb_xform.cuh); one translation unit per element type so
// the two compile in parallel.
#include "gmb_xform.cuh"

namespace gmb {
GMB_XFORM_INSTANTIATE(float)
}  // namespace gmb
