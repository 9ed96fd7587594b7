// placeholder until the fused kernels land
#include "jwc_internal.cuh"
namespace jwc {
int fast_modwt_forward(jwc_ctx*, const DeviceSlot&, cudaStream_t, const double*, double*, int64_t, int64_t, int,
                       const FilterPair&, int) { return JWC_ERR_UNSUPPORTED; }
int fast_modwt_inverse(jwc_ctx*, const DeviceSlot&, cudaStream_t, const double*, double*, int64_t, int64_t, int,
                       const FilterPair&, int) { return JWC_ERR_UNSUPPORTED; }
}
