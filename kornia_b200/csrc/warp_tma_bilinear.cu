// kornia_b200 -- instantiations of the TMA-tiled warp forward kernel, bilinear interpolation.
#include "warp_tma_host.cuh"

namespace kb200 {
int warp_tma_forward_bilinear(const TmaFwdArgs& a, cudaStream_t st) { return warp_tma_forward_impl<KB200_BILINEAR>(a, st); }
}  // namespace kb200
