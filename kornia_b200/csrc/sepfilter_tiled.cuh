// placeholder until the register-blocked separable kernel lands
#pragma once
#include "common.cuh"
namespace kb200 {
inline int sepfilter_tiled_forward(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int,
                                   int, cudaStream_t) {
  return KB200_EUNSUPPORTED;
}
}  // namespace kb200
