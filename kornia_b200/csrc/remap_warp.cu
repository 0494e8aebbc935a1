// kornia_b200 -- host dispatch of the warp-pipelined remap forward kernel (remap_warp.cuh).
#include "remap_warp.cuh"

namespace kb200 {

// Experiment knob for the persistent opt-in kernels: resident CTAs per SM the grid is sized for (default: the kernel's
// launch bounds).  KB200_GRID_PER_SM=1|2|3 -- a smaller grid means fewer, longer-running CTAs (less L2 / TMA contention).
static inline long long grid_per_sm(long long dflt) {
  const char* e = getenv("KB200_GRID_PER_SM");
  const int v = e ? atoi(e) : 0;
  return (v >= 1 && v <= dflt) ? v : dflt;
}

template <int NC, int PAD, bool ALIGN, bool LENS>
static int launch_remap_warp(const CUtensorMap& map, const RemapTiledParams& p, cudaStream_t st) {
  auto kern = remap_warp_kernel<NC, PAD, ALIGN, LENS>;
  constexpr size_t smem = (size_t)8 * NC * 72 * REMAPW_SH * 4 + 8 * sizeof(uint64_t) + 64;
  static unsigned long long configured = 0;  // per instantiation, one bit per device
  if (first_use_on_device(configured)) {
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const long long nstrips = (long long)p.B * ceil_div(p.h, 32);
  const long long cap = grid_per_sm(LENS ? 2 : 3) * sm_count();  // default = the kernel's launch bounds
  const int grid = (int)(nstrips < cap ? nstrips : cap);
  kern<<<grid, 256, smem, st>>>(map, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("remap_warp launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

int remap_warp_forward(const float* src, const float* map_x, const float* map_y, const float* lens, float* out, int B, int C, int H, int W, int h,
                       int w, int Bmap, int normalized, int interp, int pad, int align, cudaStream_t st) {
  const char* on = getenv("KB200_REMAP_V2");  // off by default: not yet run on hardware
  if (!(on && on[0] == '1')) return KB200_EUNSUPPORTED;
  const char* off = getenv("KB200_DISABLE_TMA");  // tests use it to reach the generic kernel
  if (off && off[0] == '1') return KB200_EUNSUPPORTED;
  if (interp != KB200_BILINEAR || (C != 1 && C != 3) || pad == KB200_FILL) return KB200_EUNSUPPORTED;
  if ((W % 4) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 || (long long)B * ceil_div(h, 32) > 0x7fffffffll) return KB200_EUNSUPPORTED;
  if (lens && (pad != KB200_ZEROS || !align || h != H || w != W)) return KB200_EUNSUPPORTED;
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return KB200_EUNSUPPORTED;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {72, REMAPW_SH, (cuuint32_t)C};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return KB200_EUNSUPPORTED;
  RemapTiledParams p{src, map_x, map_y, out, B, H, W, h, w, Bmap, normalized, lens};
  if (lens) return C == 3 ? launch_remap_warp<3, KB200_ZEROS, true, true>(map, p, st) : launch_remap_warp<1, KB200_ZEROS, true, true>(map, p, st);
#define KB_REMAPW_CASE(NC_, PAD_)                                                                        \
  if (C == NC_ && pad == PAD_)                                                                           \
    return align ? launch_remap_warp<NC_, PAD_, true, false>(map, p, st) : launch_remap_warp<NC_, PAD_, false, false>(map, p, st);
  KB_REMAPW_CASE(3, KB200_ZEROS)
  KB_REMAPW_CASE(3, KB200_BORDER)
  KB_REMAPW_CASE(3, KB200_REFLECTION)
  KB_REMAPW_CASE(1, KB200_ZEROS)
  KB_REMAPW_CASE(1, KB200_BORDER)
  KB_REMAPW_CASE(1, KB200_REFLECTION)
#undef KB_REMAPW_CASE
  return KB200_EUNSUPPORTED;
}

}  // namespace kb200
