// kornia_b200 -- host side of the fused SSIM kernel (ssim_tiled.cuh).
#include "ssim_tiled.cuh"

namespace kb200 {

template <int K>
static int launch(const SsimParams& p, long long tiles, cudaStream_t st) {
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    KB_CUDA(cudaFuncSetAttribute(ssim_tiled_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssim_smem_bytes<K>()));
  ssim_tiled_kernel<K><<<(unsigned)tiles, 256, ssim_smem_bytes<K>(), st>>>(p);
  return KB200_OK;
}

int ssim_tiled_forward(const float* a, const float* b, const float* taps, float* out, int planes, int H, int W, int K, float C1,
                       float C2, float eps, cudaStream_t st) {
  if (K % 2 == 0 || K < 3 || K > SSIM_MAX_K) return KB200_EUNSUPPORTED;
  KB_CHECK_ARG(K / 2 < H && K / 2 < W, "reflect border of %d needs an image larger than %d x %d", K / 2, H, W);
  SsimParams p;
  p.a = a; p.b = b; p.taps = taps; p.out = out;
  p.planes = planes; p.H = H; p.W = W;
  p.tiles_x = ceil_div(W, SSIM_TW);
  p.tiles_y = ceil_div(H, SSIM_TH);
  p.pair_ok = (W % 2 == 0) && ((uintptr_t)out % 8 == 0);
  p.C1 = C1; p.C2 = C2; p.eps = eps;
  const long long tiles = (long long)p.tiles_x * p.tiles_y * planes;
  KB_CHECK_ARG(tiles <= 0x7fffffffLL, "too many tiles (%lld)", tiles);
  int rc;
  switch (K) {
    case 3: rc = launch<3>(p, tiles, st); break;
    case 5: rc = launch<5>(p, tiles, st); break;
    case 7: rc = launch<7>(p, tiles, st); break;
    case 9: rc = launch<9>(p, tiles, st); break;
    default: rc = launch<11>(p, tiles, st); break;
  }
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("ssim_forward: kernel launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

}  // namespace kb200
