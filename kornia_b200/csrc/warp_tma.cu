// kornia_b200 -- entry of the TMA-tiled warp forward path: envelope checks, then the per-interpolation unit.
#include "warp_tma_host.cuh"

namespace kb200 {

static thread_local int g_tma_launches = 1;
int warp_tma_last_launches() { return g_tma_launches; }

constexpr long long SQUARE_MIN_PIXELS = 1ll << 20;  // below this the second launch costs more than it can save

// Returns KB200_EUNSUPPORTED when the request is outside this kernel's envelope (the caller then uses
// warp_fwd_generic): C > 4 (C > 1 and != 3 for nearest / bicubic / fill), rows not 16-byte aligned.
int warp_tma_forward(const float* src, const float* m, const float* bx, const float* by, const float* fill, float* out, int B, int C,
                     int H, int W, int h, int w, int Bm, int projective, int interp, int pad, int align, cudaStream_t st) {
  if (C < 1 || C > 4) return KB200_EUNSUPPORTED;
  if ((W % 4) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0) return KB200_EUNSUPPORTED;
  if ((long long)B * C > 0x7fffffffll || (long long)B * ((h + 31) / 32) > 0x7fffffffll) return KB200_EUNSUPPORTED;
  if (pad == KB200_FILL && !fill) return KB200_EUNSUPPORTED;
  TmaFwdArgs a{src, m, bx, by, fill, out, B, C, H, W, h, w, Bm, projective, pad, align, 0};
  g_tma_launches = 1;
  if (interp == KB200_BILINEAR && (C == 1 || C == 3) && (long long)B * h * w >= SQUARE_MIN_PIXELS && option(OPT_SQUARE_TILES)) {
    // Two footprint classes, two tile shapes (warp_tma_square.cu).  Small problems stay on one launch: they are
    // launch-latency bound and every tile that does not fit is still exact.  Option "square_tiles" = 0 restores the
    // single-kernel behaviour (tests compare the two bit for bit).
    a.only_class = CLASS_SQUARE;
    const int rc = warp_tma_forward_square(a, st);
    if (rc != KB200_OK && rc != KB200_EUNSUPPORTED) return rc;
    a.only_class = rc == KB200_OK ? CLASS_WIDE : 0;
    if (rc == KB200_OK) g_tma_launches = 2;
  }
  switch (interp) {
    case KB200_BILINEAR: return warp_tma_forward_bilinear(a, st);
    case KB200_NEAREST: return warp_tma_forward_nearest(a, st);
    case KB200_BICUBIC: return warp_tma_forward_bicubic(a, st);
  }
  return KB200_EUNSUPPORTED;
}

}  // namespace kb200
