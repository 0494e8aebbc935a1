// kornia_b200 -- host side of the uint8 ingest warp (warp_u8.cuh) and its C entry point.
#include "warp_u8_tiled.cuh"

#include <stdlib.h>

namespace kb200 {

constexpr int U8_MAX_Z = 65535;

template <int INTERP, int PAD, int KIND>
static int launch_u8(const WarpU8Params& p, cudaStream_t st) {
  const dim3 block(GEN_BX, GEN_BY);
  for (int b0 = 0; b0 < p.B; b0 += U8_MAX_Z) {
    WarpU8Params q = p;
    q.B = min(U8_MAX_Z, p.B - b0);
    q.src = p.src + (size_t)b0 * p.H * p.W * p.C;
    q.out = p.out + (size_t)b0 * p.C * p.h * p.w;
    if (p.Bm != 1) q.m = p.m + (size_t)b0 * 9;
    const dim3 grid(ceil_div(p.w, GEN_BX), ceil_div(p.h, GEN_BY), q.B);
    if (p.C == 3)
      warp_fwd_u8hwc<INTERP, PAD, KIND, 3><<<grid, block, 0, st>>>(q);
    else if (p.C == 1)
      warp_fwd_u8hwc<INTERP, PAD, KIND, 1><<<grid, block, 0, st>>>(q);
    else
      warp_fwd_u8hwc<INTERP, PAD, KIND, 0><<<grid, block, 0, st>>>(q);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("warp_u8hwc_forward: kernel launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

// ---------------------------------------------------------------- tiled kernel (warp_u8_tiled.cuh)
template <int NC, int PAD, int KIND, bool ALIGN>
static int launch_u8_tiled(const WarpU8Params& p, cudaStream_t st) {
  const dim3 grid(ceil_div(p.w, 64), ceil_div(p.h, 32), p.B);
  warp_u8_tiled_kernel<NC, PAD, KIND, ALIGN><<<grid, 256, U8T_SMEM_BYTES(NC), st>>>(p);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("warp_u8hwc_forward (tiled): kernel launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

// KB200_EUNSUPPORTED: the request is served by warp_fwd_u8hwc instead.
static int u8_tiled_forward(const WarpU8Params& p, int projective, int interp, int pad, cudaStream_t st) {
  if (!option(OPT_U8_TILED)) return KB200_EUNSUPPORTED;
  if (interp != KB200_BILINEAR || (p.C != 1 && p.C != 3 && p.C != 4)) return KB200_EUNSUPPORTED;
  // aligned 32-bit loads of whole in-image groups of 4 pixels: every image row starts on a 4-byte boundary
  if (p.W % 4 != 0 || (reinterpret_cast<uintptr_t>(p.src) & 3) != 0) return KB200_EUNSUPPORTED;
  if (p.B > 65535 || ceil_div(p.h, 32) > 65535) return KB200_EUNSUPPORTED;
#define KB_U8T_CASE(NC_, PAD_)                                                                                             \
  if (p.C == NC_ && pad == PAD_) {                                                                                         \
    if (projective) return p.align ? launch_u8_tiled<NC_, PAD_, KIND_PROJ, true>(p, st) : launch_u8_tiled<NC_, PAD_, KIND_PROJ, false>(p, st);  \
    return p.align ? launch_u8_tiled<NC_, PAD_, KIND_AFFINE, true>(p, st) : launch_u8_tiled<NC_, PAD_, KIND_AFFINE, false>(p, st);     \
  }
  KB_U8T_CASE(3, KB200_ZEROS)
  KB_U8T_CASE(3, KB200_BORDER)
  KB_U8T_CASE(3, KB200_REFLECTION)
  KB_U8T_CASE(1, KB200_ZEROS)
  KB_U8T_CASE(1, KB200_BORDER)
  KB_U8T_CASE(1, KB200_REFLECTION)
  KB_U8T_CASE(3, KB200_FILL)
  KB_U8T_CASE(1, KB200_FILL)
  KB_U8T_CASE(4, KB200_ZEROS)  // RGBA / BGRA frames
  KB_U8T_CASE(4, KB200_BORDER)
  KB_U8T_CASE(4, KB200_REFLECTION)
  KB_U8T_CASE(4, KB200_FILL)
#undef KB_U8T_CASE
  return KB200_EUNSUPPORTED;
}

template <int INTERP, int KIND>
static int by_pad(const WarpU8Params& p, int pad, cudaStream_t st) {
  switch (pad) {
    case KB200_ZEROS: return launch_u8<INTERP, KB200_ZEROS, KIND>(p, st);
    case KB200_BORDER: return launch_u8<INTERP, KB200_BORDER, KIND>(p, st);
    case KB200_REFLECTION: return launch_u8<INTERP, KB200_REFLECTION, KIND>(p, st);
    case KB200_FILL: return launch_u8<INTERP, KB200_FILL, KIND>(p, st);
  }
  set_error("bad pad %d", pad);
  return KB200_EINVAL;
}

template <int KIND>
static int by_interp(const WarpU8Params& p, int interp, int pad, cudaStream_t st) {
  switch (interp) {
    case KB200_BILINEAR: return by_pad<KB200_BILINEAR, KIND>(p, pad, st);
    case KB200_NEAREST: return by_pad<KB200_NEAREST, KIND>(p, pad, st);
    case KB200_BICUBIC: return by_pad<KB200_BICUBIC, KIND>(p, pad, st);
  }
  set_error("bad interp %d", interp);
  return KB200_EINVAL;
}

}  // namespace kb200

using namespace kb200;

int kb200_warp_u8hwc_forward(const void* src, const void* m, const void* bx, const void* by, const void* fill, void* out, int B, int C,
                             int H, int W, int h, int w, int Bm, int projective, int interp, int pad, int align_corners,
                             int normalize, void* stream) {
  KB_CHECK_ARG(src && m && bx && by && out, "null pointer argument");
  KB_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0, "non-positive shape B=%d C=%d H=%d W=%d h=%d w=%d", B, C, H, W, h, w);
  KB_CHECK_ARG((long long)H * W < (1ll << 31) && (long long)h * w < (1ll << 31), "plane too large for 32-bit in-plane offsets");
  KB_CHECK_ARG(Bm == B || Bm == 1, "matrix batch %d must be %d or 1", Bm, B);
  KB_CHECK_ARG(pad != KB200_FILL || fill, "pad=fill needs a fill vector");
  KB_CHECK_ARG(normalize >= 0 && normalize <= 2, "normalize must be 0 (raw), 1 (times 1/255) or 2 (divided by 255), got %d", normalize);
  WarpU8Params p{};
  p.src = (const unsigned char*)src; p.m = (const float*)m; p.bx = (const float*)bx; p.by = (const float*)by;
  p.fill = (const float*)fill; p.out = (float*)out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.h = h; p.w = w; p.Bm = Bm; p.align = align_corners; p.normalize = normalize;
  cudaStream_t st = (cudaStream_t)stream;
  KB_CHECK_ARG(interp >= KB200_BILINEAR && interp <= KB200_BICUBIC, "bad interp %d", interp);
  KB_CHECK_ARG(pad >= KB200_ZEROS && pad <= KB200_FILL, "bad pad %d", pad);
  const int rc = u8_tiled_forward(p, projective, interp, pad, st);
  if (rc != KB200_EUNSUPPORTED) return rc;
  return projective ? by_interp<KIND_PROJ>(p, interp, pad, st) : by_interp<KIND_AFFINE>(p, interp, pad, st);
}

int kb200_undistort_u8hwc_forward(const void* src, const void* lens, void* out, int B, int C, int H, int W, int normalize, void* stream) {
  KB_CHECK_ARG(src && lens && out, "null pointer argument");
  KB_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "non-positive shape B=%d C=%d H=%d W=%d", B, C, H, W);
  KB_CHECK_ARG((long long)H * W < (1ll << 31), "plane too large for 32-bit in-plane offsets");
  KB_CHECK_ARG(normalize >= 0 && normalize <= 2, "normalize must be 0 (raw), 1 (times 1/255) or 2 (divided by 255), got %d", normalize);
  // the tiled kernel is the only form: anything else is declined and the host converts + calls the fp32 path
  if ((C != 1 && C != 3) || W % 4 != 0 || (reinterpret_cast<uintptr_t>(src) & 3) != 0 || B > 65535 || ceil_div(H, 32) > 65535) {
    set_error("undistort_u8hwc_forward: needs C in {1,3}, W %% 4 == 0 and a 4-byte aligned image");
    return KB200_EUNSUPPORTED;
  }
  WarpU8Params p{};
  p.src = (const unsigned char*)src; p.lens = (const float*)lens; p.out = (float*)out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.h = H; p.w = W; p.Bm = B; p.align = 1; p.normalize = normalize;
  cudaStream_t st = (cudaStream_t)stream;
  return C == 3 ? launch_u8_tiled<3, KB200_ZEROS, U8_KIND_LENS, true>(p, st) : launch_u8_tiled<1, KB200_ZEROS, U8_KIND_LENS, true>(p, st);
}
