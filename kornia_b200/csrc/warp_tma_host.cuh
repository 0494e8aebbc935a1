// kornia_b200 -- host-side launch helpers of the TMA-tiled warp forward kernel (included by the
// per-interpolation translation units warp_tma_{bilinear,nearest,bicubic}.cu so they compile in parallel).
#pragma once
#include "warp_tma.cuh"

namespace kb200 {

struct TmaCfg {
  int tw, th, bw, bh, nstage, ctas_per_sm, l2promo;
};
constexpr TmaCfg TMA_CFG_DEFAULT = {64, 32, 72, 40, 2, 2, 256};  // 256-B L2 promotion: +6 % over 128 B (measured)

template <int NC, int INTERP, int PAD, bool PROJ, bool ALIGN, int TW, int TH, int BW, int BH, int NSTAGE>
static int launch_warp_tma_cfg(const CUtensorMap& map, const TmaWarpParams& p, int ctas_per_sm, cudaStream_t st) {
  auto kern = warp_fwd_tma<NC, INTERP, PAD, PROJ, ALIGN, TW, TH, BW, BH, NSTAGE>;
  constexpr size_t smem = NSTAGE * (size_t)NC * BW * BH * 4 + 2 * NSTAGE * sizeof(uint64_t) + NSTAGE * sizeof(StageInfo);
  static unsigned long long configured = 0;  // per instantiation, one bit per device
  if (first_use_on_device(configured)) {
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const long long nstrips = (long long)p.B * ceil_div(p.h, TH);
  const long long cap = (long long)ctas_per_sm * sm_count();
  const int grid = (int)(nstrips < cap ? nstrips : cap);
  kern<<<grid, TMA_THREADS, smem, st>>>(map, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("warp_fwd_tma launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

inline bool tma_cfg_env_set() {
  const char* e = getenv("KB200_TMA_CFG");
  return e && e[0];
}

// Tuning knob for experiments (RGB / zeros only): KB200_TMA_CFG="TWxTHxBWxBHxSTAGESxCTAS[xL2PROMO]"
inline TmaCfg tma_cfg(int C, int pad) {
  TmaCfg c = TMA_CFG_DEFAULT;
  const char* e = getenv("KB200_TMA_CFG");
  if (e && C == 3 && pad == KB200_ZEROS) {
    TmaCfg t = c;
    const int n = sscanf(e, "%dx%dx%dx%dx%dx%dx%d", &t.tw, &t.th, &t.bw, &t.bh, &t.nstage, &t.ctas_per_sm, &t.l2promo);
    if (n >= 6) c = t;
  }
  return c;
}

template <int NC, int INTERP, int PAD, bool PROJ, bool ALIGN>
static int launch_warp_tma(const CUtensorMap& map, const TmaWarpParams& p, const TmaCfg& c, cudaStream_t st) {
#define KB_TMA_TRY(TW_, TH_, BW_, BH_, NS_)                                                      \
  if (c.tw == TW_ && c.th == TH_ && c.bw == BW_ && c.bh == BH_ && c.nstage == NS_)               \
    return launch_warp_tma_cfg<NC, INTERP, PAD, PROJ, ALIGN, TW_, TH_, BW_, BH_, NS_>(map, p, c.ctas_per_sm, st);
  if (NC == 3 && INTERP == KB200_BILINEAR && PAD == KB200_ZEROS && PROJ && ALIGN) {  // experiment grid, headline instantiation only
    KB_TMA_TRY(64, 32, 72, 40, 3)
    KB_TMA_TRY(128, 16, 136, 24, 2)
    KB_TMA_TRY(128, 16, 136, 24, 3)
    KB_TMA_TRY(128, 32, 136, 40, 2)
    KB_TMA_TRY(64, 16, 72, 24, 2)
    KB_TMA_TRY(64, 16, 72, 24, 4)
    KB_TMA_TRY(32, 32, 40, 40, 3)
  }
#undef KB_TMA_TRY
  return launch_warp_tma_cfg<NC, INTERP, PAD, PROJ, ALIGN, 64, 32, 72, 40, 2>(map, p, c.ctas_per_sm, st);
}

// Returns KB200_EUNSUPPORTED when the request is outside this kernel's envelope (the caller then
// uses warp_fwd_generic): non-bilinear, fill padding, C > 4, rows not 16-byte aligned.

struct TmaFwdArgs {
  const float *src, *m, *bx, *by, *fill;
  float* out;
  int B, C, H, W, h, w, Bm, projective, pad, align;
  int only_class;  // 0: all samples; CLASS_WIDE / CLASS_SQUARE: only the samples of that footprint class
};

template <int INTERP>
static int warp_tma_forward_impl(const TmaFwdArgs& a, cudaStream_t st) {
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return KB200_EUNSUPPORTED;
  TmaCfg cfg = tma_cfg(a.C, a.pad);
  if (!(INTERP == KB200_BILINEAR && a.C == 3 && a.pad == KB200_ZEROS && a.projective && a.align)) cfg = TMA_CFG_DEFAULT;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B * a.C};
  const cuuint64_t strides[2] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)cfg.bw, (cuuint32_t)cfg.bh, (cuuint32_t)a.C};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.src), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       cfg.l2promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                          : (cfg.l2promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                               : (cfg.l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B)),
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return KB200_EUNSUPPORTED;
  const char* co = getenv("KB200_TMA_COPYONLY");
  TmaWarpParams p{a.src, a.m, a.bx, a.by, a.fill, a.out, a.B, a.H, a.W, a.h, a.w, a.Bm, a.align,
                  (INTERP == KB200_BILINEAR && co && co[0] == '1') ? 1 : 0, a.only_class};
  const int C = a.C, pad = a.pad;
  const bool projective = a.projective != 0, align = a.align != 0;
#define KB_TMA_CASE(NC_, PAD_)                                                                                                   \
  if (C == NC_ && pad == PAD_)                                                                                                   \
    return projective ? (align ? launch_warp_tma<NC_, INTERP, PAD_, true, true>(map, p, cfg, st)                                  \
                               : launch_warp_tma<NC_, INTERP, PAD_, true, false>(map, p, cfg, st))                                \
                      : (align ? launch_warp_tma<NC_, INTERP, PAD_, false, true>(map, p, cfg, st)                                 \
                               : launch_warp_tma<NC_, INTERP, PAD_, false, false>(map, p, cfg, st));
  KB_TMA_CASE(3, KB200_ZEROS)
  KB_TMA_CASE(3, KB200_BORDER)
  KB_TMA_CASE(3, KB200_REFLECTION)
  KB_TMA_CASE(3, KB200_FILL)
  KB_TMA_CASE(1, KB200_ZEROS)
  KB_TMA_CASE(1, KB200_BORDER)
  KB_TMA_CASE(1, KB200_REFLECTION)
  KB_TMA_CASE(1, KB200_FILL)
  if (INTERP == KB200_BILINEAR) {
    KB_TMA_CASE(4, KB200_ZEROS)
    KB_TMA_CASE(4, KB200_BORDER)
    KB_TMA_CASE(4, KB200_REFLECTION)
    KB_TMA_CASE(2, KB200_ZEROS)
    KB_TMA_CASE(2, KB200_BORDER)
    KB_TMA_CASE(2, KB200_REFLECTION)
  }
#undef KB_TMA_CASE
  return KB200_EUNSUPPORTED;
}

int warp_tma_forward_bilinear(const TmaFwdArgs& a, cudaStream_t st);
int warp_tma_forward_nearest(const TmaFwdArgs& a, cudaStream_t st);
int warp_tma_forward_bicubic(const TmaFwdArgs& a, cudaStream_t st);
// bilinear, C in {1,3}: 32 x 32 output tiles with a 56 x 56 source box, for the samples of class CLASS_SQUARE
int warp_tma_forward_square(const TmaFwdArgs& a, cudaStream_t st);

}  // namespace kb200
