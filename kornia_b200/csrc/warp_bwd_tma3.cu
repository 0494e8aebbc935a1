// kornia_b200 -- instantiations of the straight-line-unit tiled backward kernel (warp_bwd_tma3.cuh).
#include "warp_bwd_tma3.cuh"

namespace kb200 {

template <int NC, int PAD, bool PROJ, bool ALIGN, bool NEED_SRC, bool NEED_M, int UR>
static int launch3u(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, cudaStream_t st) {
  auto kern = warp_bwd_tma3<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, UR>;
  constexpr size_t per_warp = (size_t)NC * 72 * BWD_SH * 4;
  constexpr size_t smem = (size_t)TMA_CONSUMER_WARPS * per_warp * ((NEED_M ? 1 : 0) + (NEED_SRC ? 1 : 0)) + TMA_CONSUMER_WARPS * sizeof(uint64_t) + 64;
  static unsigned long long configured = 0;  // per instantiation, one bit per device
  KB_SET_SMEM_ONCE(configured, kern, smem);
  kern<<<bwd_tma_grid(p.B, p.h), BWD_THREADS, smem, st>>>(msrcwin, mgsrc, mgout, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("warp_bwd_tma3 launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

// rows per straight-line unit: option "bwd_v3" = 1 -> 2 rows (4 pixels), 2 -> 1 row (2 pixels, fewer registers)
template <int NC, int PAD, bool PROJ, bool ALIGN, bool NEED_SRC, bool NEED_M>
static int launch3(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, cudaStream_t st) {
  if (option(OPT_BWD_V3) == 2) return launch3u<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, 1>(msrcwin, mgsrc, mgout, p, st);
  return launch3u<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, 2>(msrcwin, mgsrc, mgout, p, st);
}

int launch_warp_bwd_tma3(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, int C, int pad,
                         int projective, int align, bool need_src, bool need_m, cudaStream_t st) {
  int rc = KB200_EUNSUPPORTED;
#define KB_BWD3_CASE(NC_, PAD_, PROJ_, ALIGN_)                                                              \
  if (C == NC_ && pad == PAD_ && (projective != 0) == PROJ_ && (align != 0) == ALIGN_) {                   \
    if (need_src && need_m) rc = launch3<NC_, PAD_, PROJ_, ALIGN_, true, true>(msrcwin, mgsrc, mgout, p, st);   \
    else if (need_src) rc = launch3<NC_, PAD_, PROJ_, ALIGN_, true, false>(msrcwin, mgsrc, mgout, p, st);       \
    else rc = launch3<NC_, PAD_, PROJ_, ALIGN_, false, true>(msrcwin, mgsrc, mgout, p, st);                     \
  }
#define KB_BWD3_CASES(NC_, PAD_) \
  KB_BWD3_CASE(NC_, PAD_, true, true) KB_BWD3_CASE(NC_, PAD_, true, false) KB_BWD3_CASE(NC_, PAD_, false, true) KB_BWD3_CASE(NC_, PAD_, false, false)
  KB_BWD3_CASES(3, KB200_ZEROS)
  KB_BWD3_CASES(3, KB200_BORDER)
  KB_BWD3_CASES(1, KB200_ZEROS)
  KB_BWD3_CASES(1, KB200_BORDER)
#undef KB_BWD3_CASES
#undef KB_BWD3_CASE
  return rc;
}

}  // namespace kb200
