// kornia_b200 -- instantiations of the warp-independent tiled backward kernel (warp_bwd_tma2.cuh).
#include "warp_bwd_tma2.cuh"

namespace kb200 {

template <int NC, int PAD, bool PROJ, bool ALIGN, bool NEED_SRC, bool NEED_M, bool STRIDE1, bool DYN = false>
static int launch2s(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, cudaStream_t st) {
  auto kern = warp_bwd_tma2<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, STRIDE1, DYN>;
  constexpr size_t per_warp = (size_t)NC * 72 * BWD_SH * 4;
  constexpr size_t smem = (size_t)TMA_CONSUMER_WARPS * per_warp * ((NEED_M ? 1 : 0) + (NEED_SRC ? 1 : 0)) + TMA_CONSUMER_WARPS * sizeof(uint64_t) + 64;
  static unsigned long long configured = 0;  // per instantiation, one bit per device
  KB_SET_SMEM_ONCE(configured, kern, smem);
  kern<<<bwd_tma_grid(p.B, p.h), BWD_THREADS, smem, st>>>(msrcwin, mgsrc, mgout, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("warp_bwd_tma2 launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

template <int NC, int PAD, bool PROJ, bool ALIGN, bool NEED_SRC, bool NEED_M>
static int launch2(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, bool dyn,
                   cudaStream_t st) {
  if constexpr (NC == 3)
    if (dyn) return launch2s<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, true, true>(msrcwin, mgsrc, mgout, p, st);
  if (option(OPT_BWD_STRIDE1)) return launch2s<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, true>(msrcwin, mgsrc, mgout, p, st);
  return launch2s<NC, PAD, PROJ, ALIGN, NEED_SRC, NEED_M, false>(msrcwin, mgsrc, mgout, p, st);
}

int launch_warp_bwd_tma2(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, int C, int pad,
                         int projective, int align, bool need_src, bool need_m, bool dyn, cudaStream_t st) {
  int rc = KB200_EUNSUPPORTED;
#define KB_BWD2_CASE(NC_, PAD_, PROJ_, ALIGN_)                                                              \
  if (C == NC_ && pad == PAD_ && (projective != 0) == PROJ_ && (align != 0) == ALIGN_) {                   \
    if (need_src && need_m) rc = launch2<NC_, PAD_, PROJ_, ALIGN_, true, true>(msrcwin, mgsrc, mgout, p, dyn, st);   \
    else if (need_src) rc = launch2<NC_, PAD_, PROJ_, ALIGN_, true, false>(msrcwin, mgsrc, mgout, p, dyn, st);       \
    else rc = launch2<NC_, PAD_, PROJ_, ALIGN_, false, true>(msrcwin, mgsrc, mgout, p, dyn, st);                     \
  }
#define KB_BWD2_CASES(NC_, PAD_) \
  KB_BWD2_CASE(NC_, PAD_, true, true) KB_BWD2_CASE(NC_, PAD_, true, false) KB_BWD2_CASE(NC_, PAD_, false, true) KB_BWD2_CASE(NC_, PAD_, false, false)
  KB_BWD2_CASES(3, KB200_ZEROS)
  KB_BWD2_CASES(3, KB200_BORDER)
  KB_BWD2_CASES(1, KB200_ZEROS)
  KB_BWD2_CASES(1, KB200_BORDER)
#undef KB_BWD2_CASES
#undef KB_BWD2_CASE
  return rc;
}

}  // namespace kb200
