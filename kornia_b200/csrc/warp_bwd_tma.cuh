// kornia_b200 -- tiled fused warp backward (fp32, bilinear, zeros/border, C = 3 or 1).
//
// Replaces grid_sampler_2d_backward (atomic scatter into a zero-filled tensor + a dense grad_grid)
// and the autograd of the ~15 broadcast elementwise ops of kornia/geometry/transform/imgwarp.py:
// 165-170 for the common case.  Same persistent strip/tile walk as the
// forward kernel (warp_tma.cuh).  Per 64x32 output tile:
//   * the producer warp maps the tile corners and the corners of every warp's 64x4 sub-tile, TMA-loads
//     the 72x40xC source box (only when d/dM is wanted) and publishes per-warp strip origins;
//   * d/dsrc: every consumer warp owns a private 72x8xC accumulation strip in shared memory.  The
//     four taps of a pixel are added with plain LDS/FADD/STS: within one warp instruction the lanes
//     are consecutive output columns whose floor cells are distinct (checked: a duplicate or an
//     out-of-strip tap sends that pixel to the exact global-atomic path), and taps of different
//     instructions are ordered by __syncwarp.  The strip is then added to global memory by ONE
//     cp.reduce.async.bulk.tensor (TMA reduce-add, SASS UTMAREDG.3D.ADD) per warp and tile instead
//     of twelve scattered atomics per pixel; out-of-image cells are clipped by the TMA unit.
//     (Measured on B200: a TMA reduce with a negative box coordinate traps, so strip origins are
//     clamped to >= 0; taps left/above the image are out of bounds anyway.)
//   * d/dM: taps come from the staged source box; the nine per-pixel partials accumulate in registers
//     over a whole strip segment, then warp-shuffle -> one record per (CTA, segment, warp); a
//     second kernel sums the records in a fixed order (deterministic, double accumulation).
#pragma once
#include "warp_tma.cuh"

namespace kb200 {

struct TmaBwdParams {
  const float* gout;   // (B,NC,h,w)
  const float* src;    // (B,NC,H,W)
  const float* m;      // (Bm,3,3)
  const float* bx;
  const float* by;
  float* gsrc;         // (B,NC,H,W) zero-filled by the caller, or null
  float* records;      // (grid, max_segs, 8 warps, 9) partial sums for d/dm, or null
  int* record_batch;   // (grid, max_segs) batch index of each record row, -1 = unused
  int B, H, W, h, w, Bm, max_segs;
  int debug;  // measurement aid (KB200_BWD_DEBUG): 1 skip the strip flush, 2 also skip the strip adds, 4 skip the d/dM taps
};

constexpr int BWD_SH = 8;         // rows of a warp's accumulation strip
constexpr int BWD_THREADS = 256;  // 8 warps; warp 0 doubles as the TMA issuer (a 9th warp would cap registers at 96)

struct BwdStageInfo {
  float lo_x, hi_x, lo_y, hi_y;  // source-box window (as in the forward kernel)
  unsigned k;                    // source-box index base
  int sox;                       // strip x origin (>= 0, multiple of 4)
  int soy[TMA_CONSUMER_WARPS];   // strip y origin per consumer warp (>= 0)
};

#ifdef KB200_HOST_EMU
static long long emu_exact_path_pixels = 0;
#endif

// One pixel handled entirely through global memory: exact scatter with bounds tests and, when
// requested, the coordinate gradient from global taps.  Returns (gix, giy).
template <int NC, int PAD, bool NEED_SRC, bool NEED_M>
__device__ __noinline__ float2 bwd_pixel_global(const TmaBwdParams& p, int b, int y, int x, float ix, float iy) {
#ifdef KB200_HOST_EMU
  ++emu_exact_path_pixels;  // tools/hostemu reports how many pixels left the shared-memory fast path
#endif
  const int H = p.H, W = p.W;
  const size_t splane = (size_t)H * W, oplane = (size_t)p.h * p.w;
  const float* gop = p.gout + (size_t)b * NC * oplane + (size_t)y * p.w + x;
  ix = guard_index(ix);  // NaN / inf / beyond int range -> out of bounds, as in the generic kernel
  iy = guard_index(iy);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = (x0f + 1.f) - ix, wx0 = ix - x0f, wy1 = (y0f + 1.f) - iy, wy0 = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const bool ok_nw = in_bounds(y0, x0, H, W), ok_ne = in_bounds(y0, x0 + 1, H, W);
  const bool ok_sw = in_bounds(y0 + 1, x0, H, W), ok_se = in_bounds(y0 + 1, x0 + 1, H, W);
  const int o = y0 * W + x0;
  float gix = 0.f, giy = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float go = __ldg(gop + c * oplane);
    if (NEED_SRC) {
      float* gs = p.gsrc + ((size_t)b * NC + c) * splane;
      if (ok_nw) atomicAdd(gs + o, wx1 * wy1 * go);
      if (ok_ne) atomicAdd(gs + o + 1, wx0 * wy1 * go);
      if (ok_sw) atomicAdd(gs + o + W, wx1 * wy0 * go);
      if (ok_se) atomicAdd(gs + o + W + 1, wx0 * wy0 * go);
    }
    if (NEED_M) {
      const float* s = p.src + ((size_t)b * NC + c) * splane;
      const float v_nw = ok_nw ? __ldg(s + o) : 0.f, v_ne = ok_ne ? __ldg(s + o + 1) : 0.f;
      const float v_sw = ok_sw ? __ldg(s + o + W) : 0.f, v_se = ok_se ? __ldg(s + o + W + 1) : 0.f;
      gix += go * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
      giy += go * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
    }
  }
  return make_float2(gix, giy);
}

template <int NC, int PAD, bool PROJ, bool ALIGN, bool NEED_SRC, bool NEED_M>
__global__ void __launch_bounds__(BWD_THREADS, 2) warp_bwd_tma(const __grid_constant__ CUtensorMap tmap_src,
                                                               const __grid_constant__ CUtensorMap tmap_gsrc,
                                                               const __grid_constant__ CUtensorMap tmap_gout,
                                                               const __grid_constant__ TmaBwdParams p) {
  using R = RN<float>;
  constexpr int TW = 64, TH = 32, BW = 72, BH = 40;
  constexpr int NJ = TW / 32, RPW = TH / TMA_CONSUMER_WARPS;
  constexpr int PLANE = BW * BH;
  constexpr uint32_t BOX_BYTES = NC * PLANE * 4;
  constexpr int SPLANE = BW * BWD_SH;               // one channel of a strip
  constexpr int STRIP_FLOATS = NC * SPLANE;

  extern __shared__ __align__(128) unsigned char bwd_smem[];
  float* box = reinterpret_cast<float*>(bwd_smem);                               // [NC][BH][BW] (NEED_M only)
  float* strips = box + (NEED_M ? NC * PLANE : 0);                                // [8 warps][NC][SH][BW]
  uint64_t* full = reinterpret_cast<uint64_t*>(strips + TMA_CONSUMER_WARPS * STRIP_FLOATS);
  uint64_t* empty = full + 1;
  BwdStageInfo* info = reinterpret_cast<BwdStageInfo*>(empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma::mbar_init(full, 1);
    tma::mbar_init(empty, TMA_CONSUMER_WARPS);
    tma::fence_barrier_init();
  }
  if (NEED_M) {  // mark every record row of this CTA unused; rows are claimed as segments are processed
    for (int i = threadIdx.x; i < p.max_segs; i += blockDim.x) p.record_batch[(size_t)blockIdx.x * p.max_segs + i] = -1;
  }
  __syncthreads();

  const int tiles_x = ceil_div(p.w, TW), tiles_y = ceil_div(p.h, TH);
  const Segments segs(p.B * tiles_y, tiles_x);
  const int H = p.H, W = p.W;
  const float Wm1 = (float)(W - 1), Hm1 = (float)(H - 1), Wf = (float)W, Hf = (float)H;

  // ------------------------------------------------------------------ tile look-ahead (warp 0 only)
  // Warp 0 prepares tile t+1 while everybody works on tile t: it maps the corners of the tile and of every
  // warp's 64 x RPW sub-tile (lane = 4 * w + corner), pulls the upstream-gradient tile and the source box into
  // L2, and -- once all warps have released the shared buffers -- publishes the origins and issues the TMA load.
  struct Next {
    int seg, strip, tx0, tx1, cursor, tx;
    bool live;
    int b, ty, ox, oy, soy_w;
    bool ok, fits;
  } nx{0, 0, 0, 0, 0, 0, false, 0, 0, 0, 0, 0, false, false};
  auto next_advance = [&]() {  // step to the following tile of this CTA's sequence
    if (nx.live && nx.tx + 1 < nx.tx1) {
      ++nx.tx;
      return;
    }
    nx.live = segs.get(nx.seg, nx.strip, nx.tx0, nx.tx1, nx.cursor);
    ++nx.seg;
    nx.tx = nx.tx0;
  };
  auto next_prepare = [&]() {  // corners + L2 prefetch of the tile `nx` points at (all 32 lanes of warp 0)
    if (!nx.live) return;
    const int b = nx.strip / tiles_y, ty = nx.strip - b * tiles_y;
    nx.b = b;
    nx.ty = ty;
    Mat3<float> m;
    m.load(p.m + (p.Bm == 1 ? 0 : (size_t)b * 9));
    const int wsub = lane >> 2;
    const int py = min(ty * TH + wsub * RPW + ((lane & 2) ? RPW - 1 : 0), p.h - 1);
    const int px = min(nx.tx * TW + ((lane & 1) ? TW - 1 : 0), p.w - 1);
    float gx, gy, den;
    map_point<float, PROJ>(m, __ldg(p.bx + px), __ldg(p.by + py), gx, gy, den);
    float ix = unnorm<ALIGN>(gx, Wm1, Wf), iy = unnorm<ALIGN>(gy, Hm1, Hf);
    bool ok = fabsf(ix) < 4.0e6f && fabsf(iy) < 4.0e6f;
    if (PROJ) {
      const unsigned neg = __ballot_sync(0xffffffffu, den < 0.f);
      ok = ok && (neg == 0u || neg == 0xffffffffu) && fabsf(den) > 1e-12f;
    }
    if (PAD == KB200_BORDER) {
      ix = clip_coord(ix, W);
      iy = clip_coord(iy, H);
    }
    float lo_x = ix, hi_x = ix, lo_y = iy, hi_y = iy;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
      hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    const float warp_lo_y = lo_y;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
      hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
      hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
    }
    ok = __all_sync(0xffffffffu, ok);
    const int x_lo = (int)floorf(lo_x), x_hi = (int)floorf(hi_x) + 1;
    const int y_lo = (int)floorf(lo_y), y_hi = (int)floorf(hi_y) + 1;
    const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;
    const int spare = BW - need_w - 3;
    nx.ox = (x_lo - (spare > 0 ? spare / 2 : 0)) & ~3;  // TMA: 16-byte aligned box start
    nx.oy = y_lo - (BH - need_h) / 2;
    nx.ok = ok;
    nx.fits = ok && x_hi - nx.ox + 1 <= BW && need_h <= BH;
    // strip origin of sub-tile wsub: clamped to the image (a TMA reduce cannot take negative coordinates)
    nx.soy_w = ok ? max((int)floorf(warp_lo_y), 0) : 0x20000000;
    if (tma::elect_one()) {
      tma::prefetch_3d(&tmap_gout, nx.tx * TW, ty * TH, b * NC);
      if (NEED_M && nx.fits) tma::prefetch_3d(&tmap_src, nx.ox, nx.oy, b * NC);
    }
    __syncwarp();
  };
  auto next_publish = [&]() {  // shared buffers are free: origins + the real load (warp 0, converged)
    if (!nx.live) return;
    if ((lane & 3) == 0) info->soy[lane >> 2] = nx.soy_w;
    __syncwarp();  // the elected lane's arrive (release) must cover the other lanes' writes to `info`
    if (tma::elect_one()) {
      info->sox = nx.ok ? max(nx.ox, 0) : 0x20000000;
      if (NEED_M) {
        if (nx.fits) {
          info->lo_x = (float)nx.ox;
          info->hi_x = (float)(nx.ox + BW - 1);
          info->lo_y = (float)nx.oy;
          info->hi_y = (float)(nx.oy + BH - 1);
          info->k = (unsigned)(FLOOR_MAGIC_BITS + nx.oy) * (unsigned)BW + (unsigned)(FLOOR_MAGIC_BITS + nx.ox);
          tma::mbar_arrive_expect_tx(full, BOX_BYTES);
          tma::load_3d(box, &tmap_src, full, nx.ox, nx.oy, nx.b * NC);
        } else {
          info->lo_x = info->lo_y = 1.f;
          info->hi_x = info->hi_y = 0.f;
          info->k = 0;
          tma::mbar_arrive(full);
        }
      } else {
        tma::mbar_arrive(full);
      }
    }
    __syncwarp();
  };
  if (warp == 0) {  // first tile: nothing to wait for
    if (NEED_M && tma::elect_one()) tma::prefetch_map(&tmap_src);
    next_advance();
    next_prepare();
    next_publish();
  }

  // -------------------------------------------------------------------- consumer warps
  const size_t oplane = (size_t)p.h * p.w;
  float* strip_mem = strips + warp * STRIP_FLOATS;
  const uint32_t strip_u32 = tma::smem_u32(strip_mem);
  uint32_t phase = 0;
  int seg_strip, tx0, tx1, cursor = 0;
  for (int seg = 0; segs.get(seg, seg_strip, tx0, tx1, cursor); ++seg) {
    const int b = seg_strip / tiles_y, ty = seg_strip - b * tiles_y;
    Mat3<float> m;
    m.load(p.m + (p.Bm == 1 ? 0 : (size_t)b * 9));
    const int y_base = ty * TH + warp * RPW;
    const float* gbase = p.gout + (size_t)b * NC * oplane;
    float pm[9];  // d/dm partials of this thread over the whole segment
#pragma unroll
    for (int k = 0; k < 9; ++k) pm[k] = 0.f;
    const float ux_scale = ALIGN ? Wm1 * 0.5f : Wf * 0.5f, uy_scale = ALIGN ? Hm1 * 0.5f : Hf * 0.5f;

    for (int tx = tx0; tx < tx1; ++tx) {
      if (warp == 0) {  // look one tile ahead
        next_advance();
        next_prepare();
      }
      // lane <-> output columns (2 lane, 2 lane + 1): inside one instruction the lanes are two pixels apart, so
      // their floor cells are distinct whenever the source step per output pixel exceeds 1/2
      const int x0 = tx * TW + 2 * lane;
      float bxv[NJ], cx0[NJ], cx1[NJ], cx2[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        bxv[j] = __ldg(p.bx + min(x0 + j, p.w - 1));
        cx0[j] = R::mul(m.m00, bxv[j]);
        cx1[j] = R::mul(m.m10, bxv[j]);
        cx2[j] = PROJ ? R::mul(m.m20, bxv[j]) : 0.f;
      }
      if (NEED_SRC) {
        // the previous tile's strip must have been read by the TMA unit before it is cleared
        if (lane == 0) tma::bulk_wait_read0();
        __syncwarp();
        float4* z = reinterpret_cast<float4*>(strip_mem);
        for (int e = lane; e < STRIP_FLOATS / 4; e += 32) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      tma::mbar_wait(full, phase);
      const float lo_x = info->lo_x, hi_x = info->hi_x, lo_y = info->lo_y, hi_y = info->hi_y;
      const unsigned kbox = info->k;
      const int sox = info->sox, soy = info->soy[warp];
      const uint32_t box_base = tma::smem_u32(box) - 4u * kbox;
      // strip cell (ly, lx) = (Y - MAGIC - soy, X - MAGIC - sox)
      const unsigned kstrip = (unsigned)(FLOOR_MAGIC_BITS + soy) * (unsigned)BW + (unsigned)(FLOOR_MAGIC_BITS + sox);
      const uint32_t strip_base = strip_u32 - 4u * kstrip;
      const float s_lo_x = (float)sox, s_hi_x = (float)(sox + BW - 1), s_lo_y = (float)soy, s_hi_y = (float)(soy + BWD_SH - 1);
      __syncwarp();

      // upstream gradient, software-pipelined one row ahead: both columns of a lane in one 8-byte load per channel
      float2 go_next[NC];
      {
        const float* g0 = gbase + (size_t)min(y_base, p.h - 1) * p.w + min(x0, p.w - 2);
#pragma unroll
        for (int c = 0; c < NC; ++c) go_next[c] = __ldg(reinterpret_cast<const float2*>(g0 + c * oplane));
      }
#pragma unroll 1
      for (int i = 0; i < RPW; ++i) {
        const int y = y_base + i;
        const float byr = __ldg(p.by + min(y, p.h - 1));
        const float cy0 = R::mul(m.m01, byr), cy1 = R::mul(m.m11, byr), cy2 = PROJ ? R::mul(m.m21, byr) : 0.f;
        float2 go2[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) go2[c] = go_next[c];
        if (i + 1 < RPW) {
          const float* g1 = gbase + (size_t)min(y + 1, p.h - 1) * p.w + min(x0, p.w - 2);
#pragma unroll
          for (int c = 0; c < NC; ++c) go_next[c] = __ldg(reinterpret_cast<const float2*>(g1 + c * oplane));
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int x = x0 + j;
          const bool live = y < p.h && x < p.w;  // warp-uniform in y, not in x
          const float nx = R::add(R::add(cx0[j], cy0), m.m02);
          const float ny = R::add(R::add(cx1[j], cy1), m.m12);
          float gx = nx, gy = ny, rden = 1.f;
          bool den_ok = true;
          if (PROJ) {
            const float den = R::add(R::add(cx2[j], cy2), m.m22);
            den_ok = fabsf(den) >= 8.67361738e-19f;  // 2^-60: below it the shared-reciprocal division is not exact
            rden = refined_rcp(den);
            gx = div_by_rcp(nx, den, rden);
            gy = div_by_rcp(ny, den, rden);
            if (!den_ok) {  // rare: exact library division (uniform cost is nil, the branch is almost never taken)
              gx = __fdiv_rn(nx, den);
              gy = __fdiv_rn(ny, den);
              rden = __fdiv_rn(1.f, den);
            }
          }
          float ix = unnorm<ALIGN>(gx, Wm1, Wf), iy = unnorm<ALIGN>(gy, Hm1, Hf);
          float px = 1.f, py = 1.f;  // d(padded coordinate)/d(coordinate): 0 where the border clamp is active
          if (PAD == KB200_BORDER) {
            if (!(ix > 0.f && ix < Wm1)) px = 0.f;
            if (!(iy > 0.f && iy < Hm1)) py = 0.f;
            ix = fminf(Wm1, fmaxf(ix, 0.f));
            iy = fminf(Hm1, fmaxf(iy, 0.f));
          }
          const float tX = __fadd_rd(ix, FLOOR_MAGIC), tY = __fadd_rd(iy, FLOOR_MAGIC);
          const int X = __float_as_int(tX), Y = __float_as_int(tY);
          // fast path conditions (all lanes of the warp must agree):
          //  - taps inside this warp's strip (x0+1, y0+1 included) and, for d/dM, inside the source box
          //  - no two lanes of this instruction share a floor cell: along a row the map is monotone, so
          //    duplicates are adjacent lanes
          bool ok = live && den_ok && ix >= s_lo_x && ix < s_hi_x && iy >= s_lo_y && iy < s_hi_y;
          if (NEED_M) ok = ok && ix >= lo_x && ix < hi_x && iy >= lo_y && iy < hi_y;
          const int Xl = __shfl_up_sync(0xffffffffu, X, 1), Yl = __shfl_up_sync(0xffffffffu, Y, 1);
          if (lane > 0 && Xl == X && Yl == Y) ok = false;
          float gix = 0.f, giy = 0.f;
          if (__all_sync(0xffffffffu, ok)) {
            const float x0f = R::sub(tX, FLOOR_MAGIC), y0f = R::sub(tY, FLOOR_MAGIC);
            const float wx1 = (x0f + 1.f) - ix, wx0 = ix - x0f, wy1 = (y0f + 1.f) - iy, wy0 = iy - y0f;
            const float w_nw = wx1 * wy1, w_ne = wx0 * wy1, w_sw = wx1 * wy0, w_se = wx0 * wy0;
            float go[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) go[c] = j == 0 ? go2[c].x : go2[c].y;
            if (NEED_SRC && !(p.debug & 2)) {
              const uint32_t a = ((unsigned)Y * (unsigned)BW + (unsigned)X) * 4u + strip_base;
              // tap by tap: lanes hit distinct cells inside one instruction; __syncwarp orders the taps
#pragma unroll
              for (int c = 0; c < NC; ++c) tma::sts(a + c * SPLANE * 4, tma::lds(a + c * SPLANE * 4) + w_nw * go[c]);
              __syncwarp();
#pragma unroll
              for (int c = 0; c < NC; ++c) tma::sts(a + (c * SPLANE + 1) * 4, tma::lds(a + (c * SPLANE + 1) * 4) + w_ne * go[c]);
              __syncwarp();
#pragma unroll
              for (int c = 0; c < NC; ++c) tma::sts(a + (c * SPLANE + BW) * 4, tma::lds(a + (c * SPLANE + BW) * 4) + w_sw * go[c]);
              __syncwarp();
#pragma unroll
              for (int c = 0; c < NC; ++c)
                tma::sts(a + (c * SPLANE + BW + 1) * 4, tma::lds(a + (c * SPLANE + BW + 1) * 4) + w_se * go[c]);
              __syncwarp();
            }
            if (NEED_M && !(p.debug & 4)) {
              const uint32_t t = ((unsigned)Y * (unsigned)BW + (unsigned)X) * 4u + box_base;
              // s_tap = sum_c gout[c] * src[c, tap]; then the two bilinear derivatives
              float s_nw = 0.f, s_ne = 0.f, s_sw = 0.f, s_se = 0.f;
#pragma unroll
              for (int c = 0; c < NC; ++c) {
                s_nw = fmaf(go[c], tma::lds(t + (c * PLANE) * 4), s_nw);
                s_ne = fmaf(go[c], tma::lds(t + (c * PLANE + 1) * 4), s_ne);
                s_sw = fmaf(go[c], tma::lds(t + (c * PLANE + BW) * 4), s_sw);
                s_se = fmaf(go[c], tma::lds(t + (c * PLANE + BW + 1) * 4), s_se);
              }
              gix = (s_ne - s_nw) * wy1 + (s_se - s_sw) * wy0;
              giy = (s_sw - s_nw) * wx1 + (s_se - s_ne) * wx0;
            }
          } else if (live) {
            // exact per-pixel path (the unpadded coordinate is re-clamped inside for 'border')
            const float2 g = bwd_pixel_global<NC, PAD, NEED_SRC, NEED_M>(p, b, y, x, ix, iy);
            gix = g.x;
            giy = g.y;
          }
          if (NEED_M && live) {
            const float dgx = gix * ux_scale * px, dgy = giy * uy_scale * py;
            const float ax = dgx * rden, ay = dgy * rden;
            pm[0] += ax * bxv[j]; pm[1] += ax * byr; pm[2] += ax;
            pm[3] += ay * bxv[j]; pm[4] += ay * byr; pm[5] += ay;
            if (PROJ) {
              const float az = -(ax * gx + ay * gy);
              pm[6] += az * bxv[j]; pm[7] += az * byr; pm[8] += az;
            }
          }
        }
      }
      // release the shared buffers, flush the strip
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(empty);
      if (NEED_SRC) {
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && sox < 0x10000000 && soy < 0x10000000 && !(p.debug & 3)) {
          tma::reduce_add_3d(&tmap_gsrc, strip_u32, sox, soy, b * NC);
          tma::bulk_commit();
        }
      }
      if (warp == 0) {  // when every warp has released this tile: publish + load the next one
        tma::mbar_wait(empty, phase);
        next_publish();
      }
      phase ^= 1;
    }
    if (NEED_M) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float v = pm[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        pm[k] = v;
      }
      if (lane == 0) {
        const size_t row = (size_t)blockIdx.x * p.max_segs + seg;
        float* rec = p.records + (row * TMA_CONSUMER_WARPS + warp) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) rec[k] = pm[k];
        if (warp == 0) p.record_batch[row] = b;
      }
    }
  }
  if (NEED_SRC && lane == 0) tma::bulk_wait0();  // reductions done before exit
}

// Second stage of d/dm for the tiled kernel: fixed-order sum over the record rows of each sample.
// grid = (9, Bm).  records (rows, 8, 9), record_batch (rows).
static __global__ void __launch_bounds__(256) warp_gm_reduce_records(const float* __restrict__ records, const int* __restrict__ record_batch,
                                                              float* __restrict__ gm, int rows, int Bm) {
  const int k = blockIdx.x, bm = blockIdx.y;
  double s = 0.0;
  for (int r = threadIdx.x; r < rows; r += 256) {
    const int rb = record_batch[r];
    if (rb < 0 || (Bm != 1 && rb != bm)) continue;
    const float* rec = records + (size_t)r * TMA_CONSUMER_WARPS * 9 + k;
#pragma unroll
    for (int w = 0; w < TMA_CONSUMER_WARPS; ++w) s += (double)rec[w * 9];
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gm[(size_t)bm * 9 + k] = (float)sh[0];
}

// ------------------------------------------------------------------------------------------ host
inline int bwd_tma_grid(int B, int h) {
  const long long nstrips = (long long)B * ceil_div(h, 32);
  const long long cap = 2ll * sm_count();
  return (int)(nstrips < cap ? nstrips : cap);
}
inline int bwd_tma_max_segs(int B, int h) {
  const long long nstrips = (long long)B * ceil_div(h, 32);
  return (int)(nstrips / bwd_tma_grid(B, h)) + 2;
}
inline size_t bwd_tma_workspace_bytes(int B, int h) {
  const size_t rows = (size_t)bwd_tma_grid(B, h) * bwd_tma_max_segs(B, h);
  return rows * (TMA_CONSUMER_WARPS * 9 * sizeof(float)) + rows * sizeof(int) + 256;
}

// Warp-independent variant (warp_bwd_tma2.cuh), opt-in with KB200_BWD_V2=1.  msrcwin: tensor map of `src` with the
// per-warp window box (72, BWD_SH, C).
int launch_warp_bwd_tma2(const CUtensorMap& msrcwin, const CUtensorMap& mgsrc, const CUtensorMap& mgout, const TmaBwdParams& p, int C, int pad,
                         int projective, int align, bool need_src, bool need_m, cudaStream_t st);

int warp_tma_backward(const float* gout, const float* src, const float* m, const float* bx, const float* by, float* gsrc, float* gm,
                      void* workspace, int B, int C, int H, int W, int h, int w, int Bm, int projective, int interp, int pad, int align,
                      cudaStream_t st);

}  // namespace kb200
