// kornia_b200 -- tiled remap forward, second structure: persistent CTAs, every warp its own pipeline (fp32, bilinear,
// zeros/border/reflection, C in {1,3}; LENS = fused undistort_image, see remap_tiled.cuh).
//
// Same per-pixel arithmetic as remap_tiled_kernel (bit-identical), different organisation.  There a CTA serves one
// 64 x 32 tile and is strictly sequential: map loads -> __syncthreads -> one thread reduces the bounding boxes of the
// eight warps and issues the 72x40xC box -> everybody waits -> blend; one CTA launch per tile (130k for B=64 at 1080p);
// measured 55-60 % of the 32 B/pixel roofline.  Here:
//   * a warp owns a 64 x 4 sub-tile: it reads its own map entries (coalesced), reduces their bounding box with
//     shuffles, derives its own 72 x 8 window, TMA-loads it (6.9 KB, own buffer, own mbarrier) and blends -- no
//     CTA-wide barrier, no shared round, nothing sequential across warps;
//   * the map entries of the NEXT tile are fetched before the warp waits for its window, so the two global latencies
//     of a tile (maps, window) overlap across tiles;
//   * CTAs are persistent over the Segments schedule of the forward kernel (strips of tiles, tail-balanced).
//
// Status: written after the round-1 GPU budget was spent; compiled for sm_100a, executed only on the host emulator
// (tools/hostemu), not yet on hardware.  Dispatched only when KB200_REMAP_V2=1 (remap_warp.cu).
#pragma once
#include "remap_tiled.cuh"

namespace kb200 {

constexpr int REMAPW_SH = 8;  // rows of a warp's source window

template <int NC, int PAD, bool ALIGN, bool LENS>
__global__ void __launch_bounds__(256, LENS ? 2 : 3) remap_warp_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RemapTiledParams p) {
  using R = RN<float>;
  constexpr int TW = 64, TH = 32, BW = 72, SH = REMAPW_SH;
  constexpr int NJ = 2, RPW = 4, NP = NJ * RPW;
  constexpr int PLANE = BW * SH;
  constexpr int WIN_FLOATS = NC * PLANE;
  constexpr uint32_t WIN_BYTES = WIN_FLOATS * 4;
  static_assert(WIN_BYTES % 128 == 0, "per-warp windows stay 128-byte aligned");
  constexpr bool INTERIOR = PAD == KB200_REFLECTION;
  constexpr bool PRECLAMP = PAD == KB200_BORDER;

  extern __shared__ __align__(128) unsigned char remapw_smem[];
  float* wins = reinterpret_cast<float*>(remapw_smem);                       // [8 warps][NC][SH][BW]
  uint64_t* wfull = reinterpret_cast<uint64_t*>(wins + 8 * WIN_FLOATS);      // [8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) tma::mbar_init(&wfull[i], 1);
    tma::fence_barrier_init();
    tma::prefetch_map(&tmap);
  }
  __syncthreads();  // the only CTA-wide barrier

  const int H = p.H, W = p.W;
  const float Wm1 = (float)(W - 1), Hm1 = (float)(H - 1), Wf = (float)W, Hf = (float)H;
  // conversions.py:1487-1498: factor = 2 / clamp(size - 1, eps)
  const float fx = R::div(2.f, fmaxf(Wm1, 1e-8f)), fy = R::div(2.f, fmaxf(Hm1, 1e-8f));
  const size_t oplane = (size_t)p.h * p.w, splane = (size_t)H * W;
  float* win_mem = wins + warp * WIN_FLOATS;
  const uint32_t win_u32 = tma::smem_u32(win_mem);
  uint64_t* my_full = &wfull[warp];
  uint32_t phase = 0;

  const int tiles_x = ceil_div(p.w, TW), tiles_y = ceil_div(p.h, TH);
  const Segments segs(p.B * tiles_y, tiles_x);

  // source coordinates (normalised, as the reference hands them to grid_sample) of this lane's 8 pixels of tile (b, ty, tx)
  auto fetch = [&](int b, int ty, int tx, const float (&L)[16], float (&gx)[NP], float (&gy)[NP]) {
    const float* mxp = LENS ? nullptr : p.map_x + (p.Bmap == 1 ? 0 : (size_t)b * oplane);
    const float* myp = LENS ? nullptr : p.map_y + (p.Bmap == 1 ? 0 : (size_t)b * oplane);
    const int x0 = tx * TW + lane, y_base = ty * TH + warp * RPW;
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int x = x0 + 32 * j, y = y_base + i, u = i * NJ + j;
        float a = 0.f, c = 0.f;
        if (x < p.w && y < p.h) {
          if (LENS) {
            lens_distort(L, (float)x, (float)y, a, c);
          } else {
            a = __ldg(mxp + (size_t)y * p.w + x);
            c = __ldg(myp + (size_t)y * p.w + x);
          }
          if (!p.normalized) {
            a = R::sub(R::mul(fx, a), 1.f);
            c = R::sub(R::mul(fy, c), 1.f);
          }
        }
        gx[u] = a;
        gy[u] = c;
      }
    }
  };

  int strip, tx0, tx1, cursor = 0;
  float L[16];
  float ngx[NP], ngy[NP];  // coordinates of the tile this warp works on next, fetched one tile ahead
  // first tile of this CTA
  {
    int c0 = 0, s0, a0, b0;
    if (!segs.get(0, s0, a0, b0, c0)) return;
    const int b = s0 / tiles_y;
    if (LENS) {
#pragma unroll
      for (int k = 0; k < 16; ++k) L[k] = __ldg(p.lens + (size_t)b * 16 + k);
    }
    fetch(b, s0 - b * tiles_y, a0, L, ngx, ngy);
  }
  for (int seg = 0; segs.get(seg, strip, tx0, tx1, cursor); ++seg) {
    const int b = strip / tiles_y, ty = strip - b * tiles_y;
    const int y_base = ty * TH + warp * RPW;
    const float* sp = p.src + (size_t)b * NC * splane;
    float* obase = p.out + (size_t)b * NC * oplane;
    // the segment after this one (for the look-ahead at this segment's last tile)
    int nstrip = 0, ntx0 = 0, ntx1 = 0, ncursor = cursor;
    const bool has_next_seg = segs.get(seg + 1, nstrip, ntx0, ntx1, ncursor);

    for (int tx = tx0; tx < tx1; ++tx) {
      // ---- 1. this tile's coordinates (fetched during the previous tile) -> unnormalised pixel coordinates + bounding box
      float ux[NP], uy[NP];
      float lo_x = 3.0e38f, hi_x = -3.0e38f, lo_y = 3.0e38f, hi_y = -3.0e38f;
      bool finite = true;
      const int x0 = tx * TW + lane;
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        const int x = x0 + 32 * (u % NJ), y = y_base + u / NJ;
        ux[u] = unnorm<ALIGN>(ngx[u], Wm1, Wf);
        uy[u] = unnorm<ALIGN>(ngy[u], Hm1, Hf);
        if (x < p.w && y < p.h) {
          float ix = ux[u], iy = uy[u];
          finite = finite && fabsf(ix) < 4.0e6f && fabsf(iy) < 4.0e6f;
          if (PRECLAMP) {
            ix = fminf(Wm1, fmaxf(ix, 0.f));
            iy = fminf(Hm1, fmaxf(iy, 0.f));
          }
          lo_x = fminf(lo_x, ix); hi_x = fmaxf(hi_x, ix);
          lo_y = fminf(lo_y, iy); hi_y = fmaxf(hi_y, iy);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
      }
      finite = __all_sync(0xffffffffu, finite);
      // ---- 2. this warp's window (every lane computes the same numbers)
      bool ok = finite && lo_x > -4.0e6f && hi_x < 4.0e6f && lo_y > -4.0e6f && hi_y < 4.0e6f && lo_x <= hi_x && lo_y <= hi_y;
      const int x_lo = ok ? (int)floorf(lo_x) : 0, x_hi = ok ? (int)floorf(hi_x) + 1 : 0;
      const int y_lo = ok ? (int)floorf(lo_y) : 0, y_hi = ok ? (int)floorf(hi_y) + 1 : 0;
      const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;
      const int spare = BW - need_w - 3;
      const int ox = (x_lo - (spare > 0 ? spare / 2 : 0)) & ~3;  // TMA: 16-byte aligned box start
      ok = ok && x_hi - ox + 1 <= BW && need_h <= SH;
      const int oy = y_lo - (SH - need_h) / 2;
      float wlx = 1.f, whx = 0.f, wly = 1.f, why = 0.f;  // empty window: every pixel takes the exact path
      if (ok) {
        wlx = (float)ox; whx = (float)(ox + BW - 1); wly = (float)oy; why = (float)(oy + SH - 1);
        if (INTERIOR) {
          wlx = fmaxf(wlx, 0.f); whx = fminf(whx, Wm1);
          wly = fmaxf(wly, 0.f); why = fminf(why, Hm1);
        }
      }
      __syncwarp();  // every lane has finished reading the previous tile's window
      if (ok && tma::elect_one()) {
        tma::fence_proxy_async();
        tma::mbar_arrive_expect_tx(my_full, WIN_BYTES);
        tma::load_3d(win_mem, &tmap, my_full, ox, oy, b * NC);
      }
      __syncwarp();
      // ---- 3. look ahead: the coordinates of the next tile of this CTA's schedule, in flight while the window lands
      {
        int nb = b, nty = ty, ntx = tx + 1;
        bool more = ntx < tx1;
        if (!more && has_next_seg) {
          nb = nstrip / tiles_y;
          nty = nstrip - nb * tiles_y;
          ntx = ntx0;
          more = true;
          if (LENS && nb != b) {
#pragma unroll
            for (int k = 0; k < 16; ++k) L[k] = __ldg(p.lens + (size_t)nb * 16 + k);
          }
        }
        if (more) fetch(nb, nty, ntx, L, ngx, ngy);
      }
      const unsigned kwin = (unsigned)(FLOOR_MAGIC_BITS + oy) * (unsigned)BW + (unsigned)(FLOOR_MAGIC_BITS + ox);
      const uint32_t tbase = win_u32 - 4u * kwin;
      if (ok) {  // warp-uniform
        tma::mbar_wait(my_full, phase);
        phase ^= 1;
      }
      // ---- 4. sample
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        const int x = x0 + 32 * (u % NJ), y = y_base + u / NJ;
        if (x >= p.w || y >= p.h) continue;
        float ix = ux[u], iy = uy[u];
        if (PRECLAMP) {
          ix = fminf(Wm1, fmaxf(ix, 0.f));
          iy = fminf(Hm1, fmaxf(iy, 0.f));
        }
        float* o = obase + (size_t)y * p.w + x;
        if (ix >= wlx && ix < whx && iy >= wly && iy < why) {
          const float tX = __fadd_rd(ix, FLOOR_MAGIC), tY = __fadd_rd(iy, FLOOR_MAGIC);
          const uint32_t a0 = ((unsigned)__float_as_int(tY) * (unsigned)BW + (unsigned)__float_as_int(tX)) * 4u + tbase;
          const float x0f = R::sub(tX, FLOOR_MAGIC), y0f = R::sub(tY, FLOOR_MAGIC);
          const float wx1 = R::sub(R::add(x0f, 1.f), ix), wx0 = R::sub(ix, x0f);
          const float wy1 = R::sub(R::add(y0f, 1.f), iy), wy0 = R::sub(iy, y0f);
          const float w_nw = R::mul(wx1, wy1), w_ne = R::mul(wx0, wy1), w_sw = R::mul(wx1, wy0), w_se = R::mul(wx0, wy0);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            float a = R::fma(tma::lds(a0 + (c * PLANE) * 4), w_nw, 0.f);
            a = R::fma(tma::lds(a0 + (c * PLANE + 1) * 4), w_ne, a);
            a = R::fma(tma::lds(a0 + (c * PLANE + BW) * 4), w_sw, a);
            a = R::fma(tma::lds(a0 + (c * PLANE + BW + 1) * 4), w_se, a);
            __stcs(o + c * oplane, a);
          }
        } else {
          PixelSampler<float, KB200_BILINEAR, PAD> S;
          S.prepare(ux[u], uy[u], H, W, ALIGN);
#pragma unroll
          for (int c = 0; c < NC; ++c) __stcs(o + c * oplane, S.sample(sp + c * splane));
        }
      }
    }
  }
}

}  // namespace kb200
