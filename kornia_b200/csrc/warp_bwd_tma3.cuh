// kornia_b200 -- tiled fused warp backward, third structure: straight-line pixel units on conflict-free lanes (fp32,
// bilinear, zeros/border, C = 3 or 1).
//
// Same per-pixel arithmetic and the same skeleton as warp_bwd_tma2 (every warp its own pipeline: own 64 x 4 sub-tile, own
// 72 x 8 x C accumulation strip flushed by one TMA reduce-add, own TMA-loaded window of `src` for d/dM; the coordinate chain
// of imgwarp.py:165-170 op for op).  What the ncu capture of round 2 (profiles/r2_first_bwd_ncu_digest.txt) said about the
// per-pixel loop of the earlier kernels, and what this one does about it:
//   * 254 thread-instructions per pixel of which ~95 are branches / integer glue, issue-active 43 %.  Here a UNIT of
//     UR rows x 2 columns per lane is evaluated as straight-line code -- all coordinates first (independent chains
//     interleave), one fast-path decision per pixel as a lane predicate instead of a warp vote + branch, the upstream gradient
//     of the next unit in flight while this one is processed.
//   * 50 % of the shared-memory wavefronts were bank conflicts: lanes stood two columns apart (stride-2 floats) so that the
//     lanes of one instruction never share a floor cell.  Here lane <-> column (stride 1, conflict-free); two neighbouring
//     lanes that fall into the same source cell (maps that minify, about one instruction in three on the benchmark
//     homographies) are separated by RANK: lanes of rank 0 update the strip first, rank 1 in a second round that is only
//     issued when some lane needs it (warp-uniform branch); deeper pile-ups (step < 1/2) take the exact per-pixel path.
//   * ~10 % of the pixels -- every tile on the left / top image edge whose taps reach column / row -1 -- left the fast path
//     because a TMA reduce cannot start at a negative coordinate.  Those taps are outside the image anyway ('zeros': no
//     contribution, tap value 0): here they are predicated off and the pixel stays on the fast path.
//   * the exact path is taken per LANE (divergent branch), not per warp instruction.
#pragma once
#include "warp_bwd_tma.cuh"

namespace kb200 {

template <int NC, int PAD, bool PROJ, bool ALIGN, bool NEED_SRC, bool NEED_M, int UR = 2>
__global__ void __launch_bounds__(BWD_THREADS, 2) warp_bwd_tma3(const __grid_constant__ CUtensorMap tmap_srcwin,
                                                                const __grid_constant__ CUtensorMap tmap_gsrc,
                                                                const __grid_constant__ CUtensorMap tmap_gout,
                                                                const __grid_constant__ TmaBwdParams p) {
  using R = RN<float>;
  constexpr int TW = 64, TH = 32, BW = 72;
  constexpr int NJ = TW / 32, RPW = TH / TMA_CONSUMER_WARPS;
  constexpr int U = UR * NJ;                        // pixels of a unit
  constexpr int UPT = RPW / UR;                     // units per tile and warp
  static_assert(RPW % UR == 0, "units tile the warp's rows");
  constexpr int SPLANE = BW * BWD_SH;               // one channel of a window / strip
  constexpr int STRIP_FLOATS = NC * SPLANE;
  constexpr uint32_t WIN_BYTES = STRIP_FLOATS * 4;
  static_assert((STRIP_FLOATS * 4) % 128 == 0, "per-warp buffers stay 128-byte aligned");

  extern __shared__ __align__(128) unsigned char bwd3_smem[];
  float* wins = reinterpret_cast<float*>(bwd3_smem);                                   // [8 warps][NC][SH][BW] (NEED_M only)
  float* strips = wins + (NEED_M ? TMA_CONSUMER_WARPS * STRIP_FLOATS : 0);              // [8 warps][NC][SH][BW] (NEED_SRC only)
  uint64_t* wfull = reinterpret_cast<uint64_t*>(strips + (NEED_SRC ? TMA_CONSUMER_WARPS * STRIP_FLOATS : 0));  // [8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < TMA_CONSUMER_WARPS; ++i) tma::mbar_init(&wfull[i], 1);
    tma::fence_barrier_init();
    if (NEED_M) tma::prefetch_map(&tmap_srcwin);
  }
  if (NEED_M) {  // mark every record row of this CTA unused; rows are claimed as segments are processed
    for (int i = threadIdx.x; i < p.max_segs; i += blockDim.x) p.record_batch[(size_t)blockIdx.x * p.max_segs + i] = -1;
  }
  __syncthreads();  // the only CTA-wide barrier of the kernel

  const int tiles_x = ceil_div(p.w, TW), tiles_y = ceil_div(p.h, TH);
  const Segments segs(p.B * tiles_y, tiles_x);
  const int H = p.H, W = p.W;
  const float Wm1 = (float)(W - 1), Hm1 = (float)(H - 1), Wf = (float)W, Hf = (float)H;
  const size_t oplane = (size_t)p.h * p.w;
  float* strip_mem = strips + warp * STRIP_FLOATS;
  float* win_mem = wins + warp * STRIP_FLOATS;
  const uint32_t strip_u32 = tma::smem_u32(strip_mem), win_u32 = tma::smem_u32(win_mem);
  uint64_t* my_full = &wfull[warp];
  uint32_t phase = 0;
  const float ux_scale = ALIGN ? Wm1 * 0.5f : Wf * 0.5f, uy_scale = ALIGN ? Hm1 * 0.5f : Hf * 0.5f;

  int seg_strip, tx0, tx1, cursor = 0;
  for (int seg = 0; segs.get(seg, seg_strip, tx0, tx1, cursor); ++seg) {
    const int b = seg_strip / tiles_y, ty = seg_strip - b * tiles_y;
    Mat3<float> m;
    m.load(p.m + (p.Bm == 1 ? 0 : (size_t)b * 9));
    const int y_base = ty * TH + warp * RPW;
    const float* gbase = p.gout + (size_t)b * NC * oplane;
    // per-row terms of this warp's rows, constant along the segment (same products the reference forms)
    float byr[RPW], cy0[RPW], cy1[RPW], cy2[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      byr[i] = __ldg(p.by + min(y_base + i, p.h - 1));
      cy0[i] = R::mul(m.m01, byr[i]);
      cy1[i] = R::mul(m.m11, byr[i]);
      cy2[i] = PROJ ? R::mul(m.m21, byr[i]) : 0.f;
    }
    float pm[9];  // d/dm partials of this thread over the whole segment
#pragma unroll
    for (int k = 0; k < 9; ++k) pm[k] = 0.f;

    // upstream gradient of unit `n` of this segment (n counts units across the segment's tiles): coalesced 128-byte rows
    auto load_gout = [&](int tx, int i0, float (&dst)[U][NC]) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u / NJ, j = u % NJ;
        const float* g = gbase + (size_t)min(y_base + i, p.h - 1) * p.w + min(tx * TW + lane + 32 * j, p.w - 1);
#pragma unroll
        for (int c = 0; c < NC; ++c) dst[u][c] = __ldg(g + c * oplane);
      }
    };
    float go_next[U][NC];
    load_gout(tx0, 0, go_next);

    for (int tx = tx0; tx < tx1; ++tx) {
      // ---------------------------------------------------------------- this warp's window (all lanes, replicated)
      int sox, soy;
      bool win_ok;
      {
        const int py = min(y_base + ((lane & 2) ? RPW - 1 : 0), p.h - 1);
        const int px = min(tx * TW + ((lane & 1) ? TW - 1 : 0), p.w - 1);
        float gx, gy, den;
        map_point<float, PROJ>(m, __ldg(p.bx + px), __ldg(p.by + py), gx, gy, den);
        float ix = unnorm<ALIGN>(gx, Wm1, Wf), iy = unnorm<ALIGN>(gy, Hm1, Hf);
        bool ok = fabsf(ix) < 4.0e6f && fabsf(iy) < 4.0e6f;
        if (PROJ) {  // lanes hold the four corners eight times over: the ballot is over the corners
          const unsigned neg = __ballot_sync(0xffffffffu, den < 0.f);
          ok = ok && (neg == 0u || neg == 0xffffffffu) && fabsf(den) > 1e-12f;
        }
        if (PAD == KB200_BORDER) {
          ix = clip_coord(ix, W);
          iy = clip_coord(iy, H);
        }
        float lo_x = ix, hi_x = ix, lo_y = iy;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
          hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
          lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        }
        win_ok = __all_sync(0xffffffffu, ok);
        const int x_lo = (int)floorf(lo_x), x_hi = (int)floorf(hi_x) + 1;
        const int need_w = x_hi - x_lo + 1;
        const int spare = BW - need_w - 3;
        const int ox = (x_lo - (spare > 0 ? spare / 2 : 0)) & ~3;  // TMA: 16-byte aligned start
        // origins are clamped to the image: a TMA reduce cannot take negative coordinates; taps in column / row -1 are
        // predicated off below (they are outside the image)
        sox = win_ok ? max(ox, 0) : 0x20000000;
        soy = win_ok ? max((int)floorf(lo_y), 0) : 0x20000000;
      }
      if (warp == 0 && tx + 1 < tx1 && tma::elect_one()) tma::prefetch_3d(&tmap_gout, (tx + 1) * TW, ty * TH, b * NC);
      __syncwarp();

      const int x0 = tx * TW + lane;
      float bxv[NJ], cx0[NJ], cx1[NJ], cx2[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        bxv[j] = __ldg(p.bx + min(x0 + 32 * j, p.w - 1));
        cx0[j] = R::mul(m.m00, bxv[j]);
        cx1[j] = R::mul(m.m10, bxv[j]);
        cx2[j] = PROJ ? R::mul(m.m20, bxv[j]) : 0.f;
      }
      if (NEED_M && win_ok) {
        // every lane has finished reading the previous tile's window (program order + the __syncwarp above)
        if (tma::elect_one()) {
          tma::fence_proxy_async();
          tma::mbar_arrive_expect_tx(my_full, WIN_BYTES);
          tma::load_3d(win_mem, &tmap_srcwin, my_full, sox, soy, b * NC);
        }
        __syncwarp();
      }
      if (NEED_SRC) {
        // the previous tile's strip must have been read by the TMA unit before it is cleared
        if (lane == 0) tma::bulk_wait_read0();
        __syncwarp();
        float4* z = reinterpret_cast<float4*>(strip_mem);
        for (int e = lane; e < STRIP_FLOATS / 4; e += 32) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // cell (ly, lx) of the window / strip = (Y - MAGIC - soy, X - MAGIC - sox)
      const unsigned kwin = (unsigned)(FLOOR_MAGIC_BITS + soy) * (unsigned)BW + (unsigned)(FLOOR_MAGIC_BITS + sox);
      const uint32_t strip_base = strip_u32 - 4u * kwin, win_base = win_u32 - 4u * kwin;
      // coordinates served from the window: both taps of each axis inside it; with 'zeros' a window that starts at the image
      // edge also serves coordinates in [-1, 0): the tap in column / row -1 is outside the image and is predicated off
      const bool edge_x = PAD == KB200_ZEROS && sox == 0, edge_y = PAD == KB200_ZEROS && soy == 0;
      const float s_lo_x = edge_x ? -1.f : (float)sox, s_hi_x = (float)(sox + BW - 1);
      const float s_lo_y = edge_y ? -1.f : (float)soy, s_hi_y = (float)(soy + BWD_SH - 1);
      bool waited = !(NEED_M && win_ok);
      __syncwarp();  // strip cleared by all lanes

#pragma unroll
      for (int un = 0; un < UPT; ++un) {
        const int i0 = un * UR;
        float go[U][NC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int c = 0; c < NC; ++c) go[u][c] = go_next[u][c];
        // next unit's upstream gradient: the next rows of this tile, or the first rows of the next tile
        if (un + 1 < UPT) load_gout(tx, i0 + UR, go_next);
        else if (tx + 1 < tx1) load_gout(tx + 1, 0, go_next);

        // ---- coordinates of the unit, straight-line (independent chains interleave)
        float ix[U], iy[U], gxs[U], gys[U], rdens[U];
        bool fast[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u / NJ, j = u % NJ;
          const float nx = R::add(R::add(cx0[j], cy0[i]), m.m02);
          const float ny = R::add(R::add(cx1[j], cy1[i]), m.m12);
          float gx = nx, gy = ny, rden = 1.f;
          bool den_ok = true;
          if (PROJ) {
            const float den = R::add(R::add(cx2[j], cy2[i]), m.m22);
            den_ok = fabsf(den) >= 8.67361738e-19f;  // 2^-60: below it the shared-reciprocal division is not exact
            rden = refined_rcp(den);
            gx = div_by_rcp(nx, den, rden);
            gy = div_by_rcp(ny, den, rden);
            if (!den_ok) {  // rare: exact library division
              gx = __fdiv_rn(nx, den);
              gy = __fdiv_rn(ny, den);
              rden = __fdiv_rn(1.f, den);
            }
          }
          gxs[u] = gx;
          gys[u] = gy;
          rdens[u] = rden;
          ix[u] = unnorm<ALIGN>(gx, Wm1, Wf);
          iy[u] = unnorm<ALIGN>(gy, Hm1, Hf);
          fast[u] = den_ok;
        }

        if (!waited) {  // warp-uniform: the window of this tile has landed (first unit only)
          tma::mbar_wait(my_full, phase);
          phase ^= 1;
          waited = true;
        }

        // ---- one pixel after the other: the strip updates of different pixels may touch the same cells
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u / NJ, j = u % NJ;
          const int y = y_base + i, x = x0 + 32 * j;
          const bool live = y < p.h && x < p.w;
          float cix = ix[u], ciy = iy[u];
          float px = 1.f, py = 1.f;  // d(padded coordinate)/d(coordinate): 0 where the border clamp is active
          if (PAD == KB200_BORDER) {
            if (!(cix > 0.f && cix < Wm1)) px = 0.f;
            if (!(ciy > 0.f && ciy < Hm1)) py = 0.f;
            cix = fminf(Wm1, fmaxf(cix, 0.f));
            ciy = fminf(Hm1, fmaxf(ciy, 0.f));
          }
          const float tX = __fadd_rd(cix, FLOOR_MAGIC), tY = __fadd_rd(ciy, FLOOR_MAGIC);
          const int X = __float_as_int(tX), Y = __float_as_int(tY);
          bool ok = live && fast[u] && cix >= s_lo_x && cix < s_hi_x && ciy >= s_lo_y && ciy < s_hi_y;
          // rank of this lane among neighbouring lanes that fall into the same floor cell (the map is monotone along a row,
          // so lanes sharing a cell are consecutive): 0 = first of its cell, 1 = second, 2+ = exact path
          const int Xl = __shfl_up_sync(0xffffffffu, X, 1), Yl = __shfl_up_sync(0xffffffffu, Y, 1);
          const bool same1 = lane > 0 && Xl == X && Yl == Y;
          const bool same2 = __shfl_up_sync(0xffffffffu, same1 ? 1 : 0, 1) != 0 && same1;
          ok = ok && !same2;
          const bool r1 = ok && same1;
          const bool any_r1 = NEED_SRC && __any_sync(0xffffffffu, r1);
          float gix = 0.f, giy = 0.f;
          const float x0f = R::sub(tX, FLOOR_MAGIC), y0f = R::sub(tY, FLOOR_MAGIC);
          const float wx1 = (x0f + 1.f) - cix, wx0 = cix - x0f, wy1 = (y0f + 1.f) - ciy, wy0 = ciy - y0f;
          const uint32_t cell = ((unsigned)Y * (unsigned)BW + (unsigned)X) * 4u;
          // taps in column / row -1 (only reachable on an image-edge window with 'zeros') are outside the image
          const bool west_in = !(edge_x && cix < 0.f), north_in = !(edge_y && ciy < 0.f);
          if (NEED_SRC) {
            const float w_nw = wx1 * wy1, w_ne = wx0 * wy1, w_sw = wx1 * wy0, w_se = wx0 * wy0;
            const uint32_t a = cell + strip_base;
            // tap by tap: within one round the active lanes hit distinct cells; __syncwarp orders rounds and taps
#pragma unroll
            for (int round = 0; round < 2; ++round) {
              if (round == 1 && !any_r1) break;  // warp-uniform
              const bool act = ok && (round == 0 ? !same1 : same1);
              if (act && west_in && north_in) {
#pragma unroll
                for (int c = 0; c < NC; ++c) tma::sts(a + c * SPLANE * 4, tma::lds(a + c * SPLANE * 4) + w_nw * go[u][c]);
              }
              __syncwarp();
              if (act && north_in) {
#pragma unroll
                for (int c = 0; c < NC; ++c) tma::sts(a + (c * SPLANE + 1) * 4, tma::lds(a + (c * SPLANE + 1) * 4) + w_ne * go[u][c]);
              }
              __syncwarp();
              if (act && west_in) {
#pragma unroll
                for (int c = 0; c < NC; ++c) tma::sts(a + (c * SPLANE + BW) * 4, tma::lds(a + (c * SPLANE + BW) * 4) + w_sw * go[u][c]);
              }
              __syncwarp();
              if (act) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                  tma::sts(a + (c * SPLANE + BW + 1) * 4, tma::lds(a + (c * SPLANE + BW + 1) * 4) + w_se * go[u][c]);
              }
              __syncwarp();
            }
          }
          if (NEED_M && ok) {
            const uint32_t t = cell + win_base;
            // s_tap = sum_c gout[c] * src[c, tap]; then the two bilinear derivatives
            float s_nw = 0.f, s_ne = 0.f, s_sw = 0.f, s_se = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const float v_nw = (west_in && north_in) ? tma::lds_ro(t + (c * SPLANE) * 4) : 0.f;
              const float v_ne = north_in ? tma::lds_ro(t + (c * SPLANE + 1) * 4) : 0.f;
              const float v_sw = west_in ? tma::lds_ro(t + (c * SPLANE + BW) * 4) : 0.f;
              const float v_se = tma::lds_ro(t + (c * SPLANE + BW + 1) * 4);
              s_nw = fmaf(go[u][c], v_nw, s_nw);
              s_ne = fmaf(go[u][c], v_ne, s_ne);
              s_sw = fmaf(go[u][c], v_sw, s_sw);
              s_se = fmaf(go[u][c], v_se, s_se);
            }
            gix = (s_ne - s_nw) * wy1 + (s_se - s_sw) * wy0;
            giy = (s_sw - s_nw) * wx1 + (s_se - s_ne) * wx0;
          }
          if (live && !ok) {
            // exact per-pixel path, lane by lane (bounds tests and atomics on global memory)
            const float2 g = bwd_pixel_global<NC, PAD, NEED_SRC, NEED_M>(p, b, y, x, cix, ciy);
            gix = g.x;
            giy = g.y;
          }
          __syncwarp();
          if (NEED_M && live) {
            const float dgx = gix * ux_scale * px, dgy = giy * uy_scale * py;
            const float ax = dgx * rdens[u], ay = dgy * rdens[u];
            pm[0] += ax * bxv[j]; pm[1] += ax * byr[i]; pm[2] += ax;
            pm[3] += ay * bxv[j]; pm[4] += ay * byr[i]; pm[5] += ay;
            if (PROJ) {
              const float az = -(ax * gxs[u] + ay * gys[u]);
              pm[6] += az * bxv[j]; pm[7] += az * byr[i]; pm[8] += az;
            }
          }
        }
      }
      if (!waited) {  // unreachable with UPT >= 1, kept so the barrier phase can never drift
        tma::mbar_wait(my_full, phase);
        phase ^= 1;
      }
      // flush the strip: one TMA reduce-add per warp and tile
      if (NEED_SRC) {
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && sox < 0x10000000 && soy < 0x10000000) {
          tma::reduce_add_3d(&tmap_gsrc, strip_u32, sox, soy, b * NC);
          tma::bulk_commit();
        }
      }
      __syncwarp();
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

}  // namespace kb200
