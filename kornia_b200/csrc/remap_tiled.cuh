// kornia_b200 -- tiled remap forward (fp32, bilinear, zeros/border/reflection, C in {1,3}).
//
// Replaces stack + normalize_pixel_coordinates + expand + F.grid_sample of
// kornia/geometry/transform/imgwarp.py:688-702.  One CTA per 64 x 32 output tile:
//   1. every thread reads the map entries of its 8 pixels (coalesced), applies the reference's
//      normalise -> unnormalise chain (conversions.py:1487-1498, GridSampler.h:27-35) and keeps the
//      source coordinates in registers;
//   2. the CTA reduces their bounding box (warp shuffles + one shared round), one thread aligns it and
//      issues a TMA load of the 72 x 40 x C source box (zero fill outside the image);
//   3. pixels whose taps lie inside the box are blended from shared memory (12 LDS + 12 FMA per RGB
//      pixel), the rest take the exact per-pixel global path (PixelSampler) -- results are bit-identical
//      to the generic kernel.
// Smooth maps (undistortion, flow fields, elastic grids of moderate amplitude) fit the box; a tile whose
// coordinates spread further simply runs on the exact path.  Algorithmic bytes: 8 (maps) + 12 + 12 per
// RGB pixel.
//
// LENS (undistort_image, kornia/geometry/calibration/undistort.py:183-198): the map entry of a pixel is not read but
// evaluated in registers from the 16 lens numbers of its sample -- distort_points (calibration/distort.py:137-189) op
// for op, one IEEE rounding per torch op, applied to the exact integer pixel grid create_meshgrid produces -- so the
// (B,H,W) maps and the ~45 elementwise passes that build them never exist: 24 B per RGB pixel instead of 32 + ~400.
// Tilt coefficients (a 3x3 matmul in the reference) are not covered.  Opt-in: KB200_FUSED_UNDISTORT=1 on the host.
#pragma once
#include "warp_tma.cuh"

namespace kb200 {

struct RemapTiledParams {
  const float* src;
  const float* map_x;
  const float* map_y;
  float* out;
  int B, H, W, h, w, Bmap, normalized;
  const float* lens;  // LENS only: (B,16) = fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4
};

// distort_points (calibration/distort.py:137-189 without the tilt branch) at pixel (px, py): the operations, their
// association and the python-scalar products (2 * p1 is formed first, then * x, then * y) of the reference, each
// rounded on its own as eager torch does.
__device__ __forceinline__ void lens_distort(const float (&L)[16], float px, float py, float& mx, float& my) {
  using R = RN<float>;
  const float fx = L[0], fy = L[1], cx = L[2], cy = L[3];
  const float x = R::div(R::sub(px, cx), fx), y = R::div(R::sub(py, cy), fy);
  const float r2 = R::add(R::mul(x, x), R::mul(y, y));
  const float r4 = R::mul(r2, r2);
  const float r6 = R::mul(r4, r2);
  const float num = R::add(R::add(R::add(1.f, R::mul(L[4], r2)), R::mul(L[5], r4)), R::mul(L[8], r6));
  const float den = R::add(R::add(R::add(1.f, R::mul(L[9], r2)), R::mul(L[10], r4)), R::mul(L[11], r6));
  const float rad = R::div(num, den);
  const float xy1 = R::mul(R::mul(R::mul(2.f, L[6]), x), y);                 // 2 * p1 * x * y
  const float xy2 = R::mul(R::mul(R::mul(2.f, L[7]), x), y);                 // 2 * p2 * x * y
  const float rx = R::add(r2, R::mul(R::mul(2.f, x), x));                    // r2 + 2 * x * x
  const float ry = R::add(r2, R::mul(R::mul(2.f, y), y));                    // r2 + 2 * y * y
  const float xd = R::add(R::add(R::add(R::add(R::mul(x, rad), xy1), R::mul(L[7], rx)), R::mul(L[12], r2)), R::mul(L[13], r4));
  const float yd = R::add(R::add(R::add(R::add(R::mul(y, rad), R::mul(L[6], ry)), xy2), R::mul(L[14], r2)), R::mul(L[15], r4));
  mx = R::add(R::mul(fx, xd), cx);
  my = R::add(R::mul(fy, yd), cy);
}

template <int NC, int PAD, bool ALIGN, bool LENS = false>
__global__ void __launch_bounds__(256, LENS ? 3 : 4) remap_tiled_kernel(const __grid_constant__ CUtensorMap tmap,
                                                             const __grid_constant__ RemapTiledParams p) {
  using R = RN<float>;
  constexpr int TW = 64, TH = 32, BW = 72, BH = 40;
  constexpr int NJ = 2, RPW = 4;
  constexpr int PLANE = BW * BH;
  constexpr uint32_t BOX_BYTES = NC * PLANE * 4;
  constexpr bool INTERIOR = PAD == KB200_REFLECTION;
  constexpr bool PRECLAMP = PAD == KB200_BORDER;

  extern __shared__ __align__(128) unsigned char remap_smem[];
  float* box = reinterpret_cast<float*>(remap_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(remap_smem + BOX_BYTES);
  float* red = reinterpret_cast<float*>(full + 1);  // [8 warps][4]
  float* win = red + 32;                            // lo_x, hi_x, lo_y, hi_y
  unsigned* kidx = reinterpret_cast<unsigned*>(win + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = blockIdx.x, ty = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x == 0) {
    tma::mbar_init(full, 1);
    tma::fence_barrier_init();
  }
  const int H = p.H, W = p.W;
  const float Wm1 = (float)(W - 1), Hm1 = (float)(H - 1), Wf = (float)W, Hf = (float)H;
  // conversions.py:1487-1498: factor = 2 / clamp(size - 1, eps)
  const float fx = R::div(2.f, fmaxf(Wm1, 1e-8f)), fy = R::div(2.f, fmaxf(Hm1, 1e-8f));
  const size_t oplane = (size_t)p.h * p.w;
  const float* mxp = p.map_x + (p.Bmap == 1 ? 0 : (size_t)b * oplane);
  const float* myp = p.map_y + (p.Bmap == 1 ? 0 : (size_t)b * oplane);
  const int x0 = tx * TW + lane, y_base = ty * TH + warp * RPW;
  float L[16];
  if (LENS) {
#pragma unroll
    for (int k = 0; k < 16; ++k) L[k] = __ldg(p.lens + (size_t)b * 16 + k);
  }

  // ---- 1. coordinates of this thread's pixels
  float ux[RPW * NJ], uy[RPW * NJ];  // unnormalised, un-padded (what the exact path consumes)
  float lo_x = 3.0e38f, hi_x = -3.0e38f, lo_y = 3.0e38f, hi_y = -3.0e38f;
  bool finite = true;
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int x = x0 + 32 * j, y = y_base + i;
      const int u = i * NJ + j;
      float gx = 0.f, gy = 0.f;
      if (x < p.w && y < p.h) {
        if (LENS) {
          lens_distort(L, (float)x, (float)y, gx, gy);
        } else {
          gx = __ldg(mxp + (size_t)y * p.w + x);
          gy = __ldg(myp + (size_t)y * p.w + x);
        }
        if (!p.normalized) {
          gx = R::sub(R::mul(fx, gx), 1.f);
          gy = R::sub(R::mul(fy, gy), 1.f);
        }
      }
      ux[u] = unnorm<ALIGN>(gx, Wm1, Wf);
      uy[u] = unnorm<ALIGN>(gy, Hm1, Hf);
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
  }
  // ---- 2. bounding box of the tile
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
    hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
    lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
    hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
  }
  finite = __all_sync(0xffffffffu, finite);
  if (lane == 0) {
    red[warp * 4 + 0] = finite ? lo_x : -3.0e38f;  // a non-finite coordinate anywhere disables the box
    red[warp * 4 + 1] = finite ? hi_x : 3.0e38f;
    red[warp * 4 + 2] = lo_y;
    red[warp * 4 + 3] = hi_y;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = red[0], bb = red[1], c = red[2], d = red[3];
    for (int wv = 1; wv < 8; ++wv) {
      a = fminf(a, red[wv * 4]); bb = fmaxf(bb, red[wv * 4 + 1]);
      c = fminf(c, red[wv * 4 + 2]); d = fmaxf(d, red[wv * 4 + 3]);
    }
    bool ok = a > -4.0e6f && bb < 4.0e6f && c > -4.0e6f && d < 4.0e6f && a <= bb && c <= d;
    const int x_lo = ok ? (int)floorf(a) : 0, x_hi = ok ? (int)floorf(bb) + 1 : 0;
    const int y_lo = ok ? (int)floorf(c) : 0, y_hi = ok ? (int)floorf(d) + 1 : 0;
    const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;
    const int spare = BW - need_w - 3;
    const int ox = (x_lo - (spare > 0 ? spare / 2 : 0)) & ~3;  // TMA: 16-byte aligned box start
    ok = ok && x_hi - ox + 1 <= BW && need_h <= BH;
    if (ok) {
      const int oy = y_lo - (BH - need_h) / 2;
      float wlx = (float)ox, whx = (float)(ox + BW - 1), wly = (float)oy, why = (float)(oy + BH - 1);
      if (INTERIOR) {
        wlx = fmaxf(wlx, 0.f); whx = fminf(whx, Wm1);
        wly = fmaxf(wly, 0.f); why = fminf(why, Hm1);
      }
      win[0] = wlx; win[1] = whx; win[2] = wly; win[3] = why;
      *kidx = (unsigned)(FLOOR_MAGIC_BITS + oy) * (unsigned)BW + (unsigned)(FLOOR_MAGIC_BITS + ox);
      tma::mbar_arrive_expect_tx(full, BOX_BYTES);
      tma::load_3d(box, &tmap, full, ox, oy, b * NC);
    } else {
      win[0] = win[2] = 1.f;
      win[1] = win[3] = 0.f;
      *kidx = 0;
      tma::mbar_arrive(full);
    }
  }
  tma::mbar_wait(full, 0);
  const float wlx = win[0], whx = win[1], wly = win[2], why = win[3];
  const uint32_t tbase = tma::smem_u32(box) - 4u * (*kidx);

  // ---- 3. sample
  const size_t splane = (size_t)H * W;
  const float* sp = p.src + (size_t)b * NC * splane;
  float* obase = p.out + (size_t)b * NC * oplane;
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int x = x0 + 32 * j, y = y_base + i;
      if (x >= p.w || y >= p.h) continue;
      const int u = i * NJ + j;
      float ix = ux[u], iy = uy[u];
      if (PRECLAMP) {
        ix = fminf(Wm1, fmaxf(ix, 0.f));
        iy = fminf(Hm1, fmaxf(iy, 0.f));
      }
      ix = interior_reflection<PAD, ALIGN>(ix);  // the window test and the taps below see the coordinate reflect_coord returns
      iy = interior_reflection<PAD, ALIGN>(iy);
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

int remap_tiled_forward(const float* src, const float* map_x, const float* map_y, float* out, int B, int C, int H, int W, int h, int w,
                        int Bmap, int normalized, int interp, int pad, int align, cudaStream_t st);
// undistort_image in one kernel: bilinear, zeros, align_corners=True, output size = input size; lens (B,16).
int undistort_tiled_forward(const float* src, const float* lens, float* out, int B, int C, int H, int W, cudaStream_t st);

}  // namespace kb200
