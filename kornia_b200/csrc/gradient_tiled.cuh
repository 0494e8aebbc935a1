// kornia_b200 -- image derivatives on the TMA tile loader (fp32, K in {3,5}, NOUT in {2,3}, replicate border).
//
// Same arithmetic as spatial_gradient_fwd (gradient.cuh: NOUT stencils per pixel from one window, taps as constant-bank
// operands, row-major ascending FMA order -- bit-identical), on the skeleton of filter2d_tiled_kernel: persistent CTAs
// walk strips of 128 x 32 tiles, the (128+16) x (32+K-1) input box is TMA-loaded into shared memory and double
// buffered, the replicate border (kornia/filters/sobel.py:66-68: F.pad(..., 'replicate')) is a patch of the box on edge
// tiles.  A thread produces 4 columns x R rows of every output plane from a (R+K-1)-row register window fed by aligned
// LDS.128 -- the L1-gather kernel it replaces re-reads every input row K times through L1 with tiny CTAs (measured
// 50-66 % of the HBM roofline).  MAG fuses sobel's sqrt(gx^2 + gy^2 + eps) (sobel.py:158-167).
//
// Status: written after the round-1 GPU budget was spent; compiled for sm_100a, not yet run on hardware.  Dispatched
// only when KB200_TILED_GRADIENT=1 (gradient.cu).
#pragma once
#include "filter2d_tiled.cuh"
#include "gradient.cuh"

namespace kb200 {

struct GradTiledParams {
  float* out;  // (planes,NOUT,H,W), or (planes,H,W) for the magnitude
  int H, W, planes;
  float eps;
  float taps[GRAD_MAX_OUT * GRAD_MAX_K * GRAD_MAX_K];  // [NOUT][K][K]
};

template <int K, int NOUT, bool MAG>
__global__ void __launch_bounds__(256, 3) grad_tiled_kernel(const __grid_constant__ CUtensorMap tmap,
                                                            const __grid_constant__ GradTiledParams p) {
  static_assert((K == 3 || K == 5) && NOUT >= 2 && NOUT <= GRAD_MAX_OUT && (!MAG || (NOUT == 2 && K == 3)), "stencil limits");
  constexpr int HALO = (K - 1) / 2;
  constexpr int TW = SEPT_TW, TH = SEPT_TH, BW = SEPT_BW;
  constexpr int BH = TH + K - 1;
  constexpr int COL0 = SEPT_XPAD - HALO;
  constexpr int A0 = COL0 & 3;
  constexpr int NV = (A0 + 4 + K - 1 + 3) / 4;
  constexpr int TILE_FLOATS = BH * BW;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4;
  constexpr int R = NOUT == 3 ? 2 : 4;      // output rows per thread and sweep
  constexpr int SWEEPS = TH / (8 * R);      // 8 row groups of R rows per sweep
  static_assert(TW / 4 == 32 && SWEEPS * 8 * R == TH, "thread mapping: 32 quads x 8 row groups x SWEEPS");

  extern __shared__ __align__(128) unsigned char gradt_smem[];
  float* tiles = reinterpret_cast<float*>(gradt_smem);  // [2][BH][BW]
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + 2 * TILE_FLOATS);

  const int tid = threadIdx.x;
  if (tid == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  const int tiles_x = ceil_div(p.W, TW), tiles_y = ceil_div(p.H, TH);
  const Segments segs(p.planes * tiles_y, tiles_x);

  struct Ahead {
    int seg, strip, tx0, tx1, cursor, tx, n;
    bool live;
  } ah{0, 0, 0, 0, 0, 0, 0, false};
  auto ahead_next = [&]() {
    if (ah.live && ah.tx + 1 < ah.tx1) {
      ++ah.tx;
      ++ah.n;
      return;
    }
    const bool first = !ah.live && ah.n == 0 && ah.seg == 0;
    ah.live = segs.get(ah.seg, ah.strip, ah.tx0, ah.tx1, ah.cursor);
    ++ah.seg;
    ah.tx = ah.tx0;
    if (!first) ++ah.n;
  };
  auto issue = [&]() {
    if (!ah.live) return;
    const int plane = ah.strip / tiles_y, ty = ah.strip - plane * tiles_y;
    const int s = ah.n & 1;
    tma::fence_proxy_async();
    tma::mbar_arrive_expect_tx(&full[s], TILE_BYTES);
    tma::load_3d(tiles + s * TILE_FLOATS, &tmap, &full[s], ah.tx * TW - SEPT_XPAD, ty * TH - HALO, plane);
    ahead_next();
  };
  if (tid == 0) {
    ahead_next();
    issue();
    issue();
  }

  const int q = tid & 31, rg = tid >> 5;  // quad q (columns 4q..4q+3), row group rg
  const size_t HW = (size_t)p.H * p.W;

  int n = 0, strip, tx0, tx1, cursor = 0;
  for (int seg = 0; segs.get(seg, strip, tx0, tx1, cursor); ++seg) {
    const int plane = strip / tiles_y, ty = strip - plane * tiles_y;
    const int y0 = ty * TH;
    float* oplane = p.out + (size_t)plane * (MAG ? 1 : NOUT) * HW;

    for (int tx = tx0; tx < tx1; ++tx, ++n) {
      const int s = n & 1;
      float* tile = tiles + s * TILE_FLOATS;
      tma::mbar_wait(&full[s], (n >> 1) & 1);

      {  // replicate border: cells of the box outside the image copy the nearest image cell (which lies inside the box)
        const int ox = tx * TW - SEPT_XPAD, oy = y0 - HALO;
        if (ox < 0 || oy < 0 || ox + BW > p.W || oy + BH > p.H) {
          for (int e = tid; e < BH * BW; e += 256) {
            const int r = e / BW, c = e - r * BW;
            const int gy = oy + r, gx = ox + c;
            if ((unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) continue;
            const int sr = border_index<KB200_REPLICATE>(gy, p.H) - oy, sc = border_index<KB200_REPLICATE>(gx, p.W) - ox;
            if ((unsigned)sr < (unsigned)BH && (unsigned)sc < (unsigned)BW) tile[e] = tile[sr * BW + sc];
          }
          __syncthreads();
        }
      }

      const int gx0 = tx * TW + 4 * q;
#pragma unroll
      for (int sw = 0; sw < SWEEPS; ++sw) {
        const int row0 = (sw * 8 + rg) * R;  // first output row of this thread inside the tile
        float acc[NOUT][R][4];
#pragma unroll
        for (int o = 0; o < NOUT; ++o)
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[o][r][x] = 0.f;
        const float4* src4 = reinterpret_cast<const float4*>(tile + row0 * BW + (COL0 & ~3) + 4 * q);
#pragma unroll
        for (int wr = 0; wr < R + K - 1; ++wr) {
          float win[NV * 4];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const float4 t = src4[wr * (BW / 4) + v];
            win[4 * v] = t.x; win[4 * v + 1] = t.y; win[4 * v + 2] = t.z; win[4 * v + 3] = t.w;
          }
          // window row wr is tap row i = wr - r of output row r: i ascends with wr for every output (row-major tap order)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int i = wr - r;
            if (i >= 0 && i < K) {
#pragma unroll
              for (int j = 0; j < K; ++j)
#pragma unroll
                for (int o = 0; o < NOUT; ++o) {
                  const float t = p.taps[(o * K + i) * K + j];
#pragma unroll
                  for (int x = 0; x < 4; ++x) acc[o][r][x] = __fmaf_rn(t, win[A0 + x + j], acc[o][r][x]);
                }
            }
          }
        }
        if (gx0 < p.W) {  // W % 4 == 0: whole quads
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int gy = y0 + row0 + r;
            if (gy < p.H) {
              if (MAG) {
                float m[4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
                  m[x] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(acc[0][r][x], acc[0][r][x]), __fmul_rn(acc[1][r][x], acc[1][r][x])), p.eps));
                __stcs(reinterpret_cast<float4*>(oplane + (size_t)gy * p.W + gx0), make_float4(m[0], m[1], m[2], m[3]));
              } else {
#pragma unroll
                for (int o = 0; o < NOUT; ++o)
                  __stcs(reinterpret_cast<float4*>(oplane + (size_t)o * HW + (size_t)gy * p.W + gx0),
                         make_float4(acc[o][r][0], acc[o][r][1], acc[o][r][2], acc[o][r][3]));
              }
            }
          }
        }
      }
      __syncthreads();  // tile[s] consumed by every thread
      if (tid == 0) issue();
    }
  }
}

// KB200_EUNSUPPORTED -> the caller runs spatial_gradient_fwd.
int spatial_gradient_tiled_forward(const float* x, const double* taps, float* out, int planes, int H, int W, int nout, int k, int magnitude,
                                   double eps, cudaStream_t st);

}  // namespace kb200
