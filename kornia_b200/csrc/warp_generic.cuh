// kornia_b200 -- generic (any C, any mode, fp32/fp64) fused warp / remap kernels.
//
// One thread per output pixel, channels looped in-thread, taps gathered through L1/L2 with
// per-tap bounds checks.  This is the correctness reference of the CUDA path and the fallback of
// the TMA-tiled kernel (warp_tma.cu) for shapes / modes the tile path does not cover.
//
// Replaces, per launch: create_meshgrid + ~15 broadcast elementwise kernels + stack +
// grid_sampler_2d (kornia/geometry/transform/imgwarp.py:157-174, :277-290, :293-320, :688-702).
#pragma once
#include "sampler.cuh"

namespace kb200 {

enum { KIND_AFFINE = 0, KIND_PROJ = 1, KIND_REMAP = 2 };

template <typename T>
struct WarpParams {
  const T* src;      // (B,C,H,W)
  const T* m;        // (Bm,3,3)           [affine / projective]
  const T* bx;       // (w)
  const T* by;       // (h)
  const T* map_x;    // (Bmap,h,w)         [remap]
  const T* map_y;
  const T* fill;     // (C) or null
  T* out;            // (B,C,h,w)
  int B, C, H, W, h, w;
  int Bm;            // matrix / map batch (B or 1)
  int align;
  int normalized;    // remap: maps already in [-1,1]
};

template <typename T>
struct WarpGradParams {
  WarpParams<T> p;   // p.out unused
  const T* gout;     // (B,C,h,w)
  T* gsrc;           // (B,C,H,W) zero-filled, or null
  T* partial;        // (B, nblk, 9) block partials for d/dm, or null
  T* gmap_x;         // (B,h,w) or null   [remap]
  T* gmap_y;
  int need_coord_grad;
};

constexpr int GEN_BX = 32;
constexpr int GEN_BY = 8;

// Normalised sampling coordinate of output pixel (x,y) of sample b, plus the pieces the
// backward pass needs.
template <typename T, int KIND>
struct Coord {
  T gx, gy, den, bxv, byv, fx, fy;
};

template <typename T, int KIND>
__device__ __forceinline__ Coord<T, KIND> coord_of(const WarpParams<T>& p, const Mat3<T>& m, int b, int x, int y) {
  using R = RN<T>;
  Coord<T, KIND> c;
  if (KIND == KIND_REMAP) {
    const size_t plane = (size_t)p.h * p.w;
    const size_t off = (p.Bm == 1 ? 0 : (size_t)b * plane) + (size_t)y * p.w + x;
    T mx = ldg(p.map_x + off), my = ldg(p.map_y + off);
    c.fx = c.fy = T(1);
    if (!p.normalized) {
      // conversions.py:1487-1498: factor = 2 / clamp(size - 1, eps); factor * p - 1
      c.fx = R::div(T(2), fmax(T(p.W - 1), T(1e-8)));
      c.fy = R::div(T(2), fmax(T(p.H - 1), T(1e-8)));
      mx = R::sub(R::mul(c.fx, mx), T(1));
      my = R::sub(R::mul(c.fy, my), T(1));
    }
    c.gx = mx;
    c.gy = my;
    c.den = T(1);
    c.bxv = c.byv = T(0);
  } else {
    c.bxv = ldg(p.bx + x);
    c.byv = ldg(p.by + y);
    map_point<T, KIND == KIND_PROJ>(m, c.bxv, c.byv, c.gx, c.gy, c.den);
    c.fx = c.fy = T(1);
  }
  return c;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename T, int INTERP, int PAD, int KIND>
__global__ void __launch_bounds__(GEN_BX* GEN_BY) warp_fwd_generic(const WarpParams<T> p) {
  const int x = blockIdx.x * GEN_BX + threadIdx.x;
  const int y = blockIdx.y * GEN_BY + threadIdx.y;
  const int b = blockIdx.z;
  if (x >= p.w || y >= p.h) return;
  Mat3<T> m;
  if (KIND != KIND_REMAP) m.load(p.m + (p.Bm == 1 ? 0 : (size_t)b * 9));
  const Coord<T, KIND> c = coord_of<T, KIND>(p, m, b, x, y);
  const bool align = p.align != 0;
  const int H = p.H, W = p.W;
  const size_t splane = (size_t)H * W;
  const size_t oplane = (size_t)p.h * p.w;
  const T* sp = p.src + (size_t)b * p.C * splane;
  T* op = p.out + (size_t)b * p.C * oplane + (size_t)y * p.w + x;

  PixelSampler<T, INTERP, PAD> S;
  S.prepare(unnormalize(c.gx, W, align), unnormalize(c.gy, H, align), H, W, align);
  for (int ch = 0; ch < p.C; ++ch) {
    const T v = S.sample(sp + ch * splane);
    st_stream(op + ch * oplane, S.finish(v, PAD == KB200_FILL ? ldg(p.fill + ch) : T(0)));
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int INTERP, int PAD, int KIND>
__global__ void __launch_bounds__(GEN_BX* GEN_BY) warp_bwd_generic(const WarpGradParams<T> g) {
  using R = RN<T>;
  constexpr int SPAD = (PAD == KB200_FILL) ? KB200_ZEROS : PAD;
  const WarpParams<T>& p = g.p;
  const int x = blockIdx.x * GEN_BX + threadIdx.x;
  const int y = blockIdx.y * GEN_BY + threadIdx.y;
  const int b = blockIdx.z;
  const bool live = x < p.w && y < p.h;
  const bool align = p.align != 0;
  const int H = p.H, W = p.W;
  const size_t splane = (size_t)H * W;
  const size_t oplane = (size_t)p.h * p.w;

  T dgx = T(0), dgy = T(0);  // d loss / d normalised coordinate
  Coord<T, KIND> c;
  c.gx = c.gy = c.bxv = c.byv = T(0);
  c.den = c.fx = c.fy = T(1);
  if (live) {
    Mat3<T> m;
    if (KIND != KIND_REMAP) m.load(p.m + (p.Bm == 1 ? 0 : (size_t)b * 9));
    c = coord_of<T, KIND>(p, m, b, x, y);
    const T* sp = p.src + (size_t)b * p.C * splane;
    T* gsp = g.gsrc ? g.gsrc + (size_t)b * p.C * splane : nullptr;
    const T* gop = g.gout + (size_t)b * p.C * oplane + (size_t)y * p.w + x;
    const bool want_coord = g.need_coord_grad != 0;
    // d(ix)/d(gx) of the unnormalisation (GridSampler.h:45-55)
    const T ux = align ? T(W - 1) * T(0.5) : T(W) * T(0.5);
    const T uy = align ? T(H - 1) * T(0.5) : T(H) * T(0.5);
    T ix = unnormalize(c.gx, W, align);
    T iy = unnormalize(c.gy, H, align);
    T gix = T(0), giy = T(0);

    if (INTERP == KB200_BILINEAR) {
      T px, py;
      ix = pad_coord_grad<T, SPAD>(ix, W, align, &px);
      iy = pad_coord_grad<T, SPAD>(iy, H, align, &py);
      const T x0f = R::floor(ix), y0f = R::floor(iy);
      const T wx1 = (x0f + T(1)) - ix, wx0 = ix - x0f;
      const T wy1 = (y0f + T(1)) - iy, wy0 = iy - y0f;
      const int x0 = (int)x0f, y0 = (int)y0f;
      const bool ok_nw = in_bounds(y0, x0, H, W), ok_ne = in_bounds(y0, x0 + 1, H, W);
      const bool ok_sw = in_bounds(y0 + 1, x0, H, W), ok_se = in_bounds(y0 + 1, x0 + 1, H, W);
      const int o = y0 * W + x0;
      for (int ch = 0; ch < p.C; ++ch) {
        const T go = ldg(gop + ch * oplane);
        if (gsp) {
          T* gs = gsp + ch * splane;
          if (ok_nw) atomicAdd(gs + o, wx1 * wy1 * go);
          if (ok_ne) atomicAdd(gs + o + 1, wx0 * wy1 * go);
          if (ok_sw) atomicAdd(gs + o + W, wx1 * wy0 * go);
          if (ok_se) atomicAdd(gs + o + W + 1, wx0 * wy0 * go);
        }
        if (want_coord) {
          const T* s = sp + ch * splane;
          const T f = (PAD == KB200_FILL) ? ldg(p.fill + ch) : T(0);  // out = fill + sum w (v - fill)
          const T v_nw = ok_nw ? ldg(s + o) - f : T(0);
          const T v_ne = ok_ne ? ldg(s + o + 1) - f : T(0);
          const T v_sw = ok_sw ? ldg(s + o + W) - f : T(0);
          const T v_se = ok_se ? ldg(s + o + W + 1) - f : T(0);
          gix += go * ((v_ne - v_nw) * wy1 + (v_se - v_sw) * wy0);
          giy += go * ((v_sw - v_nw) * wx1 + (v_se - v_ne) * wx0);
        }
      }
      dgx = gix * ux * px;
      dgy = giy * uy * py;
    } else if (INTERP == KB200_NEAREST) {
      ix = pad_coord<T, SPAD>(ix, W, align);
      iy = pad_coord<T, SPAD>(iy, H, align);
      const int xn = (int)R::rint(ix), yn = (int)R::rint(iy);
      if (gsp && in_bounds(yn, xn, H, W)) {
        for (int ch = 0; ch < p.C; ++ch) atomicAdd(gsp + ch * splane + yn * W + xn, ldg(gop + ch * oplane));
      }
    } else {  // bicubic
      const T fx = R::floor(ix), fy = R::floor(iy);
      const T tx = ix - fx, ty = iy - fy;
      T cx[4], cy[4], dx[4], dy[4];
      cubic_weights<T>(tx, cx);
      cubic_weights<T>(ty, cy);
      cubic_weights_grad<T>(tx, dx);
      cubic_weights_grad<T>(ty, dy);
      int xo[4], yo[4];
      bool xok[4], yok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int xi = (int)pad_coord<T, SPAD>(fx - T(1) + T(i), W, align);
        const int yi = (int)pad_coord<T, SPAD>(fy - T(1) + T(i), H, align);
        xok[i] = (unsigned)xi < (unsigned)W;
        yok[i] = (unsigned)yi < (unsigned)H;
        xo[i] = xi;
        yo[i] = yi * W;
      }
      for (int ch = 0; ch < p.C; ++ch) {
        const T go = ldg(gop + ch * oplane);
        const T f = (PAD == KB200_FILL) ? ldg(p.fill + ch) : T(0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!(yok[i] && xok[j])) continue;
            const int o = yo[i] + xo[j];
            if (gsp) atomicAdd(gsp + ch * splane + o, go * cx[j] * cy[i]);
            if (want_coord) {
              const T v = ldg(sp + ch * splane + o) - f;
              gix -= v * dx[j] * cy[i] * go;
              giy -= v * dy[i] * cx[j] * go;
            }
          }
        }
      }
      dgx = gix * ux;
      dgy = giy * uy;
    }
  }

  if (KIND == KIND_REMAP) {
    if (live && g.gmap_x) {
      const size_t off = (size_t)b * oplane + (size_t)y * p.w + x;
      g.gmap_x[off] = dgx * c.fx;
      g.gmap_y[off] = dgy * c.fy;
    }
    return;
  }
  if (!g.partial) return;

  // d/dm: nine per-pixel partials -> warp shuffle -> shared -> one row per block (deterministic)
  const T rden = (KIND == KIND_PROJ) ? T(1) / c.den : T(1);
  const T ax = dgx * rden, ay = dgy * rden;
  const T az = (KIND == KIND_PROJ) ? -(ax * c.gx + ay * c.gy) : T(0);
  T part[9] = {ax * c.bxv, ax * c.byv, ax, ay * c.bxv, ay * c.byv, ay, az * c.bxv, az * c.byv, az};
  __shared__ T red[GEN_BY][9];
  const int lane = threadIdx.x, wid = threadIdx.y;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const T s = warp_sum(live ? part[k] : T(0));
    if (lane == 0) red[wid][k] = s;
  }
  __syncthreads();
  if (wid == 0 && lane < 9) {
    T s = T(0);
#pragma unroll
    for (int r = 0; r < GEN_BY; ++r) s += red[r][lane];
    const size_t nblk = (size_t)gridDim.x * gridDim.y;
    const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    g.partial[((size_t)b * nblk + blk) * 9 + lane] = s;
  }
}

// Second stage of d/dm: fixed-order sum of the block partials (double accumulation).
// grid = (9, Bm); partial is (B, nblk, 9); Bm == 1 sums over the batch too.
template <typename T>
__global__ void __launch_bounds__(256) warp_gm_reduce(const T* __restrict__ partial, T* __restrict__ gm, int B, int Bm,
                                                      long long nblk) {
  const int k = blockIdx.x, bm = blockIdx.y;
  const long long rows = (Bm == 1) ? (long long)B * nblk : nblk;
  const T* base = partial + (Bm == 1 ? 0 : (size_t)bm * nblk * 9) + k;
  double s = 0.0;
  for (long long r = threadIdx.x; r < rows; r += 256) s += (double)base[r * 9];
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gm[(size_t)bm * 9 + k] = (T)sh[0];
}

}  // namespace kb200
