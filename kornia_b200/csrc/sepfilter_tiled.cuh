// kornia_b200 -- separable K x K filter, one pass over HBM (fp32, 'same' padding; the blur kernel).
//
// Replaces filter2d_separable's two (F.pad copy + grouped conv2d) passes (kornia/filters/filter.py:
// 205-207 -> :136-150) for square odd kernels: gaussian_blur2d(separable=True), SSIM windows, box blurs.
//
// One CTA per 128 x 32 output tile of one plane:
//   1. TMA (cp.async.bulk.tensor.2d) lands the (128+16) x (32+K-1) input box in shared memory; texels
//      outside the image arrive as zeros ('constant' border for free).  For 'reflect' / 'replicate'
//      only CTAs that touch the image border patch their out-of-image halo cells from cells of the
//      same tile (the folded source is always inside the tile).
//   2. row pass: each thread slides a K-tap window over 4 consecutive outputs (aligned LDS.128 in,
//      4 x K FMAs, STS.128 out) into a second shared tile;
//   3. column pass: each thread owns 2 adjacent columns x 8 rows and accumulates with packed
//      fma.rn.f32x2 (FFMA2) -- a column pair is a natural 64-bit register pair -- then streams out
//      with 8-byte stores.
// HBM traffic: 4 B read + 4 B written per element (24 B per RGB pixel) plus halo re-reads that hit
// in L2; the eager reference moves >= 32 B per element (two padded copies, two conv passes).
// Tap order (ascending, FMA) is the generic kernel's, so the two agree bit for bit.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include "filter_generic.cuh"
#include "warp_tma.cuh"

namespace kb200 {

constexpr int SEPT_TW = 128;
constexpr int SEPT_TH = 32;
constexpr int SEPT_XPAD = 8;                     // box starts 8 texels left of the tile: keeps the TMA start 16-B aligned
constexpr int SEPT_BW = SEPT_TW + 2 * SEPT_XPAD; // 144

struct SepTiledParams {
  const float* kx;  // (Bkx, K)
  const float* ky;  // (Bky, K)
  float* out;
  int C, H, W, Bkx, Bky;
};

__device__ __forceinline__ void tma_load_2d_plane(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  tma::load_3d(dst, map, bar, c0, c1, c2);
}

template <int K, int BORDER>
__global__ void __launch_bounds__(256) sepfilter_tiled_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const __grid_constant__ SepTiledParams p) {
  constexpr int HALO = (K - 1) / 2;
  static_assert(K % 2 == 1 && HALO <= SEPT_XPAD, "odd kernels up to 17 taps");
  constexpr int BH = SEPT_TH + K - 1;
  constexpr int BW = SEPT_BW;
  constexpr int TW = SEPT_TW, TH = SEPT_TH;
  constexpr int COL0 = SEPT_XPAD - HALO;        // tile column of the first tap of output x = 0
  constexpr int A0 = COL0 & 3;                  // its offset inside an aligned float4
  constexpr int NV = (A0 + 4 + K - 1 + 3) / 4;  // aligned float4 loads that cover the 4-output window

  extern __shared__ __align__(128) unsigned char sept_smem[];
  float* tile = reinterpret_cast<float*>(sept_smem);  // [BH][BW]
  float* mid = tile + BH * BW;                        // [BH][TW]
  uint64_t* bar = reinterpret_cast<uint64_t*>(mid + BH * TW);

  const int tid = threadIdx.x;
  const int plane = blockIdx.z, b = plane / p.C;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int ox = x0 - SEPT_XPAD, oy = y0 - HALO;  // box origin in image coordinates

  if (tid == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
  if (tid < 32 && tma::elect_one()) {
    tma::mbar_arrive_expect_tx(bar, BH * BW * 4);
    tma::load_3d(tile, &tmap, bar, ox, oy, plane);
  }
  // taps of this sample (filter.py:131,141-142: kernel index b mod Bk), fetched while the tile is in flight
  float kx[K], ky[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    kx[j] = __ldg(p.kx + (size_t)(b % p.Bkx) * K + j);
    ky[j] = __ldg(p.ky + (size_t)(b % p.Bky) * K + j);
  }
  tma::mbar_wait(bar, 0);

  if (BORDER != KB200_CONSTANT) {
    // CTA-uniform: does the box stick out of the image?
    if (ox < 0 || oy < 0 || ox + BW > p.W || oy + BH > p.H) {
      for (int e = tid; e < BH * BW; e += 256) {
        const int r = e / BW, c = e - r * BW;
        const int gy = oy + r, gx = ox + c;
        if ((unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) continue;
        const int fy = border_index<BORDER>(gy, p.H), fx = border_index<BORDER>(gx, p.W);
        const int sr = fy - oy, sc = fx - ox;
        // the folded source of every cell an output of this tile needs lies inside the tile; cells
        // further out (only reachable through the 8-texel alignment padding) are never read
        if ((unsigned)sr < (unsigned)BH && (unsigned)sc < (unsigned)BW) tile[e] = tile[sr * BW + sc];
      }
      __syncthreads();
    }
  }

  // ---------------------------------------------------------------- row pass: tile -> mid
  for (int item = tid; item < BH * (TW / 4); item += 256) {
    const int r = item / (TW / 4), q = item - r * (TW / 4);
    const float4* src4 = reinterpret_cast<const float4*>(tile + r * BW + (COL0 & ~3) + 4 * q);
    float win[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 t = src4[v];
      win[4 * v] = t.x; win[4 * v + 1] = t.y; win[4 * v + 2] = t.z; win[4 * v + 3] = t.w;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = __fmaf_rn(kx[j], win[A0 + o + j], acc[o]);
    }
    *reinterpret_cast<float4*>(mid + r * TW + 4 * q) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  __syncthreads();

  // ---------------------------------------------------------------- column pass: mid -> out
  {
    constexpr int RY = 8;
    static_assert((TW / 2) * (TH / RY) == 256, "one item per thread");
    const int cp = tid % (TW / 2), yb = tid / (TW / 2);
    const int x = x0 + 2 * cp, yrow = yb * RY;
    float2 ky2[K];
#pragma unroll
    for (int i = 0; i < K; ++i) ky2[i] = make_float2(ky[i], ky[i]);
    float2 acc[RY];
#pragma unroll
    for (int o = 0; o < RY; ++o) acc[o] = make_float2(0.f, 0.f);
    const float2* m2 = reinterpret_cast<const float2*>(mid + yrow * TW + 2 * cp);
#pragma unroll
    for (int i = 0; i < RY + K - 1; ++i) {
      const float2 v = m2[i * (TW / 2)];
#pragma unroll
      for (int o = 0; o < RY; ++o) {
        if (i - o >= 0 && i - o < K) acc[o] = __ffma2_rn(ky2[i - o], v, acc[o]);
      }
    }
    if (x < p.W) {
      float* op = p.out + (size_t)plane * p.H * p.W + (size_t)(y0 + yrow) * p.W + x;
#pragma unroll
      for (int o = 0; o < RY; ++o) {
        if (y0 + yrow + o < p.H) __stcs(reinterpret_cast<float2*>(op + (size_t)o * p.W), acc[o]);
      }
    }
  }
}

template <int K, int BORDER>
static int launch_sep_tiled(const CUtensorMap& map, const SepTiledParams& p, int planes, cudaStream_t st) {
  constexpr int BH = SEPT_TH + K - 1;
  constexpr size_t smem = (size_t)(BH * SEPT_BW + BH * SEPT_TW) * 4 + 16;
  auto kern = sepfilter_tiled_kernel<K, BORDER>;
  static bool configured = false;
  if (!configured) {
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const dim3 grid(ceil_div(p.W, SEPT_TW), ceil_div(p.H, SEPT_TH), planes);
  kern<<<grid, 256, smem, st>>>(map, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("sepfilter_tiled launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

// KB200_EUNSUPPORTED -> caller uses sepfilter_fwd_generic (circular border, 'valid', even / non-square /
// > 17-tap kernels, rows not 16-byte aligned, images narrower than the fold distance).
inline int sepfilter_tiled_forward(const float* x, const float* kx, const float* ky, float* out, int B, int C, int H, int W,
                                   int Bkx, int kw, int Bky, int kh, int border, int same, cudaStream_t st) {
  const char* off = getenv("KB200_DISABLE_TILED_FILTER");
  if (off && off[0] == '1') return KB200_EUNSUPPORTED;
  if (!same || kw != kh || (kw & 1) == 0 || kw < 3 || kw > 17 || border == KB200_CIRCULAR) return KB200_EUNSUPPORTED;
  if ((W % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 7) != 0) return KB200_EUNSUPPORTED;
  const int halo = (kw - 1) / 2;
  if (border != KB200_CONSTANT && (H <= halo || W <= halo)) return KB200_EUNSUPPORTED;  // fold source must be in the tile
  if ((long long)B * C > 65535) return KB200_EUNSUPPORTED;
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return KB200_EUNSUPPORTED;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)SEPT_BW, (cuuint32_t)(SEPT_TH + kw - 1), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return KB200_EUNSUPPORTED;
  SepTiledParams p{kx, ky, out, C, H, W, Bkx, Bky};
#define KB_SEP_CASE(K_)                                                                          \
  if (kw == K_) {                                                                                \
    if (border == KB200_CONSTANT) return launch_sep_tiled<K_, KB200_CONSTANT>(map, p, B * C, st); \
    if (border == KB200_REFLECT) return launch_sep_tiled<K_, KB200_REFLECT>(map, p, B * C, st);   \
    return launch_sep_tiled<K_, KB200_REPLICATE>(map, p, B * C, st);                              \
  }
  KB_SEP_CASE(3)
  KB_SEP_CASE(5)
  KB_SEP_CASE(7)
  KB_SEP_CASE(9)
  KB_SEP_CASE(11)
  KB_SEP_CASE(13)
  KB_SEP_CASE(15)
  KB_SEP_CASE(17)
#undef KB_SEP_CASE
  return KB200_EUNSUPPORTED;
}

}  // namespace kb200
