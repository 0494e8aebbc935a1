// kornia_b200 -- image derivatives: NOUT small stencils over one read of the image (fp32/fp64).
//
// Replaces spatial_gradient's F.pad(replicate) copy + F.conv2d with a (NOUT,1,k,k) weight
// (kornia/filters/sobel.py:59-74) and, with MAG, the whole of sobel() (sobel.py:158-167: two
// strided slices, three elementwise kernels and a sqrt on top of it).  The padded copy is never
// materialised (the replicate border is a clamp on the tap index) and every output of a pixel is
// produced from the same window in registers:
//   spatial_gradient: 4 B read + 4*NOUT B written per element (reference: >= 8 + 4 + 4*NOUT)
//   sobel magnitude : 4 B read + 4 B written per element     (reference: >= 12 + 8 + 9*4 more)
// Each thread produces 4 neighbouring outputs of one row: per tap row one aligned 16-byte load plus
// 2*(k/2) halo scalars, then NOUT*k*4 FMAs with the taps as kernel-parameter (constant bank) operands.
#pragma once
#include "filter_generic.cuh"

namespace kb200 {

constexpr int GRAD_MAX_OUT = 3;
constexpr int GRAD_MAX_K = 5;

template <typename T>
struct GradParams {
  const T* x;  // (planes,H,W)
  T* out;      // (planes,NOUT,H,W), or (planes,H,W) for the magnitude
  int planes, H, W;
  int vec;     // rows start 16-byte aligned and W % 4 == 0: vector loads / stores allowed
  T eps;       // magnitude only
  T taps[GRAD_MAX_OUT * GRAD_MAX_K * GRAD_MAX_K];  // [NOUT][K][K], correlation order
};

__device__ __forceinline__ float sqrt_rn(float v) { return __fsqrt_rn(v); }
__device__ __forceinline__ double sqrt_rn(double v) { return __dsqrt_rn(v); }

template <typename T, int K, int NOUT, bool MAG>
__global__ void __launch_bounds__(256) spatial_gradient_fwd(const __grid_constant__ GradParams<T> p) {
  static_assert(K <= GRAD_MAX_K && NOUT <= GRAD_MAX_OUT && (!MAG || NOUT == 2), "stencil limits");
  constexpr int HALO = K / 2;
  const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x0 >= p.W || y >= p.H) return;
  const bool inner = x0 - HALO >= 0 && x0 + 3 + HALO < p.W;  // no clamp needed along x
  const bool full = x0 + 3 < p.W;
  const size_t HW = (size_t)p.H * p.W;

  for (int plane = blockIdx.z; plane < p.planes; plane += gridDim.z) {
    const T* xp = p.x + (size_t)plane * HW;
    T acc[NOUT][4];
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[o][q] = T(0);

#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int sy = min(max(y + i - HALO, 0), p.H - 1);
      const T* row = xp + (size_t)sy * p.W;
      T win[4 + K - 1];
      if (inner) {
        if (sizeof(T) == 4 && p.vec) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(row + x0));
          win[HALO] = c.x; win[HALO + 1] = c.y; win[HALO + 2] = c.z; win[HALO + 3] = c.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) win[HALO + q] = ldg(row + x0 + q);
        }
#pragma unroll
        for (int e = 0; e < HALO; ++e) {
          win[e] = ldg(row + x0 - HALO + e);
          win[HALO + 4 + e] = ldg(row + x0 + 4 + e);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4 + K - 1; ++e) win[e] = ldg(row + min(max(x0 - HALO + e, 0), p.W - 1));
      }
#pragma unroll
      for (int j = 0; j < K; ++j)
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
          const T t = p.taps[(o * K + i) * K + j];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[o][q] = RN<T>::fma(t, win[q + j], acc[o][q]);
        }
    }

    if (MAG) {
      // sqrt(gx*gx + gy*gy + eps): one rounding per reference op (sobel.py:166)
      T* op = p.out + (size_t)plane * HW + (size_t)y * p.W + x0;
      T m[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        m[q] = sqrt_rn(RN<T>::add(RN<T>::add(RN<T>::mul(acc[0][q], acc[0][q]), RN<T>::mul(acc[1][q], acc[1][q])), p.eps));
      if (full && p.vec && sizeof(T) == 4) {
        __stcs(reinterpret_cast<float4*>(op), make_float4(m[0], m[1], m[2], m[3]));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (x0 + q < p.W) st_stream(op + q, m[q]);
      }
    } else {
#pragma unroll
      for (int o = 0; o < NOUT; ++o) {
        T* op = p.out + ((size_t)plane * NOUT + o) * HW + (size_t)y * p.W + x0;
        if (full && p.vec && sizeof(T) == 4) {
          __stcs(reinterpret_cast<float4*>(op), make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (x0 + q < p.W) st_stream(op + q, acc[o][q]);
        }
      }
    }
  }
}

// Adjoint of (replicate pad o NOUT correlations): gather form, deterministic.
//   gx[s] = sum_o sum_{q in preimage(s)} sum_{i,j} t_o[i,j] * gout[o][q + HALO - (i,j)]
template <typename T>
__global__ void __launch_bounds__(256) spatial_gradient_bwd(const __grid_constant__ GradParams<T> p, const T* __restrict__ gout,
                                                            T* __restrict__ gx, int K, int NOUT) {
  const int sx = blockIdx.x * 32 + threadIdx.x;
  const int sy = blockIdx.y * 8 + threadIdx.y;
  if (sx >= p.W || sy >= p.H) return;
  const int HALO = K / 2;
  const size_t HW = (size_t)p.H * p.W;
  int ylo[3], yhi[3], xlo[3], xhi[3];
  preimage<KB200_REPLICATE>(sy, p.H, HALO, HALO, ylo, yhi);
  preimage<KB200_REPLICATE>(sx, p.W, HALO, HALO, xlo, xhi);
  for (int plane = blockIdx.z; plane < p.planes; plane += gridDim.z) {
    T acc = T(0);
    for (int o = 0; o < NOUT; ++o) {
      const T* gp = gout + ((size_t)plane * NOUT + o) * HW;
      for (int ry = 0; ry < 3; ++ry)
        for (int qy = ylo[ry]; qy <= yhi[ry]; ++qy)
          for (int rx = 0; rx < 3; ++rx)
            for (int qx = xlo[rx]; qx <= xhi[rx]; ++qx)
              for (int i = 0; i < K; ++i) {
                const int y = qy + HALO - i;
                if ((unsigned)y >= (unsigned)p.H) continue;
                for (int j = 0; j < K; ++j) {
                  const int x = qx + HALO - j;
                  if ((unsigned)x >= (unsigned)p.W) continue;
                  acc = RN<T>::fma(p.taps[(o * K + i) * K + j], ldg(gp + (size_t)y * p.W + x), acc);
                }
              }
    }
    gx[(size_t)plane * HW + (size_t)sy * p.W + sx] = acc;
  }
}

int spatial_gradient_forward(const void* x, const double* taps, void* out, int planes, int H, int W, int nout, int k,
                             int magnitude, double eps, int dtype, cudaStream_t st);
int spatial_gradient_backward(const void* gout, const double* taps, void* gx, int planes, int H, int W, int nout, int k,
                              int dtype, cudaStream_t st);

}  // namespace kb200
