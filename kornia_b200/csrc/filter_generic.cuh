// kornia_b200 -- generic depthwise 2-D correlation with fused border handling (fp32/fp64).
//
// Replaces F.pad + view + F.conv2d(groups = planes) + view of kornia/filters/filter.py:136-150.
// The padded copy is never materialised: the border mode becomes index arithmetic.
#pragma once
#include "common.cuh"

namespace kb200 {

// Source index of padded position q (q counted from -front) on an axis of length n.
// Returns -1 for the zero region of 'constant'.  torch guarantees pad < n (reflect) / pad <= n
// (circular), so one fold suffices (torch/nn/functional.py pad contract used at filter.py:138).
template <int BORDER>
__device__ __forceinline__ int border_index(int q, int n) {
  if ((unsigned)q < (unsigned)n) return q;
  if (BORDER == KB200_CONSTANT) return -1;
  if (BORDER == KB200_REPLICATE) return q < 0 ? 0 : n - 1;
  if (BORDER == KB200_REFLECT) return q < 0 ? -q : 2 * (n - 1) - q;
  return q < 0 ? q + n : q - n;  // circular
}

template <typename T>
struct FilterParams {
  const T* x;       // (B,C,H,W)
  const T* k;       // (Bk,kh,kw)
  T* out;           // (B,C,Ho,Wo)
  int B, C, H, W, Bk, kh, kw;
  int top, left;    // front padding ('same') or 0 ('valid')
  int Ho, Wo;
};

template <typename T, int BORDER>
__global__ void __launch_bounds__(256) filter2d_fwd_generic(const FilterParams<T> p) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int plane = blockIdx.z;  // b*C + c
  if (x >= p.Wo || y >= p.Ho) return;
  const int b = plane / p.C;
  const T* xp = p.x + (size_t)plane * p.H * p.W;
  const T* kp = p.k + (size_t)(b % p.Bk) * p.kh * p.kw;  // filter.py:131,141-142 -> kernel b mod Bk
  T acc = T(0);
  for (int i = 0; i < p.kh; ++i) {
    const int sy = border_index<BORDER>(y + i - p.top, p.H);
    for (int j = 0; j < p.kw; ++j) {
      const int sx = border_index<BORDER>(x + j - p.left, p.W);
      const T v = (sy >= 0 && sx >= 0) ? ldg(xp + (size_t)sy * p.W + sx) : T(0);
      acc = RN<T>::fma(ldg(kp + i * p.kw + j), v, acc);
    }
  }
  st_stream(p.out + (size_t)plane * p.Ho * p.Wo + (size_t)y * p.Wo + x, acc);
}

// Padded positions q in [-front, n-1+rear] that fold onto source index s, as up to three ranges.
template <int BORDER>
__device__ __forceinline__ void preimage(int s, int n, int front, int rear, int lo[3], int hi[3]) {
  lo[0] = hi[0] = s;
  lo[1] = lo[2] = 0;
  hi[1] = hi[2] = -1;  // empty
  if (BORDER == KB200_REPLICATE) {
    if (s == 0 && front > 0) { lo[1] = -front; hi[1] = -1; }
    if (s == n - 1 && rear > 0) { lo[2] = n; hi[2] = n - 1 + rear; }
  } else if (BORDER == KB200_REFLECT) {
    if (s >= 1 && s <= front) lo[1] = hi[1] = -s;
    const int d = n - 1 - s;
    if (d >= 1 && d <= rear) lo[2] = hi[2] = n - 1 + d;
  } else if (BORDER == KB200_CIRCULAR) {
    if (s >= n - front) lo[1] = hi[1] = s - n;
    if (s <= rear - 1) lo[2] = hi[2] = s + n;
  }
}

// d/dx: adjoint of (border fold o correlate).  Gather form, deterministic (no atomics):
//   dpad[q] = sum_{i,j} k[i,j] * gout[q + front - i, ...];   gx[s] = sum_{q in preimage(s)} dpad[q]
template <typename T, int BORDER>
__global__ void __launch_bounds__(256) filter2d_bwd_input_generic(const FilterParams<T> p, const T* __restrict__ gout,
                                                                  T* __restrict__ gx, int bottom, int right) {
  const int sx = blockIdx.x * 32 + threadIdx.x;
  const int sy = blockIdx.y * 8 + threadIdx.y;
  const int plane = blockIdx.z;
  if (sx >= p.W || sy >= p.H) return;
  const int b = plane / p.C;
  const T* gp = gout + (size_t)plane * p.Ho * p.Wo;
  const T* kp = p.k + (size_t)(b % p.Bk) * p.kh * p.kw;
  int ylo[3], yhi[3], xlo[3], xhi[3];
  preimage<BORDER>(sy, p.H, p.top, bottom, ylo, yhi);
  preimage<BORDER>(sx, p.W, p.left, right, xlo, xhi);
  T acc = T(0);
  for (int ry = 0; ry < 3; ++ry)
    for (int qy = ylo[ry]; qy <= yhi[ry]; ++qy)
      for (int rx = 0; rx < 3; ++rx)
        for (int qx = xlo[rx]; qx <= xhi[rx]; ++qx) {
          // padded coordinates counted from 0: (qy + top, qx + left); out[y,x] touches pad[y+i, x+j]
          for (int i = 0; i < p.kh; ++i) {
            const int y = qy + p.top - i;
            if ((unsigned)y >= (unsigned)p.Ho) continue;
            for (int j = 0; j < p.kw; ++j) {
              const int x = qx + p.left - j;
              if ((unsigned)x >= (unsigned)p.Wo) continue;
              acc = RN<T>::fma(ldg(kp + i * p.kw + j), ldg(gp + (size_t)y * p.Wo + x), acc);
            }
          }
        }
  gx[(size_t)plane * p.H * p.W + (size_t)sy * p.W + sx] = acc;
}

// d/dk stage 1: grid = (kh*kw, B*C); each block reduces one tap over one plane -> ws[plane][tap].
template <typename T, int BORDER>
__global__ void __launch_bounds__(256) filter2d_bwd_kernel_stage1(const FilterParams<T> p, const T* __restrict__ gout,
                                                                  T* __restrict__ ws) {
  const int tap = blockIdx.x, plane = blockIdx.y;
  const int i = tap / p.kw, j = tap % p.kw;
  const T* xp = p.x + (size_t)plane * p.H * p.W;
  const T* gp = gout + (size_t)plane * p.Ho * p.Wo;
  double s = 0.0;
  const int n = p.Ho * p.Wo;
  for (int e = threadIdx.x; e < n; e += 256) {
    const int y = e / p.Wo, x = e - y * p.Wo;
    const int sy = border_index<BORDER>(y + i - p.top, p.H);
    const int sx = border_index<BORDER>(x + j - p.left, p.W);
    if (sy >= 0 && sx >= 0) s += (double)ldg(gp + e) * (double)ldg(xp + (size_t)sy * p.W + sx);
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) ws[(size_t)plane * p.kh * p.kw + tap] = (T)sh[0];
}

// stage 2: gk[bk][tap] = sum over planes (b,c) with b mod Bk == bk, fixed order.
template <typename T>
__global__ void filter2d_bwd_kernel_stage2(const T* __restrict__ ws, T* __restrict__ gk, int B, int C, int Bk, int taps) {
  const int tap = blockIdx.x * blockDim.x + threadIdx.x;
  const int bk = blockIdx.y;
  if (tap >= taps) return;
  double s = 0.0;
  for (int b = bk; b < B; b += Bk)
    for (int c = 0; c < C; ++c) s += (double)ws[((size_t)b * C + c) * taps + tap];
  gk[(size_t)bk * taps + tap] = (T)s;
}

// ------------------------------------------------------------------------------------------
// Separable filter, one pass over HBM (filter.py:205-207 fused): the input tile plus halo is
// staged in shared memory with the border fold applied by the loader, the row pass writes a
// second shared tile, the column pass streams the result out.
// ------------------------------------------------------------------------------------------
template <typename T>
struct SepParams {
  const T* x;
  const T* kx;   // (Bkx,kw)
  const T* ky;   // (Bky,kh)
  T* out;
  int B, C, H, W, Bkx, kw, Bky, kh, top, left, Ho, Wo;
};

constexpr int SEP_TW = 64;
constexpr int SEP_TH = 32;

template <typename T, int BORDER>
__global__ void __launch_bounds__(256) sepfilter_fwd_generic(const SepParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int in_w = SEP_TW + p.kw - 1, in_h = SEP_TH + p.kh - 1;
  T* tile = reinterpret_cast<T*>(smem_raw);  // [in_h][in_w]
  T* mid = tile + in_h * in_w;               // [in_h][SEP_TW]
  T* taps = mid + in_h * SEP_TW;             // [kw + kh]
  const int plane = blockIdx.z, b = plane / p.C;
  const int ox = blockIdx.x * SEP_TW, oy = blockIdx.y * SEP_TH;
  const int tid = threadIdx.x;
  const T* xp = p.x + (size_t)plane * p.H * p.W;
  for (int e = tid; e < p.kw; e += 256) taps[e] = ldg(p.kx + (size_t)(b % p.Bkx) * p.kw + e);
  for (int e = tid; e < p.kh; e += 256) taps[p.kw + e] = ldg(p.ky + (size_t)(b % p.Bky) * p.kh + e);
  for (int e = tid; e < in_h * in_w; e += 256) {
    const int r = e / in_w, c = e - r * in_w;
    const int sy = border_index<BORDER>(oy + r - p.top, p.H);
    const int sx = border_index<BORDER>(ox + c - p.left, p.W);
    tile[e] = (sy >= 0 && sx >= 0) ? ldg(xp + (size_t)sy * p.W + sx) : T(0);
  }
  __syncthreads();
  for (int e = tid; e < in_h * SEP_TW; e += 256) {
    const int r = e / SEP_TW, c = e - r * SEP_TW;
    T acc = T(0);
    for (int j = 0; j < p.kw; ++j) acc = RN<T>::fma(taps[j], tile[r * in_w + c + j], acc);
    mid[e] = acc;
  }
  __syncthreads();
  for (int e = tid; e < SEP_TH * SEP_TW; e += 256) {
    const int r = e / SEP_TW, c = e - r * SEP_TW;
    const int y = oy + r, x = ox + c;
    if (y >= p.Ho || x >= p.Wo) continue;
    T acc = T(0);
    for (int i = 0; i < p.kh; ++i) acc = RN<T>::fma(taps[p.kw + i], mid[(r + i) * SEP_TW + c], acc);
    st_stream(p.out + (size_t)plane * p.Ho * p.Wo + (size_t)y * p.Wo + x, acc);
  }
}

}  // namespace kb200
