// kornia_b200 -- coordinate transforms of the sampler (device).
//
// Semantics: ATen grid_sampler_2d as called by the reference at
// kornia/geometry/transform/imgwarp.py:174,290,316,320,702.  Spec followed:
// torch/include/ATen/native/GridSampler.h:27-35 (unnormalise), :58-60 (clip), :89-105 (reflect),
// :143-160 (compute_coordinates), UpSample.h:398-423 (cubic convolution, A = -0.75), and the
// CUDA flavour's non-finite guard (cuda/GridSampler.cuh: out-of-int-range / NaN -> -100).
#pragma once
#include "common.cuh"

namespace kb200 {

template <typename T>
__device__ __forceinline__ T unnormalize(T g, int size, bool align) {
  using R = RN<T>;
  if (align) return R::mul(R::mul(R::add(g, T(1)), T(0.5)), T(size - 1));
  return R::mul(R::sub(R::mul(R::add(g, T(1)), T(size)), T(1)), T(0.5));
}

template <typename T>
__device__ __forceinline__ T clip_coord(T c, int size) {
  return fmin(T(size - 1), fmax(c, T(0)));
}

template <typename T>
__device__ __forceinline__ T reflect_coord(T c, int twice_low, int twice_high) {
  using R = RN<T>;
  if (twice_low == twice_high) return T(0);
  const T lo = T(twice_low) * T(0.5);
  const T span = T(twice_high - twice_low) * T(0.5);
  c = R::abs(R::sub(c, lo));
  const T extra = R::fmod(c, span);
  const int flips = static_cast<int>(R::floor(R::div(c, span)));
  return (flips % 2 == 0) ? R::add(extra, lo) : R::add(R::sub(span, extra), lo);
}

// values that cannot be an index (NaN, +-inf, beyond int range) become -100: out of bounds
template <typename T>
__device__ __forceinline__ T guard_index(T c) {
  return (RN<T>::abs(c) <= T(2147483648.0)) ? c : T(-100);
}

template <typename T, int PAD>
__device__ __forceinline__ T pad_coord(T c, int size, bool align) {
  if (PAD == KB200_BORDER) {
    c = clip_coord(c, size);
  } else if (PAD == KB200_REFLECTION) {
    c = align ? reflect_coord(c, 0, 2 * (size - 1)) : reflect_coord(c, -1, 2 * size - 1);
    c = clip_coord(c, size);
  }
  return guard_index(c);
}

// Same, also returning d(out)/d(in) (GridSampler.h:176-203): borders count as out of bounds.
template <typename T, int PAD>
__device__ __forceinline__ T pad_coord_grad(T c, int size, bool align, T* mult) {
  using R = RN<T>;
  T g = T(1);
  if (PAD == KB200_REFLECTION) {
    const int tl = align ? 0 : -1;
    const int th = align ? 2 * (size - 1) : 2 * size - 1;
    if (tl == th) {
      c = T(0);
      g = T(0);
    } else {
      const T lo = T(tl) * T(0.5);
      const T span = T(th - tl) * T(0.5);
      c = R::sub(c, lo);
      T sgn = T(1);
      if (c < T(0)) {
        sgn = T(-1);
        c = -c;
      }
      const T extra = R::fmod(c, span);
      const int flips = static_cast<int>(R::floor(R::div(c, span)));
      if (flips % 2 == 0) {
        c = R::add(extra, lo);
        g = sgn;
      } else {
        c = R::add(R::sub(span, extra), lo);
        g = -sgn;
      }
    }
  }
  if (PAD == KB200_BORDER || PAD == KB200_REFLECTION) {
    const T hi = T(size - 1);
    if (c <= T(0)) {
      c = T(0);
      g = T(0);
    } else if (c >= hi) {
      c = hi;
      g = T(0);
    }
  }
  *mult = g;
  return guard_index(c);
}

// cubic convolution weights for the four taps at offsets -1, 0, 1, 2 (A = -0.75)
template <typename T>
__device__ __forceinline__ void cubic_weights(T t, T w[4]) {
  const T A = T(-0.75);
  const T x0 = t + T(1);
  w[0] = ((A * x0 - T(5) * A) * x0 + T(8) * A) * x0 - T(4) * A;
  w[1] = ((A + T(2)) * t - (A + T(3))) * t * t + T(1);
  const T u = T(1) - t;
  w[2] = ((A + T(2)) * u - (A + T(3))) * u * u + T(1);
  const T x3 = u + T(1);
  w[3] = ((A * x3 - T(5) * A) * x3 + T(8) * A) * x3 - T(4) * A;
}

// d(weights)/dt (GridSampler.h get_cubic_coefficients_grad)
template <typename T>
__device__ __forceinline__ void cubic_weights_grad(T t, T w[4]) {
  const T A = T(-0.75);
  T x = T(-1) - t;
  w[0] = (T(-3) * A * x - T(10) * A) * x - T(8) * A;
  x = -t;
  w[1] = (T(-3) * (A + T(2)) * x - T(2) * (A + T(3))) * x;
  x = T(1) - t;
  w[2] = (T(3) * (A + T(2)) * x - T(2) * (A + T(3))) * x;
  x = T(2) - t;
  w[3] = (T(3) * A * x - T(10) * A) * x + T(8) * A;
}

__device__ __forceinline__ bool in_bounds(int y, int x, int H, int W) {
  return (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
}

// Per-sample 3x3 matrix held in registers; the map of A.3 with the reference's association order.
template <typename T>
struct Mat3 {
  T m00, m01, m02, m10, m11, m12, m20, m21, m22;
  __device__ __forceinline__ void load(const T* p) {
    m00 = p[0]; m01 = p[1]; m02 = p[2];
    m10 = p[3]; m11 = p[4]; m12 = p[5];
    m20 = p[6]; m21 = p[7]; m22 = p[8];
  }
};

template <typename T, bool PROJ>
__device__ __forceinline__ void map_point(const Mat3<T>& m, T bx, T by, T& gx, T& gy, T& den) {
  using R = RN<T>;
  // imgwarp.py:167-169 / :279-280: (m_i0 * x + m_i1 * y) + m_i2, then a true division
  T nx = R::add(R::add(R::mul(m.m00, bx), R::mul(m.m01, by)), m.m02);
  T ny = R::add(R::add(R::mul(m.m10, bx), R::mul(m.m11, by)), m.m12);
  if (PROJ) {
    den = R::add(R::add(R::mul(m.m20, bx), R::mul(m.m21, by)), m.m22);
    gx = R::div(nx, den);
    gy = R::div(ny, den);
  } else {
    den = T(1);
    gx = nx;
    gy = ny;
  }
}

}  // namespace kb200
