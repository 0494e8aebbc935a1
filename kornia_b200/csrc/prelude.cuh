// kornia_b200 -- fused (B,3,3) prelude: m = inverse( N_dst @ (M @ inverse(N_src)) ).
//
// Replaces the ~35 tiny torch launches of normalize_homography + _inverse_3x3_closed_form
// (kornia/geometry/conversions.py:1717-1725, kornia/core/utils.py:159-166) with one launch when no
// gradient w.r.t. M is needed.  The arithmetic mimics what those torch CUDA kernels compute, so the
// result is bit-identical to the torch prelude on the same device (asserted by
// tests/test_parity_gpu.py::test_fused_prelude_bit_identical; the host layer falls back to the
// torch ops when M requires grad).  `variant` selects between plausible contraction orders of the
// third-party kernels; the test pins the one that matches.
#pragma once
#include "common.cuh"

namespace kb200 {

template <typename T>
__device__ __forceinline__ T cross_term(T a, T b, T c, T d, int variant) {  // a*b - c*d as at::cross_kernel rounds it
  using R = RN<T>;
  if (variant & 1) return R::fma(-c, d, R::mul(a, b));
  return R::fma(a, b, -R::mul(c, d));
}

template <typename T>
__device__ __forceinline__ void inv3_torchlike(const T A[9], T out[9], int variant) {
  using R = RN<T>;
  // columns a, b, c ; rows of the adjugate are b x c, c x a, a x b  (utils.py:159-164)
  const T a0 = A[0], a1 = A[3], a2 = A[6];
  const T b0 = A[1], b1 = A[4], b2 = A[7];
  const T c0 = A[2], c1 = A[5], c2 = A[8];
  T r[9];
  r[0] = cross_term(b1, c2, b2, c1, variant); r[1] = cross_term(b2, c0, b0, c2, variant); r[2] = cross_term(b0, c1, b1, c0, variant);
  r[3] = cross_term(c1, a2, c2, a1, variant); r[4] = cross_term(c2, a0, c0, a2, variant); r[5] = cross_term(c0, a1, c1, a0, variant);
  r[6] = cross_term(a1, b2, a2, b1, variant); r[7] = cross_term(a2, b0, a0, b2, variant); r[8] = cross_term(a0, b1, a1, b0, variant);
  // det = (col_a * row0).sum(-1): products rounded by the mul kernel, then summed left to right
  const T p0 = R::mul(a0, r[0]), p1 = R::mul(a1, r[1]), p2 = R::mul(a2, r[2]);
  T det;
  if (variant & 4) det = R::add(R::add(p0, p2), p1);       // strided two-lane reduction: (x0 + x2) + x1
  else if (variant & 8) det = R::add(p0, R::add(p1, p2));
  else det = R::add(R::add(p0, p1), p2);
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = R::div(r[i], det);
}

template <typename T>
__device__ __forceinline__ void matmul3_torchlike(const T A[9], const T Bm[9], T C[9], int variant) {
  using R = RN<T>;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T acc;
      if (variant & 2) {
        acc = R::add(R::add(R::mul(A[i * 3], Bm[j]), R::mul(A[i * 3 + 1], Bm[3 + j])), R::mul(A[i * 3 + 2], Bm[6 + j]));
      } else {  // GEMM inner loop: fused multiply-add over k, accumulator starts at zero
        acc = R::fma(A[i * 3], Bm[j], T(0));
        acc = R::fma(A[i * 3 + 1], Bm[3 + j], acc);
        acc = R::fma(A[i * 3 + 2], Bm[6 + j], acc);
      }
      C[i * 3 + j] = acc;
    }
}

template <typename T>
__global__ void warp_prelude_kernel(const T* __restrict__ M, T* __restrict__ out, int B, int rows, float sx_s, float sy_s,
                                    float sx_d, float sy_d, int variant) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  T Mh[9];
  for (int i = 0; i < rows * 3; ++i) Mh[i] = M[(size_t)b * rows * 3 + i];
  if (rows == 2) {  // convert_affinematrix_to_homography: pad a zero row, then += 1 on the corner
    Mh[6] = T(0);
    Mh[7] = T(0);
    Mh[8] = RN<T>::add(T(0), T(1));
  }
  const T Ns[9] = {T(sx_s), T(0), T(-1), T(0), T(sy_s), T(-1), T(0), T(0), T(1)};
  const T Nd[9] = {T(sx_d), T(0), T(-1), T(0), T(sy_d), T(-1), T(0), T(0), T(1)};
  T Nsi[9], X[9], Mn[9], m[9];
  inv3_torchlike<T>(Ns, Nsi, variant);
  matmul3_torchlike<T>(Mh, Nsi, X, variant);
  matmul3_torchlike<T>(Nd, X, Mn, variant);
  inv3_torchlike<T>(Mn, m, variant);
#pragma unroll
  for (int i = 0; i < 9; ++i) out[(size_t)b * 9 + i] = m[i];
}

// Backward of the prelude: given m = inverse(Nd @ M3 @ inverse(Ns)) and gm = dL/dm, returns dL/dM3 (rows x 3).
//   Mn = Nd M3 Nsi,  m = Mn^-1   =>   dL/dMn = -m^T gm m^T ,   dL/dM3 = Nd^T dL/dMn Nsi^T
// Plain arithmetic (double accumulation): gradients carry no bit-exactness contract.
template <typename T>
__global__ void warp_prelude_backward_kernel(const T* __restrict__ m, const T* __restrict__ gm, T* __restrict__ gM, int B, int rows,
                                             float sx_s, float sy_s, float sx_d, float sy_d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double mm[9], g[9], t[9], gn[9];
  for (int i = 0; i < 9; ++i) {
    mm[i] = (double)m[(size_t)b * 9 + i];
    g[i] = (double)gm[(size_t)b * 9 + i];
  }
  // t = m^T g
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[i * 3 + j] = mm[0 * 3 + i] * g[0 * 3 + j] + mm[1 * 3 + i] * g[1 * 3 + j] + mm[2 * 3 + i] * g[2 * 3 + j];
  // gn = -(t m^T)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) gn[i * 3 + j] = -(t[i * 3 + 0] * mm[j * 3 + 0] + t[i * 3 + 1] * mm[j * 3 + 1] + t[i * 3 + 2] * mm[j * 3 + 2]);
  // Nd = [[a,0,-1],[0,c,-1],[0,0,1]] ; Nsi = inverse of [[p,0,-1],[0,q,-1],[0,0,1]] = [[1/p,0,1/p],[0,1/q,1/q],[0,0,1]]
  const double a = (double)(T)sx_d, c = (double)(T)sy_d, ip = 1.0 / (double)(T)sx_s, iq = 1.0 / (double)(T)sy_s;
  // u = Nd^T gn : rows (a*gn0, c*gn1, -gn0 - gn1 + gn2)
  double u[9];
  for (int j = 0; j < 3; ++j) {
    u[0 * 3 + j] = a * gn[0 * 3 + j];
    u[1 * 3 + j] = c * gn[1 * 3 + j];
    u[2 * 3 + j] = -gn[0 * 3 + j] - gn[1 * 3 + j] + gn[2 * 3 + j];
  }
  // gM = u Nsi^T : column k of Nsi^T is row k of Nsi
  for (int i = 0; i < rows; ++i) {
    const double u0 = u[i * 3], u1 = u[i * 3 + 1], u2 = u[i * 3 + 2];
    gM[((size_t)b * rows + i) * 3 + 0] = (T)(u0 * ip + u2 * ip);
    gM[((size_t)b * rows + i) * 3 + 1] = (T)(u1 * iq + u2 * iq);
    gM[((size_t)b * rows + i) * 3 + 2] = (T)u2;
  }
}

}  // namespace kb200

namespace kb200 {

// Unit square -> quadrilateral (Heckbert), every binary op rounded on its own like the reference's
// chain of elementwise torch kernels (kornia/geometry/transform/imgwarp.py:411-441).
template <typename T>
__device__ __forceinline__ void square_to_quad(const T* __restrict__ pts, T Q[9]) {
  using R = RN<T>;
  const T x0 = pts[0], y0 = pts[1], x1 = pts[2], y1 = pts[3], x2 = pts[4], y2 = pts[5], x3 = pts[6], y3 = pts[7];
  const T dx1 = R::sub(x1, x2), dx2 = R::sub(x3, x2), sx = R::sub(R::add(R::sub(x0, x1), x2), x3);
  const T dy1 = R::sub(y1, y2), dy2 = R::sub(y3, y2), sy = R::sub(R::add(R::sub(y0, y1), y2), y3);
  const T den = R::sub(R::mul(dx1, dy2), R::mul(dy1, dx2));
  const T g = R::div(R::sub(R::mul(sx, dy2), R::mul(sy, dx2)), den);
  const T h = R::div(R::sub(R::mul(dx1, sy), R::mul(dy1, sx)), den);
  Q[0] = R::add(R::sub(x1, x0), R::mul(g, x1)); Q[1] = R::add(R::sub(x3, x0), R::mul(h, x3)); Q[2] = x0;
  Q[3] = R::add(R::sub(y1, y0), R::mul(g, y1)); Q[4] = R::add(R::sub(y3, y0), R::mul(h, y3)); Q[5] = y0;
  Q[6] = g; Q[7] = h; Q[8] = T(1);
}

// get_perspective_transform (imgwarp.py:456-462) in one launch: H = Q(dst) @ inverse(Q(src)), scaled to H[2,2] = 1.
// Replaces ~45 tiny torch launches per call on the RandomPerspective / crop_and_resize path.
template <typename T>
__global__ void perspective_from_points_kernel(const T* __restrict__ src, const T* __restrict__ dst, T* __restrict__ out, int B,
                                               int variant) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  T Qs[9], Qd[9], Qsi[9], Hm[9];
  square_to_quad<T>(src + (size_t)b * 8, Qs);
  square_to_quad<T>(dst + (size_t)b * 8, Qd);
  inv3_torchlike<T>(Qs, Qsi, variant);
  matmul3_torchlike<T>(Qd, Qsi, Hm, variant);
  const T w = Hm[8];
#pragma unroll
  for (int i = 0; i < 9; ++i) out[(size_t)b * 9 + i] = RN<T>::div(Hm[i], w);
}

}  // namespace kb200

namespace kb200 {

__device__ __forceinline__ float cos_t(float v) { return cosf(v); }
__device__ __forceinline__ double cos_t(double v) { return cos(v); }
__device__ __forceinline__ float sin_t(float v) { return sinf(v); }
__device__ __forceinline__ double sin_t(double v) { return sin(v); }

// get_rotation_matrix2d (imgwarp.py:607-622) in one launch: T(c) @ R(angle) @ S(scale) @ T(-c), the three 3x3
// products accumulated like torch's batched GEMM, the angle converted like deg2rad on the device (conversions.py:148:
// times the fp32 constant pi, then "divided" by 180 the way torch's CUDA scalar division does it).  Replaces ~35 tiny torch launches on the rotate / scale / RandomAffine path.
template <typename T>
__global__ void rotation_matrix2d_kernel(const T* __restrict__ center, const T* __restrict__ angle, const T* __restrict__ scale,
                                         T* __restrict__ out, int B, int variant) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  using R = RN<T>;
  const T cx = center[2 * b], cy = center[2 * b + 1];
  // deg2rad: (angle * pi) / 180.0 -- torch's CUDA division by a host scalar multiplies by the rounded reciprocal
  const T rad = R::mul(R::mul(angle[b], (T)3.14159265358979323846f), R::div(T(1), T(180)));
  const T c = cos_t(rad), s = sin_t(rad);
  const T to_c[9] = {T(1), T(0), cx, T(0), T(1), cy, T(0), T(0), T(1)};
  const T from_c[9] = {T(1), T(0), -cx, T(0), T(1), -cy, T(0), T(0), T(1)};
  const T rot[9] = {c, s, T(0), -s, c, T(0), T(0), T(0), T(1)};
  const T scl[9] = {R::mul(T(1), scale[2 * b]), T(0), T(0), T(0), R::mul(T(1), scale[2 * b + 1]), T(0), T(0), T(0), T(1)};
  T a[9], ab[9], abc[9];
  matmul3_torchlike<T>(to_c, rot, a, variant);
  matmul3_torchlike<T>(a, scl, ab, variant);
  matmul3_torchlike<T>(ab, from_c, abc, variant);
#pragma unroll
  for (int i = 0; i < 6; ++i) out[(size_t)b * 6 + i] = abc[i];
}

}  // namespace kb200
