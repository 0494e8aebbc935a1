// kornia_b200 -- host side of the image-derivative kernels (gradient.cuh).
#include "gradient.cuh"
#include "gradient_tiled.cuh"

namespace kb200 {

template <typename T>
static int fill_params(GradParams<T>& p, const void* x, const double* taps, void* out, int planes, int H, int W, int nout, int k,
                       double eps) {
  KB_CHECK_ARG(taps, "null taps");
  KB_CHECK_ARG(planes > 0 && H > 0 && W > 0, "bad sizes planes=%d H=%d W=%d", planes, H, W);
  KB_CHECK_ARG((k == 3 || k == 5) && nout >= 1 && nout <= GRAD_MAX_OUT, "unsupported stencil: %d outputs of %dx%d taps", nout, k, k);
  p.x = (const T*)x;
  p.out = (T*)out;
  p.planes = planes;
  p.H = H;
  p.W = W;
  p.vec = (W % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  p.eps = (T)eps;
  for (int e = 0; e < nout * k * k; ++e) p.taps[e] = (T)taps[e];  // exact: the host built them in T
  return KB200_OK;
}

template <typename T>
static int forward_t(const void* x, const double* taps, void* out, int planes, int H, int W, int nout, int k, int magnitude,
                     double eps, cudaStream_t st) {
  GradParams<T> p;
  int rc = fill_params(p, x, taps, out, planes, H, W, nout, k, eps);
  if (rc) return rc;
  KB_CHECK_ARG(x && out, "null tensor");
  KB_CHECK_ARG(!magnitude || nout == 2, "the magnitude needs exactly two derivative outputs");
  const dim3 block(32, 8), grid(ceil_div(W, 128), ceil_div(H, 8), min(planes, 65535));
  if (magnitude) {
    KB_CHECK_ARG(k == 3, "the magnitude is defined on first-order 3x3 stencils");
    spatial_gradient_fwd<T, 3, 2, true><<<grid, block, 0, st>>>(p);
  } else if (k == 3 && nout == 2) {
    spatial_gradient_fwd<T, 3, 2, false><<<grid, block, 0, st>>>(p);
  } else if (k == 3 && nout == 3) {
    spatial_gradient_fwd<T, 3, 3, false><<<grid, block, 0, st>>>(p);
  } else if (k == 5 && nout == 3) {
    spatial_gradient_fwd<T, 5, 3, false><<<grid, block, 0, st>>>(p);
  } else {
    set_error("unsupported stencil: %d outputs of %dx%d taps", nout, k, k);
    return KB200_EINVAL;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("spatial_gradient_forward: kernel launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

template <typename T>
static int backward_t(const void* gout, const double* taps, void* gx, int planes, int H, int W, int nout, int k, cudaStream_t st) {
  GradParams<T> p;
  int rc = fill_params(p, nullptr, taps, nullptr, planes, H, W, nout, k, 0.0);
  if (rc) return rc;
  KB_CHECK_ARG(gout && gx, "null tensor");
  const dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), min(planes, 65535));
  spatial_gradient_bwd<T><<<grid, block, 0, st>>>(p, (const T*)gout, (T*)gx, k, nout);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("spatial_gradient_backward: kernel launch failed: %s", cudaGetErrorString(e));
    return KB200_ECUDA;
  }
  return KB200_OK;
}

int spatial_gradient_forward(const void* x, const double* taps, void* out, int planes, int H, int W, int nout, int k,
                             int magnitude, double eps, int dtype, cudaStream_t st) {
  if (dtype == KB200_F32) {  // TMA tile loader (opt-in, gradient_tiled.cuh); anything it declines runs the kernel below
    const int rc = spatial_gradient_tiled_forward((const float*)x, taps, (float*)out, planes, H, W, nout, k, magnitude, eps, st);
    if (rc != KB200_EUNSUPPORTED) return rc;
  }
  return dtype == KB200_F32 ? forward_t<float>(x, taps, out, planes, H, W, nout, k, magnitude, eps, st)
                            : forward_t<double>(x, taps, out, planes, H, W, nout, k, magnitude, eps, st);
}

int spatial_gradient_backward(const void* gout, const double* taps, void* gx, int planes, int H, int W, int nout, int k,
                              int dtype, cudaStream_t st) {
  return dtype == KB200_F32 ? backward_t<float>(gout, taps, gx, planes, H, W, nout, k, st)
                            : backward_t<double>(gout, taps, gx, planes, H, W, nout, k, st);
}

}  // namespace kb200
