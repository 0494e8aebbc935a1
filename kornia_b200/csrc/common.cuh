// kornia_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kornia_b200.h"

namespace kb200 {

// ---------------------------------------------------------------- error plumbing (capi.cu)
void set_error(const char* fmt, ...);

#define KB_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::kb200::set_error(__VA_ARGS__);   \
      return KB200_EINVAL;               \
    }                                    \
  } while (0)

#define KB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ::kb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return KB200_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

// ---------------------------------------------------------------- separately rounded arithmetic
// The eager reference evaluates the coordinate chain as one torch kernel per op, so every
// product / sum is rounded on its own (no FMA contraction).  The _rn intrinsics are never
// contracted by nvcc, which keeps the chain bit-faithful.
template <typename T>
struct RN;

template <>
struct RN<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float floor(float a) { return floorf(a); }
  static __device__ __forceinline__ float rint(float a) { return rintf(a); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float fmod(float a, float b) { return fmodf(a, b); }
};

template <>
struct RN<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double floor(double a) { return ::floor(a); }
  static __device__ __forceinline__ double rint(double a) { return ::rint(a); }
  static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
  static __device__ __forceinline__ double fmod(double a, double b) { return ::fmod(a, b); }
};

// read-only, L1-allocating global load (gathers reuse neighbouring taps through L1)
template <typename T>
__device__ __forceinline__ T ldg(const T* p) {
  return __ldg(p);
}

// streaming store: the output is written once and never re-read by this kernel
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per device: remember, per kernel instantiation, on which devices it was applied
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

}  // namespace kb200
