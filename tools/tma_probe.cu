// Stand-alone probe: which form of a 3-D TMA tile load works on this box?
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH, int NC, bool ELECT>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int ox, int oy, int oz) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* tile = reinterpret_cast<float*>(sm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + BW * BH * NC * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  bool issue;
  if (ELECT) {
    uint32_t pred = 0;
    if (threadIdx.x < 32) asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    issue = pred != 0;
  } else {
    issue = threadIdx.x == 0;
  }
  if (issue) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(BW * BH * NC * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(tile)),
                 "l"(&tmap), "r"(s32(bar)), "r"(ox), "r"(oy), "r"(oz)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(bar)) : "memory");
  }
  for (int i = threadIdx.x; i < BW * BH * NC; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int BH, int NC, bool ELECT>
int run(EncodeFn enc, float* d_src, int W, int H, int planes, float* d_out, const std::vector<float>& h_src, int ox, int oy, int oz, const char* tag) {
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {BW, BH, NC};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", tag, (int)r); return 0; }
  size_t smem = BW * BH * NC * 4 + 64;
  auto k = probe<BW, BH, NC, ELECT>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<1, 288, smem>>>(map, d_out, ox, oy, oz);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: KERNEL FAILED: %s\n", tag, cudaGetErrorString(e)); return 2; }
  std::vector<float> h(BW * BH * NC);
  CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int c = 0; c < NC; ++c) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
    int sx = ox + x, sy = oy + y, sc = oz + c;
    float want = (sx >= 0 && sx < W && sy >= 0 && sy < H && sc >= 0 && sc < planes) ? h_src[((size_t)sc * H + sy) * W + sx] : 0.f;
    if (h[(c * BH + y) * BW + x] != want) ++bad;
  }
  printf("%s: ok, mismatches=%d\n", tag, bad);
  return 0;
}

int main() {
  int W = 384, H = 216, planes = 6;
  std::vector<float> h_src((size_t)W * H * planes);
  for (size_t i = 0; i < h_src.size(); ++i) h_src[i] = (float)(i % 100003) * 0.25f;
  float *d_src, *d_out;
  CK(cudaMalloc(&d_src, h_src.size() * 4));
  CK(cudaMalloc(&d_out, 1 << 20));
  CK(cudaMemcpy(d_src, h_src.data(), h_src.size() * 4, cudaMemcpyHostToDevice));
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)ptr;
  int rc;
  struct { int ox, oy, oz; const char* tag; } cases[] = {
      {8, 4, 0, "aligned inside"}, {9, 4, 0, "ox=9 (unaligned) inside"}, {10, 4, 0, "ox=10"}, {12, 4, 0, "ox=12 (16B aligned)"},
      {352, 4, 0, "x overrun high, aligned"}, {8, 200, 0, "y overrun high"}, {8, 4, 4, "z overrun high"}, {-4, -3, 3, "negative aligned"},
      {-3, -3, 3, "negative unaligned"}, {350, 200, 3, "high edge unaligned"}};
  for (auto& c : cases) {
    rc = run<72, 40, 3, true>(enc, d_src, W, H, planes, d_out, h_src, c.ox, c.oy, c.oz, c.tag);
    if (rc == 2) { cudaDeviceReset(); CK(cudaMalloc(&d_src, h_src.size() * 4)); CK(cudaMalloc(&d_out, 1 << 20)); CK(cudaMemcpy(d_src, h_src.data(), h_src.size() * 4, cudaMemcpyHostToDevice)); }
  }
  return rc == 2;
}
