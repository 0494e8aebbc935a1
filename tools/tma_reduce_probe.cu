// Probe: cp.reduce.async.bulk.tensor.3d .add on f32 (smem box -> global += ), incl. partially out-of-bounds boxes.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <stdlib.h>
#include <math.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH, int NC>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int ox, int oy, int oz, int reps) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* tile = reinterpret_cast<float*>(sm) + (threadIdx.x / 32) * BW * BH * NC;   // one box per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < reps; ++r) {
    for (int i = lane; i < BW * BH * NC; i += 32) tile[i] = 1.0f + 0.001f * (i % 7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint32_t pred = 0;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    if (pred) {
      asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap), "r"(s32(tile)),
                   "r"(ox), "r"(oy + warp * 3), "r"(oz)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int W = 384, H = 216, planes = 6, BW = 72, BH = 8, NC = 3, WARPS = 4;
  std::vector<float> h((size_t)W * H * planes, 0.5f);
  float* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)ptr;
  CUtensorMap map;
  cuuint64_t dims[3] = {W, H, planes}; cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {BW, BH, NC}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  struct C { int ox, oy, oz, reps; const char* tag; };
  C one{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), "case"};
  C cases[] = {one};
  size_t smem = (size_t)WARPS * BW * BH * NC * 4;
  CK(cudaFuncSetAttribute(probe<BW, BH, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> want = h;
  for (auto& c : cases) {
    probe<BW, BH, NC><<<1, WARPS * 32, smem>>>(map, c.ox, c.oy, c.oz, c.reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: KERNEL FAILED: %s\n", c.tag, cudaGetErrorString(e)); return 1; }
    for (int rep = 0; rep < c.reps; ++rep) for (int w = 0; w < WARPS; ++w) for (int ch = 0; ch < NC; ++ch) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
      int sx = c.ox + x, sy = c.oy + w * 3 + y, sc = c.oz + ch;
      if (sx >= 0 && sx < W && sy >= 0 && sy < H && sc >= 0 && sc < planes) want[((size_t)sc * H + sy) * W + sx] += 1.0f + 0.001f * (((ch * BH + y) * BW + x) % 7);
    }
    std::vector<float> got(h.size());
    CK(cudaMemcpy(got.data(), d, got.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (size_t i = 0; i < got.size(); ++i) { double er = fabs(got[i] - want[i]); if (er > 1e-4) ++bad; if (er > maxerr) maxerr = er; }
    printf("%s: ok, mismatches=%d maxerr=%g\n", c.tag, bad, maxerr);
  }
  return 0;
}
