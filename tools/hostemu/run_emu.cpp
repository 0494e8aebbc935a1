// Executes the kernel templates of kornia_b200/csrc on the CPU (hostemu.h) and compares them BIT FOR BIT with scalar
// restatements of the same arithmetic.  Covers the kernels written after the round-1 GPU budget was spent (DESIGN.md
// section 9) and, as a check of the emulator itself, the hardware-verified kernels they derive from.
//
//   make -C tools/hostemu && tools/hostemu/run_emu          (or: python tools/hostemu/run.py)
//
// x86 fmaf() is correctly rounded and the build uses -ffp-contract=off, so fp32 results are those of the device
// intrinsics (__fmaf_rn, __fmul_rn, ...).  What this does NOT cover: timing, occupancy, bank conflicts, the real TMA
// unit and the memory model of the async proxy -- a pass here means "the index arithmetic, the pipeline bookkeeping and
// the barriers are consistent", not "runs on a B200".
#include "hostemu.h"

#include "../../kornia_b200/csrc/filter2d_tiled.cuh"
#include "../../kornia_b200/csrc/gradient_tiled.cuh"
#include "../../kornia_b200/csrc/sepfilter_vwalk.cuh"
#include "../../kornia_b200/csrc/ssim_vwalk.cuh"

#include <random>
#include <string>

namespace kb200 {
alignas(128) unsigned char sept_smem[256 * 1024];
alignas(128) unsigned char sepv_smem[256 * 1024];
alignas(128) unsigned char f2d_smem[256 * 1024];
alignas(128) unsigned char gradt_smem[256 * 1024];
alignas(128) unsigned char ssimv_smem[256 * 1024];
alignas(128) float ssim_smem[64 * 1024];
void set_error(const char*, ...) {}
}  // namespace kb200

using namespace kb200;

static std::mt19937 rng(1234);
static std::vector<float> randv(size_t n, float lo = 0.f, float hi = 1.f) {
  std::uniform_real_distribution<float> d(lo, hi);
  std::vector<float> v(n);
  for (auto& x : v) x = d(rng);
  return v;
}
static float* aligned(std::vector<float>& store, size_t n) {  // 128-byte aligned view into a vector
  store.assign(n + 64, -777.f);
  return reinterpret_cast<float*>(((size_t)store.data() + 127) & ~(size_t)127);
}
static int fold(int q, int n, int border) {
  if (q >= 0 && q < n) return q;
  if (border == KB200_CONSTANT) return -1;
  if (border == KB200_REPLICATE) return q < 0 ? 0 : n - 1;
  return q < 0 ? -q : 2 * (n - 1) - q;
}
static int failures = 0;
static void compare(const std::string& what, const float* got, const float* want, size_t n) {
  size_t bad = 0, first = 0;
  for (size_t i = 0; i < n; ++i)
    if (memcmp(got + i, want + i, 4) != 0 && !(got[i] == 0.f && want[i] == 0.f)) {
      if (!bad) first = i;
      ++bad;
    }
  if (bad) {
    printf("FAIL %-72s %zu / %zu differ, first at %zu: got %.9g want %.9g\n", what.c_str(), bad, n, first, got[first], want[first]);
    ++failures;
  } else {
    printf("ok   %s\n", what.c_str());
  }
}

// ------------------------------------------------------------------------------------------ separable filter
static void ref_sepfilter(const float* x, const float* kx, const float* ky, float* out, int planes, int C, int H, int W, int Bk, int K, int border) {
  const int h = (K - 1) / 2;
  std::vector<float> mid((size_t)H * W);
  for (int p = 0; p < planes; ++p) {
    const float* xp = x + (size_t)p * H * W;
    const float* kxp = kx + (size_t)((p / C) % Bk) * K;
    for (int y = 0; y < H; ++y)
      for (int xx = 0; xx < W; ++xx) {
        float a = 0.f;
        for (int j = 0; j < K; ++j) {
          const int sx = fold(xx + j - h, W, border);
          a = fmaf(kxp[j], sx < 0 ? 0.f : xp[(size_t)y * W + sx], a);
        }
        mid[(size_t)y * W + xx] = a;
      }
    for (int y = 0; y < H; ++y)
      for (int xx = 0; xx < W; ++xx) {
        float a = 0.f;
        for (int i = 0; i < K; ++i) {
          const int sy = fold(y + i - h, H, border);
          a = fmaf(ky[i], sy < 0 ? 0.f : mid[(size_t)sy * W + xx], a);
        }
        out[(size_t)p * H * W + (size_t)y * W + xx] = a;
      }
  }
}

template <int K, int BORDER>
static void test_sepfilter(int B, int C, int H, int W, unsigned grid, bool lazy) {
  emu::lazy_tma = lazy;
  const int planes = B * C;
  std::vector<float> xs, o1s, o2s;
  float* x = aligned(xs, (size_t)planes * H * W);
  for (size_t i = 0; i < (size_t)planes * H * W; ++i) x[i] = randv(1)[0];
  auto kx = randv((size_t)B * K, -1.f, 1.f), ky = randv(K, -1.f, 1.f);
  std::vector<float> want((size_t)planes * H * W);
  ref_sepfilter(x, kx.data(), ky.data(), want.data(), planes, C, H, W, B, K, BORDER);
  float* o1 = aligned(o1s, want.size());
  float* o2 = aligned(o2s, want.size());
  const std::string tag = "K=" + std::to_string(K) + " border=" + std::to_string(BORDER) + " " + std::to_string(B) + "x" + std::to_string(C) + "x" +
                          std::to_string(H) + "x" + std::to_string(W) + " grid=" + std::to_string(grid) + (lazy ? " lazy" : " eager");
  {  // the hardware-verified strip-walking kernel: checks the emulator
    const CUtensorMap map = emu::make_map(x, W, H, planes, SEPT_BW, SEPT_TH + K - 1, 1);
    SepTiledParams p{kx.data(), ky.data(), o1, C, H, W, B, 1, planes, x, 0.f};
    emu::launch(grid, dim3(256), [&] { sepfilter_tiled_kernel<K, BORDER, false>(map, p); });
    compare("sepfilter_tiled_kernel (verified on hw) " + tag, o1, want.data(), want.size());
  }
  {
    const CUtensorMap main = emu::make_map(x, W, H, planes, SEPT_BW, SEPT_TH, 1), pro = emu::make_map(x, W, H, planes, SEPT_BW, K - 1, 1);
    SepTiledParams p{kx.data(), ky.data(), o2, C, H, W, B, 1, planes, x, 0.f};
    emu::launch(grid, dim3(256), [&] { sepfilter_vwalk_kernel<K, BORDER>(main, pro, p); });
    compare("sepfilter_vwalk_kernel                  " + tag, o2, want.data(), want.size());
  }
}

// ------------------------------------------------------------------------------------------ filter2d / pyrdown / derivatives
static void ref_filter2d(const float* x, const float* k, int nout, float* out, int planes, int H, int W, int K, int border) {
  const int h = (K - 1) / 2;
  for (int p = 0; p < planes; ++p)
    for (int o = 0; o < nout; ++o)
      for (int y = 0; y < H; ++y)
        for (int xx = 0; xx < W; ++xx) {
          float a = 0.f;
          for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
              const int sy = fold(y + i - h, H, border), sx = fold(xx + j - h, W, border);
              a = fmaf(k[(o * K + i) * K + j], (sy < 0 || sx < 0) ? 0.f : x[(size_t)p * H * W + (size_t)sy * W + sx], a);
            }
          out[(((size_t)p * nout + o) * H + y) * W + xx] = a;
        }
}

template <int BORDER>
static void test_pyrdown(int planes, int H, int W, unsigned grid, bool lazy) {
  emu::lazy_tma = lazy;
  std::vector<float> xs, os, fs;
  float* x = aligned(xs, (size_t)planes * H * W);
  for (size_t i = 0; i < (size_t)planes * H * W; ++i) x[i] = randv(1)[0];
  float taps[25];
  const float b5[5] = {1, 4, 6, 4, 1};
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) taps[i * 5 + j] = b5[i] * b5[j] / 256.f;
  std::vector<float> full((size_t)planes * H * W), want((size_t)planes * (H / 2) * (W / 2));
  ref_filter2d(x, taps, 1, full.data(), planes, H, W, 5, BORDER);
  for (int p = 0; p < planes; ++p)
    for (int y = 0; y < H / 2; ++y)
      for (int xx = 0; xx < W / 2; ++xx) {
        const float* f = full.data() + (size_t)p * H * W + (size_t)(2 * y) * W + 2 * xx;
        want[((size_t)p * (H / 2) + y) * (W / 2) + xx] = 0.25f * ((f[0] + f[1]) + (f[W] + f[W + 1]));
      }
  const std::string tag = "border=" + std::to_string(BORDER) + " " + std::to_string(planes) + "x" + std::to_string(H) + "x" + std::to_string(W) +
                          " grid=" + std::to_string(grid) + (lazy ? " lazy" : " eager");
  const CUtensorMap map = emu::make_map(x, W, H, planes, SEPT_BW, SEPT_TH + 4, 1);
  {
    float* o = aligned(fs, full.size());
    F2dTiledParams p{taps, o, 1, H, W, 1, planes};
    emu::launch(grid, dim3(256), [&] { filter2d_tiled_kernel<5, BORDER, false>(map, p); });
    compare("filter2d_tiled_kernel<5> (verified on hw) " + tag, o, full.data(), full.size());
  }
  {
    float* o = aligned(os, want.size());
    F2dTiledParams p{taps, o, 1, H, W, 1, planes};
    emu::launch(grid, dim3(256), [&] { filter2d_tiled_kernel<5, BORDER, true>(map, p); });
    compare("filter2d_tiled_kernel<5, DOWN2> (pyrdown)  " + tag, o, want.data(), want.size());
  }
}

template <int K, int NOUT, bool MAG>
static void test_gradient(int planes, int H, int W, unsigned grid, bool lazy) {
  emu::lazy_tma = lazy;
  std::vector<float> xs, os;
  float* x = aligned(xs, (size_t)planes * H * W);
  for (size_t i = 0; i < (size_t)planes * H * W; ++i) x[i] = randv(1)[0];
  GradTiledParams p;
  p.H = H; p.W = W; p.planes = planes; p.eps = 1e-6f;
  auto t = randv(NOUT * K * K, -1.f, 1.f);
  for (int e = 0; e < NOUT * K * K; ++e) p.taps[e] = (e % 4 == 1) ? 0.f : t[e];  // some exact zeros, like the Sobel stencils
  std::vector<float> d((size_t)planes * NOUT * H * W), want;
  ref_filter2d(x, p.taps, NOUT, d.data(), planes, H, W, K, KB200_REPLICATE);
  if (MAG) {
    want.resize((size_t)planes * H * W);
    for (int pl = 0; pl < planes; ++pl)
      for (size_t i = 0; i < (size_t)H * W; ++i) {
        const float gx = d[((size_t)pl * 2) * H * W + i], gy = d[((size_t)pl * 2 + 1) * H * W + i];
        want[(size_t)pl * H * W + i] = sqrtf((gx * gx + gy * gy) + p.eps);
      }
  } else {
    want = d;
  }
  p.out = aligned(os, want.size());
  const CUtensorMap map = emu::make_map(x, W, H, planes, SEPT_BW, SEPT_TH + K - 1, 1);
  emu::launch(grid, dim3(256), [&] { grad_tiled_kernel<K, NOUT, MAG>(map, p); });
  compare("grad_tiled_kernel K=" + std::to_string(K) + " NOUT=" + std::to_string(NOUT) + (MAG ? " MAG " : "     ") + std::to_string(planes) + "x" +
              std::to_string(H) + "x" + std::to_string(W) + " grid=" + std::to_string(grid) + (lazy ? " lazy" : " eager"),
          p.out, want.data(), want.size());
}

// ------------------------------------------------------------------------------------------ SSIM
template <int K>
static void test_ssim(int planes, int H, int W, unsigned grid, bool lazy) {
  emu::lazy_tma = lazy;
  std::vector<float> as, bs, o1s, o2s;
  const size_t n = (size_t)planes * H * W;
  float* a = aligned(as, n);
  float* b = aligned(bs, n);
  for (size_t i = 0; i < n; ++i) {
    a[i] = randv(1)[0];
    b[i] = std::min(1.f, std::max(0.f, a[i] + randv(1, -0.1f, 0.1f)[0]));
  }
  auto taps = randv(K, 0.f, 1.f);
  const float C1 = 1e-4f, C2 = 9e-4f, eps = 1e-12f;
  const std::string tag = "K=" + std::to_string(K) + " " + std::to_string(planes) + "x" + std::to_string(H) + "x" + std::to_string(W) +
                          " grid=" + std::to_string(grid) + (lazy ? " lazy" : " eager");
  float* o1 = aligned(o1s, n);
  {  // the hardware-verified gather kernel is the reference here
    SsimParams p;
    p.a = a; p.b = b; p.taps = taps.data(); p.out = o1; p.planes = planes; p.H = H; p.W = W;
    p.tiles_x = ceil_div(W, SSIM_TW); p.tiles_y = ceil_div(H, SSIM_TH); p.pair_ok = (W % 2 == 0); p.C1 = C1; p.C2 = C2; p.eps = eps;
    emu::launch((unsigned)(p.tiles_x * p.tiles_y * planes), dim3(256), [&] { ssim_tiled_kernel<K>(p); });
  }
  float* o2 = aligned(o2s, n);
  {
    const CUtensorMap ma = emu::make_map(a, W, H, planes, SSIMV_BW, SSIMV_TH, 1), mb = emu::make_map(b, W, H, planes, SSIMV_BW, SSIMV_TH, 1);
    const CUtensorMap pa = emu::make_map(a, W, H, planes, SSIMV_BW, K - 1, 1), pb = emu::make_map(b, W, H, planes, SSIMV_BW, K - 1, 1);
    SsimVParams p{taps.data(), o2, planes, H, W, C1, C2, eps};
    emu::launch(grid, dim3(256), [&] { ssim_vwalk_kernel<K>(ma, mb, pa, pb, p); });
  }
  compare("ssim_vwalk_kernel vs ssim_tiled_kernel " + tag, o2, o1, n);
}

int main() {
  for (int lazy = 0; lazy < 2; ++lazy) {
    // grids that do not divide the number of strips / bands: segments that start in the middle of a band
    test_sepfilter<11, KB200_REFLECT>(2, 3, 70, 132, 5, lazy);
    test_sepfilter<11, KB200_REPLICATE>(1, 2, 97, 260, 4, lazy);
    test_sepfilter<11, KB200_CONSTANT>(1, 1, 33, 4, 1, lazy);
    test_sepfilter<3, KB200_REFLECT>(2, 1, 32, 128, 2, lazy);
    test_sepfilter<5, KB200_REPLICATE>(1, 1, 6, 8, 1, lazy);
    test_sepfilter<17, KB200_REFLECT>(1, 2, 130, 140, 3, lazy);
    test_sepfilter<17, KB200_CONSTANT>(1, 1, 64, 388, 7, lazy);
    test_pyrdown<KB200_REFLECT>(3, 70, 132, 4, lazy);
    test_pyrdown<KB200_REPLICATE>(2, 34, 260, 3, lazy);
    test_pyrdown<KB200_CONSTANT>(1, 4, 4, 1, lazy);
    test_gradient<3, 2, false>(3, 70, 132, 4, lazy);
    test_gradient<3, 2, true>(2, 33, 260, 3, lazy);
    test_gradient<3, 3, false>(1, 3, 4, 1, lazy);
    test_gradient<5, 3, false>(2, 97, 136, 5, lazy);
    test_ssim<11>(3, 70, 132, 4, lazy);
    test_ssim<5>(2, 33, 8, 1, lazy);
    test_ssim<3>(1, 32, 64, 1, lazy);
    test_ssim<7>(2, 97, 260, 7, lazy);
    test_ssim<9>(1, 6, 8, 1, lazy);
  }
  printf("%s: %d failing comparisons, %lld TMA loads and %lld CTA barriers emulated\n", failures ? "FAILED" : "PASSED", failures, emu::n_tma,
         emu::n_barriers);
  return failures ? 1 : 0;
}
