// Stand-alone check + micro-benchmark of the tcgen05 kernels against the FP32 CUDA-core kernels.
//   make -C tools tc_test && ./tools/tc_test          (on a B200)
// Prints, per shape, max |tc - simt| / max|simt| and the event-timed TFLOP/s of both kernels.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../hippie_b200/csrc/kernels.cuh"

using namespace hp;

static inline int ck_(cudaError_t e, const char* f, int l) {
  if (e != cudaSuccess) {
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), f, l);
    exit(1);
  }
  return 0;
}
#define CK(x) ck_((x), __FILE__, __LINE__)

static int g_passes = 3;

static void fill(std::vector<float>& v, unsigned seed, float scale) {
  unsigned s = seed * 2654435761u + 12345u;
  for (auto& x : v) {
    s = s * 1664525u + 1013904223u;
    x = scale * (((s >> 8) & 0xFFFF) / 32768.0f - 1.0f);
  }
}

static double compare(const std::vector<float>& a, const std::vector<float>& b, double* maxref) {
  double md = 0, mr = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    md = fmax(md, fabs((double)a[i] - b[i]));
    mr = fmax(mr, fabs((double)b[i]));
  }
  *maxref = mr;
  return md;
}

template <class F>
static float time_ms(F f, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

static void test_conv(int B, int L, int Cin, int Cout, int k, int stride, bool bias, bool acc) {
  const int Lout = (L + 2 * (k == 3 ? 1 : 0) - k) / stride + 1;
  const int64_t in_floats = ((int64_t)B * (L + 2) + 2) * Cin, out_floats = ((int64_t)B * (Lout + 2) + 2) * Cout;
  std::vector<float> hx(in_floats), hw((size_t)Cout * k * Cin), hb(Cout), hc0(out_floats);
  fill(hx, 1, 1.0f), fill(hw, 2, 0.05f), fill(hb, 3, 0.5f), fill(hc0, 4, 0.3f);
  float *dx, *dw, *db, *dc1, *dc2, *dp1, *dp2;
  CK(cudaMalloc(&dx, in_floats * 4)), CK(cudaMalloc(&dw, hw.size() * 4)), CK(cudaMalloc(&db, Cout * 4));
  CK(cudaMalloc(&dc1, out_floats * 4)), CK(cudaMalloc(&dc2, out_floats * 4));
  const int64_t part_floats = ((int64_t)B * Lout / 64 + 2) * Cout * 2;
  CK(cudaMalloc(&dp1, part_floats * 4)), CK(cudaMalloc(&dp2, part_floats * 4));
  CK(cudaMemcpy(dx, hx.data(), in_floats * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dc1, hc0.data(), out_floats * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dc2, hc0.data(), out_floats * 4, cudaMemcpyHostToDevice));
  ConvGemm g{};
  g.A = dx + Cin, g.W = dw, g.bias = bias ? db : nullptr, g.M = B * Lout, g.N = Cout, g.K = k * Cin, g.Lout = Lout;
  g.in_rows = L + 2, g.in_stride = stride, g.in_off = k == 3 ? 0 : 1, g.in_C = Cin;
  g.out_rows = Lout + 2, g.out_off = 1, g.out_lstride = 1, g.accumulate = acc ? 1 : 0;
  ConvGemm g1 = g, g2 = g;
  g1.C = dc1 + Cout, g1.part = getenv("TC_NOSTATS") ? nullptr : dp1, g2.C = dc2 + Cout, g2.part = getenv("TC_NOSTATS") ? nullptr : dp2;
  launch_conv_gemm_simt(g1, 0);
  TcMap ma, mw;
  const int bn = tc_pick_bn(B, Cout, Lout, 148);
  bool ok = tc_make_act_map(&ma, g.A, Cin, g.K, Lout, g.in_rows, stride, g.in_off, B) && tc_make_weight_map(&mw, dw, Cout, g.K, bn);
  if (!ok) {
    printf("conv  B=%d L=%d %d->%d k%d s%d: tensor map creation FAILED\n", B, L, Cin, Cout, k, stride);
    return;
  }
  launch_conv_gemm_tc(g2, ma, mw, bn, B, g_passes, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> r1(out_floats), r2(out_floats);
  CK(cudaMemcpy(r1.data(), dc1, out_floats * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r2.data(), dc2, out_floats * 4, cudaMemcpyDeviceToHost));
  double mr, md = compare(r2, r1, &mr);
  g1.accumulate = g2.accumulate = 0;
  const float t1 = time_ms([&] { launch_conv_gemm_simt(g1, 0); }, 20);
  const float t2 = time_ms([&] { launch_conv_gemm_tc(g2, ma, mw, bn, B, g_passes, 0); }, 20);
  const double fl = 2.0 * g.M * g.N * g.K;
  printf("conv  B=%4d L=%3d %3d->%3d k%d s%d bias%d acc%d bn%3d | rel err %.2e | simt %7.1f us %6.1f TF | tc %7.1f us %6.1f TF\n", B,
         L, Cin, Cout, k, stride, bias, acc, bn, md / mr, t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9);
  cudaFree(dx), cudaFree(dw), cudaFree(db), cudaFree(dc1), cudaFree(dc2), cudaFree(dp1), cudaFree(dp2);
}

static void test_wgrad(int B, int L, int Cin, int Cout, int k) {
  const int R = B * (L + 2), N = k * Cin, roff = k == 3 ? -1 : 0;
  std::vector<float> hdy((size_t)R * Cout), hx(((size_t)R + 2) * Cin);
  fill(hdy, 5, 0.1f), fill(hx, 6, 1.0f);
  float *ddy, *dx, *dw1, *dw2;
  CK(cudaMalloc(&ddy, hdy.size() * 4)), CK(cudaMalloc(&dx, hx.size() * 4));
  CK(cudaMalloc(&dw1, (size_t)Cout * N * 4)), CK(cudaMalloc(&dw2, (size_t)Cout * N * 4));
  CK(cudaMemcpy(ddy, hdy.data(), hdy.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dw1, 0, (size_t)Cout * N * 4)), CK(cudaMemset(dw2, 0, (size_t)Cout * N * 4));
  WgradGemm g{};
  g.dY = ddy, g.X = dx + Cin, g.M = Cout, g.N = N, g.R = R, g.Cin = Cin, g.roff = roff;
  WgradGemm g1 = g, g2 = g;
  g1.dW = dw1, g2.dW = dw2;
  launch_wgrad_simt(g1, 148, 0);
  const int bn = (N % 128 == 0) ? 128 : 64;
  TcMap my, mx;
  bool ok = tc_make_rows_map(&my, g.dY, Cout, Cout, R, 4) && tc_make_rows_map(&mx, g.X + (int64_t)roff * Cin, Cin, N, R, bn / 32);
  if (!ok) {
    printf("wgrad B=%d L=%d %d->%d k%d: tensor map creation FAILED\n", B, L, Cin, Cout, k);
    return;
  }
  launch_wgrad_tc(g2, my, mx, bn, 148, g_passes, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> r1((size_t)Cout * N), r2((size_t)Cout * N);
  CK(cudaMemcpy(r1.data(), dw1, r1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r2.data(), dw2, r2.size() * 4, cudaMemcpyDeviceToHost));
  double mr, md = compare(r2, r1, &mr);
  const float t1 = time_ms([&] { launch_wgrad_simt(g1, 148, 0); }, 20);
  const float t2 = time_ms([&] { launch_wgrad_tc(g2, my, mx, bn, 148, g_passes, 0); }, 20);
  const double fl = 2.0 * Cout * (double)N * R;
  printf("wgrad B=%4d L=%3d %3d->%3d k%d bn%3d | rel err %.2e (tc[0..3] %.4f %.4f %.4f %.4f simt %.4f %.4f %.4f %.4f) | simt %7.1f us %6.1f TF | tc %7.1f us %6.1f TF\n",
         B, L, Cin, Cout, k, bn, md / mr, r2[0], r2[1], r2[2], r2[3], r1[0], r1[1], r1[2], r1[3], t1 * 1e3, fl / t1 / 1e9,
         t2 * 1e3, fl / t2 / 1e9);
  cudaFree(ddy), cudaFree(dx), cudaFree(dw1), cudaFree(dw2);
}

int main(int argc, char** argv) {
  std::string err;
  if (!tc_init(&err)) {
    printf("tc_init failed: %s\n", err.c_str());
    return 1;
  }
  const int B = argc > 1 ? atoi(argv[1]) : 512;
  if (getenv("TC_PASSES")) g_passes = atoi(getenv("TC_PASSES"));
  if (argc > 2) {  // probe the MN-major descriptor / TMA swizzle pairing
    const int swz[] = {3 /*128B*/, 4 /*128B_ATOM_32B*/, 5 /*ATOM_32B_FLIP_8B*/, 6 /*ATOM_64B*/};
    const unsigned lt[] = {2, 1};
    const unsigned offs[] = {4096, 1024, 512, 256, 128};
    for (int sw : swz)
      for (unsigned l : lt)
        for (unsigned lbo : offs)
          for (unsigned sbo : offs) {
            if (lbo == sbo) continue;
            tc_debug_wgrad_knobs(lbo, sbo, l, sw);
            printf("swz %d ltype %u lbo %4u sbo %4u : ", sw, l, lbo, sbo);
            test_wgrad(8, 4, 128, 128, 3);
          }
    return 0;
  }
  if (getenv("TC_DBG")) {  // where does the time go?  0 = full, 1 = no hi/lo split, 2 = one MMA pass, 3 = both
    for (int dbg : {0, 4, 1, 2, 3}) {
      tc_debug_conv(dbg);
      printf("dbg %d: ", dbg);
      test_conv(B, 7, 512, 512, 3, 1, false, false);
      printf("dbg %d: ", dbg);
      test_conv(B, 4, 512, 512, 3, 1, false, false);
      printf("dbg %d: ", dbg);
      test_conv(B, 16, 128, 128, 3, 1, false, false);
    }
    return 0;
  }
  test_wgrad(8, 4, 64, 64, 3);
  test_wgrad(8, 4, 128, 128, 3);
  test_wgrad(B, 4, 512, 512, 3);
  test_wgrad(B, 8, 256, 256, 3);
  test_wgrad(B, 32, 64, 64, 3);
  test_wgrad(B, 50, 64, 64, 3);
  test_wgrad(B, 25, 64, 128, 1);
  test_conv(8, 4, 64, 64, 3, 1, false, false);
  test_conv(B, 4, 512, 512, 3, 1, false, false);
  test_conv(B, 7, 512, 512, 3, 1, false, false);
  test_conv(B, 8, 512, 256, 3, 1, true, false);
  test_conv(B, 8, 256, 256, 3, 1, false, true);
  test_conv(B, 13, 256, 256, 3, 1, false, false);
  test_conv(B, 16, 128, 128, 3, 1, false, false);
  test_conv(B, 25, 128, 128, 3, 1, false, false);
  test_conv(B, 32, 64, 64, 3, 1, false, false);
  test_conv(B, 50, 64, 64, 3, 1, false, false);
  test_conv(B, 50, 64, 128, 3, 2, false, false);
  test_conv(B, 50, 64, 128, 1, 2, false, false);
  test_conv(B, 13, 256, 512, 3, 2, false, false);
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
