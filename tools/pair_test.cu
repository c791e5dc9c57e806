// Stand-alone check + micro-benchmark of the pair-plane tcgen05 kernels (conv_pair.cu) against the FP32 CUDA-core kernels.
//   make -C tools pair_test && ./tools/pair_test [B]          (on a B200)
// Prints, per shape, max |pair - simt| / max|simt| and the event-timed TFLOP/s of both kernels.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../hippie_b200/csrc/kernels.cuh"

using namespace hp;
enum { F16 = 0, BF16 = 1 };

static inline int ck_(cudaError_t e, const char* f, int l) {
  if (e != cudaSuccess) {
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), f, l);
    exit(1);
  }
  return 0;
}
#define CK(x) ck_((x), __FILE__, __LINE__)

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
static float* upload(const std::vector<float>& h) {
  float* d;
  CK(cudaMalloc(&d, h.size() * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  return d;
}
static void* planes_of(const float* d, int64_t n, float scale, int fmt) {
  void* p;
  CK(cudaMalloc(&p, n * 4));
  launch_to_pair(d, p, n, n, scale, fmt, 0);
  CK(cudaDeviceSynchronize());
  return p;
}

// forward conv: A = activations (a_fmt pair), B = K-major weight planes (b_fmt pair, scaled 2^8 when fp16)
static void test_conv(int B, int L, int Cin, int Cout, int k, int stride, bool bias, int a_fmt, int b_fmt) {
  const int Lout = (L + 2 * (k == 3 ? 1 : 0) - k) / stride + 1;
  const int64_t in_floats = ((int64_t)B * (L + 2) + 2) * Cin, out_floats = ((int64_t)B * (Lout + 2) + 2) * Cout;
  std::vector<float> hx(in_floats), hw((size_t)Cout * k * Cin), hb(Cout), hc0(out_floats, 0.f);
  fill(hx, 1, 1.0f), fill(hw, 2, 0.05f), fill(hb, 3, 0.5f);
  float *dx = upload(hx), *dw = upload(hw), *db = upload(hb), *dc1 = upload(hc0), *dc2 = upload(hc0), *dp1, *dp2;
  const int64_t part_floats = ((int64_t)B * Lout / 64 + 2) * Cout * 2;
  CK(cudaMalloc(&dp1, part_floats * 4)), CK(cudaMalloc(&dp2, part_floats * 4));
  CK(cudaMemset(dp1, 0, part_floats * 4)), CK(cudaMemset(dp2, 0, part_floats * 4));
  const float wscale = b_fmt == F16 ? 256.f : 1.f;
  void* px = planes_of(dx, in_floats, 1.f, a_fmt);
  void* pw = planes_of(dw, (int64_t)hw.size(), wscale, b_fmt);
  ConvGemm g{};
  g.A = dx + Cin, g.W = dw, g.bias = bias ? db : nullptr;
  g.M = B * Lout, g.N = Cout, g.K = k * Cin, g.Lout = Lout;
  g.in_rows = L + 2, g.in_stride = stride, g.in_off = k == 3 ? 0 : 1, g.in_C = Cin;
  g.out_rows = Lout + 2, g.out_off = 1, g.out_lstride = 1, g.accumulate = 0;
  ConvGemm g1 = g, g2 = g;
  g1.C = dc1 + Cout, g1.part = dp1, g2.C = dc2 + Cout, g2.part = dp2;
  const int tile1 = launch_conv_gemm_simt(g1, 0);
  TcMap ma, mw;
  const int bn = pair_pick_bn(B, Cout, Lout, 148);
  bool ok = pair_make_act_map(&ma, (const uint16_t*)px + Cin, in_floats, a_fmt, Cin, g.K, Lout, g.in_rows, stride, g.in_off, B) &&
            pair_make_w_map(&mw, pw, (int64_t)hw.size(), b_fmt, Cout, g.K, bn);
  if (!ok) {
    printf("conv: tensor map creation FAILED\n");
    return;
  }
  PairOpts o{1.f / wscale, a_fmt, b_fmt, 0, k};
  TcMap mc;
  const bool tma_out = !getenv("PT_TMA") || atoi(getenv("PT_TMA"));
  if (tma_out) {
    if (!pair_make_out_map(&mc, g2.C, Cout, Lout, g.out_rows, g.out_off, B)) {
      printf("conv: output tensor map creation FAILED\n");
      return;
    }
    o.out_map = &mc;
  }
  const int tile2 = launch_conv_pair(g2, ma, mw, bn, B, o, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> r1(out_floats), r2(out_floats);
  CK(cudaMemcpy(r1.data(), dc1, out_floats * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r2.data(), dc2, out_floats * 4, cudaMemcpyDeviceToHost));
  double mr, md = compare(r2, r1, &mr);
  // BatchNorm statistics: total sum and total M2 per column from both partial layouts
  double stat_err = 0;
  {
    const int nt1 = (g.M + tile1 - 1) / tile1, nt2 = (g.M + tile2 - 1) / tile2;
    std::vector<float> p1((size_t)nt1 * Cout * 2), p2((size_t)nt2 * Cout * 2);
    CK(cudaMemcpy(p1.data(), dp1, p1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(p2.data(), dp2, p2.size() * 4, cudaMemcpyDeviceToHost));
    for (int c = 0; c < Cout; ++c) {
      double s1 = 0, s2 = 0;
      for (int t = 0; t < nt1; ++t) s1 += p1[((size_t)t * Cout + c) * 2];
      for (int t = 0; t < nt2; ++t) s2 += p2[((size_t)t * Cout + c) * 2];
      auto m2tot = [&](std::vector<float>& p, int nt, int tile, double s) {
        double mean = s / g.M, m2 = 0;
        for (int t = 0; t < nt; ++t) {
          const int n = std::min(tile, g.M - t * tile);
          const double mt = p[((size_t)t * Cout + c) * 2] / n;
          m2 += p[((size_t)t * Cout + c) * 2 + 1] + n * (mt - mean) * (mt - mean);
        }
        return m2;
      };
      const double v1 = m2tot(p1, nt1, tile1, s1), v2 = m2tot(p2, nt2, tile2, s2);
      stat_err = fmax(stat_err, fabs(s1 - s2) / (fabs(s1) + 1e-3 * g.M));
      stat_err = fmax(stat_err, fabs(v1 - v2) / (fabs(v1) + 1e-6));
    }
  }
  const float t1 = time_ms([&] { launch_conv_gemm_simt(g1, 0); }, 20);
  const float t2 = time_ms([&] { launch_conv_pair(g2, ma, mw, bn, B, o, 0); }, 20);
  if (getenv("PT_STAMPS")) {  // phase stamps of CTA (0, 0) of one isolated launch
    unsigned long long* ds;
    CK(cudaMalloc(&ds, 64));
    CK(cudaMemset(ds, 0, 64));
    PairOpts o2 = o;
    o2.stamps = ds;
    CK(cudaDeviceSynchronize());
    launch_conv_pair(g2, ma, mw, bn, B, o2, 0);
    CK(cudaDeviceSynchronize());
    unsigned long long hs[8];
    CK(cudaMemcpy(hs, ds, 64, cudaMemcpyDeviceToHost));
    printf("   stamps ns: prologue %llu | first tile landed %llu | main loop %llu | tmem->smem %llu | stores %llu | stats %llu\n",
           hs[1] - hs[0], hs[2] - hs[1], hs[3] - hs[2], hs[4] - hs[3], hs[5] - hs[4], hs[6] - hs[5]);
    cudaFree(ds);
  }
  if (getenv("PT_STAMPS") && atoi(getenv("PT_STAMPS")) == 2) {  // every CTA: placement on the SMs and phase times
    const int nb = 128 / g.Lout, gx = (B + nb - 1) / nb, gy = g.N / bn, nc = gx * gy;
    unsigned long long* ds;
    CK(cudaMalloc(&ds, (size_t)nc * 64));
    CK(cudaMemset(ds, 0, (size_t)nc * 64));
    PairOpts o2 = o;
    o2.stamps = ds, o2.stamps_all = 1;
    CK(cudaDeviceSynchronize());
    launch_conv_pair(g2, ma, mw, bn, B, o2, 0);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> hs((size_t)nc * 8);
    CK(cudaMemcpy(hs.data(), ds, (size_t)nc * 64, cudaMemcpyDeviceToHost));
    int per_sm[256] = {0};
    unsigned long long t_min = ~0ull, t_max = 0;
    for (int c = 0; c < nc; ++c) {
      per_sm[hs[c * 8 + 7] & 255]++;
      t_min = std::min(t_min, hs[c * 8 + 0]), t_max = std::max(t_max, hs[c * 8 + 6]);
    }
    int sms = 0, hist[8] = {0};
    for (int i = 0; i < 256; ++i)
      if (per_sm[i]) ++sms, hist[std::min(per_sm[i], 7)]++;
    double ml[8] = {0}, tot[8] = {0}, fill[8] = {0}, late[8] = {0};
    int cnt[8] = {0};
    for (int c = 0; c < nc; ++c) {
      const int k = std::min(per_sm[hs[c * 8 + 7] & 255], 7);
      ml[k] += (double)(hs[c * 8 + 3] - hs[c * 8 + 2]), tot[k] += (double)(hs[c * 8 + 6] - hs[c * 8 + 0]);
      fill[k] += (double)(hs[c * 8 + 2] - hs[c * 8 + 1]), late[k] += (double)(hs[c * 8 + 0] - t_min), cnt[k]++;
    }
    printf("   all CTAs: %d CTAs on %d SMs (SMs with 1/2/3/4 CTAs: %d/%d/%d/%d), span %llu ns\n", nc, sms, hist[1], hist[2], hist[3],
           hist[4], t_max - t_min);
    for (int k = 1; k < 8; ++k)
      if (cnt[k])
        printf("     CTAs on SMs holding %d: n=%d  start +%.0f ns | first fill %.0f | main loop %.0f | CTA life %.0f ns\n", k, cnt[k],
               late[k] / cnt[k], fill[k] / cnt[k], ml[k] / cnt[k], tot[k] / cnt[k]);
    cudaFree(ds);
  }
  const double fl = 2.0 * g.M * g.N * g.K;
  printf("conv  B=%4d L=%3d %3d->%3d k%d s%d bias%d a%d b%d bn%3d | rel err %.2e stats %.1e | simt %7.1f us %6.1f TF | pair %7.1f us %6.1f TF\n",
         B, L, Cin, Cout, k, stride, bias, a_fmt, b_fmt, bn, md / mr, stat_err, t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9);
  cudaFree(dx), cudaFree(dw), cudaFree(db), cudaFree(dc1), cudaFree(dc2), cudaFree(dp1), cudaFree(dp2), cudaFree(px), cudaFree(pw);
}

// dgrad: gx = conv(dy, flipped/transposed w).  SIMT reads a transposed fp32 copy; the pair kernel reads the forward
// weight planes MN-major.
static void test_dgrad(int B, int L, int Cin, int Cout, int k, int a_fmt, int b_fmt, bool acc) {
  const int64_t dy_floats = ((int64_t)B * (L + 2) + 2) * Cout, gx_floats = ((int64_t)B * (L + 2) + 2) * Cin;
  std::vector<float> hdy(dy_floats, 0.f), hw((size_t)Cout * k * Cin), hwt((size_t)Cin * k * Cout), hg0(gx_floats, 0.f);
  fill(hw, 2, 0.05f);
  {  // realistic: zero pad rows
    std::vector<float> tmp((size_t)B * L * Cout);
    fill(tmp, 9, 1e-4f);
    for (int b = 0; b < B; ++b)
      for (int l = 0; l < L; ++l)
        for (int c = 0; c < Cout; ++c) hdy[(((size_t)b * (L + 2) + 1 + l) + 1) * Cout + c] = tmp[((size_t)b * L + l) * Cout + c];
  }
  if (acc) {
    std::vector<float> tmp((size_t)B * L * Cin);
    fill(tmp, 11, 1e-3f);
    for (int b = 0; b < B; ++b)
      for (int l = 0; l < L; ++l)
        for (int c = 0; c < Cin; ++c) hg0[(((size_t)b * (L + 2) + 1 + l) + 1) * Cin + c] = tmp[((size_t)b * L + l) * Cin + c];
  }
  for (int co = 0; co < Cout; ++co)
    for (int t = 0; t < k; ++t)
      for (int ci = 0; ci < Cin; ++ci) hwt[((size_t)ci * k + (k - 1 - t)) * Cout + co] = hw[((size_t)co * k + t) * Cin + ci];
  float *ddy = upload(hdy), *dw = upload(hw), *dwt = upload(hwt), *dg1 = upload(hg0), *dg2 = upload(hg0);
  const float wscale = b_fmt == F16 ? 256.f : 1.f;
  void* pdy = planes_of(ddy, dy_floats, 1.f, a_fmt);
  void* pw = planes_of(dw, (int64_t)hw.size(), wscale, b_fmt);
  ConvGemm g{};
  g.A = ddy + Cout, g.W = dwt, g.bias = nullptr, g.part = nullptr;
  g.M = B * L, g.N = Cin, g.K = k * Cout, g.Lout = L;
  g.in_rows = L + 2, g.in_stride = 1, g.in_off = k == 3 ? 0 : 1, g.in_C = Cout;
  g.out_rows = L + 2, g.out_off = 1, g.out_lstride = 1, g.accumulate = acc ? 1 : 0;
  ConvGemm g1 = g, g2 = g;
  g1.C = dg1 + Cin, g2.C = dg2 + Cin;
  launch_conv_gemm_simt(g1, 0);
  TcMap ma, mw;
  const int bn = pair_pick_bn(B, Cin, L, 148);
  bool ok = pair_make_act_map(&ma, (const uint16_t*)pdy + Cout, dy_floats, a_fmt, Cout, g.K, L, g.in_rows, 1, g.in_off, B) &&
            pair_make_wmn_map(&mw, pw, (int64_t)hw.size(), b_fmt, Cout, Cin, k);
  if (!ok) {
    printf("dgrad: tensor map creation FAILED\n");
    return;
  }
  PairOpts o{1.f / wscale, a_fmt, b_fmt, 1, k};
  TcMap mc;
  if (!getenv("PT_TMA") || atoi(getenv("PT_TMA"))) {
    if (!pair_make_out_map(&mc, g2.C, Cin, L, g.out_rows, g.out_off, B)) {
      printf("dgrad: output tensor map creation FAILED\n");
      return;
    }
    o.out_map = &mc;
  }
  launch_conv_pair(g2, ma, mw, bn, B, o, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> r1(gx_floats), r2(gx_floats);
  CK(cudaMemcpy(r1.data(), dg1, gx_floats * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r2.data(), dg2, gx_floats * 4, cudaMemcpyDeviceToHost));
  double mr, md = compare(r2, r1, &mr);
  g1.accumulate = g2.accumulate = 0;
  const float t1 = time_ms([&] { launch_conv_gemm_simt(g1, 0); }, 20);
  const float t2 = time_ms([&] { launch_conv_pair(g2, ma, mw, bn, B, o, 0); }, 20);
  const double fl = 2.0 * g.M * g.N * g.K;
  printf("dgrad B=%4d L=%3d %3d->%3d k%d acc%d a%d b%d bn%3d | rel err %.2e | simt %7.1f us %6.1f TF | pair %7.1f us %6.1f TF\n", B, L,
         Cin, Cout, k, acc, a_fmt, b_fmt, bn, md / mr, t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9);
  cudaFree(ddy), cudaFree(dw), cudaFree(dwt), cudaFree(dg1), cudaFree(dg2), cudaFree(pdy), cudaFree(pw);
}

static void test_wgrad(int B, int L, int Cin, int Cout, int k, int a_fmt, int b_fmt) {
  const int R = B * (L + 2), N = k * Cin, roff = k == 3 ? -1 : 0;
  std::vector<float> hdy((size_t)R * Cout), hx(((size_t)R + 2) * Cin);
  fill(hdy, 5, 1e-4f), fill(hx, 6, 1.0f);
  float *ddy = upload(hdy), *dx = upload(hx), *dw1, *dw2;
  CK(cudaMalloc(&dw1, (size_t)Cout * N * 4)), CK(cudaMalloc(&dw2, (size_t)Cout * N * 4));
  CK(cudaMemset(dw1, 0, (size_t)Cout * N * 4)), CK(cudaMemset(dw2, 0, (size_t)Cout * N * 4));
  void* pdy = planes_of(ddy, (int64_t)hdy.size(), 1.f, a_fmt);
  void* px = planes_of(dx, (int64_t)hx.size(), 1.f, b_fmt);
  WgradGemm g{};
  g.dY = ddy, g.X = dx + Cin, g.M = Cout, g.N = N, g.R = R, g.Cin = Cin, g.roff = roff;
  WgradGemm g1 = g, g2 = g;
  g1.dW = dw1, g2.dW = dw2;
  launch_wgrad_simt(g1, 148, 0);
  const int bn = (N % 128 == 0) ? 128 : 64;
  TcMap my, mx;
  bool ok = pair_make_rows_map(&my, pdy, (int64_t)hdy.size(), a_fmt, Cout, Cout, R) &&
            pair_make_rows_map(&mx, (const uint16_t*)px + Cin + (int64_t)roff * Cin, (int64_t)hx.size(), b_fmt, Cin, N, R);
  if (!ok) {
    printf("wgrad: tensor map creation FAILED\n");
    return;
  }
  TcMap mdw;
  if (!pair_make_dw_map(&mdw, dw2, Cout, N)) {
    printf("wgrad: output tensor map creation FAILED\n");
    return;
  }
  PairOpts o{1.f, a_fmt, b_fmt, 1, k};
  launch_wgrad_pair(g2, my, mx, mdw, bn, 148, o, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> r1((size_t)Cout * N), r2((size_t)Cout * N);
  CK(cudaMemcpy(r1.data(), dw1, r1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r2.data(), dw2, r2.size() * 4, cudaMemcpyDeviceToHost));
  double mr, md = compare(r2, r1, &mr);
  const float t1 = time_ms([&] { launch_wgrad_simt(g1, 148, 0); }, 20);
  const float t2 = time_ms([&] { launch_wgrad_pair(g2, my, mx, mdw, bn, 148, o, 0); }, 20);
  const double fl = 2.0 * Cout * (double)N * R;
  printf("wgrad B=%4d L=%3d %3d->%3d k%d a%d b%d bn%3d | rel err %.2e | simt %7.1f us %6.1f TF | pair %7.1f us %6.1f TF\n", B, L, Cin,
         Cout, k, a_fmt, b_fmt, bn, md / mr, t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9);
  cudaFree(ddy), cudaFree(dx), cudaFree(dw1), cudaFree(dw2), cudaFree(pdy), cudaFree(px);
}

// Two independent conv problems, alone and concurrently on two streams: do they share a chip-wide bottleneck?
struct ConvProb {
  ConvGemm g;
  TcMap ma, mw;
  int bn, B;
  PairOpts o;
};
static ConvProb make_prob(int B, int L, int Cin, int Cout, unsigned seed) {
  const int k = 3, Lout = L;
  const int64_t in_floats = ((int64_t)B * (L + 2) + 2) * Cin, out_floats = ((int64_t)B * (Lout + 2) + 2) * Cout;
  std::vector<float> hx(in_floats), hw((size_t)Cout * k * Cin);
  fill(hx, seed, 1.0f), fill(hw, seed + 1, 0.05f);
  float *dx = upload(hx), *dw = upload(hw), *dc;
  CK(cudaMalloc(&dc, out_floats * 4));
  void* px = planes_of(dx, in_floats, 1.f, F16);
  void* pw = planes_of(dw, (int64_t)hw.size(), 256.f, F16);
  ConvProb pr{};
  pr.g.A = dx + Cin, pr.g.W = dw, pr.g.C = dc + Cout, pr.g.M = B * Lout, pr.g.N = Cout, pr.g.K = k * Cin, pr.g.Lout = Lout;
  pr.g.in_rows = L + 2, pr.g.in_stride = 1, pr.g.in_off = 0, pr.g.in_C = Cin;
  pr.g.out_rows = Lout + 2, pr.g.out_off = 1, pr.g.out_lstride = 1;
  pr.bn = pair_pick_bn(B, Cout, Lout, 148), pr.B = B;
  pair_make_act_map(&pr.ma, (const uint16_t*)px + Cin, in_floats, F16, Cin, pr.g.K, Lout, L + 2, 1, 0, B);
  pair_make_w_map(&pr.mw, pw, (int64_t)hw.size(), F16, Cout, pr.g.K, pr.bn);
  pr.o = PairOpts{1.f / 256.f, F16, F16, 0, k};
  return pr;
}
static void test_concurrency(int B, int L, int C) {
  ConvProb a = make_prob(B, L, C, C, 21), b = make_prob(B, L, C, C, 31);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking), cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  const int iters = 50;
  auto run = [&](bool ua, bool ub) {
    cudaEvent_t e0, e1, j;
    cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventCreate(&j);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1);
    cudaStreamWaitEvent(s2, e0, 0);
    for (int i = 0; i < iters; ++i) {
      if (ua) launch_conv_pair(a.g, a.ma, a.mw, a.bn, a.B, a.o, s1);
      if (ub) launch_conv_pair(b.g, b.ma, b.mw, b.bn, b.B, b.o, s2);
    }
    cudaEventRecord(j, s2);
    cudaStreamWaitEvent(s1, j, 0);
    cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / iters;
  };
  run(true, true);
  const float ta = run(true, false), tb = run(false, true), tab = run(true, true);
  const int nb = 128 / L, ctas = ((B + nb - 1) / nb) * (C / 64);
  printf("concurrency B=%4d L=%2d C=%3d (%3d CTAs each): A alone %6.1f us, B alone %6.1f us, A||B %6.1f us per pair  (sum %6.1f)\n", B, L,
         C, ctas, ta, tb, tab, ta + tb);
}

int main(int argc, char** argv) {
  std::string err;
  if (!pair_init(&err)) {
    printf("pair_init failed: %s\n", err.c_str());
    return 1;
  }
  const int B = argc > 1 ? atoi(argv[1]) : 512;
  const int sel = argc > 2 ? atoi(argv[2]) : 0;
  if (sel == 4) {
    test_concurrency(128, 4, 512);   // 32 CTAs each
    test_concurrency(256, 4, 512);   // 64 CTAs each: together 128 <= 148 SMs
    test_concurrency(512, 4, 512);   // 128 CTAs each
    test_concurrency(1024, 4, 512);  // 256 CTAs each
    test_concurrency(512, 32, 64);   // K = 192
    return 0;
  }
  if (sel == 0 || sel == 1) {
    test_conv(8, 4, 64, 64, 3, 1, false, F16, F16);
    test_conv(B, 25, 64, 64, 3, 1, false, F16, F16);
    test_conv(B, 50, 64, 64, 3, 1, false, F16, F16);
    test_conv(B, 25, 64, 128, 3, 2, false, F16, F16);
    test_conv(B, 25, 64, 128, 1, 2, false, F16, F16);
    test_conv(B, 13, 128, 128, 3, 1, false, F16, F16);
    test_conv(B, 13, 128, 256, 3, 2, false, F16, F16);
    test_conv(B, 7, 256, 256, 3, 1, false, F16, F16);
    test_conv(B, 7, 256, 512, 1, 2, false, F16, F16);
    test_conv(B, 4, 512, 512, 3, 1, false, F16, F16);
    test_conv(B, 7, 512, 512, 3, 1, false, F16, F16);
    test_conv(B, 8, 512, 256, 3, 1, true, F16, F16);
    test_conv(B, 16, 256, 128, 3, 1, true, F16, F16);
    test_conv(B, 32, 128, 64, 3, 1, true, F16, F16);
    test_conv(B, 32, 64, 64, 3, 1, false, F16, F16);
    test_conv(B - 3, 7, 256, 256, 3, 1, false, F16, F16);
    test_conv(B, 7, 256, 256, 3, 1, false, BF16, BF16);
  }
  if (sel == 0 || sel == 2) {
    test_dgrad(8, 4, 64, 64, 3, F16, F16, false);
    test_dgrad(B, 4, 512, 512, 3, F16, F16, false);
    test_dgrad(B, 7, 512, 512, 3, F16, F16, true);
    test_dgrad(B, 8, 512, 256, 3, F16, F16, false);
    test_dgrad(B, 14, 128, 256, 1, F16, F16, false);
    test_dgrad(B, 25, 64, 64, 3, F16, F16, false);
    test_dgrad(B, 50, 64, 64, 3, F16, F16, true);
    test_dgrad(B, 4, 512, 512, 3, F16, F16, false);
  }
  if (sel == 9) test_conv(B, 7, 256, 256, 3, 1, false, BF16, F16);  // mixed operand formats: illegal instruction on sm_100a
  if (sel == 0 || sel == 3) {
    test_wgrad(8, 4, 64, 64, 3, F16, F16);
    test_wgrad(8, 4, 128, 128, 3, F16, F16);
    test_wgrad(B, 4, 512, 512, 3, F16, F16);
    test_wgrad(B, 7, 512, 512, 3, F16, F16);
    test_wgrad(B, 8, 256, 256, 3, F16, F16);
    test_wgrad(B, 14, 128, 256, 1, F16, F16);
    test_wgrad(B, 25, 64, 64, 3, F16, F16);
    test_wgrad(B, 50, 64, 64, 3, F16, F16);
    test_wgrad(B, 32, 128, 64, 3, F16, F16);
    test_wgrad(B, 4, 512, 512, 3, BF16, BF16);
  }
  return 0;
}
