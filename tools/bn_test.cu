// Phase timing of the self-finalizing BatchNorm apply kernel on one layer shape (tools, not product).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hippie_b200/csrc/kernels.cuh"
using namespace hp;
static inline int ck_(cudaError_t e, int l) {
  if (e != cudaSuccess) {
    printf("CUDA error %s line %d\n", cudaGetErrorString(e), l);
    exit(1);
  }
  return 0;
}
#define CK(x) ck_((x), __LINE__)

static void run(int B, int L, int C, int res) {
  const int64_t rows = (int64_t)B * (L + 2) + 2, n = rows * C;
  float *c, *r, *out, *part, *coef, *gamma, *beta, *rm, *rv;
  uint16_t* planes;
  int64_t* cnt;
  CK(cudaMalloc(&c, n * 4)), CK(cudaMalloc(&r, n * 4)), CK(cudaMalloc(&out, n * 4)), CK(cudaMalloc(&planes, n * 4));
  const int tile = (128 / L) * L, M = B * L, ntiles = (M + tile - 1) / tile;
  CK(cudaMalloc(&part, (size_t)ntiles * C * 8)), CK(cudaMalloc(&coef, C * 32)), CK(cudaMalloc(&gamma, C * 4));
  CK(cudaMalloc(&beta, C * 4)), CK(cudaMalloc(&rm, C * 4)), CK(cudaMalloc(&rv, C * 4)), CK(cudaMalloc(&cnt, 8));
  CK(cudaMemset(c, 0, n * 4)), CK(cudaMemset(r, 0, n * 4)), CK(cudaMemset(part, 0, (size_t)ntiles * C * 8));
  std::vector<float> ones(C, 1.f);
  CK(cudaMemcpy(gamma, ones.data(), C * 4, cudaMemcpyHostToDevice)), CK(cudaMemset(beta, 0, C * 4));
  CK(cudaMemset(rm, 0, C * 4)), CK(cudaMemcpy(rv, ones.data(), C * 4, cudaMemcpyHostToDevice)), CK(cudaMemset(cnt, 0, 8));
  unsigned long long* ds;
  CK(cudaMalloc(&ds, 64)), CK(cudaMemset(ds, 0, 64));
  BnApply a{};
  a.c = c + C, a.coef = coef, a.r = res ? r + C : nullptr, a.rcoef = nullptr, a.out = out + C, a.out_up = nullptr;
  a.B = B, a.L = L, a.C = C, a.slope = 0.01f, a.train = 1;
  a.fin.part = part, a.fin.set_shape(ntiles, tile, M), a.fin.C = C, a.fin.gamma = gamma, a.fin.beta = beta;
  a.fin.run_mean = rm, a.fin.run_var = rv, a.fin.run_count = cnt, a.fin.coef = coef;
  a.out_p = planes + C, a.out_ps = n;
  for (int i = 0; i < 3; ++i) launch_bn_apply(a, 148, 0);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 50; ++i) launch_bn_apply(a, 148, 0);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  a.stamps = ds;
  launch_bn_apply(a, 148, 0);
  CK(cudaDeviceSynchronize());
  unsigned long long hs[8];
  CK(cudaMemcpy(hs, ds, 64, cudaMemcpyDeviceToHost));
  printf("bn_apply B=%d L=%2d C=%3d res%d ntiles=%3d: %6.2f us/launch back-to-back | CTA(0,0): wait %llu ns, finalize %llu ns, apply %llu ns | %.1f MB moved\n",
         B, L, C, res, ntiles, ms * 1e3 / 50, hs[1] - hs[0], hs[2] - hs[1], hs[3] - hs[2], (double)M * C * (4 + 4 + 4 + (res ? 4 : 0)) / 1e6);
  cudaFree(c), cudaFree(r), cudaFree(out), cudaFree(planes), cudaFree(part), cudaFree(coef);
}
static void run_bwd(int B, int L, int C, int shortcut) {
  const int64_t rows = (int64_t)B * (L + 2) + 2, n = rows * C;
  float *g, *out, *c, *cs, *coef, *coef_s, *gamma, *dg, *db, *part, *gres, *slot;
  uint16_t* planes;
  CK(cudaMalloc(&g, n * 4)), CK(cudaMalloc(&out, n * 4)), CK(cudaMalloc(&c, n * 4)), CK(cudaMalloc(&cs, n * 4));
  CK(cudaMalloc(&gres, n * 4)), CK(cudaMalloc(&planes, n * 8)), CK(cudaMalloc(&coef, C * 32)), CK(cudaMalloc(&coef_s, C * 32));
  CK(cudaMalloc(&gamma, C * 4)), CK(cudaMalloc(&dg, C * 4)), CK(cudaMalloc(&db, C * 4)), CK(cudaMalloc(&slot, 64));
  CK(cudaMalloc(&part, (size_t)(kBnBwdMaxChunks + 1) * C * 3 * 4));
  CK(cudaMemset(g, 0, n * 4)), CK(cudaMemset(out, 0, n * 4)), CK(cudaMemset(c, 0, n * 4)), CK(cudaMemset(cs, 0, n * 4));
  CK(cudaMemset(coef, 0, C * 32)), CK(cudaMemset(coef_s, 0, C * 32)), CK(cudaMemset(gamma, 0, C * 4)), CK(cudaMemset(slot, 0, 64));
  BnBwd a{};
  a.g = g + C, a.out = out + C, a.c = c + C, a.coef = coef, a.part = part, a.B = B, a.L = L, a.C = C, a.slope = 0.01f;
  a.gamma = gamma, a.dgamma = dg, a.dbeta = db, a.dil = 1, a.Ld = L, a.inv_n = 1.0 / ((double)B * L);
  a.dc_p = planes + C, a.dc_ps = n, a.dc_slot = slot;
  if (shortcut) {
    a.cs = cs + C, a.coef_s = coef_s, a.gamma_s = gamma, a.dgamma_s = dg, a.dbeta_s = db;
    a.dcs_p = planes + 2 * n + C, a.dcs_ps = n, a.dcs_slot = slot + 8, a.dil_s = 1, a.Ld_s = L;
  } else {
    a.gres = gres + C;
  }
  for (int i = 0; i < 3; ++i) launch_bn_bwd(a, 148, 0);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 50; ++i) launch_bn_bwd(a, 148, 0);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("bn_bwd (reduce + apply) B=%d L=%2d C=%3d shortcut%d: %6.2f us per pair back-to-back | %.1f MB moved\n", B, L, C, shortcut,
         ms * 1e3 / 50, (double)B * L * C * 4 * (2 * (3 + shortcut) + 1 + shortcut + (shortcut ? 0 : 1)) / 1e6);
}

int main() {
  run(512, 50, 64, 0);
  run(512, 50, 64, 1);
  run(512, 25, 128, 0);
  run(512, 13, 256, 1);
  run(512, 7, 512, 0);
  run(512, 4, 512, 1);
  run(512, 32, 64, 1);
  run(64, 50, 64, 0);
  run_bwd(512, 50, 64, 0);
  run_bwd(512, 25, 128, 0);
  run_bwd(512, 13, 256, 1);
  run_bwd(512, 7, 512, 0);
  run_bwd(512, 4, 512, 1);
  run_bwd(512, 32, 64, 0);
  run_bwd(64, 50, 64, 0);
  return 0;
}
