// HBM-bound kernels of the cVAE step: stem conv, BatchNorm finalize / apply / backward, pooling +
// linear tails, decoder head and tail (+ MSE), gradient clipping + AdamW.  All tensors are
// channels-last padded rows (kernels.cuh).  Reference anchors are given per kernel.
#include <cstdlib>

#include "kernels.cuh"
#include "pair_fmt.cuh"
#include "pdl.cuh"

// Launch bounds of the BatchNorm kernels (256 threads).  Capping them at 112 registers (__maxnreg__) so that one of their
// CTAs fits next to two GEMM CTAs of 96 registers was measured slower (bs512 step 3.00 vs 2.98 ms, gpurun_out/r02_exp33.txt).
#ifndef HP_BN_BOUNDS
#define HP_BN_BOUNDS __launch_bounds__(256)
#endif

namespace hp {

namespace {

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }
// power of two s with bound * s in [2^13, 2^14): the scale of a gradient pair tensor whose |values| <= bound
__device__ __forceinline__ float pair_scale_from_bound(float bound) {
  if (!(bound > 0.f) || bound > 3.0e38f) return 1.f;
  const int e = (int)((__float_as_uint(bound) >> 23) & 0xFFu) - 127;
  int se = 13 - e;
  se = se < -126 ? -126 : (se > 126 ? 126 : se);
  return __uint_as_float((uint32_t)(se + 127) << 23);
}
__device__ __forceinline__ float4 fmax4abs(float4 m, float4 v) {
  return make_float4(fmaxf(m.x, fabsf(v.x)), fmaxf(m.y, fabsf(v.y)), fmaxf(m.z, fabsf(v.z)), fmaxf(m.w, fabsf(v.w)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Stem: Conv1d(1, 64, k3, s2, p1, bias=False)   reference hippie/backbones.py:78,95
// 128 logical rows per CTA; thread = (channel, row group of 4).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float xin(const float* x, int Lin, int i) { return (i >= 0 && i < Lin) ? __ldg(x + i) : 0.f; }

__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       float* __restrict__ c0, float* __restrict__ part, int B,
                                                       int Lin, int Lout) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[4][64];
  __shared__ float smean[64];
  const int co = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int M = B * Lout, m0 = blockIdx.x * 128;
  const float w0 = w[co * 3 + 0], w1 = w[co * 3 + 1], w2 = w[co * 3 + 2];
  // the thread's 32 outputs stay in registers between the two passes of the centred statistics (the loads of a row are
  // warp-uniform broadcasts; unrolled so that they are all in flight at once: 20 -> ~8 us on the ISI branch)
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int m = m0 + rg + 4 * j;
    v[j] = 0.f;
    if (m < M) {
      const int b = m / Lout, l = m - b * Lout;
      const float* xb = x + (int64_t)b * Lin;
      float t = w0 * xin(xb, Lin, 2 * l - 1);
      t = fmaf(w1, xin(xb, Lin, 2 * l), t);
      t = fmaf(w2, xin(xb, Lin, 2 * l + 1), t);
      c0[((int64_t)b * (Lout + 2) + 1 + l) * 64 + co] = t;
      v[j] = t;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];  // same order as the sequential loop (invalid rows add 0)
  if (!part) return;
  red[rg][co] = s;
  __syncthreads();
  const int nvalid = min(128, M - m0);
  float colsum = red[0][co] + red[1][co] + red[2][co] + red[3][co];
  if (rg == 0) smean[co] = colsum / (float)nvalid;
  __syncthreads();
  const float mu = smean[co];
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (m0 + rg + 4 * j < M) {
      const float d = v[j] - mu;
      q += d * d;
    }
  }
  __syncthreads();
  red[rg][co] = q;
  __syncthreads();
  if (rg == 0) {
    float m2 = red[0][co] + red[1][co] + red[2][co] + red[3][co];
    part[((int64_t)blockIdx.x * 64 + co) * 2 + 0] = colsum;
    part[((int64_t)blockIdx.x * 64 + co) * 2 + 1] = m2;
  }
}

__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dc0,
                                                         float* __restrict__ part, int B, int Lin, int Lout) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[4][192];
  const int co = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int M = B * Lout, m0 = blockIdx.x * 128;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int r = rg; r < 128; r += 4) {
    int m = m0 + r;
    if (m >= M) break;
    int b = m / Lout, l = m - b * Lout;
    const float* xb = x + (int64_t)b * Lin;
    float g = dc0[((int64_t)b * (Lout + 2) + 1 + l) * 64 + co];
    a0 = fmaf(g, xin(xb, Lin, 2 * l - 1), a0);
    a1 = fmaf(g, xin(xb, Lin, 2 * l), a1);
    a2 = fmaf(g, xin(xb, Lin, 2 * l + 1), a2);
  }
  red[rg][co * 3 + 0] = a0, red[rg][co * 3 + 1] = a1, red[rg][co * 3 + 2] = a2;
  __syncthreads();
  if (threadIdx.x < 192) {
    int i = threadIdx.x;
    part[(int64_t)blockIdx.x * 192 + i] = red[0][i] + red[1][i] + red[2][i] + red[3][i];
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ out,
                                       int accumulate) {
  pdl_trigger();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)part[(int64_t)p * n + i];
  out[i] = accumulate ? out[i] + (float)s : (float)s;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d (nn.BatchNorm1d defaults: eps 1e-5, momentum 0.1, biased variance to normalise, unbiased into
// running_var).  The producing conv leaves per-tile (sum, centred M2) partials; every CTA of the apply kernel owns a
// slab of 32 channels and first turns the partials of ITS channels into (scale, shift) -- Chan's parallel combination
// in double, fixed order -- so there is no separate finalize launch.  CTAs with blockIdx.y == 0 also publish the
// coefficients (mean / invstd are needed by the backward pass) and update the running statistics.
// ------------------------------------------------------------------------------------------------
constexpr int kSlab = 32;  // channels per CTA: 128-byte row segments

struct ChanAcc {
  double s, q;
};

// 256 threads.  Results for the slab's channels land in shared arrays sc / sh / mu (scale, shift = beta, mean).
// Per tile t with n_t rows the conv left (s_t, m2_t) = (sum, centred sum of squares); with S = sum s_t,
// Q = sum (m2_t + s_t^2 / n_t):  mean = S / M,  M2 = Q - S^2 / M  (Chan's combination written as sums; double keeps
// the cancellation harmless: relative error 1e-16 * (1 + mean^2 / var)).  No division inside the loop.
__device__ __forceinline__ void bn_finalize_slab(const BnFinalize& f, int c0, bool publish, float* sc, float* sh,
                                                 float* mu, ChanAcc (*acc)[kSlab]) {
  const int c = threadIdx.x & (kSlab - 1), slice = threadIdx.x >> 5;  // 8 slices of tiles
  const double inv_full = f.inv_tile, inv_last = f.inv_last;
  const int last = f.ntiles - 1;
  double S = 0.0, Q = 0.0;
  const float* base = f.part + ((int64_t)c0 + c) * 2;
  // everything the last step needs is requested up front: the phase is a chain of L2 round trips otherwise
  float gam = 0.f, bet = 0.f, rmean = 0.f, rvar = 0.f;
  if (slice == 0) {
    gam = f.gamma[c0 + c], bet = f.beta[c0 + c];
    if (publish) rmean = f.run_mean[c0 + c], rvar = f.run_var[c0 + c];
  }
  // 8 tiles in flight per thread: the loop is a chain of L2 round trips otherwise (up to 32 tiles per thread)
  constexpr int FL = 8;
  if (f.tot) {  // the conv epilogues accumulated (sum x, sum x^2) per channel: nothing to re-sum
    if (slice == 0) {
      double v[kBnFwdTotCopies][2];
#pragma unroll
      for (int k = 0; k < kBnFwdTotCopies; ++k) {
        const double* p = f.tot + (int64_t)k * 2 * f.C + c0 + c;
        v[k][0] = __ldcg(p), v[k][1] = __ldcg(p + f.C);
      }
#pragma unroll
      for (int k = 0; k < kBnFwdTotCopies; ++k) S += v[k][0], Q += v[k][1];
    }
  } else
  for (int t0 = slice; t0 < f.ntiles; t0 += 8 * FL) {
    float2 pv[FL];
#pragma unroll
    for (int u = 0; u < FL; ++u) {
      const int t = t0 + 8 * u;
      pv[u] = t < f.ntiles ? __ldcg(reinterpret_cast<const float2*>(base + (int64_t)t * f.C * 2)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < FL; ++u) {
      const int t = t0 + 8 * u;
      const double st = (double)pv[u].x;
      S += st;
      Q += (double)pv[u].y + st * st * (t == last ? inv_last : inv_full);
    }
  }
  if (!f.tot) {
    acc[slice][c] = ChanAcc{S, Q};
    __syncthreads();
  }
  if (slice == 0) {
    if (!f.tot)
      for (int k = 1; k < 8; ++k) S += acc[k][c].s, Q += acc[k][c].q;
    const double mean = S * f.inv_M;
    double m2 = Q - S * mean;
    m2 = m2 > 0.0 ? m2 : 0.0;
    const double var_b = m2 * f.inv_M;
    // 1 / sqrt(v) in double without the software sqrt / division: float estimate + two Newton steps (multiplies only)
    const double vv = var_b + (double)kBnEps;
    double y = (double)rsqrtf((float)vv);
    y = y * (1.5 - 0.5 * vv * y * y);
    y = y * (1.5 - 0.5 * vv * y * y);
    const float invstd = (float)y;
    const float meanf = (float)mean;
    const int cc = c0 + c;
    const float scale = gam * invstd, beta = bet;
    sc[c] = scale, sh[c] = beta, mu[c] = meanf;
    if (publish) {
      f.coef[0 * f.C + cc] = scale;
      f.coef[1 * f.C + cc] = beta;
      f.coef[2 * f.C + cc] = meanf;
      f.coef[3 * f.C + cc] = invstd;
      const float var_u = (float)(m2 * f.inv_Mm1);
      f.run_mean[cc] = (1.f - kBnMomentum) * rmean + kBnMomentum * meanf;
      f.run_var[cc] = (1.f - kBnMomentum) * rvar + kBnMomentum * var_u;
      if (cc == 0) *f.run_count += 1;
    }
  }
  __syncthreads();
}

__global__ void bn_eval_coefs_kernel(const BnEvalEntry* __restrict__ tab, const float* __restrict__ params,
                                     const float* __restrict__ run_mean, const float* __restrict__ run_var,
                                     float* __restrict__ ws) {
  const BnEvalEntry e = tab[blockIdx.x];
  float* coef = ws + e.coef_off;
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    float invstd = 1.f / sqrtf(run_var[e.run_off + c] + kBnEps);
    coef[0 * e.C + c] = params[e.gamma_off + c] * invstd;
    coef[1 * e.C + c] = params[e.beta_off + c];
    coef[2 * e.C + c] = run_mean[e.run_off + c];
    coef[3 * e.C + c] = invstd;
  }
}

// out = lrelu((c - mean)*scale + beta + residual)   reference hippie/backbones.py:37-40,66-69,95
// grid = (C / 32, row chunks); thread = (channel quad of the slab, row lane)
template <int RES>  // 0 none, 1 identity, 2 BatchNorm'd shortcut
__global__ void HP_BN_BOUNDS bn_apply_kernel(BnApply a) {
  __shared__ float s_sc[kSlab], s_sh[kSlab], s_mu[kSlab], r_sc[kSlab], r_sh[kSlab], r_mu[kSlab];
  __shared__ ChanAcc s_acc[8][kSlab];
  pdl_trigger();
  const bool stamping = a.stamps && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  auto stamp = [&](int i) {
    if (stamping) {
      unsigned long long tt;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tt));
      a.stamps[i] = tt;
    }
  };
  stamp(0);
  pdl_wait();
  stamp(1);
  const int c0 = blockIdx.x * kSlab;
  if (a.train) {
    bn_finalize_slab(a.fin, c0, blockIdx.y == 0, s_sc, s_sh, s_mu, s_acc);
    if (RES == 2) bn_finalize_slab(a.rfin, c0, blockIdx.y == 0, r_sc, r_sh, r_mu, s_acc);
  } else {
    if (threadIdx.x < kSlab) {
      const int cc = c0 + threadIdx.x;
      s_sc[threadIdx.x] = a.coef[0 * a.C + cc], s_sh[threadIdx.x] = a.coef[1 * a.C + cc];
      s_mu[threadIdx.x] = a.coef[2 * a.C + cc];
      if (RES == 2) {
        r_sc[threadIdx.x] = a.rcoef[0 * a.C + cc], r_sh[threadIdx.x] = a.rcoef[1 * a.C + cc];
        r_mu[threadIdx.x] = a.rcoef[2 * a.C + cc];
      }
    }
    __syncthreads();
  }
  stamp(2);
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const float4 sc = *reinterpret_cast<const float4*>(s_sc + q * 4), be = *reinterpret_cast<const float4*>(s_sh + q * 4);
  const float4 mu = *reinterpret_cast<const float4*>(s_mu + q * 4);
  float4 rs = sc, rb = be, rm = mu;
  if (RES == 2) {
    rs = *reinterpret_cast<const float4*>(r_sc + q * 4), rb = *reinterpret_cast<const float4*>(r_sh + q * 4);
    rm = *reinterpret_cast<const float4*>(r_mu + q * 4);
  }
  const int M = a.B * a.L;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_end = min(M, (int)(blockIdx.y + 1) * rows_per);
  const int col = c0 + q * 4;
  constexpr int U = 4;  // independent rows per thread and iteration: the kernel is latency-bound, keep loads in flight
  for (int m0 = blockIdx.y * rows_per + rl; m0 < m_end; m0 += 32 * U) {
    float4 x[U], r[U];
    int64_t row[U], ru[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int m = m0 + 32 * u;
      if (m < m_end) {
        const int b = m / a.L, l = m - b * a.L;
        row[u] = (int64_t)b * (a.L + 2) + 1 + l, ru[u] = (int64_t)b * (2 * a.L + 2) + 1 + 2 * l;
        x[u] = *reinterpret_cast<const float4*>(a.c + row[u] * a.C + col);
        if (RES != 0) r[u] = *reinterpret_cast<const float4*>(a.r + row[u] * a.C + col);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (m0 + 32 * u >= m_end) break;
      float4 y;
      y.x = fmaf(x[u].x - mu.x, sc.x, be.x), y.y = fmaf(x[u].y - mu.y, sc.y, be.y);
      y.z = fmaf(x[u].z - mu.z, sc.z, be.z), y.w = fmaf(x[u].w - mu.w, sc.w, be.w);
      if (RES == 1) {
        y.x += r[u].x, y.y += r[u].y, y.z += r[u].z, y.w += r[u].w;
      } else if (RES == 2) {
        y.x += fmaf(r[u].x - rm.x, rs.x, rb.x), y.y += fmaf(r[u].y - rm.y, rs.y, rb.y);
        y.z += fmaf(r[u].z - rm.z, rs.z, rb.z), y.w += fmaf(r[u].w - rm.w, rs.w, rb.w);
      }
      y.x = lrelu(y.x, a.slope), y.y = lrelu(y.y, a.slope), y.z = lrelu(y.z, a.slope), y.w = lrelu(y.w, a.slope);
      *reinterpret_cast<float4*>(a.out + row[u] * a.C + col) = y;
      if (a.out_p) store_pair4(a.out_p, a.out_ps, row[u] * a.C + col, y, a.flags);
      if (a.out_up) {
        *reinterpret_cast<float4*>(a.out_up + ru[u] * a.C + col) = y;
        *reinterpret_cast<float4*>(a.out_up + (ru[u] + 1) * a.C + col) = y;
      }
      if (a.up_p) {
        store_pair4(a.up_p, a.up_ps, ru[u] * a.C + col, y);
        store_pair4(a.up_p, a.up_ps, (ru[u] + 1) * a.C + col, y);
      }
    }
  }
  stamp(3);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm backward (train mode), fused with the LeakyReLU backward and the residual split.
//   g_pre = g * (out > 0 ? 1 : slope);  S1 = sum g_pre, S2 = sum g_pre * xhat (per channel)
//   dc    = gamma * invstd * (g_pre - S1/n - xhat * S2/n)
// Two launches: the reduce kernel leaves per-chunk partials and the maxima that bound |dc|; every CTA of the apply
// kernel sums the partials of its 32-channel slab itself.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_g(const BnBwd& a, int b, int l, int col) {
  if (a.g_up) {
    const int64_t ru = (int64_t)b * (2 * a.L + 2) + 1 + 2 * l;
    const float4 g0 = *reinterpret_cast<const float4*>(a.g + ru * a.C + col);
    const float4 g1 = *reinterpret_cast<const float4*>(a.g + (ru + 1) * a.C + col);
    return make_float4(g0.x + g1.x, g0.y + g1.y, g0.z + g1.z, g0.w + g1.w);
  }
  const int64_t row = (int64_t)b * (a.L + 2) + 1 + l;
  return *reinterpret_cast<const float4*>(a.g + row * a.C + col);
}
__device__ __forceinline__ float max4(float4 v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
__device__ __forceinline__ float block_max(float v, float* red) {  // 256 threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) v = fmaxf(v, red[w]);
  __syncthreads();
  return v;
}

// part[chunk][C][3] = (S1, S2, S2s);  slot = (max|g_pre|, max|xhat|, max|gamma*invstd|, 1/scale) as atomicMax targets
// (non-negative floats order like their bit patterns; the slots are zeroed once per step by the engine)
template <int U>
__global__ void HP_BN_BOUNDS bn_bwd_reduce_kernel(BnBwd a, int rows_per_cta) {
  __shared__ float4 red[3][256];
  __shared__ float mred[8];
  pdl_trigger();
  pdl_wait();
  const int C4 = a.C >> 2;
  const int RL = 256 / C4;  // row lanes
  const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
  const int M = a.B * a.L;
  const int m_begin = blockIdx.x * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, s3 = s1, gm = s1, xm = s1, xsm = s1;
  float km = 0.f, ksm = 0.f;
  if (rl < RL) {
    const float4 mu = *reinterpret_cast<const float4*>(a.coef + 2 * a.C + cq * 4);
    const float4 is = *reinterpret_cast<const float4*>(a.coef + 3 * a.C + cq * 4);
    float4 mus = mu, iss = is;
    if (a.dc_slot) {
      const float4 ga = *reinterpret_cast<const float4*>(a.gamma + cq * 4);
      km = max4(fmax4abs(make_float4(0.f, 0.f, 0.f, 0.f), make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w)));
    }
    if (a.cs) {
      mus = *reinterpret_cast<const float4*>(a.coef_s + 2 * a.C + cq * 4);
      iss = *reinterpret_cast<const float4*>(a.coef_s + 3 * a.C + cq * 4);
      if (a.dcs_slot) {
        const float4 ga = *reinterpret_cast<const float4*>(a.gamma_s + cq * 4);
        ksm = max4(fmax4abs(make_float4(0.f, 0.f, 0.f, 0.f),
                            make_float4(ga.x * iss.x, ga.y * iss.y, ga.z * iss.z, ga.w * iss.w)));
      }
    }
    for (int m0 = m_begin + rl; m0 < m_end; m0 += RL * U) {
      float4 gv[U], ov[U], xv[U], xsv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int m = m0 + RL * u;
        if (m < m_end) {
          const int b = m / a.L, l = m - b * a.L;
          const int64_t row = (int64_t)b * (a.L + 2) + 1 + l;
          gv[u] = load_g(a, b, l, cq * 4);
          ov[u] = *reinterpret_cast<const float4*>(a.out + row * a.C + cq * 4);
          xv[u] = *reinterpret_cast<const float4*>(a.c + row * a.C + cq * 4);
          if (a.cs) xsv[u] = *reinterpret_cast<const float4*>(a.cs + row * a.C + cq * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (m0 + RL * u >= m_end) break;
        float4 g = gv[u];
        const float4 o = ov[u], x = xv[u];
        g.x *= o.x > 0.f ? 1.f : a.slope, g.y *= o.y > 0.f ? 1.f : a.slope;
        g.z *= o.z > 0.f ? 1.f : a.slope, g.w *= o.w > 0.f ? 1.f : a.slope;
        s1.x += g.x, s1.y += g.y, s1.z += g.z, s1.w += g.w;
        const float4 xh = make_float4((x.x - mu.x) * is.x, (x.y - mu.y) * is.y, (x.z - mu.z) * is.z, (x.w - mu.w) * is.w);
        s2.x = fmaf(g.x, xh.x, s2.x), s2.y = fmaf(g.y, xh.y, s2.y);
        s2.z = fmaf(g.z, xh.z, s2.z), s2.w = fmaf(g.w, xh.w, s2.w);
        gm = fmax4abs(gm, g), xm = fmax4abs(xm, xh);
        if (a.cs) {
          const float4 xs = xsv[u];
          const float4 xsh = make_float4((xs.x - mus.x) * iss.x, (xs.y - mus.y) * iss.y, (xs.z - mus.z) * iss.z,
                                         (xs.w - mus.w) * iss.w);
          s3.x = fmaf(g.x, xsh.x, s3.x), s3.y = fmaf(g.y, xsh.y, s3.y);
          s3.z = fmaf(g.z, xsh.z, s3.z), s3.w = fmaf(g.w, xsh.w, s3.w);
          xsm = fmax4abs(xsm, xsh);
        }
      }
    }
  }
  red[0][threadIdx.x] = s1, red[1][threadIdx.x] = s2, red[2][threadIdx.x] = s3;
  __syncthreads();
  if (threadIdx.x < C4) {
    for (int k = 0; k < (a.cs ? 3 : 2); ++k) {
      float4 t = red[k][cq];
      for (int r = 1; r < RL; ++r) {
        float4 u = red[k][r * C4 + cq];
        t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
      }
      if (a.tot) {  // per-channel totals: the apply kernel has nothing to re-sum
        double* dst = a.tot + ((int64_t)(blockIdx.x % kBnBwdTotCopies) * a.C + cq * 4) * 3;
        atomicAdd(dst + 0 * 3 + k, (double)t.x), atomicAdd(dst + 1 * 3 + k, (double)t.y);
        atomicAdd(dst + 2 * 3 + k, (double)t.z), atomicAdd(dst + 3 * 3 + k, (double)t.w);
      } else {
        float* dst = a.part + ((int64_t)blockIdx.x * a.C + cq * 4) * 3;
        dst[0 * 3 + k] = t.x, dst[1 * 3 + k] = t.y, dst[2 * 3 + k] = t.z, dst[3 * 3 + k] = t.w;
      }
    }
  }
  if (a.dc_slot) {  // uniform branch
    const float g_all = block_max(max4(gm), mred), x_all = block_max(max4(xm), mred), k_all = block_max(km, mred);
    float xs_all = 0.f, ks_all = 0.f;
    if (a.dcs_slot) xs_all = block_max(max4(xsm), mred), ks_all = block_max(ksm, mred);
    if (threadIdx.x == 0) {
      unsigned int* s = reinterpret_cast<unsigned int*>(a.dc_slot);
      atomicMax(s + 0, __float_as_uint(g_all)), atomicMax(s + 1, __float_as_uint(x_all));
      atomicMax(s + 2, __float_as_uint(k_all));
      if (a.dcs_slot) {
        unsigned int* ss = reinterpret_cast<unsigned int*>(a.dcs_slot);
        atomicMax(ss + 0, __float_as_uint(g_all)), atomicMax(ss + 1, __float_as_uint(xs_all));
        atomicMax(ss + 2, __float_as_uint(ks_all));
      }
    }
  }
}

__device__ __forceinline__ float4 bn_dx(float4 g, float4 x, float4 mu, float4 is, float4 k, float4 m1, float4 m2) {
  float4 r;
  r.x = k.x * (g.x - m1.x - (x.x - mu.x) * is.x * m2.x);
  r.y = k.y * (g.y - m1.y - (x.y - mu.y) * is.y * m2.y);
  r.z = k.z * (g.z - m1.z - (x.z - mu.z) * is.z * m2.z);
  r.w = k.w * (g.w - m1.w - (x.w - mu.w) * is.w * m2.w);
  return r;
}
// |dc| = |k (g - m1 - xhat m2)| with |m1| <= max|g| and |m2| = |mean(g xhat)| <= max|g| mean|xhat| <= max|g|
// (mean xhat^2 = 1), so |dc| <= max|k| max|g| (2 + max|xhat|): a safe, slightly loose bound (the pair format keeps
// full precision over 18 bits of dynamic range below the scaled maximum)
__device__ __forceinline__ float dc_scale(const float* slot) {
  return pair_scale_from_bound(slot[2] * slot[0] * (2.f + slot[1]));
}

// grid = (C / 32, row chunks); thread = (channel quad of the slab, row lane)
__global__ void HP_BN_BOUNDS bn_bwd_apply_kernel(BnBwd a, int nchunks) {
  __shared__ double s_sum[8][kSlab][3];
  __shared__ float s_m1[kSlab], s_m2[kSlab], s_m3[kSlab];
  pdl_trigger();
  pdl_wait();
  const int c0 = blockIdx.x * kSlab;
  {
    const int c = threadIdx.x & (kSlab - 1), slice = threadIdx.x >> 5;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (a.tot) {
      if (slice == 0) {
        double v[kBnBwdTotCopies][3];
#pragma unroll
        for (int k = 0; k < kBnBwdTotCopies; ++k) {
          const double* p = a.tot + ((int64_t)k * a.C + c0 + c) * 3;
          v[k][0] = __ldcg(p), v[k][1] = __ldcg(p + 1), v[k][2] = __ldcg(p + 2);
        }
#pragma unroll
        for (int k = 0; k < kBnBwdTotCopies; ++k) s1 += v[k][0], s2 += v[k][1], s3 += v[k][2];
      }
    } else
    for (int t0 = slice; t0 < nchunks; t0 += 64) {  // 8 chunks (24 loads) in flight per thread
      float v[8][3];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + 8 * u;
        const float* p = a.part + ((int64_t)min(t, nchunks - 1) * a.C + c0 + c) * 3;
        const bool ok = t < nchunks;
        v[u][0] = ok ? __ldcg(p) : 0.f, v[u][1] = ok ? __ldcg(p + 1) : 0.f, v[u][2] = ok ? __ldcg(p + 2) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s1 += (double)v[u][0], s2 += (double)v[u][1], s3 += (double)v[u][2];
    }
    if (!a.tot) {
      s_sum[slice][c][0] = s1, s_sum[slice][c][1] = s2, s_sum[slice][c][2] = s3;
      __syncthreads();
    }
    if (slice == 0) {
      if (!a.tot)
        for (int k = 1; k < 8; ++k) s1 += s_sum[k][c][0], s2 += s_sum[k][c][1], s3 += s_sum[k][c][2];
      const double inv_n = a.inv_n;
      s_m1[c] = (float)(s1 * inv_n), s_m2[c] = (float)(s2 * inv_n), s_m3[c] = (float)(s3 * inv_n);
      if (blockIdx.y == 0) {
        a.dgamma[c0 + c] = (float)s2, a.dbeta[c0 + c] = (float)s1;
        if (a.cs) a.dgamma_s[c0 + c] = (float)s3, a.dbeta_s[c0 + c] = (float)s1;
      }
    }
    __syncthreads();
  }
  float sc = 1.f, scs = 1.f;
  if (a.dc_p) sc = dc_scale(a.dc_slot);
  if (a.dcs_p) scs = dc_scale(a.dcs_slot);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    if (a.dc_p) a.dc_slot[3] = 1.f / sc;
    if (a.dcs_p) a.dcs_slot[3] = 1.f / scs;
  }
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = c0 + q * 4;
  const float4 m1 = *reinterpret_cast<const float4*>(s_m1 + q * 4), m2 = *reinterpret_cast<const float4*>(s_m2 + q * 4);
  const float4 m3 = *reinterpret_cast<const float4*>(s_m3 + q * 4);
  const float4 mu = *reinterpret_cast<const float4*>(a.coef + 2 * a.C + col);
  const float4 is = *reinterpret_cast<const float4*>(a.coef + 3 * a.C + col);
  const float4 ga = *reinterpret_cast<const float4*>(a.gamma + col);
  const float4 k = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
  float4 mus = mu, iss = is, ks = k;
  if (a.cs) {
    mus = *reinterpret_cast<const float4*>(a.coef_s + 2 * a.C + col);
    iss = *reinterpret_cast<const float4*>(a.coef_s + 3 * a.C + col);
    const float4 gs = *reinterpret_cast<const float4*>(a.gamma_s + col);
    ks = make_float4(gs.x * iss.x, gs.y * iss.y, gs.z * iss.z, gs.w * iss.w);
  }
  const int M = a.B * a.L;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_end = min(M, (int)(blockIdx.y + 1) * rows_per);
  constexpr int U = 2;
  for (int m0 = blockIdx.y * rows_per + rl; m0 < m_end; m0 += 32 * U) {
    float4 gv[U], ov[U], xv[U], xsv[U];
    int bb[U], ll[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int m = m0 + 32 * u;
      if (m < m_end) {
        const int b = m / a.L, l = m - b * a.L;
        const int64_t row = (int64_t)b * (a.L + 2) + 1 + l;
        bb[u] = b, ll[u] = l;
        gv[u] = load_g(a, b, l, col);
        ov[u] = *reinterpret_cast<const float4*>(a.out + row * a.C + col);
        xv[u] = *reinterpret_cast<const float4*>(a.c + row * a.C + col);
        if (a.cs) xsv[u] = *reinterpret_cast<const float4*>(a.cs + row * a.C + col);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (m0 + 32 * u >= m_end) break;
      const int b = bb[u], l = ll[u];
      const int64_t row = (int64_t)b * (a.L + 2) + 1 + l;
      float4 g = gv[u];
      const float4 o = ov[u];
      g.x *= o.x > 0.f ? 1.f : a.slope, g.y *= o.y > 0.f ? 1.f : a.slope;
      g.z *= o.z > 0.f ? 1.f : a.slope, g.w *= o.w > 0.f ? 1.f : a.slope;
      if (a.gres) *reinterpret_cast<float4*>(a.gres + row * a.C + col) = g;
      {
        const int64_t drow = (int64_t)b * (a.Ld + 2) + 1 + (int64_t)a.dil * l;
        const float4 d = bn_dx(g, xv[u], mu, is, k, m1, m2);
        if (a.dc) *reinterpret_cast<float4*>(a.dc + drow * a.C + col) = d;
        if (a.dc_p) store_pair4(a.dc_p, a.dc_ps, drow * a.C + col, make_float4(d.x * sc, d.y * sc, d.z * sc, d.w * sc));
      }
      if (a.cs) {
        const int64_t drow = (int64_t)b * (a.Ld_s + 2) + 1 + (int64_t)a.dil_s * l;
        const float4 d = bn_dx(g, xsv[u], mus, iss, ks, m1, m3);
        if (a.dcs) *reinterpret_cast<float4*>(a.dcs + drow * a.C + col) = d;
        if (a.dcs_p)
          store_pair4(a.dcs_p, a.dcs_ps, drow * a.C + col, make_float4(d.x * scs, d.y * scs, d.z * scs, d.w * scs));
      }
    }
  }
}

__global__ void __launch_bounds__(256) pairsum_acc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B,
                                                          int L, int C) {
  pdl_trigger();
  pdl_wait();
  const int C4 = C >> 2;
  const int64_t total = (int64_t)B * L * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % C4);
    const int64_t m = i / C4;
    const int b = (int)(m / L), l = (int)(m - (int64_t)b * L);
    const int64_t row = (int64_t)b * (L + 2) + 1 + l, ru = (int64_t)b * (2 * L + 2) + 1 + 2 * l;
    float4 d = *reinterpret_cast<float4*>(dst + row * C + cq * 4);
    const float4 s0 = *reinterpret_cast<const float4*>(src + ru * C + cq * 4);
    const float4 s1 = *reinterpret_cast<const float4*>(src + (ru + 1) * C + cq * 4);
    d.x += s0.x + s1.x, d.y += s0.y + s1.y, d.z += s0.z + s1.z, d.w += s0.w + s1.w;
    *reinterpret_cast<float4*>(dst + row * C + cq * 4) = d;
  }
}

// ------------------------------------------------------------------------------------------------
// Encoder tail: adaptive_avg_pool1d(x, 1) + Linear(512 -> F)   reference hippie/backbones.py:100-102
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pool_linear_fwd_kernel(const float* __restrict__ x4, int L, int C,
                                                              const float* __restrict__ W,
                                                              const float* __restrict__ bias, int F,
                                                              float* __restrict__ pooled, float* __restrict__ h) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sp[];  // [C]
  const int b = blockIdx.x;
  const float inv = 1.f / (float)L;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += x4[((int64_t)b * (L + 2) + 1 + l) * C + c];
    s *= inv;
    sp[c] = s;
    pooled[(int64_t)b * C + c] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int f = warp; f < F; f += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(sp[c], __ldg(W + (int64_t)f * C + c), s);
    s = warp_sum(s);
    if (lane == 0) h[(int64_t)b * F + f] = s + bias[f];
  }
}

__global__ void __launch_bounds__(128) pool_linear_bwd_x_kernel(const float* __restrict__ dh,
                                                                const float* __restrict__ W, int L, int C, int F,
                                                                float* __restrict__ g_x4) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sd[];  // [F]
  const int b = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x) sd[f] = dh[(int64_t)b * F + f];
  __syncthreads();
  const float inv = 1.f / (float)L;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int f = 0; f < F; ++f) s = fmaf(sd[f], __ldg(W + (int64_t)f * C + c), s);
    s *= inv;
    for (int l = 0; l < L; ++l) g_x4[((int64_t)b * (L + 2) + 1 + l) * C + c] = s;
  }
}

// dW[j][i] = sum_b dy[b][j] * x[b][i];  db[j] = sum_b dy[b][j]   (column i == nin is the bias).
// 64 outputs per CTA, the batch reduction is split over 4 thread groups and combined in shared memory.
__global__ void __launch_bounds__(256) linear_wgrad_rows_kernel(const float* __restrict__ dy, int ldy,
                                                                const float* __restrict__ x, int ldx, int B, int nin,
                                                                int nout, float* __restrict__ dW,
                                                                float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[4][64];
  const int o = threadIdx.x & 63, bg = threadIdx.x >> 6;
  const int idx = blockIdx.x * 64 + o;
  const int j = idx / (nin + 1), i = idx - j * (nin + 1);
  float s = 0.f;
  if (j < nout) {
    // 16 samples per thread in flight: the loop is a chain of L2 round trips otherwise (54 us -> measured below)
    constexpr int U = 16;
    const bool bias = i >= nin;
    for (int b0 = bg; b0 < B; b0 += 4 * U) {
      float dv[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int b = b0 + 4 * u;
        const bool ok = b < B;
        dv[u] = ok ? __ldg(dy + (int64_t)b * ldy + j) : 0.f;
        xv[u] = (ok && !bias) ? __ldg(x + (int64_t)b * ldx + i) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) s = fmaf(dv[u], xv[u], s);
    }
  }
  red[bg][o] = s;
  __syncthreads();
  if (bg == 0 && j < nout) {
    s = (red[0][o] + red[1][o]) + (red[2][o] + red[3][o]);
    if (i < nin)
      dW[(int64_t)j * nin + i] = s;
    else if (db)
      db[j] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder head: Linear(F -> 512), unsqueeze(-1), nearest x4   reference hippie/backbones.py:129-131
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_linear_fwd_kernel(const float* __restrict__ d, int F,
                                                             const float* __restrict__ W,
                                                             const float* __restrict__ bias, int C,
                                                             float* __restrict__ t0, uint16_t* __restrict__ t0_p,
                                                             int64_t t0_ps, unsigned* __restrict__ flags) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sd[];  // [F]
  const int b = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x) sd[f] = d[(int64_t)b * F + f];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    const float* w = W + (int64_t)c * F;
    for (int f = 0; f < F; ++f) s = fmaf(sd[f], __ldg(w + f), s);
    s += bias[c];
#pragma unroll
    for (int l = 0; l < 4; ++l) t0[((int64_t)b * 6 + 1 + l) * C + c] = s;
    if (t0_p) {
      uint16_t h, lo;
      pair_split<kPairF16>(s, h, lo);
      if (flags && !(fabsf(s) < kPairF16Max)) atomicOr(flags, kFlagPairSaturated);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        t0_p[((int64_t)b * 6 + 1 + l) * C + c] = h;
        t0_p[t0_ps + ((int64_t)b * 6 + 1 + l) * C + c] = lo;
      }
    }
  }
}

__global__ void __launch_bounds__(256) dec_linear_bwd_x_kernel(const float* __restrict__ g_t0,
                                                               const float* __restrict__ W, int F, int C,
                                                               float* __restrict__ gx0, float* __restrict__ dd) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sg[];  // [C]
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 4; ++l) s += g_t0[((int64_t)b * 6 + 1 + l) * C + c];
    sg[c] = s;
    gx0[(int64_t)b * C + c] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int f = warp; f < F; f += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(sg[c], __ldg(W + (int64_t)c * F + f), s);
    s = warp_sum(s);
    if (lane == 0) dd[(int64_t)b * F + f] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder tail + MSE (+ backward down to the layer1 output)
//   reference hippie/backbones.py:117-118,136-139; hippie/model.py:465-466
// ------------------------------------------------------------------------------------------------
constexpr int kTailSPB = 4;  // samples per CTA
__global__ void __launch_bounds__(256) dec_tail_kernel(DecTail t) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* xs = sm;                      // [34][65]
  float* Wos = xs + 34 * 65;           // [Lo][65]
  float* ys = Wos + t.Lo * 65;         // [64]
  float* dds = ys + 64;                // [Lo]
  float* dys = dds + ((t.Lo + 3) & ~3);  // [66]: dys[1 + p], zero guards
  float* wcs = dys + 68;               // [192]
  __shared__ float sred[8];
  const int tid = threadIdx.x;
  for (int i = tid; i < t.Lo * 64; i += 256) Wos[(i >> 6) * 65 + (i & 63)] = t.Wo[i];
  for (int i = tid; i < 192; i += 256) wcs[i] = t.wc[i];
  if (tid < 65) xs[0 * 65 + tid] = 0.f, xs[33 * 65 + tid] = 0.f;
  if (tid == 0) dys[0] = 0.f, dys[65] = 0.f;
  const float bc = t.bc[0];
  float dw_acc = 0.f, dbc_acc = 0.f, sse_acc = 0.f;
  const float gscale = t.train ? t.loss_w * 2.f / ((float)t.B * (float)t.Lo) : 0.f;

  for (int sb = 0; sb < kTailSPB; ++sb) {
    const int b = blockIdx.x * kTailSPB + sb;
    if (b >= t.B) break;  // uniform across the CTA
    __syncthreads();
    for (int i = tid; i < 32 * 64; i += 256) xs[(1 + (i >> 6)) * 65 + (i & 63)] = t.x[((int64_t)b * 34 + 1) * 64 + i];
    __syncthreads();
    {  // y[p] = bc + sum_t sum_c u[p+t-1][c] * wc[t][c];  u[q] = x[q >> 1], zero outside [0, 64)
      const int p = tid >> 2, part = tid & 3;
      float s = 0.f;
#pragma unroll
      for (int tt = 0; tt < 3; ++tt) {
        const int q = p + tt - 1;
        const int row = q < 0 ? 0 : (q >= 64 ? 33 : 1 + (q >> 1));
#pragma unroll
        for (int c = 0; c < 16; ++c) s = fmaf(xs[row * 65 + part * 16 + c], wcs[tt * 64 + part * 16 + c], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0) {
        s += bc;
        ys[p] = s;
        if (t.y) t.y[(int64_t)b * 64 + p] = s;
      }
    }
    __syncthreads();
    float sq = 0.f;
    if (tid < t.Lo) {
      float s = 0.f;
#pragma unroll 8
      for (int p = 0; p < 64; ++p) s = fmaf(ys[p], Wos[tid * 65 + p], s);
      s += t.bo[tid];
      t.dec[(int64_t)b * t.Lo + tid] = s;
      if (t.target) {
        const float diff = s - t.target[(int64_t)b * t.Lo + tid];
        sq = diff * diff;
        const float gd = gscale * diff;
        dds[tid] = gd;
        if (t.train) t.ddec[(int64_t)b * t.Lo + tid] = gd;
      }
    }
    if (t.target) {
      sq = warp_sum(sq);
      if ((tid & 31) == 0) sred[tid >> 5] = sq;
    }
    __syncthreads();
    if (t.target && tid == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += sred[w];
      sse_acc += s;
    }
    if (!t.train) continue;
    if (tid < 64) {  // dy[p] = sum_o ddec[o] * Wo[o][p]
      float s = 0.f;
      for (int o = 0; o < t.Lo; ++o) s = fmaf(dds[o], Wos[o * 65 + tid], s);
      dys[1 + tid] = s;
      t.dy[(int64_t)b * 64 + tid] = s;
    }
    __syncthreads();
    {  // gradient w.r.t. x: du[q][c] = sum_t dy[q-t+1] * wc[t][c];  g_x[l] = du[2l] + du[2l+1]
      const int c = tid & 63, lg = tid >> 6;
      const float w0 = wcs[c], w1 = wcs[64 + c], w2 = wcs[128 + c];
      for (int l = lg; l < 32; l += 4) {
        const int q = 2 * l;  // dys index offset +1
        const float du0 = dys[1 + q + 1] * w0 + dys[1 + q] * w1 + dys[1 + q - 1] * w2;
        const float du1 = dys[1 + q + 2] * w0 + dys[1 + q + 1] * w1 + dys[1 + q] * w2;
        t.g_x[((int64_t)b * 34 + 1 + l) * 64 + c] = du0 + du1;
      }
    }
    if (tid < 192) {  // dwc[t][c] += sum_p dy[p] * u[p+t-1][c]
      const int tt = tid >> 6, c = tid & 63;
      float s = 0.f;
      for (int p = 0; p < 64; ++p) {
        const int q = p + tt - 1;
        const int row = q < 0 ? 0 : (q >= 64 ? 33 : 1 + (q >> 1));
        s = fmaf(dys[1 + p], xs[row * 65 + c], s);
      }
      dw_acc += s;
    } else if (tid == 192) {
      float s = 0.f;
      for (int p = 0; p < 64; ++p) s += dys[1 + p];
      dbc_acc += s;
    }
  }
  if (t.part) {
    float* pp = t.part + (int64_t)blockIdx.x * 196;
    if (tid < 192) pp[tid] = dw_acc;
    if (tid == 192) pp[192] = dbc_acc;
    if (tid == 0) pp[193] = sse_acc;
  }
}

// grid = Lo + 1.  block o < Lo: dWo[o][p] = sum_b ddec[b][o] * y[b][p], dbo[o];  last block: partials.
__global__ void __launch_bounds__(256) dec_tail_reduce_kernel(DecTail t, int ncta, float* dwc, float* dbc, float* dWo,
                                                              float* dbo, float* sse) {
  pdl_trigger();
  pdl_wait();
  const int o = blockIdx.x, tid = threadIdx.x;
  if (o < t.Lo) {
    if (!t.train) return;
    __shared__ float red[4][64];
    const int p = tid & 63, bg = tid >> 6;
    float s = 0.f, sb = 0.f;
    for (int b = bg; b < t.B; b += 4) {
      const float g = t.ddec[(int64_t)b * t.Lo + o];
      s = fmaf(g, t.y[(int64_t)b * 64 + p], s);
      sb += g;
    }
    red[bg][p] = s;
    __syncthreads();
    if (bg == 0) dWo[o * 64 + p] = red[0][p] + red[1][p] + red[2][p] + red[3][p];
    __syncthreads();
    if (p == 0) red[bg][0] = sb;
    __syncthreads();
    if (tid == 0) dbo[o] = red[0][0] + red[1][0] + red[2][0] + red[3][0];
  } else {
    for (int i = tid; i < 194; i += 256) {
      if (i < 193 && !t.train) continue;
      double s = 0.0;
      for (int c = 0; c < ncta; ++c) s += (double)t.part[(int64_t)c * 196 + i];
      if (i < 192)
        dwc[i] = (float)s;
      else if (i == 192)
        dbc[0] = (float)s;
      else
        *sse = (float)s;
    }
  }
}

__global__ void loss_finalize_kernel(const float* sse1, const float* sse2, const float* kl_parts, int n_kl, int B, int Lo1,
                                     int Lo2,
                                     float beta, float w1, float w2, int multimodal, float* scalars) {
  pdl_trigger();
  pdl_wait();
  const float mse1 = *sse1 / ((float)B * (float)Lo1);
  const float mse2 = multimodal ? *sse2 / ((float)B * (float)Lo2) : 0.f;
  float kl_sum = 0.f;
  for (int i = 0; i < n_kl; ++i) kl_sum += kl_parts[i];
  const float kl = kl_sum / (float)B;
  const float mse = multimodal ? w1 * mse1 + w2 * mse2 : mse1;
  scalars[0] = mse + beta * kl;
  scalars[1] = mse1;
  scalars[2] = mse2;
  scalars[3] = kl;
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ (torch/nn/utils/clip_grad.py; Lightning gradient_clip_val,
// reference scripts/train_model_with_multimodal.py:55,701) + torch.optim.AdamW (hippie/model.py:447)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float scale,
                                                    float* __restrict__ partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(g + i * 4);
    v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)  // the n % 4 tail (the flat buffers are padded to whole groups today)
    for (int64_t i = n4 * 4; i < n; ++i) s += (g[i] * scale) * (g[i] * scale);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) clip_coef_kernel(const float* __restrict__ partials, int n, float max_norm,
                                                        float* __restrict__ scalars) {
  pdl_trigger();
  pdl_wait();
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float norm = (float)sqrt(t);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);
    scalars[4] = norm;
    scalars[5] = coef;
  }
}

struct AdamCoefs {  // fp32 roundings of the double expressions torch evaluates on the host (torch/optim/adam.py)
  float decay, step_size, bc2_sqrt, step_size_cls, bc2_sqrt_cls, omb1, omb2, beta2, eps;
};
__device__ __forceinline__ void adamw_one(float g, float& p, float& m, float& v, float coef, float gscale, float decay,
                                          float omb1, float omb2, float beta2, float ss, float bs, float eps) {
  g = g * gscale;
  g *= coef;
  p = p * decay;
  m = m + omb1 * (g - m);  // lerp_(grad, 1 - beta1)
  v = v * beta2;
  v = fmaf(omb2 * g, g, v);  // addcmul_(grad, grad, value = 1 - beta2)
  const float denom = sqrtf(v) / bs + eps;
  p = p - ss * (m / denom);
}
// torch.optim.AdamW's single-tensor update (torch/optim/adam.py) over the flat buffers, four elements per thread and
// iteration (128-bit loads / stores: the kernel moves 7 x 64 MB and is HBM-bound); the class-embedding range carries
// its own step count (or is skipped when it has no gradient), groups that straddle its ends go element by element.
__global__ void __launch_bounds__(256) adamw_kernel(AdamArgs a, AdamCoefs k) {
  pdl_trigger();
  pdl_wait();
  const float coef = a.scalars[5];
  const int64_t n4 = a.n >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4 + 1; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = q * 4;
    const int cnt = q < n4 ? 4 : (int)(a.n - i0);  // the last group holds the n % 4 tail
    if (cnt <= 0) break;
    const bool in0 = i0 >= a.skip_lo && i0 < a.skip_hi, in3 = i0 + 3 >= a.skip_lo && i0 + 3 < a.skip_hi;
    if (cnt == 4 && in0 == in3 && !(i0 < a.skip_lo && i0 + 3 >= a.skip_hi)) {
      if (in0 && !a.has_cls_grad) continue;
      const float ss = in0 ? k.step_size_cls : k.step_size, bs = in0 ? k.bc2_sqrt_cls : k.bc2_sqrt;
      const float4 g = *reinterpret_cast<const float4*>(a.g + i0);
      float4 p = *reinterpret_cast<const float4*>(a.p + i0), m = *reinterpret_cast<const float4*>(a.m + i0);
      float4 v = *reinterpret_cast<const float4*>(a.v + i0);
      adamw_one(g.x, p.x, m.x, v.x, coef, a.grad_scale, k.decay, k.omb1, k.omb2, k.beta2, ss, bs, k.eps);
      adamw_one(g.y, p.y, m.y, v.y, coef, a.grad_scale, k.decay, k.omb1, k.omb2, k.beta2, ss, bs, k.eps);
      adamw_one(g.z, p.z, m.z, v.z, coef, a.grad_scale, k.decay, k.omb1, k.omb2, k.beta2, ss, bs, k.eps);
      adamw_one(g.w, p.w, m.w, v.w, coef, a.grad_scale, k.decay, k.omb1, k.omb2, k.beta2, ss, bs, k.eps);
      *reinterpret_cast<float4*>(a.p + i0) = p, *reinterpret_cast<float4*>(a.m + i0) = m;
      *reinterpret_cast<float4*>(a.v + i0) = v;
      if (a.wp_hi)
        store_pair4(a.wp_hi, a.wp_lo - a.wp_hi, i0,
                    make_float4(p.x * a.wp_scale, p.y * a.wp_scale, p.z * a.wp_scale, p.w * a.wp_scale), a.flags,
                    kFlagWeightSaturated);
    } else {
      for (int e = 0; e < cnt; ++e) {
        const int64_t i = i0 + e;
        const bool in = i >= a.skip_lo && i < a.skip_hi;
        if (in && !a.has_cls_grad) continue;
        float p = a.p[i], m = a.m[i], v = a.v[i];
        adamw_one(a.g[i], p, m, v, coef, a.grad_scale, k.decay, k.omb1, k.omb2, k.beta2, in ? k.step_size_cls : k.step_size,
                  in ? k.bc2_sqrt_cls : k.bc2_sqrt, k.eps);
        a.p[i] = p, a.m[i] = m, a.v[i] = v;
        if (a.wp_hi) {
          const float x = p * a.wp_scale;
          pair_split<kPairF16>(x, a.wp_hi[i], a.wp_lo[i]);
          if (a.flags && !(fabsf(x) < kPairF16Max)) atomicOr(a.flags, kFlagWeightSaturated);
        }
      }
    }
  }
}

// wt[ci][k-1-t][co] = w[co][t][ci]
// ------------------------------------------------------------------------------------------------
// Batch preprocessing (EphysDataset.__getitem__, reference hippie/dataloading.py:27-56): rows index[b] of a raw float64
// table -> float32, optional log(x + 1), linear interpolation with align_corners=False exactly as ATen's CPU kernel
// evaluates it (oracle/cvae_oracle.py:interp_linear):  scale = fp32(n) / size;  src = max(fma(scale, i + 0.5, -0.5), 0);
// i0 = min(floor(src), n - 1);  i1 = min(i0 + 1, n - 1);  w1 = src - i0;  w0 = 1 - w1;  out = fma(x[i0], w0, x[i1] * w1).
// The waveform path is bit-exact; log() is the correctly rounded float of the double logarithm, which differs from
// ATen's vectorised logf by one ulp for ~0.3 % of the values.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preprocess_kernel(const double* __restrict__ raw, int width,
                                                         const int64_t* __restrict__ index, int B, int size, int take_log,
                                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * size) return;
  const int b = i / size, j = i - b * size;
  const double* row = raw + (index ? index[b] : (int64_t)b) * width;
  const float scale = __fdiv_rn((float)width, (float)size);
  const float src = fmaxf(__fmaf_rn(scale, (float)j + 0.5f, -0.5f), 0.f);
  const int i0 = min((int)floorf(src), width - 1), i1 = min(i0 + 1, width - 1);
  const float w1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f), w0 = __fsub_rn(1.f, w1);
  float x0 = (float)row[i0], x1 = (float)row[i1];
  if (take_log) x0 = (float)log((double)__fadd_rn(x0, 1.f)), x1 = (float)log((double)__fadd_rn(x1, 1.f));
  out[i] = __fmaf_rn(x0, w0, __fmul_rn(x1, w1));
}

__global__ void __launch_bounds__(256) io_copy_kernel(IoCopy c) {
  const IoSeg sg = c.seg[blockIdx.y];
  const uint32_t* src = static_cast<const uint32_t*>(sg.src);
  uint32_t* dst = static_cast<uint32_t*>(sg.dst);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.words; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

__global__ void __launch_bounds__(256) refresh_wt_kernel(const WtEntry* __restrict__ tab,
                                                         const float* __restrict__ params, float* __restrict__ ws) {
  __shared__ float tile[32][33];
  const WtEntry e = tab[blockIdx.z];
  const int nco = e.cout / 32, nci = e.cin / 32;
  const int tiles = nco * nci * e.k;
  for (int tt = blockIdx.x; tt < tiles; tt += gridDim.x) {
    const int t = tt % e.k, rest = tt / e.k, cit = rest % nci, cot = rest / nci;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    __syncthreads();
    for (int r = ty; r < 32; r += 8)  // read rows co, contiguous ci
      tile[r][tx] = params[e.w_off + ((int64_t)(cot * 32 + r) * e.k + t) * e.cin + cit * 32 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)  // write rows ci, contiguous co
      ws[e.wt_off + ((int64_t)(cit * 32 + r) * e.k + (e.k - 1 - t)) * e.cout + cot * 32 + tx] = tile[tx][r];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static inline int ew_grid(int64_t work_items, int block = 256) {
  int64_t g = (work_items + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

void launch_stem_fwd(const float* x, const float* w, float* c0, float* part, int B, int Lin, int Lout, cudaStream_t s) {
  int grid = (B * Lout + 127) / 128;
  launch_pdl(stem_fwd_kernel, dim3(grid), dim3(256), 0, s, x, w, c0, part, B, Lin, Lout);
}
int launch_stem_wgrad(const float* x, const float* dc0, float* part, int B, int Lin, int Lout, cudaStream_t s) {
  int grid = (B * Lout + 127) / 128;
  launch_pdl(stem_wgrad_kernel, dim3(grid), dim3(256), 0, s, x, dc0, part, B, Lin, Lout);
  return grid;
}
void launch_reduce_partials(const float* part, int nparts, int n, float* out, int accumulate, cudaStream_t s) {
  launch_pdl(reduce_partials_kernel, dim3((n + 127) / 128), dim3(128), 0, s, part, nparts, n, out, accumulate);
}
void launch_bn_eval_coefs(const BnEvalEntry* table_dev, int n, const float* params, const float* run_mean,
                          const float* run_var, float* ws, cudaStream_t s) {
  bn_eval_coefs_kernel<<<n, 128, 0, s>>>(table_dev, params, run_mean, run_var, ws);
}
// row chunks so that slabs x chunks ~ 2 CTAs per SM, each with at least 32 rows
static inline int slab_row_chunks(int M, int C, int sm_count) {
  int chunks = (2 * sm_count) / (C / kSlab);
  const int max_chunks = (M + 31) / 32;
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks < 1 ? 1 : chunks;
}
void launch_bn_apply(const BnApply& a, int sm_count, cudaStream_t s) {
  dim3 grid(a.C / kSlab, slab_row_chunks(a.B * a.L, a.C, sm_count));
  if (!a.r)
    launch_pdl(bn_apply_kernel<0>, grid, dim3(256), 0, s, a);
  else if (!a.rcoef)
    launch_pdl(bn_apply_kernel<1>, grid, dim3(256), 0, s, a);
  else
    launch_pdl(bn_apply_kernel<2>, grid, dim3(256), 0, s, a);
}
void launch_bn_bwd(const BnBwd& a, int sm_count, cudaStream_t s) {
  const int M = a.B * a.L;
  int rows = (M + kBnBwdMaxChunks - 1) / kBnBwdMaxChunks;
  if (rows < 16) rows = 16;
  const int nchunks = (M + rows - 1) / rows;
  // HIPPIE_B200_DEBUG_SKIP bits 2 / 4: knock-out experiments (wrong results: no reduce / no apply launch), only in builds
  // made with EXTRA=-DHP_EXPERIMENTS
#ifdef HP_EXPERIMENTS
  static const int dbg_skip = getenv("HIPPIE_B200_DEBUG_SKIP") ? atoi(getenv("HIPPIE_B200_DEBUG_SKIP")) : 0;
#else
  constexpr int dbg_skip = 0;
#endif
  if (!(dbg_skip & 2)) launch_pdl(bn_bwd_reduce_kernel<4>, dim3(nchunks), dim3(256), 0, s, a, rows);
  dim3 grid(a.C / kSlab, slab_row_chunks(M, a.C, sm_count));
  if (!(dbg_skip & 4)) launch_pdl(bn_bwd_apply_kernel, grid, dim3(256), 0, s, a, nchunks);
}
void launch_pairsum_acc(const float* src, float* dst, int B, int L, int C, cudaStream_t s) {
  launch_pdl(pairsum_acc_kernel, dim3(ew_grid((int64_t)B * L * (C / 4))), dim3(256), 0, s, src, dst, B, L, C);
}
void launch_pool_linear_fwd(const float* x4, int B, int L, int C, const float* W, const float* bias, int F,
                            float* pooled, float* h, cudaStream_t s) {
  launch_pdl(pool_linear_fwd_kernel, dim3(B), dim3(128), C * sizeof(float), s, x4, L, C, W, bias, F, pooled, h);
}
void launch_pool_linear_bwd_x(const float* dh, const float* W, int B, int L, int C, int F, float* g_x4, cudaStream_t s) {
  launch_pdl(pool_linear_bwd_x_kernel, dim3(B), dim3(128), F * sizeof(float), s, dh, W, L, C, F, g_x4);
}
void launch_linear_wgrad(const float* dy, int ldy, const float* x, int ldx, int B, int nin, int nout, float* dW, float* db,
                         cudaStream_t s) {
  launch_pdl(linear_wgrad_rows_kernel, dim3((nout * (nin + 1) + 63) / 64), dim3(256), 0, s, dy, ldy, x, ldx, B, nin, nout, dW, db);
}
void launch_dec_linear_fwd(const float* d, int B, int F, const float* W, const float* bias, int C, float* t0,
                           uint16_t* t0_p, int64_t t0_ps, unsigned* flags, cudaStream_t s) {
  launch_pdl(dec_linear_fwd_kernel, dim3(B), dim3(256), F * sizeof(float), s, d, F, W, bias, C, t0, t0_p, t0_ps, flags);
}
void launch_dec_linear_bwd_x(const float* g_t0, const float* W, int B, int F, int C, float* gx0, float* dd,
                             cudaStream_t s) {
  launch_pdl(dec_linear_bwd_x_kernel, dim3(B), dim3(256), C * sizeof(float), s, g_t0, W, F, C, gx0, dd);
}
static size_t dec_tail_smem(int Lo) { return (size_t)(34 * 65 + Lo * 65 + 64 + ((Lo + 3) & ~3) + 68 + 192) * sizeof(float); }
int launch_dec_tail(const DecTail& t, cudaStream_t s) {
  const int ncta = (t.B + kTailSPB - 1) / kTailSPB;
  const size_t smem = dec_tail_smem(t.Lo);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(dec_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  launch_pdl(dec_tail_kernel, dim3(ncta), dim3(256), smem, s, t);
  return ncta;
}
void launch_dec_tail_reduce(const DecTail& t, int ncta, float* dwc, float* dbc, float* dWo, float* dbo, float* sse,
                            cudaStream_t s) {
  launch_pdl(dec_tail_reduce_kernel, dim3(t.Lo + 1), dim3(256), 0, s, t, ncta, dwc, dbc, dWo, dbo, sse);
}
void launch_loss_finalize(const float* sse1, const float* sse2, const float* kl_parts, int n_kl, int B, int Lo1, int Lo2,
                          float beta, float w1, float w2, int multimodal, float* scalars, cudaStream_t s) {
  launch_pdl(loss_finalize_kernel, dim3(1), dim3(1), 0, s, sse1, sse2, kl_parts, n_kl, B, Lo1, Lo2, beta, w1, w2, multimodal, scalars);
}
void launch_clip_adamw(const AdamArgs& a, cudaStream_t s) {
  const int nblk = 1024;
  launch_pdl(sumsq_kernel, dim3(nblk), dim3(256), 0, s, a.g, a.n, a.grad_scale, a.partials);
  launch_pdl(clip_coef_kernel, dim3(1), dim3(256), 0, s, a.partials, nblk, a.max_norm, a.scalars);
  // scalar factors exactly as torch.optim.adam._single_tensor_adam computes them (python doubles -> fp32)
  const double bc1 = 1.0 - pow(a.beta1, (double)a.step), bc2 = 1.0 - pow(a.beta2, (double)a.step);
  const int sc = a.step_cls > 0 ? a.step_cls : 1;
  const double bc1c = 1.0 - pow(a.beta1, (double)sc), bc2c = 1.0 - pow(a.beta2, (double)sc);
  AdamCoefs k{(float)(1.0 - a.lr * a.wd), (float)(a.lr / bc1), (float)sqrt(bc2), (float)(a.lr / bc1c), (float)sqrt(bc2c),
              (float)(1.0 - a.beta1), (float)(1.0 - a.beta2), (float)a.beta2, (float)a.eps};
  launch_pdl(adamw_kernel, dim3(ew_grid(a.n / 4 + 1)), dim3(256), 0, s, a, k);
}
void launch_preprocess(const double* raw, int width, const int64_t* index, int B, int size, int take_log, float* out,
                       cudaStream_t s) {
  preprocess_kernel<<<(B * size + 255) / 256, 256, 0, s>>>(raw, width, index, B, size, take_log, out);
}
void launch_io_copy(const IoCopy& c, cudaStream_t s) {
  if (c.n <= 0) return;
  io_copy_kernel<<<dim3(32, c.n), 256, 0, s>>>(c);
}
void launch_refresh_wt(const WtEntry* table_dev, int n, const float* params, float* ws, cudaStream_t s) {
  dim3 grid(64, 1, n);
  refresh_wt_kernel<<<grid, 256, 0, s>>>(table_dev, params, ws);
}

}  // namespace hp
