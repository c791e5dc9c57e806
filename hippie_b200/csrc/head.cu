// Latent head of the cVAE: condition embeddings, fusion MLP (or the unimodal encoder_fc), z_mean /
// z_log_var, reparameterisation, KL term and the decoder_fc MLPs -- forward and backward.
//
// Reference: MultiModalCVAE.encode / reparameterize / decode (hippie/model.py:397-422),
//            hippieUnimodalCVAE (hippie/model.py:46-72), KL term (hippie/model.py:472-474).
//
// ~3.6 K parameters and B x 50 activations.  BatchNorm1d over the batch couples every sample, so in training
// mode the kernels are cooperative launches: every phase is a grid-stride loop over (sample, feature) and the
// phases are separated by grid.sync().  Weights are read through L1, intermediates live in a small global
// scratch that the backward pass re-reads.  In eval mode samples are independent: each CTA takes a sample
// range and only block-level barriers are needed.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace hp {

struct HeadScratch {
  int64_t cat, f0, f1, e0, enc, mu, lv, zc, g0[2], g1[2], stats;
  int64_t dg1, dg0, dzc, dmu, dlv, denc, de0, df1, df0, dcat, total;
};

__host__ __device__ inline HeadScratch head_layout(int z, int h, int B) {
  HeadScratch L;
  const int64_t Z2 = 2 * z, D0 = 2 * Z2 + 2 * h, DZ = z + 2 * h, Bn = B;
  int64_t o = 0;
  auto take = [&](int64_t n) {
    int64_t r = o;
    o += (n + 3) & ~(int64_t)3;
    return r;
  };
  L.cat = take(Bn * D0), L.f0 = take(Bn * Z2), L.f1 = take(Bn * Z2), L.e0 = take(Bn * z), L.enc = take(Bn * z);
  L.mu = take(Bn * z), L.lv = take(Bn * z), L.zc = take(Bn * DZ);
  for (int m = 0; m < 2; ++m) L.g0[m] = take(Bn * Z2), L.g1[m] = take(Bn * Z2);
  L.stats = take(2 * (Z2 + z + 2 * Z2));
  L.dg1 = take(Bn * Z2), L.dg0 = take(Bn * Z2), L.dzc = take(Bn * DZ), L.dmu = take(Bn * z), L.dlv = take(Bn * z);
  L.denc = take(Bn * z), L.de0 = take(Bn * z), L.df1 = take(Bn * Z2), L.df0 = take(Bn * Z2), L.dcat = take(Bn * D0);
  L.total = o;
  return L;
}

int64_t head_scratch_floats(int z, int h, int B) { return head_layout(z, h, B).total; }

namespace {

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Tc {  // thread coordinates of a phase: grid-wide in cooperative (training) mode, block-wide otherwise
  int t0, ts;
  bool coop;
};
__device__ __forceinline__ void phase_sync(const Tc& tc) {
  if (tc.coop)
    cg::this_grid().sync();
  else
    __syncthreads();
}

// out[b][j] = bias[j] + sum_i in[b][i] * W[j][i]   (optionally LeakyReLU)
__device__ void ph_linear(const Tc& tc, const float* in, int ldi, const float* __restrict__ W,
                          const float* __restrict__ bias, float* out, int ldo, int nin, int nout, int b_lo, int b_hi,
                          float slope) {
  const int total = (b_hi - b_lo) * nout;
  for (int idx = tc.t0; idx < total; idx += tc.ts) {
    const int b = b_lo + idx / nout, j = idx % nout;
    const float* x = in + (int64_t)b * ldi;
    const float* w = W + (int64_t)j * nin;
    float s = 0.f;
    for (int i = 0; i < nin; ++i) s = fmaf(x[i], w[i], s);
    s += bias[j];
    out[(int64_t)b * ldo + j] = slope >= 0.f ? lrelu(s, slope) : s;
  }
}

// BatchNorm1d over the batch (training), one warp per feature, + LeakyReLU
__device__ void ph_bn_train(const Tc& tc, const float* x, int ldx, int F, int B, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float* stat, float* rm, float* rv, int64_t* cnt,
                            float* out, int ldo, float slope) {
  const int warp = tc.t0 >> 5, lane = threadIdx.x & 31, nw = tc.ts >> 5;
  for (int j = warp; j < F; j += nw) {
    float s = 0.f;
    for (int b = lane; b < B; b += 32) s += x[(int64_t)b * ldx + j];
    const float mean = warp_sum(s) / (float)B;
    float q = 0.f;
    for (int b = lane; b < B; b += 32) {
      float d = x[(int64_t)b * ldx + j] - mean;
      q = fmaf(d, d, q);
    }
    q = warp_sum(q);
    const float invstd = 1.f / sqrtf(q / (float)B + kBnEps);
    if (lane == 0) {
      stat[j] = mean, stat[F + j] = invstd;
      rm[j] = (1.f - kBnMomentum) * rm[j] + kBnMomentum * mean;
      rv[j] = (1.f - kBnMomentum) * rv[j] + kBnMomentum * (q / (float)max(B - 1, 1));
    }
    const float g = gamma[j] * invstd, be = beta[j];
    for (int b = lane; b < B; b += 32)
      out[(int64_t)b * ldo + j] = lrelu(fmaf(x[(int64_t)b * ldx + j] - mean, g, be), slope);
  }
  if (tc.t0 == 0) *cnt += 1;
}

__device__ void ph_bn_eval(const float* x, int ldx, int F, int b_lo, int b_hi, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float* rm, const float* rv, float* out, int ldo,
                           float slope) {
  const int total = (b_hi - b_lo) * F;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int b = b_lo + idx / F, j = idx % F;
    const float invstd = 1.f / sqrtf(rv[j] + kBnEps);
    out[(int64_t)b * ldo + j] = lrelu(fmaf(x[(int64_t)b * ldx + j] - rm[j], gamma[j] * invstd, beta[j]), slope);
  }
}

// dx[b][i] (=|+=) (sum_j dy[b][j] * W[j][i]) * lrelu'(mask[b][i])
__device__ void ph_dgrad(const Tc& tc, const float* dy, int ldy, const float* __restrict__ W, int nin, int nout,
                         float* dx, int ldx, const float* mask, int ldm, float slope, bool acc, int B) {
  const int total = B * nin;
  for (int idx = tc.t0; idx < total; idx += tc.ts) {
    const int b = idx / nin, i = idx % nin;
    const float* g = dy + (int64_t)b * ldy;
    float s = 0.f;
    for (int j = 0; j < nout; ++j) s = fmaf(g[j], W[(int64_t)j * nin + i], s);
    if (mask) s *= mask[(int64_t)b * ldm + i] > 0.f ? 1.f : slope;
    float* d = dx + (int64_t)b * ldx + i;
    *d = acc ? *d + s : s;
  }
}

// dW[j][i] = sum_b dy[b][j] * x[b][i];  db[j] = sum_b dy[b][j]     one warp per output, lanes over the batch
__device__ void ph_wgrad(const Tc& tc, const float* dy, int ldy, const float* x, int ldx, int nin, int nout, float* dW,
                         float* db, int B) {
  const int total = nout * (nin + 1);
  const int warp = tc.t0 >> 5, lane = threadIdx.x & 31, nw = tc.ts >> 5;
  for (int idx = warp; idx < total; idx += nw) {
    const int j = idx / (nin + 1), i = idx % (nin + 1);
    float s = 0.f;
    if (i < nin) {
      for (int b = lane; b < B; b += 32) s = fmaf(dy[(int64_t)b * ldy + j], x[(int64_t)b * ldx + i], s);
    } else {
      for (int b = lane; b < B; b += 32) s += dy[(int64_t)b * ldy + j];
    }
    s = warp_sum(s);
    if (lane == 0) {
      if (i < nin)
        dW[(int64_t)j * nin + i] = s;
      else
        db[j] = s;
    }
  }
}

// backward of y = lrelu(bn(x)) over the batch; g is the gradient w.r.t. y
__device__ void ph_bn_bwd(const Tc& tc, const float* g, int ldg, const float* y, int ldy, const float* x, int ldx,
                          const float* stat, int F, const float* __restrict__ gamma, float* dgamma, float* dbeta,
                          float* dx, int lddx, float slope, int B) {
  const int warp = tc.t0 >> 5, lane = threadIdx.x & 31, nw = tc.ts >> 5;
  for (int j = warp; j < F; j += nw) {
    const float mean = stat[j], invstd = stat[F + j];
    float s1 = 0.f, s2 = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float gp = g[(int64_t)b * ldg + j] * (y[(int64_t)b * ldy + j] > 0.f ? 1.f : slope);
      s1 += gp;
      s2 = fmaf(gp, (x[(int64_t)b * ldx + j] - mean) * invstd, s2);
    }
    s1 = warp_sum(s1), s2 = warp_sum(s2);
    if (lane == 0) dgamma[j] = s2, dbeta[j] = s1;
    const float k = gamma[j] * invstd, m1 = s1 / (float)B, m2 = s2 / (float)B;
    for (int b = lane; b < B; b += 32) {
      const float gp = g[(int64_t)b * ldg + j] * (y[(int64_t)b * ldy + j] > 0.f ? 1.f : slope);
      dx[(int64_t)b * lddx + j] = k * (gp - m1 - (x[(int64_t)b * ldx + j] - mean) * invstd * m2);
    }
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) head_fwd_kernel(HeadArgs a) {
  const HeadScratch L = head_layout(a.z, a.h, a.B);
  float* S = a.scratch;
  const HeadParams& P = a.hp;
  const float* W = a.params;
  const int z = a.z, h = a.h, Z2 = 2 * z, E = a.n_enc * Z2, D0 = E + 2 * h, DZ = z + 2 * h;
  // training: cooperative grid over the whole batch;  eval: each CTA owns a sample range
  Tc tc;
  tc.coop = a.train != 0 && gridDim.x > 1;  // one CTA (small batches): block barriers instead of grid barriers
  tc.t0 = tc.coop ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : (int)threadIdx.x;
  tc.ts = tc.coop ? (int)(gridDim.x * blockDim.x) : (int)blockDim.x;
  const int per = (a.B + gridDim.x - 1) / gridDim.x;
  const bool whole = a.train != 0;  // training statistics couple the batch: every CTA sees all samples
  const int b_lo = whole ? 0 : min(a.B, (int)blockIdx.x * per), b_hi = whole ? a.B : min(a.B, b_lo + per);
  const int nb = b_hi - b_lo;
  __shared__ float sred[32];

  // cat = [h1, (h2,) source_emb, class_emb]   (hippie/model.py:405-406, 425-426)
  for (int idx = tc.t0; idx < nb * D0; idx += tc.ts) {
    const int b = b_lo + idx / D0, j = idx % D0;
    float v;
    if (j < E)
      v = a.hin[j / Z2][(int64_t)b * Z2 + (j % Z2)];
    else if (j < E + h)
      v = W[P.src_emb + a.src[b] * h + (j - E)];
    else
      v = a.cls ? W[P.cls_emb + a.cls[b] * h + (j - E - h)] : 0.f;
    S[L.cat + (int64_t)b * D0 + j] = v;
  }
  phase_sync(tc);
  ph_linear(tc, S + L.cat, D0, W + P.f0_w, W + P.f0_b, S + L.f0, Z2, D0, Z2, b_lo, b_hi, -1.f);
  phase_sync(tc);
  if (a.train)
    ph_bn_train(tc, S + L.f0, Z2, Z2, a.B, W + P.fbn_g, W + P.fbn_b, S + L.stats, a.run_mean + P.fbn_run,
                a.run_var + P.fbn_run, a.run_count + P.fbn_cnt, S + L.f1, Z2, kSlopeHead);
  else
    ph_bn_eval(S + L.f0, Z2, Z2, b_lo, b_hi, W + P.fbn_g, W + P.fbn_b, a.run_mean + P.fbn_run, a.run_var + P.fbn_run,
               S + L.f1, Z2, kSlopeHead);
  phase_sync(tc);
  ph_linear(tc, S + L.f1, Z2, W + P.f3_w, W + P.f3_b, S + L.e0, z, Z2, z, b_lo, b_hi, -1.f);
  phase_sync(tc);
  if (P.ebn_g >= 0) {  // unimodal encoder_fc ends with BatchNorm1d(z) + LeakyReLU(0.2)  (hippie/model.py:21-28)
    if (a.train)
      ph_bn_train(tc, S + L.e0, z, z, a.B, W + P.ebn_g, W + P.ebn_b, S + L.stats + 2 * Z2, a.run_mean + P.ebn_run,
                  a.run_var + P.ebn_run, a.run_count + P.ebn_cnt, S + L.enc, z, kSlopeHead);
    else
      ph_bn_eval(S + L.e0, z, z, b_lo, b_hi, W + P.ebn_g, W + P.ebn_b, a.run_mean + P.ebn_run, a.run_var + P.ebn_run,
                 S + L.enc, z, kSlopeHead);
  } else {
    for (int idx = tc.t0; idx < nb * z; idx += tc.ts)
      S[L.enc + (int64_t)b_lo * z + idx] = S[L.e0 + (int64_t)b_lo * z + idx];
  }
  phase_sync(tc);

  // mu, logvar, z = mu + eps * exp(0.5 logvar), KL   (hippie/model.py:397-400, 408, 472)
  float klp = 0.f;
  for (int idx = tc.t0; idx < nb * z; idx += tc.ts) {
    const int b = b_lo + idx / z, i = idx % z;
    const float* e = S + L.enc + (int64_t)b * z;
    float m = 0.f, v = 0.f;
    for (int k = 0; k < z; ++k) {
      m = fmaf(e[k], W[P.zm_w + i * z + k], m);
      v = fmaf(e[k], W[P.zv_w + i * z + k], v);
    }
    m += W[P.zm_b + i], v += W[P.zv_b + i];
    S[L.mu + (int64_t)b * z + i] = m, S[L.lv + (int64_t)b * z + i] = v;
    if (a.out_mu) a.out_mu[(int64_t)b * z + i] = m;
    if (a.out_logvar) a.out_logvar[(int64_t)b * z + i] = v;
    const float ev = expf(v);
    klp += -0.5f * (1.f + v - m * m - ev);
    if (a.decode) {
      const float std = expf(0.5f * v);
      S[L.zc + (int64_t)b * DZ + i] = a.eps ? fmaf(a.eps[(int64_t)b * z + i], std, m) : m;
    }
  }
  if (a.out_enc) {
    if (a.zscore_ddof < 0) {
      for (int idx = tc.t0; idx < nb * z; idx += tc.ts)
        a.out_enc[(int64_t)b_lo * z + idx] = S[L.enc + (int64_t)b_lo * z + idx];
    } else {  // per-row z-score (scripts/train_model_with_multimodal.py:31 ddof 0; scripts/utils.py:87-88 ddof 1)
      for (int bb = tc.t0; bb < nb; bb += tc.ts) {
        const float* e = S + L.enc + (int64_t)(b_lo + bb) * z;
        float s = 0.f;
        for (int k = 0; k < z; ++k) s += e[k];
        const float mean = s / (float)z;
        float q = 0.f;
        for (int k = 0; k < z; ++k) q += (e[k] - mean) * (e[k] - mean);
        const float sd = sqrtf(q / (float)(z - a.zscore_ddof));
        for (int k = 0; k < z; ++k) a.out_enc[(int64_t)(b_lo + bb) * z + k] = (e[k] - mean) / sd;
      }
    }
  }
  klp = warp_sum(klp);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = klp;
  __syncthreads();
  if (threadIdx.x == 0 && a.kl_sum) {  // per-CTA partial; loss_finalize adds them in a fixed order (deterministic)
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
    a.kl_sum[blockIdx.x] = s;
  }
  if (!a.decode) return;

  // zc = [z, source_emb, class_emb]   (hippie/model.py:412-413)
  for (int idx = tc.t0; idx < nb * 2 * h; idx += tc.ts) {
    const int b = b_lo + idx / (2 * h), j = idx % (2 * h);
    S[L.zc + (int64_t)b * DZ + z + j] = S[L.cat + (int64_t)b * D0 + E + j];
  }
  phase_sync(tc);
  for (int m = 0; m < a.n_dec; ++m) {  // decoder_fc: Linear, LeakyReLU(.2), Linear, BatchNorm1d, LeakyReLU(.2)
    ph_linear(tc, S + L.zc, DZ, W + P.d0_w[m], W + P.d0_b[m], S + L.g0[m], Z2, DZ, Z2, b_lo, b_hi, kSlopeHead);
  }
  phase_sync(tc);
  for (int m = 0; m < a.n_dec; ++m)
    ph_linear(tc, S + L.g0[m], Z2, W + P.d2_w[m], W + P.d2_b[m], S + L.g1[m], Z2, Z2, Z2, b_lo, b_hi, -1.f);
  phase_sync(tc);
  for (int m = 0; m < a.n_dec; ++m) {
    if (a.train)
      ph_bn_train(tc, S + L.g1[m], Z2, Z2, a.B, W + P.dbn_g[m], W + P.dbn_b[m],
                  S + L.stats + 2 * (Z2 + z) + m * 2 * Z2, a.run_mean + P.dbn_run[m], a.run_var + P.dbn_run[m],
                  a.run_count + P.dbn_cnt[m], a.dout[m], Z2, kSlopeHead);
    else
      ph_bn_eval(S + L.g1[m], Z2, Z2, b_lo, b_hi, W + P.dbn_g[m], W + P.dbn_b[m], a.run_mean + P.dbn_run[m],
                 a.run_var + P.dbn_run[m], a.dout[m], Z2, kSlopeHead);
  }
}

// Cooperative launch; the two decoder_fc branches use separate gradient scratch (dg1/dg0 per branch live in the
// df1/df0 and dg1/dg0 slots) so that both run inside the same phases.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) head_bwd_kernel(HeadArgs a) {
  const HeadScratch L = head_layout(a.z, a.h, a.B);
  float* S = a.scratch;
  const HeadParams& P = a.hp;
  const float* W = a.params;
  float* G = a.grads;
  const int z = a.z, h = a.h, Z2 = 2 * z, E = a.n_enc * Z2, D0 = E + 2 * h, DZ = z + 2 * h, B = a.B;
  Tc tc;
  tc.coop = gridDim.x > 1;
  tc.t0 = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  tc.ts = (int)(gridDim.x * blockDim.x);
  // per-branch scratch: branch 0 uses (dg1, dg0), branch 1 borrows (df1, df0), which are not live yet
  const int64_t dg1o[2] = {L.dg1, L.df1}, dg0o[2] = {L.dg0, L.df0};

  for (int m = 0; m < a.n_dec; ++m) {
    const float* stat = S + L.stats + 2 * (Z2 + z) + m * 2 * Z2;
    ph_bn_bwd(tc, a.dd[m], Z2, a.dout[m], Z2, S + L.g1[m], Z2, stat, Z2, W + P.dbn_g[m], G + P.dbn_g[m],
              G + P.dbn_b[m], S + dg1o[m], Z2, kSlopeHead, B);
  }
  phase_sync(tc);
  for (int m = 0; m < a.n_dec; ++m) {
    ph_wgrad(tc, S + dg1o[m], Z2, S + L.g0[m], Z2, Z2, Z2, G + P.d2_w[m], G + P.d2_b[m], B);
    ph_dgrad(tc, S + dg1o[m], Z2, W + P.d2_w[m], Z2, Z2, S + dg0o[m], Z2, S + L.g0[m], Z2, kSlopeHead, false, B);
  }
  phase_sync(tc);
  for (int m = 0; m < a.n_dec; ++m) ph_wgrad(tc, S + dg0o[m], Z2, S + L.zc, DZ, DZ, Z2, G + P.d0_w[m], G + P.d0_b[m], B);
  ph_dgrad(tc, S + dg0o[0], Z2, W + P.d0_w[0], DZ, Z2, S + L.dzc, DZ, nullptr, 0, 0.f, false, B);
  phase_sync(tc);
  if (a.n_dec > 1) {
    ph_dgrad(tc, S + dg0o[1], Z2, W + P.d0_w[1], DZ, Z2, S + L.dzc, DZ, nullptr, 0, 0.f, true, B);
    phase_sync(tc);
  }
  // reparameterisation + KL  (hippie/model.py:397-400, 472-474): total = ... + beta * mean_b(kl_b)
  const float kscale = a.beta / (float)B;
  for (int idx = tc.t0; idx < B * z; idx += tc.ts) {
    const int b = idx / z, i = idx % z;
    const float dz = S[L.dzc + (int64_t)b * DZ + i];
    const float mu = S[L.mu + idx], lv = S[L.lv + idx];
    const float std = expf(0.5f * lv);
    S[L.dmu + idx] = dz + kscale * mu;
    S[L.dlv + idx] = dz * a.eps[idx] * 0.5f * std + kscale * 0.5f * (expf(lv) - 1.f);
  }
  phase_sync(tc);
  ph_wgrad(tc, S + L.dmu, z, S + L.enc, z, z, z, G + P.zm_w, G + P.zm_b, B);
  ph_wgrad(tc, S + L.dlv, z, S + L.enc, z, z, z, G + P.zv_w, G + P.zv_b, B);
  // denc = dmu * Wm + dlv * Wv in one pass
  for (int idx = tc.t0; idx < B * z; idx += tc.ts) {
    const int b = idx / z, i = idx % z;
    float s = 0.f;
    for (int j = 0; j < z; ++j) {
      s = fmaf(S[L.dmu + (int64_t)b * z + j], W[P.zm_w + j * z + i], s);
      s = fmaf(S[L.dlv + (int64_t)b * z + j], W[P.zv_w + j * z + i], s);
    }
    S[L.denc + idx] = s;
  }
  phase_sync(tc);
  const float* de0 = S + L.denc;
  if (P.ebn_g >= 0) {
    ph_bn_bwd(tc, S + L.denc, z, S + L.enc, z, S + L.e0, z, S + L.stats + 2 * Z2, z, W + P.ebn_g, G + P.ebn_g,
              G + P.ebn_b, S + L.de0, z, kSlopeHead, B);
    de0 = S + L.de0;
    phase_sync(tc);
  }
  ph_wgrad(tc, de0, z, S + L.f1, Z2, Z2, z, G + P.f3_w, G + P.f3_b, B);
  ph_dgrad(tc, de0, z, W + P.f3_w, Z2, z, S + L.df1, Z2, nullptr, 0, 0.f, false, B);
  phase_sync(tc);
  ph_bn_bwd(tc, S + L.df1, Z2, S + L.f1, Z2, S + L.f0, Z2, S + L.stats, Z2, W + P.fbn_g, G + P.fbn_g, G + P.fbn_b,
            S + L.df0, Z2, kSlopeHead, B);
  phase_sync(tc);
  ph_wgrad(tc, S + L.df0, Z2, S + L.cat, D0, D0, Z2, G + P.f0_w, G + P.f0_b, B);
  ph_dgrad(tc, S + L.df0, Z2, W + P.f0_w, D0, Z2, S + L.dcat, D0, nullptr, 0, 0.f, false, B);
  phase_sync(tc);
  for (int e = 0; e < a.n_enc; ++e)
    for (int idx = tc.t0; idx < B * Z2; idx += tc.ts)
      a.dh[e][idx] = S[L.dcat + (int64_t)(idx / Z2) * D0 + e * Z2 + (idx % Z2)];
  // embedding gradients (scatter-add of the two places each embedding is used): one warp per table entry
  {
    const int warp = tc.t0 >> 5, lane = threadIdx.x & 31, nw = tc.ts >> 5;
    const int n_src = a.num_sources * h, n_cls = a.cls ? a.num_classes * h : 0;
    for (int idx = warp; idx < n_src + n_cls; idx += nw) {
      const bool is_cls = idx >= n_src;
      const int e = is_cls ? idx - n_src : idx;
      const int s = e / h, k = e % h;
      const int64_t* lab = is_cls ? a.cls : a.src;
      const int off_cat = E + (is_cls ? h : 0) + k, off_zc = z + (is_cls ? h : 0) + k;
      float acc = 0.f;
      for (int b = lane; b < B; b += 32)
        if (lab[b] == s) acc += S[L.dcat + (int64_t)b * D0 + off_cat] + S[L.dzc + (int64_t)b * DZ + off_zc];
      acc = warp_sum(acc);
      if (lane == 0) G[(is_cls ? P.cls_emb : P.src_emb) + e] = acc;
    }
  }
}

}  // namespace

static int head_grid(int B, int z) {
  // the phases are short grid-stride loops separated by grid barriers: more CTAs shorten the loops (measured at bs512:
  // 16 CTAs 3.70 ms per step, 40 CTAs 3.61, 148 CTAs 3.58); one CTA per SM keeps the cooperative launch co-resident
  int g = (B * 2 * z + 63) / 64;
  if (g > 148) g = 148;
  if (g < 1) g = 1;
  return g;
}

// Up to this many (sample, feature) items the whole head runs in ONE 1024-thread CTA: ~10 phases separated by block
// barriers (~0.1 us) instead of grid barriers of a cooperative launch (~4 us each).  Measured: bs64 step 1.99 -> 1.93 ms;
// at bs512 (10 K items) one CTA is far too slow (3.6 -> 4.6 ms), so the threshold sits at the small-batch case.
constexpr int kHeadSingleCtaItems = 128 * 2 * 10;

int launch_head_fwd(const HeadArgs& a, cudaStream_t s) {
  if (a.train) {
    void* args[] = {const_cast<HeadArgs*>(&a)};
    if (a.B * 2 * a.z <= kHeadSingleCtaItems) {
      head_fwd_kernel<1024><<<1, 1024, 0, s>>>(a);
      return 1;
    }
    const int grid = head_grid(a.B, a.z);
    cudaLaunchCooperativeKernel((const void*)head_fwd_kernel<256>, dim3(grid), dim3(256), args, 0, s);
    return grid;
  }
  int grid = (a.B + 63) / 64;
  if (grid > kHeadMaxCtas) grid = kHeadMaxCtas;
  if (grid < 1) grid = 1;
  head_fwd_kernel<256><<<grid, 256, 0, s>>>(a);
  return grid;
}
void launch_head_bwd(const HeadArgs& a, cudaStream_t s) {
  void* args[] = {const_cast<HeadArgs*>(&a)};
  if (a.B * 2 * a.z <= kHeadSingleCtaItems) {
    head_bwd_kernel<1024><<<1, 1024, 0, s>>>(a);
    return;
  }
  cudaLaunchCooperativeKernel((const void*)head_bwd_kernel<256>, dim3(head_grid(a.B, a.z)), dim3(256), args, 0, s);
}

}  // namespace hp
