// Latent head of the cVAE: condition embeddings, fusion MLP (or the unimodal encoder_fc), z_mean /
// z_log_var, reparameterisation, KL term and the decoder_fc MLPs -- forward and backward.
//
// Reference: MultiModalCVAE.encode / reparameterize / decode (hippie/model.py:397-422),
//            hippieUnimodalCVAE (hippie/model.py:46-72), KL term (hippie/model.py:472-474).
//
// ~3.6 K parameters and B x 50 activations.  Every CTA owns a range of samples and walks the whole per-sample chain
// (gather -> Linear -> ... ) for them with block barriers only.  BatchNorm1d over the batch is the one place where samples
// couple: there every CTA publishes the moments of its samples, ONE grid barrier follows (cooperative launch in
// training mode), and every CTA combines all partials itself (in double).  Forward: 2 grid barriers (3 unimodal);
// backward: 2 (3 unimodal).  The weight gradients are batch reductions: every CTA adds the outer products of its own
// samples into the zeroed gradient buffer with atomics.  In eval mode samples are independent and no grid barrier is
// needed.  (The first version separated every phase by grid.sync(): ~21 barriers of ~3.4 us, 51 + 100 us per bs512 step
// against 42 + 47 us now)
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace hp {

struct HeadScratch {
  int64_t cat, f0, f1, e0, enc, mu, lv, zc, g0[2], g1[2], stats;
  int64_t dg1, dg0, dzc, dmu, dlv, denc, de0, df1, df0, dcat, dg1b, dg0b, bnpart, total;
};
constexpr int kHeadTrainCtas = 148;  // cooperative grid: at most one CTA per SM

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
  L.dg1b = take(Bn * Z2), L.dg0b = take(Bn * Z2);
  L.bnpart = take((int64_t)4 * kHeadTrainCtas * Z2 * 4);  // four BatchNorm layers x [CTA][feature] float4 moments
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

// HIPPIE_B200_HEAD_STAMPS=1 (debugging, eager mode): CTA 0 records %globaltimer at the phase boundaries
__device__ unsigned long long g_head_stamps[64];
__device__ int g_head_stamps_on = 0;
__device__ __forceinline__ void hstamp(int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && g_head_stamps_on) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_head_stamps[i] = t;
  }
}

struct Tc {  // thread coordinates of a phase (block-wide: every CTA works on its own samples)
  int t0, ts;
};
__device__ __forceinline__ void grid_sync() {
  if (gridDim.x > 1)
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

constexpr int kHeadMaxF = 512;  // 2 * z_dim <= 512 (hippie_create)

// (count, mean, centred sum of squares) merged with Chan's formula
__device__ __forceinline__ void mom_merge(double& n, double& mean, double& m2, double nb, double meanb, double m2b) {
  if (nb == 0.0) return;
  const double nt = n + nb, d = meanb - mean;
  mean += d * (nb / nt);
  m2 += m2b + d * d * (n * nb / nt);
  n = nt;
}

// BatchNorm1d over the batch (training), step 1: moments of this CTA's samples -> part[feature] = (n, mean, M2, -)
template <int THREADS>
__device__ void bn_partial(const float* x, int ldx, int F, int b_lo, int b_hi, float* part, double (*s_mom)[3]) {
  const int Fp = min(F, THREADS), nsl = THREADS / Fp;
  const int jj = threadIdx.x % Fp, sl = threadIdx.x / Fp;
  for (int jb = 0; jb < F; jb += Fp) {
    const int j = jb + jj;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    if (sl < nsl && j < F) {
      float sum = 0.f;
      int cnt = 0;
      for (int b = b_lo + sl; b < b_hi; b += nsl) sum += x[(int64_t)b * ldx + j], ++cnt;
      if (cnt) {
        const float mu = sum / (float)cnt;
        float q = 0.f;
        for (int b = b_lo + sl; b < b_hi; b += nsl) {
          const float d = x[(int64_t)b * ldx + j] - mu;
          q = fmaf(d, d, q);
        }
        n = cnt, mean = mu, m2 = q;
      }
    }
    s_mom[threadIdx.x][0] = n, s_mom[threadIdx.x][1] = mean, s_mom[threadIdx.x][2] = m2;
    __syncthreads();
    if (sl == 0 && j < F) {
      for (int k = 1; k < nsl; ++k) mom_merge(n, mean, m2, s_mom[k * Fp + jj][0], s_mom[k * Fp + jj][1], s_mom[k * Fp + jj][2]);
      reinterpret_cast<float4*>(part)[j] = make_float4((float)n, (float)mean, (float)m2, 0.f);
    }
    __syncthreads();
  }
}

// step 2 (after the grid barrier): combine the partials of all CTAs -- total count and sum first, then
// M2 = sum_c (M2_c + n_c (mean_c - mean)^2); no divisions inside the loops.  s_stat[j] = mean, s_stat[F + j] = invstd;
// CTA 0 publishes them for the backward pass and updates the running statistics (momentum 0.1, unbiased variance)
template <int THREADS>
__device__ void bn_finish(const float* part, int ncta, int F, float* stat, float* rm, float* rv, int64_t* cnt,
                          double (*s_mom)[3], float* s_stat) {
  const int Fp = min(F, THREADS), nsl = THREADS / Fp;
  const int jj = threadIdx.x % Fp, sl = threadIdx.x / Fp;
  for (int jb = 0; jb < F; jb += Fp) {
    const int j = jb + jj;
    const bool live = sl < nsl && j < F;
    // the thread's partials stay in registers between the two passes when they fit (one L2 round trip)
    constexpr int KP = 16;
    const bool in_regs = (ncta + nsl - 1) / nsl <= KP;
    float4 pv[KP];
    double n = 0.0, sum = 0.0;
    if (live) {
      if (in_regs) {
#pragma unroll
        for (int u = 0; u < KP; ++u) {
          const int c = sl + u * nsl;
          int cr = c + (int)blockIdx.x;  // rotated per reader: all CTAs read the same partials at the same time
          cr = cr >= ncta ? cr - ncta : cr;
          pv[u] = c < ncta ? __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)cr * F + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < KP; ++u) n += (double)pv[u].x, sum += (double)pv[u].x * (double)pv[u].y;
      } else {
        for (int c = sl; c < ncta; c += nsl) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)c * F + j);
          n += (double)v.x, sum += (double)v.x * (double)v.y;
        }
      }
    }
    s_mom[threadIdx.x][0] = n, s_mom[threadIdx.x][1] = sum;
    __syncthreads();
    n = 0.0, sum = 0.0;
    for (int k = 0; k < nsl; ++k) n += s_mom[k * Fp + jj][0], sum += s_mom[k * Fp + jj][1];
    const double mean = sum / n;
    double m2 = 0.0;
    if (live) {
      if (in_regs) {
#pragma unroll
        for (int u = 0; u < KP; ++u) {
          const double d = (double)pv[u].y - mean;
          m2 += (double)pv[u].z + (double)pv[u].x * d * d;
        }
      } else {
        for (int c = sl; c < ncta; c += nsl) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)c * F + j);
          const double d = (double)v.y - mean;
          m2 += (double)v.z + (double)v.x * d * d;
        }
      }
    }
    s_mom[threadIdx.x][2] = m2;
    __syncthreads();
    if (sl == 0 && j < F) {
      for (int k = 1; k < nsl; ++k) m2 += s_mom[k * Fp + jj][2];
      const float var = (float)(m2 / n);
      const float invstd = 1.f / sqrtf(var + kBnEps);
      s_stat[j] = (float)mean, s_stat[F + j] = invstd;
      if (blockIdx.x == 0) {
        stat[j] = (float)mean, stat[F + j] = invstd;
        rm[j] = (1.f - kBnMomentum) * rm[j] + kBnMomentum * (float)mean;
        rv[j] = (1.f - kBnMomentum) * rv[j] + kBnMomentum * (float)(m2 / fmax(n - 1.0, 1.0));
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cnt += 1;
}

// step 3: normalise + LeakyReLU for the CTA's samples
__device__ void bn_apply_own(const Tc& tc, const float* x, int ldx, int F, int b_lo, int b_hi,
                             const float* __restrict__ gamma, const float* __restrict__ beta, const float* s_stat,
                             float* out, int ldo, float slope) {
  const int total = (b_hi - b_lo) * F;
  for (int idx = tc.t0; idx < total; idx += tc.ts) {
    const int b = b_lo + idx / F, j = idx % F;
    out[(int64_t)b * ldo + j] = lrelu(fmaf(x[(int64_t)b * ldx + j] - s_stat[j], gamma[j] * s_stat[F + j], beta[j]), slope);
  }
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

// dx[b][i] (=|+=) (sum_j dy[b][j] * W[j][i]) * lrelu'(mask[b][i])   for the CTA's samples
__device__ void ph_dgrad(const Tc& tc, const float* dy, int ldy, const float* __restrict__ W, int nin, int nout,
                         float* dx, int ldx, const float* mask, int ldm, float slope, bool acc, int b_lo, int b_hi) {
  const int total = (b_hi - b_lo) * nin;
  for (int idx = tc.t0; idx < total; idx += tc.ts) {
    const int b = b_lo + idx / nin, i = idx % nin;
    const float* g = dy + (int64_t)b * ldy;
    float s = 0.f;
    for (int j = 0; j < nout; ++j) s = fmaf(g[j], W[(int64_t)j * nin + i], s);
    if (mask) s *= mask[(int64_t)b * ldm + i] > 0.f ? 1.f : slope;
    float* d = dx + (int64_t)b * ldx + i;
    *d = acc ? *d + s : s;
  }
}

// Weight gradients of every Linear layer of the head: dW[j][i] = sum_b dy[b][j] * x[b][i], db[j] = sum_b dy[b][j].
// Every CTA adds the contribution of ITS samples to the (zeroed) gradient buffer with atomics, one thread per output
// element over the flattened list of all layers -- no grid barrier and no pass over other CTAs' data.  (A grid-wide
// reduction pass, one warp per output with lanes over the batch, took 26-42 us at bs512: hundreds of warps walk the
// same few rows.)  The CTAs start at different outputs so that their atomics do not queue on the same addresses.
struct WgJob {
  const float* dy;
  const float* x;
  float* dW;
  float* db;
  int ldy, ldx, nin, nout;
};
__device__ void ph_wgrad_own(const Tc& tc, const WgJob* jobs, int njobs, int b_lo, int b_hi) {
  int total = 0;
  for (int k = 0; k < njobs; ++k) total += jobs[k].nout * (jobs[k].nin + 1);
  const int rot = (int)((blockIdx.x * 997u) % (unsigned)total);
  constexpr int Q = 4;  // outputs per thread and pass: their loads are in flight together
  for (int t = tc.t0; t < total; t += Q * tc.ts) {
    const float* pdy[Q];
    const float* px[Q];
    float* dst[Q];
    int sdy[Q], sx[Q];
    float acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int tq = t + q * tc.ts;
      int r = tq + rot;
      r = r >= total ? r - total : r;
      r = tq < total ? r : 0;
      int k = 0;
      while (r >= jobs[k].nout * (jobs[k].nin + 1)) r -= jobs[k].nout * (jobs[k].nin + 1), ++k;
      const WgJob& J = jobs[k];
      const int j = r / (J.nin + 1), i = r % (J.nin + 1);
      const bool bias = i == J.nin;
      pdy[q] = J.dy + j, sdy[q] = J.ldy;
      px[q] = bias ? nullptr : J.x + i, sx[q] = J.ldx;
      dst[q] = tq < total ? (bias ? J.db + j : J.dW + (int64_t)j * J.nin + i) : nullptr;
      acc[q] = 0.f;
    }
    for (int b = b_lo; b < b_hi; ++b) {
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float xv = px[q] ? px[q][(int64_t)b * sx[q]] : 1.f;
        acc[q] = fmaf(pdy[q][(int64_t)b * sdy[q]], xv, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (dst[q]) atomicAdd(dst[q], acc[q]);
  }
}

// backward of y = lrelu(bn(x)) over the batch, g = gradient w.r.t. y.  Step 1: (sum g', sum g' * xhat) over the CTA's
// samples -> part[feature] = (s1, s2, -, -)
template <int THREADS>
__device__ void bnb_partial(const float* g, int ldg, const float* y, int ldy, const float* x, int ldx, const float* stat,
                            int F, int b_lo, int b_hi, float slope, float* part, double (*s_mom)[3]) {
  const int Fp = min(F, THREADS), nsl = THREADS / Fp;
  const int jj = threadIdx.x % Fp, sl = threadIdx.x / Fp;
  for (int jb = 0; jb < F; jb += Fp) {
    const int j = jb + jj;
    float s1 = 0.f, s2 = 0.f;
    if (sl < nsl && j < F) {
      const float mean = stat[j], invstd = stat[F + j];
      for (int b = b_lo + sl; b < b_hi; b += nsl) {
        const float gp = g[(int64_t)b * ldg + j] * (y[(int64_t)b * ldy + j] > 0.f ? 1.f : slope);
        s1 += gp;
        s2 = fmaf(gp, (x[(int64_t)b * ldx + j] - mean) * invstd, s2);
      }
    }
    s_mom[threadIdx.x][0] = s1, s_mom[threadIdx.x][1] = s2;
    __syncthreads();
    if (sl == 0 && j < F) {
      double t1 = s1, t2 = s2;
      for (int k = 1; k < nsl; ++k) t1 += s_mom[k * Fp + jj][0], t2 += s_mom[k * Fp + jj][1];
      reinterpret_cast<float4*>(part)[j] = make_float4((float)t1, (float)t2, 0.f, 0.f);
    }
    __syncthreads();
  }
}

// step 2 (after the grid barrier): totals over all CTAs; s_stat[j] = mean_b(g'), s_stat[F + j] = mean_b(g' * xhat);
// CTA 0 writes dgamma / dbeta
template <int THREADS>
__device__ void bnb_finish(const float* part, int ncta, int F, int B, float* dgamma, float* dbeta, double (*s_mom)[3],
                           float* s_stat) {
  const int Fp = min(F, THREADS), nsl = THREADS / Fp;
  const int jj = threadIdx.x % Fp, sl = threadIdx.x / Fp;
  for (int jb = 0; jb < F; jb += Fp) {
    const int j = jb + jj;
    double t1 = 0.0, t2 = 0.0;
    if (sl < nsl && j < F)
      for (int c = sl; c < ncta; c += nsl) {
        int cr = c + (int)blockIdx.x;
        cr = cr >= ncta ? cr - ncta : cr;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)cr * F + j);
        t1 += (double)v.x, t2 += (double)v.y;
      }
    s_mom[threadIdx.x][0] = t1, s_mom[threadIdx.x][1] = t2;
    __syncthreads();
    if (sl == 0 && j < F) {
      for (int k = 1; k < nsl; ++k) t1 += s_mom[k * Fp + jj][0], t2 += s_mom[k * Fp + jj][1];
      s_stat[j] = (float)(t1 / (double)B), s_stat[F + j] = (float)(t2 / (double)B);
      if (blockIdx.x == 0) dgamma[j] = (float)t2, dbeta[j] = (float)t1;
    }
    __syncthreads();
  }
}

// step 3: dx for the CTA's samples
__device__ void bnb_apply_own(const Tc& tc, const float* g, int ldg, const float* y, int ldy, const float* x, int ldx,
                              const float* stat, int F, int b_lo, int b_hi, const float* __restrict__ gamma,
                              const float* s_stat, float* dx, int lddx, float slope) {
  const int total = (b_hi - b_lo) * F;
  for (int idx = tc.t0; idx < total; idx += tc.ts) {
    const int b = b_lo + idx / F, j = idx % F;
    const float mean = stat[j], invstd = stat[F + j];
    const float gp = g[(int64_t)b * ldg + j] * (y[(int64_t)b * ldy + j] > 0.f ? 1.f : slope);
    dx[(int64_t)b * lddx + j] = gamma[j] * invstd * (gp - s_stat[j] - (x[(int64_t)b * ldx + j] - mean) * invstd * s_stat[F + j]);
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) head_fwd_kernel(HeadArgs a) {
  hstamp(0);
  const HeadScratch L = head_layout(a.z, a.h, a.B);
  float* S = a.scratch;
  const HeadParams& P = a.hp;
  const float* W = a.params;
  const int z = a.z, h = a.h, Z2 = 2 * z, E = a.n_enc * Z2, D0 = E + 2 * h, DZ = z + 2 * h;
  const Tc tc{(int)threadIdx.x, (int)blockDim.x};
  const int per = (a.B + gridDim.x - 1) / gridDim.x;  // every CTA owns a sample range
  const int b_lo = min(a.B, (int)blockIdx.x * per), b_hi = min(a.B, b_lo + per);
  const int nb = b_hi - b_lo;
  const int ncta = gridDim.x;
  __shared__ float sred[32];
  __shared__ double s_mom[THREADS][3];
  __shared__ float s_stat[2 * kHeadMaxF];
  float* part = S + L.bnpart;
  const int64_t part_layer = (int64_t)kHeadTrainCtas * Z2 * 4;  // one region per BatchNorm layer

  // cat = [h1, (h2,) source_emb, class_emb]   (hippie/model.py:405-406, 425-426)
  for (int idx = tc.t0; idx < nb * D0; idx += tc.ts) {
    const int b = b_lo + idx / D0, j = idx % D0;
    float v;
    if (j < E) {
      v = a.stage == kHeadDecodeOnly ? 0.f : a.hin[j / Z2][(int64_t)b * Z2 + (j % Z2)];
    } else if (j < E + h) {
      if (a.emb_src_in) {  // encode() / decode() of the module API: the caller passes the embedding rows themselves
        v = a.emb_src_in[(int64_t)b * h + (j - E)];
      } else {  // nn.Embedding raises IndexError for such a label; here: flag it (hippie_device_flags) and read row 0
        int64_t r = a.src[b];
        if (r < 0 || r >= a.num_sources) {
          if (a.flags) atomicOr(a.flags, kFlagSourceLabel);
          r = 0;
        }
        v = W[P.src_emb + r * h + (j - E)];
      }
    } else {
      if (a.emb_cls_in) {
        v = a.emb_cls_in[(int64_t)b * h + (j - E - h)];
      } else if (a.cls) {
        int64_t r = a.cls[b];
        if (r < 0 || r >= a.num_classes) {
          if (a.flags) atomicOr(a.flags, kFlagClassLabel);
          r = 0;
        }
        v = W[P.cls_emb + r * h + (j - E - h)];
      } else {
        v = 0.f;
      }
    }
    S[L.cat + (int64_t)b * D0 + j] = v;
  }
  __syncthreads();
  hstamp(1);  // S
  if (a.stage != kHeadDecodeOnly) {
  ph_linear(tc, S + L.cat, D0, W + P.f0_w, W + P.f0_b, S + L.f0, Z2, D0, Z2, b_lo, b_hi, -1.f);
  __syncthreads();
  hstamp(2);  // S
  if (a.train) {
    bn_partial<THREADS>(S + L.f0, Z2, Z2, b_lo, b_hi, part + (int64_t)blockIdx.x * Z2 * 4, s_mom);
    grid_sync();
    hstamp(3);  // G
    bn_finish<THREADS>(part, ncta, Z2, S + L.stats, a.run_mean + P.fbn_run, a.run_var + P.fbn_run, a.run_count + P.fbn_cnt,
                       s_mom, s_stat);
    bn_apply_own(tc, S + L.f0, Z2, Z2, b_lo, b_hi, W + P.fbn_g, W + P.fbn_b, s_stat, S + L.f1, Z2, kSlopeHead);
  } else {
    ph_bn_eval(S + L.f0, Z2, Z2, b_lo, b_hi, W + P.fbn_g, W + P.fbn_b, a.run_mean + P.fbn_run, a.run_var + P.fbn_run,
               S + L.f1, Z2, kSlopeHead);
  }
  __syncthreads();
  hstamp(4);  // S
  ph_linear(tc, S + L.f1, Z2, W + P.f3_w, W + P.f3_b, S + L.e0, z, Z2, z, b_lo, b_hi, -1.f);
  __syncthreads();
  hstamp(5);  // S
  if (P.ebn_g >= 0) {  // unimodal encoder_fc ends with BatchNorm1d(z) + LeakyReLU(0.2)  (hippie/model.py:21-28)
    if (a.train) {
      float* pe = part + part_layer;
      bn_partial<THREADS>(S + L.e0, z, z, b_lo, b_hi, pe + (int64_t)blockIdx.x * z * 4, s_mom);
      grid_sync();
      hstamp(6);  // G
      bn_finish<THREADS>(pe, ncta, z, S + L.stats + 2 * Z2, a.run_mean + P.ebn_run, a.run_var + P.ebn_run,
                         a.run_count + P.ebn_cnt, s_mom, s_stat);
      bn_apply_own(tc, S + L.e0, z, z, b_lo, b_hi, W + P.ebn_g, W + P.ebn_b, s_stat, S + L.enc, z, kSlopeHead);
    } else {
      ph_bn_eval(S + L.e0, z, z, b_lo, b_hi, W + P.ebn_g, W + P.ebn_b, a.run_mean + P.ebn_run, a.run_var + P.ebn_run,
                 S + L.enc, z, kSlopeHead);
    }
  } else {
    for (int idx = tc.t0; idx < nb * z; idx += tc.ts)
      S[L.enc + (int64_t)b_lo * z + idx] = S[L.e0 + (int64_t)b_lo * z + idx];
  }
  __syncthreads();
  hstamp(7);  // S

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
  hstamp(8);  // S
  if (threadIdx.x == 0 && a.kl_sum) {  // per-CTA partial; loss_finalize adds them in a fixed order (deterministic)
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
    a.kl_sum[blockIdx.x] = s;
  }
  if (!a.decode) return;
  }  // stage != kHeadDecodeOnly
  if (a.stage == kHeadDecodeOnly) {  // decode(z, source_emb, class_emb): z is an input
    for (int idx = tc.t0; idx < nb * z; idx += tc.ts) {
      const int b = b_lo + idx / z, i = idx % z;
      S[L.zc + (int64_t)b * DZ + i] = a.z_in[(int64_t)b * z + i];
    }
  }

  // zc = [z, source_emb, class_emb]   (hippie/model.py:412-413)
  for (int idx = tc.t0; idx < nb * 2 * h; idx += tc.ts) {
    const int b = b_lo + idx / (2 * h), j = idx % (2 * h);
    S[L.zc + (int64_t)b * DZ + z + j] = S[L.cat + (int64_t)b * D0 + E + j];
  }
  __syncthreads();
  hstamp(9);  // S
  // decoder_fc: Linear, LeakyReLU(.2), Linear, BatchNorm1d, LeakyReLU(.2); both branches share every phase
  for (int idx = tc.t0; idx < a.n_dec * nb * Z2; idx += tc.ts) {
    const int m = idx / (nb * Z2), r = idx - m * nb * Z2;
    const int b = b_lo + r / Z2, j = r % Z2;
    const float* x = S + L.zc + (int64_t)b * DZ;
    const float* w = W + P.d0_w[m] + (int64_t)j * DZ;
    float sum = 0.f;
    for (int i = 0; i < DZ; ++i) sum = fmaf(x[i], w[i], sum);
    S[L.g0[m] + (int64_t)b * Z2 + j] = lrelu(sum + W[P.d0_b[m] + j], kSlopeHead);
  }
  __syncthreads();
  hstamp(10);  // S
  for (int idx = tc.t0; idx < a.n_dec * nb * Z2; idx += tc.ts) {
    const int m = idx / (nb * Z2), r = idx - m * nb * Z2;
    const int b = b_lo + r / Z2, j = r % Z2;
    const float* x = S + L.g0[m] + (int64_t)b * Z2;
    const float* w = W + P.d2_w[m] + (int64_t)j * Z2;
    float sum = 0.f;
    for (int i = 0; i < Z2; ++i) sum = fmaf(x[i], w[i], sum);
    S[L.g1[m] + (int64_t)b * Z2 + j] = sum + W[P.d2_b[m] + j];
  }
  __syncthreads();
  hstamp(11);  // S
  if (a.train) {
    for (int m = 0; m < a.n_dec; ++m)
      bn_partial<THREADS>(S + L.g1[m], Z2, Z2, b_lo, b_hi, part + (2 + m) * part_layer + (int64_t)blockIdx.x * Z2 * 4, s_mom);
    grid_sync();
    hstamp(12);  // G
    for (int m = 0; m < a.n_dec; ++m) {
      bn_finish<THREADS>(part + (2 + m) * part_layer, ncta, Z2, S + L.stats + 2 * (Z2 + z) + m * 2 * Z2,
                         a.run_mean + P.dbn_run[m], a.run_var + P.dbn_run[m], a.run_count + P.dbn_cnt[m], s_mom, s_stat);
      bn_apply_own(tc, S + L.g1[m], Z2, Z2, b_lo, b_hi, W + P.dbn_g[m], W + P.dbn_b[m], s_stat, a.dout[m], Z2, kSlopeHead);
      __syncthreads();  // s_stat is reused by the second branch
      hstamp(13);  // S
    }
  } else {
    for (int m = 0; m < a.n_dec; ++m)
      ph_bn_eval(S + L.g1[m], Z2, Z2, b_lo, b_hi, W + P.dbn_g[m], W + P.dbn_b[m], a.run_mean + P.dbn_run[m],
                 a.run_var + P.dbn_run[m], a.dout[m], Z2, kSlopeHead);
  }
  hstamp(40);
}

// Cooperative launch.  Same sample ownership as the forward kernel; the weight gradients (batch reductions) come last.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) head_bwd_kernel(HeadArgs a) {
  hstamp(0);
  const HeadScratch L = head_layout(a.z, a.h, a.B);
  float* S = a.scratch;
  const HeadParams& P = a.hp;
  const float* W = a.params;
  float* G = a.grads;
  const int z = a.z, h = a.h, Z2 = 2 * z, E = a.n_enc * Z2, D0 = E + 2 * h, DZ = z + 2 * h, B = a.B;
  const Tc tc{(int)threadIdx.x, (int)blockDim.x};
  const int per = (B + gridDim.x - 1) / gridDim.x;
  const int b_lo = min(B, (int)blockIdx.x * per), b_hi = min(B, b_lo + per);
  const int nb = b_hi - b_lo;
  const int ncta = gridDim.x;
  __shared__ double s_mom[THREADS][3];
  __shared__ float s_stat[2 * kHeadMaxF];
  float* part = S + L.bnpart;
  const int64_t part_layer = (int64_t)kHeadTrainCtas * Z2 * 4;
  const int64_t dg1o[2] = {L.dg1, L.dg1b}, dg0o[2] = {L.dg0, L.dg0b};  // every gradient stays live for the wgrad phase

  for (int m = 0; m < a.n_dec; ++m)
    bnb_partial<THREADS>(a.dd[m], Z2, a.dout[m], Z2, S + L.g1[m], Z2, S + L.stats + 2 * (Z2 + z) + m * 2 * Z2, Z2, b_lo,
                         b_hi, kSlopeHead, part + (2 + m) * part_layer + (int64_t)blockIdx.x * Z2 * 4, s_mom);
  grid_sync();
  hstamp(1);  // G
  for (int m = 0; m < a.n_dec; ++m) {
    bnb_finish<THREADS>(part + (2 + m) * part_layer, ncta, Z2, B, G + P.dbn_g[m], G + P.dbn_b[m], s_mom, s_stat);
    bnb_apply_own(tc, a.dd[m], Z2, a.dout[m], Z2, S + L.g1[m], Z2, S + L.stats + 2 * (Z2 + z) + m * 2 * Z2, Z2, b_lo, b_hi,
                  W + P.dbn_g[m], s_stat, S + dg1o[m], Z2, kSlopeHead);
    __syncthreads();
    hstamp(2);  // S
  }
  for (int idx = tc.t0; idx < a.n_dec * nb * Z2; idx += tc.ts) {  // dg0 = (dg1 W_d2) * lrelu'(g0), both branches
    const int m = idx / (nb * Z2), r = idx - m * nb * Z2;
    const int b = b_lo + r / Z2, i = r % Z2;
    const float* g = S + dg1o[m] + (int64_t)b * Z2;
    const float* w = W + P.d2_w[m];
    float sum = 0.f;
    for (int j = 0; j < Z2; ++j) sum = fmaf(g[j], w[(int64_t)j * Z2 + i], sum);
    sum *= S[L.g0[m] + (int64_t)b * Z2 + i] > 0.f ? 1.f : kSlopeHead;
    S[dg0o[m] + (int64_t)b * Z2 + i] = sum;
  }
  __syncthreads();
  hstamp(3);  // S
  for (int idx = tc.t0; idx < nb * DZ; idx += tc.ts) {  // dzc = sum over the branches of dg0 W_d0
    const int b = b_lo + idx / DZ, i = idx % DZ;
    float sum = 0.f;
    for (int m = 0; m < a.n_dec; ++m) {
      const float* g = S + dg0o[m] + (int64_t)b * Z2;
      const float* w = W + P.d0_w[m];
      for (int j = 0; j < Z2; ++j) sum = fmaf(g[j], w[(int64_t)j * DZ + i], sum);
    }
    S[L.dzc + (int64_t)b * DZ + i] = sum;
  }
  __syncthreads();
  hstamp(4);  // S
  // reparameterisation + KL  (hippie/model.py:397-400, 472-474): total = ... + beta * mean_b(kl_b)
  const float kscale = a.beta / (float)B;
  for (int i0 = tc.t0; i0 < nb * z; i0 += tc.ts) {
    const int idx = b_lo * z + i0;
    const int b = idx / z, i = idx % z;
    const float dz = S[L.dzc + (int64_t)b * DZ + i];
    const float mu = S[L.mu + idx], lv = S[L.lv + idx];
    const float std = expf(0.5f * lv);
    S[L.dmu + idx] = dz + kscale * mu;
    S[L.dlv + idx] = dz * a.eps[idx] * 0.5f * std + kscale * 0.5f * (expf(lv) - 1.f);
  }
  __syncthreads();
  hstamp(6);  // S
  // denc = dmu * Wm + dlv * Wv in one pass
  for (int i0 = tc.t0; i0 < nb * z; i0 += tc.ts) {
    const int idx = b_lo * z + i0;
    const int b = idx / z, i = idx % z;
    float s = 0.f;
    for (int j = 0; j < z; ++j) {
      s = fmaf(S[L.dmu + (int64_t)b * z + j], W[P.zm_w + j * z + i], s);
      s = fmaf(S[L.dlv + (int64_t)b * z + j], W[P.zv_w + j * z + i], s);
    }
    S[L.denc + idx] = s;
  }
  __syncthreads();
  hstamp(7);  // S
  const float* de0 = S + L.denc;
  if (P.ebn_g >= 0) {
    float* pe = part + part_layer;
    bnb_partial<THREADS>(S + L.denc, z, S + L.enc, z, S + L.e0, z, S + L.stats + 2 * Z2, z, b_lo, b_hi, kSlopeHead,
                         pe + (int64_t)blockIdx.x * z * 4, s_mom);
    grid_sync();
    hstamp(8);  // G
    bnb_finish<THREADS>(pe, ncta, z, B, G + P.ebn_g, G + P.ebn_b, s_mom, s_stat);
    bnb_apply_own(tc, S + L.denc, z, S + L.enc, z, S + L.e0, z, S + L.stats + 2 * Z2, z, b_lo, b_hi, W + P.ebn_g, s_stat,
                  S + L.de0, z, kSlopeHead);
    de0 = S + L.de0;
    __syncthreads();
    hstamp(9);  // S
  }
  ph_dgrad(tc, de0, z, W + P.f3_w, Z2, z, S + L.df1, Z2, nullptr, 0, 0.f, false, b_lo, b_hi);
  __syncthreads();
  hstamp(10);  // S
  bnb_partial<THREADS>(S + L.df1, Z2, S + L.f1, Z2, S + L.f0, Z2, S + L.stats, Z2, b_lo, b_hi, kSlopeHead,
                       part + (int64_t)blockIdx.x * Z2 * 4, s_mom);
  grid_sync();
  hstamp(11);  // G
  bnb_finish<THREADS>(part, ncta, Z2, B, G + P.fbn_g, G + P.fbn_b, s_mom, s_stat);
  bnb_apply_own(tc, S + L.df1, Z2, S + L.f1, Z2, S + L.f0, Z2, S + L.stats, Z2, b_lo, b_hi, W + P.fbn_g, s_stat, S + L.df0,
                Z2, kSlopeHead);
  __syncthreads();
  hstamp(12);  // S
  ph_dgrad(tc, S + L.df0, Z2, W + P.f0_w, D0, Z2, S + L.dcat, D0, nullptr, 0, 0.f, false, b_lo, b_hi);
  __syncthreads();
  hstamp(13);  // S
  for (int e = 0; e < a.n_enc; ++e)
    for (int i0 = tc.t0; i0 < nb * Z2; i0 += tc.ts) {
      const int idx = b_lo * Z2 + i0;
      a.dh[e][idx] = S[L.dcat + (int64_t)(idx / Z2) * D0 + e * Z2 + (idx % Z2)];
    }

  // ---- weight gradients: every CTA adds the outer products of its own samples (atomics; the buffer is zeroed per step)
  {
    WgJob jobs[8];
    int nj = 0;
    auto add = [&](const float* dy, int ldy, const float* x, int ldx, int nin, int nout, int64_t w, int64_t b) {
      jobs[nj++] = WgJob{dy, x, G + w, G + b, ldy, ldx, nin, nout};
    };
    add(S + L.df0, Z2, S + L.cat, D0, D0, Z2, P.f0_w, P.f0_b);
    for (int m = 0; m < a.n_dec; ++m) {
      add(S + dg0o[m], Z2, S + L.zc, DZ, DZ, Z2, P.d0_w[m], P.d0_b[m]);
      add(S + dg1o[m], Z2, S + L.g0[m], Z2, Z2, Z2, P.d2_w[m], P.d2_b[m]);
    }
    add(de0, z, S + L.f1, Z2, Z2, z, P.f3_w, P.f3_b);
    add(S + L.dmu, z, S + L.enc, z, z, z, P.zm_w, P.zm_b);
    add(S + L.dlv, z, S + L.enc, z, z, z, P.zv_w, P.zv_b);
    ph_wgrad_own(tc, jobs, nj, b_lo, b_hi);
  }
  // embedding gradients: scatter-add of the two places each embedding row is used (cat and zc)
  for (int idx = tc.t0; idx < nb * 2 * h; idx += tc.ts) {
    const int b = b_lo + idx / (2 * h), k2 = idx % (2 * h);
    const bool is_cls = k2 >= h;
    if (is_cls && !a.cls) continue;
    const int k = is_cls ? k2 - h : k2;
    const int64_t row = is_cls ? a.cls[b] : a.src[b];
    if (row < 0 || row >= (is_cls ? a.num_classes : a.num_sources)) continue;  // flagged by the forward kernel
    const float v = S[L.dcat + (int64_t)b * D0 + E + k2] + S[L.dzc + (int64_t)b * DZ + z + k2];
    atomicAdd(G + (is_cls ? P.cls_emb : P.src_emb) + row * h + k, v);
  }
  hstamp(40);
}

}  // namespace

// Cooperative grid of the training-mode kernels: a few samples per CTA keeps the per-sample chain short, and the CTAs
// also share the grid-wide weight-gradient phase.  At most one CTA per SM (co-residency of the cooperative launch).
static int head_grid(int B) {
  int per = (B + kHeadTrainCtas - 1) / kHeadTrainCtas;
  if (per < 2) per = 2;
  int g = (B + per - 1) / per;
  return g < 1 ? 1 : g;
}

static void head_stamps_dump(const char* what, cudaStream_t s) {
  cudaStreamSynchronize(s);
  unsigned long long h[64];
  cudaMemcpyFromSymbol(h, g_head_stamps, sizeof(h));
  fprintf(stderr, "%s stamps (us since start):", what);
  for (int i = 1; i < 64; ++i)
    if (h[i] >= h[0] && h[i] - h[0] < 10000000ULL) fprintf(stderr, " %d:%.1f", i, (double)(h[i] - h[0]) * 1e-3);
  fprintf(stderr, "\n");
  unsigned long long z[64] = {0};
  cudaMemcpyToSymbol(g_head_stamps, z, sizeof(z));
}
static bool head_stamps_enabled() {
  static int on = -1;
  if (on < 0) {
    on = getenv("HIPPIE_B200_HEAD_STAMPS") ? 1 : 0;
    cudaMemcpyToSymbol(g_head_stamps_on, &on, sizeof(int));
  }
  return on == 1;
}

int launch_head_fwd(const HeadArgs& a, cudaStream_t s) {
  const bool st = head_stamps_enabled();
  struct Dump { bool on; cudaStream_t s; ~Dump() { if (on) head_stamps_dump("head_fwd", s); } } dump{st && a.train, s};
  if (a.train) {
    void* args[] = {const_cast<HeadArgs*>(&a)};
    const int grid = head_grid(a.B);
    if (grid == 1)
      head_fwd_kernel<256><<<1, 256, 0, s>>>(a);
    else
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
  struct Dump { bool on; cudaStream_t s; ~Dump() { if (on) head_stamps_dump("head_bwd", s); } } dump{head_stamps_enabled(), s};
  void* args[] = {const_cast<HeadArgs*>(&a)};
  const int grid = head_grid(a.B);
  if (grid == 1)
    head_bwd_kernel<256><<<1, 256, 0, s>>>(a);
  else
    cudaLaunchCooperativeKernel((const void*)head_bwd_kernel<256>, dim3(grid), dim3(256), args, 0, s);
}

}  // namespace hp
