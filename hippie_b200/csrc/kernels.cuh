// Launchers of the sm_100a kernels behind libhippie_b200.so (internal header).
//
// Data layout (DESIGN.md "Data layout in HBM"): every activation / gradient tensor is channels-last
// with one zero row on either side of each sample:  row(b, l) = b * (L + 2) + 1 + l,  C floats per row.
// A k=3, pad=1 convolution therefore reads, for output row (b, l), the 3*Cin CONTIGUOUS floats that
// start at input row b*(Lin+2) + l*stride -- the implicit-GEMM A row needs no im2col and no bounds
// checks.  Conv1d weights are stored [Cout][k][Cin] so the GEMM B row (K index = t*Cin + ci) is
// contiguous too.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace hp {

constexpr float kSlopeBackbone = 0.01f;  // F.leaky_relu default  (reference hippie/backbones.py:37,40,66,69,95)
constexpr float kSlopeHead = 0.2f;       // nn.LeakyReLU(0.2)     (reference hippie/model.py:367,381,384,388,391)
constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;

// ---- implicit-GEMM convolution:  C[m, n] (+)= sum_k A[row(m), k] * W[n, k]  (+ bias[n]) ----------------
struct ConvGemm {
  const float* A;     // input tensor base (row 0)
  const float* W;     // [N][K] K-contiguous
  const float* bias;  // [N] or null
  float* C;           // output tensor base (row 0)
  float* part;        // BatchNorm statistics partials [mtile][N][2] = (sum, centred M2), or null
  double* tot = nullptr;  // tcgen05 path: per-channel totals [copies][2][N] = (sum x, sum x^2), added with double
                          // atomics (zeroed once per step); replaces `part`
  int M, N, K;        // M = B * Lout logical rows
  int Lout;           // logical rows per sample
  int in_rows;        // padded rows per sample of A  (Lin + 2)
  int in_stride;      // convolution stride (rows)
  int in_off;         // first tap's padded-row offset (0 for k3/p1, 1 for k1/p0)
  int in_C;           // floats per input row
  int out_rows;       // padded rows per sample of C  (Lout_tensor + 2)
  int out_off;        // padded-row offset of logical row 0 (1)
  int out_lstride;    // padded rows between consecutive logical rows (1)
  int accumulate;     // C += result
};
// returns the number of logical rows per statistics tile (BM) it used
int launch_conv_gemm_simt(const ConvGemm& g, cudaStream_t s);

// ---- TMA tensor maps of the tcgen05 path (conv_pair.cu) are built on the host at first use ---------------
struct alignas(64) TcMap {
  unsigned char opaque[128];  // CUtensorMap
};

// ---- weight gradient:  dW[m, n] += sum_r dY[r, m] * X[(r + roff) * Cin + n]  (split-K, atomics) -------
struct WgradGemm {
  const float* dY;  // [R][M]  (R = B * (L + 2) padded rows, pad rows are zero)
  const float* X;   // input tensor base (row 0); has a finite guard row on either side
  float* dW;        // [M][N]  (N = k * Cin), zero-initialised by the caller
  int M, N, R;
  int Cin;
  int roff;  // -1 for k3/p1, 0 for k1/p0
};
void launch_wgrad_simt(const WgradGemm& g, int sm_count, cudaStream_t s);

// ---- 16-bit pair planes (conv_pair.cu): operands are (hi, lo) planes written by the producing kernels ----
// Eval-mode BatchNorm folded into the conv epilogue (running statistics are known before the conv runs):
//   y = lrelu((acc - mean) * scale + beta + res)  written as fp32 (optional), as fp16 pair planes (optional) and as the
//   nearest x2 up-sampled pair planes (optional) -- the same arithmetic, in the same order, as bn_apply_kernel.
struct EvalFold {
  const float* coef = nullptr;  // [0] = scale, [1] = beta, [2] = mean (C floats each); null = no folding
  const float* res = nullptr;   // residual tensor (same layout as the output) or null
  float slope = 1.f;            // LeakyReLU slope; 1 = no activation (shortcut branch)
  int write_f32 = 1;
  uint16_t* out_p = nullptr;
  int64_t out_ps = 0;
  uint16_t* up_p = nullptr;  // planes of the up-sampled copy: rows 2l, 2l+1 of a [B][2L+2][C] tensor
  int64_t up_ps = 0;
  unsigned* flags = nullptr;  // device error flags (kFlagPairSaturated)
};
struct PairOpts {
  float out_scale;   // applied to the accumulator sum in the epilogue (2^-8 when B holds scaled weights)
  int a_fmt, b_fmt;  // kPairF16 / kPairBF16 (pair_fmt.cuh)
  int b_mn;          // conv: 1 = B tiles read MN-major from the forward weight planes [co][t][ci] (dgrad)
  int taps;          // conv, b_mn = 1: kernel size of the layer
  const float* dyn_scale = nullptr;  // device float multiplied into out_scale (1 / scale of a gradient pair tensor)
  unsigned long long* stamps = nullptr;  // tools/pair_test: 8 device slots for %globaltimer phase stamps of CTA (0, 0)
  int stamps_all = 0;  // tools/pair_test: every CTA writes 8 slots (slot 7 = %smid)
  const EvalFold* fold = nullptr;        // conv forward in eval mode
  const TcMap* out_map = nullptr;        // conv: fp32 output tensor map (pair_make_out_map) -> TMA stores of whole tiles
};
bool pair_init(std::string* err);
bool pair_make_act_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int in_C, int K, int Lout,
                       int in_rows, int in_stride, int in_off, int max_batch);
bool pair_make_w_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int N, int K, int bn);
bool pair_make_wmn_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int Cout, int Cin, int k);
bool pair_make_rows_map(TcMap* out, const void* planes, int64_t plane_stride, int fmt, int64_t row_elems, int channels,
                        int rows);
bool pair_make_out_map(TcMap* out, float* C, int N, int Lout, int out_rows, int out_off, int batch);
bool pair_make_dw_map(TcMap* out, float* dW, int M, int N);
int pair_pick_bn(int B, int N, int Lout, int sm_count);
int launch_conv_pair(const ConvGemm& g, const TcMap& mapA, const TcMap& mapB, int bn, int B, const PairOpts& o,
                     cudaStream_t s);
void launch_wgrad_pair(const WgradGemm& g, const TcMap& mapDY, const TcMap& mapX, const TcMap& mapDW, int bn, int sm_count,
                       const PairOpts& o, cudaStream_t s);
// planes[0 .. n) = hi(src * scale), planes[plane_stride .. plane_stride + n) = lo
void launch_to_pair(const float* src, void* planes, int64_t plane_stride, int64_t n, float scale, int fmt, cudaStream_t s,
                    unsigned* flags = nullptr);


// ---- stem conv (Cin = 1, k3, s2, p1)  reference hippie/backbones.py:78,95 -----------------------------
void launch_stem_fwd(const float* x, const float* w /*[64][3]*/, float* c0, float* part, int B, int Lin, int Lout,
                     cudaStream_t s);
// dW partials: part[cta][64*3]; reduce with launch_reduce_partials
int launch_stem_wgrad(const float* x, const float* dc0, float* part, int B, int Lin, int Lout, cudaStream_t s);

// out[i] (+)= sum_p part[p][i]   (double accumulation)
void launch_reduce_partials(const float* part, int nparts, int n, float* out, int accumulate, cudaStream_t s);

// ---- BatchNorm (train-mode statistics over B*L rows)  ---------------------------------------------------
// coef layout per BatchNorm: 8 arrays of C floats: [0]=scale [1]=shift [2]=mean [3]=invstd [4]=k [5]=m1 [6]=m2 [7]=spare
struct BnFinalize {
  const float* part;  // [ntiles][C][2]
  const double* tot = nullptr;  // [copies][2][C] totals (sum x, sum x^2) when the conv epilogues accumulated them
  int ntiles, tile_rows, M, C;
  // host-computed reciprocals (double division / sqrt are ~0.2 us software sequences on the device, and the finalize
  // phase sits on the critical path of every BatchNorm): 1/M, 1/max(M-1,1), 1/tile_rows, 1/(rows of the last tile)
  double inv_M, inv_Mm1, inv_tile, inv_last;
  void set_shape(int ntiles_, int tile_rows_, int M_) {
    ntiles = ntiles_, tile_rows = tile_rows_, M = M_;
    inv_M = 1.0 / M_, inv_Mm1 = 1.0 / (M_ > 1 ? M_ - 1 : 1), inv_tile = 1.0 / tile_rows_;
    inv_last = 1.0 / (M_ - (ntiles_ - 1) * tile_rows_);
  }
  const float* gamma;
  const float* beta;
  float* run_mean;
  float* run_var;
  int64_t* run_count;
  float* coef;
};
// eval mode, all BatchNorms of the model in one launch: scale/shift from the running statistics
struct BnEvalEntry {
  int64_t gamma_off, beta_off, run_off, coef_off;
  int C;
};
void launch_bn_eval_coefs(const BnEvalEntry* table_dev, int n, const float* params, const float* run_mean,
                          const float* run_var, float* ws, cudaStream_t s);

// out = lrelu(c*scale+shift (+ r | + r*rscale+rshift)); optional second copy nearest-upsampled x2
struct BnApply {
  const float* c;
  const float* coef;
  const float* r;      // residual tensor or null
  const float* rcoef;  // coef of the residual's BatchNorm, or null for an identity shortcut
  float* out;
  float* out_up;  // [B][2L+2][C] or null
  int B, L, C;
  float slope;
  // train = 1: (scale, shift) come from the conv's statistics partials (fin / rfin: coef, running buffers, part,
  // ntiles, tile_rows, M, C filled in); train = 0: from coef / rcoef (eval mode: bn_eval_coefs)
  int train = 0;
  BnFinalize fin{};
  BnFinalize rfin{};
  // fp16 pair planes of `out` / of the up-sampled copy (row 0 of the hi plane; lo plane `*_ps` elements later), or null
  uint16_t* out_p = nullptr;
  int64_t out_ps = 0;
  uint16_t* up_p = nullptr;
  int64_t up_ps = 0;
  unsigned long long* stamps = nullptr;  // tools/bn_test: %globaltimer phase stamps of CTA (0, 0)
  unsigned* flags = nullptr;             // device error flags: kFlagPairSaturated when |out| exceeds the fp16 range
};
void launch_bn_apply(const BnApply& a, int sm_count, cudaStream_t s);

// backward of  out = lrelu(bn(c) + [bn_s(cs) | identity]) :
//   g_pre = g * (out > 0 ? 1 : slope);  per channel S1 = sum g_pre, S2 = sum g_pre * xhat, S2s likewise for cs
struct BnBwd {
  const float* g;    // gradient w.r.t. out; if g_up != 0 it is a [B][2L+2][C] tensor and g(b,l) = g[2l] + g[2l+1]
  int g_up;
  const float* out;
  const float* c;
  float* coef;
  const float* cs;  // shortcut conv output or null
  float* coef_s;
  float* part;  // [nchunks][C][3] = (S1, S2, S2s)
  double* tot = nullptr;  // [C][3] totals added with double atomics by the reduce kernel (zeroed once per step)
  int B, L, C;
  float slope;
  // finalize: parameter gradients
  const float* gamma;
  const float* gamma_s;
  float* dgamma;
  float* dbeta;
  float* dgamma_s;
  float* dbeta_s;
  // apply
  float* dc;  // gradient w.r.t. c, written at row b*(Ld+2)+1+dil*l
  int dil, Ld;
  float* dcs;
  int dil_s, Ld_s;
  float* gres;  // if non-null: g_pre is written here (identity shortcut), same layout as out
  double inv_n = 0.0;  // 1 / (B * L), computed on the host
  // pair path: dc / dcs may be null; the gradients go to fp16 pair planes scaled by a power of two chosen from an upper
  // bound of max|dc|.  slot = (max|g_pre|, max|xhat|, max|gamma*invstd|, 1 / scale): [0..2] are atomicMax targets that
  // the engine zeroes once per step, [3] is read by the dgrad / wgrad epilogues
  uint16_t* dc_p = nullptr;
  int64_t dc_ps = 0;
  float* dc_slot = nullptr;
  uint16_t* dcs_p = nullptr;
  int64_t dcs_ps = 0;
  float* dcs_slot = nullptr;
};
void launch_bn_bwd(const BnBwd& a, int sm_count, cudaStream_t s);  // reduce + apply (2 launches)
constexpr int kBnBwdLaunches = 2;
constexpr int kBnBwdTotCopies = 8;  // the reduce CTAs spread their atomics over this many copies of the [C][3] totals
constexpr int kBnFwdTotCopies = 1;  // copies of the conv epilogues' [2][C] totals (copy = M-tile index mod copies); 8 measured no gain
constexpr int kBnBwdMaxChunks = 128;  // every apply CTA re-sums the partials of its 32 channels

// dst[b,l,c] += src[b,2l,c] + src[b,2l+1,c]     (backward of nearest x2 up-sampling)
void launch_pairsum_acc(const float* src, float* dst, int B, int L, int C, cudaStream_t s);

// ---- encoder tail: adaptive_avg_pool1d + Linear(512 -> F)   reference hippie/backbones.py:100-102 -------
void launch_pool_linear_fwd(const float* x4, int B, int L, int C, const float* W, const float* bias, int F,
                            float* pooled, float* h, cudaStream_t s);
// dpool -> g_x4 (overwrite)
void launch_pool_linear_bwd_x(const float* dh, const float* W, int B, int L, int C, int F, float* g_x4, cudaStream_t s);
// dW[j][i] = sum_b dy[b][j] * x[b][i], db[j] = sum_b dy[b][j]: the weight gradients of the Linear layers next to the
// backbones; off the critical path (the engine launches them on the weight-gradient streams)
void launch_linear_wgrad(const float* dy, int ldy, const float* x, int ldx, int B, int nin, int nout, float* dW, float* db,
                         cudaStream_t s);

// ---- decoder head: Linear(F -> 512) + unsqueeze + nearest x4   reference hippie/backbones.py:129-131 ---
void launch_dec_linear_fwd(const float* d, int B, int F, const float* W, const float* bias, int C, float* t0,
                           uint16_t* t0_p /*fp16 pair planes or null*/, int64_t t0_ps, unsigned* flags, cudaStream_t s);
void launch_dec_linear_bwd_x(const float* g_t0, const float* W, int B, int F, int C, float* gx0, float* dd,
                             cudaStream_t s);

// ---- decoder tail: nearest x2 -> conv(64->1,k3,bias) -> view -> Linear(64 -> Lo) + MSE  -----------------
//      reference hippie/backbones.py:117-118,136-139 and hippie/model.py:465-466
struct DecTail {
  const float* x;  // layer1 output [B][32+2][64]
  const float* wc;
  const float* bc;  // conv weight [1][3][64], bias [1]
  const float* Wo;
  const float* bo;      // linear_out [Lo][64], [Lo]
  const float* target;  // [B][Lo] or null (no loss)
  float* dec;           // [B][Lo] (scratch or user output)
  int B, Lo;
  // training extras (null in eval)
  float* y;       // [B][64] conv output (saved)
  float* ddec;    // [B][Lo]
  float* dy;      // [B][64]
  float* g_x;     // gradient w.r.t. x, [B][34][64] layout
  float* part;    // [ncta][196]: 192 dwc, 1 dbc, 1 sum of squared error
  float loss_w;   // modality weight w1 / w2
  int train;
};
int launch_dec_tail(const DecTail& t, cudaStream_t s);  // returns ncta
// reduces the partials: dwc, dbc -> grads; dWo, dbo from ddec,y; sse -> scal_sse (a float)
void launch_dec_tail_reduce(const DecTail& t, int ncta, float* dwc, float* dbc, float* dWo, float* dbo, float* sse,
                            cudaStream_t s);

// ---- latent head (embeddings, fusion MLP, mu/logvar, reparameterise, KL, decoder_fc) -------------------
//      reference hippie/model.py:402-432 (multimodal), :46-72 (unimodal)
struct HeadParams {  // offsets (floats) into the flat parameter / gradient buffers; -1 = absent
  int64_t f0_w, f0_b, fbn_g, fbn_b, f3_w, f3_b, ebn_g, ebn_b;  // fusion_encoder / encoder_fc (.0,.1,.3,.4)
  int64_t src_emb, cls_emb, zm_w, zm_b, zv_w, zv_b;
  int64_t d0_w[2], d0_b[2], d2_w[2], d2_b[2], dbn_g[2], dbn_b[2];
  int64_t fbn_run, ebn_run, dbn_run[2];  // offsets into bn_mean/bn_var; count index = *_cnt
  int fbn_cnt, ebn_cnt, dbn_cnt[2];
};
struct HeadArgs {
  HeadParams hp;
  const float* params;
  float* grads;
  float* run_mean;
  float* run_var;
  int64_t* run_count;
  int z, h, n_enc, n_dec;  // h = class_hidden_dim; n_enc/n_dec = 2 multimodal, 1 unimodal
  int num_sources, num_classes;
  int B;
  const float* hin[2];  // encoder outputs [B][2z]
  const int64_t* src;
  const int64_t* cls;  // null -> zeros
  const float* eps;    // [B][z]
  float* scratch;      // head workspace (see head_scratch_floats)
  float* dout[2];      // decoder_fc outputs [B][2z]
  float* out_enc;
  float* out_mu;
  float* out_logvar;  // optional user outputs [B][z]
  float* kl_sum;      // [number of CTAs] partial sums of kl_b (added in order by launch_loss_finalize)
  int train;          // batch statistics + running update
  int decode;         // 0 = stop after mu/logvar (embedding pass)
  int zscore_ddof;    // -1 none; else z-score out_enc rows in place
  // module API (MultiModalCVAE.encode / .decode, hippie/model.py:402-422): explicit embedding rows [B][h] instead of the
  // label gathers, z [B][z] as an input, and which part of the head runs
  const float* emb_src_in = nullptr;
  const float* emb_cls_in = nullptr;
  const float* z_in = nullptr;
  int stage = 0;               // kHeadAll (decode = 0 stops after mu / logvar) or kHeadDecodeOnly
  unsigned* flags = nullptr;   // device error flags (kFlag*), sticky until hippie_device_flags reads them
  // backward only
  const float* dd[2];  // gradients w.r.t. decoder_fc outputs
  float* dh[2];        // gradients w.r.t. encoder outputs
  float beta;
};
constexpr int kHeadAll = 0, kHeadDecodeOnly = 2;
// device error flags (include/hippie_b200.h: HIPPIE_FLAG_*)
constexpr unsigned kFlagSourceLabel = 1u, kFlagClassLabel = 2u, kFlagPairSaturated = 4u, kFlagWeightSaturated = 8u;
int64_t head_scratch_floats(int z, int h, int B);
constexpr int kHeadMaxCtas = 592;
int launch_head_fwd(const HeadArgs& a, cudaStream_t s);  // returns the number of CTAs (= KL partials written)
void launch_head_bwd(const HeadArgs& a, cudaStream_t s);

// ---- loss scalars: total = w1*mse1 + w2*mse2 + beta*kl ---------------------------------------------------
void launch_loss_finalize(const float* sse1, const float* sse2, const float* kl_parts, int n_kl, int B, int Lo1, int Lo2,
                          float beta, float w1, float w2, int multimodal, float* scalars, cudaStream_t s);

// ---- gradient clipping + AdamW over the flat buffers ----------------------------------------------------
struct AdamArgs {
  float* p;
  float* g;
  float* m;
  float* v;
  int64_t n;
  int64_t skip_lo, skip_hi;  // class_embedding range, stepped with step_cls (or skipped when has_cls_grad = 0)
  double lr, beta1, beta2, eps, wd;  // torch keeps the hyper-parameters as Python doubles: 1 - beta is taken in double
  float max_norm, grad_scale;
  int step, step_cls, has_cls_grad;
  float* partials;  // >= 1024 floats
  float* scalars;   // [4] = grad norm, [5] = clip coefficient
  // optional: fp16 pair planes of the parameter buffer (tensor-core GEMM operands), rewritten for every updated element
  // while it is in a register (the separate 64 MB-read / 64 MB-write refresh pass of the next step goes away)
  uint16_t* wp_hi;
  uint16_t* wp_lo;
  float wp_scale;
  unsigned* flags;  // receives kFlagWeightSaturated when an updated weight leaves the range of the hi plane
};
void launch_clip_adamw(const AdamArgs& a, cudaStream_t s);
constexpr int kClipAdamLaunches = 3;

// rows index[b] (null: b) of a raw float64 table -> [B][size] float32: cast, optional log(x + 1), linear interpolation
void launch_preprocess(const double* raw, int width, const int64_t* index, int B, int size, int take_log, float* out,
                       cudaStream_t s);

// up to 8 segment copies of 4-byte words in one launch (input staging / output delivery around a graph replay)
struct IoSeg {
  const void* src;
  void* dst;
  int64_t words;
};
struct IoCopy {
  IoSeg seg[8];
  int n;
};
void launch_io_copy(const IoCopy& c, cudaStream_t s);

// transposed + tap-flipped copies of the conv weights for dgrad:  wt[ci][k-1-t][co] = w[co][t][ci]
struct WtEntry {
  int64_t w_off, wt_off;
  int cout, cin, k;
};
void launch_refresh_wt(const WtEntry* table_dev, int n, const float* params, float* ws, cudaStream_t s);

}  // namespace hp
