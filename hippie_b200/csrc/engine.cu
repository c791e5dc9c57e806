// libhippie_b200.so -- C ABI (include/hippie_b200.h) and the layer program of the HIPPIE cVAE.
//
// The engine owns no tensor memory: parameters, gradients, AdamW state and BatchNorm buffers are
// flat device buffers handed over by hippie_bind(); activations live in a caller-provided workspace.
// The network structure below restates, as launch sequences over channels-last padded tensors,
//   ResNet18Enc / BasicBlockEnc      reference hippie/backbones.py:19-41, 73-103
//   ResNet18Dec / BasicBlockDec      reference hippie/backbones.py:44-70, 106-141
//   MultiModalCVAE / unimodal twin   reference hippie/model.py:350-432, 12-72
//   training_step loss               reference hippie/model.py:454-482
// and their backward passes (autograd in the reference).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hippie_b200.h"
#include "kernels.cuh"
#include "pair_fmt.cuh"

using namespace hp;

namespace {

struct Param {
  std::string name;
  int ndim = 0;
  int64_t shape[3] = {0, 0, 0};
  int64_t off = 0, numel = 0;
  int layout = HIPPIE_LAYOUT_NATIVE;
};
struct BNInfo {
  std::string name;
  int C = 0;
  int gamma = -1, beta = -1;  // param indices
  int64_t run_off = 0;        // offset into bn_mean / bn_var
  int64_t coef_off = 0;       // workspace offset of the 8*C coefficient block
  int64_t part_off = -1;      // workspace offset of this BatchNorm's statistics partials [tiles][C][2]
  int64_t tot_off = -1;       // workspace offset of its totals: forward [2][C] doubles, then backward [C][3] doubles
  int ntiles = 0, tile_rows = 0, M = 0;  // shape of the partials the producing conv left (set at launch time)
  bool totals = false;                   // the producing conv accumulated totals instead of per-tile partials
};
struct Act {
  std::string name;
  int L = 0, C = 0;
  int64_t off = -1;   // workspace float offset of padded row 0 of the fp32 tensor (-1: no fp32 copy)
  int64_t poff = -1;  // workspace float offset of padded row 0 of the fp16 hi plane (-1: no pair planes)
  int64_t pstride = 0;  // elements between the hi and the lo plane
  int64_t slot = -1;  // gradient pair tensors: workspace offset of (bound, scale, 1 / scale)
};
enum { kF32 = 1, kPlanes = 2, kSlot = 4 };
struct Conv {
  int w = -1, b = -1;
  int cin = 0, cout = 0, k = 3, stride = 1;
  int64_t wt_off = -1;
  int id = -1;
};
struct ConvMaps {  // TMA tensor maps of the tcgen05 path, built at first use after hippie_bind
  TcMap a_fwd, w_k, a_dg, w_mn, wg_dy, wg_x, c_fwd, c_dg, dw;
  bool fwd_ready = false, dg_ready = false, dw_ready = false;
  const float *c_fwd_ptr = nullptr, *c_dg_ptr = nullptr;  // the output tensors the store maps describe ...
  int c_fwd_B = -1, c_dg_B = -1;                          // ... and the batch size they end at
  int wg_B = -1;  // the wgrad maps bound the reduction rows, so they depend on the batch size
};
struct EncBlock {
  Conv c1, c2, cs;
  int bn1 = -1, bn2 = -1, bns = -1;
  bool down = false;
  int x = -1, c1o = -1, a1 = -1, c2o = -1, cso = -1, out = -1;
  int g_a1 = -1, dc1 = -1, dc2 = -1, dcs = -1;
};
struct Encoder {
  std::string prefix;
  int stem_w = -1, bn0 = -1, Lin = 0, L0 = 0;
  int c0 = -1, a0 = -1, dc0 = -1;
  EncBlock blk[8];
  int lin_w = -1, lin_b = -1;
  int64_t pooled = 0, h = 0, dh = 0;  // workspace offsets
};
struct DecBlock {
  Conv c2, c1, cs;
  int bn2 = -1, bn1 = -1, bns = -1;
  bool up = false;
  int x = -1, x_up = -1, c2o = -1, a2 = -1, a2_up = -1, c1o = -1, cso = -1, out = -1, out_up = -1;
  int dc2 = -1, dc1 = -1, dcs = -1, g_a2 = -1, g_a2_up = -1, g_x_up = -1;
};
struct Decoder {
  std::string prefix;
  int lin_w = -1, lin_b = -1, t0 = -1;
  DecBlock blk[8];
  int wc = -1, bc = -1, lo_w = -1, lo_b = -1, Lo = 0;
  int64_t d = 0, dd = 0, gx0 = 0, y = 0, ddec = 0, dy = 0, dec = 0, part = 0;  // workspace offsets
};
struct Branch {
  cudaStream_t st = nullptr;
  float* part = nullptr;   // BatchNorm statistics / stem / tail partials
  float* bpart = nullptr;  // BatchNorm backward partials
  cudaStream_t wst = nullptr;  // weight gradients are off the critical path: they run here, behind an event
  cudaStream_t hst = nullptr;  // helper stream of the branch: the shortcut convolution of a down / up block runs beside the main path
};

}  // namespace

struct hippie_engine {
  hippie_cfg cfg{};
  std::string err;
  int sm_count = 148;
  std::vector<Param> params;
  std::map<std::string, int> pidx;
  std::vector<BNInfo> bns;
  std::map<std::string, int> bidx;
  std::vector<Act> acts;
  std::map<int, int> gact;  // activation index -> gradient tensor index
  int64_t param_floats = 0, bn_floats = 0, ws_floats = 0;
  int64_t grad_split = 0;  // first parameter offset that does not belong to an encoder backbone
  // gradient ranges in the order the backward pass completes them (hippie_grad_bounds): encoder e owns
  // [enc_begin[e], enc_deep[e]) = stem + layer1 + layer2 (final after part 3) and [enc_deep[e], enc_end[e]) = layer3 +
  // layer4 + linear (final after part 2)
  int64_t enc_begin[2] = {0, 0}, enc_deep[2] = {0, 0}, enc_end[2] = {0, 0};
  int n_enc = 0, n_dec = 0;
  Encoder enc[2];
  Decoder dec[2];
  HeadParams headp{};
  int64_t head_scratch = 0, scal_off = 0, part_off[2] = {0, 0}, bpart_off[2] = {0, 0}, adam_part = 0;
  int64_t part_floats = 0, bpart_floats = 0, tot_off = 0, tot_floats = 0;
  int64_t wt_table_off = 0, bn_table_off = 0;
  std::vector<WtEntry> wt_table;
  std::vector<BnEvalEntry> bn_table;
  int cls_emb = -1;
  // bound buffers
  float *P = nullptr, *G = nullptr, *M1 = nullptr, *M2 = nullptr, *bn_mean = nullptr, *bn_var = nullptr;
  int64_t* bn_count = nullptr;
  float* ws = nullptr;
  bool bound = false;
  cudaStream_t side = nullptr;
  cudaStream_t wside[2] = {nullptr, nullptr};
  cudaStream_t hside[2] = {nullptr, nullptr};  // Branch::hst
  // Batch sizes up to this run the shortcut convolutions on the helper streams.  Every fork / join costs the kernels behind
  // it their programmatic launch edge, so it pays where kernels are short and the SMs are far from full (bs64 step
  // 1.476 -> 1.453 ms, bs64 embedding pass 0.284 -> 0.267 ms) and loses at bs512 (3.015 -> 3.06 ms); HIPPIE_B200_SHORTCUT_STREAM
  int helper_max_batch = 128;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  cudaEvent_t next_event() { return ev_pool[ev_next++ % ev_pool.size()]; }
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int launches = 0;
  int n_convs = 0;
  bool failed = false;  // a tensor-map encode failed while launching (reported by the entry point)
  // ---- CUDA graphs: one instantiated graph per call signature, replayed from staged inputs ----------------
  struct CallArgs {
    int mode;  // 0 train_fwd_bwd, 1 train_forward, 2 eval_forward, 3 embed, 4 + p = part p (0..3) of a split train_fwd_bwd
    const float *x1, *x2;
    const int64_t *src, *cls;
    const float* eps;
    int B;
    float beta, w1, w2;
    int zscore;
    float *scalars, *enc, *mu, *lv, *d1, *d2;
  };
  struct GraphKey {
    int mode, B, flags, zscore;
    float beta, w1, w2;
    bool operator<(const GraphKey& o) const {
      return std::memcmp(this, &o, sizeof(GraphKey)) < 0;
    }
  };
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    int seen = 0;  // 0 = never called (first call runs eagerly: lazy initialisation), -1 = capture failed, stay eager
  };
  std::map<GraphKey, GraphEntry> graphs;
  bool use_graphs = true;
  unsigned long long graph_flags = cudaGraphInstantiateFlagUseNodePriority;  // HIPPIE_B200_GRAPH_PRIO=0: every node at the caller's priority
  static constexpr size_t kMaxGraphs = 32;  // distinct call signatures kept as instantiated graphs
  cudaStream_t cap = nullptr;  // capture stream (the caller's stream may be the legacy default stream)
  // Data-parallel exchange without splitting the step (hippie_train_fwd_bwd_part, part 4): the whole step stays ONE
  // graph; `xs` collects, without ever holding up the backward chain, the points at which a slice of the gradient
  // buffer is final (chain stream + weight-gradient streams) and records ev_slice[k] -- as an EXTERNAL event-record node
  // when the step is captured -- for the caller's exchange stream to wait on (hippie_slice_wait).
  cudaStream_t xs = nullptr;
  cudaEvent_t ev_slice[2] = {nullptr, nullptr};
  bool export_ev = false, export_capturing = false;
  std::vector<cudaEvent_t> deep_ev;
  int64_t st_x1 = 0, st_x2 = 0, st_src = 0, st_cls = 0, st_eps = 0, st_scal = 0, st_enc = 0, st_mu = 0, st_lv = 0,
          st_d1 = 0, st_d2 = 0;
  void clear_graphs() {
    for (auto& kv : graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
  }
  std::vector<ConvMaps> cmaps;
  bool tma_epilogue = true;  // conv outputs leave through TMA tensor stores (HIPPIE_B200_TMA_STORE=0: per-thread stores)
  bool planes_fresh = false;  // the weight pair planes equal the parameter buffer (refresh_weights)
  bool planes_keep = true;    // HIPPIE_B200_KEEP_PLANES=0: convert the whole buffer in every forward-type call (round 1)
  bool use_tc = false;  // tcgen05 implicit GEMMs over fp16 pair planes (conv_path 0); false = FP32 CUDA-core GEMMs
  std::string tc_note;
  int64_t flags_off = 0;  // device error flags (HIPPIE_FLAG_*), sticky until hippie_device_flags clears them
  unsigned* flags() { return reinterpret_cast<unsigned*>(ws + flags_off); }
  int64_t wp_off = 0;     // weight pair planes: hi plane at ws + wp_off (as halfs), lo plane param_floats elements later
  int64_t slots_off = 0;  // (max|g|, max|xhat|, max|k|, 1/scale) per gradient pair tensor, zeroed once per step
  int n_slots = 0;
  static constexpr int kMaxSlots = 256;
  // profiling aid (bench.py roofline): CUDA events around every implicit-GEMM launch
  struct ProfRec {
    int kind;
    cudaEvent_t e0, e1;
    double flop;
  };
  bool profiling = false;
  std::vector<ProfRec> prof;
  cudaEvent_t prof_begin(Branch& br) {
    if (!profiling) return nullptr;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, br.st);
    return e;
  }
  void prof_end(cudaEvent_t e0, int kind, double flop, Branch& br) {
    if (!profiling) return;
    cudaEvent_t e1;
    cudaEventCreate(&e1);
    cudaEventRecord(e1, br.st);
    prof.push_back({kind, e0, e1, flop});
  }

  // ---- construction -------------------------------------------------------------------------
  int64_t take(int64_t floats) {
    int64_t r = ws_floats;
    ws_floats += (floats + 63) & ~(int64_t)63;
    return r;
  }
  // stand-alone backbones (cfg.multimodal = 2 / 3) have no module prefix: ".conv1.weight" -> "conv1.weight"
  static std::string nm(const std::string& s) { return (!s.empty() && s[0] == '.') ? s.substr(1) : s; }
  int pix(const std::string& name) const { return pidx.at(nm(name)); }
  int bix(const std::string& name) const { return bidx.at(nm(name)); }
  int add_param(const std::string& name_, int ndim, int64_t s0, int64_t s1, int64_t s2, int layout) {
    const std::string name = nm(name_);
    Param p;
    p.name = name, p.ndim = ndim, p.shape[0] = s0, p.shape[1] = s1, p.shape[2] = s2, p.layout = layout;
    p.numel = s0 * (ndim > 1 ? s1 : 1) * (ndim > 2 ? s2 : 1);
    p.off = param_floats;
    param_floats += (p.numel + 7) & ~(int64_t)7;  // 16-byte aligned in the fp16 pair planes too
    pidx[name] = (int)params.size();
    params.push_back(p);
    return (int)params.size() - 1;
  }
  void spec_conv(const std::string& n, int cout, int cin, int k, bool bias) {
    add_param(n + ".weight", 3, cout, cin, k, HIPPIE_LAYOUT_CONV_OKI);
    if (bias) add_param(n + ".bias", 1, cout, 0, 0, HIPPIE_LAYOUT_NATIVE);
  }
  void spec_bn(const std::string& n_, int c) {
    const std::string n = nm(n_);
    BNInfo b;
    b.name = n, b.C = c;
    b.gamma = add_param(n + ".weight", 1, c, 0, 0, HIPPIE_LAYOUT_NATIVE);
    b.beta = add_param(n + ".bias", 1, c, 0, 0, HIPPIE_LAYOUT_NATIVE);
    b.run_off = bn_floats;
    bn_floats += (c + 3) & ~3;
    bidx[n] = (int)bns.size();
    bns.push_back(b);
  }
  void spec_linear(const std::string& n, int nout, int nin) {
    add_param(n + ".weight", 2, nout, nin, 0, HIPPIE_LAYOUT_NATIVE);
    add_param(n + ".bias", 1, nout, 0, 0, HIPPIE_LAYOUT_NATIVE);
  }
  // ResNet18Enc.__init__ (hippie/backbones.py:74-84); BasicBlockEnc.__init__ (:20-34)
  void spec_encoder(const std::string& p, int z) {
    spec_conv(p + ".conv1", 64, 1, 3, false);
    spec_bn(p + ".bn1", 64);
    int in_planes = 64;
    const int planes[4] = {64, 128, 256, 512}, strides[4] = {1, 2, 2, 2};
    for (int li = 0; li < 4; ++li)
      for (int bi = 0; bi < 2; ++bi) {
        const int s = bi == 0 ? strides[li] : 1, out = in_planes * s;
        const std::string q = p + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
        spec_conv(q + ".conv1", out, in_planes, 3, false);
        spec_bn(q + ".bn1", out);
        spec_conv(q + ".conv2", out, out, 3, false);
        spec_bn(q + ".bn2", out);
        if (s != 1) {
          spec_conv(q + ".shortcut.0", out, in_planes, 1, false);
          spec_bn(q + ".shortcut.1", out);
        }
        in_planes = planes[li];
      }
    spec_linear(p + ".linear", 2 * z, 512);
  }
  // ResNet18Dec.__init__ (hippie/backbones.py:107-126): _make_layer reverses the strides
  void spec_decoder(const std::string& p, int z, int output_size) {
    spec_linear(p + ".linear", 512, 2 * z);
    int in_planes = 512;
    const int lis[4] = {4, 3, 2, 1}, planes[4] = {256, 128, 64, 64}, strides[4] = {2, 2, 2, 1};
    for (int i = 0; i < 4; ++i) {
      for (int bi = 0; bi < 2; ++bi) {
        const int s = bi == 0 ? 1 : strides[i], out = in_planes / s;
        const std::string q = p + ".layer" + std::to_string(lis[i]) + "." + std::to_string(bi);
        spec_conv(q + ".conv2", in_planes, in_planes, 3, false);
        spec_bn(q + ".bn2", in_planes);
        if (s == 1) {
          spec_conv(q + ".conv1", out, in_planes, 3, false);
          spec_bn(q + ".bn1", out);
        } else {
          spec_conv(q + ".conv1.conv", out, in_planes, 3, true);
          spec_bn(q + ".bn1", out);
          spec_conv(q + ".shortcut.0.conv", out, in_planes, 3, true);
          spec_bn(q + ".shortcut.1", out);
        }
      }
      in_planes = planes[i];
    }
    spec_conv(p + ".conv1.conv", 1, 64, 3, true);
    spec_linear(p + ".linear_out", output_size, 64);
  }
  bool full_model() const { return cfg.multimodal == 0 || cfg.multimodal == 1; }
  void build_spec() {
    const int z = cfg.z_dim, h = cfg.class_hidden_dim;
    if (cfg.multimodal == HIPPIE_KIND_ENCODER) {  // ResNet18Enc as a module of its own (hippie/backbones.py:73-103)
      spec_encoder("", z);
    } else if (cfg.multimodal == HIPPIE_KIND_DECODER) {  // ResNet18Dec (hippie/backbones.py:106-141)
      spec_decoder("", z, cfg.len_wave);
    } else if (cfg.multimodal) {  // MultiModalCVAE.__init__ (hippie/model.py:352-395)
      spec_encoder("encoder_mod1", z);
      spec_encoder("encoder_mod2", z);
      spec_linear("fusion_encoder.0", 2 * z, 4 * z + 2 * h);
      spec_bn("fusion_encoder.1", 2 * z);
      spec_linear("fusion_encoder.3", z, 2 * z);
      add_param("source_embedding.weight", 2, cfg.num_sources, h, 0, HIPPIE_LAYOUT_NATIVE);
      cls_emb = add_param("class_embedding.weight", 2, cfg.num_classes, h, 0, HIPPIE_LAYOUT_NATIVE);
      spec_linear("z_mean", z, z);
      spec_linear("z_log_var", z, z);
      for (const char* m : {"decoder_fc_mod1", "decoder_fc_mod2"}) {
        spec_linear(std::string(m) + ".0", 2 * z, z + 2 * h);
        spec_linear(std::string(m) + ".2", 2 * z, 2 * z);
        spec_bn(std::string(m) + ".3", 2 * z);
      }
      spec_decoder("decoder_mod1", z, cfg.len_wave);
      spec_decoder("decoder_mod2", z, cfg.len_isi);
    } else {  // hippieUnimodalCVAE.__init__ (hippie/model.py:13-44)
      spec_encoder("encoder", z);
      spec_linear("encoder_fc.0", 2 * z, 2 * z + 2 * h);
      spec_bn("encoder_fc.1", 2 * z);
      spec_linear("encoder_fc.3", z, 2 * z);
      spec_bn("encoder_fc.4", z);
      add_param("source_embedding.weight", 2, cfg.num_sources, h, 0, HIPPIE_LAYOUT_NATIVE);
      cls_emb = add_param("class_embedding.weight", 2, cfg.num_classes, h, 0, HIPPIE_LAYOUT_NATIVE);
      spec_linear("z_mean", z, z);
      spec_linear("z_log_var", z, z);
      spec_linear("decoder_fc.0", 2 * z, z + 2 * h);
      spec_linear("decoder_fc.2", 2 * z, 2 * z);
      spec_bn("decoder_fc.3", 2 * z);
      spec_decoder("decoder", z, cfg.len_wave);
    }
    param_floats = (param_floats + 63) & ~(int64_t)63;
  }

  bool pair_mode() const { return cfg.conv_path != 1; }
  // kinds: kF32 = fp32 tensor, kPlanes = fp16 pair planes (only materialised on the tcgen05 path), kSlot = scale slot
  int act(const std::string& name_, int L, int C, int kinds = kF32) {
    std::string name = name_;
    for (const char* pre : {"g:.", "d:."})  // gradient tensors of a prefix-less backbone
      if (name.rfind(pre, 0) == 0) name = name.substr(0, 2) + name.substr(3);
    name = nm(name);
    Act a;
    a.name = name, a.L = L, a.C = C;
    const int64_t rows = (int64_t)cfg.max_batch * (L + 2) + 2;  // + one finite guard row on either side
    if (!pair_mode()) kinds = kF32;
    if (kinds & kF32) a.off = take(rows * C) + C;
    if (kinds & kPlanes) {
      a.pstride = rows * C;
      a.poff = take(rows * C) + C / 2;  // 2 planes x 2 bytes = rows * C floats; row 0 starts C halfs in
    }
    if (kinds & kSlot) a.slot = slots_off + 4 * (n_slots++);
    acts.push_back(a);
    return (int)acts.size() - 1;
  }
  uint16_t* PL(int i) { return reinterpret_cast<uint16_t*>(ws + acts[i].poff); }
  // statistics partials of the BatchNorm that follows conv output `a`: tiles of >= 64 logical rows
  void bn_attach(int bn, int a) {
    bns[bn].part_off = take((((int64_t)cfg.max_batch * acts[a].L + 63) / 64 + 1) * acts[a].C * 2);
  }
  int grad_of(int a) {
    if (cfg.inference_only) return -1;
    auto it = gact.find(a);
    if (it != gact.end()) return it->second;
    int g = act("g:" + acts[a].name, acts[a].L, acts[a].C);
    gact[a] = g;
    return g;
  }
  int gact_or(int a) { return cfg.inference_only ? -1 : act("d:" + acts[a].name, acts[a].L, acts[a].C); }
  Conv mkconv(const std::string& n, int cout, int cin, int k, int stride, bool bias, bool need_wt) {
    Conv c;
    c.w = pix(n + ".weight");
    c.b = bias ? pix(n + ".bias") : -1;
    c.cin = cin, c.cout = cout, c.k = k, c.stride = stride;
    c.id = n_convs++;
    if (need_wt && !cfg.inference_only && !pair_mode()) {
      c.wt_off = take((int64_t)cout * cin * k);
      WtEntry e;
      e.w_off = params[c.w].off, e.wt_off = c.wt_off, e.cout = cout, e.cin = cin, e.k = k;
      wt_table.push_back(e);
    }
    return c;
  }
  static int conv_len(int L, int k, int s, int p) { return (L + 2 * p - k) / s + 1; }

  void build_encoder(Encoder& E, const std::string& p, int Lin, bool train_tensors) {
    const int z = cfg.z_dim;
    E.prefix = p, E.Lin = Lin, E.L0 = conv_len(Lin, 3, 2, 1);
    E.stem_w = pix(p + ".conv1.weight");
    E.bn0 = bix(p + ".bn1");
    E.c0 = act(p + ".conv1", E.L0, 64);
    bn_attach(E.bn0, E.c0);
    E.a0 = act(p + ".stem", E.L0, 64, kF32 | kPlanes);
    if (train_tensors) E.dc0 = gact_or(E.c0);
    int x = E.a0, L = E.L0, in_planes = 64;
    const int planes[4] = {64, 128, 256, 512}, strides[4] = {1, 2, 2, 2};
    for (int li = 0; li < 4; ++li)
      for (int bi = 0; bi < 2; ++bi) {
        EncBlock& b = E.blk[li * 2 + bi];
        const int s = bi == 0 ? strides[li] : 1, out = in_planes * s;
        const std::string q = p + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
        const int Lout = conv_len(L, 3, s, 1);
        b.down = s != 1;
        b.c1 = mkconv(q + ".conv1", out, in_planes, 3, s, false, true);
        b.c2 = mkconv(q + ".conv2", out, out, 3, 1, false, true);
        b.bn1 = bix(q + ".bn1"), b.bn2 = bix(q + ".bn2");
        b.x = x;
        b.c1o = act(q + ".conv1", Lout, out);
        b.a1 = act(q + ".a1", Lout, out, kF32 | kPlanes);
        b.c2o = act(q + ".conv2", Lout, out);
        if (b.down) {
          b.cs = mkconv(q + ".shortcut.0", out, in_planes, 1, s, false, true);
          b.bns = bix(q + ".shortcut.1");
          b.cso = act(q + ".shortcut", Lout, out);
        }
        b.out = act(q, Lout, out, (li * 2 + bi) < 7 ? (kF32 | kPlanes) : kF32);  // the last block feeds the pooling only
        bn_attach(b.bn1, b.c1o), bn_attach(b.bn2, b.c2o);
        if (b.down) bn_attach(b.bns, b.cso);
        if (train_tensors) {
          grad_of(b.x), grad_of(b.out);
          b.g_a1 = act("g:" + q + ".a1", Lout, out);
          // gradients w.r.t. conv outputs are GEMM operands only: scaled fp16 pair planes on the tcgen05 path
          b.dc2 = act("d:" + q + ".conv2", Lout, out, kPlanes | kSlot);
          // gradients of stride-2 convs are stored zero-dilated at the INPUT resolution, so their
          // dgrad / wgrad run as stride-1 problems (DESIGN.md "Backward of strided convs")
          b.dc1 = act("d:" + q + ".conv1", b.down ? L : Lout, out, kPlanes | kSlot);
          if (b.down) b.dcs = act("d:" + q + ".shortcut", L, out, kPlanes | kSlot);
        }
        x = b.out, L = Lout, in_planes = planes[li];
      }
    E.lin_w = pix(p + ".linear.weight"), E.lin_b = pix(p + ".linear.bias");
    E.pooled = take((int64_t)cfg.max_batch * 512);
    E.h = take((int64_t)cfg.max_batch * 2 * z);
    E.dh = take((int64_t)cfg.max_batch * 2 * z);
  }

  void build_decoder(Decoder& D, const std::string& p, int Lo, bool train_tensors) {
    const int z = cfg.z_dim;
    D.prefix = p, D.Lo = Lo;
    D.lin_w = pix(p + ".linear.weight"), D.lin_b = pix(p + ".linear.bias");
    D.t0 = act(p + ".linear", 4, 512, kF32 | kPlanes);
    int x = D.t0, L = 4, in_planes = 512;
    const int lis[4] = {4, 3, 2, 1}, planes[4] = {256, 128, 64, 64}, strides[4] = {2, 2, 2, 1};
    for (int i = 0; i < 4; ++i) {
      for (int bi = 0; bi < 2; ++bi) {
        DecBlock& b = D.blk[i * 2 + bi];
        const int s = bi == 0 ? 1 : strides[i], out = in_planes / s;
        const std::string q = p + ".layer" + std::to_string(lis[i]) + "." + std::to_string(bi);
        b.up = s != 1;
        const int Lout = L * s;
        b.x = x;
        b.c2 = mkconv(q + ".conv2", in_planes, in_planes, 3, 1, false, true);
        b.bn2 = bix(q + ".bn2"), b.bn1 = bix(q + ".bn1");
        b.c2o = act(q + ".conv2", L, in_planes);
        b.a2 = act(q + ".a2", L, in_planes, b.up ? kF32 : (kF32 | kPlanes));
        if (b.up) {
          b.c1 = mkconv(q + ".conv1.conv", out, in_planes, 3, 1, true, true);
          b.cs = mkconv(q + ".shortcut.0.conv", out, in_planes, 3, 1, true, true);
          b.bns = bix(q + ".shortcut.1");
          // the up-sampled copies are conv inputs only: pair planes, no fp32 tensor, on the tcgen05 path
          b.a2_up = act(q + ".a2_up", Lout, in_planes, kPlanes);
          b.x_up = act(q + ".x_up", Lout, in_planes, kPlanes);
          b.cso = act(q + ".shortcut", Lout, out);
        } else {
          b.c1 = mkconv(q + ".conv1", out, in_planes, 3, 1, false, true);
        }
        b.c1o = act(q + ".conv1", Lout, out);
        b.out = act(q, Lout, out, (i * 2 + bi) < 7 ? (kF32 | kPlanes) : kF32);  // the last block feeds the tail kernel
        bn_attach(b.bn2, b.c2o), bn_attach(b.bn1, b.c1o);
        if (b.up) bn_attach(b.bns, b.cso);
        if (train_tensors) {
          grad_of(b.x), grad_of(b.out);
          b.dc2 = act("d:" + q + ".conv2", L, in_planes, kPlanes | kSlot);
          b.dc1 = act("d:" + q + ".conv1", Lout, out, kPlanes | kSlot);
          if (b.up) {
            b.dcs = act("d:" + q + ".shortcut", Lout, out, kPlanes | kSlot);
            b.g_a2_up = act("g:" + q + ".a2_up", Lout, in_planes);
            b.g_x_up = act("g:" + q + ".x_up", Lout, in_planes);
          } else {
            b.g_a2 = act("g:" + q + ".a2", L, in_planes);
          }
        }
        x = b.out, L = Lout;
      }
      in_planes = planes[i];
    }
    for (int k = 0; k + 1 < 8; ++k)
      if (D.blk[k + 1].up) D.blk[k].out_up = D.blk[k + 1].x_up;
    D.wc = pix(p + ".conv1.conv.weight"), D.bc = pix(p + ".conv1.conv.bias");
    D.lo_w = pix(p + ".linear_out.weight"), D.lo_b = pix(p + ".linear_out.bias");
    const int64_t mb = cfg.max_batch;
    D.d = take(mb * 2 * z), D.dd = take(mb * 2 * z), D.gx0 = take(mb * 512);
    D.y = take(mb * 64), D.ddec = take(mb * Lo), D.dy = take(mb * 64), D.dec = take(mb * Lo);
    D.part = take(((mb + 3) / 4 + 1) * 196);
  }

  int64_t off_of(const std::string& n) { return params[pix(n)].off; }
  void build() {
    build_spec();
    slots_off = take(4 * kMaxSlots);
    const bool tr = !cfg.inference_only && full_model();  // stand-alone backbones are forward-only
    const int z = cfg.z_dim;
    flags_off = take(64);
    if (cfg.multimodal == HIPPIE_KIND_ENCODER) {
      n_enc = 1, n_dec = 0;
      build_encoder(enc[0], "", cfg.len_wave, false);
    } else if (cfg.multimodal == HIPPIE_KIND_DECODER) {
      n_enc = 0, n_dec = 1;
      build_decoder(dec[0], "", cfg.len_wave, false);
    } else if (cfg.multimodal) {
      n_enc = n_dec = 2;
      build_encoder(enc[0], "encoder_mod1", cfg.len_wave, tr);
      build_encoder(enc[1], "encoder_mod2", cfg.len_isi, tr);
      build_decoder(dec[0], "decoder_mod1", cfg.len_wave, tr);
      build_decoder(dec[1], "decoder_mod2", cfg.len_isi, tr);
    } else {
      n_enc = n_dec = 1;
      build_encoder(enc[0], "encoder", cfg.len_wave, tr);
      build_decoder(dec[0], "decoder", cfg.len_wave, tr);
    }
    // head parameter map
    HeadParams& H = headp;
    memset(&H, 0xff, sizeof(H));  // all -1
    if (full_model()) {
      const std::string f = cfg.multimodal ? "fusion_encoder" : "encoder_fc";
      H.f0_w = off_of(f + ".0.weight"), H.f0_b = off_of(f + ".0.bias");
      grad_split = H.f0_w;
      for (int e = 0; e < 2; ++e) {
        if (e < n_enc) {
          const std::string pre = cfg.multimodal ? "encoder_mod" + std::to_string(e + 1) : "encoder";
          enc_begin[e] = off_of(pre + ".conv1.weight"), enc_deep[e] = off_of(pre + ".layer3.0.conv1.weight");
        } else {
          enc_begin[e] = enc_deep[e] = grad_split;
        }
      }
      enc_end[0] = n_enc == 2 ? enc_begin[1] : grad_split, enc_end[1] = grad_split;
      H.fbn_g = off_of(f + ".1.weight"), H.fbn_b = off_of(f + ".1.bias");
      H.f3_w = off_of(f + ".3.weight"), H.f3_b = off_of(f + ".3.bias");
      H.fbn_run = bns[bix(f + ".1")].run_off, H.fbn_cnt = bix(f + ".1");
      if (!cfg.multimodal) {
        H.ebn_g = off_of(f + ".4.weight"), H.ebn_b = off_of(f + ".4.bias");
        H.ebn_run = bns[bix(f + ".4")].run_off, H.ebn_cnt = bix(f + ".4");
      }
      H.src_emb = off_of("source_embedding.weight"), H.cls_emb = off_of("class_embedding.weight");
      H.zm_w = off_of("z_mean.weight"), H.zm_b = off_of("z_mean.bias");
      H.zv_w = off_of("z_log_var.weight"), H.zv_b = off_of("z_log_var.bias");
      for (int m = 0; m < n_dec; ++m) {
        const std::string d = cfg.multimodal ? "decoder_fc_mod" + std::to_string(m + 1) : "decoder_fc";
        H.d0_w[m] = off_of(d + ".0.weight"), H.d0_b[m] = off_of(d + ".0.bias");
        H.d2_w[m] = off_of(d + ".2.weight"), H.d2_b[m] = off_of(d + ".2.bias");
        H.dbn_g[m] = off_of(d + ".3.weight"), H.dbn_b[m] = off_of(d + ".3.bias");
        H.dbn_run[m] = bns[bix(d + ".3")].run_off, H.dbn_cnt[m] = bix(d + ".3");
      }
    }
    head_scratch = take(head_scratch_floats(z, cfg.class_hidden_dim, cfg.max_batch));
    scal_off = take(64 + kHeadMaxCtas);  // [0],[1] squared-error sums, [64..) KL partials per head CTA
    // partial-sum buffers.  BatchNorm statistics: [tiles of >= 64 logical rows][C][2]; the stem's weight-gradient
    // partials ([tiles of 128 rows][192]) also fit.  BatchNorm backward: at most kBnBwdMaxChunks chunks x C x 3.
    part_floats = 4096;
    for (auto& a : acts)
      part_floats = std::max<int64_t>(part_floats, (((int64_t)cfg.max_batch * a.L + 63) / 64 + 1) * a.C * 2);
    bpart_floats = (int64_t)(kBnBwdMaxChunks + 1) * 512 * 6;
    for (int i = 0; i < 2; ++i) part_off[i] = take(part_floats), bpart_off[i] = take(bpart_floats);
    {  // per-channel BatchNorm totals (double atomics), one contiguous region zeroed once per step
      int64_t n = 0;
      for (auto& b : bns) b.tot_off = n, n += (4 * kBnFwdTotCopies + 6 * kBnBwdTotCopies) * (int64_t)b.C;  // (2 + 3) x copies x C doubles
      tot_floats = n;
      tot_off = take(tot_floats);
      for (auto& b : bns) b.tot_off += tot_off;
    }
    adam_part = take(1024);
    {
      const int64_t mb = cfg.max_batch;
      st_x1 = take(mb * cfg.len_wave), st_x2 = take(mb * std::max(cfg.len_isi, 1)), st_src = take(mb * 2);
      st_cls = take(mb * 2), st_eps = take(mb * z), st_scal = take(8), st_enc = take(mb * z), st_mu = take(mb * z);
      st_lv = take(mb * z), st_d1 = take(mb * cfg.len_wave), st_d2 = take(mb * std::max(cfg.len_isi, 1));
    }
    if (pair_mode()) {
      wp_off = take(param_floats);  // 2 planes x 2 bytes per parameter
    }
    for (auto& b : bns) {
      b.coef_off = take(8 * (int64_t)b.C);
      BnEvalEntry e;
      e.gamma_off = params[b.gamma].off, e.beta_off = params[b.beta].off, e.run_off = b.run_off, e.coef_off = b.coef_off;
      e.C = b.C;
      bn_table.push_back(e);
    }
    wt_table_off = take((int64_t)(wt_table.size() * sizeof(WtEntry) + 3) / 4 + 4);
    bn_table_off = take((int64_t)(bn_table.size() * sizeof(BnEvalEntry) + 3) / 4 + 4);
  }

  // ---- helpers ---------------------------------------------------------------------------------
  float* A(int i) { return acts[i].off >= 0 ? ws + acts[i].off : nullptr; }
  uint16_t* WP() { return reinterpret_cast<uint16_t*>(ws + wp_off); }
  float* slot(int i) { return ws + acts[i].slot; }
  float* coef(int bn) { return ws + bns[bn].coef_off; }
  float* Pp(int p) { return P + params[p].off; }
  float* Gp(int p) { return G + params[p].off; }

  // eval mode on the tcgen05 path: the BatchNorm that follows a conv (running statistics), the residual, the LeakyReLU
  // and the pair-plane conversion run in the conv's epilogue -- no bn_apply launches, no fp32 round trip of the conv output
  bool fold_eval(bool train) const { return !train && use_tc && !profiling; }
  // out = lrelu_slope(bn(conv(in)) + res): `out` is the ACTIVATION tensor (fp32 written only when the caller needs it)
  void conv_fwd_folded(const Conv& cv, int in, int bn, int res, int out, int out_up, float slope, bool write_f32, int B,
                       Branch& br) {
    EvalFold f{};
    f.coef = coef(bn), f.res = res >= 0 ? A(res) : nullptr, f.slope = slope, f.write_f32 = write_f32 ? 1 : 0;
    f.flags = flags();
    if (acts[out].poff >= 0) f.out_p = PL(out), f.out_ps = acts[out].pstride;
    if (out_up >= 0 && acts[out_up].poff >= 0) f.up_p = PL(out_up), f.up_ps = acts[out_up].pstride;
    conv_fwd(cv, in, out, -1, B, false, br, &f);
  }
  void conv_fwd(const Conv& cv, int in, int out, int bn, int B, bool train, Branch& br, const EvalFold* fold = nullptr) {
    ConvGemm g{};
    g.A = A(in), g.W = Pp(cv.w), g.bias = cv.b >= 0 ? Pp(cv.b) : nullptr, g.C = A(out);
    g.part = (train && bn >= 0) ? ws + bns[bn].part_off : nullptr;
    if (train && bn >= 0 && use_tc) g.tot = reinterpret_cast<double*>(ws + bns[bn].tot_off), bns[bn].totals = true;
    g.M = B * acts[out].L, g.N = cv.cout, g.K = cv.k * cv.cin, g.Lout = acts[out].L;
    g.in_rows = acts[in].L + 2, g.in_stride = cv.stride, g.in_off = cv.k == 3 ? 0 : 1, g.in_C = cv.cin;
    g.out_rows = acts[out].L + 2, g.out_off = 1, g.out_lstride = 1, g.accumulate = 0;
    cudaEvent_t pe = prof_begin(br);
    int tile;
    if (use_tc) {
      ConvMaps& m = cmaps[cv.id];
      const uint16_t* wpl = WP() + params[cv.w].off;
      if (!m.fwd_ready) {
        bool ok = pair_make_act_map(&m.a_fwd, PL(in), acts[in].pstride, kPairF16, g.in_C, g.K, g.Lout, g.in_rows,
                                    g.in_stride, g.in_off, cfg.max_batch);
        ok = ok && pair_make_w_map(&m.w_k, wpl, param_floats, kPairF16, g.N, g.K, pair_pick_bn(B, g.N, g.Lout, sm_count));
        if (!ok) return (void)(err = "cuTensorMapEncodeTiled failed (conv forward)", failed = true);
        m.fwd_ready = true;
      }
      if (!fold && (m.c_fwd_ptr != g.C || m.c_fwd_B != B)) {  // fp32 output: tiles leave through TMA stores
        if (!pair_make_out_map(&m.c_fwd, g.C, g.N, g.Lout, g.out_rows, g.out_off, B))
          return (void)(err = "cuTensorMapEncodeTiled failed (conv forward output)", failed = true);
        m.c_fwd_ptr = g.C, m.c_fwd_B = B;
      }
      const int bn_tile = pair_pick_bn(B, g.N, g.Lout, sm_count);
      PairOpts o{1.f / kWeightPairScale, kPairF16, kPairF16, 0, cv.k};
      o.fold = fold;
      if (!fold && tma_epilogue) o.out_map = &m.c_fwd;
      tile = launch_conv_pair(g, m.a_fwd, m.w_k, bn_tile, B, o, br.st);
    } else {
      tile = launch_conv_gemm_simt(g, br.st);
    }
    prof_end(pe, 0, 2.0 * g.M * g.N * g.K, br);
    ++launches;
    if (train && bn >= 0) bns[bn].ntiles = (g.M + tile - 1) / tile, bns[bn].tile_rows = tile, bns[bn].M = g.M;
  }
  // the apply kernel turns the partials the conv left into coefficients (and running statistics) itself
  BnFinalize finalize_args(int bn) {
    BnFinalize f{};
    f.part = ws + bns[bn].part_off, f.C = bns[bn].C;
    if (bns[bn].totals) f.tot = reinterpret_cast<const double*>(ws + bns[bn].tot_off);
    f.set_shape(bns[bn].ntiles, bns[bn].tile_rows, bns[bn].M);
    f.gamma = Pp(bns[bn].gamma), f.beta = Pp(bns[bn].beta);
    f.run_mean = bn_mean + bns[bn].run_off, f.run_var = bn_var + bns[bn].run_off, f.run_count = bn_count + bn;
    f.coef = coef(bn);
    return f;
  }
  void apply(int c, int bn, int r, int rbn, int out, int out_up, int B, bool train, Branch& br) {
    BnApply a{};
    a.train = train ? 1 : 0;
    if (train) {
      a.fin = finalize_args(bn);
      if (rbn >= 0) a.rfin = finalize_args(rbn);
    }
    a.c = A(c), a.coef = coef(bn), a.r = r >= 0 ? A(r) : nullptr, a.rcoef = rbn >= 0 ? coef(rbn) : nullptr;
    a.out = A(out), a.out_up = out_up >= 0 ? A(out_up) : nullptr;
    a.B = B, a.L = acts[c].L, a.C = acts[c].C, a.slope = kSlopeBackbone;
    a.flags = flags();
    if (use_tc) {
      if (acts[out].poff >= 0) a.out_p = PL(out), a.out_ps = acts[out].pstride;
      if (out_up >= 0 && acts[out_up].poff >= 0) a.up_p = PL(out_up), a.up_ps = acts[out_up].pstride;
    }
    cudaEvent_t pe = prof_begin(br);
    launch_bn_apply(a, sm_count, br.st);
    prof_end(pe, 3, 0.0, br);
    ++launches;
  }
  // dgrad as a stride-1 convolution of the (dilated) output gradient with the transposed weights
  void dgrad(const Conv& cv, int dy, int gx, bool accumulate, int B, Branch& br) {
    ConvGemm g{};
    g.A = A(dy), g.W = cv.wt_off >= 0 ? ws + cv.wt_off : nullptr, g.bias = nullptr, g.C = A(gx), g.part = nullptr;
    g.M = B * acts[gx].L, g.N = cv.cin, g.K = cv.k * cv.cout, g.Lout = acts[gx].L;
    g.in_rows = acts[dy].L + 2, g.in_stride = 1, g.in_off = cv.k == 3 ? 0 : 1, g.in_C = cv.cout;
    g.out_rows = acts[gx].L + 2, g.out_off = 1, g.out_lstride = 1, g.accumulate = accumulate ? 1 : 0;
    cudaEvent_t pe = prof_begin(br);
    if (use_tc) {
      ConvMaps& m = cmaps[cv.id];
      if (!m.dg_ready) {
        bool ok = pair_make_act_map(&m.a_dg, PL(dy), acts[dy].pstride, kPairF16, g.in_C, g.K, g.Lout, g.in_rows, 1,
                                    g.in_off, cfg.max_batch);
        ok = ok && pair_make_wmn_map(&m.w_mn, WP() + params[cv.w].off, param_floats, kPairF16, cv.cout, cv.cin, cv.k);
        if (!ok) return (void)(err = "cuTensorMapEncodeTiled failed (conv dgrad)", failed = true);
        m.dg_ready = true;
      }
      if (m.c_dg_ptr != g.C || m.c_dg_B != B) {
        if (!pair_make_out_map(&m.c_dg, g.C, g.N, g.Lout, g.out_rows, g.out_off, B))
          return (void)(err = "cuTensorMapEncodeTiled failed (conv dgrad output)", failed = true);
        m.c_dg_ptr = g.C, m.c_dg_B = B;
      }
      const int bn_tile = pair_pick_bn(B, g.N, g.Lout, sm_count);
      PairOpts o{1.f / kWeightPairScale, kPairF16, kPairF16, 1, cv.k};
      o.dyn_scale = slot(dy) + 3;
      if (tma_epilogue) o.out_map = &m.c_dg;
      launch_conv_pair(g, m.a_dg, m.w_mn, bn_tile, B, o, br.st);
    } else {
      launch_conv_gemm_simt(g, br.st);
    }
    // algorithmic FLOP: a stride-2 conv's output gradient is stored zero-dilated, half of the M x K products are zeros
    prof_end(pe, 1, 2.0 * g.M * g.N * g.K / (cv.stride == 2 ? 2.0 : 1.0), br);
    ++launches;
  }
  // The shortcut convolution of a down-sampling / up-sampling block (forward: 1x1 or resize conv of the block input;
  // backward: its dgrad) depends only on the block's input (forward) or on the block's first BatchNorm backward, so it
  // runs on the branch's helper stream beside conv -> BatchNorm -> conv of the main path and rejoins before the
  // kernel that needs it: six fewer GEMMs per branch on each of the forward and backward critical chains.
  Branch fork_helper(Branch& br) {
    Branch hb = br;
    if (!br.hst) return hb;
    hb.st = br.hst;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, br.st);
    cudaStreamWaitEvent(br.hst, e, 0);
    return hb;
  }
  void join_helper(Branch& br) {
    if (!br.hst) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, br.hst);
    cudaStreamWaitEvent(br.st, e, 0);
  }
  // the weight-gradient stream of the branch waits for everything issued on the branch stream so far
  void side_wait(Branch& br) {
    if (!br.wst || br.wst == br.st) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, br.st);
    cudaStreamWaitEvent(br.wst, e, 0);
  }
  void wgrad(const Conv& cv, int dy, int x, int B, Branch& br) {
    WgradGemm g{};
    g.dY = A(dy), g.X = A(x), g.dW = Gp(cv.w);
    g.M = cv.cout, g.N = cv.k * cv.cin, g.R = B * (acts[dy].L + 2), g.Cin = cv.cin, g.roff = cv.k == 3 ? -1 : 0;
    cudaStream_t wst = br.wst ? br.wst : br.st;
    side_wait(br);  // dY is final once the stream reaches this point; everything the wgrad reads stays untouched
    cudaEvent_t pe = prof_begin(br);
#ifdef HP_EXPERIMENTS  // knock-out runs of DESIGN 4.2 / 4.3 (wrong results): only in builds made with EXTRA=-DHP_EXPERIMENTS
    static const int dbg_skip = getenv("HIPPIE_B200_DEBUG_SKIP") ? atoi(getenv("HIPPIE_B200_DEBUG_SKIP")) : 0;
    if (dbg_skip & 1) return;  // no weight gradients
#endif
    if (use_tc) {
      ConvMaps& m = cmaps[cv.id];
      if (m.wg_B != B) {
        bool ok = pair_make_rows_map(&m.wg_dy, PL(dy), acts[dy].pstride, kPairF16, g.M, g.M, g.R);
        ok = ok && pair_make_rows_map(&m.wg_x, PL(x) + (int64_t)g.roff * g.Cin, acts[x].pstride, kPairF16, g.Cin, g.N, g.R);
        if (!ok) return (void)(err = "cuTensorMapEncodeTiled failed (conv wgrad)", failed = true);
        m.wg_B = B;
      }
      if (!m.dw_ready) {
        if (!pair_make_dw_map(&m.dw, g.dW, g.M, g.N))
          return (void)(err = "cuTensorMapEncodeTiled failed (conv wgrad output)", failed = true);
        m.dw_ready = true;
      }
      PairOpts o{1.f, kPairF16, kPairF16, 1, cv.k};
      o.dyn_scale = slot(dy) + 3;
      launch_wgrad_pair(g, m.wg_dy, m.wg_x, m.dw, pair_pick_bn(B, g.N, 1, sm_count), sm_count, o, wst);
    } else {
      launch_wgrad_simt(g, sm_count, wst);
    }
    // algorithmic FLOPs: only the B*Lout real output rows contribute (pad / dilation rows are zeros)
    prof_end(pe, 2, 2.0 * (double)g.M * g.N * (double)B * acts[dy].L / (cv.stride == 2 ? 2.0 : 1.0), br);
    ++launches;
  }
  void bn_bwd(int g, bool g_up, int out, int c, int bn, int cs, int bnsi, int dc, int dil, int dcs, int dil_s, int gres,
              int B, Branch& br) {
    BnBwd a{};
    a.g = A(g), a.g_up = g_up ? 1 : 0, a.out = A(out), a.c = A(c), a.coef = coef(bn);
    a.cs = cs >= 0 ? A(cs) : nullptr, a.coef_s = cs >= 0 ? coef(bnsi) : nullptr;
    a.part = br.bpart, a.B = B, a.L = acts[out].L, a.C = acts[out].C, a.slope = kSlopeBackbone;
    a.tot = reinterpret_cast<double*>(ws + bns[bn].tot_off) + 2 * kBnFwdTotCopies * (int64_t)bns[bn].C;
    a.inv_n = 1.0 / ((double)B * acts[out].L);
    a.gamma = Pp(bns[bn].gamma), a.dgamma = Gp(bns[bn].gamma), a.dbeta = Gp(bns[bn].beta);
    if (cs >= 0) a.gamma_s = Pp(bns[bnsi].gamma), a.dgamma_s = Gp(bns[bnsi].gamma), a.dbeta_s = Gp(bns[bnsi].beta);
    a.dc = A(dc), a.dil = dil, a.Ld = acts[dc].L;
    if (cs >= 0) a.dcs = A(dcs), a.dil_s = dil_s, a.Ld_s = acts[dcs].L;
    if (use_tc && acts[dc].poff >= 0) a.dc_p = PL(dc), a.dc_ps = acts[dc].pstride, a.dc_slot = slot(dc);
    if (use_tc && cs >= 0 && acts[dcs].poff >= 0) a.dcs_p = PL(dcs), a.dcs_ps = acts[dcs].pstride, a.dcs_slot = slot(dcs);
    a.gres = gres >= 0 ? A(gres) : nullptr;
    cudaEvent_t pe = prof_begin(br);
    launch_bn_bwd(a, sm_count, br.st);
    prof_end(pe, 4, 0.0, br);
    launches += kBnBwdLaunches;
  }

  void encoder_fwd(Encoder& E, const float* x, int B, bool train, Branch& br) {
    cudaEvent_t pe7 = prof_begin(br);
    launch_stem_fwd(x, Pp(E.stem_w), A(E.c0), train ? ws + bns[E.bn0].part_off : nullptr, B, E.Lin, E.L0, br.st);
    ++launches;
    prof_end(pe7, 7, 0.0, br);
    if (train) bns[E.bn0].ntiles = (B * E.L0 + 127) / 128, bns[E.bn0].tile_rows = 128, bns[E.bn0].M = B * E.L0;
    apply(E.c0, E.bn0, -1, -1, E.a0, -1, B, train, br);
    for (int i = 0; i < 8; ++i) {
      EncBlock& b = E.blk[i];
      if (fold_eval(train)) {
        if (b.down) {  // shortcut: BatchNorm, no LeakyReLU
          Branch hb = fork_helper(br);
          conv_fwd_folded(b.cs, b.x, b.bns, -1, b.cso, -1, 1.f, true, B, hb);
        }
        conv_fwd_folded(b.c1, b.x, b.bn1, -1, b.a1, -1, kSlopeBackbone, false, B, br);  // a1: planes only
        if (b.down) join_helper(br);
        conv_fwd_folded(b.c2, b.a1, b.bn2, b.down ? b.cso : b.x, b.out, -1, kSlopeBackbone, true, B, br);
        continue;
      }
      if (b.down) {
        Branch hb = fork_helper(br);
        conv_fwd(b.cs, b.x, b.cso, b.bns, B, train, hb);
      }
      conv_fwd(b.c1, b.x, b.c1o, b.bn1, B, train, br);
      apply(b.c1o, b.bn1, -1, -1, b.a1, -1, B, train, br);
      conv_fwd(b.c2, b.a1, b.c2o, b.bn2, B, train, br);
      if (b.down) {
        join_helper(br);
        apply(b.c2o, b.bn2, b.cso, b.bns, b.out, -1, B, train, br);
      } else {
        apply(b.c2o, b.bn2, b.x, -1, b.out, -1, B, train, br);
      }
    }
    const int last = E.blk[7].out;
    cudaEvent_t pe8 = prof_begin(br);
    launch_pool_linear_fwd(A(last), B, acts[last].L, 512, Pp(E.lin_w), Pp(E.lin_b), 2 * cfg.z_dim, ws + E.pooled,
                           ws + E.h, br.st);
    prof_end(pe8, 8, 0.0, br);
    ++launches;
  }
  // phase 0 = whole backward; 1 = deep half (Linear, layer4, layer3); 2 = shallow half (layer2, layer1, stem)
  void encoder_bwd(Encoder& E, const float* x, int B, Branch& br, int phase = 0) {
    const int last = E.blk[7].out;
    if (phase != 2) {
      cudaEvent_t pe9 = prof_begin(br);
      side_wait(br);  // dh and pooled are final: the Linear's weight gradient runs beside the backbone backward
      launch_linear_wgrad(ws + E.dh, 2 * cfg.z_dim, ws + E.pooled, 512, B, 512, 2 * cfg.z_dim, Gp(E.lin_w), Gp(E.lin_b),
                          br.wst ? br.wst : br.st);
      launch_pool_linear_bwd_x(ws + E.dh, Pp(E.lin_w), B, acts[last].L, 512, 2 * cfg.z_dim, A(gact.at(last)), br.st);
      prof_end(pe9, 9, 0.0, br);
      launches += 2;
    }
    const int i_hi = phase == 2 ? 3 : 7, i_lo = phase == 1 ? 4 : 0;
    for (int i = i_hi; i >= i_lo; --i) {
      EncBlock& b = E.blk[i];
      const int gx = gact.at(b.x), gout = gact.at(b.out);
      bn_bwd(gout, false, b.out, b.c2o, b.bn2, b.down ? b.cso : -1, b.bns, b.dc2, 1, b.dcs, 2, b.down ? -1 : gx, B, br);
      if (b.down) {  // the shortcut's dgrad (first writer of gx) beside dgrad(conv2) -> BatchNorm backward
        Branch hb = fork_helper(br);
        dgrad(b.cs, b.dcs, gx, false, B, hb);
        wgrad(b.cs, b.dcs, b.x, B, br);
      }
      dgrad(b.c2, b.dc2, b.g_a1, false, B, br);
      wgrad(b.c2, b.dc2, b.a1, B, br);
      bn_bwd(b.g_a1, false, b.a1, b.c1o, b.bn1, -1, -1, b.dc1, b.down ? 2 : 1, -1, 1, -1, B, br);
      if (b.down) join_helper(br);
      dgrad(b.c1, b.dc1, gx, true, B, br);
      wgrad(b.c1, b.dc1, b.x, B, br);
      if (export_ev && i == 4) {  // the deep half of this encoder's gradients is final behind these two points
        for (cudaStream_t st : {br.st, br.wst ? br.wst : br.st}) {
          cudaEvent_t e = next_event();
          cudaEventRecord(e, st);
          deep_ev.push_back(e);
        }
      }
    }
    if (phase == 1) return;
    bn_bwd(gact.at(E.a0), false, E.a0, E.c0, E.bn0, -1, -1, E.dc0, 1, -1, 1, -1, B, br);
    cudaEvent_t pe10 = prof_begin(br);
    const int np = launch_stem_wgrad(x, A(E.dc0), br.part, B, E.Lin, E.L0, br.st);
    launch_reduce_partials(br.part, np, 192, Gp(E.stem_w), 0, br.st);
    prof_end(pe10, 10, 0.0, br);
    launches += 2;
  }

  void decoder_fwd(Decoder& D, int B, bool train, Branch& br) {
    cudaEvent_t pe11 = prof_begin(br);
    launch_dec_linear_fwd(ws + D.d, B, 2 * cfg.z_dim, Pp(D.lin_w), Pp(D.lin_b), 512, A(D.t0),
                          use_tc ? PL(D.t0) : nullptr, acts[D.t0].pstride, flags(), br.st);
    prof_end(pe11, 11, 0.0, br);
    ++launches;
    for (int i = 0; i < 8; ++i) {
      DecBlock& b = D.blk[i];
      if (fold_eval(train)) {
        if (b.up) {
          Branch hb = fork_helper(br);
          conv_fwd_folded(b.cs, b.x_up, b.bns, -1, b.cso, -1, 1.f, true, B, hb);
        }
        conv_fwd_folded(b.c2, b.x, b.bn2, -1, b.a2, b.a2_up, kSlopeBackbone, false, B, br);  // a2 / a2_up: planes only
        if (b.up) {
          join_helper(br);
          conv_fwd_folded(b.c1, b.a2_up, b.bn1, b.cso, b.out, b.out_up, kSlopeBackbone, true, B, br);
        } else {
          conv_fwd_folded(b.c1, b.a2, b.bn1, b.x, b.out, b.out_up, kSlopeBackbone, true, B, br);
        }
        continue;
      }
      if (b.up) {
        Branch hb = fork_helper(br);
        conv_fwd(b.cs, b.x_up, b.cso, b.bns, B, train, hb);
      }
      conv_fwd(b.c2, b.x, b.c2o, b.bn2, B, train, br);
      apply(b.c2o, b.bn2, -1, -1, b.a2, b.a2_up, B, train, br);
      if (b.up) {
        conv_fwd(b.c1, b.a2_up, b.c1o, b.bn1, B, train, br);
        join_helper(br);
        apply(b.c1o, b.bn1, b.cso, b.bns, b.out, b.out_up, B, train, br);
      } else {
        conv_fwd(b.c1, b.a2, b.c1o, b.bn1, B, train, br);
        apply(b.c1o, b.bn1, b.x, -1, b.out, b.out_up, B, train, br);
      }
    }
  }
  DecTail tail_args(Decoder& D, const float* target, float* dec_out, int B, bool train, float loss_w) {
    DecTail t{};
    t.x = A(D.blk[7].out), t.wc = Pp(D.wc), t.bc = Pp(D.bc), t.Wo = Pp(D.lo_w), t.bo = Pp(D.lo_b);
    t.target = target, t.dec = dec_out ? dec_out : ws + D.dec, t.B = B, t.Lo = D.Lo;
    t.part = ws + D.part, t.loss_w = loss_w, t.train = train ? 1 : 0;
    if (train) {
      t.y = ws + D.y, t.ddec = ws + D.ddec, t.dy = ws + D.dy, t.g_x = A(gact.at(D.blk[7].out));
    }
    return t;
  }
  void decoder_tail(Decoder& D, int m, const float* target, float* dec_out, int B, bool train, float loss_w, Branch& br) {
    DecTail t = tail_args(D, target, dec_out, B, train, loss_w);
    cudaEvent_t pe12 = prof_begin(br);
    const int ncta = launch_dec_tail(t, br.st);
    float* sse = ws + scal_off + m;
    if (train)
      launch_dec_tail_reduce(t, ncta, Gp(D.wc), Gp(D.bc), Gp(D.lo_w), Gp(D.lo_b), sse, br.st);
    else
      launch_dec_tail_reduce(t, ncta, nullptr, nullptr, nullptr, nullptr, sse, br.st);
    prof_end(pe12, 12, 0.0, br);
    launches += 2;
  }
  void decoder_bwd(Decoder& D, int B, Branch& br) {
    for (int i = 7; i >= 0; --i) {
      DecBlock& b = D.blk[i];
      const int gx = gact.at(b.x), gout = gact.at(b.out);
      if (b.up) {
        bn_bwd(gout, false, b.out, b.c1o, b.bn1, b.cso, b.bns, b.dc1, 1, b.dcs, 1, -1, B, br);
        {  // the shortcut's dgrad beside dgrad(conv1) -> BatchNorm backward -> dgrad(conv2)
          Branch hb = fork_helper(br);
          dgrad(b.cs, b.dcs, b.g_x_up, false, B, hb);
          wgrad(b.cs, b.dcs, b.x_up, B, br);
        }
        dgrad(b.c1, b.dc1, b.g_a2_up, false, B, br);
        wgrad(b.c1, b.dc1, b.a2_up, B, br);
        bn_bwd(b.g_a2_up, true, b.a2, b.c2o, b.bn2, -1, -1, b.dc2, 1, -1, 1, -1, B, br);
        dgrad(b.c2, b.dc2, gx, false, B, br);
        wgrad(b.c2, b.dc2, b.x, B, br);
        join_helper(br);
        cudaEvent_t pe14 = prof_begin(br);
        launch_pairsum_acc(A(b.g_x_up), A(gx), B, acts[b.x].L, acts[b.x].C, br.st);
        prof_end(pe14, 14, 0.0, br);
        ++launches;
      } else {
        bn_bwd(gout, false, b.out, b.c1o, b.bn1, -1, -1, b.dc1, 1, -1, 1, gx, B, br);
        dgrad(b.c1, b.dc1, b.g_a2, false, B, br);
        wgrad(b.c1, b.dc1, b.a2, B, br);
        bn_bwd(b.g_a2, false, b.a2, b.c2o, b.bn2, -1, -1, b.dc2, 1, -1, 1, -1, B, br);
        dgrad(b.c2, b.dc2, gx, true, B, br);
        wgrad(b.c2, b.dc2, b.x, B, br);
      }
    }
    cudaEvent_t pe13 = prof_begin(br);
    launch_dec_linear_bwd_x(A(gact.at(D.t0)), Pp(D.lin_w), B, 2 * cfg.z_dim, 512, ws + D.gx0, ws + D.dd, br.st);
    side_wait(br);
    // dW[c][f] = sum_b gx0[b][c] * d[b][f];  db[c] = sum_b gx0[b][c]
    launch_linear_wgrad(ws + D.gx0, 512, ws + D.d, 2 * cfg.z_dim, B, 2 * cfg.z_dim, 512, Gp(D.lin_w), Gp(D.lin_b),
                        br.wst ? br.wst : br.st);
    prof_end(pe13, 13, 0.0, br);
    launches += 2;
  }

  HeadArgs head_args(int B, const int64_t* src, const int64_t* cls, const float* eps, bool train, bool decode,
                     float* out_enc, float* out_mu, float* out_logvar, float beta, int zscore) {
    HeadArgs a{};
    a.hp = headp, a.params = P, a.grads = G, a.run_mean = bn_mean, a.run_var = bn_var, a.run_count = bn_count;
    a.z = cfg.z_dim, a.h = cfg.class_hidden_dim, a.n_enc = n_enc, a.n_dec = n_dec;
    a.num_sources = cfg.num_sources, a.num_classes = cfg.num_classes, a.B = B;
    for (int e = 0; e < n_enc; ++e) a.hin[e] = ws + enc[e].h, a.dh[e] = ws + enc[e].dh;
    for (int m = 0; m < n_dec; ++m) a.dout[m] = ws + dec[m].d, a.dd[m] = ws + dec[m].dd;
    a.src = src, a.cls = cls, a.eps = eps, a.scratch = ws + head_scratch;
    a.out_enc = out_enc, a.out_mu = out_mu, a.out_logvar = out_logvar;
    a.kl_sum = ws + scal_off + 64, a.train = train ? 1 : 0, a.decode = decode ? 1 : 0, a.zscore_ddof = zscore;
    a.beta = beta;
    a.flags = flags();
    return a;
  }

  // tcgen05 path: fp16 pair planes of the whole parameter buffer (scaled 2^8), used K-major by the forward convs and
  // MN-major by dgrad.  FP32 path: transposed + tap-flipped copies for dgrad.
  // The planes stay valid between calls: hippie_clip_adamw rewrites the planes of every weight it updates, and whoever
  // else writes the parameter buffer reports it (hippie_params_changed) -- only then is the 64 MB buffer converted again.
  void refresh_weights(bool backward, cudaStream_t main) {
    if (use_tc) {
      if (planes_fresh && planes_keep) return;
      launch_to_pair(P, WP(), param_floats, param_floats, kWeightPairScale, kPairF16, main, flags());
      ++launches;
      planes_fresh = true;
    } else if (backward && !wt_table.empty()) {
      launch_refresh_wt(reinterpret_cast<const WtEntry*>(ws + wt_table_off), (int)wt_table.size(), P, ws, main);
      ++launches;
    }
  }
  void fork(cudaStream_t main) {
    cudaEventRecord(ev_fork, main);
    cudaStreamWaitEvent(side, ev_fork, 0);
  }
  void join(cudaStream_t main) {
    cudaEventRecord(ev_join, side);
    cudaStreamWaitEvent(main, ev_join, 0);
  }
  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
  }
  int validate(int B, const void* x1, const void* x2, const void* src) {
    if (!bound) return fail(-2, "hippie_bind has not been called");
    if (!full_model()) return fail(-7, "stand-alone backbone engine: only hippie_encoder_forward / hippie_decoder_forward apply");
    if (B < 1 || B > cfg.max_batch) return fail(-3, "B outside [1, max_batch]");
    if (!x1 || !src || (cfg.multimodal && !x2)) return fail(-4, "null input pointer");
    return 0;
  }
  int validate_module_call(int B, bool train) {
    if (!bound) return fail(-2, "hippie_bind has not been called");
    if (B < 1 || B > cfg.max_batch) return fail(-3, "B outside [1, max_batch]");
    if (train && B < 2) return fail(-3, "training-mode BatchNorm needs B >= 2");
    return 0;
  }

  // full forward (+ loss, + backward when train)
  int run(bool train, bool backward, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
          const float* eps, int B, float beta, float w1, float w2, float* scalars_out, float* out_enc, float* out_mu,
          float* out_logvar, float* out_dec1, float* out_dec2, cudaStream_t main, int part = -1) {
    launches = 0;
    Branch b0{main, ws + part_off[0], ws + bpart_off[0]}, b1{profiling ? main : side, ws + part_off[1], ws + bpart_off[1]};
    if (backward && !profiling) b0.wst = wside[0], b1.wst = wside[1];
    if (B <= helper_max_batch && !profiling) b0.hst = hside[0], b1.hst = hside[1];
    const float* xin[2] = {x1, x2};
    export_ev = part == 4 && !profiling;  // whole step, slice events exported
    if (part == 4) part = -1;
    if (export_ev) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(main, &cs);
      export_capturing = cs != cudaStreamCaptureStatusNone;
      deep_ev.clear();
    }
    if (part >= 1) {  // later parts of a split step: the encoders' backward pass (1 = whole, 2 = deep half, 3 = shallow half)
      encoders_bwd(xin, B, b0, b1, main, part == 1 ? 0 : part - 1);
      if (failed) return fail(-9, err);
      return check("train_fwd_bwd (part >= 1)");
    }
    float* dec_out[2] = {out_dec1, out_dec2};
    const float lw[2] = {cfg.multimodal ? w1 : 1.f, w2};
    cudaMemsetAsync(ws + scal_off, 0, 64 * sizeof(float), main);
    if (train) cudaMemsetAsync(ws + tot_off, 0, tot_floats * sizeof(float), main);
    if (backward) {
      // (zeroing the 64 MB gradient buffer on a weight-gradient stream under the forward pass was measured slower: the
      // event wait it needs in front of the first gradient writer costs that kernel its programmatic launch edge)
      cudaMemsetAsync(G, 0, param_floats * sizeof(float), main);
      if (use_tc) cudaMemsetAsync(ws + slots_off, 0, 4 * kMaxSlots * sizeof(float), main);
    }
    refresh_weights(backward, main);
    if (!train) {
      launch_bn_eval_coefs(reinterpret_cast<const BnEvalEntry*>(ws + bn_table_off), (int)bn_table.size(), P, bn_mean,
                           bn_var, ws, main);
      ++launches;
    }
    const bool two = n_enc == 2;
    if (two) fork(main);
    encoder_fwd(enc[0], xin[0], B, train, b0);
    if (two) {
      encoder_fwd(enc[1], xin[1], B, train, b1);
      join(main);
    }
    HeadArgs ha = head_args(B, src, cls, eps, train, true, out_enc, out_mu, out_logvar, beta, -1);
    cudaEvent_t pe5 = prof_begin(b0);
    const int n_kl = launch_head_fwd(ha, main);
    prof_end(pe5, 5, 0.0, b0);
    ++launches;
    if (two) fork(main);
    for (int m = 0; m < n_dec; ++m) {
      Branch& br = m == 0 ? b0 : b1;
      decoder_fwd(dec[m], B, train, br);
      decoder_tail(dec[m], m, xin[m], dec_out[m], B, backward, lw[m], br);
      if (backward) decoder_bwd(dec[m], B, br);
    }
    if (two) join(main);
    if (scalars_out) {
      launch_loss_finalize(ws + scal_off + 0, ws + scal_off + 1, ws + scal_off + 64, n_kl, B, dec[0].Lo,
                           cfg.multimodal ? dec[1].Lo : 1, beta, w1, w2, cfg.multimodal, scalars_out, main);
      ++launches;
    }
    if (backward) {
      cudaEvent_t pe6 = prof_begin(b0);
      launch_head_bwd(ha, main);
      prof_end(pe6, 6, 0.0, b0);
      ++launches;
      if (part == 0) {
        join_wgrad(main);  // decoder + head gradients are final when part 0 completes
      } else {
        if (export_ev) {  // latent head + decoders: final from here on
          if (n_dec == 2)
            signal_slice(0, {main, wside[0], wside[1]});
          else
            signal_slice(0, {main, wside[0]});
        }
        encoders_bwd(xin, B, b0, b1, main);
        if (export_ev) {  // xs rejoins the caller's stream (a captured step must end with every forked stream joined)
          cudaEvent_t e = next_event();
          cudaEventRecord(e, xs);
          cudaStreamWaitEvent(main, e, 0);
          export_ev = false;
        }
      }
    }
    if (failed) return fail(-9, err);
    return check(train ? "train_fwd_bwd" : "eval_forward");
  }
  // slice k of the gradient buffer is final once everything enqueued so far on `streams` / recorded in `evs` is done
  void signal_slice(int k, std::initializer_list<cudaStream_t> streams, const std::vector<cudaEvent_t>& evs = {}) {
    for (cudaStream_t st : streams) {
      cudaEvent_t e = next_event();
      cudaEventRecord(e, st);
      cudaStreamWaitEvent(xs, e, 0);
    }
    for (cudaEvent_t e : evs) cudaStreamWaitEvent(xs, e, 0);
    if (export_capturing)
      cudaEventRecordWithFlags(ev_slice[k], xs, cudaEventRecordExternal);
    else
      cudaEventRecord(ev_slice[k], xs);
  }
  void join_wgrad(cudaStream_t main) {  // the weight-gradient streams rejoin the caller's stream
    if (profiling) return;
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e = next_event();
      cudaEventRecord(e, wside[i]);
      cudaStreamWaitEvent(main, e, 0);
    }
  }
  void encoders_bwd(const float* const* xin, int B, Branch& b0, Branch& b1, cudaStream_t main, int phase = 0) {
    const bool two = n_enc == 2;
    if (two) fork(main);
    encoder_bwd(enc[0], xin[0], B, b0, phase);
    if (two) encoder_bwd(enc[1], xin[1], B, b1, phase);
    if (export_ev) signal_slice(1, {}, deep_ev);  // layer3 + layer4 + Linear of every encoder (recorded in encoder_bwd)
    if (two) join(main);
    join_wgrad(main);
  }

  // encoders + latent head only, eval-mode BatchNorm (get_embeddings_multimodal, scripts/...:22-34)
  int run_embed(const float* x1, const float* x2, const int64_t* src, const int64_t* cls, int B, int zscore_ddof,
                float* out_enc, float* out_mu, float* out_logvar, cudaStream_t main) {
    launches = 0;
    Branch b0{main, ws + part_off[0], ws + bpart_off[0]}, b1{side, ws + part_off[1], ws + bpart_off[1]};
    if (B <= helper_max_batch) b0.hst = hside[0], b1.hst = hside[1];
    launch_bn_eval_coefs(reinterpret_cast<const BnEvalEntry*>(ws + bn_table_off), (int)bn_table.size(), P, bn_mean, bn_var,
                         ws, main);
    ++launches;
    refresh_weights(false, main);
    const bool two = n_enc == 2;
    if (two) fork(main);
    encoder_fwd(enc[0], x1, B, false, b0);
    if (two) {
      encoder_fwd(enc[1], x2, B, false, b1);
      join(main);
    }
    HeadArgs ha = head_args(B, src, cls, nullptr, false, false, out_enc, out_mu, out_logvar, 0.f, zscore_ddof);
    ha.kl_sum = nullptr;
    launch_head_fwd(ha, main);
    ++launches;
    if (failed) return fail(-9, err);
    return check("hippie_embed");
  }

  // ---- module-level API (eager launches; not on the training hot path) -----------------------------------------
  void prepare_forward(bool train, cudaStream_t main) {
    launches = 0;
    if (train) cudaMemsetAsync(ws + tot_off, 0, tot_floats * sizeof(float), main);
    refresh_weights(false, main);
    if (!train) {
      launch_bn_eval_coefs(reinterpret_cast<const BnEvalEntry*>(ws + bn_table_off), (int)bn_table.size(), P, bn_mean,
                           bn_var, ws, main);
      ++launches;
    }
  }
  // ResNet18Enc.forward (hippie/backbones.py:94-103): x [B,1,Lin] -> [B,2z]
  int run_encoder(int which, const float* x, int B, bool train, float* out, cudaStream_t main) {
    prepare_forward(train, main);
    Branch b0{main, ws + part_off[0], ws + bpart_off[0]};
    encoder_fwd(enc[which], x, B, train, b0);
    cudaMemcpyAsync(out, ws + enc[which].h, (size_t)B * 2 * cfg.z_dim * sizeof(float), cudaMemcpyDeviceToDevice, main);
    if (failed) return fail(-9, err);
    return check("hippie_encoder_forward");
  }
  // ResNet18Dec.forward (hippie/backbones.py:128-141): d [B,2z] -> [B,1,Lo]
  int run_decoder(int which, const float* d, int B, bool train, float* out, cudaStream_t main) {
    prepare_forward(train, main);
    Branch b0{main, ws + part_off[0], ws + bpart_off[0]};
    Decoder& D = dec[which];
    cudaMemcpyAsync(ws + D.d, d, (size_t)B * 2 * cfg.z_dim * sizeof(float), cudaMemcpyDeviceToDevice, main);
    decoder_fwd(D, B, train, b0);
    decoder_tail(D, which, nullptr, out, B, false, 1.f, b0);
    if (failed) return fail(-9, err);
    return check("hippie_decoder_forward");
  }
  // MultiModalCVAE.encode (hippie/model.py:402-408) / hippieUnimodalCVAE.encode (:50-56): embeddings are inputs
  int run_encode(const float* x1, const float* x2, const float* emb_src, const float* emb_cls, int B, bool train,
                 float* out_enc, float* out_mu, float* out_logvar, cudaStream_t main) {
    prepare_forward(train, main);
    Branch b0{main, ws + part_off[0], ws + bpart_off[0]}, b1{side, ws + part_off[1], ws + bpart_off[1]};
    const bool two = n_enc == 2;
    if (two) fork(main);
    encoder_fwd(enc[0], x1, B, train, b0);
    if (two) {
      encoder_fwd(enc[1], x2, B, train, b1);
      join(main);
    }
    HeadArgs ha = head_args(B, nullptr, nullptr, nullptr, train, false, out_enc, out_mu, out_logvar, 0.f, -1);
    ha.kl_sum = nullptr, ha.emb_src_in = emb_src, ha.emb_cls_in = emb_cls;
    launch_head_fwd(ha, main);
    ++launches;
    if (failed) return fail(-9, err);
    return check("hippie_encode");
  }
  // MultiModalCVAE.decode (hippie/model.py:410-422) / hippieUnimodalCVAE.decode (:58-61): z is an input
  int run_decode(const float* z_in, const float* emb_src, const float* emb_cls, int B, bool train, float* out_dec1,
                 float* out_dec2, cudaStream_t main) {
    prepare_forward(train, main);
    Branch b0{main, ws + part_off[0], ws + bpart_off[0]}, b1{side, ws + part_off[1], ws + bpart_off[1]};
    HeadArgs ha = head_args(B, nullptr, nullptr, nullptr, train, true, nullptr, nullptr, nullptr, 0.f, -1);
    ha.kl_sum = nullptr, ha.emb_src_in = emb_src, ha.emb_cls_in = emb_cls, ha.z_in = z_in, ha.stage = kHeadDecodeOnly;
    launch_head_fwd(ha, main);
    ++launches;
    const bool two = n_dec == 2;
    float* dec_out[2] = {out_dec1, out_dec2};
    if (two) fork(main);
    for (int m = 0; m < n_dec; ++m) {
      Branch& br = m == 0 ? b0 : b1;
      decoder_fwd(dec[m], B, train, br);
      decoder_tail(dec[m], m, nullptr, dec_out[m], B, false, 1.f, br);
    }
    if (two) join(main);
    if (failed) return fail(-9, err);
    return check("hippie_decode");
  }

  int exec(const CallArgs& a, cudaStream_t main) {
    if (a.mode == 3) return run_embed(a.x1, a.x2, a.src, a.cls, a.B, a.zscore, a.enc, a.mu, a.lv, main);
    const bool backward = a.mode == 0 || a.mode >= 4;
    return run(a.mode != 2, backward, a.x1, a.x2, a.src, a.cls, a.eps, a.B, a.beta, a.w1, a.w2, a.scalars, a.enc, a.mu, a.lv,
               a.d1, a.d2, main, a.mode >= 4 ? a.mode - 4 : -1);
  }

  // Graph-aware dispatch.  Inputs are copied into fixed staging buffers (one launch), the captured launch sequence is
  // replayed, requested outputs are copied out (one launch).  The first call of a signature runs eagerly so that every
  // lazy initialisation (tensor maps, kernel attributes) happens outside a capture.
  int call(const CallArgs& a, cudaStream_t main) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (use_graphs && !profiling) cudaStreamIsCapturing(main, &cs);
    if (!use_graphs || profiling || cs != cudaStreamCaptureStatusNone) return exec(a, main);
    GraphKey key;
    std::memset(&key, 0, sizeof(key));
    key.mode = a.mode, key.B = a.B, key.zscore = a.zscore, key.beta = a.beta, key.w1 = a.w1, key.w2 = a.w2;
    key.flags = (a.cls ? 1 : 0) | (a.eps ? 2 : 0) | (a.scalars ? 4 : 0) | (a.enc ? 8 : 0) | (a.mu ? 16 : 0) |
                (a.lv ? 32 : 0) | (a.d1 ? 64 : 0) | (a.d2 ? 128 : 0) | (planes_fresh && planes_keep ? 256 : 0);
    if (graphs.size() >= kMaxGraphs && !graphs.count(key)) return exec(a, main);  // e.g. a beta schedule: stay eager
    GraphEntry& e = graphs[key];
    if (e.seen <= 0) {
      if (e.seen == 0) e.seen = 1;
      return exec(a, main);
    }
    const int z = cfg.z_dim;
    IoCopy in{};
    auto add = [](IoCopy& c, const void* src, void* dst, int64_t words) {
      if (src && dst && words > 0) c.seg[c.n++] = IoSeg{src, dst, words};
    };
    add(in, a.x1, ws + st_x1, (int64_t)a.B * cfg.len_wave);
    if (cfg.multimodal) add(in, a.x2, ws + st_x2, (int64_t)a.B * cfg.len_isi);
    add(in, a.src, ws + st_src, (int64_t)a.B * 2);
    add(in, a.cls, ws + st_cls, (int64_t)a.B * 2);
    add(in, a.eps, ws + st_eps, (int64_t)a.B * z);
    launch_io_copy(in, main);
    if (!e.exec) {
      CallArgs s = a;
      s.x1 = ws + st_x1, s.x2 = cfg.multimodal ? ws + st_x2 : nullptr;
      s.src = reinterpret_cast<const int64_t*>(ws + st_src);
      s.cls = a.cls ? reinterpret_cast<const int64_t*>(ws + st_cls) : nullptr;
      s.eps = a.eps ? ws + st_eps : nullptr;
      s.scalars = a.scalars ? ws + st_scal : nullptr;
      s.enc = a.enc ? ws + st_enc : nullptr, s.mu = a.mu ? ws + st_mu : nullptr, s.lv = a.lv ? ws + st_lv : nullptr;
      s.d1 = a.d1 ? ws + st_d1 : nullptr, s.d2 = a.d2 ? ws + st_d2 : nullptr;
      cudaGraph_t g = nullptr;
      cudaError_t ce = cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed);
      int rc = ce == cudaSuccess ? exec(s, cap) : (int)ce;
      if (ce == cudaSuccess) ce = cudaStreamEndCapture(cap, &g);
      if (rc == 0 && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&e.exec, g, graph_flags);
      if (g) cudaGraphDestroy(g);
      if (rc != 0 || ce != cudaSuccess || !e.exec) {  // stay eager for this signature
        cudaGetLastError();
        e.exec = nullptr, e.seen = -1, failed = false;
        return exec(a, main);
      }
      e.launches = launches;
    }
    cudaGraphLaunch(e.exec, main);
    // a graph captured from stale planes contains the conversion (later parts of a split step contain none)
    if (use_tc && (a.mode <= 4 || a.mode == 8)) planes_fresh = true;
    IoCopy out{};
    add(out, ws + st_scal, a.scalars, a.mode == 3 ? 0 : 4);
    add(out, ws + st_enc, a.enc, (int64_t)a.B * z), add(out, ws + st_mu, a.mu, (int64_t)a.B * z);
    add(out, ws + st_lv, a.lv, (int64_t)a.B * z);
    if (a.mode != 3) {
      add(out, ws + st_d1, a.d1, (int64_t)a.B * cfg.len_wave);
      if (cfg.multimodal) add(out, ws + st_d2, a.d2, (int64_t)a.B * cfg.len_isi);
    }
    if (out.n) launch_io_copy(out, main);
    launches = e.launches + 1 + (out.n ? 1 : 0);
    return check("graph launch");
  }
};

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int hippie_abi_version(void) { return 3; }

int hippie_create(const hippie_cfg* cfg, hippie_handle* out) {
  if (!cfg || !out) return -1;
  *out = nullptr;
  if (cfg->z_dim < 1 || cfg->z_dim > 256 || cfg->class_hidden_dim < 1 || cfg->num_sources < 1 || cfg->num_classes < 1 ||
      cfg->len_wave < 4 || (cfg->multimodal == 1 && cfg->len_isi < 4) || cfg->max_batch < 1 || cfg->multimodal < 0 ||
      cfg->multimodal > HIPPIE_KIND_DECODER)
    return -1;
  hippie_engine* e = new hippie_engine();
  e->cfg = *cfg;
  e->build();
  *out = e;
  return 0;
}

void hippie_destroy(hippie_handle h) {
  if (!h) return;
  h->clear_graphs();
  if (h->side) cudaStreamDestroy(h->side);
  if (h->cap) cudaStreamDestroy(h->cap);
  if (h->xs) cudaStreamDestroy(h->xs);
  for (int i = 0; i < 2; ++i)
    if (h->ev_slice[i]) cudaEventDestroy(h->ev_slice[i]);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->wside[i]) cudaStreamDestroy(h->wside[i]);
    if (h->hside[i]) cudaStreamDestroy(h->hside[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

const char* hippie_last_error(hippie_handle h) { return h ? h->err.c_str() : "null handle"; }

int hippie_num_params(hippie_handle h) { return h ? (int)h->params.size() : -1; }
int64_t hippie_param_floats(hippie_handle h) { return h ? h->param_floats : -1; }
int hippie_param_info(hippie_handle h, int idx, char* name, int name_cap, int64_t* offset, int64_t* numel, int32_t* ndim,
                      int64_t* shape3, int32_t* layout) {
  if (!h || idx < 0 || idx >= (int)h->params.size()) return -1;
  const Param& p = h->params[idx];
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", p.name.c_str());
  if (offset) *offset = p.off;
  if (numel) *numel = p.numel;
  if (ndim) *ndim = p.ndim;
  if (shape3) shape3[0] = p.shape[0], shape3[1] = p.shape[1], shape3[2] = p.shape[2];
  if (layout) *layout = p.layout;
  return 0;
}
int hippie_num_bn(hippie_handle h) { return h ? (int)h->bns.size() : -1; }
int64_t hippie_bn_floats(hippie_handle h) { return h ? h->bn_floats : -1; }
int hippie_bn_info(hippie_handle h, int idx, char* name, int name_cap, int64_t* offset, int64_t* channels) {
  if (!h || idx < 0 || idx >= (int)h->bns.size()) return -1;
  const BNInfo& b = h->bns[idx];
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", b.name.c_str());
  if (offset) *offset = b.run_off;
  if (channels) *channels = b.C;
  return 0;
}
size_t hippie_workspace_bytes(hippie_handle h) { return h ? (size_t)h->ws_floats * sizeof(float) : 0; }
int hippie_num_tensors(hippie_handle h) { return h ? (int)h->acts.size() : -1; }  // offset -1: no fp32 copy exists
int hippie_tensor_info(hippie_handle h, int idx, char* name, int name_cap, int64_t* offset, int32_t* L, int32_t* C,
                       int32_t* pad) {
  if (!h || idx < 0 || idx >= (int)h->acts.size()) return -1;
  const Act& a = h->acts[idx];
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", a.name.c_str());
  if (offset) *offset = a.off;
  if (L) *L = a.L;
  if (C) *C = a.C;
  if (pad) *pad = 1;
  return 0;
}

int hippie_bind(hippie_handle h, float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* bn_mean,
                float* bn_var, int64_t* bn_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return -1;
  if (!params || !bn_mean || !bn_var || !bn_count || !workspace) return h->fail(-4, "null buffer");
  if (!h->cfg.inference_only && h->full_model() && (!grads || !exp_avg || !exp_avg_sq))
    return h->fail(-4, "training engine needs grads and AdamW state");
  if (workspace_bytes < (size_t)h->ws_floats * sizeof(float)) return h->fail(-5, "workspace too small");
  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess) return h->fail((int)ce, std::string("cudaGetDevice: ") + cudaGetErrorString(ce));
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return h->fail(-6, "libhippie_b200 is built for sm_100a (B200) only");
  cudaStream_t st = (cudaStream_t)stream;
  h->P = params, h->G = grads, h->M1 = exp_avg, h->M2 = exp_avg_sq;
  h->bn_mean = bn_mean, h->bn_var = bn_var, h->bn_count = bn_count, h->ws = (float*)workspace;
  h->clear_graphs();
  h->planes_fresh = false;
  if (const char* g = getenv("HIPPIE_B200_KEEP_PLANES")) h->planes_keep = atoi(g) != 0;
  if (const char* g = getenv("HIPPIE_B200_SHORTCUT_STREAM")) h->helper_max_batch = atoi(g);
  if (const char* g = getenv("HIPPIE_B200_GRAPHS")) h->use_graphs = atoi(g) != 0;
  if (const char* g = getenv("HIPPIE_B200_GRAPH_PRIO")) h->graph_flags = atoi(g) ? cudaGraphInstantiateFlagUseNodePriority : 0;
  if (const char* g = getenv("HIPPIE_B200_TMA_STORE")) h->tma_epilogue = atoi(g) != 0;
  if (!h->side) {
    // Stream priorities, measured (gpurun_out/r02_exp39.txt, r02_exp40.txt): giving the branch chains the HIGHEST priority and
    // the weight-gradient streams the lowest -- the obvious choice for a critical path -- starves the weight gradients until
    // they run as an exposed tail behind the chains: eager bs512 step 3.79 ms, graph replay with per-node priorities 3.21 ms.
    // The other way round (weight gradients take free SM slots first, the chains' few CTAs fit in between) gives 3.16 ms eager
    // and 3.01 ms replayed; replay without node priorities (everything at the caller's priority) 3.03 ms.
    int prio_chain = 0, prio_wgrad = 0;
    cudaDeviceGetStreamPriorityRange(&prio_chain, &prio_wgrad);  // (least, greatest): chains least, weight gradients greatest
    if (const char* g = getenv("HIPPIE_B200_PRIO_INVERT"))  // 0: round-1 assignment (chains above the weight gradients)
      if (!atoi(g)) std::swap(prio_chain, prio_wgrad);
    int prio_side = prio_chain;
    if (const char* g = getenv("HIPPIE_B200_PRIO"))  // experiment: "<chain 0>,<chain 1>,<weight gradients>"
      sscanf(g, "%d,%d,%d", &prio_chain, &prio_side, &prio_wgrad);
    cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_side);
    cudaStreamCreateWithPriority(&h->cap, cudaStreamNonBlocking, prio_chain);
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithPriority(&h->wside[i], cudaStreamNonBlocking, prio_wgrad);
    cudaStreamCreateWithPriority(&h->hside[0], cudaStreamNonBlocking, prio_chain);
    cudaStreamCreateWithPriority(&h->hside[1], cudaStreamNonBlocking, prio_side);
    cudaStreamCreateWithPriority(&h->xs, cudaStreamNonBlocking, prio_chain);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&h->ev_slice[i], cudaEventDisableTiming);
    h->ev_pool.resize(512);
    for (auto& e : h->ev_pool) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  }
  cudaMemsetAsync(workspace, 0, (size_t)h->ws_floats * sizeof(float), st);
  h->cmaps.assign(h->n_convs, ConvMaps{});
  h->use_tc = false, h->failed = false;
  if (h->pair_mode()) {
    h->use_tc = pair_init(&h->tc_note);
    if (!h->use_tc) return h->fail(-8, "tcgen05 path unavailable: " + h->tc_note);
  }
  if (!h->wt_table.empty())
    cudaMemcpyAsync(h->ws + h->wt_table_off, h->wt_table.data(), h->wt_table.size() * sizeof(WtEntry),
                    cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(h->ws + h->bn_table_off, h->bn_table.data(), h->bn_table.size() * sizeof(BnEvalEntry),
                  cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  h->bound = true;
  return h->check("hippie_bind");
}

int hippie_train_fwd_bwd(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                         const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out, float* out_enc,
                         float* out_mu, float* out_logvar, float* out_dec1, float* out_dec2, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate(B, x1, x2, src)) return rc;
  if (h->cfg.inference_only) return h->fail(-7, "inference-only engine");
  if (!eps) return h->fail(-4, "eps is required for training (reparameterisation noise)");
  if (B < 2) return h->fail(-3, "training-mode BatchNorm needs B >= 2");
  hippie_engine::CallArgs a{0, x1, x2, src, cls, eps, B, beta, w1, w2, -1, scalars_out, out_enc, out_mu, out_logvar, out_dec1, out_dec2};
  return h->call(a, (cudaStream_t)stream);
}

int hippie_train_fwd_bwd_part(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                              const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out,
                              int32_t part, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate(B, x1, x2, src)) return rc;
  if (h->cfg.inference_only) return h->fail(-7, "inference-only engine");
  if (part < 0 || part > 4) return h->fail(-3, "part must be 0..4");
  if (!eps) return h->fail(-4, "eps is required for training (reparameterisation noise)");
  if (B < 2) return h->fail(-3, "training-mode BatchNorm needs B >= 2");
  hippie_engine::CallArgs a{4 + part, x1, x2, src, cls, eps, B, beta, w1, w2, -1, (part == 0 || part == 4) ? scalars_out : nullptr,
                            nullptr, nullptr, nullptr, nullptr, nullptr};
  return h->call(a, (cudaStream_t)stream);
}

int hippie_slice_wait(hippie_handle h, int32_t slice, void* stream) {
  if (!h) return -1;
  if (!h->bound) return h->fail(-2, "hippie_bind has not been called");
  if (slice < 0 || slice > 1) return h->fail(-3, "slice must be 0 (latent head + decoders) or 1 (deep encoder halves)");
  cudaError_t ce = cudaStreamWaitEvent((cudaStream_t)stream, h->ev_slice[slice], 0);
  return ce == cudaSuccess ? 0 : h->fail((int)ce, std::string("hippie_slice_wait: ") + cudaGetErrorString(ce));
}

int64_t hippie_grad_split(hippie_handle h) { return h ? h->grad_split : -1; }

int hippie_grad_bounds(hippie_handle h, int64_t* bounds) {
  if (!h || !bounds) return -1;
  for (int e = 0; e < 2; ++e) bounds[3 * e] = h->enc_begin[e], bounds[3 * e + 1] = h->enc_deep[e], bounds[3 * e + 2] = h->enc_end[e];
  return 0;
}

int hippie_eval_forward(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                        const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out, float* out_enc,
                        float* out_mu, float* out_logvar, float* out_dec1, float* out_dec2, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate(B, x1, x2, src)) return rc;
  hippie_engine::CallArgs a{2, x1, x2, src, cls, eps, B, beta, w1, w2, -1, scalars_out, out_enc, out_mu, out_logvar, out_dec1, out_dec2};
  return h->call(a, (cudaStream_t)stream);
}

int hippie_train_forward(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                         const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out, float* out_enc,
                         float* out_mu, float* out_logvar, float* out_dec1, float* out_dec2, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate(B, x1, x2, src)) return rc;
  if (B < 2) return h->fail(-3, "training-mode BatchNorm needs B >= 2");
  hippie_engine::CallArgs a{1, x1, x2, src, cls, eps, B, beta, w1, w2, -1, scalars_out, out_enc, out_mu, out_logvar, out_dec1, out_dec2};
  return h->call(a, (cudaStream_t)stream);
}

int hippie_embed(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls, int32_t B,
                 int32_t zscore_ddof, float* out_enc, float* out_mu, float* out_logvar, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate(B, x1, x2, src)) return rc;
  hippie_engine::CallArgs a{3, x1, x2, src, cls, nullptr, B, 0.f, 0.f, 0.f, zscore_ddof, nullptr, out_enc, out_mu, out_logvar, nullptr, nullptr};
  return h->call(a, (cudaStream_t)stream);
}

int hippie_encoder_forward(hippie_handle h, int32_t which, const float* x, int32_t B, int32_t train, float* out, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate_module_call(B, train != 0)) return rc;
  if (which < 0 || which >= h->n_enc) return h->fail(-3, "no such encoder");
  if (!x || !out) return h->fail(-4, "null pointer");
  return h->run_encoder(which, x, B, train != 0, out, (cudaStream_t)stream);
}

int hippie_decoder_forward(hippie_handle h, int32_t which, const float* d, int32_t B, int32_t train, float* out, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate_module_call(B, train != 0)) return rc;
  if (which < 0 || which >= h->n_dec) return h->fail(-3, "no such decoder");
  if (!d || !out) return h->fail(-4, "null pointer");
  return h->run_decoder(which, d, B, train != 0, out, (cudaStream_t)stream);
}

int hippie_encode(hippie_handle h, const float* x1, const float* x2, const float* source_emb, const float* class_emb,
                  int32_t B, int32_t train, float* out_enc, float* out_mu, float* out_logvar, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate_module_call(B, train != 0)) return rc;
  if (!h->full_model()) return h->fail(-7, "stand-alone backbone engine");
  if (!x1 || (h->cfg.multimodal && !x2) || !source_emb || !class_emb) return h->fail(-4, "null input pointer");
  return h->run_encode(x1, x2, source_emb, class_emb, B, train != 0, out_enc, out_mu, out_logvar, (cudaStream_t)stream);
}

int hippie_decode(hippie_handle h, const float* z, const float* source_emb, const float* class_emb, int32_t B, int32_t train,
                  float* out_dec1, float* out_dec2, void* stream) {
  if (!h) return -1;
  if (int rc = h->validate_module_call(B, train != 0)) return rc;
  if (!h->full_model()) return h->fail(-7, "stand-alone backbone engine");
  if (!z || !source_emb || !class_emb || !out_dec1 || (h->cfg.multimodal && !out_dec2)) return h->fail(-4, "null pointer");
  return h->run_decode(z, source_emb, class_emb, B, train != 0, out_dec1, out_dec2, (cudaStream_t)stream);
}

int hippie_device_flags(hippie_handle h, uint32_t* flags_out, int32_t clear, void* stream) {
  if (!h || !flags_out) return -1;
  if (!h->bound) return h->fail(-2, "hippie_bind has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce = cudaMemcpyAsync(flags_out, h->flags(), sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess && clear) ce = cudaMemsetAsync(h->flags(), 0, sizeof(uint32_t), st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);  // a diagnostic call: this one does wait for the stream
  return ce == cudaSuccess ? 0 : h->fail((int)ce, std::string("hippie_device_flags: ") + cudaGetErrorString(ce));
}

int hippie_clip_adamw(hippie_handle h, double lr, double beta1, double beta2, double eps, double weight_decay, float max_norm,
                      float grad_scale, int32_t step, int32_t step_cls, int32_t has_cls_grad, float* scalars_out,
                      void* stream) {
  if (!h) return -1;
  if (!h->bound) return h->fail(-2, "hippie_bind has not been called");
  if (h->cfg.inference_only || !h->full_model()) return h->fail(-7, "inference-only engine");
  if (!scalars_out) return h->fail(-4, "scalars_out is required");
  if (step < 1) return h->fail(-3, "step is 1-based");
  AdamArgs a{};
  a.p = h->P, a.g = h->G, a.m = h->M1, a.v = h->M2, a.n = h->param_floats;
  a.skip_lo = h->params[h->cls_emb].off, a.skip_hi = a.skip_lo + h->params[h->cls_emb].numel;
  a.lr = lr, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.wd = weight_decay, a.max_norm = max_norm;
  a.grad_scale = grad_scale, a.step = step, a.step_cls = step_cls, a.has_cls_grad = has_cls_grad;
  a.partials = h->ws + h->adam_part, a.scalars = scalars_out;
  if (h->use_tc && h->planes_fresh && h->planes_keep) {  // keep the weight pair planes current (stale planes are converted in full by the next forward)
    a.wp_hi = h->WP(), a.wp_lo = h->WP() + h->param_floats, a.wp_scale = kWeightPairScale, a.flags = h->flags();
  }
  launch_clip_adamw(a, (cudaStream_t)stream);
  h->launches = kClipAdamLaunches;
  return h->check("hippie_clip_adamw");
}

int hippie_params_changed(hippie_handle h) {
  if (!h) return -1;
  h->planes_fresh = false;
  return 0;
}

int hippie_preprocess_batch(const double* wave_raw, int32_t wave_width, const double* isi_raw, int32_t isi_width,
                            const int64_t* index, int32_t B, float* x1, int32_t len_wave, float* x2, int32_t len_isi,
                            void* stream) {
  if (B == 0) return 0;
  if (B < 0 || (wave_raw && (!x1 || wave_width < 1 || len_wave < 1)) || (isi_raw && (!x2 || isi_width < 1 || len_isi < 1)))
    return -1;
  if (wave_raw) launch_preprocess(wave_raw, wave_width, index, B, len_wave, 0, x1, (cudaStream_t)stream);
  if (isi_raw) launch_preprocess(isi_raw, isi_width, index, B, len_isi, 1, x2, (cudaStream_t)stream);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int hippie_last_launch_count(hippie_handle h) { return h ? h->launches : -1; }

int hippie_conv_path_in_use(hippie_handle h) { return h ? (h->use_tc ? 2 : 1) : -1; }

int hippie_profile(hippie_handle h, int enable) {
  if (!h) return -1;
  for (auto& r : h->prof) cudaEventDestroy(r.e0), cudaEventDestroy(r.e1);
  h->prof.clear();
  h->profiling = enable != 0;
  return 0;
}

int hippie_profile_read(hippie_handle h, int kind, double* total_ms, double* total_flop, int* launches) {
  if (!h || !total_ms || !total_flop || !launches) return -1;
  *total_ms = 0, *total_flop = 0, *launches = 0;
  for (auto& r : h->prof) {
    if (r.kind != kind) continue;
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    *total_ms += ms, *total_flop += r.flop, *launches += 1;
  }
  return h->check("hippie_profile_read");
}

}  // extern "C"
