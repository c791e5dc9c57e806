/*
 * hippie_b200.h -- C ABI of libhippie_b200.so, the sm_100a engine behind HIPPIE's cVAE hot path.
 *
 * The reference (aghatpande/HIPPIE) has no native code and no FFI: its hot path is the Python
 * call chain  MultiModalCVAETrainModule.training_step -> MultiModalCVAE.forward -> ResNet18Enc /
 * ResNet18Dec -> torch ATen, followed by Lightning's clip_grad_norm_ and torch.optim.AdamW.step.
 * Each entry point below replaces one link of that chain; the reference interface it stands in
 * for is cited as /root/reference/<file>:<line>.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes and PODs only; no torch types.  All tensor pointers are DEVICE
 *     pointers owned by the caller (PyTorch); the engine never allocates, frees or resizes them.
 *   - every launch goes to the caller-supplied `stream` (a cudaStream_t passed as void*); internal
 *     side streams are forked from / joined to it with events.  No hidden host syncs, no host reads
 *     of device data.  From the second call of a signature (entry point, B, which optional pointers
 *     are given, beta / w1 / w2) on, the launch sequence is replayed as a CUDA graph from staged
 *     inputs (HIPPIE_B200_GRAPHS=0 disables that); calls made while the caller is capturing its own
 *     graph are recorded into the caller's graph instead.
 *   - return value: 0 = ok, <0 = argument/config error, >0 = cudaError_t.  hippie_last_error()
 *     returns a message for the most recent non-zero return of that handle.
 *   - one handle per (process, device); not thread-safe.
 *   - activations live in the caller's workspace as channels-last rows [B][L+2][C] with one zero
 *     row on either side of every sample (DESIGN.md "Data layout in HBM").
 */
#ifndef HIPPIE_B200_H
#define HIPPIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef struct hippie_engine* hippie_handle;

/* Constructor arguments of MultiModalCVAE (hippie/model.py:352) / hippieUnimodalCVAE (:13). */
typedef struct hippie_cfg {
  int32_t z_dim;
  int32_t class_hidden_dim;
  int32_t num_sources;
  int32_t num_classes;
  int32_t len_wave;   /* output_size_wave (50); the single output_size when unimodal          */
  int32_t len_isi;    /* output_size_isi (100); ignored when unimodal                          */
  int32_t multimodal; /* 1 = MultiModalCVAE, 0 = hippieUnimodalCVAE; HIPPIE_KIND_ENCODER / _DECODER = a
                         ResNet18Enc / ResNet18Dec on its own (hippie/backbones.py:73-141; forward-only engines:
                         len_wave = input length / output_size, parameter names without a module prefix)       */
  int32_t max_batch;  /* largest B any call will pass; sizes the workspace                      */
  int32_t inference_only; /* 1 = no gradient tensors in the workspace (embedding engines)      */
  int32_t conv_path;  /* 0 = tcgen05 implicit GEMMs over fp16 pair planes (bind fails when TMA tensor maps are
                         unavailable), 1 = FP32 CUDA-core implicit GEMMs (the parity yardstick of the former)     */
} hippie_cfg;

enum { HIPPIE_KIND_UNIMODAL = 0, HIPPIE_KIND_MULTIMODAL = 1, HIPPIE_KIND_ENCODER = 2, HIPPIE_KIND_DECODER = 3 };

/* Device-side error flags (hippie_device_flags): conditions the reference turns into exceptions or that silently lose
 * precision here.  Sticky until cleared. */
enum {
  HIPPIE_FLAG_SOURCE_LABEL = 1,    /* a source label outside [0, num_sources): nn.Embedding raises IndexError; the kernels
                                      read row 0 and skip the gradient of that sample                                    */
  HIPPIE_FLAG_CLASS_LABEL = 2,     /* likewise for class labels and num_classes                                          */
  HIPPIE_FLAG_PAIR_SATURATED = 4,  /* an activation fed to a GEMM exceeded the fp16 range of the pair planes (65504)     */
  HIPPIE_FLAG_WEIGHT_SATURATED = 8 /* a parameter exceeded the range of the scaled weight planes (|w| >= 255.9)          */
};

/* Layout kinds of a parameter inside the flat buffer. */
enum {
  HIPPIE_LAYOUT_NATIVE = 0,  /* same element order as the torch tensor                          */
  HIPPIE_LAYOUT_CONV_OKI = 1 /* Conv1d weight stored [Cout][k][Cin]; torch shape is [Cout][Cin][k] */
};

int hippie_abi_version(void);

int hippie_create(const hippie_cfg* cfg, hippie_handle* out);
void hippie_destroy(hippie_handle h);
const char* hippie_last_error(hippie_handle h);

/* ---- layout enumeration, in state_dict() order (hippie/model.py:360-395 construction order) ----
 * Parameters: `offset`/`numel` in floats inside the flat params / grads / exp_avg / exp_avg_sq
 * buffers (each hippie_param_floats() long; offsets are 32-byte aligned, gaps stay zero).
 * `shape` is the torch shape (ndim <= 3).  */
int hippie_num_params(hippie_handle h);
int64_t hippie_param_floats(hippie_handle h);
int hippie_param_info(hippie_handle h, int idx, char* name, int name_cap, int64_t* offset, int64_t* numel,
                      int32_t* ndim, int64_t* shape3, int32_t* layout);
/* BatchNorm1d buffers: running_mean at bn_mean[offset..offset+C), running_var likewise in bn_var,
 * num_batches_tracked at bn_count[idx] (int64). */
int hippie_num_bn(hippie_handle h);
int64_t hippie_bn_floats(hippie_handle h);
int hippie_bn_info(hippie_handle h, int idx, char* name, int name_cap, int64_t* offset, int64_t* channels);

/* Workspace the caller must provide (bytes, for cfg.max_batch). */
size_t hippie_workspace_bytes(hippie_handle h);

/* Named activation tensors inside the workspace (debug / parity tests): offset in floats,
 * positions L, channels C, pad rows per side (row of (b,l) = b*(L+2*pad)+pad+l). */
int hippie_num_tensors(hippie_handle h);
int hippie_tensor_info(hippie_handle h, int idx, char* name, int name_cap, int64_t* offset, int32_t* L,
                       int32_t* C, int32_t* pad);

/* Replaces nn.Module parameter/buffer ownership + torch.optim.AdamW state (hippie/model.py:447):
 * hands the engine the flat device buffers.  Zero-fills the workspace (async on `stream`). */
int hippie_bind(hippie_handle h, float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* bn_mean,
                float* bn_var, int64_t* bn_count, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces training_step + zero_grad + loss.backward() (hippie/model.py:454-482).
 *   x1 [B,1,len_wave] f32, x2 [B,1,len_isi] f32 (NULL when unimodal), src int64 [B],
 *   cls int64 [B] or NULL (class embedding := 0, hippie/model.py:426), eps f32 [B,z] = the N(0,1)
 *   draw of reparameterize (hippie/model.py:397-400).
 * Writes the gradient of the loss w.r.t. every parameter into `grads` (zero-filled first),
 * updates BatchNorm running statistics, and writes scalars_out[0..3] = total, mse1, mse2, kl_mean
 * (device floats).  out_* (device, may be NULL): enc [B,z], mu [B,z], logvar [B,z],
 * dec1 [B,len_wave], dec2 [B,len_isi]. */
int hippie_train_fwd_bwd(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                         const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out,
                         float* out_enc, float* out_mu, float* out_logvar, float* out_dec1, float* out_dec2,
                         void* stream);

/* The same step in two stream-ordered parts, for data-parallel callers that overlap the gradient exchange with the
 * backward pass (reference: Lightning's implicit DDP does the same with its gradient buckets):
 *   part 0 = forward + loss + decoder backward + latent-head backward; when it completes, grads[hippie_grad_split(h) ..
 *            hippie_param_floats(h)) are final (latent head + both decoders, 52 % of the parameters);
 *   part 1 = encoder backward; afterwards grads[0 .. hippie_grad_split(h)) are final.
 * Calling part 0 then part 1 with the same arguments is equivalent to hippie_train_fwd_bwd (parts 2, 3: below). */
int hippie_train_fwd_bwd_part(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                              const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out,
                              int32_t part, void* stream);
/* part 4 = the WHOLE step as one launch sequence (one CUDA graph) that additionally publishes, without interrupting the
 * backward pass, the two points at which a slice of the gradient buffer becomes final:
 *   slice 0 = grads[hippie_grad_split(h) ..)            (latent head + decoders), after the latent-head backward;
 *   slice 1 = the deep halves [deep_e, end_e) of hippie_grad_bounds(h)   (layer3, layer4, Linear of every encoder).
 * hippie_slice_wait(h, k, s) makes stream s wait for slice k of the most recent part-4 call (cudaStreamWaitEvent on an
 * event the step records -- as an external event-record node when the step is replayed from its graph), so that an
 * all-reduce issued on s overlaps the rest of the backward pass; the shallow remainder is final when the call's own
 * stream reaches the end of the step.  hippie_b200/parallel.py:train_step_overlapped. */
int hippie_slice_wait(hippie_handle h, int32_t slice, void* stream);
int64_t hippie_grad_split(hippie_handle h);

/* A finer split of the encoder backward for the same purpose: instead of part 1 call
 *   part 2 = the deep half of every encoder (Linear, layer4, layer3: 94 % of an encoder's parameters), then
 *   part 3 = the shallow half (layer2, layer1, stem).
 * hippie_grad_bounds fills bounds[6] = {begin, deep, end} per encoder (second triple empty for the unimodal model):
 * grads[deep .. end) are final after part 2, grads[begin .. deep) after part 3.  Parts 0, 2, 3 in this order are
 * equivalent to hippie_train_fwd_bwd. */
int hippie_grad_bounds(hippie_handle h, int64_t* bounds);

/* Replaces Lightning's gradient_clip_val (scripts/train_model_with_multimodal.py:55,701 ->
 * torch.nn.utils.clip_grad_norm_) followed by torch.optim.AdamW.step (hippie/model.py:447).
 *   grad_scale multiplies every gradient first (1/world after a summing all-reduce);
 *   max_norm <= 0 disables clipping; step / step_cls are the 1-based AdamW step counts of the
 *   ordinary parameters and of class_embedding.weight; has_cls_grad = 0 leaves class_embedding
 *   untouched (torch skips params whose grad is None).
 * scalars_out[4] = total gradient norm (after grad_scale), scalars_out[5] = clip coefficient. */
int hippie_clip_adamw(hippie_handle h, double lr, double beta1, double beta2, double eps, double weight_decay,
                      float max_norm, float grad_scale, int32_t step, int32_t step_cls, int32_t has_cls_grad,
                      float* scalars_out, void* stream);

/* The tensor-core path keeps a derived copy of the parameter buffer (fp16 pair planes of every weight) in the
 * workspace.  It is written in full by the first forward-type call after hippie_bind and kept current by
 * hippie_clip_adamw (each updated weight is re-split while it is in a register), so that the steady-state step and the
 * chunks of an embedding pass do not re-convert 64 MB of unchanged parameters per call.  The caller owns the
 * parameter buffer: after writing into it by any other means (the counterpart of load_state_dict, in-place
 * initialisation, a parameter broadcast -- hippie/model.py has no such call on the hot path) call
 * hippie_params_changed; the next forward-type call converts the whole buffer again.  hippie_b200/engine.py does this
 * from the version counter of the flat parameter tensor. */
int hippie_params_changed(hippie_handle h);

/* Replaces module.eval(); module(batch) (hippie/model.py:510-520): full forward with running
 * statistics.  Outputs as above (any may be NULL); scalars_out (may be NULL) gets the validation
 * loss terms of validation_step (hippie/model.py:484-508). */
int hippie_eval_forward(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                        const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out,
                        float* out_enc, float* out_mu, float* out_logvar, float* out_dec1, float* out_dec2,
                        void* stream);

/* Replaces module.train(); module(batch) without a backward pass (MultiModalCVAE.forward in training mode,
 * hippie/model.py:424-432): batch statistics, running-statistics update, no gradients.  Same arguments as
 * hippie_eval_forward. */
int hippie_train_forward(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                         const float* eps, int32_t B, float beta, float w1, float w2, float* scalars_out,
                         float* out_enc, float* out_mu, float* out_logvar, float* out_dec1, float* out_dec2,
                         void* stream);

/* Replaces get_embeddings_multimodal's model(sample)[0] (scripts/train_model_with_multimodal.py:
 * 22-34): encoders + fusion only (the reference also runs both decoders and discards them).
 * zscore_ddof: -1 = raw `encoded`; 0 / 1 = per-row z-score with that ddof fused in. */
int hippie_embed(hippie_handle h, const float* x1, const float* x2, const int64_t* src, const int64_t* cls,
                 int32_t B, int32_t zscore_ddof, float* out_enc, float* out_mu, float* out_logvar, void* stream);

/* Module-level forward calls (eager launches, no gradients; `train` != 0: batch statistics + running-statistics update).
 *   hippie_encoder_forward  replaces ResNet18Enc.forward (hippie/backbones.py:94-103): x [B,1,L] -> out [B,2z];
 *   hippie_decoder_forward  replaces ResNet18Dec.forward (hippie/backbones.py:128-141): d [B,2z] -> out [B,1,output_size].
 *     `which` selects encoder_mod1 / encoder_mod2 (decoder_mod1 / decoder_mod2) of a full model; 0 otherwise.
 *   hippie_encode  replaces MultiModalCVAE.encode (hippie/model.py:402-408) / hippieUnimodalCVAE.encode (:50-56): the
 *     embedding ROWS source_emb / class_emb [B, class_hidden_dim] are inputs, as in the reference;
 *   hippie_decode  replaces .decode (hippie/model.py:410-422 / :58-61): z [B,z] is an input. */
int hippie_encoder_forward(hippie_handle h, int32_t which, const float* x, int32_t B, int32_t train, float* out, void* stream);
int hippie_decoder_forward(hippie_handle h, int32_t which, const float* d, int32_t B, int32_t train, float* out, void* stream);
int hippie_encode(hippie_handle h, const float* x1, const float* x2, const float* source_emb, const float* class_emb,
                  int32_t B, int32_t train, float* out_enc, float* out_mu, float* out_logvar, void* stream);
int hippie_decode(hippie_handle h, const float* z, const float* source_emb, const float* class_emb, int32_t B, int32_t train,
                  float* out_dec1, float* out_dec2, void* stream);

/* Reads (and optionally clears) the HIPPIE_FLAG_* word the kernels set.  Unlike every other entry point this one waits
 * for `stream` (one 4-byte device-to-host copy); call it at epoch boundaries or after a suspicious result.  The
 * reference raises IndexError inside nn.Embedding for a bad label (hippie/model.py:425-426). */
int hippie_device_flags(hippie_handle h, uint32_t* flags_out, int32_t clear, void* stream);

/* Replaces EphysDataset / EphysDatasetLabeled.__getitem__ + DataLoader collation for one batch (hippie/dataloading.py:
 * 27-56, 74-101; normalize=False as every reference script passes): rows index[0..B) (NULL = rows 0..B-1) of raw,
 * device-resident float64 tables as pandas reads them -> float32 cast, log(isi + 1), linear interpolation
 * (align_corners=False) to len_wave / len_isi samples, written as x1 [B,1,len_wave] / x2 [B,1,len_isi].  Either table may
 * be NULL.  Stateless (no handle).  The waveform path is bit-exact with the reference's CPU result; the ISI path is
 * within one ulp of it before interpolation (ATen's vectorised logf is not correctly rounded). */
int hippie_preprocess_batch(const double* wave_raw, int32_t wave_width, const double* isi_raw, int32_t isi_width,
                            const int64_t* index, int32_t B, float* x1, int32_t len_wave, float* x2, int32_t len_isi,
                            void* stream);

/* Replaces KNeighborsClassifier(n_neighbors=k).fit(train).kneighbors(query) of the stage-3 evaluation
 * (scripts/train_model_with_multimodal.py:916-924, 929-931; scikit-learn, Euclidean metric, uniform weights): for every
 * row of query [n_query, dim] the k <= 32 nearest rows of train [n_train, dim] by squared Euclidean distance evaluated in
 * double (float32 inputs widened, summed in feature order), ascending, equal distances by ascending row index.
 * out_index [n_query, k] int64; out_sqdist [n_query, k] double (may be NULL; sklearn reports its square root).
 * Stateless (no handle). */
int hippie_knn_neighbors(const float* train, int64_t n_train, const float* query, int64_t n_query, int32_t dim,
                         int32_t k, int64_t* out_index, double* out_sqdist, void* stream);

/* Replaces .predict(test), balanced_accuracy_score and confusion_matrix for every k in [k_lo, k_hi] at once
 * (scripts/train_model_with_multimodal.py:919-934).  neighbors [n_query, k_stride] from hippie_knn_neighbors
 * (k_hi <= k_stride); train_class [n_train] / true_class [n_query] are dense class indices in [0, n_classes <= 128)
 * over the sorted union of the labels (what LabelEncoder / np.unique produce).  Majority vote, ties to the smallest
 * class.  out_pred [nk, n_query] int64 (may be NULL); out_confusion [nk, n_classes, n_classes] int64 rows = true class
 * (zeroed here); out_balanced_accuracy [nk] double = mean recall over the classes present in true_class, summed in
 * numpy's order.  true_class, out_confusion, out_balanced_accuracy may be NULL together (prediction only). */
int hippie_knn_evaluate(const int64_t* neighbors, int64_t n_query, int32_t k_stride, const int64_t* train_class,
                        const int64_t* true_class, int32_t n_classes, int32_t k_lo, int32_t k_hi, int64_t* out_pred,
                        int64_t* out_confusion, double* out_balanced_accuracy, void* stream);

/* Number of kernel launches issued by the most recent call of each kind (bench.py gpu_launches). */
int hippie_last_launch_count(hippie_handle h);

/* 2 = the tcgen05 pair-plane implicit GEMMs serve conv forward / dgrad / wgrad, 1 = the FP32 CUDA-core GEMMs do. */
int hippie_conv_path_in_use(hippie_handle h);

/* Measurement aid for bench.py (not part of the reference interface): when enabled, every implicit-GEMM launch
 * (kind 0 = conv forward, 1 = dgrad, 2 = wgrad) of the following calls is bracketed by CUDA events on the stream it
 * is launched on, and both branches run on the caller's stream so the per-launch times are not shared.
 * hippie_profile_read sums, per kind, the event durations and the algorithmic FLOPs (2*M*N*K over real rows). */
int hippie_profile(hippie_handle h, int enable);
int hippie_profile_read(hippie_handle h, int kind, double* total_ms, double* total_flop, int* launches);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HIPPIE_B200_H */
