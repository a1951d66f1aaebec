/*
 * snb.h - C ABI of libsnb.so, the B200 (sm_100a) implementation of the ray-rendering hot
 * path of wagnva/semantic-nerf-for-satellite-data.
 *
 * Boundary.  The reference is pure Python/PyTorch; the functions below are what a binding for
 * this path would call (the ctypes stub a reference maintainer would add is in INTEGRATION.md,
 * the in-tree one is semantic-nerf-for-satellite-data_b200/_lib.py).  Each entry names the
 * reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions.
 *   - All pointers are DEVICE pointers unless the name ends in _host.  Buffers are owned by the
 *     caller (PyTorch's allocator); nothing here allocates device memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - Every function returns 0 on success, a negative SNB_ERR_* for a rejected argument, or a
 *     positive cudaError_t.  snb_last_error() returns a thread-local description.
 *   - There is no CPU fallback: without a CUDA device every compute entry fails.
 *   - bf16 buffers are passed as void* (raw __nv_bfloat16).
 */
#ifndef SNB_H_
#define SNB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_VERSION 200

enum {
  SNB_OK = 0,
  SNB_ERR_INVALID = -1,     /* null pointer / negative size / misaligned buffer          */
  SNB_ERR_UNSUPPORTED = -2, /* shape outside what the kernels implement (e.g. S < 2)      */
  SNB_ERR_NO_DEVICE = -3,   /* no sm_100 device / driver entry point missing              */
  SNB_ERR_WORKSPACE = -4    /* workspace too small                                        */
};

/* model kinds: baseline/models/satnerf.py:101 (raw xyz input) and
 * semantic/models/rs_semantic.py:139 (positional mapping, semantic head) */
/* SNB_MODEL_NERF: vanilla NeRF as baseline/pipelines/nerf.py:26-34 builds it (baseline/models/nerf.py:98-212): positional
 * encoding of xyz (10) and of the view direction (4), ReLU activations, outputs [rgb | sigma].  Its plans hold the trunk, sigma,
 * feats and the rgb head only; the packed sun column is pinned to 1, so the compositing kernel's irradiance is 1; the 24 encoded view-direction values ride a 32-column `aux` row (snb_nerf_aux). */
enum { SNB_MODEL_SATNERF = 0, SNB_MODEL_SEMANTIC = 1, SNB_MODEL_NERF = 2,
       SNB_MODEL_SNERF = 3 /* S-NeRF (baseline/models/snerf.py:104-243): SatNeRF without the uncertainty head and embedding */ };

/* which heads a pass evaluates (bit mask).  SNB_HEADS_ALL is the reference's forward();
 * SNB_HEADS_SOLAR is what the solar-correction pass keeps (semantic/components/rendering.py:76-78);
 * SNB_HEADS_DEPTH is what the depth-supervision batch consumes (baseline/components/loss.py:40). */
enum {
  SNB_HEAD_SIGMA = 1, SNB_HEAD_RGB = 2, SNB_HEAD_SUN = 4, SNB_HEAD_BETA = 8, SNB_HEAD_SEM = 16, SNB_HEAD_SKY = 32,
  SNB_HEADS_ALL = 63, SNB_HEADS_SOLAR = 1 | 4, SNB_HEADS_DEPTH = 1
};

int snb_version(void);
const char* snb_last_error(void);
/* number of SMs of the current device; <0 on error */
int snb_device_sms(void);

/* ------------------------------------------------------------------------------------------------
 * K1  sample generation + encoding
 * replaces framework/components/rendering.py:84-116 (sample_rays), baseline/models/commons.py:58-74
 * (Mapping.forward), the nn.Embedding lookup + int cast of semantic/components/rendering.py:35-45,
 * the repeat_interleave broadcasts of semantic/models/rs_semantic.py:42-61 and sky_color(sun_d)
 * (rs_semantic.py:222-227, a per-ray function).
 *
 *   rays   (N,8) f32  [o(3) d(3) near far]          framework/components/rays.py:7-38
 *   extras (N,4) f32  [sun_d(3) ts]                 framework/components/rays.py:41-64
 *   u      (N,S) f32  jitter in [0,1) or NULL -> in-kernel Philox4x32-10 keyed on (seed, ray_offset+ray, sample)
 *   z_vals (N,S) f32  written (z_given=0) or read (z_given=1: the reference's given_z_vals)
 *   enc    bf16 (N*S, enc_ld): hi+lo is the two-term bf16 split of the fp32 encoding (k0 = 3 raw xyz /
 *          60 posenc).  satnerf: enc_ld = 64, [hi(3) | hi(3) | lo(3) | 0];  semantic: enc_ld = 128,
 *          [hi(60) | lo(60) | 0(8)].  NULL to skip.
 *   enc_sc same for the solar-correction points o + sun_d*z. NULL to skip.
 *   aux    bf16 (N*S,16): [1, sun_d(3), t(tau), 0..] - per-ray terms of the heads as a K-segment.
 *   sky    f32  (N,3): sigmoid(W2 relu(W1 sun_d + b1) + b2).  NULL to skip.
 * ---------------------------------------------------------------------------------------------- */
/*   t_steps (S) f32  linspace(0,1,S) exactly as the reference's host computes it (rendering.py:95);
 *          passed in so the stratification bins are bit-identical to the reference's. */
/*   seed_dev: optional DEVICE uint64; when non-NULL the Philox key is read from it instead of `seed`, so a captured CUDA
 *          graph of the training step draws fresh jitter on every replay. */
int snb_sample_encode(const float* rays, const float* extras, const float* u, uint64_t seed,
                      const uint64_t* seed_dev, uint64_t ray_offset, const float* t_steps, const float* t_table, int vocab, int tau,
                      const float* sky_w1, const float* sky_b1, const float* sky_w2, const float* sky_b2,
                      int sky_hidden, int n_rays, int n_samples, int model_kind, int z_given,
                      float* z_vals, void* enc, void* enc_sc, void* aux, float* sky, void* stream);

/* Model.forward entry on caller-supplied points (no ray structure): encodes xyz (P,3), builds aux
 * from per-point sun_d (P,3) and t (P,tau), sky per point (P,3).
 * replaces the input side of satnerf.py:208-222 / rs_semantic.py:260-277,325-328. */
int snb_encode_points(const float* xyz, const float* sun_d, const float* t, int tau,
                      const float* sky_w1, const float* sky_b1, const float* sky_w2, const float* sky_b2,
                      int sky_hidden, int n_points, int model_kind, void* enc, void* aux, float* sky,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  the MLP (trunk + heads), forward and backward, as tcgen05 GEMMs
 * replaces SatNeRF.forward (baseline/models/satnerf.py:208-255) and RSSemanticNeRF.forward/.sigma/
 * .semantic (semantic/models/rs_semantic.py:260-340) and their autograd.
 * ---------------------------------------------------------------------------------------------- */
typedef struct snb_model snb_model; /* opaque host object: architecture + packed-weight layout */

/* variant (semantic model only, 0 = the shipped rs_semantic.toml): head-input variants of semantic/models/rs_semantic.py -
 * SNB_VARIANT_TJ_FOR_S: the semantic head reads cat(f, t) (`use_tj_for_s`, :207-211,330-338);
 * SNB_VARIANT_TJ_INSTEAD_OF_BETA: the colour head reads cat(f, t) (`use_tj_instead_of_beta`, :186-189,287-288).
 * Both are four more weight columns of their hidden block against the aux K-segment [1, sun_d, t].
 * SNB_VARIANT_SEPARATE_BETA_S: a second uncertainty head `semantic_beta_from_xyz` (`use_separate_beta_for_s`, :228-237,
 * 297-303): cat(f, t) -> 256 -> softplus, packed column 9, the class scores move to columns 10.. (n_classes <= 9).
 * SNB_VARIANT_SEPARATE_TJ_S: the semantic head (with TJ_FOR_S) and the semantic uncertainty head read a SECOND embedding t_s
 * (`use_separate_tj_for_semantic`, :300-301,334-335): the caller passes both tables side by side as one (vocab, 2 tau) table
 * to snb_sample_encode / a (P, 2 tau) `t` to snb_encode_points / snb_mlp_forward_fp32 (their `tau` argument is 2 tau then),
 * t_s lands in aux columns 4+tau..4+2tau and its gradient in the same columns of g_aux.
 * SNB_VARIANT_FULL_FEATURES (SatNeRF and the semantic model): `fc_use_full_features` - the head hidden layers and sky_color
 * are fc_units = 512 wide instead of 256 (satnerf.py:123-124, rs_semantic.py:147-148).
 * SNB_VARIANT_RELU (SatNeRF and the semantic model): `activation_function != "siren"` (rs_semantic.py:150,158; SatNeRF's
 * `siren=False`, satnerf.py:127,146) - every hidden activation is a ReLU, the first trunk layer has no w0 = 30.
 * t_embedding_tau: width of the per-image embedding (`t_embedding_tau`, 4 in every shipped TOML; 0 = 4): at most 12, 6 with
 * SNB_VARIANT_SEPARATE_TJ_S - the per-ray inputs travel in 16 aux columns [1 | sun_d | t | t_s].
 * mapping_pos_n_freq (semantic model; 0 = 10): frequencies of the positional mapping of xyz (`mapping_pos_n_freq`, 1..10).
 * snb_sample_encode always writes the 10-frequency row; a model with fewer has 6 L input columns in its first and skip
 * layers and zero packed weights for the frequencies it does not have. */
enum { SNB_VARIANT_TJ_FOR_S = 1, SNB_VARIANT_TJ_INSTEAD_OF_BETA = 2, SNB_VARIANT_SEPARATE_BETA_S = 4,
       SNB_VARIANT_SEPARATE_TJ_S = 8, SNB_VARIANT_FULL_FEATURES = 16, SNB_VARIANT_RELU = 32 };
int snb_model_create(snb_model** out, int model_kind, int n_classes, int semantic_sigmoid, int variant, int t_embedding_tau,
                     int mapping_pos_n_freq);
void snb_model_destroy(snb_model* m);
/* number of fp32 parameters / the offset table: parameters live in ONE flat fp32 buffer in the
 * reference state_dict order; names[i], offsets[i], rows[i], cols[i]. */
int64_t snb_model_param_count(const snb_model* m);
int snb_model_num_tensors(const snb_model* m);
int snb_model_tensor_info(const snb_model* m, int i, const char** name, int64_t* offset, int* rows, int* cols);
/* bytes of the packed bf16 weight image (forward + transposed copies) */
size_t snb_model_packed_bytes(const snb_model* m);
/* fp32 flat params -> packed bf16 image (call after every optimiser step) */
int snb_model_pack(const snb_model* m, const float* params, void* packed, void* stream);

/* workspace (activations saved for backward, scratch) for P points; train=0: inference only */
size_t snb_mlp_workspace_bytes(const snb_model* m, int64_t n_points, int train);

/* forward: enc/aux/sky from K1 -> out (P, n_out) f32 packed [rgb 0:3 | sigma 3 | sun 4 | sky 5:8 | beta 8 | sem 9:]
 * (rs_semantic.py:291-311).  Columns of heads not in head_mask are written as 0.
 * rows_per_ray: sky/aux are per ray when >0 (sky indexed by row / rows_per_ray), per point when 0. */
int snb_mlp_forward(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                    int64_t n_points, const void* enc, const void* aux, const float* sky,
                    int rows_per_ray, int head_mask, int train, float* out, void* stream);

/* backward: g_out (P, n_out) f32 (gradient w.r.t. `out`) -> accumulates into grads (flat fp32, same
 * layout as params), g_aux (P,16) f32 [.., d t(tau) at cols 4..] (NULL to skip), g_sky (P or N,3) summed
 * by the caller.  Must follow snb_mlp_forward(train=1) on the same workspace. */
/* bucket_events: NULL, or 3 cudaEvent_t (entries may be NULL).  The weight gradients complete in three buckets - heads,
 * trunk layers 4-7, trunk layers 0-3 (flat ranges: snb_model_grad_buckets) - and event b is recorded on `stream` as soon
 * as bucket b of `grads` is final, so a data-parallel caller can start that bucket's all-reduce on a side stream while
 * the remaining weight-gradient GEMMs still run ("bucketed ... overlapped with K2 backward"). */
int snb_mlp_backward(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                     int64_t n_points, const void* enc, const void* aux, const float* out,
                     const float* g_out, int head_mask, float* grads, float* g_aux, void* const* bucket_events,
                     void* stream);
/* One training step's main pass AND its solar-correction pass (baseline/components/rendering.py:103-118: the same model
 * evaluated at the points along the sun direction, sigma and sun heads only) on ONE workspace of n_points + n_solar_points
 * rows (snb_mlp_workspace_bytes of the sum): rows [0, n_points) of enc / out / g_out are the main pass (all heads), rows
 * [n_points, n_points + n_solar_points) the solar pass; aux row i serves both (n_solar_points <= n_points; 0 = no solar pass).
 * Forward = the two passes' chained launches.  Backward = the two passes' dgrad chains, then ONE weight-gradient GEMM per
 * layer both passes share (trunk, sun layers, head outputs) over all rows, one set of folded-layer products and one
 * unpack per bucket - instead of two of each with snb_mlp_backward called per pass.  Results equal the two-call form. */
int snb_mlp_forward_with_solar(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                               int64_t n_points, int64_t n_solar_points, const void* enc, const void* aux,
                               const float* sky, int rows_per_ray, float* out, void* stream);
int snb_mlp_backward_with_solar(const snb_model* m, const void* packed, void* workspace, size_t workspace_bytes,
                                int64_t n_points, int64_t n_solar_points, const void* enc, const void* aux,
                                const float* out, const float* g_out, float* grads, float* g_aux,
                                void* const* bucket_events, void* stream);
/* flat element ranges [lo[b], hi[b]) of the three gradient buckets, in completion order */
int snb_model_grad_buckets(const snb_model* m, int64_t* lo3, int64_t* hi3);

/* NeRF only: aux (P, 32) bf16 = [1, sin/cos(2^k d)_{k<4} (24, commons.py:68-74 order), 0 x 7] from the per-ray view
 * directions dirs (N, 3) f32 (row stride `stride` floats), broadcast over the ray's n_samples samples.  Replaces
 * Mapping(4, 3)(input_dir) + repeat_interleave (baseline/models/nerf.py:33-37,197-199). */
int snb_nerf_aux(const float* dirs, int stride, int n_rays, int n_samples, void* aux32, void* stream);

/* fp32 verification mode ("fp32 mode" of the parity contract: rgb / depth within 1e-3 of the reference's fp32 CPU path
 * for ANY weights, not only at the initialisers).  Same forward as snb_mlp_forward - Model.forward of
 * satnerf.py:208-255 / rs_semantic.py:260-340 - with fp32 inputs, the fp32 weights straight from the flat parameter buffer
 * and fp32 FMA accumulation on the CUDA cores (no bf16 anywhere, no tensor cores: B200 has no fp32 MMA).  Inference only.
 *   xyz   (P,3) f32 sample positions (o + d*z, or o + sun_d*z for the solar-correction pass)
 *   sun_d (R,3), t (R,tau), sky (R,3) f32: R = P / rows_per_ray rows when rows_per_ray > 1 (one per ray, broadcast over
 *         the ray's samples like repeat_interleave in rs_semantic.py:42-61), else one row per point
 *   out   (P, n_out) f32, packed like snb_mlp_forward; heads outside head_mask are written as 0 */
size_t snb_mlp_fp32_workspace_bytes(const snb_model* m, int64_t n_points);
int snb_mlp_forward_fp32(const snb_model* m, const float* params, void* workspace, size_t workspace_bytes,
                         int64_t n_points, const float* xyz, const float* sun_d, const float* t, const float* sky,
                         int rows_per_ray, int head_mask, float* out, void* stream);

/* sky_color / embedding parameter gradients (tiny per-ray kernel).
 *   sky   (N,3) f32 from K1 (NULL: skip the sky_color gradients)
 *   g_out (P, n_out) f32 - columns 5:8 (sky) are summed per ray
 *   g_aux (P,16) f32 from snb_mlp_backward - columns 4..4+tau are summed per ray and scattered into
 *         g_t_table (vocab,tau) by ts = extras[:,3]  (NULL: skip the embedding gradient)
 * replaces autograd through sky_color (satnerf.py:188-193,248) and nn.Embedding
 * (semantic/components/rendering.py:42). */
int snb_ray_param_backward(const snb_model* m, const float* params, const float* extras, const float* sky,
                           const float* g_out, const float* g_aux, int n_rays, int n_samples, int n_out,
                           int tau, int vocab, float* grads, float* g_t_table, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3  per-ray compositing
 * replaces framework/util/rendering.py:4-34 (convert_sigmas) and the tails of `inference`
 * (satnerf.py:73-96, rs_semantic.py:81-126, _logit_to_label :131-136) and their autograd.
 * ---------------------------------------------------------------------------------------------- */
/* flags: SNB_COMPOSITE_NO_CLAMP - the composited colour is NOT clamped to [0, 1]: NeRF's inference
 * (baseline/models/nerf.py:73-86) returns the raw sum, SatNeRF / S-NeRF / semantic clamp (satnerf.py:79, rs_semantic.py:103);
 * SNB_COMPOSITE_BETA_S - packed column 9 is the separate semantic uncertainty (`use_separate_beta_for_s`,
 * semantic/models/rs_semantic.py:90-96): the class scores start at column 10 instead of 9; the fused loss uses its composited
 * value for the uncertainty-weighted semantic loss and adds its log term (loss_terms[7], without the constant 3/2 lambda_s) */
enum { SNB_COMPOSITE_NO_CLAMP = 1, SNB_COMPOSITE_BETA_S = 2 };
int snb_composite_forward(const float* out, const float* z_vals, int n_rays, int n_samples, int n_out,
                          int n_classes, int flags, float* rgb, float* depth, float* weights, float* transparency,
                          float* sem_logits, int64_t* sem_label, void* stream);

/* any g_* may be NULL (treated as zero).  g_out_direct: gradient flowing straight into the
 * per-sample tensors (albedo/sun/sky/beta/sigmas views of `out`).  Writes g_out (P, n_out). */
int snb_composite_backward(const float* out, const float* z_vals, int n_rays, int n_samples, int n_out,
                           int n_classes, int flags, const float* g_rgb, const float* g_depth, const float* g_weights,
                           const float* g_transparency, const float* g_sem_logits,
                           const float* g_out_direct, float* g_out, void* stream);

/* K3 + losses fused: composite forward, the loss terms that sit directly on its outputs and the
 * composite backward in one pass per ray; the per-sample weights / transparency / beta / sun_sc tensors are never
 * materialised.  Replaces, for the training step, snb_composite_forward + the loss modules + snb_composite_backward:
 *   mode 0 (main pass)   SNerfLoss colour term (color = 0, baseline/components/loss.py:71-94) or SatNerfLoss's
 *                        uncertainty_aware_loss (color = 1, loss.py:16-27) + SemanticLoss (cross-entropy with
 *                        ignore_index, semantic/components/loss.py:35-65) + SemanticCarRegLoss (loss.py:117-157)
 *   mode 1 (solar pass)  solar_correction terms 2 and 3 (baseline/components/loss.py:4-13)
 *   mode 2 (depth batch) DepthLoss (baseline/components/loss.py:30-47)
 *   mode 3 (statistics)  pre-pass of the uncertainty-weighted semantic loss: loss_terms[0] += sum of the per-ray
 *                        cross-entropies, loss_terms[1] += sum_r 1 / (2 beta_r^2); g_out may be NULL (nothing is written).
 *                        Pass `counts + 4` as loss_terms, then run mode 0 with sem_unc != 0 and the same `counts`
 *                        (float[6]; data parallel: all-reduce entries 4-5 with the rest).
 * gt_rgb (N,3) f32; labels (N) i64 or NULL; depth_gt / depth_w (N) f32 (depth_w NULL = 1);
 * ray_mask (N) u8 or NULL: the reference's `semantic_sparsity_mask` (semantic/dataset/semantic_dataset.py:65,87, passed to
 *   all semantic losses by semantic/components/training_step.py:58-88) - rays with mask 0 enter neither the cross-entropy nor
 *   the car regularisation;
 * counts: device float[>=2] = {rays in the cross-entropy mean, rays in the car-regularisation mean} as snb_label_counts
 *   produces them (NULL without labels).  Data parallel: all-reduce (sum) them first and pass inv_n = 1 / GLOBAL rays, so the
 *   per-rank terms add up to the global-batch loss and the summed gradients are the global-batch gradients;
 * g_out (P, n_out) f32 = d(sum of the terms)/d(out); loss_terms: device float[8], ACCUMULATED:
 * [0] colour [1] log-beta without its constant 3/2 [2] cross-entropy [3] car reg [4] sc term 2 [5] sc term 3 [6] depth
 * [7] log-beta of the separate semantic uncertainty (without its constant 3/2 lambda_s). */
typedef struct snb_loss_params {
  int mode, color;
  float beta_min, inv_n;     /* inv_n = 1 / (rays of the batch): the means of the reference losses */
  float lambda_s;
  int ignore_index;
  float lambda_c;
  int car_label;
  float lambda_sc, lambda_ds;
  int flags;                 /* SNB_COMPOSITE_* */
  int sem_unc;               /* mode 0: 0 = SemanticLoss; 1 = SemanticUncertaintyLoss (`use_beta_for_s`, semantic/components/
                              * loss.py:6-32,68-114: lambda_s * CE_mean * mean_r 1 / (2 beta_r^2)); 2 = the same with beta
                              * detached (`detach_beta_for_s`).  Needs the statistics of a mode-3 pre-pass in counts[4:6]. */
} snb_loss_params;
int snb_composite_loss(const float* out, const float* z_vals, int n_rays, int n_samples, int n_out, int n_classes,
                       const float* gt_rgb, const int64_t* labels, const uint8_t* ray_mask, const float* depth_gt,
                       const float* depth_w, const float* counts, const snb_loss_params* p, float* g_out,
                       float* loss_terms, void* stream);

/* The masked-mean denominators of the semantic losses, with exactly the predicates snb_composite_loss applies (ACCUMULATED
 * into counts, device float[3]; zero it first):
 *   counts[0] += rays with mask != 0, label != ignore_index and 0 <= label < n_classes   (CrossEntropyLoss mean, loss.py:41-55)
 *   counts[1] += rays with mask != 0 and label == car_label                               (SemanticCarRegLoss, loss.py:131-147)
 *   counts[2] += labels outside [0, n_classes) that are not ignore_index - torch's CrossEntropyLoss raises for those; the
 *                caller checks this entry where it can afford a device read (the ray table does it once per table). */
int snb_label_counts(const int64_t* labels, const uint8_t* ray_mask, int n_rays, int n_classes, int ignore_index,
                     int car_label, float* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * optimiser step on the flat buffers (Adam, torch.optim.Adam semantics, weight_decay = 0:
 * baseline/pipelines/base_ray_pipeline.py:246-269).  grad_scale multiplies the gradient first
 * (1/world_size after the all-reduce). */
/* step_dev: optional DEVICE int; when non-NULL the bias corrections use *step_dev instead of `step` (CUDA-graph replay). */
int snb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, const int* step_dev, float grad_scale,
                  void* stream);

/* tuning / test hook: run the MLP as ONE chained persistent kernel per pass (1, the default; inter-layer
 * activations are read back from L2) or as one GEMM launch per layer (0).  Both give the same results; the
 * parity tests compare them at sizes where every SM pair carries several row blocks.  Returns the old value. */
int snb_set_chained_mlp(int on);

/* ------------------------------------------------------------------------------------------------
 * measurement hooks (bench.py): count kernel launches made by this library and time every GEMM
 * launch with CUDA events on the launching stream.  snb_profile_end synchronises the device and
 * returns the summed GEMM device time (ms), the number of GEMM launches, all launches, and the MACs
 * the GEMM launches executed (padded tile work, for reference next to the algorithmic count). */
void snb_profile_begin(int time_gemms);
/* the running launch counter, and a way to credit launches that did not pass through the host entry points: a step captured
 * into a CUDA graph launches its kernels on every replay without calling them (the caller adds the captured count). */
int64_t snb_profile_launch_count(void);
void snb_profile_add_launches(int64_t n);
int snb_profile_end(double* gemm_ms, int64_t* gemm_launches, int64_t* total_launches, double* gemm_macs);

/* test hook (host only, no device needed): the row blocks SM pair `pair` of `n_pairs` carries through a chained launch of
 * one or two passes (n_blocks1 = 0: one), in execution order, as (pass << 24 | block) words; returns their number.
 * shift1: the rotation of the second pass's dealing (the library uses n_blocks0 % n_pairs). */
int snb_chain_schedule(int n_blocks0, int n_blocks1, int shift1, int n_pairs, int pair, int* out, int cap);

/* ------------------------------------------------------------------------------------------------
 * test hook: one bf16 GEMM through the tcgen05 kernel.  D = A (M,K) x B (N,K)^T (+ epilogue).
 * a_mn / b_mn = 1: operand stored (K,M) / (K,N) row-major (the wgrad form).
 * epi: 0 sin(w0*(acc+bias)) -> out0 bf16 [+ out1 = (M, N/32) u32 sign mask: bit c of row r = cos(..) < 0],
 *      1 acc+bias -> bf16, 2 acc*mul -> bf16 [out1 = sign mask: acc * w0 * (+-)sqrt(1 - mul^2), the SIREN
 *      derivative rebuilt from the saved activation mul = sin(..)]  (0-2: N % 256 == 0, K-major operands),
 *      4 f32 rows (N==16), 5 f32 += acc (split-K reduce).
 * ---------------------------------------------------------------------------------------------- */
int snb_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int N, int K,
                  int a_mn, int b_mn, int epi, void* out0, void* out1, int64_t ldo, const void* mul,
                  const float* bias, float w0, int splits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB_H_ */
