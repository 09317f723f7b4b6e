/* ns2vc_b200 — C-ABI of the B200-native NS2VC denoiser hot path.
 *
 * The reference (adelacvg/NS2VC) is pure Python/PyTorch and defines no FFI; its boundary for this
 * path is the Python object graph.  This header is the C boundary the Python
 * drop-in modules (ns2vc_b200/unet.py, fused.py) bind with ctypes; each entry point names the
 * reference interface it stands in for.  Plain C types, raw device pointers + cudaStream_t,
 * int return (0 = ok, <0 = error; ns2vc_last_error() returns the message).  No exceptions or C++
 * types cross the ABI; the caller owns every buffer, the handle owns only its packed weights and, per
 * (B, T, S, workspace) it has seen, a launch program with ~20 KB of static device tables.
 * All calls are stream-ordered.  The FIRST prepare_cond / forward / time_table for a new (B, T, S, workspace)
 * builds that program (host work, one cudaMalloc, one host-to-device copy): it must not run under stream
 * capture - run a shape once eagerly, after which its calls allocate nothing, never synchronise and are
 * capturable in a CUDA graph.  One workspace may serve several shapes one after another (prepare_cond again
 * after a switch).  One handle per device, one run at a time; not thread-safe per handle.
 */
#ifndef NS2VC_B200_H
#define NS2VC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ns2vc_unet ns2vc_unet;
typedef void* ns2vc_stream;   /* cudaStream_t */

#define NS2VC_MAX_LEVELS 8

/* Constructor arguments of UNet1DConditionModel (reference unet1d/unet_1d_condition.py:151-203,
 * as called from model.py:391-400). */
typedef struct ns2vc_unet_cfg {
  int in_channels;            /* conv_in input channels = latent + content (356)                 */
  int latent_channels;        /* leading channels that change every step (100); the rest is the  */
                              /* step-invariant content embedding whose conv_in share is hoisted  */
  int out_channels;           /* 100                                                              */
  int n_levels;
  int block_out_channels[NS2VC_MAX_LEVELS];
  int layers_per_block[NS2VC_MAX_LEVELS];
  int down_has_attn[NS2VC_MAX_LEVELS];   /* CrossAttnDownBlock2D (1) / DownBlock2D (0)           */
  int up_has_attn[NS2VC_MAX_LEVELS];     /* CrossAttnUpBlock2D (1) / UpBlock2D (0)               */
  int num_heads;              /* reference: attention_head_dim reinterpreted as head count (:219) */
  int cross_attention_dim;
  int norm_num_groups;
  float norm_eps;
  int time_scale_shift;       /* resnet_time_scale_shift == 'scale_shift'                         */
  int add_embed_text;         /* addition_embed_type == 'text'                                    */
  int add_embed_heads;        /* addition_embed_type_num_heads (64)                               */
  int flip_sin_to_cos;
  float freq_shift;
} ns2vc_unet_cfg;

const char* ns2vc_last_error(void);

/* Build the layer plan for cfg (reference ctor, unet_1d_condition.py:421-559). */
int ns2vc_unet_create(const ns2vc_unet_cfg* cfg, ns2vc_unet** out);
void ns2vc_unet_destroy(ns2vc_unet* h);

/* state_dict contract (reference key names and shapes). */
int ns2vc_unet_num_weights(const ns2vc_unet* h);
int ns2vc_unet_weight_info(const ns2vc_unet* h, int i, const char** name, int64_t shape[4], int* ndim);
/* Copy one parameter (device fp32, contiguous) into the handle: load_state_dict per key. */
int ns2vc_unet_load_weight(ns2vc_unet* h, const char* key, const float* dptr, const int64_t* shape, int ndim,
                           ns2vc_stream stream);
/* Pack every contraction weight into the tcgen05 operand layout (bf16 hi/lo, swizzled tiles);
 * fails if any key is missing (strict load, reference inference/infer_tool.py:27). */
int ns2vc_unet_finalize(ns2vc_unet* h, ns2vc_stream stream);

int ns2vc_unet_workspace_bytes(const ns2vc_unet* h, int B, int T, int S, size_t* bytes);

/* Step-invariant conditioning (the part of Diffusion_Encoder.forward / UNet forward that does not
 * depend on x or t: model.py:406-411, unet_1d_condition.py:816-818, 869-870, cross-attention K/V
 * attention_processor.py:1017-1020).
 *   content : [B, in_channels-latent_channels, T] fp32, batch stride content_bstride floats (may be NULL if 0 ch)
 *   prompt  : [B, S, cross_attention_dim] fp32 contiguous (encoder_hidden_states)
 *   mask    : [B, S] uint8 (1 = attend), or NULL (no encoder_attention_mask)                    */
int ns2vc_unet_prepare_cond(ns2vc_unet* h, const float* content, long long content_bstride, const float* prompt,
                            const uint8_t* mask, int B, int T, int S, void* ws, ns2vc_stream stream);

/* UNet1DConditionModel.forward (unet_1d_condition.py:743-1037) given prepared conditioning.
 *   x [B, latent_channels, T] fp32 (batch stride x_bstride floats), t [B] fp32 -> out [B, out_channels, T] */
int ns2vc_unet_forward(ns2vc_unet* h, const float* x, long long x_bstride, const float* t, float* out, int B, int T,
                       int S, void* ws, ns2vc_stream stream);

/* The timestep path of many forwards at once (the sampler knows every evaluation time when the run starts):
 * Timesteps -> TimestepEmbedding (+ the pooled prompt embedding) -> time_emb_proj of all resnets
 * (reference embeddings.py:24-64, 157-218; unet_1d_condition.py:825-883; resnet.py:619-629).
 *   t_rows [n_rows] fp32 device, n_rows = steps x B in step-major order (row r belongs to batch entry r % B)
 *   table  device buffer of ns2vc_unet_time_table_floats(h, n_rows) floats; its first n_rows x film_width floats are the
 *          FiLM rows, the rest is scratch.  Needs ns2vc_unet_prepare_cond() on the same (B,T,S,workspace) first.        */
int ns2vc_unet_film_width(const ns2vc_unet* h);
size_t ns2vc_unet_time_table_floats(const ns2vc_unet* h, int n_rows);
int ns2vc_unet_time_table(ns2vc_unet* h, const float* t_rows, int n_rows, float* table, int B, int T, int S, void* ws,
                          ns2vc_stream stream);
/* ns2vc_unet_forward with the timestep path taken from B rows of such a table (film_rows = table + step * B * film_width). */
int ns2vc_unet_forward_film(ns2vc_unet* h, const float* x, long long x_bstride, const float* film_rows, float* out, int B,
                            int T, int S, void* ws, ns2vc_stream stream);

/* Per-step sampler math fused into one element-wise kernel (bit-exact fp32 op order).
 * DPM-Solver++(2M): model_wrapper x_start->noise (sampler/dpm_solver.py:291-292), data_prediction_fn
 * (:437-439), dpm_solver_first_update (:569-576), multistep_dpm_solver_second_update (:813-831). */
typedef struct ns2vc_dpm_coef {
  float alpha_s, sigma_s, c_x, c_m, c_d, inv_r0;
  int order;                  /* 0: x0 round trip only; 1 / 2: + first / second order update       */
} ns2vc_dpm_coef;
/* nan_flag (device int, may be NULL): set to 1 when x holds a NaN - the reference asserts on that every denoiser call
 * (model.py:404); the fused loop checks the flag once after the run instead of syncing every step. */
int ns2vc_dpm_step(const float* x, const float* unet_out, const float* m_prev, const ns2vc_dpm_coef* c, float* m_cur,
                   float* x_next, size_t n, int* nan_flag, ns2vc_stream stream);

/* UniPC-bh2 (sampler/uni_pc.py:471-588): corrector at t and predictor to the next time. */
typedef struct ns2vc_unipc_coef {
  float alpha_t, sigma_t, c_x, c_m, ab, rk, rho0, rho1;
  int corr_order;
  float n_c_x, n_c_m, nab, nrk;
  int pred_order;
} ns2vc_unipc_coef;
int ns2vc_unipc_step(const float* x_prev, const float* x_eval, const float* unet_out, const float* m0, const float* m1,
                     const ns2vc_unipc_coef* c, float* m_t, float* x_t, float* x_pred, size_t n, int* nan_flag,
                     ns2vc_stream stream);

/* Bit-exact index helpers (host, no GPU): nearest-neighbour source index of F.interpolate(size=)
 * (reference resnet.py:160) and the stride-2 conv length rule (resnet.py:200). */
int ns2vc_nearest_index(int t_in, int t_out, int* idx /* [t_out] */);
int ns2vc_down_length(int t);

/* encoder_attention_mask -> additive attention bias, (1 - m) * -10000 (reference unet_1d_condition.py:816-818): the kernel
 * prepare_cond runs, exposed for the bit-exact test.  mask [n] uint8 device, bias [n] fp32 device. */
int ns2vc_mask_bias(const uint8_t* mask, int n, float* bias, ns2vc_stream stream);

/* Diagnostics used by the parity tests. */
int ns2vc_unet_num_taps(const ns2vc_unet* h);
int ns2vc_unet_tap_info(const ns2vc_unet* h, int i, const char** name, int* level, int* channels);
int ns2vc_unet_set_tap(ns2vc_unet* h, int i, float* dst /* device [B, T_level, C] token-major, or NULL */);
const char* ns2vc_unet_plan_string(const ns2vc_unet* h);
int ns2vc_unet_launch_count(const ns2vc_unet* h);  /* kernels launched by the last forward */
const char* ns2vc_build_info(void);

/* Per-kernel-kind device timing (CUDA events around every launch on the caller's stream); used by
 * bench.py for the roofline line.  Off by default; never enable inside a timed region. */
int ns2vc_unet_set_profiling(ns2vc_unet* h, int on);
/* In-kernel stamps of CTA (0,0) of every GEMM launch of the next forwards, 32 slots per launch: [0,8) %globaltimer
 * (entry, prologue, PDL wait, first stage full, MMAs issued, accumulator ready, epilogue done, exit), [8,16) SM-clock
 * stamps of the epilogue sub-steps, [16,22) %globaltimer stamps of the fused prep (dependency wait done, loads issued,
 * affine ready, rows stored, all preps done, published to the cluster).  NULL disables. */
int ns2vc_unet_set_trace(ns2vc_unet* h, unsigned long long* device_buf, int n_gemms);
/* Diagnostics: attention launch i of the next forward writes per-key-tile SM-clock stamps of its CTA (0,0,0)
 * to device_buf[2048*i ...] ([16 tiles][16 slots], then [start ns, end ns, SM id] of up to 597 CTAs; see attention_v2.cu).
 * NULL disables. */
int ns2vc_unet_set_attn_trace(ns2vc_unet* h, unsigned long long* device_buf, int n_launches);
/* [min entry, max exit] %globaltimer of every launch of the next forwards (buffer pre-set to {~0, 0} pairs). */
int ns2vc_unet_set_span_trace(ns2vc_unet* h, unsigned long long* device_buf, int n_launches);
int ns2vc_unet_launch_kind(const ns2vc_unet* h, int launch_index);   /* index into ns2vc_profile_kind_name */
int ns2vc_profile_num_kinds(void);
const char* ns2vc_profile_kind_name(int kind);
int ns2vc_unet_profile_read(ns2vc_unet* h, int kind, double* ms_total, long long* launches);
int ns2vc_unet_profile_dump(ns2vc_unet* h, const char* csv_path);   /* one row per launch */
int ns2vc_unet_profile_reset(ns2vc_unet* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Condition encoders: `Pre_model.infer` (reference model.py:359-377) - the step immediately BEFORE the denoiser
 * (SURVEY.md 8(f) rank 1): ref_enc (TextTimeEmbedding, unet1d/embeddings.py:421-434), PromptEncoder and PhoneEncoder
 * (model.py:98-190: ConvLayer -> n x EncSALayer (operations.py:784-821: LayerNorm, 8-head self-attention with key padding,
 * LayerNorm, k=9 conv-FFN) -> ConvLayer -> LayerNorm), all frames past an utterance's length exactly zero.
 * Same conventions as the denoiser handle above: raw device pointers, caller-owned workspace, stream-ordered, int errors.
 * The first call for a new (B, T, S, workspace) builds the launch program on the host (no device allocation). */
typedef struct ns2vc_pre ns2vc_pre;
typedef struct ns2vc_pre_cfg {       /* reference config.json "phoneme_encoder" / "prompt_encoder" (model.py:332-340)        */
  int phone_in, phone_hidden, phone_out, phone_layers;       /* PhoneEncoder(in_channels, hidden_channels, out_channels, n_layers) */
  int prompt_in, prompt_hidden, prompt_out, prompt_layers;   /* PromptEncoder(...)                                          */
  int ref_dim;                       /* TextTimeEmbedding(100, 100, 1): width of the mel prompt (= prompt_in)                 */
  int ref_heads;                     /* 1                                                                                     */
  int n_heads;                       /* EncSALayer(c, 8, ...) (operations.py:961)                                             */
  int ffn_kernel;                    /* 9 (operations.py:963)                                                                 */
} ns2vc_pre_cfg;
int ns2vc_pre_create(const ns2vc_pre_cfg* cfg, ns2vc_pre** out);
void ns2vc_pre_destroy(ns2vc_pre* h);
int ns2vc_pre_num_weights(const ns2vc_pre* h);                                  /* state_dict contract: reference key names / shapes */
int ns2vc_pre_weight_info(const ns2vc_pre* h, int i, const char** name, int64_t shape[4], int* ndim);
int ns2vc_pre_load_weight(ns2vc_pre* h, const char* key, const float* dptr, const int64_t* shape, int ndim, ns2vc_stream stream);
int ns2vc_pre_finalize(ns2vc_pre* h, ns2vc_stream stream);                      /* strict: fails on a missing key                   */
int ns2vc_pre_workspace_bytes(const ns2vc_pre* h, int B, int T, int S, size_t* bytes);
/* Pre_model.infer:
 *   c [B, phone_in, T] fp32, refer [B, prompt_in, S] fp32 (contiguous), lengths / refer_lengths [B] int64 (device, each >= 1)
 *   -> content [B, T, phone_out], prompt [B, S, prompt_out] fp32 token-major (the reference returns the [T, B, C] / [S, B, C]
 *      views of the same values: model.py:145, 189).                                                                        */
int ns2vc_pre_infer(ns2vc_pre* h, const float* c, const float* refer, const int64_t* lengths, const int64_t* refer_lengths,
                    float* content, float* prompt, int B, int T, int S, void* ws, ns2vc_stream stream);
/* Diagnostics for the parity tests: per-layer activations (token-major [B, rows, channels]; rows = 1 for the speaker vector). */
int ns2vc_pre_num_taps(const ns2vc_pre* h);
int ns2vc_pre_tap_info(const ns2vc_pre* h, int i, const char** name, int* rows, int* channels);
int ns2vc_pre_set_tap(ns2vc_pre* h, int i, float* dst);
int ns2vc_pre_launch_count(const ns2vc_pre* h);   /* kernels launched by the last infer */

#ifdef __cplusplus
}
#endif
#endif /* NS2VC_B200_H */
