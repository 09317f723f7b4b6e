// Shared declarations for the ns2vc_b200 denoiser engine (sm_100a only).
//
// Internal activation layout: TOKEN-MAJOR fp32 [B, T_l, C] ("rows" = B*T_l tokens, channels
// contiguous).  The reference keeps [B, C, T] and permutes around every transformer block
// (reference unet1d/transformer_1d.py:264, 289); here the convs are implicit GEMMs over
// token rows (tap j of a k=3 conv is the same matrix shifted by j-1 rows), so no permute exists.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

namespace ns2vc {

// ---------------------------------------------------------------------------------------------
// Error handling: C-ABI returns negative codes; the message is kept per thread.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define NS_CHECK_CUDA(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ns2vc::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),    \
                       cudaGetErrorString(_e));                                                \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

#define NS_REQUIRE(cond, ...)                                                                  \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      ns2vc::set_error(__VA_ARGS__);                                                           \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

// ---------------------------------------------------------------------------------------------
// GEMM / implicit-conv operator descriptor (shared by the tcgen05 kernel and the SIMT debug
// kernel so both see byte-identical problem statements).
//
//   out[m, n] = epilogue( sum_seg sum_{c < nch} A_seg(m, c) * W[kofs_seg + c, n] )
//
// A_seg(m, c): m -> (b, t);  u = t*stride + tap;  zero if u outside [0, T_virt);
//              r = rowmap ? rowmap[u] : u;  x = src[(b*T_src + r)*ld + ch0 + c];
//              then the segment's transform (zero padding is applied AFTER the transform,
//              as a conv pads the normalised activations: reference resnet.py:597-612).
// ---------------------------------------------------------------------------------------------
enum AMode : int {
  A_RAW = 0,          // x
  A_AFFINE = 1,       // x * p0[b, aoff+c] + p1[b, aoff+c]              (GroupNorm folded)
  A_AFFINE_SILU = 2,  // silu(x * p0 + p1)                              (GroupNorm[+FiLM] + SiLU)
  A_LN = 3,           // (x - mean_m) * rstd_m * p1[c] + p2[c];  p0 = rowstats [M_src, 2]
};

struct ASeg {
  const float* src;   // token-major [B, T_src, ld]
  const float* p0;
  const float* p1;
  const float* p2;
  int ld;             // floats per source row
  int ch0;            // first source channel of this segment
  int nch;            // valid channels (k >= nch inside the segment's k-blocks reads as 0)
  int nkb;            // number of 64-wide k-blocks this segment occupies
  int tap;            // row offset (-1, 0, +1)
  int mode;           // AMode
  int aoff;           // channel offset into the affine arrays (second concat source)
  int ald;            // row pitch of the affine arrays (floats per batch entry)
};

enum EpiFlags : int {
  EPI_BIAS = 1,       // + bias[n]
  EPI_RESIDUAL = 2,   // + res[m*res_ld + n]
  EPI_GEGLU = 4,      // packed N pairs 64 value | 64 gate columns: out = (v+bv) * gelu_erf(g+bg)
  EPI_OUT_NCT = 8,    // store out[b, n, t] (channel-major, n < n_valid) instead of token-major
  EPI_ROWBIAS = 16,   // + rowbias[b*rowbias_ld + n]   (per-sample bias, time_embedding 'default')
};

constexpr int kMaxSeg = 8;

struct GemmOp {
  ASeg seg[kMaxSeg];
  int nseg;
  int nkb_total;      // sum of seg[i].nkb
  int B, T_out, T_src, T_virt;
  int stride;
  const int* rowmap;  // nearest-upsample index table [T_virt] or nullptr
  // B operand
  const __nv_bfloat16* w_hi;   // packed [kb][Npad][64] (128B-swizzled rows)
  const __nv_bfloat16* w_lo;
  const float* w_f32;          // debug SIMT backend: [K_pad][Npad] fp32 (nullptr unless enabled)
  int N;                       // packed output columns (multiple of 64)
  // epilogue
  int flags;
  const float* bias;           // [N] (GEGLU: [2*N_out] in the reference's value|gate order)
  const float* rowbias;
  int rowbias_ld;
  const float* res;
  int res_ld;
  float* out;
  int out_ld;                  // token-major pitch, or (EPI_OUT_NCT) unused
  int n_valid;                 // logical output columns written (<= N, or N/2 for GEGLU)
};

// Launchers (each returns 0 or a negative error code; all stream-ordered, no host sync).
int launch_gemm_tc(const GemmOp& op, cudaStream_t st);
int launch_gemm_simt(const GemmOp& op, cudaStream_t st);

// Weight packing (device side, load time).  Source W is the reference parameter layout
// [n_rows, cin_total, ktaps] fp32 (ktaps = 1 for nn.Linear / 1x1 conv).
struct PackSeg {
  const float* w;     // device pointer
  int n_rows;         // rows of w used (output channels from this tensor)
  int cin_total;
  int ktaps;
  int tap;            // which tap (0..ktaps-1)
  int cin0, ncin;     // channel range of w's input axis covered by this segment
  int n_dst0;         // first packed column
  int kb0;            // first k-block in the packed K order
  int nkb;
  int geglu_half;     // 0: plain. >0: interleave value/gate (value rows [0,half), gate rows [half,2*half))
};
int launch_pack_b(const PackSeg& ps, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, float* w_f32, int Npad,
                  cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Attention
// ---------------------------------------------------------------------------------------------
struct AttnOp {
  const float* q; int q_ld;       // [B, Tq, q_ld], head h at column h*dh
  const float* k; int k_ld;       // [B, Tk, k_ld]
  const float* v; int v_ld;
  const float* bias;              // additive [B, Tk] or nullptr      (reference: 0 / -10000)
  float* out; int out_ld;         // [B, Tq, out_ld]
  int B, H, Tq, Tk, dh;
  float scale;                    // dh^-0.5
};
int launch_attention(const AttnOp& op, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Norm statistics and small kernels (kernels_misc.cu)
// ---------------------------------------------------------------------------------------------
// GroupNorm statistics over a (possibly two-source, channel-concatenated) token-major tensor,
// finalised into a per-(b, c) affine:  scale = rstd*gamma*(1+film_s), shift = (beta - mean*rstd*gamma)*(1+film_s) + film_b
struct GnOp {
  const float* src1; int ld1; int C1;
  const float* src2; int ld2; int C2;    // src2 may be nullptr (C2 = 0)
  int B, T, G;
  float eps;
  const float* gamma; const float* beta; // [C1+C2]
  const float* film;  int film_ld;       // nullptr or [B, film_ld]: scale at film[b, c], shift at film[b, C + c]
  float* scale; float* shift;            // [B, C]
  double* acc;                           // [B*G*2] zero on entry, zero on exit
  unsigned* counter;                     // [B*G]   zero on entry, zero on exit
};
int launch_gn_affine(const GnOp& op, cudaStream_t st);

int launch_ln_stats(const float* x, int ld, int M, int C, float eps, float* stats /*[M,2]*/, cudaStream_t st);
int launch_ln_apply(const float* x, int ld, int M, int C, float eps, const float* gamma, const float* beta,
                    float* y, int y_ld, cudaStream_t st);

// [B, C, T] (batch stride bstride) -> token-major [B, T, ldo] (channels >= C zero-filled up to Cpad)
int launch_nct_to_tokens(const float* x, long long bstride, int B, int C, int T, float* out, int ldo, int Cpad,
                         cudaStream_t st);
// token-major [B, T, ld] -> [B, C, T]
int launch_tokens_to_nct(const float* x, int ld, int B, int C, int T, float* out, cudaStream_t st);

enum LinIn : int { LIN_RAW = 0, LIN_SILU = 1, LIN_SINUSOID = 2 };
// Small-M linear: out[m, n] = f(x[m, :]) . W[n, :] + bias[n] (+ add[m, n]);  W row-major [N, K].
// LIN_SINUSOID: x is t[m] and the input row is the K-wide sinusoidal embedding
// (reference embeddings.py:24-64, flip_sin_to_cos / freq_shift as given).
struct LinOp {
  const float* x; int x_ld; int M; int K;
  const float* W; const float* bias; int N;
  const float* add; int add_ld;
  float* out; int out_ld;
  int in_mode; int flip_sin_to_cos; float freq_shift;
  int out_silu;
};
int launch_small_linear(const LinOp& op, cudaStream_t st);

// AttentionPooling pieces (reference embeddings.py:499-546)
int launch_pool_class_token(const float* xn /*[B,S,C] LN'd*/, const float* pos /*[C]*/, int B, int S, int C,
                            float* tokens /*[B,S+1,C]: row0 = class token, rows 1.. = xn*/, cudaStream_t st);
int launch_pool_attend(const float* q /*[B,C]*/, const float* kv /*[B,S+1,2C] k|v*/, int B, int S1, int C, int heads,
                       float* out /*[B,C]*/, cudaStream_t st);
int launch_mask_bias(const uint8_t* mask, int n, float* bias, cudaStream_t st);

// Fused sampler steps (element-wise, bit-exact op order; see kernels_misc.cu)
struct DpmStepCoef {   // DPM-Solver++(2M): one post-UNet step
  float alpha_s, sigma_s;      // at the time the UNet was evaluated (x0 round trip)
  float c_x;                   // sigma_t / sigma_s
  float c_m;                   // alpha_t * expm1(-h)
  float c_d;                   // 0.5 * c_m
  float inv_r0;                // 1 / r0        (order 2 only)
  int order;                   // 0: round trip only, 1: first-order update, 2: second-order update
};
int launch_dpm_step(const float* x, const float* unet_out, const float* m_prev, const DpmStepCoef& c,
                    float* m_cur, float* x_next, size_t n, cudaStream_t st);

struct UniPcStepCoef {  // UniPC-bh2, data prediction: corrector at t (+ predictor to t_next)
  float alpha_t, sigma_t;      // x0 round trip at t
  // corrector at t from (x_prev at t_p0, m0, m1):  x_t = xbar - ab*(rho0*D1 + rho1*(m_t - m0))
  float c_x, c_m;              // xbar = c_x * x_prev - c_m * m0
  float ab;                    // alpha_t * B_h
  float rk;                    // D1 = (m1 - m0) / rk   (order-2 corrector)
  float rho0, rho1;
  int corr_order;              // 0: no corrector (first call: history only), 1, 2
  // predictor to t_next from (x_t, m_t, m0):  x_pred = nbar - nab*(0.5*D1n), D1n = (m0 - m_t)/nrk
  float n_c_x, n_c_m, nab, nrk;
  int pred_order;              // 0: none, 1: xbar only, 2: with D1n
};
int launch_unipc_step(const float* x_prev, const float* x_eval, const float* unet_out, const float* m0,
                      const float* m1, const UniPcStepCoef& c, float* m_t, float* x_t, float* x_pred, size_t n,
                      cudaStream_t st);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace ns2vc
