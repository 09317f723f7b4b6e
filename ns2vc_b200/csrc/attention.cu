// Scaled-dot-product attention (reference attention_processor.py:1025-1036 -> F.scaled_dot_product_attention):
//   out = softmax(q k^T * dh^-0.5 + bias) v,   8 heads, dh = C/8, no positional term.
// v1: fp32 flash-style kernel on the FMA pipes — one query row per thread, K/V tiles staged in
// shared memory (all threads read the same K/V element: broadcast, conflict-free), online softmax
// over 16-key chunks with exp2f and the scale*log2(e) folded into q.
// Self-attention: k/v come from the fused QKV buffer; cross-attention: from the per-utterance
// K/V cache with the additive mask bias (0 / -10000, NOT -inf: reference unet_1d_condition.py:817).
#include "common.cuh"
#include <math.h>

namespace ns2vc {

constexpr int kAttnThreads = 128;   // query rows per CTA
constexpr int kAttnKeys = 64;       // keys per shared-memory tile
constexpr int kChunk = 16;

template <int DH>
__global__ void __launch_bounds__(kAttnThreads) attn_kernel(AttnOp op) {
  __shared__ float Ks[kAttnKeys][DH];
  __shared__ float Vs[kAttnKeys][DH];
  __shared__ float Bs[kAttnKeys];
  const int b = blockIdx.z, h = blockIdx.y;
  const int tq = blockIdx.x * kAttnThreads + threadIdx.x;
  const bool qv = tq < op.Tq;
  const float qs = op.scale * 1.4426950408889634f;       // fold log2(e): softmax via exp2
  float q[DH], o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; o[d] = 0.f; }
  if (qv) {
    const float* qp = op.q + ((long long)b * op.Tq + tq) * op.q_ld + h * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] = qp[d] * qs;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int k0 = 0; k0 < op.Tk; k0 += kAttnKeys) {
    const int nk = min(kAttnKeys, op.Tk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kAttnKeys * DH; i += kAttnThreads) {
      const int j = i / DH, d = i % DH;
      float kv = 0.f, vv = 0.f;
      if (j < nk) {
        const long long row = (long long)b * op.Tk + k0 + j;
        kv = op.k[row * op.k_ld + h * DH + d];
        vv = op.v[row * op.v_ld + h * DH + d];
      }
      Ks[j][d] = kv;
      Vs[j][d] = vv;
    }
    for (int j = threadIdx.x; j < kAttnKeys; j += kAttnThreads)
      Bs[j] = (j < nk) ? (op.bias ? op.bias[(long long)b * op.Tk + k0 + j] * 1.4426950408889634f : 0.f) : -INFINITY;
    __syncthreads();
#pragma unroll 1
    for (int c0 = 0; c0 < kAttnKeys; c0 += kChunk) {
      if (c0 >= nk) break;
      float s[kChunk];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < kChunk; ++j) {
        float a = Bs[c0 + j];
#pragma unroll
        for (int d = 0; d < DH; ++d) a = fmaf(q[d], Ks[c0 + j][d], a);
        s[j] = a;
        cmax = fmaxf(cmax, a);
      }
      const float mnew = fmaxf(mrun, cmax);
      const float corr = exp2f(mrun - mnew);           // mrun = -inf on the first chunk -> 0
      lrun *= corr;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
      for (int j = 0; j < kChunk; ++j) {
        const float p = exp2f(s[j] - mnew);            // padded keys: s = -inf -> 0
        lrun += p;
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = fmaf(p, Vs[c0 + j][d], o[d]);
      }
      mrun = mnew;
    }
  }
  if (qv) {
    const float inv = 1.0f / lrun;
    if (op.out) {
      float* po = op.out + ((long long)b * op.Tq + tq) * op.out_ld + h * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) po[d] = o[d] * inv;
    }
    if (op.out_hi) {
      __nv_bfloat16* ph = op.out_hi + ((long long)b * op.Tq + tq) * op.out_split_ld + h * DH;
      __nv_bfloat16* pl = op.out_lo + ((long long)b * op.Tq + tq) * op.out_split_ld + h * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        const float v = o[d] * inv;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        ph[d] = hi;
        pl[d] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
    }
  }
}

template <int DH>
static int launch_dh(const AttnOp& op, cudaStream_t st) {
  dim3 grid(ceil_div(op.Tq, kAttnThreads), op.H, op.B);
  attn_kernel<DH><<<grid, kAttnThreads, 0, st>>>(op);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attention launch failed: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

int launch_attention(const AttnOp& op, cudaStream_t st) {
  if (op.Tk <= 0 || op.Tq <= 0) { set_error("attention: empty sequence"); return -1; }
  switch (op.dh) {
    case 4: return launch_dh<4>(op, st);
    case 8: return launch_dh<8>(op, st);
    case 12: return launch_dh<12>(op, st);
    case 16: return launch_dh<16>(op, st);
    case 24: return launch_dh<24>(op, st);
    case 32: return launch_dh<32>(op, st);
    case 48: return launch_dh<48>(op, st);
    case 64: return launch_dh<64>(op, st);
    default: set_error("attention: head dim %d not supported (4,8,12,16,24,32,48,64)", op.dh); return -1;
  }
}

}  // namespace ns2vc
