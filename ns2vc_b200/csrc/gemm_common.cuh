// Device helpers shared by the tcgen05 GEMM, the SIMT debug GEMM and the prep kernels.
#pragma once
#include "common.cuh"

namespace ns2vc {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// erf-GELU, as F.gelu default (reference attention.py:295)
// erf via Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far inside the fp32 parity budget); the libm
// erff costs ~3x the instructions and the GEGLU epilogue is issue-latency bound.
__device__ __forceinline__ float gelu_erf_f(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float erf_abs = 1.0f - p * t * __expf(-x * x);
  return 0.5f * v * (1.0f + copysignf(erf_abs, v));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// 8 fp32 values -> 16 bytes of bf16 hi and 16 bytes of bf16 lo (x = hi + lo to ~2^-17 relative)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  float h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
  hi.x = pack_bf16x2(h[0], h[1]); hi.y = pack_bf16x2(h[2], h[3]);
  hi.z = pack_bf16x2(h[4], h[5]); hi.w = pack_bf16x2(h[6], h[7]);
  lo.x = pack_bf16x2(v[0] - h[0], v[1] - h[1]); lo.y = pack_bf16x2(v[2] - h[2], v[3] - h[3]);
  lo.z = pack_bf16x2(v[4] - h[4], v[5] - h[5]); lo.w = pack_bf16x2(v[6] - h[6], v[7] - h[7]);
}

// One A element of a segment, as the TMA path sees it (zero outside the source).
__device__ __forceinline__ float a_fetch_split(const GemmOp& op, const GSeg& s, int b, int t, int c) {
  const SplitBuf& src = op.src[s.src];
  const int r = t + s.tap, ch = s.c0 + c;
  if (r < 0 || r >= src.T || ch >= src.C) return 0.f;
  const long long off = ((long long)b * src.T + r) * src.ld + ch;
  return __bfloat162float(src.hi[off]) + __bfloat162float(src.lo[off]);
}

// Epilogue for one accumulator value at (m = b*T_out + t, logical column n).  For GEGLU the caller
// passes the value accumulator in `acc` and the gate accumulator in `acc_gate`.
__device__ __forceinline__ float epi_value(const GemmOp& op, int b, long long m, int n, float acc, float acc_gate) {
  float v = acc;
  if (op.flags & EPI_GEGLU) {
    const int half = op.n_valid;   // = 4C
    v = (acc + __ldg(op.bias + n)) * gelu_erf_f(acc_gate + __ldg(op.bias + half + n));
  } else if (op.flags & EPI_BIAS) {
    v += __ldg(op.bias + n);
  }
  if (op.flags & EPI_ROWBIAS) v += __ldg(op.rowbias + (long long)b * op.rowbias_ld + n);
  if (op.flags & EPI_RESIDUAL) v += __ldg(op.res + m * op.res_ld + n);
  return v;
}

}  // namespace ns2vc
