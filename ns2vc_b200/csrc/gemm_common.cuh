// Device helpers shared by the tcgen05 GEMM and the SIMT debug GEMM: A-operand row mapping,
// the fused normalisation/activation transforms, and the epilogue math.
#pragma once
#include "common.cuh"

namespace ns2vc {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// erf-GELU, as F.gelu default (reference attention.py:295)
__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// Source row for output row (b, t) and a segment tap; returns -1 for a zero-padded row.
__device__ __forceinline__ long long a_src_row(const GemmOp& op, int b, int t, int tap) {
  const int u = t * op.stride + tap;
  if (u < 0 || u >= op.T_virt) return -1;
  const int r = op.rowmap ? __ldg(op.rowmap + u) : u;
  return (long long)b * op.T_src + r;
}

// One A element (used by the SIMT kernel and as the definition the vector paths must match).
__device__ __forceinline__ float a_fetch(const GemmOp& op, const ASeg& s, int b, long long srow, int c) {
  if (srow < 0 || c >= s.nch) return 0.f;
  float x = __ldg(s.src + srow * s.ld + s.ch0 + c);
  switch (s.mode) {
    case A_AFFINE:
      x = fmaf(x, __ldg(s.p0 + (long long)b * s.ald + s.aoff + c), __ldg(s.p1 + (long long)b * s.ald + s.aoff + c));
      break;
    case A_AFFINE_SILU:
      x = fmaf(x, __ldg(s.p0 + (long long)b * s.ald + s.aoff + c), __ldg(s.p1 + (long long)b * s.ald + s.aoff + c));
      x = silu_f(x);
      break;
    case A_LN: {
      const float mean = __ldg(s.p0 + 2 * srow), rstd = __ldg(s.p0 + 2 * srow + 1);
      x = (x - mean) * rstd * __ldg(s.p1 + s.ch0 + c) + __ldg(s.p2 + s.ch0 + c);
      break;
    }
    default: break;
  }
  return x;
}

// Epilogue for one accumulator value at (m = b*T_out + t, packed column n).  For GEGLU the caller
// passes the value accumulator in `acc` and the gate accumulator in `acc_gate`, and n is the
// logical output column.
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
