// Small HBM/L2-bound kernels of the denoiser step: layout conversion, GroupNorm/LayerNorm
// statistics, the timestep/FiLM GEMV, AttentionPooling pieces, and the fused sampler updates.
// All are coalesced/vectorised; none is worth tensor cores.
#include "gemm_common.cuh"
#include "prep_common.cuh"
#include "launch.cuh"
#include <cstdarg>
#include <cstdio>
#include <math.h>

namespace ns2vc {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

#define NS_LAUNCH_CHECK()                                                                      \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      set_error("%s:%d launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e));        \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Layout conversion.  32x32 smem tile transpose, coalesced on both sides.
// ---------------------------------------------------------------------------------------------
__global__ void nct_to_tokens_kernel(const float* __restrict__ x, long long bstride, int C, int T,
                                     float* __restrict__ out, int ldo, int Cpad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xb = x + (long long)b * bstride;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? xb[(long long)c * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < Cpad) out[((long long)b * T + t) * ldo + c] = tile[threadIdx.x][i];
  }
}
int launch_nct_to_tokens(const float* x, long long bstride, int B, int C, int T, float* out, int ldo, int Cpad,
                         cudaStream_t st) {
  dim3 grid(ceil_div(T, 32), ceil_div(Cpad, 32), B), block(32, 8);
  nct_to_tokens_kernel<<<grid, block, 0, st>>>(x, bstride, C, T, out, ldo, Cpad);
  NS_LAUNCH_CHECK();
  return 0;
}

__global__ void tokens_to_nct_kernel(const float* __restrict__ x, int ld, int C, int T, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? x[((long long)b * T + t) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) out[((long long)b * C + c) * T + t] = tile[threadIdx.x][i];
  }
}
int launch_tokens_to_nct(const float* x, int ld, int B, int C, int T, float* out, cudaStream_t st) {
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B), block(32, 8);
  tokens_to_nct_kernel<<<grid, block, 0, st>>>(x, ld, C, T, out);
  NS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm row statistics (one warp per row; two-pass in registers, C <= 2048).
// ---------------------------------------------------------------------------------------------
template <bool APPLY>
__global__ void __launch_bounds__(256) ln_kernel(const float* __restrict__ x, int ld, int M, int C, float eps,
                                                 float* __restrict__ stats, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float* __restrict__ y, int y_ld) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long long)row * ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { float d = xr[c] - mean; q += d * d; }
  const float var = warp_sum(q) / (float)C;
  const float rstd = 1.0f / sqrtf(var + eps);
  if (APPLY) {
    float* yr = y + (long long)row * y_ld;
    for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
  } else if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
}
int launch_ln_stats(const float* x, int ld, int M, int C, float eps, float* stats, cudaStream_t st) {
  ln_kernel<false><<<ceil_div(M, 8), 256, 0, st>>>(x, ld, M, C, eps, stats, nullptr, nullptr, nullptr, 0);
  NS_LAUNCH_CHECK();
  return 0;
}
int launch_ln_apply(const float* x, int ld, int M, int C, float eps, const float* gamma, const float* beta, float* y,
                    int y_ld, cudaStream_t st) {
  ln_kernel<true><<<ceil_div(M, 8), 256, 0, st>>>(x, ld, M, C, eps, nullptr, gamma, beta, y, y_ld);
  NS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Activation prep: concat(src1, src2) -> [per-(b,c) affine (+SiLU)] -> bf16 hi/lo split, optionally a
// second untransformed split (the resnet's 1x1 shortcut operand).  One thread = 8 channels of one
// row (two 16-byte loads, 16-byte hi + lo stores); rows may be remapped (stride-2 decimation for
// the downsample convs, nearest-upsample index table).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_split_kernel(PrepOp op) {
  span_begin(op.span);
  pdl_trigger();
  extern __shared__ float aff[];                         // [2][C] scale | shift of this block's batch entry
  const int C = op.C1 + op.C2;
  float pg[kPrepSlots], pb[kPrepSlots];
  prep_fetch_norm_weights(op, C, pg, pb, threadIdx.x, blockDim.x);   // weights: before griddepcontrol.wait
  pdl_wait();
  const int b = blockIdx.y;
  const int chunks = op.out.ld >> 3;                     // 8-channel chunks per output row (incl. zero padding)
  const int total = op.T_dst * chunks;
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  // the first chunk's loads are in flight while the block derives the GroupNorm affine
  PrepChunk k0;
  const bool have = i < total;
  if (have) prep_load_at(op, b, C, i / chunks, i % chunks, k0);
  float fs[kPrepSlots], fb[kPrepSlots];
  prep_fetch_film(op, op.gn.film, b, C, fs, fb, threadIdx.x, blockDim.x);
  prep_affine(op, b, C, C, aff, pg, pb, fs, fb, threadIdx.x, blockDim.x, BlockSync());
  if (have) prep_finish(op, b, C, aff, k0);
  for (i += stride; i < total; i += stride) {
    PrepChunk k;
    prep_load_at(op, b, C, i / chunks, i % chunks, k);
    prep_finish(op, b, C, aff, k);
  }
  span_end(op.span);
}
int launch_prep_split(const PrepOp& op, cudaStream_t st) {
  if ((op.out.ld & 7) || (op.raw.hi && op.raw.ld != op.out.ld)) { set_error("prep_split: bad pitch"); return -1; }
  const int C = op.C1 + op.C2;
  if (op.mode != PREP_RAW && !op.scale && (C % op.gn.G)) { set_error("prep_split: %d channels not divisible by %d groups", C, op.gn.G); return -1; }
  const int total = op.T_dst * (op.out.ld >> 3);
  int bx = (total + 255) / 256;
  const int cap = (148 * 8 + op.B - 1) / op.B;           // ~8 blocks per SM over the whole grid
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const size_t smem = (op.mode != PREP_RAW) ? (size_t)prep_affine_floats(C) * sizeof(float) : 0;
  if (smem > 48 * 1024 || (op.mode != PREP_RAW && !op.scale && (C > kPrepSlots * 256 || op.gn.G > 64))) { set_error("prep_split: C=%d too large", C); return -1; }
  cudaError_t e = launch_k(prep_split_kernel, dim3(bx, op.B), dim3(256), smem, st, op);
  if (e != cudaSuccess) { set_error("prep_split launch failed: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

// LayerNorm + split: one warp per row, two-pass statistics in registers (C <= 1024).
__global__ void __launch_bounds__(256) ln_split_kernel(const float* __restrict__ x, int ld, int M, int C, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       SplitBuf out, unsigned long long* span) {
  span_begin(span);
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  constexpr int kMaxChunks = 4;                            // 8-channel chunks per lane: C <= 32*8*4
  // gamma / beta are weights: fetch them while the producer of x is still running (before griddepcontrol.wait)
  float gm[kMaxChunks][8], bt[kMaxChunks][8];
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int c0 = (lane + 32 * k) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { gm[k][j] = 0.f; bt[k][j] = 0.f; }
    if (c0 + 8 <= C && ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
      gm[k][0] = g0.x; gm[k][1] = g0.y; gm[k][2] = g0.z; gm[k][3] = g0.w; gm[k][4] = g1.x; gm[k][5] = g1.y; gm[k][6] = g1.z; gm[k][7] = g1.w;
      bt[k][0] = b0.x; bt[k][1] = b0.y; bt[k][2] = b0.z; bt[k][3] = b0.w; bt[k][4] = b1.x; bt[k][5] = b1.y; bt[k][6] = b1.z; bt[k][7] = b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c0 + j < C) { gm[k][j] = __ldg(gamma + c0 + j); bt[k][j] = __ldg(beta + c0 + j); }
    }
  }
  pdl_wait();
  if (row >= M) { span_end(span); return; }
  const float* xr = x + (long long)row * ld;
  float v[kMaxChunks][8];
  const int chunks = (C + 7) >> 3;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ck = lane + 32 * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
    if (ck < chunks) {
      const int c0 = ck * 8;
      if (c0 + 8 <= C && ((ld & 3) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xr + c0)), b4 = __ldg(reinterpret_cast<const float4*>(xr + c0) + 1);
        v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w; v[k][4] = b4.x; v[k][5] = b4.y; v[k][6] = b4.z; v[k][7] = b4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (c0 + j < C) v[k][j] = xr[c0 + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[k][j];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ck = lane + 32 * k;
    if (ck < chunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (ck * 8 + j < C) { const float d = v[k][j] - mean; q += d * d; }
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
  const int ochunks = out.ld >> 3;
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ck = lane + 32 * k;
    if (ck < ochunks) {
      const int c0 = ck * 8;
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        y[j] = (c < C) ? (v[k][j] - mean) * rstd * gm[k][j] + bt[k][j] : 0.f;
      }
      uint4 hi, lo;
      split8(y, hi, lo);
      *reinterpret_cast<uint4*>(out.hi + (long long)row * out.ld + c0) = hi;
      *reinterpret_cast<uint4*>(out.lo + (long long)row * out.ld + c0) = lo;
    }
  }
  span_end(span);
}
int launch_ln_split(const float* x, int ld, int M, int C, float eps, const float* gamma, const float* beta, SplitBuf out,
                    cudaStream_t st, unsigned long long* span) {
  if (C > 1024 || (out.ld & 7) || out.ld > 1024) { set_error("ln_split: C=%d / pitch %d unsupported", C, out.ld); return -1; }
  launch_k(ln_split_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, st, x, ld, M, C, eps, gamma, beta, out, span);
  NS_LAUNCH_CHECK();
  return 0;
}

// [B, C, T] fp32 -> split token-major [B, T, out.ld]; 32x32 smem transpose, zero-fills c >= C.
__global__ void nct_to_split_kernel(const float* __restrict__ x, long long bstride, int C, int T, SplitBuf out,
                                    const char* warm, long long warm_bytes) {
  pdl_trigger();
  // first kernel of a forward: pull this step's FiLM rows (a slice of the run's timestep table, cold in L2) towards L2 so that
  // the 22 conv2 launches that read them later do not each wait for HBM
  if (warm) {
    const long long line = ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (blockDim.x * blockDim.y) + threadIdx.y * blockDim.x + threadIdx.x;
    if (line * 128 < warm_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(warm + line * 128));
  }
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xb = x + (long long)b * bstride;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? xb[(long long)c * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < out.ld) {
      const float v = tile[threadIdx.x][i];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const long long off = ((long long)b * T + t) * out.ld + c;
      out.hi[off] = h;
      out.lo[off] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}
int launch_nct_to_split(const float* x, long long bstride, int B, int C, int T, SplitBuf out, cudaStream_t st, const void* warm,
                        long long warm_bytes) {
  dim3 grid(ceil_div(T, 32), ceil_div(out.ld, 32), B), block(32, 8);
  launch_k(nct_to_split_kernel, grid, block, 0, st, x, bstride, C, T, out, (const char*)warm, warm_bytes);
  NS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Small-M linear (timestep MLP, batched FiLM projections, AttentionPooling projections).
// One warp per output column n; the input rows (<= 8 at a time) live in shared memory.
// HBM-bound on W (read once per launch when M <= 8).
// ---------------------------------------------------------------------------------------------
constexpr int kLinRows = 8;
__global__ void __launch_bounds__(256) small_linear_kernel(LinOp op) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float xs[];   // [kLinRows][K]
  const int m0 = blockIdx.y * kLinRows;
  const int rows = min(kLinRows, op.M - m0);
  const int K = op.K;
  const bool vec_in = op.in_mode != LIN_SINUSOID && op.x_ld == K && (K & 3) == 0 && (reinterpret_cast<uintptr_t>(op.x) & 15) == 0;
  if (vec_in) {
    // contiguous input rows: all 16-byte loads of this thread are in flight before the first use
    const float4* xin = reinterpret_cast<const float4*>(op.x + (long long)m0 * K);
    const int n4 = rows * K / 4;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; v[u] = (i < n4) ? __ldg(xin + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n4) {
          if (op.in_mode == LIN_SILU) {
            v[u].x = v[u].x / (1.0f + expf(-v[u].x)); v[u].y = v[u].y / (1.0f + expf(-v[u].y));
            v[u].z = v[u].z / (1.0f + expf(-v[u].z)); v[u].w = v[u].w / (1.0f + expf(-v[u].w));
          }
          reinterpret_cast<float4*>(xs)[i] = v[u];
        }
      }
    }
  }
  for (int i = threadIdx.x; !vec_in && i < rows * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    float v;
    if (op.in_mode == LIN_SINUSOID) {
      // reference embeddings.py:41-59: emb = t * exp(-ln(1e4) * i / (half - shift)); [sin | cos], flipped
      const int half = K / 2;
      const float t = op.x[(long long)(m0 + r) * op.x_ld];
      if (k >= 2 * half) {
        v = 0.f;
      } else {
        bool first = k < half;
        int i2 = first ? k : k - half;
        float ex = (-9.210340371976184f * (float)i2) / ((float)half - op.freq_shift);
        float arg = t * expf(ex);
        bool use_cos = op.flip_sin_to_cos ? first : !first;
        v = use_cos ? cosf(arg) : sinf(arg);
      }
    } else {
      v = op.x[(long long)(m0 + r) * op.x_ld + k];
      if (op.in_mode == LIN_SILU) v = v / (1.0f + expf(-v));
    }
    xs[r * K + k] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= op.N) return;
  const float* wr = op.W + (long long)n * K;
  float acc[kLinRows];
#pragma unroll
  for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
  if ((K & 127) == 0 && (reinterpret_cast<uintptr_t>(wr) & 15) == 0) {
    // the weight row is the only HBM traffic: issue up to four 16-byte loads per lane before the first FMA so a warp
    // pays one memory latency per 512 weights instead of one per 32
    for (int k0 = 0; k0 < K; k0 += 512) {
      float4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * 128 + lane * 4;
        w[u] = (k < K) ? __ldg(reinterpret_cast<const float4*>(wr + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * 128 + lane * 4;
        if (k < K) {
#pragma unroll
          for (int r = 0; r < kLinRows; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + r * K + k);
            acc[r] = fmaf(xv.x, w[u].x, acc[r]); acc[r] = fmaf(xv.y, w[u].y, acc[r]);
            acc[r] = fmaf(xv.z, w[u].z, acc[r]); acc[r] = fmaf(xv.w, w[u].w, acc[r]);
          }
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float w = __ldg(wr + k);
#pragma unroll
      for (int r = 0; r < kLinRows; ++r) acc[r] = fmaf(xs[r * K + k], w, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kLinRows; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float bv = op.bias ? op.bias[n] : 0.f;
    for (int r = 0; r < rows; ++r) {
      float v = acc[r] + bv;
      if (op.add) v += op.add[(long long)(op.add_rows > 0 ? (m0 + r) % op.add_rows : (m0 + r)) * op.add_ld + n];
      if (op.out_silu) v = v / (1.0f + expf(-v));
      op.out[(long long)(m0 + r) * op.out_ld + n] = v;
    }
  }
}
int launch_small_linear(const LinOp& op, cudaStream_t st) {
  size_t smem = (size_t)kLinRows * op.K * sizeof(float);
  if (smem > 48 * 1024) { set_error("small_linear: K=%d too large", op.K); return -1; }
  dim3 grid(ceil_div(op.N, 8), ceil_div(op.M, kLinRows));
  launch_k(small_linear_kernel, grid, dim3(256), smem, st, op);
  NS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// AttentionPooling pieces (reference embeddings.py:499-546); once per utterance.
// ---------------------------------------------------------------------------------------------
__global__ void pool_class_token_kernel(const float* __restrict__ xn, const float* __restrict__ pos, int S, int C,
                                        float* __restrict__ tokens) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < S; ++t) s += xn[((long long)b * S + t) * C + c];
    tokens[((long long)b * (S + 1)) * C + c] = s / (float)S + pos[c];
    for (int t = 0; t < S; ++t) tokens[((long long)b * (S + 1) + 1 + t) * C + c] = xn[((long long)b * S + t) * C + c];
  }
}
int launch_pool_class_token(const float* xn, const float* pos, int B, int S, int C, float* tokens, cudaStream_t st) {
  pool_class_token_kernel<<<B, 256, 0, st>>>(xn, pos, S, C, tokens);
  NS_LAUNCH_CHECK();
  return 0;
}

// one warp per (b, head): softmax over S1 keys of (q*s).(k*s), s = dph^-1/4; out = sum_j w_j v_j
__global__ void pool_attend_kernel(const float* __restrict__ q, const float* __restrict__ kv, int S1, int C, int heads,
                                   float* __restrict__ out) {
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dph = C / heads;
  const int lane = threadIdx.x;
  const float sc = 1.0f / sqrtf(sqrtf((float)dph));
  const float* qh = q + (long long)b * C + h * dph;
  float mx = -INFINITY;
  for (int j = lane; j < S1; j += 32) {
    const float* kr = kv + ((long long)b * S1 + j) * 2 * C + h * dph;
    float s = 0.f;
    for (int d = 0; d < dph; ++d) s += (qh[d] * sc) * (kr[d] * sc);
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float den = 0.f;
  float acc[16];
  for (int d = 0; d < 16; ++d) acc[d] = 0.f;
  for (int j = lane; j < S1; j += 32) {
    const float* kr = kv + ((long long)b * S1 + j) * 2 * C + h * dph;
    float s = 0.f;
    for (int d = 0; d < dph; ++d) s += (qh[d] * sc) * (kr[d] * sc);
    float p = expf(s - mx);
    den += p;
    const float* vr = kr + C;
    for (int d = 0; d < dph && d < 16; ++d) acc[d] += p * vr[d];
  }
  den = warp_sum(den);
  for (int d = 0; d < dph && d < 16; ++d) {
    float a = warp_sum(acc[d]);
    if (lane == 0) out[(long long)b * C + h * dph + d] = a / den;
  }
}
int launch_pool_attend(const float* q, const float* kv, int B, int S1, int C, int heads, float* out, cudaStream_t st) {
  if (C % heads || C / heads > 16) { set_error("pool_attend: dim/head %d/%d unsupported", C, heads); return -1; }
  pool_attend_kernel<<<B * heads, 32, 0, st>>>(q, kv, S1, C, heads, out);
  NS_LAUNCH_CHECK();
  return 0;
}

// bool mask -> additive bias, bit-exact with (1 - m) * -10000 (reference unet_1d_condition.py:817)
__global__ void mask_bias_kernel(const uint8_t* __restrict__ mask, int n, float* __restrict__ bias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bias[i] = (1.0f - (mask[i] ? 1.0f : 0.0f)) * -10000.0f;
}
int launch_mask_bias(const uint8_t* mask, int n, float* bias, cudaStream_t st) {
  mask_bias_kernel<<<ceil_div(n, 256), 256, 0, st>>>(mask, n, bias);
  NS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused sampler steps.  Every arithmetic op uses the round-to-nearest intrinsics so that the
// compiler cannot contract a*b+c into an FMA: the reference evaluates each product and sum as a
// separate fp32 tensor op (dpm_solver.py:291-292, 437-439, 569-576, 813-831), and the result
// here is bit-identical to that sequence.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float x0_round_trip(float x, float o, float alpha, float sigma) {
  // noise = (x - alpha*out)/sigma  (model_wrapper, x_start)  ;  x0 = (x - sigma*noise)/alpha
  float noise = __fdiv_rn(__fsub_rn(x, __fmul_rn(alpha, o)), sigma);
  return __fdiv_rn(__fsub_rn(x, __fmul_rn(sigma, noise)), alpha);
}

__global__ void __launch_bounds__(256) dpm_step_kernel(const float* __restrict__ x, const float* __restrict__ o,
                                                       const float* __restrict__ mp, DpmStepCoef c,
                                                       float* __restrict__ mc, float* __restrict__ xn, size_t n, int* nan_flag) {
  pdl_trigger();
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < n; i += stride) {
    const float xv = x[i];
    bad |= (xv != xv);                                     // the reference asserts on NaN in the denoiser input every call (model.py:404)
    const float m0 = x0_round_trip(xv, o[i], c.alpha_s, c.sigma_s);
    mc[i] = m0;
    if (c.order == 0) continue;
    float r = __fsub_rn(__fmul_rn(c.c_x, xv), __fmul_rn(c.c_m, m0));
    if (c.order == 2) {
      float d1 = __fmul_rn(c.inv_r0, __fsub_rn(m0, mp[i]));
      r = __fsub_rn(r, __fmul_rn(c.c_d, d1));
    }
    xn[i] = r;
  }
  if (bad && nan_flag) atomicOr(nan_flag, 1);
}
int launch_dpm_step(const float* x, const float* unet_out, const float* m_prev, const DpmStepCoef& c, float* m_cur,
                    float* x_next, size_t n, int* nan_flag, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(dpm_step_kernel, dim3(blocks), dim3(256), 0, st, x, unet_out, m_prev, c, m_cur, x_next, n, nan_flag);
  NS_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) unipc_step_kernel(const float* __restrict__ xp, const float* __restrict__ xe,
                                                         const float* __restrict__ o, const float* __restrict__ m0p,
                                                         const float* __restrict__ m1p, UniPcStepCoef c,
                                                         float* __restrict__ mt_out, float* __restrict__ xt_out,
                                                         float* __restrict__ xpred_out, size_t n, int* nan_flag) {
  pdl_trigger();
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < n; i += stride) {
    const float xev = xe[i];
    bad |= (xev != xev);                                   // NaN guard of the denoiser input (model.py:404)
    const float mt = x0_round_trip(xev, o[i], c.alpha_t, c.sigma_t);
    mt_out[i] = mt;
    float xt = xev;
    float m0 = 0.f;
    if (c.corr_order > 0) {
      m0 = m0p[i];
      // uni_pc.py:533-536, 561-567
      const float xbar = __fsub_rn(__fmul_rn(c.c_x, xp[i]), __fmul_rn(c.c_m, m0));
      const float d1t = __fsub_rn(mt, m0);
      float inner;
      if (c.corr_order == 2) {
        const float d1 = __fdiv_rn(__fsub_rn(m1p[i], m0), c.rk);
        inner = __fadd_rn(__fmul_rn(c.rho0, d1), __fmul_rn(c.rho1, d1t));
      } else {
        inner = __fmul_rn(c.rho1, d1t);     // 0 + 0.5*D1_t
      }
      xt = __fsub_rn(xbar, __fmul_rn(c.ab, inner));
      xt_out[i] = xt;
    }
    if (c.pred_order > 0) {
      const float nbar = __fsub_rn(__fmul_rn(c.n_c_x, xt), __fmul_rn(c.n_c_m, mt));
      float xpred = nbar;
      if (c.pred_order == 2) {
        const float d1n = __fdiv_rn(__fsub_rn(m0, mt), c.nrk);
        xpred = __fsub_rn(nbar, __fmul_rn(c.nab, __fmul_rn(0.5f, d1n)));
      }
      xpred_out[i] = xpred;
    }
  }
  if (bad && nan_flag) atomicOr(nan_flag, 1);
}
int launch_unipc_step(const float* x_prev, const float* x_eval, const float* unet_out, const float* m0,
                      const float* m1, const UniPcStepCoef& c, float* m_t, float* x_t, float* x_pred, size_t n,
                      int* nan_flag, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(unipc_step_kernel, dim3(blocks), dim3(256), 0, st, x_prev, x_eval, unet_out, m0, m1, c, m_t, x_t, x_pred, n, nan_flag);
  NS_LAUNCH_CHECK();
  return 0;
}

}  // namespace ns2vc
