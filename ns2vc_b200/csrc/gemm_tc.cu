// tcgen05 implicit-GEMM kernel for every contraction of the denoiser step (k=3 convs, 1x1 convs,
// linears), sm_100a.
//
//   D[128 x BN] (fp32, TMEM) += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (3xBF16 split)
//
// fp32-level parity with the reference (rtol 1e-3 / atol 1e-4) cannot be met by single-pass
// bf16/tf32 MMAs, so both operands are split x = hi + lo (bf16 each) and
// three kind::f16 MMAs accumulate into the same TMEM tile (the lo*lo term, ~2^-16 relative, is
// dropped).
//
// Warp roles (320 threads, 1 CTA/SM, 3-stage mbarrier ring of 64 KB stages):
//   warps 0-7  A producers: read the fp32 activations (token-major rows, shifted per conv tap,
//              strided / index-mapped for down/up-sampling, two base pointers for channel
//              concats), apply the fused GroupNorm(+FiLM)+SiLU or LayerNorm transform, split to
//              bf16 hi/lo and write the K-major SWIZZLE_128B shared-memory image of the UMMA A
//              operand.  After the main loop the same warps run the epilogue:
//              TMEM -> registers (tcgen05.ld 32x32b) -> bias / GEGLU / residual -> global.
//   warp 8     B producer: one cp.async.bulk (UBLKCP) per hi/lo tile from the pre-swizzled packed
//              weights; also owns the TMEM allocation.
//   warp 9     MMA issuer: one elected lane issues tcgen05.mma and tcgen05.commit.
#include "gemm_common.cuh"
#include <cstdio>

namespace ns2vc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kProdWarps = 8;
constexpr int kThreads = (kProdWarps + 2) * 32;
constexpr int kATileBytes = BM * BK * 2;                // 16 KB: one bf16 [128 x 64] A tile (hi or lo)

template <int BN_> struct TileCfg {
  static constexpr int BN = BN_;
  static constexpr int kBTileBytes = BN_ * BK * 2;      // one bf16 [BN x 64] B tile (hi or lo)
  static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
  static constexpr int kStages = (BN_ == 128) ? 3 : 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
  static constexpr uint32_t kTmemCols = BN_;
  // Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1,
  // B=bf16 [10,13)=1, A/B K-major, N>>3 at [17,23), M>>4 at [24,29).
  static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("ns2vc gemm_tc: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major; 1) | [32,46) SBO>>4 = 1024>>4
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---- A-producer helpers.  One thread owns 8 consecutive channels (one 16-byte bf16 chunk) of 4 rows.
// Must agree element-for-element with a_fetch() in gemm_common.cuh.
struct Row8 { float v[8]; };

__device__ __forceinline__ void load_raw8(const ASeg& s, long long srow, int c0, bool vec, Row8& r) {
  const float* px = s.src + srow * s.ld + s.ch0 + c0;
  if (vec) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(px));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(px) + 1);
    r.v[0] = x0.x; r.v[1] = x0.y; r.v[2] = x0.z; r.v[3] = x0.w; r.v[4] = x1.x; r.v[5] = x1.y; r.v[6] = x1.z; r.v[7] = x1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = (c0 + j < s.nch) ? __ldg(px + j) : 0.f;
  }
}
__device__ __forceinline__ void load_vec8(const float* p, int c0, int nch, bool vec, float* o) {
  if (vec) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p)), a1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a0.x; o[1] = a0.y; o[2] = a0.z; o[3] = a0.w; o[4] = a1.x; o[5] = a1.y; o[6] = a1.z; o[7] = a1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (c0 + j < nch) ? __ldg(p + j) : 0.f;
  }
}
__device__ __forceinline__ void store_split8(uint8_t* a_hi, uint8_t* a_lo, int r, int chunk, const float* v) {
  float h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
  uint4 hi, lo;
  hi.x = pack_bf16x2(h[0], h[1]); hi.y = pack_bf16x2(h[2], h[3]);
  hi.z = pack_bf16x2(h[4], h[5]); hi.w = pack_bf16x2(h[6], h[7]);
  lo.x = pack_bf16x2(v[0] - h[0], v[1] - h[1]); lo.y = pack_bf16x2(v[2] - h[2], v[3] - h[3]);
  lo.z = pack_bf16x2(v[4] - h[4], v[5] - h[5]); lo.w = pack_bf16x2(v[6] - h[6], v[7] - h[7]);
  const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(a_hi + off) = hi;
  *reinterpret_cast<uint4*>(a_lo + off) = lo;
}

template <int BN_>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmOp op) {
  using Cfg = TileCfg<BN_>;
  constexpr int BN = Cfg::BN;
  constexpr int kStages = Cfg::kStages;
  constexpr int kStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;             // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kStages * kStageBytes + 8 * (2 * kStages + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = op.B * op.T_out;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;                           // first packed column of this tile
  const int nkb = op.nkb_total;

  if (warp == 9 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), kProdWarps + 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProdWarps) {
    // ===================== A producers =====================
    const int chunk = tid & 7;                              // 8 channels = one 16-byte bf16 chunk
    const int rsub = tid >> 3;                              // 0..31
    int rb[4], rt[4];
    bool rv[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int m = m0 + p * 32 + rsub;
      rv[p] = m < M;
      rb[p] = rv[p] ? m / op.T_out : 0;
      rt[p] = rv[p] ? m % op.T_out : 0;
    }
    const bool same_b = (rb[0] == rb[1]) && (rb[0] == rb[2]) && (rb[0] == rb[3]);
    int si = 0, kbl = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int stage = kb % kStages;
      const uint32_t parity = (uint32_t)((kb / kStages) & 1);
      const ASeg& s = op.seg[si];
      const int c0 = kbl * 64 + chunk * 8;
      const bool chan_ok = c0 < s.nch;
      const bool full = c0 + 8 <= s.nch;
      const bool vec = full && (((s.ld | s.ch0) & 3) == 0);
      // ---- phase 1: issue every global load of this k-block before touching the data
      long long srow[4];
      Row8 x[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        srow[p] = (rv[p] && chan_ok) ? a_src_row(op, rb[p], rt[p], s.tap) : -1;
        if (srow[p] >= 0) load_raw8(s, srow[p], c0, vec, x[p]);
      }
      float pa[8], pb[8];                                   // affine scale/shift or LN gamma/beta
      float2 st[4];
      const int mode = s.mode;
      if (chan_ok) {
        if (mode == A_AFFINE || mode == A_AFFINE_SILU) {
          const bool pvec = full && (((s.ald | s.aoff) & 3) == 0);
          load_vec8(s.p0 + (long long)rb[0] * s.ald + s.aoff + c0, c0, s.nch, pvec, pa);
          load_vec8(s.p1 + (long long)rb[0] * s.ald + s.aoff + c0, c0, s.nch, pvec, pb);
        } else if (mode == A_LN) {
          const bool pvec = full && ((s.ch0 & 3) == 0);
          load_vec8(s.p1 + s.ch0 + c0, c0, s.nch, pvec, pa);
          load_vec8(s.p2 + s.ch0 + c0, c0, s.nch, pvec, pb);
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (srow[p] >= 0) st[p] = __ldg(reinterpret_cast<const float2*>(s.p0) + srow[p]);
        }
      }
      mbar_wait(empty_bar(stage), parity ^ 1u);
      uint8_t* a_hi = smem + stage * kStageBytes;
      uint8_t* a_lo = a_hi + kATileBytes;
      // ---- phase 2: transform, split, store
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int r = p * 32 + rsub;
        float v[8];
        if (srow[p] < 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = x[p].v[j];
          if (mode == A_AFFINE || mode == A_AFFINE_SILU) {
            float sc[8], sh[8];
            if (same_b || rb[p] == rb[0]) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { sc[j] = pa[j]; sh[j] = pb[j]; }
            } else {
              const bool pvec = full && (((s.ald | s.aoff) & 3) == 0);
              load_vec8(s.p0 + (long long)rb[p] * s.ald + s.aoff + c0, c0, s.nch, pvec, sc);
              load_vec8(s.p1 + (long long)rb[p] * s.ald + s.aoff + c0, c0, s.nch, pvec, sh);
            }
            if (mode == A_AFFINE_SILU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = silu_f(fmaf(v[j], sc[j], sh[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
            }
          } else if (mode == A_LN) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (v[j] - st[p].x) * st[p].y * pa[j] + pb[j];
          }
          if (!full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (c0 + j >= s.nch) v[j] = 0.f;
          }
        }
        store_split8(a_hi, a_lo, r, chunk, v);
      }
      fence_proxy_async();                                  // generic-proxy stores -> visible to UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      if (++kbl == s.nkb) { kbl = 0; ++si; }
    }

    // ===================== epilogue =====================
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                                 // TMEM lane quarter this warp may access
    const int hh = warp >> 2;                               // column half
    const int r = q * 32 + lane;
    const long long m = (long long)m0 + r;
    const bool mv = m < M;
    const int b = mv ? (int)(m / op.T_out) : 0;
    const int t = mv ? (int)(m % op.T_out) : 0;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if (op.flags & EPI_GEGLU) {
      if (BN == 128) {
        float val[32], gate[32];
        tmem_ld32(trow + (uint32_t)(hh * 32), val);
        tmem_ld32(trow + (uint32_t)(64 + hh * 32), gate);
        const int nbase = blockIdx.y * 64 + hh * 32;        // logical output column
        if (mv) {
          float* po = op.out + m * op.out_ld + nbase;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = epi_value(op, b, m, nbase + j + 0, val[j + 0], gate[j + 0]);
            o.y = epi_value(op, b, m, nbase + j + 1, val[j + 1], gate[j + 1]);
            o.z = epi_value(op, b, m, nbase + j + 2, val[j + 2], gate[j + 2]);
            o.w = epi_value(op, b, m, nbase + j + 3, val[j + 3], gate[j + 3]);
            *reinterpret_cast<float4*>(po + j) = o;
          }
        }
      }
    } else {
      constexpr int kChunks = BN / 64;                      // 32-column chunks per warp
#pragma unroll 1
      for (int cc = 0; cc < kChunks; ++cc) {
        float acc[32];
        const int col = hh * (BN / 2) + cc * 32;
        tmem_ld32(trow + (uint32_t)col, acc);
        const int nbase = n0 + col;
        if (!mv || nbase >= op.n_valid) continue;
        if (op.flags & EPI_OUT_NCT) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nbase + j;
            if (n < op.n_valid) op.out[((long long)b * op.n_valid + n) * op.T_out + t] = epi_value(op, b, m, n, acc[j], 0.f);
          }
        } else if (nbase + 32 <= op.n_valid && ((op.out_ld & 3) == 0) && (!(op.flags & EPI_RESIDUAL) || (op.res_ld & 3) == 0) &&
                   !(op.flags & EPI_ROWBIAS)) {
          float* po = op.out + m * op.out_ld + nbase;
          const float4* pr = (op.flags & EPI_RESIDUAL) ? reinterpret_cast<const float4*>(op.res + m * op.res_ld + nbase) : nullptr;
          const float4* pbias = (op.flags & EPI_BIAS) ? reinterpret_cast<const float4*>(op.bias + nbase) : nullptr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            if (pbias) { const float4 bv = __ldg(pbias + j); o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w; }
            if (pr) { const float4 rv4 = __ldg(pr + j); o.x += rv4.x; o.y += rv4.y; o.z += rv4.z; o.w += rv4.w; }
            *reinterpret_cast<float4*>(po + 4 * j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nbase + j;
            if (n < op.n_valid) op.out[m * op.out_ld + n] = epi_value(op, b, m, n, acc[j], 0.f);
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===================== B producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % kStages;
        const uint32_t parity = (uint32_t)((kb / kStages) & 1);
        mbar_wait(empty_bar(stage), parity ^ 1u);
        const uint32_t b_hi = base + stage * kStageBytes + 2 * kATileBytes;
        const uint32_t b_lo = b_hi + Cfg::kBTileBytes;
        mbar_arrive_expect_tx(full_bar(stage), 2u * Cfg::kBTileBytes);
        const size_t eoff = ((size_t)kb * op.N + n0) * 64;
        bulk_g2s(b_hi, op.w_hi + eoff, Cfg::kBTileBytes, full_bar(stage));
        bulk_g2s(b_lo, op.w_lo + eoff, Cfg::kBTileBytes, full_bar(stage));
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % kStages;
        const uint32_t parity = (uint32_t)((kb / kStages) & 1);
        mbar_wait(full_bar(stage), parity);
        tc_fence_after();
        const uint32_t a_hi = base + stage * kStageBytes;
        const uint32_t a_lo = a_hi + kATileBytes;
        const uint32_t b_hi = a_hi + 2 * kATileBytes;
        const uint32_t b_lo = b_hi + Cfg::kBTileBytes;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t dah = umma_desc(a_hi + k * 32), dal = umma_desc(a_lo + k * 32);
          const uint64_t dbh = umma_desc(b_hi + k * 32), dbl = umma_desc(b_lo + k * 32);
          umma_bf16(tmem_base, dah, dbh, Cfg::kIdesc, (kb | k) != 0 ? 1u : 0u);
          umma_bf16(tmem_base, dah, dbl, Cfg::kIdesc, 1u);
          umma_bf16(tmem_base, dal, dbh, Cfg::kIdesc, 1u);
        }
        umma_commit(empty_bar(stage));                      // frees this smem stage when the MMAs retire
      }
      umma_commit(tmem_full_bar);                           // accumulator complete -> epilogue
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

template <int BN_>
static int launch_bn(const GemmOp& op, cudaStream_t st) {
  using Cfg = TileCfg<BN_>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc: cannot set %d B dynamic smem: %s", Cfg::kSmemBytes, cudaGetErrorString(e)); return -2; }
    attr_set = true;
  }
  const int M = op.B * op.T_out;
  dim3 grid(ceil_div(M, BM), op.N / BN_);
  gemm_tc_kernel<BN_><<<grid, kThreads, Cfg::kSmemBytes, st>>>(op);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("gemm_tc launch failed: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

int launch_gemm_tc(const GemmOp& op, cudaStream_t st) {
  if (op.N % 128) { set_error("gemm_tc: packed N=%d is not a multiple of 128", op.N); return -1; }
  if (op.nkb_total <= 0) { set_error("gemm_tc: empty K"); return -1; }
  const int M = op.B * op.T_out;
  // N-tile: 128 wide when that already fills the 148 SMs, else 64 wide (twice the CTAs; the
  // GEGLU epilogue pairs value|gate inside a 128-column block and needs BN = 128).
  const int ctas128 = ceil_div(M, BM) * (op.N / 128);
  if ((op.flags & EPI_GEGLU) || ctas128 >= 120) return launch_bn<128>(op, st);
  return launch_bn<64>(op, st);
}

}  // namespace ns2vc
