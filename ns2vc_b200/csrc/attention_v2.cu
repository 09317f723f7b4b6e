// tcgen05 flash attention, v2 (product path): TMA-fed and warp-specialised.
//   out = softmax(q k^T * dh^-0.5 + bias) v      (reference attention_processor.py:1025-1036)
//
// CTA = 128 queries of one (batch, head); key tiles of 64; 288 threads:
//   warp 0      one elected thread (elect.sync) is TMA producer and MMA issuer at once.  Producer: Q once, then a ring of
//               {K_hi, K_lo, V_hi, V_lo} tiles.  q / k / v are the bf16 hi/lo "split" tensors the projection GEMMs'
//               epilogues wrote (token-major), so a tile is a plain 3-D box {head channels, 64 keys, 1 batch} - no
//               conversion, no transpose.  MMA issuer: S[128x64] = Q K^T (A, B K-major) and O_tile[128xdh] = P V with V
//               as an MN-major B operand (dh contiguous, keys = MMA K dimension), each as 3 bf16 MMAs over the hi/lo
//               splits, fp32 accumulators in TMEM.  x_hi*[y_hi ; y_lo] is one instruction of twice the N (two TMEM column
//               groups), x_lo*y_hi a second one: 2 MMAs per k-step instead of 3.  The ring stage of tile j-1 is refilled
//               right before PV(j) is issued (PV(j-1) has retired by then: the softmax warps waited for it).
//   warps 1-8   online softmax, two threads per query row (32 score columns and dh/2 output columns
//               each): exp2 with scale*log2(e) folded into one FFMA, P written as a SWIZZLE_128B A operand.
// The roles only meet through mbarriers: S(j+1) is issued as soon as the softmax warps have
// pulled S(j) out of TMEM, and P(j) V(j) runs under softmax(j+1), so neither MMA latency nor a
// CTA-wide barrier sits on the per-tile critical path.
// Shared-memory rows of the Q/K/V tiles are `PB` bytes wide (32/64/128 = the TMA box width and the
// UMMA swizzle mode), so a dh=16 head moves 32 B per key instead of a padded 128 B row.
#include "gemm_common.cuh"
#include "tc_common.cuh"
#include "launch.cuh"
#include <math.h>
#include <cstdlib>

namespace ns2vc {

namespace {

constexpr int kQ = 128, kKeys = 64;
constexpr int kThreadsV2 = 288;                          // warp 0: TMA producer + MMA issuer (one elected thread); warps 1-8: softmax

template <int DHP, int PB, bool PF16, bool BIAS> struct ACfg {
  static constexpr int NST = (PB == 128) ? 2 : 3;
  static constexpr int kQBytes = kQ * PB;            // Q hi (lo follows)
  static constexpr int kTBytes = kKeys * PB;         // one K or V tile (hi or lo)
  static constexpr int kStageBytes = 4 * kTBytes;    // K_hi | K_lo | V_hi | V_lo
  static constexpr int kPBytes = kQ * 128;           // P [128 x 64] 16-bit: fp16 (PF16), or bf16 hi with lo following
  static constexpr int kOffQ = 0;
  static constexpr int kOffP = 2 * kQBytes;
  static constexpr int kOffKV = kOffP + (PF16 ? 1 : 2) * kPBytes;
  static constexpr int kOffBar = kOffKV + NST * kStageBytes;
  static constexpr int kOffXch = kOffBar + 256;      // [3: parity 0 / parity 1 / row sums][2 halves][128 rows] floats
  static constexpr int kOffBias = kOffXch + 3072;    // additive bias * log2(e) (or -inf past Tk) for up to kBiasKeys keys
  static constexpr int kBiasKeys = (PB == 32) ? 768 : 1024;   // d_h = 16: 3 KB so that the biased (cross-attention) variant also fits 4 CTAs / SM
  static constexpr int kSmem = kOffBias + (BIAS ? 4 * kBiasKeys : 0) + 1024 /*alignment slack*/;
  static constexpr int NO = PB / 2;                  // channels per V tile row = width of one O column group
  // TMEM: S = x_hi*[y_hi ; y_lo] lands in two column groups when the K tiles are issued as one N = 128 operand (SC);
  // for 32-byte head rows (dh = 16) S stays three plain N = 64 MMAs so that 128 columns (-> 3 CTAs / SM) suffice.
  static constexpr bool SC = (PB != 32);
  static constexpr int kSCols = SC ? 128 : 64;
  static constexpr int kOCols = 2 * NO;              // one O buffer: [0,NO) hi*hi + lo*hi, [NO,2NO) hi*lo
  static constexpr int kTmemCols = (kSCols + 2 * kOCols <= 128) ? 128 : (kSCols + 2 * kOCols <= 256) ? 256 : 512;
  static constexpr int kBySmem = (kSmem <= 56 * 1024) ? 4 : (kSmem <= 74 * 1024) ? 3 : (kSmem <= 113 * 1024) ? 2 : 1;
  static constexpr int kMinCtas = (512 / kTmemCols < kBySmem) ? 512 / kTmemCols : kBySmem;
};

// Shared-memory matrix descriptor for a tile whose rows are PB bytes (SWIZZLE_<PB>B), 8-row groups dense.
template <int PB>
__device__ __forceinline__ uint64_t desc_pb(uint32_t saddr, uint32_t lbo_bytes) {
  constexpr uint64_t layout = (PB == 32) ? 6 : (PB == 64) ? 4 : 2;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)((8 * PB) >> 4) << 32) |
         (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int N>
__device__ __forceinline__ void tmem_ld_nw(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld_nw<8>(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }
// Per-tile role timeline of CTA (0,0,0) (scripts/trace_attn.py): compiled in only with -DNS2VC_ATTN_TRACE (NVCC_EXTRA of build.sh) -
// eight predicated stamps per key tile are ~7 % of the softmax loop's issue slots and a live pointer in a 56-register kernel.
#ifdef NS2VC_ATTN_TRACE
#define ATRACE(j, slot) do { if (tr && (j) < 16) tr[(j) * 16 + (slot)] = clk(); } while (0)
#else
#define ATRACE(j, slot) do { } while (0)
#endif
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int DHP, int PB, bool BIAS, bool PF16>
__global__ void __launch_bounds__(kThreadsV2, ACfg<DHP, PB, PF16, BIAS>::kMinCtas) attn_v2_kernel(const __grid_constant__ AttnOp op) {
  using C = ACfg<DHP, PB, PF16, BIAS>;
  constexpr int NST = C::NST;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar0 = base + C::kOffBar;
  const uint32_t q_full = bar0, s_full = bar0 + 8, s_empty = bar0 + 16, p_full = bar0 + 24, o_full = bar0 + 32;
  auto kv_full = [&](int s) { return bar0 + 40u + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 40u + 8u * (NST + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::kOffBar + 40 + 16 * NST);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQ;
  const int dh = op.dh;
  const int ntiles = (op.Tk + kKeys - 1) / kKeys;

  span_begin(op.span);
  // diagnostics: [start, end, SM id] of every CTA at trace[256 + 3 * linear CTA id] (globaltimer ns)
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (op.trace && tid == 0 && cta_lin < 597) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    op.trace[256 + 3 * cta_lin] = gtime_ns();
    op.trace[256 + 3 * cta_lin + 2] = smid;
  }
#ifdef NS2VC_ATTN_TRACE
  unsigned long long* tr = (op.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (warp == 0 || (warp == 1 && lane == 0))) ? op.trace : nullptr;
#endif
  if (tid == 0) {
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(s_empty, 256); mbar_init(p_full, 256); mbar_init(o_full, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), C::kTmemCols);
  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) prefetch_tmap(&op.tm[i]);
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + C::kSCols;   // O: two buffers of kOCols columns (tile parity)
  const uint32_t sQ = base + C::kOffQ, sP = base + C::kOffP, sKV = base + C::kOffKV;

  // Warp 0: ONE elected thread streams the tiles AND issues the MMAs (eight softmax warps + this one = 288 threads, so that
  // four CTAs fit an SM).  Under elect.sync ptxas issues the uniform-datapath instructions (UTMALDG, UTCHMMA, UTCBAR) back to
  // back; as two divergent lanes of one warp each of them sat in an ELECT / BRA.U.ANY serialisation loop.
  if (warp == 0) {
    if (elect_one()) {
      auto load_kv = [&](int j) {
        const int stage = j % NST;
        const uint32_t dst = sKV + stage * C::kStageBytes;
        mbar_arrive_expect_tx(kv_full(stage), (uint32_t)C::kStageBytes);
        tma_load_3d(dst, &op.tm[2], op.k_c0 + h * dh, j * kKeys, b, kv_full(stage));
        tma_load_3d(dst + C::kTBytes, &op.tm[3], op.k_c0 + h * dh, j * kKeys, b, kv_full(stage));
        tma_load_3d(dst + 2 * C::kTBytes, &op.tm[4], op.v_c0 + h * dh, j * kKeys, b, kv_full(stage));
        tma_load_3d(dst + 3 * C::kTBytes, &op.tm[5], op.v_c0 + h * dh, j * kKeys, b, kv_full(stage));
      };
      pdl_wait();                                           // q / k / v are the previous kernels' outputs
      mbar_arrive_expect_tx(q_full, 2u * C::kQBytes);
      tma_load_3d(sQ, &op.tm[0], op.q_c0 + h * dh, q0, b, q_full);
      tma_load_3d(sQ + C::kQBytes, &op.tm[1], op.q_c0 + h * dh, q0, b, q_full);
      for (int j = 0; j < NST && j < ntiles; ++j) { ATRACE(j, 12); load_kv(j); }
      // Every product is hi*hi + hi*lo + lo*hi.  The hi and lo tiles of K (and of V) are adjacent in shared memory,
      // so X_hi x [Y_hi ; Y_lo] is ONE instruction of twice the N whose result lands in two TMEM column groups;
      // X_lo x Y_hi accumulates into the first group and the softmax warps add the groups.
      constexpr uint32_t idS2 = umma_idesc_bf16(kQ, 2 * kKeys), idS1 = umma_idesc_bf16(kQ, kKeys);                    // A, B K-major
      constexpr uint32_t idO2 = umma_idesc_bf16(kQ, 2 * C::NO) | (1u << 16), idO1 = umma_idesc_bf16(kQ, C::NO) | (1u << 16);   // B (= V) MN-major
      auto issue_S = [&](int j) {
        const uint32_t kst = sKV + (j % NST) * C::kStageBytes;
#pragma unroll
        for (int k = 0; k < DHP / 16; ++k) {
          const uint64_t qh = desc_pb<PB>(sQ + k * 32, 16), ql = desc_pb<PB>(sQ + C::kQBytes + k * 32, 16);
          const uint64_t kh = desc_pb<PB>(kst + k * 32, 16);                   // rows [0,64) = K_hi, rows [64,128) = K_lo
          if (C::SC) {
            umma_bf16(tS, qh, kh, idS2, k != 0 ? 1u : 0u);
            umma_bf16(tS, ql, kh, idS1, 1u);
          } else {
            const uint64_t kl = desc_pb<PB>(kst + C::kTBytes + k * 32, 16);
            umma_bf16(tS, qh, kh, idS1, k != 0 ? 1u : 0u);
            umma_bf16(tS, qh, kl, idS1, 1u);
            umma_bf16(tS, ql, kh, idS1, 1u);
          }
        }
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_S(0);
      // Refill of the ring stage tile j-1 occupied (with tile j-1+NST), once PV(j-1) has retired.  Three stages: after the
      // wait for P(j) - the softmax warps waited for PV(j-1) before they stored P(j), so the stage is free without waiting.
      // Two stages (128-byte head rows): tile j+1 itself goes there, so it is refilled first thing in iteration j.
      auto refill = [&](int j) {
        if (j >= 1 && j - 1 + NST < ntiles) {
          mbar_wait(kv_empty((j - 1) % NST), (uint32_t)(((j - 1) / NST) & 1));
          ATRACE(j - 1 + NST, 12);
          load_kv(j - 1 + NST);
        }
      };
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t par = (uint32_t)(j & 1);
        if (NST < 3) refill(j);
        if (j + 1 < ntiles) {
          mbar_wait(kv_full((j + 1) % NST), (uint32_t)(((j + 1) / NST) & 1));
          mbar_wait(s_empty, par);                          // every softmax thread holds S(j) in registers
          tc_fence_after();
          ATRACE(j, 8);
          issue_S(j + 1);
          ATRACE(j, 9);
        }
        mbar_wait(p_full, par);                             // P(j) is in shared memory, O_tile(j-1) has been read
        tc_fence_after();
        ATRACE(j, 10);
        if (NST >= 3) refill(j);
        const uint32_t vst = sKV + (j % NST) * C::kStageBytes + 2 * C::kTBytes;
#pragma unroll
        for (int k = 0; k < kKeys / 16; ++k) {
          const uint64_t ph = umma_desc(sP + k * 32);
          const uint64_t vh = desc_pb<PB>(vst + k * 16 * PB, C::kTBytes);      // channel group 0 = V_hi, group 1 (+LBO) = V_lo
          if (PF16) {
            umma_bf16(tO + par * C::kOCols, ph, vh, idO2 & ~((7u << 7) | (7u << 10)), k != 0 ? 1u : 0u);   // A = fp16 P, B = fp16 [V_hi | V_lo] (format fields 0)
          } else {
            const uint64_t pl = umma_desc(sP + C::kPBytes + k * 32);
            umma_bf16(tO + par * C::kOCols, ph, vh, idO2, k != 0 ? 1u : 0u);
            umma_bf16(tO + par * C::kOCols, pl, vh, idO1, 1u);
          }
        }
        umma_commit(o_full);
        umma_commit(kv_empty(j % NST));                     // K(j), V(j) consumed
        ATRACE(j, 11);
      }
    }
  } else {
    // ===================== softmax warps =====================
    constexpr int OH = DHP / 2;                             // output columns per thread
    const int qtr = warp & 3, hf = (warp - 1) >> 2;
    const int r = qtr * 32 + lane;                          // query row = TMEM lane
    const uint32_t lane_base = ((uint32_t)(qtr * 32)) << 16;
    float* xch = reinterpret_cast<float*>(smem + C::kOffXch);
    float* bias_s = reinterpret_cast<float*>(smem + C::kOffBias);
    const float qscale = op.scale * 1.4426950408889634f;
    pdl_wait();                                             // the mask bias and the output buffers belong to earlier kernels
    if (BIAS) {                                             // additive mask bias * log2(e); -inf past Tk
      const float* bias = op.bias + (long long)b * op.Tk;
      for (int i = tid - 32; i < ntiles * kKeys; i += 256)
        bias_s[i] = (i < op.Tk) ? __ldg(bias + i) * 1.4426950408889634f : -INFINITY;
      asm volatile("bar.sync 5, 256;" ::: "memory");
    }
    float o[OH];
#pragma unroll
    for (int d = 0; d < OH; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    auto add_o_tile = [&](uint32_t buf, float scale) {      // o = (o + O_tile[buf] (TMEM)) * scale
      float ot[OH], ou[OH];
      const uint32_t to = tO + buf * C::kOCols + lane_base + hf * OH;
#pragma unroll
      for (int d0 = 0; d0 < OH; d0 += 8) {
        tmem_ld_nw<8>(to + d0, ot + d0);
        tmem_ld_nw<8>(to + C::NO + d0, ou + d0);
      }
      tmem_wait_ld();
      const unsigned long long s2 = pk2(scale, scale);
#pragma unroll
      for (int d = 0; d < OH; d += 2)
        upk2(fmul2(fadd2(pk2(o[d], o[d + 1]), fadd2(pk2(ot[d], ot[d + 1]), pk2(ou[d], ou[d + 1]))), s2), o[d], o[d + 1]);
    };

    for (int j = 0; j < ntiles; ++j) {
      const uint32_t par = (uint32_t)(j & 1);
      ATRACE(j, 0);
      mbar_wait_quiet(s_full, par);
      tc_fence_after();
      ATRACE(j, 1);
      float sv[32];
      if (C::SC) tmem_ld32_sum(tS + lane_base + hf * 32, tS + lane_base + kKeys + hf * 32, sv);
      else tmem_ld32(tS + lane_base + hf * 32, sv);
      tc_fence_before();
      mbar_arrive(s_empty);
      ATRACE(j, 2);
      const int kbase = j * kKeys + hf * 32;
      float mt = -INFINITY;
      const unsigned long long qs2 = pk2(qscale, qscale);
      if (BIAS) {
        const float4* bp = reinterpret_cast<const float4*>(bias_s + kbase);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 bb = bp[c4];
          upk2(ffma2(pk2(sv[4 * c4 + 0], sv[4 * c4 + 1]), qs2, pk2(bb.x, bb.y)), sv[4 * c4 + 0], sv[4 * c4 + 1]);
          upk2(ffma2(pk2(sv[4 * c4 + 2], sv[4 * c4 + 3]), qs2, pk2(bb.z, bb.w)), sv[4 * c4 + 2], sv[4 * c4 + 3]);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) mt = fmaxf(mt, sv[c]);
      } else {
        if (j == ntiles - 1 && (op.Tk & (kKeys - 1))) {     // keys past Tk were zero-filled by TMA: mask them
          const int nvalid = op.Tk - kbase;
#pragma unroll
          for (int c = 0; c < 32; ++c) if (c >= nvalid) sv[c] = -INFINITY;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) mt = fmaxf(mt, sv[c]);
        mt *= qscale;                                       // qscale > 0: max commutes with the scaling
      }
      xch[(par * 2 + hf) * 128 + r] = mt;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qtr) : "memory");          // the two warps of this lane quarter
      ATRACE(j, 3);
      const float m_new = fmaxf(m_run, fmaxf(mt, xch[(par * 2 + (hf ^ 1)) * 128 + r]));
      const float corr = ex2f(m_run - m_new);
      const unsigned long long nm2 = pk2(-m_new, -m_new);
      unsigned long long lt2 = pk2(0.f, 0.f);
      // PV(j-1) was issued as soon as P(j-1) was complete, i.e. before these warps even picked up S(j): by now it has
      // retired (the P buffer is free, O_tile(j-1) complete).  Waiting for it HERE lets every 8-column group of P go to
      // shared memory as soon as it is computed instead of staying live in registers across the wait (the kernel runs
      // at 56 registers per thread for four CTAs per SM).
      if (j > 0) {
        mbar_wait_quiet(o_full, par ^ 1u);
        tc_fence_after();
      }
      ATRACE(j, 5);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t ph2[4];
#pragma unroll
        for (int c = 8 * c8; c < 8 * c8 + 8; c += 2) {
          float a, bq;
          if (BIAS) upk2(fadd2(pk2(sv[c], sv[c + 1]), nm2), a, bq);
          else upk2(ffma2(pk2(sv[c], sv[c + 1]), qs2, nm2), a, bq);
          sv[c] = ex2f(a); sv[c + 1] = ex2f(bq);
          if (PF16) {                                       // the weights ARE the fp16-rounded values: numerator and row sum agree
            const uint32_t h2 = pack_f16x2(sv[c], sv[c + 1]);
            ph2[(c >> 1) & 3] = h2;
            sv[c] = f16lo_to_f32(h2); sv[c + 1] = f16hi_to_f32(h2);
          }
          lt2 = fadd2(lt2, pk2(sv[c], sv[c + 1]));
        }
        const int ck = hf * 4 + c8;
        const int off = r * 128 + ((ck ^ (r & 7)) << 4);
        if (PF16) {
          *reinterpret_cast<uint4*>(smem + C::kOffP + off) = make_uint4(ph2[0], ph2[1], ph2[2], ph2[3]);
        } else {
          uint4 hi, lo;
          split8(sv + 8 * c8, hi, lo);
          *reinterpret_cast<uint4*>(smem + C::kOffP + off) = hi;
          *reinterpret_cast<uint4*>(smem + C::kOffP + C::kPBytes + off) = lo;
        }
      }
      float lt, lt_hi;
      upk2(lt2, lt, lt_hi);
      lt += lt_hi;
      l_run = l_run * corr + lt;
      m_run = m_new;
      ATRACE(j, 4);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
      ATRACE(j, 6);
      // O_tile(j-1) (the other TMEM buffer than the one PV(j) is about to fill) is relative to the previous running max
      if (j > 0) add_o_tile(par ^ 1u, corr);
      ATRACE(j, 7);
    }
    mbar_wait_quiet(o_full, (uint32_t)((ntiles - 1) & 1));
    tc_fence_after();
    add_o_tile((uint32_t)((ntiles - 1) & 1), 1.0f);

    // total row sum = the two halves' partial sums
    xch[(4 + hf) * 128 + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qtr) : "memory");
    const float l_tot = l_run + xch[(4 + (hf ^ 1)) * 128 + r];
    if (q0 + r < op.Tq) {
      const float inv = 1.0f / l_tot;
      const long long orow = (long long)b * op.Tq + q0 + r;
      const int dbase = hf * OH;
      if (op.out) {
        float* po = op.out + orow * op.out_ld + h * dh;
#pragma unroll
        for (int d = 0; d < OH; ++d) if (dbase + d < dh) po[dbase + d] = o[d] * inv;
      }
      if (op.out_hi) {
        __nv_bfloat16* ph = op.out_hi + orow * op.out_split_ld + h * dh + dbase;
        __nv_bfloat16* pl = op.out_lo + orow * op.out_split_ld + h * dh + dbase;
        if (((op.out_split_ld | (h * dh)) & 7) == 0) {      // dh % 16 == 0 here, so dbase % 8 == 0
#pragma unroll
          for (int d0 = 0; d0 < OH; d0 += 8) {
            float v[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) v[d] = o[d0 + d] * inv;
            uint4 hi, lo;
            split8(v, hi, lo);
            *reinterpret_cast<uint4*>(ph + d0) = hi;
            *reinterpret_cast<uint4*>(pl + d0) = lo;
          }
        } else {
#pragma unroll
          for (int d = 0; d < OH; ++d) {
            const float v = o[d] * inv;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            ph[d] = hi;
            pl[d] = __float2bfloat16_rn(v - __bfloat162float(hi));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  span_end(op.span);
  if (op.trace && tid == 0 && cta_lin < 597) op.trace[256 + 3 * cta_lin + 1] = gtime_ns();
  if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int DHP, int PB, bool BIAS, bool PF16>
int launch_v2b(const AttnOp& op, cudaStream_t st) {
  using C = ACfg<DHP, PB, PF16, BIAS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_v2_kernel<DHP, PB, BIAS, PF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) { set_error("attention v2: cannot set %d B dynamic smem: %s", C::kSmem, cudaGetErrorString(e)); return -2; }
    // ask for the largest shared-memory carve-out: the CTA count per SM is what the smem budget above was sized for
    cudaFuncSetAttribute(attn_v2_kernel<DHP, PB, BIAS, PF16>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (getenv("NS2VC_DEBUG")) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attn_v2_kernel<DHP, PB, BIAS, PF16>, kThreadsV2, (size_t)C::kSmem);
      fprintf(stderr, "ns2vc: attn_v2<%d,%d,bias=%d,f16=%d> smem %d B, planned %d CTAs/SM, occupancy API says %d\n", DHP, PB, (int)BIAS, (int)PF16, C::kSmem, C::kMinCtas, nb);
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(op.Tq, kQ), op.H, op.B);
  cudaError_t e = launch_k(attn_v2_kernel<DHP, PB, BIAS, PF16>, grid, dim3(kThreadsV2), (size_t)C::kSmem, st, op);
  if (e != cudaSuccess) { set_error("attention v2 launch failed: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}
// Softmax weights: fp16 (one P*[V_hi|V_lo] MMA per k-step; the weights are DEFINED as the rounded values, so the result
// is an exact weighted mean with weights perturbed by <= 2^-12 relative) or bf16 hi/lo split (NS2VC_ATTN_P=split).
bool p_fp16() { return attention_v2_p_fp16(); }
template <int DHP, int PB>
int launch_v2(const AttnOp& op, cudaStream_t st) {
  if (op.bias) {
    if (ceil_div(op.Tk, kKeys) * kKeys > ACfg<DHP, PB, true, true>::kBiasKeys) { set_error("attention v2: %d biased keys exceed the staged-bias capacity", op.Tk); return -1; }
    return (p_fp16() && !op.p_split) ? launch_v2b<DHP, PB, true, true>(op, st) : launch_v2b<DHP, PB, true, false>(op, st);
  }
  return (p_fp16() && !op.p_split) ? launch_v2b<DHP, PB, false, true>(op, st) : launch_v2b<DHP, PB, false, false>(op, st);
}

int natural_pb(int dh) { return dh == 16 ? 32 : dh == 32 ? 64 : 128; }

}  // namespace

bool attention_v2_p_fp16() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NS2VC_ATTN_P"); v = (e && e[0] == 's') ? 0 : 1; }
  return v == 1;
}

bool attention_v2_supported(int dh, int Tk, bool biased) {
  if (!(dh == 16 || dh == 32 || dh == 48 || dh == 64)) return false;
  return !biased || ceil_div(Tk, kKeys) * kKeys <= (dh == 16 ? 768 : 1024);   // the additive bias of a row of keys is staged in shared memory
}

int encode_attn_tmaps(AttnOp& op) {
  op.pb = natural_pb(op.dh);                                // (padded 128-byte rows for every head dim were measured slower in r01)
  const int bc = op.pb / 2;                                 // box width in channels
  int rc = 0;
  if ((rc = encode_tmap_rows(&op.tm[0], op.qs.hi, op.qs.C, op.qs.T, op.B, op.qs.ld, bc, kQ, op.pb))) return rc;
  if ((rc = encode_tmap_rows(&op.tm[1], op.qs.lo, op.qs.C, op.qs.T, op.B, op.qs.ld, bc, kQ, op.pb))) return rc;
  if ((rc = encode_tmap_rows(&op.tm[2], op.ks.hi, op.ks.C, op.ks.T, op.B, op.ks.ld, bc, kKeys, op.pb))) return rc;
  if ((rc = encode_tmap_rows(&op.tm[3], op.ks.lo, op.ks.C, op.ks.T, op.B, op.ks.ld, bc, kKeys, op.pb))) return rc;
  if ((rc = encode_tmap_rows(&op.tm[4], op.vs.hi, op.vs.C, op.vs.T, op.B, op.vs.ld, bc, kKeys, op.pb))) return rc;
  if ((rc = encode_tmap_rows(&op.tm[5], op.vs.lo, op.vs.C, op.vs.T, op.B, op.vs.ld, bc, kKeys, op.pb))) return rc;
  return 0;
}

int launch_attention_v2(const AttnOp& op, cudaStream_t st) {
  if (op.Tk <= 0 || op.Tq <= 0) { set_error("attention: empty sequence"); return -1; }
  if (op.qs.T != op.Tq || op.ks.T != op.Tk || op.vs.T != op.Tk) { set_error("attention v2: split buffer / sequence length mismatch"); return -1; }
  const int dh = op.dh;
  if (op.pb == 128) {
    if (dh == 48) return launch_v2<48, 128>(op, st);
    if (dh == 64) return launch_v2<64, 128>(op, st);
  } else if (op.pb == 64 && dh == 32) {
    return launch_v2<32, 64>(op, st);
  } else if (op.pb == 32 && dh == 16) {
    return launch_v2<16, 32>(op, st);
  }
  set_error("attention v2: unsupported head dim %d / box %d", dh, op.pb);
  return -1;
}

}  // namespace ns2vc
