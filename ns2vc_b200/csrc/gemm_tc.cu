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
// A operand: the pre-normalised "split" activations [B, T, C] (bf16 hi / lo), fetched by TMA
// (cp.async.bulk.tensor.3d, SWIZZLE_128B) as {64 channels x 128 rows} boxes of ONE batch entry.
// A conv tap is the same box shifted by one row; rows before 0 / past T and channels past C are
// zero-filled by the TMA unit, which is the conv's zero padding and the ragged-edge handling at
// once.  Channel concats are just two tensor maps.
// B operand: weights pre-packed as the swizzled smem image, one cp.async.bulk per hi/lo tile.
//
// Warp roles (192 threads, 1 CTA/SM, 3-4 stage mbarrier ring):
//   warp 0     TMA producer (one elected lane)
//   warp 1     MMA issuer (one elected lane): tcgen05.mma x12 per k-block, tcgen05.commit
//   warps 2-5  epilogue: TMEM -> registers (tcgen05.ld 32x32b) -> bias / GEGLU / residual ->
//              fp32 token-major, fp32 channel-major, or bf16 hi/lo split for the next GEMM
#include "gemm_common.cuh"
#include <cuda.h>
#include <cstdio>

namespace ns2vc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 4;
constexpr int kThreads = (2 + kEpiWarps) * 32;
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 3-D tiled TMA load: box {64 ch, 128 rows, 1 batch} at (c, t, b); out-of-range coordinates are zero-filled
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const TMap* tmap, int c, int t, int b, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(t), "r"(b)
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

// Store 32 consecutive output columns of one row in the layout(s) the op asks for.
__device__ __forceinline__ void store_chunk(const GemmOp& op, int b, int t, long long m, int nbase, const float* val) {
  if (op.flags & EPI_OUT_NCT) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = nbase + j;
      if (n < op.n_valid) op.out[((long long)b * op.n_valid + n) * op.T_out + t] = val[j];
    }
    return;
  }
  const bool fullc = nbase + 32 <= op.n_valid;
  if (op.flags & EPI_OUT_F32) {
    float* po = op.out + m * op.out_ld + nbase;
    if (fullc && ((op.out_ld & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(po + 4 * j) = make_float4(val[4 * j], val[4 * j + 1], val[4 * j + 2], val[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (nbase + j < op.n_valid) po[j] = val[j];
    }
  }
  if (op.flags & EPI_OUT_SPLIT) {
    __nv_bfloat16* ph = op.out_hi + m * op.out_split_ld + nbase;
    __nv_bfloat16* pl = op.out_lo + m * op.out_split_ld + nbase;
    if (fullc && ((op.out_split_ld & 7) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 hi, lo;
        split8(val + 8 * j, hi, lo);
        *reinterpret_cast<uint4*>(ph + 8 * j) = hi;
        *reinterpret_cast<uint4*>(pl + 8 * j) = lo;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nbase + j < op.n_valid) {
          const __nv_bfloat16 h = __float2bfloat16_rn(val[j]);
          ph[j] = h;
          pl[j] = __float2bfloat16_rn(val[j] - __bfloat162float(h));
        }
    }
  }
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
  const int tiles_per_batch = (op.T_out + BM - 1) / BM;
  const int b = blockIdx.x / tiles_per_batch;
  const int t0 = (blockIdx.x % tiles_per_batch) * BM;
  const int n0 = blockIdx.y * BN;                           // first packed column of this tile
  const int nkb = op.nkb_total;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int si = 0, kbl = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % kStages;
        const uint32_t parity = (uint32_t)((kb / kStages) & 1);
        mbar_wait(empty_bar(stage), parity ^ 1u);
        const GSeg& s = op.seg[si];
        const uint32_t a_hi = base + stage * kStageBytes;
        const uint32_t a_lo = a_hi + kATileBytes;
        const uint32_t b_hi = a_lo + kATileBytes;
        const uint32_t b_lo = b_hi + Cfg::kBTileBytes;
        mbar_arrive_expect_tx(full_bar(stage), 2u * kATileBytes + 2u * Cfg::kBTileBytes);
        const int c = s.c0 + kbl * 64;
        tma_load_3d(a_hi, &op.tmap[2 * s.src], c, t0 + s.tap, b, full_bar(stage));
        tma_load_3d(a_lo, &op.tmap[2 * s.src + 1], c, t0 + s.tap, b, full_bar(stage));
        const size_t eoff = ((size_t)kb * op.N + n0) * 64;
        bulk_g2s(b_hi, op.w_hi + eoff, Cfg::kBTileBytes, full_bar(stage));
        bulk_g2s(b_lo, op.w_lo + eoff, Cfg::kBTileBytes, full_bar(stage));
        if (++kbl == s.nkb) { kbl = 0; ++si; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % kStages;
        const uint32_t parity = (uint32_t)((kb / kStages) & 1);
        mbar_wait(full_bar(stage), parity);
        tc_fence_after();
        const uint32_t a_hi = base + stage * kStageBytes;
        const uint32_t a_lo = a_hi + kATileBytes;
        const uint32_t b_hi = a_lo + kATileBytes;
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
  } else {
    // ===================== epilogue (warps 2..5) =====================
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int t = t0 + r;
    const bool mv = t < op.T_out;
    const long long m = (long long)b * op.T_out + t;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if (op.flags & EPI_GEGLU) {
      if (BN == 128) {
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          float val[32], gate[32];
          tmem_ld32(trow + (uint32_t)(hh * 32), val);
          tmem_ld32(trow + (uint32_t)(64 + hh * 32), gate);
          const int nbase = blockIdx.y * 64 + hh * 32;      // logical output column
          if (mv) {
#pragma unroll
            for (int j = 0; j < 32; ++j) val[j] = epi_value(op, b, m, nbase + j, val[j], gate[j]);
            store_chunk(op, b, t, m, nbase, val);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        float acc[32];
        tmem_ld32(trow + (uint32_t)(cc * 32), acc);
        const int nbase = n0 + cc * 32;
        if (!mv || nbase >= op.n_valid) continue;
        const bool fullc = nbase + 32 <= op.n_valid;
        if (fullc && !(op.flags & EPI_ROWBIAS) && (!(op.flags & EPI_RESIDUAL) || (op.res_ld & 3) == 0)) {
          if (op.flags & EPI_BIAS) {
            const float4* pb = reinterpret_cast<const float4*>(op.bias + nbase);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 v = __ldg(pb + j); acc[4 * j] += v.x; acc[4 * j + 1] += v.y; acc[4 * j + 2] += v.z; acc[4 * j + 3] += v.w; }
          }
          if (op.flags & EPI_RESIDUAL) {
            const float4* pr = reinterpret_cast<const float4*>(op.res + m * op.res_ld + nbase);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 v = __ldg(pr + j); acc[4 * j] += v.x; acc[4 * j + 1] += v.y; acc[4 * j + 2] += v.z; acc[4 * j + 3] += v.w; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (nbase + j < op.n_valid) acc[j] = epi_value(op, b, m, nbase + j, acc[j], 0.f);
        }
        store_chunk(op, b, t, m, nbase, acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Host: TMA descriptor encoding (driver entry point fetched through the runtime; no -lcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode_one(TMap* out, const __nv_bfloat16* base, const SplitBuf& s, int B) {
  static_assert(sizeof(CUtensorMap) == sizeof(TMap), "CUtensorMap size");
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return -2; }
  if ((s.ld & 7) || (reinterpret_cast<uintptr_t>(base) & 15)) { set_error("split buffer not TMA-aligned (ld=%d)", s.ld); return -1; }
  const cuuint64_t gdim[3] = {(cuuint64_t)s.C, (cuuint64_t)s.T, (cuuint64_t)B};
  const cuuint64_t gstr[2] = {(cuuint64_t)s.ld * 2, (cuuint64_t)s.T * s.ld * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)BM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base), gdim, gstr,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) C=%d T=%d B=%d ld=%d", (int)r, s.C, s.T, B, s.ld); return -2; }
  return 0;
}

int encode_tmaps(GemmOp& op) {
  for (int i = 0; i < op.nsrc; ++i) {
    int rc = encode_one(&op.tmap[2 * i], op.src[i].hi, op.src[i], op.B);
    if (rc) return rc;
    rc = encode_one(&op.tmap[2 * i + 1], op.src[i].lo, op.src[i], op.B);
    if (rc) return rc;
  }
  return 0;
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
  dim3 grid(op.B * ceil_div(op.T_out, BM), op.N / BN_);
  gemm_tc_kernel<BN_><<<grid, kThreads, Cfg::kSmemBytes, st>>>(op);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("gemm_tc launch failed: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

int launch_gemm_tc(const GemmOp& op, cudaStream_t st) {
  if (op.N % 128) { set_error("gemm_tc: packed N=%d is not a multiple of 128", op.N); return -1; }
  if (op.nkb_total <= 0) { set_error("gemm_tc: empty K"); return -1; }
  // N-tile: 128 wide when that already fills the 148 SMs, else 64 wide (twice the CTAs; the
  // GEGLU epilogue pairs value|gate inside a 128-column block and needs BN = 128).
  const int ctas128 = op.B * ceil_div(op.T_out, BM) * (op.N / 128);
  if ((op.flags & EPI_GEGLU) || ctas128 >= 120) return launch_bn<128>(op, st);
  return launch_bn<64>(op, st);
}

}  // namespace ns2vc
