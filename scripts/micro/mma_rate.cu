// Micro-benchmark: SM cycles per 16-wide k-step of tcgen05.mma (SS mode, M = 128, bf16) for different ways of issuing the
// 3xBF16 products A_hi*[B_hi|B_lo] and A_lo*B_hi: into ONE accumulator region (each MMA depends on the previous one), into
// separate regions (independent chains), with even / odd k-steps on separate regions, and with row-shifted A views.
#include "../../ns2vc_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <vector>
using namespace ns2vc;

// per group: n MMAs of width N_i into TMEM column C_i (compile-time: the issue loop must be as tight as the product's)
template <int SH, int N0, int C0, int N1 = 0, int C1 = 0, int N2 = 0, int C2 = 0, int N3 = 0, int C3 = 0>
__global__ void __launch_bounds__(128) k(long long* out, int reps) {
  extern __shared__ uint8_t raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < 60 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(raw_ + (base - raw))[i] = 0x3c003c00u + i;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(smem_u32(&tslot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    const uint32_t a0 = base + 128 * SH, b0 = base + 24576;
    t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = umma_desc(a0 + kk * 32), db = umma_desc(b0 + kk * 32);
        umma_bf16(tm + C0, da, db, umma_idesc_bf16(128, N0), (r | kk) ? 1u : 0u);
        if constexpr (N1 > 0) umma_bf16(tm + C1, da, db, umma_idesc_bf16(128, N1), (r | kk) ? 1u : 0u);
        if constexpr (N2 > 0) umma_bf16(tm + C2, da, db, umma_idesc_bf16(128, N2), (r | kk) ? 1u : 0u);
        if constexpr (N3 > 0) umma_bf16(tm + C3, da, db, umma_idesc_bf16(128, N3), (r | kk) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
    t1 = clock64();
  }
  mbar_wait(smem_u32(&bar), 0);
  t2 = clock64();
  tc_fence_after();
  if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 512);
}

template <int SH, int N0, int C0, int N1 = 0, int C1 = 0, int N2 = 0, int C2 = 0, int N3 = 0, int C3 = 0>
void run(long long* d, const char* name) {
  auto kf = k<SH, N0, C0, N1, C1, N2, C2, N3, C3>;
  cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int reps : {16, 128}) {
    long long h[2];
    for (int it = 0; it < 2; ++it) { kf<<<148, 128, 64 * 1024>>>(d, reps); cudaDeviceSynchronize(); }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    auto e = cudaGetLastError();
    printf("%-88s groups=%3d  issue %6lld  done %6lld cyc => %6.1f cyc per group %s\n", name, reps, h[0], h[1], (double)h[1] / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  // one "group" = the MMAs of one k-step (or of two k-steps for the even/odd patterns: then cycles per k-step = half)
  run<0, 128, 0, 64, 0>(d, "BN=64 now: hi*[Bh|Bl] -> [0,128), lo*Bh -> [0,64) (dependent)        [1 k-step]");
  run<0, 128, 0, 64, 128>(d, "BN=64: lo*Bh into its own region [128,192)                          [1 k-step]");
  run<1, 128, 0, 64, 128>(d, "  same, A view shifted by one row                                    [1 k-step]");
  run<2, 128, 0, 64, 128>(d, "  same, A view shifted by two rows                                   [1 k-step]");
  run<0, 128, 0, 64, 128, 128, 192, 64, 320>(d, "BN=64: own lo region + even/odd k-steps on separate regions        [2 k-steps]");
  run<0, 128, 0, 64, 0, 128, 192, 64, 192>(d, "BN=64: dependent pair, even/odd k-steps on separate regions         [2 k-steps]");
  run<0, 64, 0, 64, 64, 64, 128>(d, "BN=64: three N=64 MMAs into three regions                           [1 k-step]");
  run<0, 64, 0>(d, "single N=64 chain                                                   [1 MMA]");
  run<0, 128, 0>(d, "single N=128 chain                                                  [1 MMA]");
  run<0, 256, 0>(d, "single N=256 chain                                                  [1 MMA]");
  run<0, 64, 0, 64, 64>(d, "two independent N=64 chains                                         [2 MMAs]");
  run<0, 64, 0, 64, 64, 64, 128, 64, 192>(d, "four independent N=64 chains                                        [4 MMAs]");
  run<0, 128, 0, 128, 128>(d, "two independent N=128 chains                                        [2 MMAs]");
  run<0, 256, 0, 128, 0>(d, "BN=128 now: hi*[Bh|Bl] -> [0,256), lo*Bh -> [0,128) (dependent)     [1 k-step]");
  run<0, 256, 0, 128, 256>(d, "BN=128: lo*Bh into its own region [256,384)                        [1 k-step]");
  run<0, 256, 0, 128, 0, 256, 256, 128, 256>(d, "BN=128: dependent pair, even/odd k-steps on separate regions        [2 k-steps]");
  return 0;
}
