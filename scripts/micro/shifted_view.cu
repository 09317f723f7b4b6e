// Micro-test: can one 128B-swizzled K-major smem panel of 136 rows serve the three row-shifted A tiles of a k=3 conv?
// A tile for tap j = rows [j, j+128) of the panel: descriptor start address = panel + 128*j, matrix-base-offset field = ?
// D = A_view(j) x B^T with B = first 64 rows of the identity => D[i][n] must equal panel[i + j][n].
#include "../../ns2vc_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <vector>
using namespace ns2vc;

__global__ void __launch_bounds__(128) k(float* out, int shift, int bo, int mode) {
  extern __shared__ uint8_t raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = raw_ + (base - raw);
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  // panel: 136 rows x 64 bf16 (128 B), element (r, c) = r + c/64 (distinct per row and column), stored with the address-based 128B swizzle
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);
  for (int i = tid; i < 136 * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c / 8) ^ (r & 7)) << 4) + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16((float)r + (float)c / 64.0f);
  }
  // B: [64 n][64 k] identity, swizzled, at offset 20480 (1024-aligned)
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i / 64, c = i % 64;
    const int off = 20480 + n * 128 + (((c / 8) ^ (n & 7)) << 4) + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16(n == c ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(smem_u32(&tslot), 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (tid == 0) {
    const uint32_t a0 = base + 128 * shift;
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t da = umma_desc(a0 + kk * 32);
      if (mode == 1) da = umma_desc(base + kk * 32) + (uint64_t)((128 * shift) >> 4);   // plain address add on an aligned descriptor (same bits)
      da |= (uint64_t)(bo & 7) << 49;
      const uint64_t db = umma_desc(base + 20480 + kk * 32);
      umma_bf16(tm, da, db, umma_idesc_bf16(128, 64), kk ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  float v[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + h * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 64);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  std::vector<float> h(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
  for (int shift = 0; shift < 3; ++shift)
    for (int bo = 0; bo < 8; ++bo) {
      if (bo != 0 && bo != shift && bo != ((8 - shift) & 7)) continue;
      cudaMemset(d, 0, 128 * 64 * 4);
      k<<<1, 128, 40960>>>(d, shift, bo, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d base_offset %d: CUDA error %s\n", mode, shift, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0; int first = -1;
      for (int i = 0; i < 128; ++i) for (int n = 0; n < 64; ++n) {
        const float want = (float)__nv_bfloat16((float)(i + shift) + (float)n / 64.0f);
        if (h[i * 64 + n] != want) { if (first < 0) first = i * 64 + n; ++bad; }
      }
      printf("mode %d shift %d base_offset %d: %s (%d mismatches", mode, shift, bo, bad ? "WRONG" : "ok", bad);
      if (bad) printf(", first at row %d col %d: got %.4f", first / 64, first % 64, h[first]);
      printf(")\n");
    }
  return 0;
}
