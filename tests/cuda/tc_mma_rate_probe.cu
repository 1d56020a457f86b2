// Hardware probe 3: cycles per tcgen05.mma (M=128, K=16, bf16) issued back to back, as a function of N and of the
// A-operand layout: K-major SWIZZLE_128B rows (each MMA reads a 32-byte slice of 128 rows 128 B apart) versus
// K-major SWIZZLE_32B rows (dense 32-byte rows, 4 KB contiguous per MMA).  Operands are garbage (timing only).
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../longterm360fov_b200/csrc/tc_common.cuh"
using namespace tc;

__global__ void rate(int N, int a_row_bytes, int b_row_bytes, int nmma, int distinct, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  for (uint32_t i = tid * 16; i < 160 * 1024; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem_raw + (base - raw) + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tptr;
  if (tid == 0) {
    auto desc = [](uint32_t addr, uint32_t row_bytes) {
      const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
      return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(((8 * row_bytes) >> 4) & 0x3FFFu) << 32) |
             (1ull << 46) | ((uint64_t)layout << 61);
    };
    const uint32_t idesc = idesc_bf16_f32(128, N, 0, 0);
    const uint32_t a0 = base, b0 = base + 96 * 1024;
    // 8 descriptor pairs precomputed in registers, loop unrolled by 8: the issue path is MMA instructions only
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t aoff = a_row_bytes == 128 ? (k & 3) * 32 + (k >> 2) * 128 : k * 128 * 32;
      const uint32_t boff = b_row_bytes == 128 ? (k & 3) * 32 : (k & 3) * 256 * 32;
      ad[k] = desc(a0 + aoff, a_row_bytes);
      bd[k] = desc(b0 + boff, b_row_bytes);
    }
    const long long t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(td, ad[k], bd[k], idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(td, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int nmma = 2000;
  const int Ns[6] = {16, 32, 64, 128, 192, 256};
  for (int arb = 128; arb >= 32; arb /= 4)
    for (int brb = 128; brb >= 32; brb /= 4) {
      printf("A rows %3d B, B rows %3d B:", arb, brb);
      for (int ni = 0; ni < 6; ++ni) {
        rate<<<1, 128, 200 * 1024>>>(Ns[ni], arb, brb, nmma, 8, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("  N=%d: %.1f cyc/MMA (issue %.1f)", Ns[ni], (double)h[1] / nmma, (double)h[0] / nmma);
      }
      printf("\n");
    }
  return 0;
}
