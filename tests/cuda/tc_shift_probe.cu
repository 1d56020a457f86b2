// Hardware probe: can a K-major SWIZZLE_128B operand tile be addressed at a start address shifted by
// whole 128-byte rows (not a multiple of the 1024-byte swizzle repeat)?  If so, the taps of a
// convolution become descriptor offsets into ONE staged activation tile (no im2col replication).
// Variants: data swizzled by absolute address / relative to the shifted start; descriptor
// base_offset field (bits 49..51) = 0 or (start >> 7) & 7.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../longterm360fov_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int ROWS = 144, N = 32;

__global__ void probe(const __nv_bfloat16* Afull, const __nv_bfloat16* Bm, float* D, int shift, int rel_swizzle,
                      int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* As = smem;                 // ROWS x 128 B
  uint8_t* Bs = smem + 32 * 1024;     // N x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < ROWS * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    const int phase = rel_swizzle ? ((r - shift) & 7) : (r & 7);
    uint4 v = *reinterpret_cast<const uint4*>(Afull + r * 64 + c * 8);
    *reinterpret_cast<uint4*>(As + r * 128 + ((c ^ phase) << 4)) = v;
  }
  for (int idx = tid; idx < N * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    uint4 v = *reinterpret_cast<const uint4*>(Bm + r * 64 + c * 8);
    *reinterpret_cast<uint4*>(Bs + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tptr;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N, 0, 0);
    const uint32_t a0 = base + shift * 128;
    for (int k4 = 0; k4 < 4; ++k4) {
      uint64_t ad = smem_desc_sw128(a0 + k4 * 32, 16, 1024);
      if (use_base_offset) ad |= (uint64_t)((a0 >> 7) & 7) << 49;
      const uint64_t bd = smem_desc_sw128(base + 32 * 1024 + k4 * 32, 16, 1024);
      umma_bf16(td, ad, bd, idesc, k4 > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid < 128) {
    const int warp = tid >> 5;
    float v[32];
    tmem_ld16(td + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + 16, v + 16);
    tmem_ld_wait();
    for (int j = 0; j < N; ++j) D[tid * N + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(td, 32);
}

// Second probe: the same question for 64-byte and 32-byte swizzled K-major tiles (rows of 32 / 16 bf16),
// data written with the absolute-address swizzle  phys = a ^ (((a >> 7) & mask) << 4).
__global__ void probe_narrow(const __nv_bfloat16* Afull, const __nv_bfloat16* Bm, float* D, int shift, int row_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* As = smem;
  uint8_t* Bs = smem + 32 * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  const int cpr = row_bytes / 16;                       // 16-byte chunks per row: 4 (SW64) or 2 (SW32)
  const uint32_t mask = row_bytes == 64 ? 3u : 1u;
  const uint32_t layout = row_bytes == 64 ? 4u : 6u;    // SWIZZLE_64B / SWIZZLE_32B
  const int kelems = row_bytes / 2;                     // K per row
  for (int idx = tid; idx < ROWS * cpr; idx += blockDim.x) {
    const int r = idx / cpr, c = idx % cpr;
    uint4 v = *reinterpret_cast<const uint4*>(Afull + r * 64 + c * 8);
    const uint32_t a = (uint32_t)(r * row_bytes + c * 16);
    *reinterpret_cast<uint4*>(As + (a ^ (((a >> 7) & mask) << 4))) = v;
  }
  for (int idx = tid; idx < N * cpr; idx += blockDim.x) {
    const int r = idx / cpr, c = idx % cpr;
    uint4 v = *reinterpret_cast<const uint4*>(Bm + r * 64 + c * 8);
    const uint32_t a = (uint32_t)(r * row_bytes + c * 16);
    *reinterpret_cast<uint4*>(Bs + (a ^ (((a >> 7) & mask) << 4))) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tptr;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N, 0, 0);
    const uint32_t sbo = 8 * row_bytes;
    auto desc = [&](uint32_t addr) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
             ((uint64_t)layout << 61);
    };
    for (int k4 = 0; k4 < kelems / 16; ++k4)
      umma_bf16(td, desc(base + shift * row_bytes + k4 * 32), desc(base + 32 * 1024 + k4 * 32), idesc, k4 > 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid < 128) {
    const int warp = tid >> 5;
    float v[32];
    tmem_ld16(td + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + 16, v + 16);
    tmem_ld_wait();
    for (int j = 0; j < N; ++j) D[tid * N + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(td, 32);
}

int main() {
  std::vector<__nv_bfloat16> hA(ROWS * 64), hB(N * 64);
  std::vector<float> fA(ROWS * 64), fB(N * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(x); fA[i] = x; }
  for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(x); fB[i] = x; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hD(128 * N);
  for (int rel = 0; rel < 2; ++rel)
    for (int bo = 0; bo < 2; ++bo) {
      printf("data swizzle %s, base_offset %s :", rel ? "relative-to-start" : "absolute", bo ? "(start>>7)&7" : "0");
      for (int shift = 0; shift <= 10; ++shift) {
        probe<<<1, 128, 64 * 1024>>>(dA, dB, dD, shift, rel, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < 128; ++i)
          for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < 64; ++k) acc += (double)fA[(i + shift) * 64 + k] * fB[n * 64 + k];
            err = fmax(err, fabs(acc - hD[i * N + n]));
          }
        printf(" s%d:%s", shift, err < 1e-3 ? "OK" : "bad");
      }
      printf("\n");
    }
  cudaFuncSetAttribute(probe_narrow, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rb = 64; rb >= 32; rb /= 2) {
    printf("K-major SWIZZLE_%dB, absolute-address swizzle, row shifts:", rb);
    for (int shift = 0; shift <= 12; ++shift) {
      probe_narrow<<<1, 128, 64 * 1024>>>(dA, dB, dD, shift, rb);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
          double acc = 0;
          for (int k = 0; k < rb / 2; ++k) acc += (double)fA[(i + shift) * 64 + k] * fB[n * 64 + k];
          err = fmax(err, fabs(acc - hD[i * N + n]));
        }
      printf(" s%d:%s", shift, err < 1e-3 ? "OK" : "bad");
    }
    printf("\n");
  }
  return 0;
}
