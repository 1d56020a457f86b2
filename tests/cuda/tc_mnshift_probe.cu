// Hardware probe 2: MN-major operands addressed at a start shifted by whole rows.
//   D[m][n] = sum_k A[k][m] * B[k + shift][n],   A: MN-major SWIZZLE_128B (M = 128, two 64-wide groups),
//   B: MN-major tile whose rows are pixels (K index) holding N channels, row = 128 / 64 / 32 bytes
//   (SWIZZLE_128B / 64B / 32B), data written with the absolute-address swizzle.
// If row shifts work, the taps of a convolution weight gradient become descriptor offsets into ONE
// staged activation tile (no per-tap gather).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../longterm360fov_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int KROWS = 64, BROWS = 96, M = 128;

__global__ void probe(const __nv_bfloat16* Am, const __nv_bfloat16* Bm, float* D, int shift, int row_bytes, int N,
                      int a_shift) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* As = smem;                 // 2 groups x BROWS rows x 128 B (group stride 16 KB)
  uint8_t* Bs = smem + 32 * 1024;     // BROWS rows x row_bytes
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  // A[k][m]: row k, group g = m / 64, chunk c = (m % 64) / 8
  for (int idx = tid; idx < BROWS * 16; idx += blockDim.x) {
    const int r = idx >> 4, c16 = idx & 15, g = c16 >> 3, c = c16 & 7;
    uint4 v = *reinterpret_cast<const uint4*>(Am + r * M + c16 * 8);
    const uint32_t a = (uint32_t)(r * 128 + c * 16);
    *reinterpret_cast<uint4*>(As + g * 16384 + (a ^ (((a >> 7) & 7u) << 4))) = v;
  }
  const int cpr = row_bytes / 16;
  const uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  for (int idx = tid; idx < BROWS * cpr; idx += blockDim.x) {
    const int r = idx / cpr, c = idx % cpr;
    uint4 v = *reinterpret_cast<const uint4*>(Bm + r * 64 + c * 8);
    const uint32_t a = (uint32_t)(r * row_bytes + c * 16);
    *reinterpret_cast<uint4*>(Bs + (a ^ (((a >> 7) & mask) << 4))) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tptr;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16_f32(M, N, 1, 1);
    const uint32_t sbo_b = 8 * row_bytes;
    auto bdesc = [&](uint32_t addr) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((sbo_b >> 4) & 0x3FFFu) << 16) |
             ((uint64_t)((sbo_b >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
    };
    for (int k4 = 0; k4 < KROWS / 16; ++k4) {
      const uint64_t ad = smem_desc_sw128(base + (a_shift + k4 * 16) * 128, 16384, 1024);
      umma_bf16(td, ad, bdesc(base + 32 * 1024 + (shift + k4 * 16) * row_bytes), idesc, k4 > 0);
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
    for (int j = 0; j < 32; ++j) D[tid * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(td, 64);
}

// Probe 3: N spans NGRP groups (one swizzle-atom width each) whose stride LBO is ONE ROW: group g of the B tile is
// the same staged region shifted by g rows, i.e. D[m][g*GW + c] = sum_k A[k][m] * B[k + shift + g][c].
__global__ void probe_lbo(const __nv_bfloat16* Am, const __nv_bfloat16* Bm, float* D, int shift, int row_bytes,
                          int ngrp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* As = smem;
  uint8_t* Bs = smem + 32 * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < BROWS * 16; idx += blockDim.x) {
    const int r = idx >> 4, c16 = idx & 15, g = c16 >> 3, c = c16 & 7;
    uint4 v = *reinterpret_cast<const uint4*>(Am + r * M + c16 * 8);
    const uint32_t a = (uint32_t)(r * 128 + c * 16);
    *reinterpret_cast<uint4*>(As + g * 16384 + (a ^ (((a >> 7) & 7u) << 4))) = v;
  }
  const int cpr = row_bytes / 16;
  const uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  for (int idx = tid; idx < BROWS * cpr; idx += blockDim.x) {
    const int r = idx / cpr, c = idx % cpr;
    uint4 v = *reinterpret_cast<const uint4*>(Bm + r * 64 + c * 8);
    const uint32_t a = (uint32_t)(r * row_bytes + c * 16);
    *reinterpret_cast<uint4*>(Bs + (a ^ (((a >> 7) & mask) << 4))) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tptr;
  const int GW = row_bytes / 2, N = ngrp * GW;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16_f32(M, N, 1, 1);
    const uint32_t sbo_b = 8 * row_bytes, lbo_b = row_bytes;
    auto bdesc = [&](uint32_t addr) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_b >> 4) & 0x3FFFu) << 16) |
             ((uint64_t)((sbo_b >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
    };
    for (int k4 = 0; k4 < KROWS / 16; ++k4) {
      const uint64_t ad = smem_desc_sw128(base + (k4 * 16) * 128, 16384, 1024);
      umma_bf16(td, ad, bdesc(base + 32 * 1024 + (shift + k4 * 16) * row_bytes), idesc, k4 > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid < 128) {
    const int warp = tid >> 5;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[tid * 512 + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(td, 512);
}

int main() {
  std::vector<__nv_bfloat16> hA(BROWS * M), hB(BROWS * 64);
  std::vector<float> fA(BROWS * M), fB(BROWS * 64);
  srand(2);
  for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(x); fA[i] = x; }
  for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(x); fB[i] = x; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hD(128 * 32);
  const int rbs[4] = {128, 128, 64, 32}, ns[4] = {64 > 32 ? 32 : 32, 16, 32, 16};
  for (int v = 0; v < 4; ++v)
    for (int as = 0; as <= 3; as += 3) {
      const int rb = rbs[v], N = ns[v];
      printf("B MN-major SWIZZLE_%dB N=%d, A row shift %d, B row shifts:", rb, N, as);
      for (int shift = 0; shift <= 12; ++shift) {
        probe<<<1, 128, 64 * 1024>>>(dA, dB, dD, shift, rb, N, as);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < KROWS; ++k) acc += (double)fA[(k + as) * M + m] * fB[(k + shift) * 64 + n];
            err = fmax(err, fabs(acc - hD[m * 32 + n]));
          }
        printf(" s%d:%s", shift, err < 1e-3 ? "OK" : "bad");
      }
      printf("\n");
    }
  {
    float* dD2; cudaMalloc(&dD2, 128 * 512 * 4);
    std::vector<float> hD2(128 * 512);
    cudaFuncSetAttribute(probe_lbo, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int rb3[3] = {128, 64, 32}, ng3[3] = {3, 5, 5};
    for (int v = 0; v < 3; ++v) {
      const int rb = rb3[v], ngrp = ng3[v], GW = rb / 2, N = ngrp * GW;
      printf("B MN-major SWIZZLE_%dB, N = %d groups x %d with LBO = one row (%d B), start shifts:", rb, ngrp, GW, rb);
      for (int shift = 0; shift <= 9; ++shift) {
        cudaMemset(dD2, 0, 128 * 512 * 4);
        probe_lbo<<<1, 128, 64 * 1024>>>(dA, dB, dD2, shift, rb, ngrp);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD2.data(), dD2, hD2.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < N; ++n) {
            const int g = n / GW, c = n % GW;
            double acc = 0;
            for (int k = 0; k < KROWS; ++k) acc += (double)fA[k * M + m] * fB[(k + shift + g) * 64 + c];
            err = fmax(err, fabs(acc - hD2[m * 512 + n]));
          }
        printf(" s%d:%s", shift, err < 1e-3 ? "OK" : "bad");
      }
      printf("\n");
    }
  }
  return 0;
}
