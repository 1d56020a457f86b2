// Blackwell (sm_100a) building blocks shared by the tensor-core kernels of libfov360:
// mbarrier, bulk-copy (TMA engine), proxy fences, TMEM allocation, tcgen05.mma / commit / ld
// wrappers and the shared-memory / instruction descriptor encoders.
//
// Operand tiles live in shared memory in the canonical 128-byte-swizzled layouts:
//   K-major  : row r (an M or N index) is a 128-byte line holding 64 bf16 along K; the
//              16-byte chunk c of row r is stored at chunk position (c ^ (r & 7)); groups of
//              8 rows are 1024 B apart (SBO).
//   MN-major : row r (a K index) is a 128-byte line holding 64 bf16 along M/N; same chunk
//              swizzle; groups of 8 K rows are 1024 B apart (SBO) and 64-wide M/N groups are
//              LBO bytes apart.
// Tiles must start on a 1024-byte boundary.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug must fail loudly (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s at 1.9 GHz
      printf("libfov360: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---- proxies / bulk copy ---------------------------------------------------------------
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA-engine bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__host__ __device__ inline uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// ---- descriptors -------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// 32-bit instruction descriptor, kind::f16, bf16 x bf16 -> fp32.
//   [4,6) D fmt = 1 (f32) | [7,10) A fmt = 1 (bf16) | [10,13) B fmt = 1 (bf16) | bit 15 A major | bit 16 B major
//   (0 = K-major, 1 = MN-major) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ inline uint32_t idesc_bf16_f32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn_major & 1) << 15) | ((uint32_t)(b_mn_major & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued tcgen05.mma of this thread complete -> one arrival on the mbarrier
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: the calling warp reads its own 32-lane quarter; thread i gets lane
// (warp%4)*32+i, 16 consecutive fp32 columns starting at the column in taddr.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: thread i writes lane (warp%4)*32+i, 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- bf16 splitting -------------------------------------------------------------------------
// x ~= t0 + t1 + t2 with each term a bf16: 1 term = plain bf16, 2 terms ~ 16 mantissa bits,
// 3 terms ~ 24 bits (fp32).  Returns term `s` of x.
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
template <int NS>
__device__ __forceinline__ void bf16_split(float x, float (&t)[NS]) {
  float r = x;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    t[s] = bf16_round(r);
    r -= t[s];
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// Split four floats into NS bf16 terms, packed as the 8 bytes each term's tile stores:
// one cvt.rn.bf16x2 per pair and term, residuals formed by bit-expanding the packed bf16 back.
template <int NS>
__device__ __forceinline__ void split4(float4 v, uint2 (&out)[NS]) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const uint32_t p01 = pack_bf16x2(v.x, v.y), p23 = pack_bf16x2(v.z, v.w);
    out[s] = make_uint2(p01, p23);
    if (s + 1 < NS) {
      v.x -= __uint_as_float(p01 << 16); v.y -= __uint_as_float(p01 & 0xffff0000u);
      v.z -= __uint_as_float(p23 << 16); v.w -= __uint_as_float(p23 & 0xffff0000u);
    }
  }
}
// 4-element gather with the widest load the tensor's alignment allows (vec = 4, 2 or 1 floats)
__device__ __forceinline__ float4 ldg_vec4(const float* __restrict__ p, int nvalid, int vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec == 4) {
    v = __ldg(reinterpret_cast<const float4*>(p));
  } else if (vec == 2) {
    if (nvalid >= 2) { const float2 a = __ldg(reinterpret_cast<const float2*>(p)); v.x = a.x; v.y = a.y; }
    if (nvalid >= 4) { const float2 b = __ldg(reinterpret_cast<const float2*>(p + 2)); v.z = b.x; v.w = b.y; }
  } else {
    if (nvalid > 0) v.x = __ldg(p);
    if (nvalid > 1) v.y = __ldg(p + 1);
    if (nvalid > 2) v.z = __ldg(p + 2);
    if (nvalid > 3) v.w = __ldg(p + 3);
  }
  return v;
}

// the same gather through L2 only (ld.global.cg): for tensors another kernel is writing WHILE this one runs (layer
// wavefront of the persistent ConvLSTM kernels) - __ldg / L1-cached loads may return a stale line there
__device__ __forceinline__ float4 ldcg_vec4(const float* p, int nvalid, int vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec == 4) {
    v = __ldcg(reinterpret_cast<const float4*>(p));
  } else if (vec == 2) {
    if (nvalid >= 2) { const float2 a = __ldcg(reinterpret_cast<const float2*>(p)); v.x = a.x; v.y = a.y; }
    if (nvalid >= 4) { const float2 b = __ldcg(reinterpret_cast<const float2*>(p + 2)); v.z = b.x; v.w = b.y; }
  } else {
    if (nvalid > 0) v.x = __ldcg(p);
    if (nvalid > 1) v.y = __ldcg(p + 1);
    if (nvalid > 2) v.z = __ldcg(p + 2);
    if (nvalid > 3) v.w = __ldcg(p + 3);
  }
  return v;
}

// ---- layer wavefront: per (image group, timestep) flags in global memory between two concurrently running kernels ----
// producer: every thread that stored data calls wave_publish_arrive (its stores are device-visible before it counts in);
// the last of `n` arrivals raises the flag.  consumer: wave_wait spins (bounded: a broken co-residency assumption ends
// in a trap, not in a hang), then reads the data with ldcg loads.
__device__ __forceinline__ void wave_publish_arrive(int* smem_counter, int n, int* flag) {
  __threadfence();
  if (atomicAdd_block(smem_counter, 1) == n - 1) {
    *smem_counter = 0;                 // the next step's arrivals are ordered behind this one by the kernel's own barriers
    __threadfence();
    *reinterpret_cast<volatile int*>(flag) = 1;
  }
}
__device__ __forceinline__ void wave_wait(const int* flag) {
  const volatile int* f = reinterpret_cast<const volatile int*>(flag);
  unsigned spins = 0;
  while (*f == 0) {
    __nanosleep(40);
    if (++spins > (1u << 23)) __trap();          // ~0.5 s: the producer kernel never ran beside this one
  }
  __threadfence();
}

// ---- shared by the shifted-tap kernels -------------------------------------------------------
// exp-based activations for the fused epilogue (abs error ~1e-7, far inside the bf16-split budget)
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float fast_rec(int rec, float x) {
  if (rec == 0 /* FOV_REC_HARD_SIGMOID */) return fminf(fmaxf(0.2f * x + 0.5f, 0.0f), 1.0f);
  return __fdividef(1.0f, 1.0f + __expf(-x));
}

// descriptor = high word (SBO, version 1, swizzle mode) | low word (start address, LBO = 16 B)
constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
// same with an explicit leading-dimension byte offset (MN-major tiles: stride between the N / M groups)
__device__ __forceinline__ uint64_t desc_at_lbo(uint32_t hi, uint32_t saddr, uint32_t lbo_bytes) {
  return ((uint64_t)hi << 32) | (uint64_t)(((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16));
}
__device__ __forceinline__ uint64_t desc_at(uint32_t hi, uint32_t saddr) {
  return ((uint64_t)hi << 32) | (uint64_t)(((saddr >> 4) & 0x3FFFu) | (1u << 16));
}


}  // namespace tc
