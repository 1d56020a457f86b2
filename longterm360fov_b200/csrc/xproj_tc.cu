// Time-batched fc-LSTM input projection on tensor cores, fed by TMA (north_star kernel 1).
//
//   P[(b,t)][0..4H) = x[b,t,:] . W          W = the Keras LSTM `kernel` (in, 4H), gate blocks i,f,c,o
//
// for EVERY timestep of a phase whose inputs are known in advance (the encoder, a teacher-forced decoder:
// mycode/FoV_seq2seq.py:82-95), in one launch, so that the persistent recurrent kernel (lstm_seq2seq_tc.cu) only
// carries the K = 64 recurrent GEMM per step and adds P in its gate epilogue.  That is what lets the 90-wide raw
// encoder input of FoV_seq2seq run on tcgen05 at all: [W ; U] for in = 90 at two bf16 terms plus the operand tiles
// exceed the 227 KB of shared memory, U alone does not.
//
// Data path
//   * A tile = 128 GEMM rows = 64 sequences x 2 consecutive timesteps.  x (B,T,in) fp32 is viewed as the tensor
//     (2*in, T/2, B) whose strides (2*in*4, T*in*4 bytes) are multiples of 16 bytes although a single row (in*4 bytes:
//     360 for in = 90, 24 for in = 6) is not - ONE cp.async.bulk.tensor.3d (SASS UTMALDG) with box (2*in, 1, 64)
//     lands the 128 rows contiguously in shared memory; sequences beyond B are zero-filled by the TMA unit.
//   * the fp32 rows are split in shared memory into bf16 terms in the K-major 128-byte (64-byte for a 32-wide tail
//     chunk) swizzled operand layout; W is converted once per CTA the same way (B operand, N = 4H = 256).
//   * tcgen05.mma, M = 128, N = 256, fp32 accumulator in TMEM; all significant cross products of the terms.
//   * epilogue: the thread of a row reads its 256 columns with tcgen05.ld and stores float4s into the TILED
//     workspace  P[b/128][t][col/4][b%128][4]  - the consumer kernel's thread = sequence mapping then reads
//     512 contiguous bytes per warp instruction, and this kernel's stores are 256-byte contiguous runs.
//   * training: the x part of the saved [h | x | 0] rows (fov_lstm_saved.xh) is written from the staged tile,
//     coalesced, so the recurrent kernel never touches x.
// Persistent: grid = min(tiles, SMs); the TMA load of tile i+1 is in flight during the MMAs and the epilogue of tile i.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kG = 256;          // 4H
constexpr int kRows = 128;       // GEMM rows per tile
constexpr int kSeq = 64;         // sequences per tile
constexpr int kTP = 2;           // timesteps per tile
constexpr int kThr = 256;

struct XpParams {
  int B, T, in, in_p, K_xh;      // K_xh: row width of the saved [h | x | 0] tensor (0: not written)
  int nchunk, cw[2], koff[2];    // K chunks: widths 64 / 32 / 16 and first k of each
  uint32_t a_off[2], w_off[2], a_term[2], w_term[2], row_bytes[2], swz[2], desc_hi[2];
  uint32_t stage_off, stage_bytes, data_bytes;
  int ntiles, tiles_per_seqblk;
  const float* W;                // (in, 256)
  float* P;                      // tiled workspace
  float* xh;                     // optional (B,T,K_xh)
};

struct XpBook {
  uint64_t full, mma_done;
  uint32_t tmem_ptr;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(kThr, 1) xproj_tc_kernel(const __grid_constant__ CUtensorMap xmap, const XpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  XpBook* bk = reinterpret_cast<XpBook*>(smem + p.data_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&xmap);
    mbar_init(smem_u32(&bk->full), 1);
    mbar_init(smem_u32(&bk->mma_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), kG);
    tmem_relinquish();
  }
  __syncthreads();
  // first tile in flight while the weights are converted
  auto issue_load = [&](int i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int sb = tile / p.tiles_per_seqblk, j = tile - sb * p.tiles_per_seqblk;
    const uint32_t bar = smem_u32(&bk->full);
    mbar_arrive_expect_tx(bar, p.stage_bytes);
    tma_load_3d(base + p.stage_off, &xmap, 0, j, sb * kSeq, bar);
  };
  if (tid == 0 && my_tiles > 0) issue_load(0);

  // ---- W (in, 256) -> bf16 terms, K-major swizzled rows (row n = gate column n), per K chunk ----
  for (int c = 0; c < p.nchunk; ++c) {
    const int c8n = p.cw[c] / 8;
    for (int idx = tid; idx < kG * c8n; idx += kThr) {
      const int n = idx % kG, c8 = idx / kG;          // consecutive threads = consecutive n: coalesced reads of W rows
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = p.koff[c] + c8 * 8 + j;
        v[j] = k < p.in ? __ldg(&p.W[(size_t)k * kG + n]) : 0.0f;
      }
      uint2 lo[NS], hi[NS];
      split4<NS>(make_float4(v[0], v[1], v[2], v[3]), lo);
      split4<NS>(make_float4(v[4], v[5], v[6], v[7]), hi);
      const uint32_t a0 = (uint32_t)n * p.row_bytes[c] + (uint32_t)c8 * 16u;
      const uint32_t so = a0 ^ (((a0 >> 7) & p.swz[c]) << 4);
#pragma unroll
      for (int s = 0; s < NS; ++s)
        *reinterpret_cast<uint4*>(smem + p.w_off[c] + s * p.w_term[c] + so) = make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;
  const float* stage = reinterpret_cast<const float*>(smem + p.stage_off);
  const uint32_t idesc = idesc_bf16_f32(kRows, kG, 0, 0);

  for (int i = 0; i < my_tiles; ++i) {
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int sb = tile / p.tiles_per_seqblk, j = tile - sb * p.tiles_per_seqblk;
    const int b0 = sb * kSeq, t0 = j * kTP;
    mbar_wait(smem_u32(&bk->full), (uint32_t)i & 1u);
    // ---- staged fp32 rows -> bf16 terms in the A operand tiles: (row, 8-wide k group) items ----
    const int c8tot = p.in_p / 8;
    for (int idx = tid; idx < kRows * c8tot; idx += kThr) {
      const int row = idx % kRows, c8g = idx / kRows;          // consecutive threads = consecutive rows
      const int k0 = c8g * 8;
      const int c = (p.nchunk > 1 && k0 >= p.koff[1]) ? 1 : 0;
      const int c8 = (k0 - p.koff[c]) >> 3;
      const float* src = stage + row * p.in + k0;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (k0 + q < p.in) ? src[q] : 0.0f;
      uint2 lo[NS], hi[NS];
      split4<NS>(make_float4(v[0], v[1], v[2], v[3]), lo);
      split4<NS>(make_float4(v[4], v[5], v[6], v[7]), hi);
      const uint32_t a0 = (uint32_t)row * p.row_bytes[c] + (uint32_t)c8 * 16u;
      const uint32_t so = a0 ^ (((a0 >> 7) & p.swz[c]) << 4);
#pragma unroll
      for (int s = 0; s < NS; ++s)
        *reinterpret_cast<uint4*>(smem + p.a_off[c] + s * p.a_term[c] + so) = make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
    }
    // ---- training: x part (+ zero pad) of the saved [h | x | 0] rows, one warp per row, coalesced ----
    if (p.xh) {
      const int xw = p.K_xh - 64;                               // x columns incl. the pad
      for (int row = warp; row < kRows; row += kThr / 32) {
        const int b = b0 + (row >> 1), t = t0 + (row & 1);
        if (b < p.B) {
          float* dst = p.xh + ((size_t)b * p.T + t) * p.K_xh + 64;
          for (int k = lane; k < xw; k += 32) dst[k] = k < p.in ? stage[row * p.in + k] : 0.0f;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      if (i + 1 < my_tiles) issue_load(i + 1);                  // the staging buffer is free again
      tc_fence_after();
      uint32_t acc = 0;
      for (int c = 0; c < p.nchunk; ++c) {
        for (int k16 = 0; k16 < p.cw[c] / 16; ++k16) {
#pragma unroll
          for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
            for (int sa = 0; sa <= sum; ++sa) {
              const int sb_ = sum - sa;
              umma_bf16(tmem_d, desc_at(p.desc_hi[c], base + p.a_off[c] + sa * p.a_term[c] + k16 * 32),
                        desc_at(p.desc_hi[c], base + p.w_off[c] + sb_ * p.w_term[c] + k16 * 32), idesc, acc);
              acc = 1;
            }
          }
        }
      }
      umma_commit(smem_u32(&bk->mma_done));
    }
    mbar_wait(smem_u32(&bk->mma_done), (uint32_t)i & 1u);
    tc_fence_after();
    // ---- epilogue: row m = (sequence m/2, timestep m%2); warp w reads TMEM lanes 32*(w&3).., column half w>>2 ----
    {
      const int q = warp & 3, half = warp >> 2;
      const int m = q * 32 + lane;
      const int b = b0 + (m >> 1), t = t0 + (m & 1);
      const uint32_t t_row = tmem_d + ((uint32_t)(q * 32) << 16);
      float* dst = p.P + ((((size_t)(b >> 7) * p.T + t) * 64) * 128 + (size_t)(b & 127)) * 4;
#pragma unroll 2
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        float v[32];
        tmem_ld16(t_row + c0, v);
        tmem_ld16(t_row + c0 + 16, v + 16);
        tmem_ld_wait();
        if (b < p.B) {
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4)
            *reinterpret_cast<float4*>(dst + (size_t)(c0 / 4 + g4) * 512) =
                make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                            // accumulator and A tiles free for the next tile
    tc_fence_after();
  }
  if (warp == 1) tmem_dealloc(tmem_d, kG);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)sym;
  }
  return fn;
}

int xp_plan(int B, int T, int in, int NS, XpParams* out, size_t* smem) {
  XpParams p{};
  p.B = B; p.T = T; p.in = in;
  p.in_p = (in + 15) / 16 * 16;
  if (in <= 0 || p.in_p > 128 || T % kTP != 0 || (kTP * in) % 4 != 0 || kTP * in > 256) return FOV_ERR_UNSUPPORTED;
  if (p.in_p <= 64) {
    p.nchunk = 1; p.cw[0] = p.in_p <= 16 ? 16 : (p.in_p <= 32 ? 32 : 64); p.koff[0] = 0;
  } else {
    const int rem = p.in_p - 64;
    p.nchunk = 2; p.cw[0] = 64; p.koff[0] = 0; p.cw[1] = rem <= 16 ? 16 : (rem <= 32 ? 32 : 64); p.koff[1] = 64;
  }
  p.in_p = p.koff[p.nchunk - 1] + p.cw[p.nchunk - 1];
  uint32_t off = 0;
  for (int c = 0; c < p.nchunk; ++c) {
    p.row_bytes[c] = (uint32_t)p.cw[c] * 2u;
    p.swz[c] = p.row_bytes[c] == 128 ? 7u : (p.row_bytes[c] == 64 ? 3u : 1u);
    const uint32_t layout = p.row_bytes[c] == 128 ? 2u : (p.row_bytes[c] == 64 ? 4u : 6u);
    p.desc_hi[c] = ((8u * p.row_bytes[c]) >> 4) | (1u << 14) | (layout << 29);
    p.w_term[c] = (uint32_t)kG * p.row_bytes[c];
    p.w_off[c] = off; off += NS * p.w_term[c];
  }
  for (int c = 0; c < p.nchunk; ++c) {
    p.a_term[c] = (uint32_t)kRows * p.row_bytes[c];
    p.a_off[c] = off; off += NS * p.a_term[c];
  }
  p.stage_off = (off + 1023u) & ~1023u;
  p.stage_bytes = (uint32_t)(kRows * in * 4);
  p.data_bytes = (p.stage_off + p.stage_bytes + 1023u) & ~1023u;
  *smem = (size_t)p.data_bytes + sizeof(XpBook) + 1024;
  if (*smem > 227 * 1024) return FOV_ERR_UNSUPPORTED;
  p.tiles_per_seqblk = T / kTP;
  p.ntiles = ((B + kSeq - 1) / kSeq) * p.tiles_per_seqblk;
  *out = p;
  return FOV_OK;
}

template <int NS>
int xp_launch(const CUtensorMap& map, const XpParams& p, size_t smem, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(xproj_tc_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      fov_set_error("xproj: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  const int grid = p.ntiles < fov_num_sms() ? p.ntiles : fov_num_sms();
  xproj_tc_kernel<NS><<<grid, kThr, smem, st>>>(map, p);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

size_t lstm_xproj_ws_floats(int B, int T) { return (size_t)((B + 127) / 128) * 128 * (size_t)T * kG; }

bool lstm_xproj_supported(int B, int T, int in, int math, const float* x) {
  XpParams p;
  size_t smem;
  if (math < FOV_MATH_BF16 || math > FOV_MATH_BF16X3) return false;
  if (x && (uintptr_t)x % 16 != 0) return false;
  if ((long long)B * T * in >= (1LL << 31)) return false;
  return xp_plan(B, T, in, math, &p, &smem) == FOV_OK && get_encode() != nullptr;
}

// P (tiled, lstm_xproj_ws_floats floats) = x (B,T,in) . W (in,256); xh: optional saved [h | x | 0] rows of width K_xh
int lstm_xproj_run(int B, int T, int in, int math, const float* x, const float* W, float* P, float* xh, int K_xh,
                   cudaStream_t st) {
  XpParams p;
  size_t smem;
  int rc = xp_plan(B, T, in, math, &p, &smem);
  if (rc) { fov_set_error("lstm_xproj: unsupported shape (B=%d T=%d in=%d)", B, T, in); return rc; }
  FOV_CHECK_ARG(x && W && P && (uintptr_t)x % 16 == 0 && (uintptr_t)P % 16 == 0, "bad / unaligned pointers");
  p.W = W; p.P = P; p.xh = xh; p.K_xh = xh ? K_xh : 0;
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { fov_set_error("lstm_xproj: cuTensorMapEncodeTiled is not available"); return FOV_ERR_UNSUPPORTED; }
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)(kTP * in), (cuuint64_t)(T / kTP), (cuuint64_t)B};
  const cuuint64_t gstr[2] = {(cuuint64_t)(kTP * in) * 4, (cuuint64_t)T * in * 4};
  const cuuint32_t box[3] = {(cuuint32_t)(kTP * in), 1u, (cuuint32_t)kSeq};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { fov_set_error("lstm_xproj: cuTensorMapEncodeTiled failed (%d)", (int)cr); return FOV_ERR_CUDA; }
  switch (math) {
    case 1: return xp_launch<1>(map, p, smem, st);
    case 2: return xp_launch<2>(map, p, smem, st);
    default: return xp_launch<3>(map, p, smem, st);
  }
}
