// Tensor-core (tcgen05 + TMEM) weight gradient of the convolution family for sm_100a.
//
//   gw[k][n] += sum_m A(m,k) * dY[m][n]        k = (tap, channel), m = pixel, n = output channel
//   gbias[n] += sum_m dY[m][n]
//
// As a tcgen05 GEMM the reduction index is the pixel: D[128 k][BLOCK_N n] += A'[k][pixel] B'[n][pixel].
// Both operands are contiguous along their M/N index in HBM (channels-last), so they are staged
// as MN-major 128-byte-swizzled tiles: one shared-memory row per pixel, 64 bf16 along k (or n).
// 256 producer threads gather 64 pixels per stage (coalesced float4 along channels), split every
// value into NS bf16 terms and keep the loads of the next pixel block in flight while the
// current one is converted; one thread issues the MMAs; the pixel range is split across
// blockIdx.z and the partial tiles are added to gw with coalesced red.global.add.
// The bias gradient is a by-product of the dY gather (per-thread column sums).
//
// Replaces the weight-gradient ops TF1 generates for Dense / Conv2D / ConvLSTM2D
// (mycode/others_LSTM_span_whole.py:88-109, mycode/convlstm_seq2seq.py:100-126,175-181).
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int WG_M = 128;            // k rows per CTA (MMA M)
constexpr int WG_PIX_MAX = 64;       // pixels per pipeline stage: 64 (n tiles <= 128 wide) or 32 (256-wide n tiles)
constexpr int kProd = 512;           // producer threads (warps 0-15)
constexpr int kTabWarp = 16, kMmaWarp = 17;
constexpr int kThr = 576;            // + table warp (16) + MMA warp (17)
constexpr int NBMAX = 4;             // dy float4 loads per thread per pixel block (x rows: PIX / 16)
constexpr int kStagesMax = 3;
constexpr int kPfDist = 4;           // L2 prefetch distance of the 1x1-layer operand rows, in 64-pixel blocks
constexpr int RS = 36;

struct WgParams {
  const float* x; long long x_outer, x_inner; int x_pix_stride, Cin, Cin_p, kw, taps, dil_h, dil_w, pad_h, pad_w, x_vec;
  const float* dy; long long dy_outer, dy_inner; int dy_pix_stride, Cout, dy_vec;
  int H, W, HW, T_inner, M;
  int BLOCK_N, tmem_cols, stages, stage_bytes, b_term_bytes, data_bytes;
  int nblocks, blocks_per_split, dbg;
  float* gw; float* gbias;
};

struct RowTab {
  const float* xp[WG_PIX_MAX];    // address of the row's own pixel in x (channel 0)
  const float* dyp[WG_PIX_MAX];   // address of the row's pixel in dy (channel 0)
  int pyx[WG_PIX_MAX];            // (y << 16) | x; y = 0x4000 marks a row past the end
};
struct Book {
  RowTab tab[kStagesMax];
  int dstrow[2 * WG_M];             // row of gw each accumulator row adds to, or -1 (two M tiles at most)
  uint64_t full[kStagesMax], empty[kStagesMax], tabrdy[kStagesMax], tmem_full;
  uint32_t tmem_ptr;
};

// loads of out-of-range taps / rows are redirected here, so the gather needs no predicated loads
__device__ float4 g_zero_page[4];

// diagnostics: cycles CTA (0,0,0) spent per producer phase / in the MMA thread (fov_debug_wgrad_read)
__device__ unsigned long long g_wg_timeline[8];

// MT = M tiles (128 k rows each) per CTA: with two, dY is staged once for twice the k rows (half the dY re-reads)
template <int NS, bool FAST, int WG_PIX, int MT>
__global__ void __launch_bounds__(kThr, 1) tc_wgrad_kernel(const WgParams p) {
  constexpr int WG_GRP = WG_PIX * 128;   // bytes of one 64-wide MN group (WG_PIX pixel rows x 128 B)
  constexpr int NA = WG_PIX * MT / 16;   // x float4 loads per thread per pixel block
  constexpr int KT = WG_M * MT;          // k rows of this CTA
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  Book* bk = reinterpret_cast<Book*>(smem + p.data_bytes);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int k_tile = blockIdx.x, n0 = blockIdx.y * p.BLOCK_N;
  const int S = p.stages;
  const int blk0 = blockIdx.z * p.blocks_per_split;
  int nblk = p.nblocks - blk0;
  if (nblk > p.blocks_per_split) nblk = p.blocks_per_split;
  constexpr int A_TERM = 2 * MT * WG_GRP;   // KT k values = 2 MT 64-wide groups

  if (tid < KT) {
    const int k = k_tile * KT + tid;
    const int tap = k / p.Cin_p, ci = k - tap * p.Cin_p;
    bk->dstrow[tid] = (tap < p.taps && ci < p.Cin) ? tap * p.Cin + ci : -1;
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bk->full[s]), kProd);
      mbar_init(smem_u32(&bk->empty[s]), 1);
      mbar_init(smem_u32(&bk->tabrdy[s]), 1);
    }
    mbar_init(smem_u32(&bk->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == kTabWarp) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int f4 = p.BLOCK_N >> 2;          // float4 per dY row of this n tile (4..32)
  const int nb = (WG_PIX * f4 + kProd - 1) / kProd;   // dY loads per thread per block (1..4)
  const int b_c4 = tid % f4;              // fixed float4 column of this thread
  const int b_row0 = tid / f4, b_rstep = kProd / f4;

  if (warp < kProd / 32) {
    // ---------------- producers ----------------
    const int q = tid % (32 * MT), rsub = tid / (32 * MT);   // A': float4 slot along k, row phase (0..16/MT-1)
    constexpr int ASTEP = 16 / MT;                            // rows between a thread's x loads
    const int k = k_tile * KT + q * 4;
    const int tap = k / p.Cin_p, ci = k - tap * p.Cin_p;
    int a_nvalid = (tap < p.taps) ? (p.Cin - ci) : 0;
    a_nvalid = a_nvalid < 0 ? 0 : (a_nvalid > 4 ? 4 : a_nvalid);
    const int ty = tap / p.kw, tx = tap - ty * p.kw;
    const int tdy = ty * p.dil_h - p.pad_h, tdx = tx * p.dil_w - p.pad_w;
    const int a_delta = (tdy * p.W + tdx) * p.x_pix_stride + ci;
    // rows are 16*j + rsub (A') and b_row0 + j*b_rstep with b_rstep % 8 == 0 (B'): constant swizzle phase
    const uint32_t a_off = (uint32_t)(q >> 4) * WG_GRP + (uint32_t)(q & 1) * 8u + (uint32_t)rsub * 128u +
                           ((uint32_t)(((q & 15) >> 1) ^ (rsub & 7)) << 4);
    const int bcol = n0 + b_c4 * 4;
    int b_nvalid = p.Cout - bcol;
    b_nvalid = b_nvalid < 0 ? 0 : (b_nvalid > 4 ? 4 : b_nvalid);
    const uint32_t b_off = (uint32_t)(b_c4 >> 4) * WG_GRP + (uint32_t)(b_c4 & 1) * 8u + (uint32_t)b_row0 * 128u +
                           ((uint32_t)(((b_c4 & 15) >> 1) ^ (b_row0 & 7)) << 4);

    // FAST (every tensor 16-byte aligned): unconditional float4 loads, out-of-range ones redirected to a
    // zero page, so all loads of a block are in flight at once and nothing is merged after them
    const float* zp = reinterpret_cast<const float*>(g_zero_page);
    auto issue = [&](int stage, float4 (&va)[NA], float4 (&vb)[NBMAX]) {
      const RowTab& t = bk->tab[stage];
      if (FAST) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          const int r = j * ASTEP + rsub;
          const int pyx = t.pyx[r];
          const unsigned yy = (unsigned)((pyx >> 16) + tdy), xx = (unsigned)((pyx & 0xffff) + tdx);
          const bool ok = a_nvalid > 0 && yy < (unsigned)p.H && xx < (unsigned)p.W;
          va[j] = __ldg(reinterpret_cast<const float4*>(ok ? t.xp[r] + a_delta : zp));
        }
#pragma unroll
        for (int j = 0; j < NBMAX; ++j) {
          const int r = b_row0 + j * b_rstep;
          const bool ok = j < nb && r < WG_PIX && b_nvalid > 0 && (t.pyx[r & (WG_PIX - 1)] >> 16) != 0x4000;
          vb[j] = __ldg(reinterpret_cast<const float4*>(ok ? t.dyp[r & (WG_PIX - 1)] + bcol : zp));
        }
        return;
      }
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const int r = j * ASTEP + rsub;
        const int pyx = t.pyx[r];
        const unsigned yy = (unsigned)((pyx >> 16) + tdy), xx = (unsigned)((pyx & 0xffff) + tdx);
        va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_nvalid > 0 && yy < (unsigned)p.H && xx < (unsigned)p.W)
          va[j] = ldg_vec4(t.xp[r] + a_delta, a_nvalid, p.x_vec);
      }
#pragma unroll
      for (int j = 0; j < NBMAX; ++j) {
        vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nb && b_row0 + j * b_rstep < WG_PIX) {
          const int r = b_row0 + j * b_rstep;
          if (b_nvalid > 0 && (t.pyx[r] >> 16) != 0x4000) vb[j] = ldg_vec4(t.dyp[r] + bcol, b_nvalid, p.dy_vec);
        }
      }
    };
    auto store = [&](int stage, const float4 (&va)[NA], const float4 (&vb)[NBMAX]) {
      uint8_t* a_tile = smem + (size_t)stage * p.stage_bytes;
      uint8_t* b_tile = a_tile + NS * A_TERM;
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        uint2 pk[NS];
        split4<NS>(va[j], pk);
#pragma unroll
        for (int s = 0; s < NS; ++s) *reinterpret_cast<uint2*>(a_tile + s * A_TERM + j * (ASTEP * 128) + a_off) = pk[s];
      }
#pragma unroll
      for (int j = 0; j < NBMAX; ++j) {
        if (j < nb && b_row0 + j * b_rstep < WG_PIX) {
          uint2 pk[NS];
          split4<NS>(vb[j], pk);
#pragma unroll
          for (int s = 0; s < NS; ++s)
            *reinterpret_cast<uint2*>(b_tile + s * p.b_term_bytes + j * b_rstep * 128 + b_off) = pk[s];
          bsum.x += vb[j].x; bsum.y += vb[j].y; bsum.z += vb[j].z; bsum.w += vb[j].w;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bk->full[stage]));
    };
    float4 a0[NA], b0[NBMAX], a1[NA], b1[NBMAX];
    mbar_wait(smem_u32(&bk->tabrdy[0]), 0);
    issue(0, a0, b0);
    for (int i = 0; i < nblk; i += 2) {
      // block i lives in (a0,b0); prefetch block i+1 into (a1,b1), then the roles swap
      if (i + 1 < nblk) {
        const int s1 = (i + 1) % S;
        mbar_wait(smem_u32(&bk->tabrdy[s1]), (uint32_t)((i + 1) / S) & 1u);
        issue(s1, a1, b1);
      }
      store(i % S, a0, b0);
      if (i + 1 < nblk) {
        if (i + 2 < nblk) {
          const int s2 = (i + 2) % S;
          mbar_wait(smem_u32(&bk->tabrdy[s2]), (uint32_t)((i + 2) / S) & 1u);
          issue(s2, a0, b0);
        }
        store((i + 1) % S, a1, b1);
      }
    }

  } else if (warp == kTabWarp) {
    // ---------------- row-table builder: pixel -> (addresses, y, x) for each stage ----------------
    // lane owns rows lane and lane+32; their pixel advances by 64 per block, tracked with carries
    // (one division set at the start, none in the loop).
    const int dq = WG_PIX / p.HW, dr = WG_PIX - dq * p.HW;       // 64 pixels = dq images + dr pixels
    const int dry = dr / p.W, drx = dr - dry * p.W;
    const int dno = dq / p.T_inner, dni = dq - dno * p.T_inner;
    constexpr int RPL = WG_PIX / 32;                            // rows per lane
    int no[RPL], ni[RPL], py[RPL], px[RPL];
    bool live[RPL];
#pragma unroll
    for (int h = 0; h < RPL; ++h) {
      const int m = blk0 * WG_PIX + lane + h * 32;               // may exceed M: rows past the end stay dead
      const int n = m / p.HW, pix = m - n * p.HW;
      no[h] = n / p.T_inner; ni[h] = n - no[h] * p.T_inner;
      py[h] = pix / p.W; px[h] = pix - py[h] * p.W;
    }
    for (int i = 0; i < nblk; ++i) {
      const int stage = i % S;
      mbar_wait(smem_u32(&bk->empty[stage]), ((uint32_t)(i / S) & 1u) ^ 1u);
      RowTab& t = bk->tab[stage];
#pragma unroll
      for (int h = 0; h < RPL; ++h) {
        const int r = lane + h * 32;
        live[h] = (blk0 + i) * WG_PIX + r < p.M;
        if (live[h]) {
          const int pix = py[h] * p.W + px[h];
          t.pyx[r] = (py[h] << 16) | px[h];
          t.xp[r] = p.x + (long long)no[h] * p.x_outer + (long long)ni[h] * p.x_inner + (long long)pix * p.x_pix_stride;
          t.dyp[r] = p.dy + (long long)no[h] * p.dy_outer + (long long)ni[h] * p.dy_inner +
                     (long long)pix * p.dy_pix_stride;
        } else {
          t.pyx[r] = 0x4000 << 16; t.xp[r] = nullptr; t.dyp[r] = nullptr;
        }
        // advance this row by 64 pixels
        px[h] += drx; py[h] += dry;
        if (px[h] >= p.W) { px[h] -= p.W; ++py[h]; }
        int carry = 0;
        if (py[h] >= p.H) { py[h] -= p.H; carry = 1; }
        ni[h] += dni + carry; no[h] += dno;
        if (ni[h] >= p.T_inner) { ni[h] -= p.T_inner; ++no[h]; }
        if (ni[h] >= p.T_inner) { ni[h] -= p.T_inner; ++no[h]; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bk->tabrdy[stage]));
      // L2 prefetch of the operand rows kPfDist blocks ahead (1x1 layers: one contiguous row segment per pixel).  The
      // producers keep only ONE block of loads in flight in registers; with the rows already in L2 that is enough.
      if (p.taps == 1 && i + kPfDist < nblk) {
#pragma unroll
        for (int h = 0; h < RPL; ++h) {
          const long long m = (long long)(blk0 + i + kPfDist) * WG_PIX + lane + h * 32;
          if (m < p.M) {
            const int n = (int)(m / p.HW), pix = (int)(m - (long long)n * p.HW);
            const int nno = n / p.T_inner, nni = n - nno * p.T_inner;
            const float* xr = p.x + (long long)nno * p.x_outer + (long long)nni * p.x_inner + (long long)pix * p.x_pix_stride +
                              k_tile * KT;
            const float* dr_ = p.dy + (long long)nno * p.dy_outer + (long long)nni * p.dy_inner +
                               (long long)pix * p.dy_pix_stride + n0;
            const int xa = p.Cin - k_tile * KT, da = p.Cout - n0;
#pragma unroll
            for (int l = 0; l < KT / 32; ++l)
              if (l * 32 < xa) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr + l * 32));
            for (int l = 0; l * 32 < p.BLOCK_N; ++l)
              if (l * 32 < da) asm volatile("prefetch.global.L2 [%0];" ::"l"(dr_ + l * 32));
          }
        }
      }
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(WG_M, p.BLOCK_N, 1, 1);
      const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long t_full = 0;
      const long long t_begin = clock64();
      if (dbg) g_wg_timeline[3] = (unsigned long long)nblk;
      for (int i = 0; i < nblk; ++i) {
        const int stage = i % S;
        const long long tw = clock64();
        mbar_wait(smem_u32(&bk->full[stage]), (uint32_t)(i / S) & 1u);
        t_full += clock64() - tw;
        tc_fence_after();
        const uint32_t a_base = base + (uint32_t)stage * p.stage_bytes;
        const uint32_t b_base = a_base + NS * A_TERM;
#pragma unroll
        for (int k4 = 0; k4 < WG_PIX / 16; ++k4) {
#pragma unroll
          for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
            for (int sa = 0; sa <= sum; ++sa) {
              const int sb = sum - sa;
              // 16 pixel rows per K step = 2048 B; LBO = stride between 64-wide M/N groups, SBO = 8 pixel rows
              const uint64_t bd = smem_desc_sw128(b_base + sb * p.b_term_bytes + k4 * 2048, WG_GRP, 1024);
              const uint32_t acc = (i > 0 || k4 > 0 || sum != NS - 1 || sa > 0) ? 1u : 0u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t ad = smem_desc_sw128(a_base + sa * A_TERM + mt * 2 * WG_GRP + k4 * 2048, WG_GRP, 1024);
                umma_bf16(tmem_d + (uint32_t)(mt * p.BLOCK_N), ad, bd, idesc, acc);
              }
            }
          }
        }
        umma_commit(smem_u32(&bk->empty[stage]));
      }
      umma_commit(smem_u32(&bk->tmem_full));
      if (dbg) { g_wg_timeline[4] = (unsigned long long)t_full; g_wg_timeline[5] = (unsigned long long)(clock64() - t_begin); }
    }
  }

  // ---------------- epilogue: warps 0-3 add the partial tile into gw ----------------
  if (warp < 4 && p.gw) {
    mbar_wait(smem_u32(&bk->tmem_full), 0);
    tc_fence_after();
    const int r0 = warp * 32;
    const uint32_t t_row = tmem_d + ((uint32_t)r0 << 16);
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * RS);
    for (int mt = 0; mt < MT; ++mt)
    for (int c0 = 0; c0 < p.BLOCK_N; c0 += 32) {
      float v[32];
      tmem_ld16(t_row + mt * p.BLOCK_N + c0, v);
      tmem_ld16(t_row + mt * p.BLOCK_N + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(&stg[lane * RS + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      const int col = n0 + c0 + lane;
      const bool col_ok = (c0 + lane < p.BLOCK_N) && (col < p.Cout);
      for (int rr = 0; rr < 32; ++rr) {
        const int dr = bk->dstrow[mt * WG_M + r0 + rr];
        if (dr >= 0 && col_ok) atomicAdd(p.gw + (long long)dr * p.Cout + col, stg[rr * RS + lane]);
      }
      __syncwarp();
    }
  }
  if (warp < kProd / 32 && p.gbias && k_tile == 0) {
    const int bcol = n0 + b_c4 * 4;
    if (bcol < p.Cout) atomicAdd(p.gbias + bcol, bsum.x);
    if (bcol + 1 < p.Cout) atomicAdd(p.gbias + bcol + 1, bsum.y);
    if (bcol + 2 < p.Cout) atomicAdd(p.gbias + bcol + 2, bsum.z);
    if (bcol + 3 < p.Cout) atomicAdd(p.gbias + bcol + 3, bsum.w);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTabWarp) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

int vec_of(const void* ptr, long long a, long long b, long long c, long long d) {
  auto al = [&](long long m) {
    return ((uintptr_t)ptr % (4 * m) == 0) && a % m == 0 && b % m == 0 && c % m == 0 && d % m == 0;
  };
  if (al(4)) return 4;
  if (al(2)) return 2;
  return 1;
}

template <int NS, bool FAST, int PIX, int MT>
int launch(const WgParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<NS, FAST, PIX, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      fov_set_error("tc_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  tc_wgrad_kernel<NS, FAST, PIX, MT><<<grid, kThr, smem, st>>>(p);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

static int g_wg_debug = 0, g_wg_narrow = 0, g_wg_single_m = 0;
extern "C" void fov_debug_wgrad_single_m(int on) { g_wg_single_m = on; }
extern "C" void fov_debug_wgrad_narrow(int on) { g_wg_narrow = on; }
extern "C" void fov_debug_wgrad_enable(int on) { g_wg_debug = on; }
extern "C" int fov_debug_wgrad_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_wg_timeline, sizeof(unsigned long long) * 8);
}

int tc_wgrad_run(const TcWgrad& c, cudaStream_t st) {
  FOV_CHECK_ARG(c.math >= 1 && c.math <= 3, "math must be 1..3 bf16 terms");
  FOV_CHECK_ARG(c.x && c.dy && (c.gw || c.gbias), "NULL pointer");
  FOV_CHECK_ARG(c.N_img > 0 && c.H > 0 && c.W > 0 && c.Cin > 0 && c.Cout > 0 && c.T_inner > 0, "bad shape");
  const long long M = (long long)c.N_img * c.H * c.W;
  FOV_CHECK_ARG(M < (1LL << 31) - WG_PIX_MAX, "too many pixels for 32-bit indexing");
  WgParams p{};
  p.x = c.x; p.x_outer = c.x_outer; p.x_inner = c.x_inner; p.x_pix_stride = c.x_pix_stride;
  p.Cin = c.Cin; p.Cin_p = (c.Cin + 7) / 8 * 8; p.kw = c.kw; p.taps = c.kh * c.kw;
  p.dil_h = c.dil_h; p.dil_w = c.dil_w; p.pad_h = c.pad_h; p.pad_w = c.pad_w;
  p.x_vec = vec_of(c.x, c.x_outer, c.x_inner, c.x_pix_stride, c.Cin);
  p.dy = c.dy; p.dy_outer = c.dy_outer; p.dy_inner = c.dy_inner; p.dy_pix_stride = c.dy_pix_stride; p.Cout = c.Cout;
  p.dy_vec = vec_of(c.dy, c.dy_outer, c.dy_inner, c.dy_pix_stride, c.Cout);
  p.H = c.H; p.W = c.W; p.HW = c.H * c.W; p.T_inner = c.T_inner; p.M = (int)M;
  p.gw = c.gw; p.gbias = c.gbias; p.dbg = g_wg_debug;
  // n tile: 16/32/64/128 wide with 64-pixel stages, or 256 wide with 32-pixel stages when Cout > 128 (x is then read
  // by half as many n tiles; the dY row of a thread must keep a fixed float4 column)
  int bn = 16;
  while (bn < c.Cout && bn < 128) bn <<= 1;
  const bool wide = c.Cout > 128 && !g_wg_narrow;
  if (wide) bn = 256;
  const int WG_PIX = wide ? 32 : 64, WG_GRP = WG_PIX * 128;
  p.BLOCK_N = bn;
  const int n_tiles = (c.Cout + bn - 1) / bn;
  // two M tiles per CTA on the wide path when there are at least two: dY is then re-read by half as many CTAs
  const int MT = (wide && p.taps * p.Cin_p > WG_M && !g_wg_single_m) ? 2 : 1;
  const int k_tiles = c.gw ? (p.taps * p.Cin_p + WG_M * MT - 1) / (WG_M * MT) : 1;
  p.tmem_cols = (int)tmem_cols_for(bn * MT);
  p.b_term_bytes = (bn + 63) / 64 * WG_GRP;
  p.stage_bytes = c.math * (2 * MT * WG_GRP + p.b_term_bytes);
  int stg = (227 * 1024 - (int)sizeof(Book) - 2048) / p.stage_bytes;
  if (stg > kStagesMax) stg = kStagesMax;
  FOV_CHECK_ARG(stg >= 2, "tile does not fit shared memory");
  p.stages = stg;
  p.data_bytes = stg * p.stage_bytes;
  if (p.data_bytes < 4 * 32 * RS * 4) p.data_bytes = 4 * 32 * RS * 4;
  p.data_bytes = (p.data_bytes + 1023) / 1024 * 1024;
  const size_t smem = (size_t)p.data_bytes + sizeof(Book) + 1024;
  p.nblocks = (int)((M + WG_PIX - 1) / WG_PIX);
  // split the pixel reduction so that ~2 waves of CTAs cover the SMs, at least 8 blocks per CTA
  long long want = (2LL * fov_num_sms() + (long long)k_tiles * n_tiles - 1) / ((long long)k_tiles * n_tiles);
  long long max_splits = (p.nblocks + 7) / 8;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  p.blocks_per_split = (int)((p.nblocks + want - 1) / want);
  const int splits = (p.nblocks + p.blocks_per_split - 1) / p.blocks_per_split;
  dim3 grid((unsigned)k_tiles, (unsigned)n_tiles, (unsigned)splits);
  const bool fast = p.x_vec == 4 && p.dy_vec == 4 && p.Cout % 4 == 0;
  if (wide && MT == 2) {
    if (fast) {
      if (c.math == 1) return launch<1, true, 32, 2>(p, grid, smem, st);
      if (c.math == 2) return launch<2, true, 32, 2>(p, grid, smem, st);
      return launch<3, true, 32, 2>(p, grid, smem, st);
    }
    if (c.math == 1) return launch<1, false, 32, 2>(p, grid, smem, st);
    if (c.math == 2) return launch<2, false, 32, 2>(p, grid, smem, st);
    return launch<3, false, 32, 2>(p, grid, smem, st);
  }
  if (wide) {
    if (fast) {
      if (c.math == 1) return launch<1, true, 32, 1>(p, grid, smem, st);
      if (c.math == 2) return launch<2, true, 32, 1>(p, grid, smem, st);
      return launch<3, true, 32, 1>(p, grid, smem, st);
    }
    if (c.math == 1) return launch<1, false, 32, 1>(p, grid, smem, st);
    if (c.math == 2) return launch<2, false, 32, 1>(p, grid, smem, st);
    return launch<3, false, 32, 1>(p, grid, smem, st);
  }
  if (fast) {
    if (c.math == 1) return launch<1, true, 64, 1>(p, grid, smem, st);
    if (c.math == 2) return launch<2, true, 64, 1>(p, grid, smem, st);
    return launch<3, true, 64, 1>(p, grid, smem, st);
  }
  if (c.math == 1) return launch<1, false, 64, 1>(p, grid, smem, st);
  if (c.math == 2) return launch<2, false, 64, 1>(p, grid, smem, st);
  return launch<3, false, 64, 1>(p, grid, smem, st);
}

namespace {
TcWgrad wgrad_from_cfg(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw, float* gbias, int math) {
  TcWgrad c{};
  c.x = x; c.x_outer = cfg->x_img_stride; c.x_inner = 0; c.x_pix_stride = cfg->x_pix_stride; c.Cin = cfg->Cin;
  c.kh = cfg->kh; c.kw = cfg->kw; c.dil_h = cfg->dil_h; c.dil_w = cfg->dil_w; c.pad_h = cfg->pad_h; c.pad_w = cfg->pad_w;
  c.dy = dy; c.dy_outer = cfg->y_img_stride; c.dy_inner = 0; c.dy_pix_stride = cfg->y_pix_stride; c.Cout = cfg->Cout;
  c.N_img = cfg->N; c.T_inner = 1; c.H = cfg->H; c.W = cfg->W;
  c.math = math; c.gw = gw; c.gbias = gbias;
  return c;
}
}  // namespace

extern "C" size_t fov_conv_wgrad_ws_bytes(const fov_conv_cfg* cfg, int math) {
  if (!cfg || math < 1 || math > 3 || cfg->N <= 0 || cfg->H <= 0 || cfg->W <= 0 || cfg->Cin <= 0 || cfg->Cout <= 0) return 0;
  float dummy = 0.0f;
  return tc_wgrad_planes_ws_bytes(wgrad_from_cfg(cfg, &dummy, &dummy, &dummy, nullptr, math));
}

// with a workspace of fov_conv_wgrad_ws_bytes() bytes: wide k x k convolutions take the TMA-fed kernel of
// wgrad_planes_tc.cu (operands converted once into bf16 planes); everything else is fov_conv2d_bwd_weight_tc
extern "C" int fov_conv2d_bwd_weight_tc_ws(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw,
                                           float* gbias, void* ws, int math, void* stream) {
  FOV_CHECK_ARG(cfg != nullptr, "cfg is NULL");
  if (ws && gw && x && dy && fov_conv_wgrad_ws_bytes(cfg, math) > 0) {
    int rc = tc_wgrad_planes_run(wgrad_from_cfg(cfg, x, dy, gw, gbias, math), ws, (cudaStream_t)stream);
    if (rc || !gbias || (cfg->Cout + 63) / 64 * 64 <= 2048) return rc;
    return fov_conv2d_bwd_weight(cfg, x, dy, nullptr, gbias, stream);      // very wide dY: separate column-sum kernel
  }
  return fov_conv2d_bwd_weight_tc(cfg, x, dy, gw, gbias, math, stream);
}

extern "C" int fov_conv2d_bwd_weight_tc(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw,
                                        float* gbias, int math, void* stream) {
  FOV_CHECK_ARG(cfg != nullptr, "cfg is NULL");
  FOV_CHECK_ARG(cfg->N > 0 && cfg->H > 0 && cfg->W > 0 && cfg->Cin > 0 && cfg->Cout > 0, "bad shape");
  FOV_CHECK_ARG(cfg->kh > 0 && cfg->kw > 0 && cfg->dil_h > 0 && cfg->dil_w > 0, "bad kernel/dilation");
  FOV_CHECK_ARG(dy != nullptr && (gw == nullptr || x != nullptr), "NULL pointer");
  if (!gw) return fov_conv2d_bwd_weight(cfg, x, dy, nullptr, gbias, stream);   // bias only: column-sum kernel
  TcWgrad c{};
  c.x = x; c.x_outer = cfg->x_img_stride; c.x_inner = 0; c.x_pix_stride = cfg->x_pix_stride; c.Cin = cfg->Cin;
  c.kh = cfg->kh; c.kw = cfg->kw; c.dil_h = cfg->dil_h; c.dil_w = cfg->dil_w; c.pad_h = cfg->pad_h; c.pad_w = cfg->pad_w;
  c.dy = dy; c.dy_outer = cfg->y_img_stride; c.dy_inner = 0; c.dy_pix_stride = cfg->y_pix_stride; c.Cout = cfg->Cout;
  c.N_img = cfg->N; c.T_inner = 1; c.H = cfg->H; c.W = cfg->W;
  c.math = math; c.gw = gw; c.gbias = gbias;
  return tc_wgrad_run(c, (cudaStream_t)stream);
}
