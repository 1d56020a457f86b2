// Tensor-core (tcgen05 + TMEM) implicit-GEMM convolution family for sm_100a.
//
//   D[m][n] = sum_k A(m,k) * Wp[k][n]      m = pixel, n = output channel,
//                                          k = (segment, tap, channel) of the im2col row
//
// Shifted-tap design: the im2col matrix is never built, not even in shared memory.
//   * Pixels are numbered in a zero-PADDED linear frame: image n, row y, column x sits at
//     L = n*Hp*Wp + (y+PLh)*Wp + (x+PLw), where the frame is padded by the 'same' padding of the
//     convolution.  In this frame a kernel tap (ty,tx) is ONE constant row offset
//     (ty*dil_h - pad_h)*Wp + (tx*dil_w - pad_w) for every pixel, and TF 'same' zero padding
//     is just the zero rows of the frame.
//   * A CTA owns 128 consecutive frame positions (the D rows; pad positions are computed and
//     dropped).  Its producers stage the activations of those positions plus the halo ONCE:
//     one 128-byte, 128B-swizzled shared-memory row per position holding up to 64 channels
//     (fp32 -> 1..3 bf16 terms, split on the fly).
//   * tcgen05.mma operand descriptors address shared memory by absolute address (the 128B swizzle
//     is a function of the address bits: tests/cuda/tc_shift_probe.cu), so the A operand of tap t
//     is the SAME staged tile with its start address moved by shift(t) rows: kh*kw MMAs per
//     16-channel slice, no replicated data, each activation element loaded and converted once.
//   * Weights are repacked once per call into the swizzled image of every (n-tile, k-block) B tile
//     and streamed through an mbarrier ring by TMA-engine bulk copies (UBLKCP).
//   * Accumulation in TMEM (fp32); with NS bf16 terms per operand the NS(NS+1)/2 significant cross
//     products are issued (bf16x2: 3 MMAs, ~1e-5 max-abs error; bf16x3: 6 MMAs).
//   * Epilogues read the accumulator with tcgen05.ld and go through a per-warp shared-memory
//     transpose so that global accesses are coalesced float4 row segments:
//       TC_EPI_CONV: y = act(acc + bias + beta*y)
//       TC_EPI_LSTM: ConvLSTM2D gate algebra + cell update of one timestep, A = [x_t | h_{t-1}]
//                    (keras ConvLSTM2D, mycode/others_LSTM_span_whole.py:88-100,
//                     mycode/convlstm_seq2seq.py:100-126,146-165)
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                    // bf16 per 128-byte swizzle row / weight k-block
constexpr int kProducers = 256;
constexpr int kEpiWarps = 8;                   // warps 0-7: producers, then epilogue (warp w: TMEM lanes 32*(w&3)..)
constexpr int kThreads = 320;                  // + weight-copy warp (8) + MMA warp (9)
constexpr int kMaxBStages = 4;
constexpr int kMaxRegionRows = 384;            // 128 (or 2 x 128) D rows + halo
constexpr int kMaxMT = 2;                      // M tiles (128 frame positions each) per CTA
constexpr int kMaxTaps = 64;
constexpr int CONV_RS = 36;                    // conv epilogue staging row stride (floats)
enum { O_OUT = 0, O_H = 1, O_GATES = 2, O_CPREV = 3, O_DENSE = 4 };

struct DevSeg {
  const float* x;
  long long outer, inner;
  int pix_stride, Cin, Cin_p, cp_log2, cw, nch, lpr_log2, kw, taps, dil_h, dil_w, pad_h, pad_w, k_begin, vec;
  int minshift, R, region_off, nbuf;
  int row_bytes, swz_mask, term_bytes;   // staged row = cw bf16 (32/64/128 B) in the matching swizzle mode
  uint32_t desc_hi;                      // high word of the A descriptors of this segment
};

struct DevParams {
  DevSeg seg[2];
  int nseg, mode_b, dbg, n_tiles, pipe, MT;
  int H, W, Hp, Wp, PLh, PLw, HpWp, T_inner, N_img, total_pos;
  int K_total, KB, Cout, BLOCK_N, b_stages, tmem_cols, b_off, b_stage_bytes, data_bytes;
  const uint8_t* wpk;
  // conv epilogue
  const float* bias;
  float* y;
  long long y_outer, y_inner;
  int y_pix_stride, act, vec_out;
  float beta;
  // lstm epilogue
  int F, rec_act;
  const float* c_prev; long long cp_outer, cp_inner;
  float* c_out;        long long c_outer, c_inner;
  float* h_out;        long long h_outer, h_inner; int h_pix_stride;
  float* gates_out;    long long g_outer, g_inner;
  float *hT, *cT;
};

// shared-memory bookkeeping placed after the data area (activation regions, weight ring / staging)
struct Book {
  int rp[2][kMaxRegionRows];            // per region row: element offset of that position's pixel, -1 = zero row
  int off_o[5][kMaxMT * BLOCK_M];       // epilogue element offsets of the D rows (image + pixel), see O_*
  int valid[kMaxMT * BLOCK_M];          // D row is a real pixel
  float bias_s[256];                    // LSTM epilogue: the 4F gate biases
  int tapshift[2][kMaxTaps];            // row offset of every tap relative to the region start
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages], a_full[2], a_empty[2], tmem_full;
  uint32_t tmem_ptr;
};

// optional per-CTA phase timestamps (fov_debug_timeline): [cta][0..5] = start, setup done, activations
// staged, accumulator complete, epilogue done, smid
__device__ unsigned long long g_tc_timeline[256 * 8];

__device__ __forceinline__ long long img_off(int n, int T_inner, long long outer, long long inner) {
  const int no = n / T_inner;
  return (long long)no * outer + (long long)(n - no * T_inner) * inner;
}

template <int NS, int EPI>
__global__ void __launch_bounds__(kThreads, 2) tc_conv_kernel(const DevParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  Book* bk = reinterpret_cast<Book*>(smem + p.data_bytes);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // 1-D grid, n tile fastest: the CTAs that share the activations of an M tile run together, so the rows they
  // stage come from L2 after the first touch (backward-data of the wide dense layers has 15 n tiles per M tile)
  const int m_tile = (int)(blockIdx.x / (unsigned)p.n_tiles);
  const int n_tile = (int)(blockIdx.x - (unsigned)m_tile * (unsigned)p.n_tiles);
  const int MT = p.MT;                          // 128-row accumulator tiles of this CTA (2: wide layers, weights read once per 256 rows)
  const int L0 = m_tile * BLOCK_M * MT;         // first frame position of this tile
  const int n0 = n_tile * p.BLOCK_N;
  const int S = p.b_stages;
  const uint32_t b_tile_bytes = (uint32_t)p.BLOCK_N * 128u;
  const bool dbg = p.dbg && m_tile < 256 && n_tile == 0 && tid == 0;
  if (dbg) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tc_timeline[m_tile * 8 + 0] = clock64();
    g_tc_timeline[m_tile * 8 + 5] = smid;
  }

  // ---------------- setup: frame position -> pixel tables ----------------
  for (int dr = tid; dr < BLOCK_M * MT; dr += kThreads) {
    const int L = L0 + dr;
    int ok = 0;
    if (L < p.total_pos) {
      const int n = L / p.HpWp, rem = L - n * p.HpWp;
      const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
      const int y = yp - p.PLh, x = xp - p.PLw;
      if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W) {
        ok = 1;
        const int pix = y * p.W + x;
        if (EPI == TC_EPI_CONV) {
          bk->off_o[O_OUT][dr] = (int)(img_off(n, p.T_inner, p.y_outer, p.y_inner) + (long long)pix * p.y_pix_stride);
        } else {
          bk->off_o[O_OUT][dr] = (int)(img_off(n, p.T_inner, p.c_outer, p.c_inner) + (long long)pix * p.F);
          bk->off_o[O_H][dr] = (int)(img_off(n, p.T_inner, p.h_outer, p.h_inner) + (long long)pix * p.h_pix_stride);
          bk->off_o[O_GATES][dr] = (int)(img_off(n, p.T_inner, p.g_outer, p.g_inner) + (long long)pix * 4 * p.F);
          bk->off_o[O_CPREV][dr] = (int)(img_off(n, p.T_inner, p.cp_outer, p.cp_inner) + (long long)pix * p.F);
          bk->off_o[O_DENSE][dr] = (n * p.H * p.W + pix) * p.F;
        }
      }
    }
    bk->valid[dr] = ok;
  }
  for (int s = 0; s < p.nseg; ++s) {
    const DevSeg& sg = p.seg[s];
    for (int i = tid; i < sg.R; i += kThreads) {
      const int L = L0 + sg.minshift + i;
      int ptr = -1;
      if (L >= 0 && L < p.total_pos && sg.x != nullptr) {
        const int n = L / p.HpWp, rem = L - n * p.HpWp;
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const int y = yp - p.PLh, x = xp - p.PLw;
        if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W)
          ptr = (int)(img_off(n, p.T_inner, sg.outer, sg.inner) + (long long)(y * p.W + x) * sg.pix_stride);
      }
      bk->rp[s][i] = ptr;
    }
    for (int t = tid; t < sg.taps; t += kThreads) {
      const int ty = t / sg.kw, tx = t - ty * sg.kw;
      bk->tapshift[s][t] = (ty * sg.dil_h - sg.pad_h) * p.Wp + (tx * sg.dil_w - sg.pad_w) - sg.minshift;
    }
  }
  if (EPI == TC_EPI_LSTM && tid < 4 * p.F) bk->bias_s[tid] = __ldg(&p.bias[tid]);
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bk->b_full[s]), 1);
      mbar_init(smem_u32(&bk->b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bk->a_full[s]), kProducers);
      mbar_init(smem_u32(&bk->a_empty[s]), 1);
    }
    mbar_init(smem_u32(&bk->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;
  if (dbg) g_tc_timeline[m_tile * 8 + 1] = clock64();

  if (warp < kEpiWarps) {
    // ---------------- activation producers ----------------
    // Stage one 64-channel slice of a segment: region row i <- frame position L0 + minshift + i.
    auto load_region = [&](int s, int cc, int buf) {
      const DevSeg& sg = p.seg[s];
      uint8_t* reg = smem + sg.region_off + (size_t)buf * NS * sg.term_bytes;
      const int lg = sg.lpr_log2, lpr = 1 << lg;
      const int c4 = tid & (lpr - 1);
      const int row0 = tid >> lg, rstep = kProducers >> lg;        // rstep is a multiple of 8 rows
      const int ch = cc * BLOCK_K + c4 * 4;
      int nvalid = sg.Cin - ch;
      nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
      // absolute-address swizzle of this thread's slot in row0; rows row0 + 8k keep the same phase
      const uint32_t a0 = (uint32_t)row0 * sg.row_bytes + (uint32_t)c4 * 8u;
      const uint32_t st_off = a0 ^ (((a0 >> 7) & (uint32_t)sg.swz_mask) << 4);
      const int* rp = bk->rp[s];
      const float* xb = sg.x + ch;
      for (int r = row0; r < sg.R; r += 8 * rstep) {
        float4 v[8];
        if (sg.vec == 4) {
          // aligned fast path, free of value merges: all 8 loads stay in flight
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = r + j * rstep;
            const int off = (row < sg.R && nvalid > 0) ? rp[row] : -1;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(xb + off));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = r + j * rstep;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < sg.R && nvalid > 0) {
              const int off = rp[row];
              if (off >= 0) v[j] = ldg_vec4(xb + off, nvalid, sg.vec);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = r + j * rstep;
          if (row < sg.R) {
            uint2 pk[NS];
            split4<NS>(v[j], pk);
#pragma unroll
            for (int t = 0; t < NS; ++t)
              *reinterpret_cast<uint2*>(reg + t * sg.term_bytes + (row - row0) * sg.row_bytes + st_off) = pk[t];
          }
        }
      }
    };
    if (!p.mode_b) {
      for (int s = 0; s < p.nseg; ++s) load_region(s, 0, 0);
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bk->a_full[0]));
      if (dbg) g_tc_timeline[m_tile * 8 + 2] = clock64();
    } else if (p.pipe) {
      // wide dense / 1x1 layers (one 8-row batch per thread and chunk, aligned rows): the loads of chunk cc+1 are
      // issued before chunk cc is converted, so the HBM latency hides behind the conversion and the ring wait
      const DevSeg& sg = p.seg[0];
      const int nch = sg.nch;
      const int lg = sg.lpr_log2, lpr = 1 << lg;
      const int c4 = tid & (lpr - 1);
      const int row0 = tid >> lg, rstep = kProducers >> lg;
      const uint32_t a0 = (uint32_t)row0 * sg.row_bytes + (uint32_t)c4 * 8u;
      const uint32_t st_off = a0 ^ (((a0 >> 7) & (uint32_t)sg.swz_mask) << 4);
      int offs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = row0 + j * rstep;
        offs[j] = row < sg.R ? bk->rp[0][row] : -1;
      }
      auto issue = [&](int cc, float4 (&v)[8]) {
        const int ch = cc * BLOCK_K + c4 * 4;
        const bool ok = ch < sg.Cin;                       // Cin % 4 == 0 on this path
        const float* xb = sg.x + ch;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && offs[j] >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(xb + offs[j]));
        }
      };
      auto commit = [&](int buf, const float4 (&v)[8]) {
        uint8_t* reg = smem + sg.region_off + (size_t)buf * NS * sg.term_bytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (row0 + j * rstep < sg.R) {
            uint2 pk[NS];
            split4<NS>(v[j], pk);
#pragma unroll
            for (int t = 0; t < NS; ++t)
              *reinterpret_cast<uint2*>(reg + t * sg.term_bytes + (j * rstep) * sg.row_bytes + st_off) = pk[t];
          }
        }
      };
      float4 va[8], vb[8];
      issue(0, va);
      for (int cc = 0; cc < nch; cc += 2) {
        if (cc + 1 < nch) issue(cc + 1, vb);
        mbar_wait(smem_u32(&bk->a_empty[0]), ((uint32_t)(cc >> 1) & 1u) ^ 1u);
        commit(0, va);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bk->a_full[0]));
        if (cc + 1 < nch) {
          if (cc + 2 < nch) issue(cc + 2, va);
          mbar_wait(smem_u32(&bk->a_empty[1]), ((uint32_t)(cc >> 1) & 1u) ^ 1u);
          commit(1, vb);
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&bk->a_full[1]));
        }
      }
    } else {
      const int nch = p.seg[0].nch;
      for (int cc = 0; cc < nch; ++cc) {
        const int buf = cc & 1;
        mbar_wait(smem_u32(&bk->a_empty[buf]), ((uint32_t)(cc >> 1) & 1u) ^ 1u);
        load_region(0, cc, buf);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bk->a_full[buf]));
      }
    }
  } else if (warp == 8) {
    // ---------------- weight producer: one bulk copy per k-block, in MMA consumption order ----------------
    if (lane == 0) {
      const uint8_t* src = p.wpk + (size_t)n_tile * p.KB * NS * b_tile_bytes;
      const int taps = p.seg[0].taps, nch = p.seg[0].nch;
      int cc = 0, tap = 0;
      for (int it = 0; it < p.KB; ++it) {
        const int kb = p.mode_b ? tap * nch + cc : it;
        if (p.mode_b && ++tap == taps) { tap = 0; ++cc; }
        const int stage = it % S;
        mbar_wait(smem_u32(&bk->b_empty[stage]), ((uint32_t)(it / S) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&bk->b_full[stage]);
        mbar_arrive_expect_tx(bar, NS * b_tile_bytes);
        bulk_g2s(base + p.b_off + (uint32_t)stage * p.b_stage_bytes, src + (size_t)kb * NS * b_tile_bytes,
                 NS * b_tile_bytes, bar);
      }
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(BLOCK_M, p.BLOCK_N, 0, 0);
      uint32_t first = 1;
      // one 16-wide K step: A = staged region of (seg, buffer) shifted by the tap, B = weight ring slot
      const uint32_t mt_rows = (uint32_t)BLOCK_M * (uint32_t)p.seg[0].row_bytes;     // second accumulator tile: 128 rows further
      auto kstep = [&](uint32_t a_hi, uint32_t a_addr, uint32_t a_term, uint32_t b_addr) {
#pragma unroll
        for (int sum = NS - 1; sum >= 0; --sum) {     // smallest cross terms first, hi*hi last
#pragma unroll
          for (int sa = 0; sa <= sum; ++sa) {
            const int sb = sum - sa;
            const uint64_t bd = desc_at(kDescHi128, b_addr + sb * b_tile_bytes);
            umma_bf16(tmem_d, desc_at(a_hi, a_addr + sa * a_term), bd, idesc, first ^ 1u);
            if (MT == 2)
              umma_bf16(tmem_d + (uint32_t)p.BLOCK_N, desc_at(a_hi, a_addr + sa * a_term + mt_rows), bd, idesc, first ^ 1u);
            first = 0;
          }
        }
      };
      if (!p.mode_b) {
        mbar_wait(smem_u32(&bk->a_full[0]), 0);
        tc_fence_after();
        for (int kb = 0; kb < p.KB; ++kb) {
          const int stage = kb % S;
          mbar_wait(smem_u32(&bk->b_full[stage]), (uint32_t)(kb / S) & 1u);
          tc_fence_after();
          const uint32_t b0 = base + p.b_off + (uint32_t)stage * p.b_stage_bytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int k = kb * BLOCK_K + k4 * 16;
            if (k < p.K_total) {
              const int s = (p.nseg > 1 && k >= p.seg[1].k_begin) ? 1 : 0;
              const DevSeg& sg = p.seg[s];
              const int kk = k - sg.k_begin;
              const int tap = kk >> sg.cp_log2, c0 = kk & (sg.Cin_p - 1);
              const uint32_t a_addr = base + sg.region_off + (uint32_t)bk->tapshift[s][tap] * sg.row_bytes + (uint32_t)c0 * 2u;
              kstep(sg.desc_hi, a_addr, (uint32_t)sg.term_bytes, b0 + k4 * 32);
            }
          }
          umma_commit(smem_u32(&bk->b_empty[stage]));
        }
      } else {
        const DevSeg& sg = p.seg[0];
        const uint32_t a_term = (uint32_t)sg.term_bytes;
        int it = 0;
        for (int cc = 0; cc < sg.nch; ++cc) {
          const int buf = cc & 1;
          mbar_wait(smem_u32(&bk->a_full[buf]), (uint32_t)(cc >> 1) & 1u);
          tc_fence_after();
          const uint32_t areg = base + sg.region_off + (uint32_t)buf * NS * a_term;
          for (int tap = 0; tap < sg.taps; ++tap, ++it) {
            const int stage = it % S;
            mbar_wait(smem_u32(&bk->b_full[stage]), (uint32_t)(it / S) & 1u);
            tc_fence_after();
            const uint32_t b0 = base + p.b_off + (uint32_t)stage * p.b_stage_bytes;
            const uint32_t a0 = areg + (uint32_t)bk->tapshift[0][tap] * 128u;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) kstep(sg.desc_hi, a0 + k4 * 32, a_term, b0 + k4 * 32);
            umma_commit(smem_u32(&bk->b_empty[stage]));
          }
          umma_commit(smem_u32(&bk->a_empty[buf]));
        }
      }
      umma_commit(smem_u32(&bk->tmem_full));
    }
  }

  // ---------------- epilogue (warps 0-7; warp w owns TMEM lanes / D rows 32*(w&3).., every other column pass) -------
  // all MMAs have completed when tmem_full fires, so the data area is free: the staging rows alias it.
  if (warp < kEpiWarps) {
    const int q = warp & 3, half = warp >> 2;
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * CONV_RS);     // 32 rows x 32 floats (+4 pad)

    if (EPI == TC_EPI_CONV) {
      mbar_wait(smem_u32(&bk->tmem_full), 0);
      tc_fence_after();
      if (dbg) g_tc_timeline[m_tile * 8 + 3] = clock64();
      for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * BLOCK_M + q * 32;
      const uint32_t t_row = tmem_d + (uint32_t)(mt * p.BLOCK_N) + ((uint32_t)(q * 32) << 16);
      for (int c0 = half * 32; c0 < p.BLOCK_N; c0 += 64) {
        float v[32];
        tmem_ld16(t_row + c0, v);
        tmem_ld16(t_row + c0 + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(&stg[lane * CONV_RS + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        if (p.vec_out) {
          const int c4 = (lane & 7) * 4;
          const int col = n0 + c0 + c4;
          const bool col_ok = (c0 + c4 < p.BLOCK_N) && (col < p.Cout);
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll 2
          for (int it = 0; it < 32; it += 4) {
            const int rr = it + (lane >> 3), row = r0 + rr;
            if (col_ok && bk->valid[row]) {
              float* dst = p.y + bk->off_o[O_OUT][row] + col;
              float4 a = *reinterpret_cast<const float4*>(&stg[rr * CONV_RS + c4]);
              a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
              if (p.beta != 0.0f) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                a.x += p.beta * o.x; a.y += p.beta * o.y; a.z += p.beta * o.z; a.w += p.beta * o.w;
              }
              if (p.act != FOV_ACT_LINEAR) {
                a.x = fov_act(p.act, a.x); a.y = fov_act(p.act, a.y); a.z = fov_act(p.act, a.z); a.w = fov_act(p.act, a.w);
              }
              *reinterpret_cast<float4*>(dst) = a;
            }
          }
        } else {
          // Cout not a multiple of 4 (e.g. the 198-wide reconstruction head): one column per lane, 8 rows per batch
          // with the loads / stores grouped so that the batch has no per-row control flow
          const int col = n0 + c0 + lane;
          const bool col_ok = (c0 + lane < p.BLOCK_N) && (col < p.Cout);
          const float bv = (col_ok && p.bias) ? __ldg(&p.bias[col]) : 0.0f;
          for (int rr0 = 0; rr0 < 32; rr0 += 8) {
            float val[8];
            float* dst[8];
            bool ok[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int row = r0 + rr0 + j;
              ok[j] = col_ok && bk->valid[row] != 0;
              dst[j] = p.y + (ok[j] ? bk->off_o[O_OUT][row] + col : 0);
              val[j] = stg[(rr0 + j) * CONV_RS + lane] + bv;
            }
            if (p.beta != 0.0f) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (ok[j]) val[j] += p.beta * *dst[j];
            }
            if (p.act == FOV_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) val[j] = fmaxf(val[j], 0.0f);
            } else if (p.act == FOV_ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 8; ++j) val[j] = tanhf(val[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (ok[j]) *dst[j] = val[j];
          }
        }
        __syncwarp();
      }
      }
    } else {
      const int r0 = q * 32;
      const uint32_t t_row = tmem_d + ((uint32_t)r0 << 16);
      // Fused ConvLSTM step.  Accumulator columns: [i | f | c~ | o], each F wide, processed 8 channels per
      // pass; the two warps of a lane quarter take alternate passes.  Row-per-thread values go through the
      // staging rows two 8-wide segments at a time and leave as coalesced float4 row segments.
      const int F = p.F;
      constexpr int CW = 8;
      const int npass = F / CW;
      const int srow = lane >> 1, sc4 = (lane & 1) * 4;        // 2 float4 lanes per 8-wide row segment
      // prefetch c_{t-1} of this warp's first two passes while the MMAs are still running
      float4 cpre[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          cpre[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int pass = half + 2 * u;
          const int row = r0 + i * 16 + srow;
          if (p.c_prev && pass < npass && bk->valid[row])
            cpre[u][i] = __ldg(reinterpret_cast<const float4*>(p.c_prev + bk->off_o[O_CPREV][row] + pass * CW + sc4));
        }
      mbar_wait(smem_u32(&bk->tmem_full), 0);
      tc_fence_after();
      if (dbg) g_tc_timeline[m_tile * 8 + 3] = clock64();
      const int row_a = r0 + srow, row_b = r0 + 16 + srow;
      const bool ok_a = bk->valid[row_a] != 0, ok_b = bk->valid[row_b] != 0;
      // coalesced store of one staged 8-wide segment
      auto seg_store = [&](int stage_col, float* dst, int oidx, int dst_col, float* dense) {
        if (!dst) return;
        if (ok_a) {
          const float4 v = *reinterpret_cast<const float4*>(&stg[srow * CONV_RS + stage_col + sc4]);
          *reinterpret_cast<float4*>(dst + bk->off_o[oidx][row_a] + dst_col + sc4) = v;
          if (dense) *reinterpret_cast<float4*>(dense + bk->off_o[O_DENSE][row_a] + dst_col + sc4) = v;
        }
        if (ok_b) {
          const float4 v = *reinterpret_cast<const float4*>(&stg[(16 + srow) * CONV_RS + stage_col + sc4]);
          *reinterpret_cast<float4*>(dst + bk->off_o[oidx][row_b] + dst_col + sc4) = v;
          if (dense) *reinterpret_cast<float4*>(dense + bk->off_o[O_DENSE][row_b] + dst_col + sc4) = v;
        }
      };
      float* myrow = stg + lane * CONV_RS;
      for (int pass = half, u = 0; pass < npass; pass += 2, ++u) {
        const int cpass = pass * CW;
        // ---- c_{t-1}: coalesced block -> staging -> one row per thread ----
        float cpv[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) cpv[j] = 0.0f;
        if (p.c_prev) {
          float4 va, vb;
          if (u < 2) {
            va = u == 0 ? cpre[0][0] : cpre[1][0];
            vb = u == 0 ? cpre[0][1] : cpre[1][1];
          } else {
            va = vb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok_a) va = __ldg(reinterpret_cast<const float4*>(p.c_prev + bk->off_o[O_CPREV][row_a] + cpass + sc4));
            if (ok_b) vb = __ldg(reinterpret_cast<const float4*>(p.c_prev + bk->off_o[O_CPREV][row_b] + cpass + sc4));
          }
          *reinterpret_cast<float4*>(&stg[srow * CONV_RS + sc4]) = va;
          *reinterpret_cast<float4*>(&stg[(16 + srow) * CONV_RS + sc4]) = vb;
          __syncwarp();
          const float4 t0 = *reinterpret_cast<const float4*>(&myrow[0]);
          const float4 t1 = *reinterpret_cast<const float4*>(&myrow[4]);
          cpv[0] = t0.x; cpv[1] = t0.y; cpv[2] = t0.z; cpv[3] = t0.w;
          cpv[4] = t1.x; cpv[5] = t1.y; cpv[6] = t1.z; cpv[7] = t1.w;
          __syncwarp();
        }
        // ---- gates from TMEM ----
        float g[4][CW];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) tmem_ld8(t_row + gi * F + cpass, g[gi]);
        tmem_ld_wait();
        float cn[CW], hn[CW];
        // recurrent-activation choice hoisted out of the channel loop (a per-call branch serialises the channel chains)
        if (p.rec_act == 0) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int ch = cpass + j;
            const float ai = fminf(fmaxf(0.2f * (g[0][j] + bk->bias_s[ch]) + 0.5f, 0.0f), 1.0f);
            const float af = fminf(fmaxf(0.2f * (g[1][j] + bk->bias_s[F + ch]) + 0.5f, 0.0f), 1.0f);
            const float ag = fast_tanh(g[2][j] + bk->bias_s[2 * F + ch]);
            const float ao = fminf(fmaxf(0.2f * (g[3][j] + bk->bias_s[3 * F + ch]) + 0.5f, 0.0f), 1.0f);
            cn[j] = af * cpv[j] + ai * ag;
            hn[j] = ao * fast_tanh(cn[j]);
            g[0][j] = ai; g[1][j] = af; g[2][j] = ag; g[3][j] = ao;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int ch = cpass + j;
            const float ai = __fdividef(1.0f, 1.0f + __expf(-(g[0][j] + bk->bias_s[ch])));
            const float af = __fdividef(1.0f, 1.0f + __expf(-(g[1][j] + bk->bias_s[F + ch])));
            const float ag = fast_tanh(g[2][j] + bk->bias_s[2 * F + ch]);
            const float ao = __fdividef(1.0f, 1.0f + __expf(-(g[3][j] + bk->bias_s[3 * F + ch])));
            cn[j] = af * cpv[j] + ai * ag;
            hn[j] = ao * fast_tanh(cn[j]);
            g[0][j] = ai; g[1][j] = af; g[2][j] = ag; g[3][j] = ao;
          }
        }
        // ---- rounds of staged segments: (c, h), then the four activated gates ----
        auto put = [&](int col, const float (&a)[CW]) {
          *reinterpret_cast<float4*>(&myrow[col]) = make_float4(a[0], a[1], a[2], a[3]);
          *reinterpret_cast<float4*>(&myrow[col + 4]) = make_float4(a[4], a[5], a[6], a[7]);
        };
        put(0, cn); put(8, hn);
        if (p.gates_out) { put(16, g[0]); put(24, g[1]); }
        __syncwarp();
        seg_store(0, p.c_out, O_OUT, cpass, p.cT);
        seg_store(8, p.h_out, O_H, cpass, p.hT);
        if (p.gates_out) {
          seg_store(16, p.gates_out, O_GATES, cpass, nullptr);
          seg_store(24, p.gates_out, O_GATES, F + cpass, nullptr);
          __syncwarp();
          put(0, g[2]); put(8, g[3]);
          __syncwarp();
          seg_store(0, p.gates_out, O_GATES, 2 * F + cpass, nullptr);
          seg_store(8, p.gates_out, O_GATES, 3 * F + cpass, nullptr);
        }
        __syncwarp();
      }
    }
  }

  if (dbg) g_tc_timeline[m_tile * 8 + 4] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// Weight packing: Keras-layout fp32 weights -> bf16 terms in the swizzled smem image of every
// (n-tile, k-block) B tile; k = k_begin(seg) + tap*Cin_p + channel.  One thread per 16-byte chunk.
// ---------------------------------------------------------------------------------------------
struct PackSeg {
  const float* w;
  int Cin, Cin_p, taps, k_begin, mode;
};
struct PackParams {
  PackSeg seg[2];
  int nseg, KB, Cout, BLOCK_N, n_tiles, NS;
  uint8_t* out;
};

__global__ void __launch_bounds__(256) tc_pack_kernel(const PackParams p) {
  const long long total = (long long)p.n_tiles * p.KB * p.NS * p.BLOCK_N * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cphys = (int)(idx & 7);
    long long r = idx >> 3;
    const int nl = (int)(r % p.BLOCK_N); r /= p.BLOCK_N;
    const int s = (int)(r % p.NS); r /= p.NS;
    const int kb = (int)(r % p.KB);
    const int nt = (int)(r / p.KB);
    const int c = cphys ^ (nl & 7);
    const int n = nt * p.BLOCK_N + nl;
    uint32_t outw[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float term[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = kb * BLOCK_K + c * 8 + e2 * 2 + h;
        const int si = (p.nseg > 1 && k >= p.seg[1].k_begin) ? 1 : 0;
        const PackSeg& sg = p.seg[si];
        const int kk = k - sg.k_begin;
        const int tap = kk / sg.Cin_p, ci = kk - tap * sg.Cin_p;
        float wv = 0.0f;
        if (tap < sg.taps && ci < sg.Cin && n < p.Cout) {
          if (sg.mode == 0) wv = __ldg(&sg.w[((long long)tap * sg.Cin + ci) * p.Cout + n]);
          else wv = __ldg(&sg.w[((long long)(sg.taps - 1 - tap) * p.Cout + n) * sg.Cin + ci]);
        }
        float rem = wv, t = 0.0f;
        for (int i = 0; i <= s; ++i) { t = bf16_round(rem); rem -= t; }
        term[h] = t;
      }
      outw[e2] = pack_bf16x2(term[0], term[1]);
    }
    *reinterpret_cast<uint4*>(p.out + idx * 16) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
  }
}

struct SegPlan {
  int Cin_p, cp_log2, cw, nch, lpr_log2, k_begin, taps, minshift, R, region_off, nbuf, row_bytes, term_bytes;
};
struct Plan {
  SegPlan sp[2];
  int PLh, PLw, Hp, Wp, mode_b, MT;
  int K_total, KB, BLOCK_N, n_tiles, NS, b_stages, b_stage_bytes, b_off, tmem_cols, data_bytes;
  long long total_pos;
  size_t smem_bytes, ws_bytes;
};

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

int make_plan_mt(const TcConv& c, Plan* pl, int MT);

// Wide single-term convolutions (the 512 / 1024-channel heads in bf16 mode) stream their weights from L2 at
// 64 B / cycle / SM at the tensor peak - above what L2 delivers to 148 SMs at once (~43 B / cycle / SM).  Two 128-row
// accumulator tiles per CTA (256 frame positions, 2 x BLOCK_N TMEM columns) read every weight tile once per 256 rows.
int make_plan(const TcConv& c, Plan* pl) {
  if (c.epi == TC_EPI_CONV && c.nseg == 1 && c.math == 1 && c.seg[0].Cin > 64 && c.Cout >= 128 && c.no_mt2 <= 0) {
    Plan p2;
    if (make_plan_mt(c, &p2, 2) == FOV_OK &&
        (c.no_mt2 < 0 || p2.total_pos >= 2LL * BLOCK_M * 2 * fov_num_sms() / p2.n_tiles)) {
      *pl = p2;
      return FOV_OK;
    }
    fov_set_error("");
  }
  return make_plan_mt(c, pl, 1);
}

int make_plan_mt(const TcConv& c, Plan* pl, int MT) {
  pl->MT = MT;
  FOV_CHECK_ARG(c.nseg == 1 || c.nseg == 2, "nseg must be 1 or 2");
  FOV_CHECK_ARG(c.math >= 1 && c.math <= 3, "math must be 1..3 bf16 terms");
  FOV_CHECK_ARG(c.N_img > 0 && c.H > 0 && c.W > 0 && c.Cout > 0 && c.T_inner > 0, "bad shape");
  pl->NS = c.math;
  // common zero-padded frame
  int PLh = 0, PHh = 0, PLw = 0, PHw = 0;
  for (int s = 0; s < c.nseg; ++s) {
    const TcSeg& g = c.seg[s];
    FOV_CHECK_ARG(g.Cin > 0 && g.kh > 0 && g.kw > 0 && g.dil_h > 0 && g.dil_w > 0 && g.pad_h >= 0 && g.pad_w >= 0,
                  "bad segment");
    FOV_CHECK_ARG(g.kh * g.kw <= kMaxTaps, "too many kernel taps");
    const int hh = (g.kh - 1) * g.dil_h - g.pad_h, hw = (g.kw - 1) * g.dil_w - g.pad_w;
    PLh = g.pad_h > PLh ? g.pad_h : PLh; PLw = g.pad_w > PLw ? g.pad_w : PLw;
    PHh = hh > PHh ? hh : PHh; PHw = hw > PHw ? hw : PHw;
  }
  pl->PLh = PLh; pl->PLw = PLw;
  pl->Hp = c.H + PLh + PHh; pl->Wp = c.W + PLw + PHw;
  pl->total_pos = (long long)c.N_img * pl->Hp * pl->Wp;
  FOV_CHECK_ARG(pl->total_pos < (1LL << 31) - 2 * kMaxRegionRows, "too many pixels for 32-bit indexing");
  int k = 0, any_multi = 0, region_bytes = 0;
  for (int s = 0; s < c.nseg; ++s) {
    const TcSeg& g = c.seg[s];
    SegPlan& sp = pl->sp[s];
    sp.Cin_p = g.Cin <= 16 ? 16 : (g.Cin <= 32 ? 32 : (g.Cin + 63) / 64 * 64);
    sp.cp_log2 = ilog2(sp.Cin_p);       // only used when Cin_p <= 64 (a power of two there)
    sp.cw = sp.Cin_p < 64 ? sp.Cin_p : 64;
    sp.nch = sp.Cin_p <= 64 ? 1 : sp.Cin_p / 64;
    sp.lpr_log2 = ilog2(sp.cw / 4);
    sp.taps = g.kh * g.kw;
    sp.k_begin = k;
    k += sp.taps * sp.Cin_p;
    sp.minshift = -(g.pad_h * pl->Wp + g.pad_w);
    const int maxshift = ((g.kh - 1) * g.dil_h - g.pad_h) * pl->Wp + ((g.kw - 1) * g.dil_w - g.pad_w);
    sp.R = (BLOCK_M * MT + maxshift - sp.minshift + 7) / 8 * 8;
    FOV_CHECK_ARG(sp.R <= kMaxRegionRows, "convolution halo too large for the shifted-tap kernel");
    sp.nbuf = sp.nch > 1 ? 2 : 1;
    any_multi |= sp.nch > 1;
    sp.row_bytes = sp.cw * 2;
    sp.term_bytes = (sp.R * sp.row_bytes + 1023) / 1024 * 1024;
    sp.region_off = region_bytes;
    region_bytes += sp.nbuf * pl->NS * sp.term_bytes;
  }
  FOV_CHECK_ARG(!(c.nseg == 2 && any_multi), "two-segment GEMM needs <= 64 channels per segment");
  pl->mode_b = any_multi;
  pl->K_total = k;
  pl->KB = (k + BLOCK_K - 1) / BLOCK_K;
  // epilogue staging (aliases the data area)
  int staging;
  if (c.epi == TC_EPI_LSTM) {
    FOV_CHECK_ARG(c.Cout % 4 == 0, "LSTM epilogue needs Cout = 4F");
    const int F = c.Cout / 4, CW = F < 16 ? F : 16;
    FOV_CHECK_ARG(F == 8 || F == 16 || F == 32 || F == 64, "LSTM epilogue supports F in {8,16,32,64}");
    (void)CW;
    staging = kEpiWarps * 32 * CONV_RS * 4;
  } else {
    staging = kEpiWarps * 32 * CONV_RS * 4;
  }
  // n tiling + weight ring depth
  const int kUsable = 227 * 1024, book = (int)sizeof(Book) + 1024 + 1024;
  const int n_pad = (c.Cout + 15) / 16 * 16;
  int max_n = c.epi == TC_EPI_LSTM ? c.Cout : 256;
  for (;;) {
    pl->n_tiles = (n_pad + max_n - 1) / max_n;
    pl->BLOCK_N = c.epi == TC_EPI_LSTM ? c.Cout : ((n_pad + pl->n_tiles - 1) / pl->n_tiles + 15) / 16 * 16;
    pl->b_stage_bytes = pl->NS * pl->BLOCK_N * 128;
    int st = (kUsable - book - region_bytes) / pl->b_stage_bytes;
    // dense layers (1x1 kernel, wide Cin) are bound by the activation staging, which is repeated per n tile: prefer
    // ONE n tile with a two-slot weight ring over two n tiles with three slots
    const bool dense_like = c.nseg == 1 && c.seg[0].kh * c.seg[0].kw == 1 && c.seg[0].Cin >= 128;
    const int want = pl->KB < 3 ? pl->KB : (dense_like ? 2 : 3);
    if (st >= want || max_n <= 32 || c.epi == TC_EPI_LSTM) {
      if (st > kMaxBStages) st = kMaxBStages;
      if (st > pl->KB) st = pl->KB;
      FOV_CHECK_ARG(st >= 1, "tile does not fit shared memory");
      pl->b_stages = st;
      break;
    }
    max_n /= 2;
  }
  // several CTAs per SM hide the load latency of short K loops: do not take more ring slots than needed
  if (pl->KB <= 6 && pl->b_stages > 2) pl->b_stages = 2;
  FOV_CHECK_ARG(MT * pl->BLOCK_N <= 512, "accumulator tiles exceed tensor memory");
  pl->tmem_cols = (int)tmem_cols_for(MT * pl->BLOCK_N);
  pl->b_off = region_bytes;
  pl->data_bytes = region_bytes + pl->b_stages * pl->b_stage_bytes;
  if (pl->data_bytes < staging) pl->data_bytes = staging;
  pl->data_bytes = (pl->data_bytes + 1023) / 1024 * 1024;
  pl->smem_bytes = (size_t)pl->data_bytes + sizeof(Book) + 1024;
  FOV_CHECK_ARG(pl->smem_bytes <= (size_t)kUsable, "tile does not fit shared memory");
  pl->ws_bytes = (size_t)pl->n_tiles * pl->KB * pl->NS * pl->BLOCK_N * 128;
  return FOV_OK;
}

int pick_vec(const TcSeg& g) {
  auto al = [&](long long m) {
    return ((uintptr_t)g.x % (4 * m) == 0) && (g.pix_stride % m == 0) && (g.img_outer % m == 0) &&
           (g.img_inner % m == 0) && (g.Cin % m == 0);
  };
  if (al(4)) return 4;
  if (al(2)) return 2;
  return 1;
}

template <int NS, int EPI>
int launch_conv(const DevParams& dp, const Plan& pl, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<NS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(227 * 1024));
    if (e != cudaSuccess) {
      fov_set_error("tc_conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  const long long m_tiles = (pl.total_pos + BLOCK_M * pl.MT - 1) / (BLOCK_M * pl.MT);
  if (m_tiles * pl.n_tiles >= (1LL << 31)) { fov_set_error("tc_conv: grid too large"); return FOV_ERR_ARG; }
  dim3 grid((unsigned)(m_tiles * pl.n_tiles));
  tc_conv_kernel<NS, EPI><<<grid, kThreads, pl.smem_bytes, st>>>(dp);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

static int g_tc_debug = 0;
static int g_conv_no_mt2 = 0;
// diagnostics / A-B: two accumulator tiles per CTA in the wide single-term convolutions: 0 = never, 1 = when the grid
// still fills the machine (default), 2 = whenever the tile fits (small test shapes)
extern "C" void fov_debug_conv_mt2(int mode) { g_conv_no_mt2 = mode == 0 ? 1 : (mode == 2 ? -1 : 0); }
// diagnostics (not part of include/fov360.h): per-CTA phase timestamps of the next conv launches
extern "C" void fov_debug_timeline_enable(int on) { g_tc_debug = on; }
extern "C" int fov_debug_timeline_read(unsigned long long* out, int n_words) {
  return (int)cudaMemcpyFromSymbol(out, g_tc_timeline, sizeof(unsigned long long) * (size_t)n_words);
}

// plan of a whole-K-resident shifted-tap GEMM, for the persistent ConvLSTM kernels (convlstm_seq_tc.cu)
int tc_conv_step_plan(const TcConv& c, TcStepPlan* out) {
  Plan pl;
  int rc = make_plan(c, &pl);
  if (rc) return rc;
  FOV_CHECK_ARG(pl.n_tiles == 1, "not a single-n-tile problem");
  *out = TcStepPlan{};
  out->nseg = c.nseg;
  for (int s = 0; s < c.nseg; ++s) {
    const TcSeg& g = c.seg[s];
    const SegPlan& sp = pl.sp[s];
    TcStepSeg& d = out->seg[s];
    d.Cin = g.Cin; d.Cin_p = sp.Cin_p; d.cp_log2 = sp.cp_log2; d.cw = sp.cw; d.lpr_log2 = sp.lpr_log2;
    d.row_bytes = sp.row_bytes; d.term_bytes = sp.term_bytes; d.R = sp.R; d.minshift = sp.minshift;
    d.swz_mask = sp.row_bytes == 128 ? 7 : (sp.row_bytes == 64 ? 3 : 1);
    const uint32_t layout = sp.row_bytes == 128 ? 2u : (sp.row_bytes == 64 ? 4u : 6u);
    d.desc_hi = ((uint32_t)(8 * sp.row_bytes) >> 4) | (1u << 14) | (layout << 29);
    d.taps = sp.taps; d.kw = g.kw; d.dil_h = g.dil_h; d.dil_w = g.dil_w; d.pad_h = g.pad_h; d.pad_w = g.pad_w;
    d.k_begin = sp.k_begin; d.nch = sp.nch;
  }
  out->mode_b = pl.mode_b;
  out->Hp = pl.Hp; out->Wp = pl.Wp; out->PLh = pl.PLh; out->PLw = pl.PLw;
  out->K_total = pl.K_total; out->KB = pl.KB; out->BLOCK_N = pl.BLOCK_N; out->NS = pl.NS;
  out->w_bytes = pl.ws_bytes;
  return FOV_OK;
}

bool tc_conv_supported(const TcConv& c) {
  Plan pl;
  return make_plan(c, &pl) == FOV_OK;
}

size_t tc_conv_ws_bytes(const TcConv& c) {
  Plan pl;
  if (make_plan(c, &pl)) return 0;
  return pl.ws_bytes + 256;
}

int tc_conv_pack(const TcConv& c, cudaStream_t st) {
  Plan pl;
  int rc = make_plan(c, &pl);
  if (rc) return rc;
  FOV_CHECK_ARG(c.ws != nullptr, "NULL workspace");
  PackParams pp{};
  pp.nseg = c.nseg; pp.KB = pl.KB; pp.Cout = c.Cout; pp.BLOCK_N = pl.BLOCK_N; pp.n_tiles = pl.n_tiles; pp.NS = pl.NS;
  pp.out = reinterpret_cast<uint8_t*>(((uintptr_t)c.ws + 255) & ~(uintptr_t)255);
  for (int s = 0; s < c.nseg; ++s) {
    FOV_CHECK_ARG(c.seg[s].w != nullptr, "NULL weights");
    pp.seg[s] = PackSeg{c.seg[s].w, c.seg[s].Cin, pl.sp[s].Cin_p, pl.sp[s].taps, pl.sp[s].k_begin, c.seg[s].w_mode};
  }
  const long long total = (long long)(pl.ws_bytes / 16);
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * fov_num_sms()) blocks = 8LL * fov_num_sms();
  tc_pack_kernel<<<(int)blocks, 256, 0, st>>>(pp);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

int tc_conv_run(const TcConv& c, cudaStream_t st) {
  Plan pl;
  int rc = make_plan(c, &pl);
  if (rc) return rc;
  FOV_CHECK_ARG(c.ws != nullptr, "NULL workspace");
  if (!c.prepacked && (rc = tc_conv_pack(c, st))) return rc;

  DevParams dp{};
  dp.nseg = c.nseg; dp.mode_b = pl.mode_b; dp.dbg = g_tc_debug; dp.n_tiles = pl.n_tiles; dp.MT = pl.MT;
  for (int s = 0; s < c.nseg; ++s) {
    const TcSeg& g = c.seg[s];
    const SegPlan& sp = pl.sp[s];
    DevSeg& d = dp.seg[s];
    // g.x == NULL: the segment reads as zeros (ConvLSTM step 0 with a zero initial state)
    d.x = g.x; d.outer = g.img_outer; d.inner = g.img_inner; d.pix_stride = g.pix_stride;
    d.Cin = g.Cin; d.Cin_p = sp.Cin_p; d.cp_log2 = sp.cp_log2; d.cw = sp.cw; d.nch = sp.nch; d.lpr_log2 = sp.lpr_log2;
    d.kw = g.kw; d.taps = sp.taps; d.dil_h = g.dil_h; d.dil_w = g.dil_w; d.pad_h = g.pad_h; d.pad_w = g.pad_w;
    d.k_begin = sp.k_begin; d.vec = pick_vec(g);
    d.minshift = sp.minshift; d.R = sp.R; d.region_off = sp.region_off; d.nbuf = sp.nbuf;
    d.row_bytes = sp.row_bytes; d.term_bytes = sp.term_bytes;
    d.swz_mask = sp.row_bytes == 128 ? 7 : (sp.row_bytes == 64 ? 3 : 1);
    const uint32_t layout = sp.row_bytes == 128 ? 2u : (sp.row_bytes == 64 ? 4u : 6u);
    d.desc_hi = ((uint32_t)(8 * sp.row_bytes) >> 4) | (1u << 14) | (layout << 29);
    // the staging tables hold 32-bit element offsets
    const long long span = (long long)((c.N_img + c.T_inner - 1) / c.T_inner) * (g.img_outer > 0 ? g.img_outer : 1) +
                           (long long)c.T_inner * (g.img_inner > 0 ? g.img_inner : 0) +
                           (long long)c.H * c.W * g.pix_stride;
    FOV_CHECK_ARG(span < (1LL << 31), "activation tensor too large for 32-bit offsets");
  }
  dp.H = c.H; dp.W = c.W; dp.Hp = pl.Hp; dp.Wp = pl.Wp; dp.PLh = pl.PLh; dp.PLw = pl.PLw; dp.HpWp = pl.Hp * pl.Wp;
  dp.T_inner = c.T_inner; dp.N_img = c.N_img; dp.total_pos = (int)pl.total_pos;
  dp.K_total = pl.K_total; dp.KB = pl.KB; dp.Cout = c.Cout; dp.BLOCK_N = pl.BLOCK_N; dp.b_stages = pl.b_stages;
  dp.tmem_cols = pl.tmem_cols; dp.b_off = pl.b_off; dp.b_stage_bytes = pl.b_stage_bytes; dp.data_bytes = pl.data_bytes;
  dp.wpk = reinterpret_cast<const uint8_t*>(((uintptr_t)c.ws + 255) & ~(uintptr_t)255);
  // software-pipelined producer: multi-chunk segment, aligned rows, one 8-row batch per thread and chunk
  dp.pipe = pl.mode_b && dp.seg[0].vec == 4 && dp.seg[0].x != nullptr &&
            dp.seg[0].R <= 8 * (kProducers >> dp.seg[0].lpr_log2);
  dp.bias = c.bias;
  auto span_ok = [&](long long outer, long long inner, long long pix_stride) {
    return (long long)((c.N_img + c.T_inner - 1) / c.T_inner) * (outer > 0 ? outer : 1) + (long long)c.T_inner * inner +
               (long long)c.H * c.W * pix_stride < (1LL << 31);
  };
  if (c.epi == TC_EPI_CONV) {
    FOV_CHECK_ARG(c.y != nullptr, "NULL output");
    FOV_CHECK_ARG(span_ok(c.y_outer, c.y_inner, c.y_pix_stride), "output tensor too large for 32-bit offsets");
    dp.y = c.y; dp.y_outer = c.y_outer; dp.y_inner = c.y_inner; dp.y_pix_stride = c.y_pix_stride;
    dp.act = c.act; dp.beta = c.beta;
    dp.vec_out = ((uintptr_t)c.y % 16 == 0) && (c.y_pix_stride % 4 == 0) && (c.y_outer % 4 == 0) &&
                 (c.y_inner % 4 == 0) && (c.Cout % 4 == 0) && (!c.bias || (uintptr_t)c.bias % 16 == 0);
  } else {
    FOV_CHECK_ARG(c.bias && c.c_out && c.h_out, "NULL LSTM epilogue pointer");
    FOV_CHECK_ARG(span_ok(c.c_outer, c.c_inner, c.Cout / 4) && span_ok(c.h_outer, c.h_inner, c.h_pix_stride) &&
                      span_ok(c.g_outer, c.g_inner, c.Cout) && span_ok(c.cp_outer, c.cp_inner, c.Cout / 4) &&
                      (long long)c.N_img * c.H * c.W * (c.Cout / 4) < (1LL << 31),
                  "state tensors too large for 32-bit offsets");
    dp.F = c.Cout / 4; dp.rec_act = c.rec_act;
    dp.c_prev = c.c_prev; dp.cp_outer = c.cp_outer; dp.cp_inner = c.cp_inner;
    dp.c_out = c.c_out; dp.c_outer = c.c_outer; dp.c_inner = c.c_inner;
    dp.h_out = c.h_out; dp.h_outer = c.h_outer; dp.h_inner = c.h_inner; dp.h_pix_stride = c.h_pix_stride;
    dp.gates_out = c.gates_out; dp.g_outer = c.g_outer; dp.g_inner = c.g_inner;
    dp.hT = c.hT; dp.cT = c.cT;
    auto a16 = [](const void* q) { return (uintptr_t)q % 16 == 0; };
    FOV_CHECK_ARG(a16(c.c_prev) && a16(c.c_out) && a16(c.h_out) && a16(c.gates_out) && a16(c.hT) && a16(c.cT) &&
                      c.h_pix_stride % 4 == 0 && c.h_outer % 4 == 0 && c.h_inner % 4 == 0 && c.cp_outer % 4 == 0 &&
                      c.cp_inner % 4 == 0 && c.c_outer % 4 == 0 && c.c_inner % 4 == 0 && c.g_outer % 4 == 0 &&
                      c.g_inner % 4 == 0,
                  "LSTM epilogue needs 16-byte aligned state / output tensors");
  }
  if (c.epi == TC_EPI_CONV) {
    if (pl.NS == 1) return launch_conv<1, TC_EPI_CONV>(dp, pl, st);
    if (pl.NS == 2) return launch_conv<2, TC_EPI_CONV>(dp, pl, st);
    return launch_conv<3, TC_EPI_CONV>(dp, pl, st);
  }
  if (pl.NS == 1) return launch_conv<1, TC_EPI_LSTM>(dp, pl, st);
  if (pl.NS == 2) return launch_conv<2, TC_EPI_LSTM>(dp, pl, st);
  return launch_conv<3, TC_EPI_LSTM>(dp, pl, st);
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
namespace {
int conv_from_cfg(const fov_conv_cfg* cfg, int math, TcConv* c) {
  FOV_CHECK_ARG(cfg != nullptr, "cfg is NULL");
  FOV_CHECK_ARG(cfg->N > 0 && cfg->H > 0 && cfg->W > 0 && cfg->Cin > 0 && cfg->Cout > 0, "bad shape");
  FOV_CHECK_ARG(cfg->kh > 0 && cfg->kw > 0 && cfg->dil_h > 0 && cfg->dil_w > 0, "bad kernel/dilation");
  *c = TcConv{};
  c->nseg = 1;
  c->N_img = cfg->N; c->T_inner = 1; c->H = cfg->H; c->W = cfg->W;
  c->math = math;
  c->epi = TC_EPI_CONV;
  c->no_mt2 = g_conv_no_mt2;
  return FOV_OK;
}
void fwd_seg(const fov_conv_cfg* cfg, TcConv* c) {
  TcSeg& g = c->seg[0];
  g.img_outer = cfg->x_img_stride; g.img_inner = 0; g.pix_stride = cfg->x_pix_stride;
  g.Cin = cfg->Cin; g.kh = cfg->kh; g.kw = cfg->kw; g.dil_h = cfg->dil_h; g.dil_w = cfg->dil_w;
  g.pad_h = cfg->pad_h; g.pad_w = cfg->pad_w; g.w_mode = 0;
  c->Cout = cfg->Cout;
}
void bwd_seg(const fov_conv_cfg* cfg, TcConv* c) {
  TcSeg& g = c->seg[0];
  g.img_outer = cfg->y_img_stride; g.img_inner = 0; g.pix_stride = cfg->y_pix_stride;
  g.Cin = cfg->Cout; g.kh = cfg->kh; g.kw = cfg->kw; g.dil_h = cfg->dil_h; g.dil_w = cfg->dil_w;
  g.pad_h = (cfg->kh - 1) * cfg->dil_h - cfg->pad_h; g.pad_w = (cfg->kw - 1) * cfg->dil_w - cfg->pad_w;
  g.w_mode = 1;
  c->Cout = cfg->Cin;
}
}  // namespace

extern "C" size_t fov_conv_tc_ws_bytes(const fov_conv_cfg* cfg, int math, int bwd_data) {
  TcConv c;
  if (conv_from_cfg(cfg, math, &c)) return 0;
  if (bwd_data) bwd_seg(cfg, &c); else fwd_seg(cfg, &c);
  return tc_conv_ws_bytes(c);
}

extern "C" int fov_conv2d_fwd_tc(const fov_conv_cfg* cfg, const float* x, const float* w, const float* bias,
                                 float* y, void* ws, int math, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(x && w && y && ws, "NULL pointer");
  fwd_seg(cfg, &c);
  c.seg[0].x = x; c.seg[0].w = w;
  c.ws = ws;
  c.bias = bias; c.y = y; c.y_outer = cfg->y_img_stride; c.y_inner = 0; c.y_pix_stride = cfg->y_pix_stride;
  c.act = cfg->act; c.beta = cfg->beta;
  return tc_conv_run(c, (cudaStream_t)stream);
}

// Weights packed once, reused by several calls (the 10 decoder steps of convlstm_seq2seq run the same three head
// convolutions: mycode/convlstm_seq2seq.py:211-238): fov_conv_tc_pack fills ws, the *_packed calls read it.
extern "C" int fov_conv_tc_pack(const fov_conv_cfg* cfg, const float* w, void* ws, int math, int bwd_data, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(w && ws, "NULL pointer");
  if (bwd_data) bwd_seg(cfg, &c); else fwd_seg(cfg, &c);
  c.seg[0].w = w;
  c.ws = ws;
  return tc_conv_pack(c, (cudaStream_t)stream);
}

extern "C" int fov_conv2d_fwd_tc_packed(const fov_conv_cfg* cfg, const float* x, const float* bias, float* y,
                                        const void* ws, int math, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(x && y && ws, "NULL pointer");
  fwd_seg(cfg, &c);
  c.seg[0].x = x;
  c.ws = const_cast<void*>(ws); c.prepacked = 1;
  c.bias = bias; c.y = y; c.y_outer = cfg->y_img_stride; c.y_inner = 0; c.y_pix_stride = cfg->y_pix_stride;
  c.act = cfg->act; c.beta = cfg->beta;
  return tc_conv_run(c, (cudaStream_t)stream);
}

extern "C" int fov_conv2d_bwd_data_tc_packed(const fov_conv_cfg* cfg, const float* dy, float* dx, const void* ws,
                                             int math, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(dy && dx && ws, "NULL pointer");
  bwd_seg(cfg, &c);
  c.seg[0].x = dy;
  c.ws = const_cast<void*>(ws); c.prepacked = 1;
  c.bias = nullptr; c.y = dx; c.y_outer = cfg->x_img_stride; c.y_inner = 0; c.y_pix_stride = cfg->x_pix_stride;
  c.act = FOV_ACT_LINEAR; c.beta = cfg->beta;
  return tc_conv_run(c, (cudaStream_t)stream);
}

extern "C" int fov_conv2d_bwd_data_tc(const fov_conv_cfg* cfg, const float* dy, const float* w, float* dx,
                                      void* ws, int math, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(dy && w && dx && ws, "NULL pointer");
  bwd_seg(cfg, &c);
  c.seg[0].x = dy; c.seg[0].w = w;
  c.ws = ws;
  c.bias = nullptr; c.y = dx; c.y_outer = cfg->x_img_stride; c.y_inner = 0; c.y_pix_stride = cfg->x_pix_stride;
  c.act = FOV_ACT_LINEAR; c.beta = cfg->beta;
  return tc_conv_run(c, (cudaStream_t)stream);
}
