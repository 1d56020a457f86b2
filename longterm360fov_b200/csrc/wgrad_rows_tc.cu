// Fused ConvLSTM2D weight gradient on tensor cores (tcgen05 + TMEM): gK, gR and gb of one layer in ONE launch,
// without any im2col gather.
//
//   gK[tap][ci][n] += sum_{b,t,pix} x_t  [pix + tap][ci] * dZ_t[pix][n]
//   gR[tap][cf][n] += sum_{b,t,pix} h_t-1[pix + tap][cf] * dZ_t[pix][n]        gb[n] += sum dZ_t[pix][n]
//
// As a GEMM the reduction index is the pixel.  Pixels are numbered in the zero-padded linear frame of the
// shifted-tap convolution (conv_tc.cu), 128 consecutive frame positions per pipeline stage, and
//   D[n][(seg, tap, c)] += A[pos][n] * B_seg[pos + shift(tap)][c]
// with BOTH operands MN-major (rows = positions = the MMA K index, contiguous along channels exactly as they
// lie in HBM): A = the dZ tile (M = 4F gate channels), B = the staged activation rows of a segment.  Two
// address tricks verified on hardware (tests/cuda/tc_mnshift_probe.cu) remove the gather:
//   * the swizzle is a function of the absolute shared-memory address, so an operand may start at any row:
//     a kernel tap is a row offset of the SAME staged region;
//   * the stride between the N groups of a descriptor (LBO) may be ONE ROW: the kw taps of a kernel row are
//     the overlapping N groups of a single MMA, N = kw * row width.
// So every activation / dZ element is loaded from HBM and converted to bf16 terms exactly once per layer, and
// a k-step costs one MMA per segment and kernel row (x3 for the bf16x2 cross terms).  The accumulators
// (4F lanes x taps*(Cx+F) columns) stay in TMEM for the whole kernel; each CTA walks a strided set of tiles and
// adds its partial result into the gradients at the end (coalesced red.global.add).
//
// Replaces the ConvLSTM2D kernel / recurrent_kernel / bias gradient ops of TF1's BPTT
// (mycode/others_LSTM_span_whole.py:88-102).
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

// Two CTAs share an SM (launch bounds below): while one waits on its HBM loads the other converts / issues MMAs.
constexpr int kTile = 64;                  // frame positions per stage (4 MMA K-steps of 16)
constexpr int kProd = 256;                 // producer threads (warps 0-7)
constexpr int kTabWarp = 8, kMmaWarp = 9;  // warp 8 owns the TMEM allocation
constexpr int kThr = 320;
constexpr int kStages = 2;
constexpr int kPfTiles = 3;                // L2 prefetch distance, in tiles of this CTA
constexpr int kMaxFrameRows = 192;         // 128 + halo
constexpr int kAGroup = kTile * 128;       // bytes of one 64-channel group of the dZ tile (per term)

struct WrSeg {
  const float* x; long long b_stride, t_stride; int pix_stride;      // images (b, t + t_shift)
  const float* x0; long long x0_b_stride; int x0_pix_stride;         // images used when t + t_shift < 0 (may be NULL)
  int t_shift, Cin, vec, vec0;
  int cw, lpr_log2, row_bytes, swz_mask, term_bytes, R, minshift;
  int kh, kw, dil_h, dil_w, pad_h, pad_w, col0, group_taps;
  uint32_t desc_hi;
  float* gw;
};

struct WrParams {
  WrSeg seg[2];
  int nseg;
  const float* dz; long long z_b, z_t;      // (B,T,HW,Cout) dense rows of Cout floats
  int Cout, a_groups, a_lpr_log2;
  int B, T, H, W, HW, Hp, Wp, PLh, PLw, HpWp;
  long long total_pos;
  int dbg, ntiles, frame_rows, frame_min;        // table rows per tile: positions L0 + frame_min + i
  uint32_t a_term, stage_bytes, data_bytes, tmem_cols;
  float* gbias;
};

struct WrTab {
  int offz[kTile];                          // element offset of the dZ row, -1 = zero row
  int offs[2][kMaxFrameRows];               // per segment: >= 0 main tensor, <= -2 the t<0 tensor (-off-2), -1 zero
};

constexpr int kMaxOps = 64;
constexpr int kSI = 5;                    // activation-row float4 per producer thread, segment and tile
// one MMA chain per (segment, kernel row or tap): everything the issuing thread needs, precomputed once
struct WrOp { uint32_t b_rel, b_term, desc_hi, lbo, dcol, idesc, pad0, pad1; };

struct WrBook {
  WrTab tab[kStages];
  WrOp ops[kMaxOps];
  uint64_t full[kStages], empty[kStages], tmem_full;
  uint32_t tmem_ptr;
};

// diagnostics (CTA 0): [0] producer wait (barrier + empty), [1] loads+dZ stage, [2] segment rows, [3] producer total,
// [4] MMA wait full, [5] MMA issue, [6] MMA total, [7] tiles
__device__ unsigned long long g_wr_timeline[8];

template <int NS>
__global__ void __launch_bounds__(kThr, 2) tc_wgrad_rows_kernel(const WrParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  WrBook* bk = reinterpret_cast<WrBook*>(smem + p.data_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bk->full[s]), kProd);
      mbar_init(smem_u32(&bk->empty[s]), 1);
    }
    mbar_init(smem_u32(&bk->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == kTabWarp) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), p.tmem_cols);
    tmem_relinquish();
  }
  // zero rows / unused channel slots stay zero: clear the operand stages once
  for (uint32_t i = (uint32_t)tid * 16u; i < (uint32_t)kStages * p.stage_bytes; i += (uint32_t)kThr * 16u)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int alg = p.a_lpr_log2, alpr = 1 << alg;          // float4 per dZ row
  const int a_c4 = tid & (alpr - 1);

  if (warp < kProd / 32) {
    // ---------------- producers ----------------
    // dZ tile: thread owns float4 column a_c4 of rows a_row0 + j * a_rstep (a_rstep % 8 == 0: constant swizzle phase)
    const int a_row0 = tid >> alg, a_rstep = kProd >> alg;
    const int a_n = kTile / a_rstep;                       // rows per thread: 8 (Cout 128), 4 (64), 2 (32)
    const uint32_t a_in = (uint32_t)a_row0 * 128u + (uint32_t)(a_c4 & 15) * 8u;
    const uint32_t a_off = (uint32_t)(a_c4 >> 4) * kAGroup + (a_in ^ (((a_in >> 7) & 7u) << 4));
    const bool a_ok = a_c4 * 4 < p.Cout;
    const bool dbg = p.dbg && blockIdx.x == 0 && tid == 0;
    long long tw = 0, tl = 0, ts2 = 0, k0 = 0, k1 = 0, k2 = 0;
    const long long t_begin = clock64();
    for (int i = 0; i < my_tiles; ++i) {
      const int stage = i % kStages;
      const uint32_t ph = (uint32_t)(i / kStages) & 1u;
      if (dbg) k0 = clock64();
      // ---- row table of this tile, built cooperatively: frame row -> element offsets (32-bit math) ----
      WrTab& tb = bk->tab[stage];
      if (tid < p.frame_rows) {
        const int r = tid;
        const long long L = (long long)(blockIdx.x + (long long)i * gridDim.x) * kTile + p.frame_min + r;
        int oz = -1, o0 = -1, o1 = -1;
        if (L >= 0 && L < p.total_pos) {
          const unsigned Lu = (unsigned)L;
          const unsigned n = Lu / (unsigned)p.HpWp, rem = Lu - n * (unsigned)p.HpWp;
          const unsigned yp = rem / (unsigned)p.Wp, xp = rem - yp * (unsigned)p.Wp;
          const int y = (int)yp - p.PLh, x = (int)xp - p.PLw;
          if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W) {
            const int b = (int)(n / (unsigned)p.T), t = (int)n - b * p.T, pix = y * p.W + x;
            oz = (int)((long long)b * p.z_b + (long long)t * p.z_t + (long long)pix * p.Cout);
            for (int s = 0; s < p.nseg; ++s) {
              const WrSeg& sg = p.seg[s];
              const int ts = t + sg.t_shift;
              int o = -1;
              if (ts >= 0) o = (int)((long long)b * sg.b_stride + (long long)ts * sg.t_stride + (long long)pix * sg.pix_stride);
              else if (sg.x0) o = -2 - (int)((long long)b * sg.x0_b_stride + (long long)pix * sg.x0_pix_stride);
              if (s == 0) o0 = o; else o1 = o;
            }
          }
        }
        const int zr = r + p.frame_min;                    // dZ rows are frame rows 0..127 of the tile
        if (zr >= 0 && zr < kTile) tb.offz[zr] = oz;
        tb.offs[0][r] = o0;
        tb.offs[1][r] = o1;
      } else if (tid >= 128 && tid - 128 < p.frame_rows && i + kPfTiles < my_tiles) {
        // otherwise idle threads: L2 prefetch of the rows of the tile kPfTiles ahead (the producers keep one tile of
        // loads in flight in registers; with the rows already in L2 that hides the HBM latency)
        const int r = tid - 128;
        const long long L = (long long)(blockIdx.x + (long long)(i + kPfTiles) * gridDim.x) * kTile + p.frame_min + r;
        if (L >= 0 && L < p.total_pos) {
          const unsigned Lu = (unsigned)L;
          const unsigned n = Lu / (unsigned)p.HpWp, rem = Lu - n * (unsigned)p.HpWp;
          const unsigned yp = rem / (unsigned)p.Wp, xp = rem - yp * (unsigned)p.Wp;
          const int y = (int)yp - p.PLh, x = (int)xp - p.PLw;
          if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W) {
            const int b = (int)(n / (unsigned)p.T), t = (int)n - b * p.T, pix = y * p.W + x;
            const int zr = r + p.frame_min;
            if (zr >= 0 && zr < kTile) {
              const float* z = p.dz + (long long)b * p.z_b + (long long)t * p.z_t + (long long)pix * p.Cout;
              for (int l = 0; l < p.Cout; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(z + l));
            }
            for (int s = 0; s < p.nseg; ++s) {
              const WrSeg& sg = p.seg[s];
              const int ts = t + sg.t_shift;
              if (ts >= 0) {
                const float* xr = sg.x + (long long)b * sg.b_stride + (long long)ts * sg.t_stride + (long long)pix * sg.pix_stride;
                for (int l = 0; l < sg.Cin; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr + l));
              }
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
      mbar_wait(smem_u32(&bk->empty[stage]), ph ^ 1u);
      uint8_t* st = smem + (size_t)stage * p.stage_bytes;
      if (dbg) { k1 = clock64(); tw += k1 - k0; }
      // ---- all loads of the tile first (dZ rows + the activation rows of each segment), then convert + store ----
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < a_n && a_ok) {
          const int off = tb.offz[a_row0 + j * a_rstep];
          if (off >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(p.dz + off + a_c4 * 4));
        }
      }
      float4 vs[2][kSI];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const WrSeg& sg = p.seg[s];
        const int lg = sg.lpr_log2, lpr = 1 << lg;
        const int n_items = s < p.nseg ? sg.R << lg : 0;
        const int fr0 = sg.minshift - p.frame_min;
#pragma unroll
        for (int j = 0; j < kSI; ++j) {
          const int idx = tid + j * kProd;
          vs[s][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < n_items) {
            const int row = idx >> lg, ch = (idx & (lpr - 1)) * 4;
            const int off = tb.offs[s][row + fr0];
            int nv = sg.Cin - ch;
            nv = nv > 4 ? 4 : nv;
            if (nv > 0) {
              if (off >= 0) vs[s][j] = ldg_vec4(sg.x + off + ch, nv, sg.vec);
              else if (off <= -2) vs[s][j] = ldg_vec4(sg.x0 + (-off - 2) + ch, nv, sg.vec0);
            }
          }
        }
      }
      if (dbg) { k2 = clock64(); tl += k2 - k1; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < a_n) {
          bsum.x += v[j].x; bsum.y += v[j].y; bsum.z += v[j].z; bsum.w += v[j].w;
          uint2 pk[NS];
          split4<NS>(v[j], pk);
#pragma unroll
          for (int t = 0; t < NS; ++t)
            *reinterpret_cast<uint2*>(st + t * p.a_term + a_off + (uint32_t)(j * a_rstep) * 128u) = pk[t];
        }
      }
      uint32_t reg_off = NS * p.a_term;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s < p.nseg) {
          const WrSeg& sg = p.seg[s];
          const int lg = sg.lpr_log2, lpr = 1 << lg;
          const int n_items = sg.R << lg;
#pragma unroll
          for (int j = 0; j < kSI; ++j) {
            const int idx = tid + j * kProd;
            if (idx < n_items) {
              const uint32_t a0 = (uint32_t)(idx >> lg) * (uint32_t)sg.row_bytes + (uint32_t)(idx & (lpr - 1)) * 8u;
              const uint32_t so = a0 ^ (((a0 >> 7) & (uint32_t)sg.swz_mask) << 4);
              uint2 pk[NS];
              split4<NS>(vs[s][j], pk);
#pragma unroll
              for (int t = 0; t < NS; ++t) *reinterpret_cast<uint2*>(st + reg_off + t * sg.term_bytes + so) = pk[t];
            }
          }
          reg_off += NS * sg.term_bytes;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bk->full[stage]));
      if (dbg) ts2 += clock64() - k2;
    }
    if (dbg) { g_wr_timeline[0] = tw; g_wr_timeline[1] = tl; g_wr_timeline[2] = ts2; g_wr_timeline[3] = clock64() - t_begin; g_wr_timeline[7] = my_tiles; }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // A = dZ tile, MN-major SWIZZLE_128B: LBO = stride between the 64-channel groups (0: one group, the upper
      // 64 accumulator lanes duplicate the lower ones and are ignored), SBO = 8 rows
      const uint32_t a_lbo = p.a_groups > 1 ? (uint32_t)kAGroup : 0u;
      // op table: one entry per (segment, kernel row | tap)
      int nops = 0;
      {
        uint32_t reg = NS * p.a_term;
        for (int s = 0; s < p.nseg; ++s) {
          const WrSeg& sg = p.seg[s];
          const int nmma = sg.group_taps ? sg.kh : sg.kh * sg.kw;
          const int ncol = sg.group_taps ? sg.kw * sg.cw : sg.cw;
          for (int m = 0; m < nmma; ++m) {
            const int ty = sg.group_taps ? m : m / sg.kw, tx = sg.group_taps ? 0 : m - ty * sg.kw;
            const int shift = (ty * sg.dil_h - sg.pad_h) * p.Wp + (tx * sg.dil_w - sg.pad_w) - sg.minshift;
            WrOp& o = bk->ops[nops++];
            o.b_rel = reg + (uint32_t)shift * sg.row_bytes;
            o.b_term = (uint32_t)sg.term_bytes; o.desc_hi = sg.desc_hi; o.lbo = (uint32_t)sg.row_bytes;
            o.dcol = (uint32_t)(sg.col0 + m * ncol); o.idesc = idesc_bf16_f32(128, ncol, 1, 1);
          }
          reg += NS * sg.term_bytes;
        }
      }
      uint32_t first = 1;
      const bool dbg = p.dbg && blockIdx.x == 0;
      long long mw = 0, mi = 0;
      const long long m_begin = clock64();
      for (int i = 0; i < my_tiles; ++i) {
        const int stage = i % kStages;
        const long long q0 = clock64();
        mbar_wait(smem_u32(&bk->full[stage]), (uint32_t)(i / kStages) & 1u);
        tc_fence_after();
        const long long q1 = clock64();
        mw += q1 - q0;
        const uint32_t a_base = base + (uint32_t)stage * p.stage_bytes;
        for (int e = 0; e < nops; ++e) {
          const uint4 o0 = *reinterpret_cast<const uint4*>(&bk->ops[e]);            // b_rel, b_term, desc_hi, lbo
          const uint2 o1 = *reinterpret_cast<const uint2*>(&bk->ops[e].dcol);      // dcol, idesc
          const uint32_t d = tmem_d + o1.x;
          const uint32_t b0 = a_base + o0.x;
          const uint32_t kstep_b = 16u * o0.w;                                      // 16 rows
#pragma unroll
          for (int k16 = 0; k16 < kTile / 16; ++k16) {
#pragma unroll
            for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
              for (int sa = 0; sa <= sum; ++sa) {
                const int sb = sum - sa;
                const uint64_t ad = smem_desc_sw128(a_base + sa * p.a_term + k16 * 2048, a_lbo, 1024);
                umma_bf16(d, ad, desc_at_lbo(o0.z, b0 + k16 * kstep_b + sb * o0.y, o0.w), o1.y,
                          (first && k16 == 0 && sum == NS - 1 && sa == 0) ? 0u : 1u);
              }
            }
          }
        }
        first = 0;
        umma_commit(smem_u32(&bk->empty[stage]));
        mi += clock64() - q1;
      }
      umma_commit(smem_u32(&bk->tmem_full));
      if (dbg) { g_wr_timeline[4] = mw; g_wr_timeline[5] = mi; g_wr_timeline[6] = clock64() - m_begin; }
    }
  }

  // ---------------- epilogue: add the CTA's partial gradients (producer warps: lane quarter w&3, column part w>>2) ----
  if (warp < kProd / 32 && my_tiles > 0) {
    mbar_wait(smem_u32(&bk->tmem_full), 0);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;
    const int n = q * 32 + lane;
    const uint32_t t_row = tmem_d + ((uint32_t)(q * 32) << 16);
    for (int s = 0; s < p.nseg; ++s) {
      const WrSeg& sg = p.seg[s];
      if (!sg.gw) continue;
      const int taps = sg.kh * sg.kw;
      const int nchunk = taps * sg.cw / 8;
      for (int c = part; c < nchunk; c += kProd / 128) {
        const int col = c * 8, tap = col / sg.cw, ci0 = col - tap * sg.cw;
        if (ci0 >= sg.Cin) continue;
        float v[8];
        tmem_ld8(t_row + (uint32_t)(sg.col0 + col), v);
        tmem_ld_wait();
        if (n < p.Cout) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (ci0 + j < sg.Cin) atomicAdd(sg.gw + ((size_t)tap * sg.Cin + ci0 + j) * p.Cout + n, v[j]);
        }
      }
    }
    if (p.gbias && a_c4 * 4 < p.Cout) {
      atomicAdd(p.gbias + a_c4 * 4 + 0, bsum.x);
      atomicAdd(p.gbias + a_c4 * 4 + 1, bsum.y);
      atomicAdd(p.gbias + a_c4 * 4 + 2, bsum.z);
      atomicAdd(p.gbias + a_c4 * 4 + 3, bsum.w);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTabWarp) tmem_dealloc(tmem_d, p.tmem_cols);
}

int ilog2i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

int pick_vec(const float* ptr, long long a, long long b, long long c, int cin) {
  auto al = [&](long long m) {
    return ((uintptr_t)ptr % (4 * m) == 0) && (a % m == 0) && (b % m == 0) && (c % m == 0) && (cin % m == 0);
  };
  return al(4) ? 4 : (al(2) ? 2 : 1);
}

int plan(const TcWgradRows& c, WrParams* out, size_t* smem_bytes) {
  FOV_CHECK_ARG(c.math >= 1 && c.math <= 3, "math must be 1..3 bf16 terms");
  FOV_CHECK_ARG(c.nseg == 1 || c.nseg == 2, "nseg must be 1 or 2");
  FOV_CHECK_ARG(c.Cout == 32 || c.Cout == 64 || c.Cout == 128, "Cout must be 32/64/128");
  FOV_CHECK_ARG(c.B > 0 && c.T > 0 && c.H > 0 && c.W > 0 && c.dz, "bad shape");
  FOV_CHECK_ARG((uintptr_t)c.dz % 16 == 0, "dZ must be 16-byte aligned");
  WrParams p{};
  p.nseg = c.nseg;
  int PLh = 0, PHh = 0, PLw = 0, PHw = 0;
  for (int s = 0; s < c.nseg; ++s) {
    const TcWgradRowsSeg& g = c.seg[s];
    FOV_CHECK_ARG(g.x && g.Cin > 0 && g.Cin <= 64 && g.kh > 0 && g.kw > 0 && g.dil_h > 0 && g.dil_w > 0, "bad segment");
    const int hh = (g.kh - 1) * g.dil_h - g.pad_h, hw = (g.kw - 1) * g.dil_w - g.pad_w;
    PLh = g.pad_h > PLh ? g.pad_h : PLh; PLw = g.pad_w > PLw ? g.pad_w : PLw;
    PHh = hh > PHh ? hh : PHh; PHw = hw > PHw ? hw : PHw;
  }
  p.B = c.B; p.T = c.T; p.H = c.H; p.W = c.W; p.HW = c.H * c.W;
  p.PLh = PLh; p.PLw = PLw; p.Hp = c.H + PLh + PHh; p.Wp = c.W + PLw + PHw; p.HpWp = p.Hp * p.Wp;
  p.total_pos = (long long)c.B * c.T * p.HpWp;
  p.ntiles = (int)((p.total_pos + kTile - 1) / kTile);
  p.dz = c.dz; p.z_t = (long long)p.HW * c.Cout; p.z_b = p.z_t * c.T;
  p.Cout = c.Cout; p.a_groups = c.Cout > 64 ? 2 : 1; p.a_lpr_log2 = ilog2i(c.Cout / 4);
  p.a_term = (uint32_t)p.a_groups * kAGroup;
  p.gbias = c.gbias;
  FOV_CHECK_ARG((long long)c.B * c.T * p.HW * c.Cout < (1LL << 31) && p.total_pos < (1LL << 31) - 2 * kMaxFrameRows,
                "dZ too large for 32-bit offsets");
  int col = 0, fmin = 0, fmax = 0;
  uint32_t stage = (uint32_t)c.math * p.a_term;
  for (int s = 0; s < c.nseg; ++s) {
    const TcWgradRowsSeg& g = c.seg[s];
    WrSeg& d = p.seg[s];
    d.x = g.x; d.b_stride = g.b_stride; d.t_stride = g.t_stride; d.pix_stride = g.pix_stride;
    d.x0 = g.x0; d.x0_b_stride = g.x0_b_stride; d.x0_pix_stride = g.x0_pix_stride;
    d.t_shift = g.t_shift; d.Cin = g.Cin;
    d.vec = pick_vec(g.x, g.b_stride, g.t_stride, g.pix_stride, g.Cin);
    d.vec0 = g.x0 ? pick_vec(g.x0, g.x0_b_stride, 0, g.x0_pix_stride, g.Cin) : 1;
    d.cw = g.Cin <= 16 ? 16 : (g.Cin <= 32 ? 32 : 64);
    d.lpr_log2 = ilog2i(d.cw / 4);
    d.row_bytes = d.cw * 2;
    d.swz_mask = d.row_bytes == 128 ? 7 : (d.row_bytes == 64 ? 3 : 1);
    d.kh = g.kh; d.kw = g.kw; d.dil_h = g.dil_h; d.dil_w = g.dil_w; d.pad_h = g.pad_h; d.pad_w = g.pad_w;
    d.minshift = -(g.pad_h * p.Wp + g.pad_w);
    const int maxshift = ((g.kh - 1) * g.dil_h - g.pad_h) * p.Wp + ((g.kw - 1) * g.dil_w - g.pad_w);
    d.R = (kTile + maxshift - d.minshift + 7) / 8 * 8;
    d.term_bytes = (d.R * d.row_bytes + 1023) / 1024 * 1024;
    // the kw taps of a kernel row as overlapping N groups of one MMA (LBO = one row) when they are adjacent rows
    d.group_taps = (g.dil_w == 1 && g.kw * d.cw <= 256) ? 1 : 0;
    const uint32_t layout = d.row_bytes == 128 ? 2u : (d.row_bytes == 64 ? 4u : 6u);
    // MN-major descriptor high word: SBO = 8 rows (the low word carries the address and LBO = one row)
    d.desc_hi = ((uint32_t)(8 * d.row_bytes) >> 4) | (1u << 14) | (layout << 29);
    d.col0 = col;
    col += g.kh * g.kw * d.cw;
    d.gw = g.gw;
    fmin = d.minshift < fmin ? d.minshift : fmin;
    fmax = maxshift > fmax ? maxshift : fmax;
    stage += (uint32_t)c.math * (uint32_t)d.term_bytes;
    const long long span = (long long)c.B * (g.b_stride > 0 ? g.b_stride : 1) + (long long)c.T * g.t_stride;
    FOV_CHECK_ARG(span < (1LL << 31), "activation tensor too large for 32-bit offsets");
  }
  FOV_CHECK_ARG(col <= 512, "taps x channels exceed the 512 TMEM columns");
  int nops = 0;
  for (int s = 0; s < c.nseg; ++s) {
    nops += p.seg[s].group_taps ? p.seg[s].kh : p.seg[s].kh * p.seg[s].kw;
    FOV_CHECK_ARG((p.seg[s].R << p.seg[s].lpr_log2) <= kSI * kProd, "activation rows too wide for the producer batch");
  }
  FOV_CHECK_ARG(nops <= kMaxOps, "too many kernel rows / taps");
  p.frame_min = fmin;
  p.frame_rows = (kTile + fmax - fmin + 7) / 8 * 8;
  FOV_CHECK_ARG(p.frame_rows <= kMaxFrameRows, "halo too large");
  for (int s = 0; s < c.nseg; ++s)
    FOV_CHECK_ARG(p.seg[s].R + p.seg[s].minshift - fmin <= p.frame_rows, "internal: region outside the frame table");
  p.stage_bytes = (stage + 1023) / 1024 * 1024;
  p.data_bytes = kStages * p.stage_bytes;
  p.tmem_cols = tmem_cols_for(col);
  *smem_bytes = (size_t)p.data_bytes + sizeof(WrBook) + 1024;
  FOV_CHECK_ARG(*smem_bytes <= 227 * 1024, "stages do not fit shared memory");
  *out = p;
  return FOV_OK;
}

template <int NS>
int launch(const WrParams& p, int grid, size_t smem, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_rows_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      fov_set_error("tc_wgrad_rows: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  tc_wgrad_rows_kernel<NS><<<grid, kThr, smem, st>>>(p);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

static int g_wr_disable = 0, g_wr_dbg = 0, g_wr_full_grid = 0;
// diagnostics / A-B: 1 = one tile per CTA even when there are few tiles
extern "C" void fov_debug_wgrad_rows_full_grid(int on) { g_wr_full_grid = on; }
extern "C" void fov_debug_wgrad_rows_timeline(int on) { g_wr_dbg = on; }
extern "C" int fov_debug_wgrad_rows_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_wr_timeline, sizeof(unsigned long long) * 8);
}
extern "C" void fov_debug_wgrad_rows(int enable) { g_wr_disable = !enable; }

bool tc_wgrad_rows_supported(const TcWgradRows& c) {
  if (g_wr_disable) return false;
  WrParams p;
  size_t smem;
  const bool ok = plan(c, &p, &smem) == FOV_OK;
  fov_set_error("");
  return ok;
}

int tc_wgrad_rows_run(const TcWgradRows& c, cudaStream_t st) {
  WrParams p;
  size_t smem;
  int rc = plan(c, &p, &smem);
  if (rc) return rc;
  p.dbg = g_wr_dbg;
  int grid = p.ntiles < 2 * fov_num_sms() ? p.ntiles : 2 * fov_num_sms();
  // every CTA ends with a red.global.add of its whole accumulator (4F x taps*(Cx+F) floats onto the same addresses):
  // with few tiles (small batches) give each CTA at least kMinTiles of them - fewer, longer CTAs beat 296 contending
  // epilogues (B=32: 94 us per layer with one tile per CTA)
  constexpr int kMinTiles = 4;
  if (!g_wr_full_grid && p.ntiles < kMinTiles * grid) grid = (p.ntiles + kMinTiles - 1) / kMinTiles;
  if (grid < 1) grid = 1;
  switch (c.math) {
    case 1: return launch<1>(p, grid, smem, st);
    case 2: return launch<2>(p, grid, smem, st);
    default: return launch<3>(p, grid, smem, st);
  }
}
