// Persistent ConvLSTM2D layer on tensor cores (tcgen05 + TMEM): ONE launch runs every timestep.
//
// Applies when whole images fit a 128-row MMA tile (zero-padded frame Hp*Wp <= 128 positions, e.g. the
// (1, num_user-1) "others" images of mycode/others_LSTM_span_whole.py:80-102).  Samples are independent, so a
// group of G = 128 / (Hp*Wp) images never exchanges data with another group and can live on one SM for the
// whole sequence:
//   * the packed gate weights [K ; R] (every k-block, all bf16 terms) are bulk-copied into shared memory ONCE;
//   * h_{t-1} never leaves the SM: the epilogue of step t-1 writes h as bf16 terms straight into the swizzled
//     operand rows the recurrent taps of step t read (shifted-tap descriptors, see conv_tc.cu);
//   * c_{t-1} stays in the registers of the thread that owns the pixel row;
//   * x_{t+1} is fetched into registers while the MMAs of step t run, and stored once they have completed;
//   * per step and group: one tcgen05.mma chain  D[128 x 4F] = [x_t taps | h_{t-1} taps] x [K ; R]  into the
//     group's TMEM columns, then the gate algebra / cell update on the accumulator (tcgen05.ld) and the stores
//     of h_t (strided, into the channel-concatenated sequence), c_t and - when training - the activated gates.
// A CTA runs NG image groups with separate TMEM columns and barriers, so the epilogue of one group overlaps the
// MMAs of the other.  No per-step launch, no h/c round trip through HBM.
//
// Replaces the ConvLSTM2D time loop Keras runs as a TF while_loop.
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"
#include <type_traits>

// diagnostics / A-B (fov_debug_seq_spread): 1 = keep full image groups at small batches; shared with convlstm_seq_bwd_tc.cu
int g_fov_seq_no_spread = 0;

namespace {

using namespace tc;

int g_seq_wpg = 0;               // diagnostics: force 4 or 8 worker warps per group (0 = choose)

constexpr int kRows = 128;       // D rows (frame positions) of one image group
constexpr int kSeqMaxTaps = 64;
constexpr int kXI = 9;           // float4 of the next input frame each worker thread keeps in flight

struct SeqSeg {
  int Cin, Cin_p, cp_log2, lpr_log2, row_bytes, swz_mask, term_bytes, R, minshift;
  int taps, kw, dil_h, dil_w, pad_h, pad_w, k_begin;
  uint32_t desc_hi;
};

struct SeqParams {
  SeqSeg seg[2];                 // [0] = layer input x_t (HBM), [1] = h_{t-1} (on chip)
  const float* x; long long x_b, x_t; int x_pix, x_vec;
  int B, T, H, W, HW, Hp, Wp, PLh, PLw, HpWp, G, rec_act, training, dbg;
  int K_total, KB;
  uint32_t w_bytes, kb_bytes, grp_bytes, act_off, stg_off, data_bytes, tmem_cols;
  const uint8_t* wpk;
  const float* bias;
  const float *h0, *c0;          // optional dense (B,HW,F)
  float* hseq; long long h_b, h_t; int h_pix;
  float* gates; long long z_b, z_t;       // (B,T,HW,4F) activated gates (training only)
  float* cseq; long long c_b, c_t;        // (B,T,HW,F)
  float *hT, *cT;                // optional dense (B,HW,F)
  // layer wavefront (optional): [image group][T] flags.  wait_flags: x_t is being written by the layer below while this
  // kernel runs - wait for its flag, read it through L2; set_flags: raise the flag once h_t is in global memory
  const int* wait_flags;
  int* set_flags;
};

constexpr int kStgStride = 36;   // floats per row of a warp's 32 x 32 staging tile (144 B: conflict-free float4 rows)
constexpr int kStgBytes = 32 * kStgStride * 4;

constexpr int kSeqMaxSteps = 128;   // k16 steps of the gate GEMM
struct SeqBook {
  uint4 ops[kSeqMaxSteps];           // per k16 step: a_rel (within the group's operand area), a_term, desc_hi, b_rel
  int tapshift[2][kSeqMaxTaps];
  float bias_s[256];             // [pass][gate][8]: the 32 biases a pass needs are contiguous
  uint64_t w_full, a_full[2], tmem_full[2];
  uint32_t tmem_ptr;
  int wave_cnt[2];               // arrivals of a group's worker threads at the current step's flag
};

// diagnostics: cycles CTA 0 spent per phase (fov_debug_seq_read): [0] worker wait tmem_full, [1] phase A, [2] phase B,
// [3] x store + arrive, [4] total worker loop, [5] MMA wait a_full, [6] MMA issue, [7] MMA total
__device__ unsigned long long g_seq_timeline[16];   // [8..12]: phase A split: tmem ld, gate math, h store, tmem st, pair barrier

// WPG = worker warps per image group: 4 (one thread per accumulator row does all F channels) or 8 (the two warps of
// a TMEM lane quarter split the channel passes and the copy-out chunks: twice the threads for the gate algebra)
// WAVE: layer-wavefront instantiation (per-step flags to / from the neighbouring layers' kernels; one group per CTA).  A
// template parameter, not a run-time test: the flag code costs the large-batch instantiations registers and ~3 %.
template <int NS, int F, int NG, int WPG, bool WAVE = false>
__global__ void __launch_bounds__(32 * (WPG * NG + 2), NG == 1 ? 2 : 1) convlstm_seq_fwd_kernel(const SeqParams p) {
  constexpr int kWWarp = WPG * NG, kMmaWarp = WPG * NG + 1, kThr = 32 * (WPG * NG + 2);
  constexpr int N4F = 4 * F;
  constexpr int GT = WPG * 32;                       // worker threads per group
  constexpr int XI = WPG == 8 ? 5 : kXI;             // float4 of the next input frame per worker thread
  constexpr int TW = WPG == 8 ? 16 : 32;             // accumulator columns per copy-out chunk
  constexpr int STS = TW + 4;                        // floats per row of a warp's staging tile
  constexpr int NP = F / 8;                          // 8-channel passes
  constexpr int PPW = WPG == 8 ? (NP + 1) / 2 : NP;  // passes per warp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  SeqBook* bk = reinterpret_cast<SeqBook*>(smem + p.data_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---------------- setup ----------------
  for (int s = 0; s < 2; ++s) {
    const SeqSeg& sg = p.seg[s];
    for (int t = tid; t < sg.taps; t += kThr) {
      const int ty = t / sg.kw, tx = t - ty * sg.kw;
      bk->tapshift[s][t] = (ty * sg.dil_h - sg.pad_h) * p.Wp + (tx * sg.dil_w - sg.pad_w) - sg.minshift;
    }
  }
  if (tid < N4F) bk->bias_s[((tid & (F - 1)) >> 3) * 32 + (tid / F) * 8 + (tid & 7)] = __ldg(&p.bias[tid]);
  if (tid < 2) bk->wave_cnt[tid] = 0;
  if (warp == kMmaWarp && lane == 0) {
    mbar_init(smem_u32(&bk->w_full), 1);
    for (int g = 0; g < NG; ++g) {
      mbar_init(smem_u32(&bk->a_full[g]), GT);
      mbar_init(smem_u32(&bk->tmem_full[g]), 1);
    }
    fence_mbar_init();
  }
  if (warp == kWWarp) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), p.tmem_cols);
    tmem_relinquish();
  }
  // operand rows of pad positions / unused channels stay zero for the whole sequence
  for (uint32_t i = (uint32_t)tid * 16u; i < (uint32_t)NG * p.grp_bytes; i += (uint32_t)kThr * 16u)
    *reinterpret_cast<uint4*>(smem + p.act_off + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  if (warp < WPG * NG) {
    // ---------------- workers: thread = one frame position (D row) [x one half of the channel passes at WPG = 8] ----
    const int g = warp / WPG, wg = warp % WPG, q = wg & 3, half = wg >> 2;
    const int wt = q * 32 + lane;                    // my accumulator row / TMEM lane
    const int gtd = wg * 32 + lane;                  // my index among the group's worker threads
    const int b0 = (blockIdx.x * NG + g) * p.G;
    uint8_t* xreg = smem + p.act_off + (uint32_t)g * p.grp_bytes;
    uint8_t* hreg = xreg + NS * p.seg[0].term_bytes;
    const SeqSeg& sx = p.seg[0];
    const SeqSeg& sh = p.seg[1];
    const int npos = p.G * p.HpWp;

    // my D row -> pixel
    bool valid = false;
    int b = 0, pix = 0;
    if (wt < npos) {
      const int bi = wt / p.HpWp, rem = wt - bi * p.HpWp;
      const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
      const int y = yp - p.PLh, x = xp - p.PLw;
      b = b0 + bi;
      if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && b < p.B) { valid = true; pix = y * p.W + x; }
    }
    // 32-bit element offsets of my row (host-checked), -1 = not a pixel
    const int off_c = valid ? (int)((long long)b * p.c_b + (long long)pix * F) : -1;
    const int off_g = valid ? (int)((long long)b * p.z_b + (long long)pix * N4F) : -1;
    const int off_h = valid ? (int)((long long)b * p.h_b + (long long)pix * p.h_pix) : -1;
    const int off_d = valid ? (int)(((long long)b * p.HW + pix) * F) : -1;
    float* stg = reinterpret_cast<float*>(smem + p.stg_off) + (size_t)warp * (32 * STS);
    // my row of the recurrent operand region, absolute-address swizzle (region bases are 1024-byte aligned)
    const uint32_t hrow = (uint32_t)(wt - sh.minshift) * (uint32_t)sh.row_bytes;

    // input staging items: idx -> (region row, 4-channel slot)
    const int lg = sx.lpr_log2, lpr = 1 << lg;
    const int n_items = sx.R * lpr;
    int goff[XI];
#pragma unroll
    for (int j = 0; j < XI; ++j) {
      goff[j] = -1;
      const int idx = gtd + j * GT;
      if (idx < n_items) {
        const int row = idx >> lg, ch = (idx & (lpr - 1)) * 4;
        const int pos = sx.minshift + row;
        if (pos >= 0 && pos < npos && ch < sx.Cin) {
          const int bi = pos / p.HpWp, rem = pos - bi * p.HpWp;
          const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
          const int y = yp - p.PLh, x = xp - p.PLw;
          if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && b0 + bi < p.B)
            goff[j] = (int)((long long)(b0 + bi) * p.x_b + (long long)(y * p.W + x) * p.x_pix + ch);
        }
      }
    }
    const int wgrp = blockIdx.x * NG + g;            // my image group: index of its wavefront flags
    auto x_load = [&](int t, float4 (&xv)[XI]) {
      const float* xt = p.x + (long long)t * p.x_t;
      if (WAVE && p.wait_flags) wave_wait(p.wait_flags + (long long)wgrp * p.T + t);     // the layer below has stored x_t
#pragma unroll
      for (int j = 0; j < XI; ++j) {
        xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (goff[j] >= 0) {
          const int ch = ((gtd + j * GT) & (lpr - 1)) * 4;
          int nv = sx.Cin - ch;
          nv = nv > 4 ? 4 : nv;
          xv[j] = (WAVE && p.wait_flags) ? ldcg_vec4(xt + goff[j], nv, p.x_vec) : ldg_vec4(xt + goff[j], nv, p.x_vec);
        }
      }
    };
    auto x_store = [&](const float4 (&xv)[XI]) {
#pragma unroll
      for (int j = 0; j < XI; ++j) {
        if (goff[j] >= 0) {
          const int idx = gtd + j * GT;
          const uint32_t a0 = (uint32_t)(idx >> lg) * (uint32_t)sx.row_bytes + (uint32_t)(idx & (lpr - 1)) * 8u;
          const uint32_t so = a0 ^ (((a0 >> 7) & (uint32_t)sx.swz_mask) << 4);
          uint2 pk[NS];
          split4<NS>(xv[j], pk);
#pragma unroll
          for (int s = 0; s < NS; ++s) *reinterpret_cast<uint2*>(xreg + s * sx.term_bytes + so) = pk[s];
        }
      }
    };
    // one 8-channel slice of h for my row -> bf16 terms in the recurrent operand region
    auto h_store = [&](int pass, const float (&hn)[8]) {
      const uint32_t a0 = hrow + (uint32_t)pass * 16u;
      const uint32_t so = a0 ^ (((a0 >> 7) & (uint32_t)sh.swz_mask) << 4);
      uint2 lo[NS], hi[NS];
      split4<NS>(make_float4(hn[0], hn[1], hn[2], hn[3]), lo);
      split4<NS>(make_float4(hn[4], hn[5], hn[6], hn[7]), hi);
#pragma unroll
      for (int s = 0; s < NS; ++s)
        *reinterpret_cast<uint4*>(hreg + s * sh.term_bytes + so) = make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
    };

    // initial state
    // my channel passes: pass = half + 2 * pi (WPG = 8) or pi (WPG = 4); cell state of pass pi in cst[8 * pi ..]
    float cst[PPW * 8];
#pragma unroll
    for (int j = 0; j < PPW * 8; ++j) cst[j] = 0.0f;
#pragma unroll
    for (int pi = 0; pi < PPW; ++pi) {
      const int pass = WPG == 8 ? half + 2 * pi : pi;
      if (pass < NP && valid) {
        if (p.c0) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(p.c0 + off_d + pass * 8));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(p.c0 + off_d + pass * 8 + 4));
          cst[pi * 8 + 0] = v0.x; cst[pi * 8 + 1] = v0.y; cst[pi * 8 + 2] = v0.z; cst[pi * 8 + 3] = v0.w;
          cst[pi * 8 + 4] = v1.x; cst[pi * 8 + 5] = v1.y; cst[pi * 8 + 6] = v1.z; cst[pi * 8 + 7] = v1.w;
        }
        if (p.h0) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(p.h0 + off_d + pass * 8));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(p.h0 + off_d + pass * 8 + 4));
          const float hn[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          h_store(pass, hn);
        }
      }
    }

    float4 xv[XI];
    x_load(0, xv);
    x_store(xv);
    fence_proxy_async_smem();
    mbar_arrive(smem_u32(&bk->a_full[g]));

    // TMEM columns of this group: [0,4F) gate accumulators (then activated gates), [4F,5F) c_t, [5F,6F) h_t
    const uint32_t t_row = tmem_d + (uint32_t)(g * 6 * F) + ((uint32_t)(q * 32) << 16);
    // Coalesced copy-out of WD accumulator columns of the warp's 32 rows: TMEM -> registers (row per thread) ->
    // staging tile -> float4 row segments, 128 / WD rows per store instruction (full 128-byte lines at WD = 32).
    auto copy_out = [&](auto wd_tag, uint32_t col, float* dst, int my_off, float* dst2, int my_off2) {
      constexpr int WD = decltype(wd_tag)::value;
      constexpr int LPR = WD / 4, RPI = 32 / LPR;
      float v[WD];
#pragma unroll
      for (int j = 0; j < WD; j += 8) tmem_ld8(t_row + col + j, v + j);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < WD; j += 4)
        *reinterpret_cast<float4*>(&stg[lane * STS + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      const int rsub = lane / LPR, c4 = (lane % LPR) * 4;
      // all shuffles and tile reads first (independent), then the predicated stores: no per-iteration control flow
      float4 a[LPR];
      int o[LPR], o2[LPR];
#pragma unroll
      for (int it = 0; it < LPR; ++it) {
        const int r = it * RPI + rsub;
        o[it] = __shfl_sync(0xffffffffu, my_off, r);
        o2[it] = __shfl_sync(0xffffffffu, my_off2, r);
        a[it] = *reinterpret_cast<const float4*>(&stg[r * STS + c4]);
      }
#pragma unroll
      for (int it = 0; it < LPR; ++it)
        if (o[it] >= 0) *reinterpret_cast<float4*>(dst + o[it] + c4) = a[it];
      if (dst2) {
#pragma unroll
        for (int it = 0; it < LPR; ++it)
          if (o[it] >= 0) *reinterpret_cast<float4*>(dst2 + o2[it] + c4) = a[it];
      }
      __syncwarp();
    };
    const bool dbg = p.dbg && blockIdx.x == 0 && tid == 0;
    long long tw = 0, ta = 0, tb = 0, tx = 0, c0k = 0, c1k = 0, c2k = 0, c3k = 0, a_ld = 0, a_math = 0, a_hst = 0, a_st = 0;
    const long long t_begin = clock64();
    for (int t = 0; t < p.T; ++t) {
      if (t + 1 < p.T) x_load(t + 1, xv);
      if (dbg) c0k = clock64();
      mbar_wait(smem_u32(&bk->tmem_full[g]), (uint32_t)t & 1u);
      tc_fence_after();
      if (dbg) { c1k = clock64(); tw += c1k - c0k; }
      const bool last = t == p.T - 1;
      // ---- phase A: gate algebra on my row, 8 channels per pass; results go back to TMEM ----
#pragma unroll
      for (int pi = 0; pi < PPW; ++pi) {
        const int pass = WPG == 8 ? half + 2 * pi : pi;
        if (pass >= NP) continue;                     // F = 8 at WPG = 8: the upper warp has no pass
        const int cp = pass * 8;
        float gt[4][8];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) tmem_ld8(t_row + gi * F + cp, gt[gi]);
        const float* bp = &bk->bias_s[pass * 32];
        long long d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        if (dbg) d0 = clock64();
        tmem_ld_wait();
        if (dbg) { d1 = clock64(); a_ld += d1 - d0; }
        float cn[8], hn[8];
        // the recurrent-activation choice is hoisted out of the channel loop: a per-call branch would fence the eight
        // independent channel chains off from each other (no ILP for the MUFU latencies)
        if (p.rec_act == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ai = fminf(fmaxf(0.2f * (gt[0][j] + bp[j]) + 0.5f, 0.0f), 1.0f);
            const float af = fminf(fmaxf(0.2f * (gt[1][j] + bp[8 + j]) + 0.5f, 0.0f), 1.0f);
            const float ag = fast_tanh(gt[2][j] + bp[16 + j]);
            const float ao = fminf(fmaxf(0.2f * (gt[3][j] + bp[24 + j]) + 0.5f, 0.0f), 1.0f);
            cn[j] = af * cst[pi * 8 + j] + ai * ag;
            hn[j] = ao * fast_tanh(cn[j]);
            cst[pi * 8 + j] = cn[j];
            gt[0][j] = ai; gt[1][j] = af; gt[2][j] = ag; gt[3][j] = ao;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ai = __fdividef(1.0f, 1.0f + __expf(-(gt[0][j] + bp[j])));
            const float af = __fdividef(1.0f, 1.0f + __expf(-(gt[1][j] + bp[8 + j])));
            const float ag = fast_tanh(gt[2][j] + bp[16 + j]);
            const float ao = __fdividef(1.0f, 1.0f + __expf(-(gt[3][j] + bp[24 + j])));
            cn[j] = af * cst[pi * 8 + j] + ai * ag;
            hn[j] = ao * fast_tanh(cn[j]);
            cst[pi * 8 + j] = cn[j];
            gt[0][j] = ai; gt[1][j] = af; gt[2][j] = ag; gt[3][j] = ao;
          }
        }
        if (dbg) { d2 = clock64(); a_math += d2 - d1; }
        if (valid) h_store(pass, hn);
        if (dbg) { d3 = clock64(); a_hst += d3 - d2; }
        if (p.training) {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) tmem_st8(t_row + gi * F + cp, gt[gi]);
        }
        tmem_st8(t_row + 4 * F + cp, cn);
        tmem_st8(t_row + 5 * F + cp, hn);
        if (dbg) a_st += clock64() - d3;
      }
      long long e0 = 0;
      if (dbg) e0 = clock64();
      tmem_st_wait();
      if (dbg) a_st += clock64() - e0;
      if (WPG == 8) {
        // the copy-out chunks below read columns written by the partner warp of my lane quarter
        tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g * 4 + q) : "memory");
        tc_fence_after();
      }
      if (dbg) { c2k = clock64(); ta += c2k - c1k; }
      // ---- phase B: coalesced stores of the activated gates, c_t and h_t, TW accumulator columns per chunk; at
      //      WPG = 8 the two warps of a lane quarter take alternate chunks ----
      {
        int k = 0;
        if (p.training) {
          float* g_dst = p.gates + (long long)t * p.z_t;
#pragma unroll
          for (int c0 = 0; c0 < N4F; c0 += TW, ++k)
            if (WPG == 4 || (k & 1) == half)
              copy_out(std::integral_constant<int, TW>{}, (uint32_t)c0, g_dst + c0, off_g, nullptr, 0);
        }
        constexpr int WD = F < TW ? F : TW;
        float* c_dst = p.cseq + (long long)t * p.c_t;
        float* h_dst = p.hseq + (long long)t * p.h_t;
#pragma unroll
        for (int c0 = 0; c0 < F; c0 += WD) {
          if (WPG == 4 || (k & 1) == half)
            copy_out(std::integral_constant<int, WD>{}, (uint32_t)(4 * F + c0), c_dst + c0, off_c,
                     (last && p.cT) ? p.cT + c0 : nullptr, off_d);
          ++k;
          if (WPG == 4 || (k & 1) == half)
            copy_out(std::integral_constant<int, WD>{}, (uint32_t)(5 * F + c0), h_dst + c0, off_h,
                     (last && p.hT) ? p.hT + c0 : nullptr, off_d);
          ++k;
        }
      }
      if (WAVE && p.set_flags) wave_publish_arrive(&bk->wave_cnt[g], GT, p.set_flags + (long long)wgrp * p.T + t);
      if (dbg) { c3k = clock64(); tb += c3k - c2k; }
      if (t + 1 < p.T) {
        x_store(xv);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(smem_u32(&bk->a_full[g]));
      }
      if (dbg) tx += clock64() - c3k;
    }
    if (dbg) {
      g_seq_timeline[0] = tw; g_seq_timeline[1] = ta; g_seq_timeline[2] = tb; g_seq_timeline[3] = tx;
      g_seq_timeline[4] = clock64() - t_begin;
      g_seq_timeline[8] = a_ld; g_seq_timeline[9] = a_math; g_seq_timeline[10] = a_hst; g_seq_timeline[11] = a_st;
    }
  } else if (warp == kWWarp) {
    // ---------------- weights: every k-block, once ----------------
    if (lane == 0) {
      const uint32_t bar = smem_u32(&bk->w_full);
      mbar_arrive_expect_tx(bar, p.w_bytes);
      for (int kb = 0; kb < p.KB; ++kb)
        bulk_g2s(base + (uint32_t)kb * p.kb_bytes, p.wpk + (size_t)kb * p.kb_bytes, p.kb_bytes, bar);
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(kRows, N4F, 0, 0);
      const uint32_t b_term = (uint32_t)N4F * 128u;
      // operand addresses of every k16 step, computed once
      int nsteps = 0;
      for (int kb = 0; kb < p.KB; ++kb)
        for (int k4 = 0; k4 < 4; ++k4) {
          const int k = kb * 64 + k4 * 16;
          if (k >= p.K_total) break;
          const int s = k >= p.seg[1].k_begin ? 1 : 0;
          const SeqSeg& sg = p.seg[s];
          const int kk = k - sg.k_begin;
          const int tap = kk >> sg.cp_log2, c0 = kk & (sg.Cin_p - 1);
          const int ty = tap / sg.kw, tx = tap - ty * sg.kw;
          const int shift = (ty * sg.dil_h - sg.pad_h) * p.Wp + (tx * sg.dil_w - sg.pad_w) - sg.minshift;
          bk->ops[nsteps++] = make_uint4((s ? (uint32_t)NS * p.seg[0].term_bytes : 0u) + (uint32_t)shift * sg.row_bytes +
                                             (uint32_t)c0 * 2u,
                                         (uint32_t)sg.term_bytes, sg.desc_hi, (uint32_t)kb * p.kb_bytes + (uint32_t)k4 * 32u);
        }
      mbar_wait(smem_u32(&bk->w_full), 0);
      const bool dbg = p.dbg && blockIdx.x == 0;
      long long mw = 0, mi = 0;
      const long long m_begin = clock64();
      for (int t = 0; t < p.T; ++t) {
        for (int g = 0; g < NG; ++g) {
          const long long k0 = clock64();
          mbar_wait(smem_u32(&bk->a_full[g]), (uint32_t)t & 1u);
          tc_fence_after();
          const long long k1 = clock64();
          mw += k1 - k0;
          const uint32_t greg = base + p.act_off + (uint32_t)g * p.grp_bytes;
          const uint32_t d = tmem_d + (uint32_t)(g * 6 * F);
          for (int e = 0; e < nsteps; ++e) {
            const uint4 o = bk->ops[e];
#pragma unroll
            for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
              for (int sa = 0; sa <= sum; ++sa) {
                const int sb = sum - sa;
                umma_bf16(d, desc_at(o.z, greg + o.x + sa * o.y), desc_at(kDescHi128, base + o.w + sb * b_term), idesc,
                          (e > 0 || sum != NS - 1 || sa > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(smem_u32(&bk->tmem_full[g]));
          mi += clock64() - k1;
        }
      }
      if (dbg) { g_seq_timeline[5] = mw; g_seq_timeline[6] = mi; g_seq_timeline[7] = clock64() - m_begin; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWWarp) tmem_dealloc(tmem_d, p.tmem_cols);
}

struct SeqPlan {
  TcStepPlan sp;
  int G, NG, WPG;
  uint32_t grp_bytes, act_off, stg_off, data_bytes, tmem_cols;
  size_t smem_bytes;
};

int seq_plan(const fov_convlstm_cfg* c, const TcConv& step, SeqPlan* out) {
  SeqPlan pl{};
  int rc = tc_conv_step_plan(step, &pl.sp);
  if (rc) return rc;
  const TcStepPlan& sp = pl.sp;
  FOV_CHECK_ARG(sp.nseg == 2 && !sp.mode_b, "two-segment step with <= 64 channels per segment expected");
  const int F = c->F;
  FOV_CHECK_ARG(F == 8 || F == 16 || F == 32 || F == 64, "F must be 8/16/32/64");
  pl.G = kRows / (sp.Hp * sp.Wp);
  FOV_CHECK_ARG(pl.G >= 1, "image larger than one MMA tile");
  // Small batches (the reference trains at 32 / 64): the recurrence is latency bound and most SMs idle, so spread the
  // images - as few per group as still fills the machine, one group per CTA.  An MMA step costs the same for 1 or 3
  // images (M = 128 either way) but the gate algebra and the copies of a step shrink with the group.
  const int sms = fov_num_sms() / (c->wave_layers > 1 ? c->wave_layers : 1);    // layer wavefront: the stack shares the SMs
  if (!g_fov_seq_no_spread) {
    const int g_fill = (c->B + sms - 1) / sms;
    if (g_fill < pl.G) pl.G = g_fill < 1 ? 1 : g_fill;
  }
  const bool spread = !g_fov_seq_no_spread && (c->B + pl.G - 1) / pl.G <= sms;
  const int x_items = sp.seg[0].R * (1 << sp.seg[0].lpr_log2);
  FOV_CHECK_ARG(x_items <= kXI * kRows, "input frame too wide to prefetch");
  FOV_CHECK_ARG(sp.seg[0].taps <= kSeqMaxTaps && sp.seg[1].taps <= kSeqMaxTaps && sp.K_total / 16 <= kSeqMaxSteps,
                "too many taps");
  pl.grp_bytes = (uint32_t)sp.NS * (uint32_t)(sp.seg[0].term_bytes + sp.seg[1].term_bytes);
  pl.act_off = (uint32_t)((sp.w_bytes + 1023) / 1024 * 1024);
  const size_t book = sizeof(SeqBook) + 1024;
  const size_t kUsable = 227 * 1024;
  // 8 worker warps per group (16-column staging tiles) when the input frame fits their prefetch registers and the
  // shared memory allows it, else 4 (32-column tiles)
  // preference: configurations that overlap MMA and epilogue (two CTAs per SM, or two groups per CTA) first, 8 worker
  // warps before 4 when a warp gets at least one full pass (F >= 16)
  int found = 0;
  for (int round = 0; round < 2 && !found; ++round) {          // round 0: overlapping configurations only
    for (int wpg = 8; wpg >= 4 && !found; wpg -= 4) {
      if (g_seq_wpg && wpg != g_seq_wpg) continue;
      if (wpg == 8 && (x_items > 5 * 256 || F < 16)) continue;
      const size_t tile = (size_t)32 * ((wpg == 8 ? 16 : 32) + 4) * 4;
      const size_t grp = pl.grp_bytes + wpg * tile;                      // operands + the staging tiles of the group
      const size_t one = pl.act_off + grp + book, two = one + grp;
      int ng = 0;
      if (spread && one <= kUsable) ng = 1;                              // every group gets an SM of its own
      else if (2 * (one + 1024) <= 228 * 1024) ng = 1;                   // two CTAs per SM overlap each other
      else if (two <= kUsable && 2 * 6 * F <= 512) ng = 2;               // one CTA per SM, two groups inside
      else if (round == 1 && one <= kUsable) ng = 1;
      if (!ng) continue;
      found = 1;
      pl.WPG = wpg; pl.NG = ng;
      pl.stg_off = pl.act_off + (uint32_t)ng * pl.grp_bytes;
      pl.data_bytes = pl.stg_off + (uint32_t)(ng * wpg * tile);
    }
  }
  if (!found) { fov_set_error("persistent ConvLSTM: weights + operands exceed shared memory"); return FOV_ERR_ARG; }
  pl.smem_bytes = pl.data_bytes + book;
  pl.tmem_cols = tmem_cols_for(pl.NG * 6 * F);
  // 32-bit element offsets inside the kernel
  const long long HW = (long long)c->H * c->W;
  FOV_CHECK_ARG((long long)c->B * c->x_b_stride < (1LL << 31) && (long long)c->B * c->h_b_stride < (1LL << 31) &&
                    (long long)c->B * c->T * HW * 4 * F < (1LL << 31),
                "tensors too large for 32-bit offsets");
  *out = pl;
  return FOV_OK;
}

template <int NS, int F, int NG, int WPG, bool WAVE = false>
int launch_seq(const SeqParams& p, const SeqPlan& pl, int grid, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(convlstm_seq_fwd_kernel<NS, F, NG, WPG, WAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      fov_set_error("convlstm_seq: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  convlstm_seq_fwd_kernel<NS, F, NG, WPG, WAVE><<<grid, 32 * (WPG * NG + 2), pl.smem_bytes, st>>>(p);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

template <int NS, int F>
int launch_seq_ng(const SeqParams& p, const SeqPlan& pl, int grid, cudaStream_t st) {
  if (p.wait_flags || p.set_flags)        // layer wavefront: one group per CTA (checked by the caller)
    return pl.WPG == 8 ? launch_seq<NS, F, 1, 8, true>(p, pl, grid, st) : launch_seq<NS, F, 1, 4, true>(p, pl, grid, st);
  if (pl.WPG == 8) return pl.NG == 2 ? launch_seq<NS, F, 2, 8>(p, pl, grid, st) : launch_seq<NS, F, 1, 8>(p, pl, grid, st);
  return pl.NG == 2 ? launch_seq<NS, F, 2, 4>(p, pl, grid, st) : launch_seq<NS, F, 1, 4>(p, pl, grid, st);
}
template <int NS>
int launch_seq_f(int F, const SeqParams& p, const SeqPlan& pl, int grid, cudaStream_t st) {
  switch (F) {
    case 8: return launch_seq_ng<NS, 8>(p, pl, grid, st);
    case 16: return launch_seq_ng<NS, 16>(p, pl, grid, st);
    case 32: return launch_seq_ng<NS, 32>(p, pl, grid, st);
    default: return launch_seq_ng<NS, 64>(p, pl, grid, st);
  }
}

}  // namespace

static int g_seq_disable = 0, g_seq_dbg = 0;
extern "C" void fov_debug_seq_wpg(int wpg) { g_seq_wpg = wpg; }
extern "C" void fov_debug_seq_spread(int on) { g_fov_seq_no_spread = !on; }
extern "C" void fov_debug_seq_enable(int on) { g_seq_dbg = on; }
extern "C" int fov_debug_seq_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_seq_timeline, sizeof(unsigned long long) * 16);
}
// diagnostics / A-B testing: 1 = always run the per-timestep launches
extern "C" void fov_debug_convlstm_persistent(int enable) { g_seq_disable = !enable; }

bool tc_convlstm_seq_supported(const fov_convlstm_cfg* c, const TcConv& step) {
  if (g_seq_disable) return false;
  SeqPlan pl;
  const bool ok = seq_plan(c, step, &pl) == FOV_OK;
  fov_set_error("");
  return ok;
}

// image groups (= CTAs) of the persistent forward when it runs one group per CTA, else 0 (fov_convlstm_wave_groups)
int tc_convlstm_seq_wave_groups(const fov_convlstm_cfg* c, const TcConv& step) {
  if (g_seq_disable) return 0;
  SeqPlan pl;
  const bool ok = seq_plan(c, step, &pl) == FOV_OK;
  fov_set_error("");
  if (!ok || pl.NG != 1) return 0;
  return (c->B + pl.G - 1) / pl.G;
}

// step: the fused two-segment step GEMM of this layer (convlstm.cu step_conv) with ws = the packed-weight workspace
int tc_convlstm_seq_fwd(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const TcConv& step, cudaStream_t st) {
  SeqPlan pl;
  int rc = seq_plan(c, step, &pl);
  if (rc) return rc;
  if (!c->ws_prepacked && (rc = tc_conv_pack(step, st))) return rc;
  const TcStepPlan& sp = pl.sp;
  SeqParams p{};
  for (int s = 0; s < 2; ++s) {
    const TcStepSeg& a = sp.seg[s];
    SeqSeg& d = p.seg[s];
    d.Cin = a.Cin; d.Cin_p = a.Cin_p; d.cp_log2 = a.cp_log2; d.lpr_log2 = a.lpr_log2; d.row_bytes = a.row_bytes;
    d.swz_mask = a.swz_mask; d.term_bytes = a.term_bytes; d.R = a.R; d.minshift = a.minshift; d.taps = a.taps;
    d.kw = a.kw; d.dil_h = a.dil_h; d.dil_w = a.dil_w; d.pad_h = a.pad_h; d.pad_w = a.pad_w; d.k_begin = a.k_begin;
    d.desc_hi = a.desc_hi;
  }
  const int HW = c->H * c->W, F = c->F;
  p.x = io->x; p.x_b = c->x_b_stride; p.x_t = c->x_t_stride; p.x_pix = c->x_pix_stride;
  auto al = [&](long long m) {
    return ((uintptr_t)io->x % (4 * m) == 0) && (c->x_pix_stride % m == 0) && (c->x_b_stride % m == 0) &&
           (c->x_t_stride % m == 0) && (c->Cin % m == 0);
  };
  p.x_vec = al(4) ? 4 : (al(2) ? 2 : 1);
  p.B = c->B; p.T = c->T; p.H = c->H; p.W = c->W; p.HW = HW; p.Hp = sp.Hp; p.Wp = sp.Wp; p.PLh = sp.PLh; p.PLw = sp.PLw;
  p.HpWp = sp.Hp * sp.Wp; p.G = pl.G; p.rec_act = c->rec_act; p.training = c->training; p.dbg = g_seq_dbg;
  p.K_total = sp.K_total; p.KB = sp.KB;
  p.w_bytes = (uint32_t)sp.w_bytes; p.kb_bytes = (uint32_t)(sp.NS * sp.BLOCK_N * 128);
  p.grp_bytes = pl.grp_bytes; p.act_off = pl.act_off; p.stg_off = pl.stg_off; p.data_bytes = pl.data_bytes; p.tmem_cols = pl.tmem_cols;
  p.wpk = reinterpret_cast<const uint8_t*>(((uintptr_t)step.ws + 255) & ~(uintptr_t)255);
  p.bias = io->bias; p.h0 = io->h0; p.c0 = io->c0;
  p.hseq = io->hseq; p.h_b = c->h_b_stride; p.h_t = c->h_t_stride; p.h_pix = c->h_pix_stride;
  p.gates = io->gates; p.z_t = (long long)HW * 4 * F; p.z_b = p.z_t * c->T;
  p.cseq = io->cseq; p.c_t = (long long)HW * F; p.c_b = p.c_t * c->T;
  p.hT = io->hT; p.cT = io->cT;
  p.wait_flags = io->wave_wait; p.set_flags = io->wave_set;
  FOV_CHECK_ARG(!(io->wave_wait || io->wave_set) || pl.NG == 1, "wavefront flags need one image group per CTA");
  auto a16 = [](const void* q) { return (uintptr_t)q % 16 == 0; };
  FOV_CHECK_ARG(a16(io->hseq) && a16(io->gates) && a16(io->cseq) && a16(io->hT) && a16(io->cT) && a16(io->h0) &&
                    a16(io->c0) && c->h_pix_stride % 4 == 0 && c->h_b_stride % 4 == 0 && c->h_t_stride % 4 == 0,
                "persistent ConvLSTM needs 16-byte aligned state / output tensors");
  const int groups = (c->B + pl.G - 1) / pl.G;
  const int grid = (groups + pl.NG - 1) / pl.NG;
  switch (sp.NS) {
    case 1: return launch_seq_f<1>(F, p, pl, grid, st);
    case 2: return launch_seq_f<2>(F, p, pl, grid, st);
    default: return launch_seq_f<3>(F, p, pl, grid, st);
  }
}
