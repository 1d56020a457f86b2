// Persistent ConvLSTM2D BPTT on tensor cores (tcgen05 + TMEM): ONE launch walks every timestep backwards.
//
// Same residency idea as the forward kernel (convlstm_seq_tc.cu): a group of G whole images (zero-padded frame,
// G * Hp*Wp <= 128 positions) stays on one SM for the whole sequence.
//   * R^T (the flipped / transposed recurrent kernel, every k-block, all bf16 terms) is bulk-copied to shared
//     memory once;
//   * per step t (T-1 .. 0) the 256 worker threads of the group compute the gate gradients
//         dh = dh_ext[t] + dh_rec ;  dc_t = dh * o * (1 - tanh(c_t)^2) + dc ;  dZ_t = (di, df, dg, do)
//     with a CHANNEL-fastest thread mapping, so every global access (saved gates, c_t, c_{t-1}, dh_ext, dZ_t) is a
//     coalesced float4 row segment; dc stays in registers;
//   * dZ_t is written to HBM (in place of the saved gates: the weight-gradient and input-gradient kernels read
//     it later) AND, split into bf16 terms, into the swizzled operand rows of the recurrent backward-data GEMM
//         dh_rec_{t-1}[pos][F] = sum_taps dZ_t[pos + shift][4F] x R^T        (shifted-tap descriptors, conv_tc.cu)
//     whose accumulator lives in TMEM; the workers read it back through a small shared-memory transpose
//     (accumulator rows are one-row-per-thread, the gate algebra wants channel-fastest);
//   * the next step's saved tensors are fetched into registers BEFORE waiting for the MMAs of the current one.
// No per-step launch (the per-timestep path needs 2 launches per step) and no dh/dc round trip through HBM.
//
// Replaces the reverse ConvLSTM2D time loop of TF1's BPTT.
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

extern int g_fov_seq_no_spread;      // convlstm_seq_tc.cu

namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr int kWorkers = 256;            // warps 0-7: (row, 4-channel) items, channel fastest
constexpr int kBWWarp = 8, kBMmaWarp = 9;
constexpr int kBThr = 320;
constexpr int kMaxSteps = 256;           // k16 steps of the GEMM (taps * 4F / 16)
constexpr int kMaxMmas = 512;            // MMAs of one timestep's chain (k16 steps x term products)
constexpr int kMaxItems = 8;             // (row, c4) items per worker thread: 128 * F/4 / 256

struct BwdParams {
  int B, T, H, W, HW, Hp, Wp, PLh, PLw, HpWp, G, rec_act, dbg;
  int F, Cin_p, nch, row_bytes, swz_mask, term_bytes, R, minshift, taps, kw, pad_h, pad_w;
  uint32_t desc_hi;
  int K_total, KB, BLOCK_N, stack;
  uint32_t w_bytes, kb_bytes, chunk_bytes, act_off, dh_off, mma_off, data_bytes, tmem_cols;
  const uint8_t* wpk;
  float* gates; long long z_b, z_t;              // in: activated gates, out: dZ   (B,T,HW,4F)
  const float* cseq; long long c_b, c_t;         // (B,T,HW,F)
  const float* c0;                               // optional dense (B,HW,F)
  const float* dhseq; long long dh_b, dh_t; int dh_pix;   // optional gradient w.r.t. the hidden sequence
  const float *dhT, *dcT;                        // optional dense (B,HW,F)
  float* dc0;                                    // optional dense (B,HW,F)
  // fused input gradient (optional): dx_t = conv^T(dZ_t, K) rides in the same MMAs as extra accumulator columns
  float* dx; long long x_b, x_t; int x_pix, Cin, Cp, Fp, dx_accumulate;
  const uint8_t* wpk2;                           // packed K^T (same k order), NULL: no dx
  uint32_t kb1_bytes, kb2_bytes;                 // bytes of one (k-block, term) tile of R^T / K^T
  // fused weight gradient (template WG): gK / gR accumulate in TMEM over the whole sequence,
  //   D[n][(tap, c)] += sum_pos dZ_t[pos][n] * [x_t | h_{t-1}][pos + tap][c]
  // with the dZ operand rows that are already in shared memory as the MN-major A operand; x_t / h_{t-1} rows are staged
  // next to them (MN-major B, the kw taps = overlapping N groups, LBO = one row: wgrad_rows_tc.cu)
  const float* xin; long long xi_b, xi_t; int xi_pix, xi_cin;   // layer input (B,T,HW,Cin), strided
  const float* hin; long long hi_b, hi_t; int hi_pix;           // hidden sequence (B,T,HW,F), strided: h_{t-1} = hin[t-1]
  float *gK, *gR, *gB;
  uint32_t x_off, h_off, wtab_off, x_term, h_term, x_desc_hi, h_desc_hi;
  int x_cw, wg_col0, write_dz;
  // layer wavefront (optional): [image group][T] flags.  wait_flags: dhseq[t] is being completed by the BPTT of the
  // layer above (its fused dx) while this kernel runs - wait for its flag, read it through L2; set_flags: raise the flag
  // of step t once this kernel's dx_t has been added into dx
  const int* wait_flags;
  int* set_flags;
};

struct BwdBook {
  uint64_t w_full, a_full, tmem_full, wg_done;
  uint32_t tmem_ptr;
};

__device__ unsigned long long g_bwd_timeline[8];

// DXN: accumulator columns of the fused input gradient each worker thread reads back (0: no fused dx)
// WG: the layer's weight gradient rides in this kernel (one CTA per SM variants only: it needs its own TMEM columns)
// WAVE: layer-wavefront instantiation (per-step flags from / to the neighbouring layers' BPTT kernels); a template
// parameter so that the large-batch instantiations do not carry the flag code
template <int NS, int F, int DXN, bool WG, bool WAVE = false>
__global__ void __launch_bounds__(kBThr, (F <= 16) ? 2 : 1) convlstm_seq_bwd_kernel(const BwdParams p) {
  constexpr int N4F = 4 * F;
  constexpr int LPR = F / 4;                       // float4 per row of an F-channel tensor
  constexpr int NIT = kRows * LPR / kWorkers;      // items per worker thread (1, 2, 4, 8)
  constexpr int RSTEP = kWorkers / LPR;            // rows between the items of a thread
  constexpr int DHS = F + 4;                       // floats per row of the dh transpose tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  BwdBook* bk = reinterpret_cast<BwdBook*>(smem + p.data_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsteps = p.K_total / 16;

  if (warp == kBMmaWarp && lane == 0) {
    mbar_init(smem_u32(&bk->w_full), 1);
    mbar_init(smem_u32(&bk->a_full), kWorkers);
    mbar_init(smem_u32(&bk->tmem_full), 1);
    mbar_init(smem_u32(&bk->wg_done), 1);
    fence_mbar_init();
    // Every MMA of a timestep's chain, fully resolved once (A/B descriptor low words, instruction descriptor,
    // accumulate flag): the issue loop is one 16-byte table read + one tcgen05.mma.  A lone thread computing
    // descriptors costs ~100 cycles per MMA; the tensor pipe needs 45-64 (tests/cuda/tc_mma_rate_probe.cu).
    // k16 step e: k = tap * Cin_p + channel; channel chunk cc = 64-wide operand region.
    uint4* mtab = reinterpret_cast<uint4*>(smem + p.mma_off);
    const uint32_t b_term = (uint32_t)p.BLOCK_N * 128u;
    int n = 0;
    for (int e = 0; e < nsteps; ++e) {
      const int k = e * 16;
      const int tap = k / p.Cin_p, ci0 = k - tap * p.Cin_p;
      const int cc = ci0 >> 6, c0 = ci0 & 63;
      const int ty = tap / p.kw, tx = tap - ty * p.kw;
      const int shift = (ty - p.pad_h) * p.Wp + (tx - p.pad_w) - p.minshift;
      const int kb = p.nch > 1 ? tap * p.nch + cc : k >> 6;
      const uint32_t a_addr = base + p.act_off + (uint32_t)cc * p.chunk_bytes + (uint32_t)shift * p.row_bytes + (uint32_t)c0 * 2u;
      const uint32_t b_addr = base + (uint32_t)kb * p.kb_bytes + (uint32_t)((k >> 4) & 3) * 32u;
      if (p.stack) {
        for (int sa = 0; sa < NS; ++sa)
          mtab[n++] = make_uint4((((a_addr + sa * p.term_bytes) >> 4) & 0x3FFFu) | (1u << 16),
                                 ((b_addr >> 4) & 0x3FFFu) | (1u << 16),
                                 idesc_bf16_f32(kRows, (NS - sa) * p.BLOCK_N, 0, 0), n > 0 ? 1u : 0u);
      } else {
        for (int sum = NS - 1; sum >= 0; --sum)
          for (int sa = 0; sa <= sum; ++sa)
            mtab[n++] = make_uint4((((a_addr + sa * p.term_bytes) >> 4) & 0x3FFFu) | (1u << 16),
                                   (((b_addr + (sum - sa) * b_term) >> 4) & 0x3FFFu) | (1u << 16),
                                   idesc_bf16_f32(kRows, p.BLOCK_N, 0, 0), n > 0 ? 1u : 0u);
      }
    }
  }
  if (warp == kBWWarp) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), p.tmem_cols);
    tmem_relinquish();
  }
  // dZ rows of pad positions stay zero for the whole sequence
  for (uint32_t i = (uint32_t)tid * 16u; i < (uint32_t)p.nch * p.chunk_bytes; i += (uint32_t)kBThr * 16u)
    *reinterpret_cast<uint4*>(smem + p.act_off + i) = make_uint4(0u, 0u, 0u, 0u);
  if (WG) {                                       // x / h operand tiles: pad rows and unused channel slots stay zero
    const uint32_t nb = NS * (p.x_term + p.h_term);
    for (uint32_t i = (uint32_t)tid * 16u; i < nb; i += (uint32_t)kBThr * 16u)
      *reinterpret_cast<uint4*>(smem + p.x_off + i) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  if (warp < kWorkers / 32) {
    // ---------------- workers ----------------
    const int b0 = blockIdx.x * p.G;
    const int npos = p.G * p.HpWp;
    float* dh_s = reinterpret_cast<float*>(smem + p.dh_off);          // [128 rows][F + 4]
    const int c4 = tid % LPR, ch = c4 * 4;
    const int row0 = tid / LPR;
    // my items: rows row0 + j * RSTEP, channels ch .. ch+3
    int off_g[NIT], off_c[NIT], off_h[NIT], off_d[NIT], off_xc[NIT];
    uint32_t so[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
      const int m = row0 + j * RSTEP;
      off_g[j] = off_c[j] = off_h[j] = off_d[j] = off_xc[j] = -1;
      if (m < npos) {
        const int bi = m / p.HpWp, rem = m - bi * p.HpWp;
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const int y = yp - p.PLh, x = xp - p.PLw;
        const int b = b0 + bi;
        if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && b < p.B) {
          const int pix = y * p.W + x;
          off_g[j] = (int)((long long)b * p.z_b + (long long)pix * N4F) + ch;
          off_c[j] = (int)((long long)b * p.c_b + (long long)pix * F) + ch;
          off_h[j] = (int)((long long)b * p.dh_b + (long long)pix * p.dh_pix) + ch;
          off_d[j] = (int)(((long long)b * p.HW + pix) * F) + ch;
          if (DXN > 0) off_xc[j] = (int)((long long)b * p.x_b + (long long)pix * p.x_pix) + ch;
        }
      }
      so[j] = (uint32_t)(m - p.minshift) * (uint32_t)p.row_bytes;      // my row of the dZ operand region
    }
    // fused weight gradient: my (row, column group) items of the x_t / h_{t-1} operand tiles - 128 rows x 8 groups
    // (float2 of x, float4 of h) over 256 threads = 4 rows per thread, one column group
    constexpr int XH = WG ? 4 : 1;
    int off_xi[XH], off_hi[XH];
    uint32_t sx[XH], sh[XH];
    float2 vx[XH];
    float4 vhh[XH];
    if (WG) {
      // bias gradient for free: channel x_cw - 1 of the x tile is a constant 1 on every real pixel row (written once,
      // the per-step stores only touch channels < Cin), so accumulator column (centre tap, x_cw - 1) = sum_pos dZ[pos][n]
      if (tid < kRows && tid < npos) {
        const int m = tid;
        const int bi = m / p.HpWp, rem = m - bi * p.HpWp;
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const int y = yp - p.PLh, x = xp - p.PLw;
        if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && b0 + bi < p.B) {
          const uint32_t a = (uint32_t)(m - p.minshift) * (uint32_t)(p.x_cw * 2) + (uint32_t)(p.x_cw - 1) * 2u;
          const uint32_t xm = p.x_cw * 2 == 128 ? 7u : (p.x_cw * 2 == 64 ? 3u : 1u);
          *reinterpret_cast<uint16_t*>(smem + p.x_off + (a ^ (((a >> 7) & xm) << 4))) = 0x3F80u;     // bf16 1.0, term 0
        }
      }
      const int cgp = tid & 7;
#pragma unroll
      for (int j = 0; j < XH; ++j) {
        const int m = (tid >> 3) + j * 32;
        off_xi[j] = off_hi[j] = -1;
        if (m < npos) {
          const int bi = m / p.HpWp, rem = m - bi * p.HpWp;
          const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
          const int y = yp - p.PLh, x = xp - p.PLw;
          const int b = b0 + bi;
          if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && b < p.B) {
            const int pix = y * p.W + x;
            if (cgp * 2 < p.xi_cin) off_xi[j] = (int)((long long)b * p.xi_b + (long long)pix * p.xi_pix) + cgp * 2;
            if (cgp * 4 < F) off_hi[j] = (int)((long long)b * p.hi_b + (long long)pix * p.hi_pix) + cgp * 4;
          }
        }
        const uint32_t r = (uint32_t)(m - p.minshift);
        const uint32_t ax = r * (uint32_t)(p.x_cw * 2) + (uint32_t)cgp * 4u;     // float2 -> 2 bf16 = 4 bytes
        const uint32_t xm = p.x_cw * 2 == 128 ? 7u : (p.x_cw * 2 == 64 ? 3u : 1u);
        sx[j] = ax ^ (((ax >> 7) & xm) << 4);
        const uint32_t ah = r * (uint32_t)(F * 2) + (uint32_t)cgp * 8u;           // float4 -> 4 bf16 = 8 bytes
        const uint32_t hm = F * 2 == 128 ? 7u : (F * 2 == 64 ? 3u : 1u);
        sh[j] = ah ^ (((ah >> 7) & hm) << 4);
      }
    }
    auto load_xh = [&](int t) {
#pragma unroll
      for (int j = 0; j < XH; ++j) {
        vx[j] = make_float2(0.f, 0.f);
        vhh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (off_xi[j] >= 0) vx[j] = __ldg(reinterpret_cast<const float2*>(p.xin + (long long)t * p.xi_t + off_xi[j]));
        if (off_hi[j] >= 0 && t > 0)
          vhh[j] = __ldg(reinterpret_cast<const float4*>(p.hin + (long long)(t - 1) * p.hi_t + off_hi[j]));
      }
    };
    float4 dc[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
      dc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.dcT && off_d[j] >= 0) dc[j] = __ldg(reinterpret_cast<const float4*>(p.dcT + off_d[j]));
    }
    // accumulator read-back: warp w reads TMEM lanes 32*(w&3).., columns [(w>>2) * F/2, +F/2)
    const int q = warp & 3, half = warp >> 2;
    constexpr int HC = F / 2 < 8 ? 8 : F / 2;                         // columns per warp (8 at F = 8: both halves read all)
    const uint32_t t_row = tmem_d + ((uint32_t)(q * 32) << 16);
    // fused dx: my accumulator row (frame position 32q + lane) and my half of its Cin columns
    constexpr int kDxMax = DXN > 0 ? DXN : 8;
    constexpr int dxn = DXN;                                            // = Cin / 2 columns per warp half
    const int dxc0 = half * dxn;
    const bool dx_me = DXN > 0;
    float dxr[kDxMax];
    auto dx_read = [&]() {            // warp-collective TMEM read of my dx columns (8 at a time)
#pragma unroll
      for (int c = 0; c < kDxMax; c += 8)
        if (c < dxn) tmem_ld8(t_row + (uint32_t)(p.Fp + dxc0 + c), dxr + c);
      tmem_ld_wait();
      if (p.stack) {
#pragma unroll
        for (int sb = 1; sb < NS; ++sb) {
          float w[kDxMax];
#pragma unroll
          for (int c = 0; c < kDxMax; c += 8)
            if (c < dxn) tmem_ld8(t_row + (uint32_t)(sb * p.BLOCK_N + p.Fp + dxc0 + c), w + c);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < kDxMax; ++c)
            if (c < dxn) dxr[c] += w[c];
        }
      }
    };
    // dx leaves through the dh transpose tile in two phases of F columns (Cin = 2F: warp half h holds phase h),
    // so the global read-modify-write is a coalesced, channel-fastest float4 row segment like everything else
    auto dx_store = [&](int t) {
      float* dxt = p.dx + (long long)t * p.x_t;
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");     // every dh read of this step has been done
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        if (half == ph) {
          float* dst = dh_s + (q * 32 + lane) * DHS;
#pragma unroll
          for (int c = 0; c < kDxMax; c += 4)
            if (c < dxn) *reinterpret_cast<float4*>(dst + c) = make_float4(dxr[c], dxr[c + 1], dxr[c + 2], dxr[c + 3]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          if (off_xc[j] >= 0) {
            float4 v = *reinterpret_cast<const float4*>(dh_s + (row0 + j * RSTEP) * DHS + ch);
            float* g = dxt + off_xc[j] + ph * F;
            if (p.dx_accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(g);
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *reinterpret_cast<float4*>(g) = v;
          }
        }
        if (WAVE && p.set_flags && ph == 1) __threadfence();                  // my dx_t stores are device-visible
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      }
      if (WAVE && p.set_flags && tid == 0) {
        __threadfence();
        *reinterpret_cast<volatile int*>(p.set_flags + (long long)blockIdx.x * p.T + t) = 1;
      }
    };
    const bool dbg = (p.dbg & 1) && blockIdx.x == 0 && tid == 0;
    long long tw = 0, tcmp = 0, k0 = 0, k1 = 0, twg = 0, tepi = 0;
    const long long t_begin = clock64();

    float4 vg[NIT][4], vc[NIT], vp[NIT], vh[NIT];
    // saved tensors of item j at step t -> registers
    auto load_item = [&](int t, int j) {
      if (WAVE && p.wait_flags && j == 0) wave_wait(p.wait_flags + (long long)blockIdx.x * p.T + t);   // dhseq[t] is complete
      const float* gt = p.gates + (long long)t * p.z_t;
      const float* ct = p.cseq + (long long)t * p.c_t;
      {
        const bool ok = off_g[j] >= 0;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi)
          vg[j][gi] = ok ? __ldg(reinterpret_cast<const float4*>(gt + off_g[j] + gi * F)) : make_float4(0.f, 0.f, 0.f, 0.f);
        vc[j] = ok ? __ldg(reinterpret_cast<const float4*>(ct + off_c[j])) : make_float4(0.f, 0.f, 0.f, 0.f);
        vp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          if (t > 0) vp[j] = __ldg(reinterpret_cast<const float4*>(ct - p.c_t + off_c[j]));
          else if (p.c0) vp[j] = __ldg(reinterpret_cast<const float4*>(p.c0 + off_d[j]));
        }
        vh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && p.dhseq) {
          const float4* src = reinterpret_cast<const float4*>(p.dhseq + (long long)t * p.dh_t + off_h[j]);
          vh[j] = (WAVE && p.wait_flags) ? __ldcg(src) : __ldg(src);
        }
      }
    };
#pragma unroll
    for (int j = 0; j < NIT; ++j) load_item(p.T - 1, j);
    if (WG) load_xh(p.T - 1);
    for (int t = p.T - 1; t >= 0; --t) {
      // ---- dh_rec of this step: dhT at the last step, else the accumulator of the GEMM issued at step t+1 ----
      float4 dr[NIT];
      if (t == p.T - 1) {
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          dr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.dhT && off_d[j] >= 0) dr[j] = __ldg(reinterpret_cast<const float4*>(p.dhT + off_d[j]));
        }
      } else {
        if (dbg) k0 = clock64();
        mbar_wait(smem_u32(&bk->tmem_full), (uint32_t)(p.T - 2 - t) & 1u);
        tc_fence_after();
        if (dbg) { k1 = clock64(); tw += k1 - k0; }
        if (F >= 16 || half == 0) {
          float v[HC];
#pragma unroll
          for (int c = 0; c < HC; c += 8) tmem_ld8(t_row + (uint32_t)((F >= 16 ? half * HC : 0) + c), v + c);
          tmem_ld_wait();
          if (p.stack) {
#pragma unroll
            for (int sb = 1; sb < NS; ++sb) {
              float w[HC];
#pragma unroll
              for (int c = 0; c < HC; c += 8)
                tmem_ld8(t_row + (uint32_t)(sb * p.BLOCK_N + (F >= 16 ? half * HC : 0) + c), w + c);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < HC; ++c) v[c] += w[c];
            }
          }
          float* dst = dh_s + (q * 32 + lane) * DHS + (F >= 16 ? half * HC : 0);
#pragma unroll
          for (int c = 0; c < HC; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
        if (dx_me) dx_read();                      // dx_{t+1}, stored after this step's operands are handed over
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
#pragma unroll
        for (int j = 0; j < NIT; ++j)
          dr[j] = *reinterpret_cast<const float4*>(dh_s + (row0 + j * RSTEP) * DHS + ch);
      }
      // ---- gate gradients of my items ----
      float* gz = p.gates + (long long)t * p.z_t;
      const bool hard = p.rec_act == FOV_REC_HARD_SIGMOID;
      auto rec_grad = [hard](float a) {             // derivative of the recurrent activation through its output
        const float gh = (a > 0.0f && a < 1.0f) ? 0.2f : 0.0f, gs = a * (1.0f - a);
        return hard ? gh : gs;
      };
#pragma unroll
      for (int j = 0; j < NIT; ++j) {
        const bool ok = off_g[j] >= 0;               // not a pixel: inputs read as zeros, nothing is stored
        const float gi_[4] = {vg[j][0].x, vg[j][0].y, vg[j][0].z, vg[j][0].w};
        const float gf_[4] = {vg[j][1].x, vg[j][1].y, vg[j][1].z, vg[j][1].w};
        const float gg_[4] = {vg[j][2].x, vg[j][2].y, vg[j][2].z, vg[j][2].w};
        const float go_[4] = {vg[j][3].x, vg[j][3].y, vg[j][3].z, vg[j][3].w};
        const float ct_[4] = {vc[j].x, vc[j].y, vc[j].z, vc[j].w};
        const float cp_[4] = {vp[j].x, vp[j].y, vp[j].z, vp[j].w};
        const float dh_[4] = {vh[j].x + dr[j].x, vh[j].y + dr[j].y, vh[j].z + dr[j].z, vh[j].w + dr[j].w};
        float dcv[4] = {dc[j].x, dc[j].y, dc[j].z, dc[j].w};
        float dz[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // branch-free (selects only): the 16 element chains of a thread stay interleavable
          const float tc_ = fast_tanh(ct_[e]);
          const float dog = dh_[e] * tc_;
          const float dct = fmaf(dh_[e] * go_[e], 1.0f - tc_ * tc_, dcv[e]);
          dcv[e] = dct * gf_[e];
          dz[0][e] = dct * gg_[e] * rec_grad(gi_[e]);
          dz[1][e] = dct * cp_[e] * rec_grad(gf_[e]);
          dz[2][e] = dct * gi_[e] * (1.0f - gg_[e] * gg_[e]);
          dz[3][e] = dog * rec_grad(go_[e]);
        }
        dc[j] = make_float4(dcv[0], dcv[1], dcv[2], dcv[3]);
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) vg[j][gi] = make_float4(dz[gi][0], dz[gi][1], dz[gi][2], dz[gi][3]);
      }
      // the weight-gradient MMAs of step t+1 still read the operand rows: they ran under the gate algebra above
      if (WG && t < p.T - 1) {
        const long long w0 = dbg ? clock64() : 0;
        mbar_wait(smem_u32(&bk->wg_done), (uint32_t)(p.T - 2 - t) & 1u);
        if (dbg) twg += clock64() - w0;
      }
#pragma unroll
      for (int j = 0; j < NIT; ++j) {
        const bool ok = off_g[j] >= 0;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const float4 z4 = vg[j][gi];
          if (!ok) continue;
          if (!WG || p.write_dz) *reinterpret_cast<float4*>(gz + off_g[j] + gi * F) = z4;
          // bf16 terms into the operand rows: channel gi*F + ch of chunk (channel / 64)
          const int zc = gi * F + ch;
          const uint32_t a0 = so[j] + (uint32_t)(zc & 63) * 2u;
          const uint32_t sw = a0 ^ (((a0 >> 7) & (uint32_t)p.swz_mask) << 4);
          uint2 pk[NS];
          split4<NS>(z4, pk);
          uint8_t* dstc = smem + p.act_off + (uint32_t)(zc >> 6) * p.chunk_bytes;
#pragma unroll
          for (int s = 0; s < NS; ++s) *reinterpret_cast<uint2*>(dstc + s * p.term_bytes + sw) = pk[s];
        }
      }
      if (WG) {                                    // x_t / h_{t-1} rows of the weight-gradient B operand
#pragma unroll
        for (int j = 0; j < XH; ++j) {
          if (off_xi[j] >= 0) {
            float r0 = vx[j].x, r1 = vx[j].y;
#pragma unroll
            for (int s_ = 0; s_ < NS; ++s_) {
              const uint32_t pk = pack_bf16x2(r0, r1);
              *reinterpret_cast<uint32_t*>(smem + p.x_off + s_ * p.x_term + sx[j]) = pk;
              r0 -= __uint_as_float(pk << 16); r1 -= __uint_as_float(pk & 0xffff0000u);
            }
          }
          if (off_hi[j] >= 0) {
            uint2 pk[NS];
            split4<NS>(vhh[j], pk);
#pragma unroll
            for (int s_ = 0; s_ < NS; ++s_) *reinterpret_cast<uint2*>(smem + p.h_off + s_ * p.h_term + sh[j]) = pk[s_];
          }
        }
      }
      if (t > 0 || p.wpk2 || WG) {
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bk->a_full));
      }
      if (t > 0) {                                 // in flight while the GEMM of this step runs
#pragma unroll
        for (int j = 0; j < NIT; ++j) load_item(t - 1, j);
        if (WG) load_xh(t - 1);
      }
      if (DXN > 0 && t < p.T - 1) dx_store(t + 1);
      if (t == 0 && p.dc0) {
#pragma unroll
        for (int j = 0; j < NIT; ++j)
          if (off_d[j] >= 0) *reinterpret_cast<float4*>(p.dc0 + off_d[j]) = dc[j];
      }
      if (dbg) tcmp += clock64() - k1;
    }
    if (p.wpk2) {                                 // dx_0 from the GEMM issued at step 0
      mbar_wait(smem_u32(&bk->tmem_full), (uint32_t)(p.T - 1) & 1u);
      tc_fence_after();
      if (DXN > 0) { dx_read(); dx_store(0); }
    }
    if (WG) {
      // ---- partial gK / gR / gb of this CTA -> global (lanes = consecutive n: coalesced red.global.add) ----
      const long long e0 = dbg ? clock64() : 0;
      mbar_wait(smem_u32(&bk->wg_done), (uint32_t)(p.T - 1) & 1u);
      tc_fence_after();
      const int n = q * 32 + lane;
      const int kw = p.kw;
      const int xcols = kw * p.x_cw, hcols = kw * F;
      for (int c0 = half * 8; c0 < xcols + hcols; c0 += 16) {
        float v[8];
        tmem_ld8(t_row + (uint32_t)(p.wg_col0 + c0), v);
        tmem_ld_wait();
        if (n < N4F) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            if (c < xcols) {
              const int tap = c / p.x_cw, ci = c - tap * p.x_cw;
              if (ci < p.xi_cin) atomicAdd(p.gK + ((size_t)tap * p.xi_cin + ci) * N4F + n, v[j]);
              else if (ci == p.x_cw - 1 && tap == p.pad_w && p.gB) atomicAdd(p.gB + n, v[j]);      // the ones channel
            } else {
              const int ch_ = c - xcols, tap = ch_ / F, cf = ch_ - tap * F;
              atomicAdd(p.gR + ((size_t)tap * F + cf) * N4F + n, v[j]);
            }
          }
        }
      }
      if (dbg) tepi = clock64() - e0;
    }
    if (dbg) {
      g_bwd_timeline[0] = tw; g_bwd_timeline[1] = tcmp; g_bwd_timeline[2] = clock64() - t_begin;
      g_bwd_timeline[3] = twg; g_bwd_timeline[6] = tepi;
    }
  } else if (warp == kBWWarp) {
    // ---------------- weights, once: B tile of a (k-block, term) = [Fp rows of R^T | Cp rows of K^T] x 128 B ----
    // Copied in 16-row (2048-byte) pieces, one piece per lane and round: larger pieces into the interleaved layout
    // faulted on hardware (4096-byte K^T tiles next to 2048-byte R^T tiles), 2048-byte pieces do not.
    const uint32_t bar = smem_u32(&bk->w_full);
    if (lane == 0) mbar_arrive_expect_tx(bar, p.w_bytes);
    __syncwarp();
    const int ppt = (p.Fp + p.Cp) >> 4;                      // pieces per (k-block, term) tile
    const int npieces = p.KB * NS * ppt;
    for (int i = lane; i < npieces; i += 32) {
      const int tile = i / ppt, r16 = i - tile * ppt;        // tile = kb * NS + s
      const uint32_t dst = base + (uint32_t)tile * (p.kb1_bytes + p.kb2_bytes) + (uint32_t)r16 * 2048u;
      const uint8_t* src = r16 * 16 < p.Fp ? p.wpk + (size_t)tile * p.kb1_bytes + (size_t)r16 * 2048
                                           : p.wpk2 + (size_t)tile * p.kb2_bytes + (size_t)(r16 * 16 - p.Fp) * 128;
      bulk_g2s(dst, src, 2048u, bar);
    }
  } else {
    // ---------------- MMA issuer: dh_rec_{t-1} = conv^T(dZ_t, R), one chain per step ----------------
    if (lane == 0) {
      const int nmma = nsteps * (p.stack ? NS : NS * (NS + 1) / 2);
      const uint4* mtab = reinterpret_cast<const uint4*>(smem + p.mma_off);
      mbar_wait(smem_u32(&bk->w_full), 0);
      const bool dbg = (p.dbg & 1) && blockIdx.x == 0;
      long long mw = 0, mi = 0;
      // fused weight gradient: the MMAs of one step, resolved once: A = dZ rows (MN-major, M = 4F channels, K = 16
      // positions), B = x / h rows with the kw taps as overlapping N groups
      int nwg = 0;
      uint4* wtab = reinterpret_cast<uint4*>(smem + p.wtab_off);
      if (WG) {
        const uint32_t a_lbo = p.nch > 1 ? p.chunk_bytes : 0u;
        const int nk16 = (p.G * p.HpWp + 15) / 16;                 // K steps that hold positions of this group
        for (int k16 = 0; k16 < nk16; ++k16)
          for (int sum = NS - 1; sum >= 0; --sum)
            for (int sa = 0; sa <= sum; ++sa) {
              const int sb = sum - sa;
              const uint64_t ad = smem_desc_sw128(base + p.act_off + (uint32_t)sa * p.term_bytes +
                                                      (uint32_t)(-p.minshift + k16 * 16) * 128u, a_lbo, 1024);
              const uint32_t xrow = (uint32_t)(p.x_cw * 2), hrow = (uint32_t)(F * 2);
              const uint64_t bx = desc_at_lbo(p.x_desc_hi, base + p.x_off + (uint32_t)sb * p.x_term + (uint32_t)(k16 * 16) * xrow, xrow);
              const uint64_t bh = desc_at_lbo(p.h_desc_hi, base + p.h_off + (uint32_t)sb * p.h_term + (uint32_t)(k16 * 16) * hrow, hrow);
              wtab[nwg++] = make_uint4((uint32_t)ad, (uint32_t)bx, idesc_bf16_f32(kRows, p.kw * p.x_cw, 1, 1), (uint32_t)p.wg_col0);
              wtab[nwg++] = make_uint4((uint32_t)ad, (uint32_t)bh, idesc_bf16_f32(kRows, p.kw * F, 1, 1),
                                       (uint32_t)(p.wg_col0 + p.kw * p.x_cw) | 0x80000000u);
            }
      }
      const uint32_t a_hi_wg = (uint32_t)(smem_desc_sw128(0, 0, 1024) >> 32);
      const int t_last = (p.wpk2 || WG) ? 0 : 1;
      for (int t = p.T - 1; t >= t_last; --t) {
        const long long q0 = clock64();
        mbar_wait(smem_u32(&bk->a_full), (uint32_t)(p.T - 1 - t) & 1u);
        tc_fence_after();
        const long long q1 = clock64();
        mw += q1 - q0;
        if (WG && t == 0 && !p.wpk2) {             // no dh_rec / dx chain at the first timestep: weight gradient only
          for (int i = 0; i < nwg; ++i) {
            const uint4 m = wtab[i];
            umma_bf16(tmem_d + (m.w & 0x7fffffffu), ((uint64_t)a_hi_wg << 32) | m.x,
                      ((uint64_t)((m.w >> 31) ? p.h_desc_hi : p.x_desc_hi) << 32) | m.y, m.z, (t == p.T - 1 && i < 2) ? 0u : 1u);
          }
          umma_commit(smem_u32(&bk->wg_done));
          continue;
        }
        // stacked terms: the bf16 terms of a weight tile are adjacent in shared memory, so ONE MMA with
        // N = (NS - sa) * BLOCK_N multiplies A term sa by the B terms 0 .. NS-1-sa (NS MMAs per k step instead of
        // NS (NS + 1) / 2; product (sa, sb) lands in accumulator columns [sb * BLOCK_N, +BLOCK_N), summed at read-back)
#pragma unroll 4
        for (int i = 0; i < nmma; ++i) {
          const uint4 m = mtab[i];
          umma_bf16(tmem_d, ((uint64_t)p.desc_hi << 32) | m.x, ((uint64_t)kDescHi128 << 32) | m.y, m.z, m.w);
        }
        umma_commit(smem_u32(&bk->tmem_full));
        if (WG) {                                  // after the chain: these run while the workers do the next step's gate algebra
          for (int i = 0; i < nwg; ++i) {
            const uint4 m = wtab[i];
            umma_bf16(tmem_d + (m.w & 0x7fffffffu), ((uint64_t)a_hi_wg << 32) | m.x,
                      ((uint64_t)((m.w >> 31) ? p.h_desc_hi : p.x_desc_hi) << 32) | m.y, m.z, (t == p.T - 1 && i < 2) ? 0u : 1u);
          }
          umma_commit(smem_u32(&bk->wg_done));
        }
        mi += clock64() - q1;
      }
      if (dbg) { g_bwd_timeline[4] = mw; g_bwd_timeline[5] = mi; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBWWarp) tmem_dealloc(tmem_d, p.tmem_cols);
}

struct BwdPlan {
  TcStepPlan sp, sp2;
  int G, Fp, Cp, fuse_dx, stack;
  int wg, x_cw;                                  // fused weight gradient
  uint32_t x_off, h_off, wtab_off, x_term, h_term, wg_col0;
  size_t w_bytes;
  uint32_t chunk_bytes, act_off, dh_off, mma_off, data_bytes, tmem_cols;
  size_t smem_bytes;
};

int bwd_plan(const fov_convlstm_cfg* c, const TcConv& rT, const TcConv* kT, BwdPlan* out, bool want_wg = false) {
  BwdPlan pl{};
  int rc = tc_conv_step_plan(rT, &pl.sp);
  if (rc) return rc;
  pl.Fp = pl.sp.BLOCK_N; pl.Cp = 0; pl.fuse_dx = 0;
  if (kT) {
    // dx rides along when K^T has the k order / taps of R^T (same kernel size, no dilation) and Cin splits into
    // 8-column pieces per warp half
    if ((rc = tc_conv_step_plan(*kT, &pl.sp2))) return rc;
    const TcStepSeg &a = pl.sp.seg[0], &b = pl.sp2.seg[0];
    FOV_CHECK_ARG(pl.sp2.nseg == 1 && pl.sp2.K_total == pl.sp.K_total && pl.sp2.KB == pl.sp.KB && b.Cin_p == a.Cin_p &&
                      b.taps == a.taps && b.kw == a.kw && b.pad_h == a.pad_h && b.pad_w == a.pad_w && b.dil_h == 1 &&
                      b.dil_w == 1 && b.minshift == a.minshift && pl.sp2.Hp == pl.sp.Hp && pl.sp2.Wp == pl.sp.Wp &&
                      (c->Cin == 16 || c->Cin == 32) && c->Cin == 2 * c->F,
                  "input gradient cannot share the recurrent GEMM");
    pl.Cp = pl.sp2.BLOCK_N; pl.fuse_dx = 1;
  }
  pl.w_bytes = (size_t)pl.sp.KB * pl.sp.NS * (pl.Fp + pl.Cp) * 128;
  const TcStepPlan& sp = pl.sp;
  const TcStepSeg& sg = sp.seg[0];
  const int F = c->F;
  FOV_CHECK_ARG(sp.nseg == 1, "single segment expected");
  FOV_CHECK_ARG(F == 8 || F == 16 || F == 32 || F == 64, "F must be 8/16/32/64");
  pl.G = kRows / (sp.Hp * sp.Wp);
  FOV_CHECK_ARG(pl.G >= 1, "image larger than one MMA tile");
  if (!g_fov_seq_no_spread) {        // small batches: as few images per CTA as still fills the machine (see convlstm_seq_tc.cu)
    const int sms = fov_num_sms() / (c->wave_layers > 1 ? c->wave_layers : 1);  // layer wavefront: the stack shares the SMs
    const int g_fill = (c->B + sms - 1) / sms;
    if (g_fill < pl.G) pl.G = g_fill < 1 ? 1 : g_fill;
  }
  FOV_CHECK_ARG(sp.K_total / 16 <= kMaxSteps && (sp.K_total / 16) * (sp.NS * (sp.NS + 1) / 2) <= kMaxMmas, "too many k steps");
  FOV_CHECK_ARG(kRows * (F / 4) / kWorkers <= kMaxItems, "too many channels");
  pl.chunk_bytes = (uint32_t)sp.NS * (uint32_t)sg.term_bytes;
  pl.act_off = (uint32_t)((pl.w_bytes + 1023) / 1024 * 1024);
  pl.dh_off = pl.act_off + (uint32_t)sg.nch * pl.chunk_bytes;
  pl.mma_off = pl.dh_off + (uint32_t)(kRows * (F + 4) * 4);
  pl.data_bytes = (pl.mma_off + (uint32_t)((sp.K_total / 16) * (sp.NS * (sp.NS + 1) / 2) * 16) + 1023u) / 1024u * 1024u;
  pl.smem_bytes = pl.data_bytes + sizeof(BwdBook) + 1024;
  FOV_CHECK_ARG(pl.smem_bytes <= 227 * 1024, "persistent BPTT: weights + operands exceed shared memory");
  pl.stack = (sp.NS > 1 && sp.NS * (pl.Fp + pl.Cp) <= 256) ? 1 : 0;
  pl.tmem_cols = tmem_cols_for((pl.stack ? sp.NS : 1) * (pl.Fp + pl.Cp));
  // Fused weight gradient: only the one-CTA-per-SM variant (F = 32: its accumulators need kw * (x_cw + F) more TMEM
  // columns, which two resident CTAs cannot both have), 1 x kw kernels, no fused dx, an even input width <= 16.
  pl.wg = 0;
  if (want_wg && !kT && F == 32 && c->kh == 1 && c->dil_w == 1 && c->Cin <= 14 && c->Cin % 2 == 0 &&
      c->x_pix_stride % 2 == 0 && c->x_b_stride % 2 == 0 && c->x_t_stride % 2 == 0) {
    pl.x_cw = 16;
    const uint32_t base_cols = (uint32_t)((pl.stack ? sp.NS : 1) * (pl.Fp + pl.Cp));
    const uint32_t cols = base_cols + (uint32_t)(c->kw * (pl.x_cw + F));
    const uint32_t R = (uint32_t)sg.R;
    pl.x_term = (R * (uint32_t)(pl.x_cw * 2) + 1023u) / 1024u * 1024u;
    pl.h_term = (R * (uint32_t)(F * 2) + 1023u) / 1024u * 1024u;
    pl.x_off = pl.data_bytes;
    pl.h_off = pl.x_off + (uint32_t)sp.NS * pl.x_term;
    pl.wtab_off = pl.h_off + (uint32_t)sp.NS * pl.h_term;
    const uint32_t nent = (uint32_t)(((pl.G * sp.Hp * sp.Wp + 15) / 16) * (sp.NS * (sp.NS + 1) / 2) * 2);
    const uint32_t data2 = (pl.wtab_off + nent * 16u + 1023u) / 1024u * 1024u;
    if (cols <= 512 && c->kw * (pl.x_cw + F) <= 448 && c->kw * pl.x_cw <= 256 && c->kw * F <= 256 &&
        data2 + sizeof(BwdBook) + 1024 <= 227 * 1024) {
      pl.wg = 1;
      pl.wg_col0 = base_cols;
      pl.tmem_cols = tmem_cols_for((int)cols);
      pl.data_bytes = data2;
      pl.smem_bytes = pl.data_bytes + sizeof(BwdBook) + 1024;
    }
  }
  FOV_CHECK_ARG(!pl.fuse_dx || (long long)c->B * c->x_b_stride < (1LL << 31), "input too large for 32-bit offsets");
  const long long HW = (long long)c->H * c->W;
  FOV_CHECK_ARG((long long)c->B * c->T * HW * 4 * F < (1LL << 31) && (long long)c->B * c->h_b_stride < (1LL << 31),
                "tensors too large for 32-bit offsets");
  *out = pl;
  return FOV_OK;
}

template <int NS, int F, int DXN, bool WG = false, bool WAVE = false>
int launch_bwd(const BwdParams& p, const BwdPlan& pl, int grid, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(convlstm_seq_bwd_kernel<NS, F, DXN, WG, WAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      fov_set_error("convlstm_seq_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  convlstm_seq_bwd_kernel<NS, F, DXN, WG, WAVE><<<grid, kBThr, pl.smem_bytes, st>>>(p);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
template <int NS, int F>
int launch_bwd_dx(const BwdParams& p, const BwdPlan& pl, int grid, cudaStream_t st) {
  if (p.wait_flags || p.set_flags) {      // layer wavefront
    switch (p.wpk2 ? p.Cin / 2 : 0) {
      case 0: return launch_bwd<NS, F, 0, false, true>(p, pl, grid, st);
      case 8: return launch_bwd<NS, F, 8, false, true>(p, pl, grid, st);
      case 16: return launch_bwd<NS, F, 16, false, true>(p, pl, grid, st);
      default: fov_set_error("persistent BPTT: unsupported fused dx width"); return FOV_ERR_UNSUPPORTED;
    }
  }
  switch (p.wpk2 ? p.Cin / 2 : 0) {
    case 0: return launch_bwd<NS, F, 0>(p, pl, grid, st);
    case 8: return launch_bwd<NS, F, 8>(p, pl, grid, st);
    case 16: return launch_bwd<NS, F, 16>(p, pl, grid, st);
    default: fov_set_error("persistent BPTT: unsupported fused dx width"); return FOV_ERR_UNSUPPORTED;
  }
}
template <int NS>
int launch_bwd_f(int F, const BwdParams& p, const BwdPlan& pl, int grid, cudaStream_t st) {
  if (pl.wg) return launch_bwd<NS, 32, 0, true>(p, pl, grid, st);
  switch (F) {
    case 8: return launch_bwd_dx<NS, 8>(p, pl, grid, st);
    case 16: return launch_bwd_dx<NS, 16>(p, pl, grid, st);
    case 32: return launch_bwd_dx<NS, 32>(p, pl, grid, st);
    default: return launch_bwd_dx<NS, 64>(p, pl, grid, st);
  }
}

}  // namespace

// Weight gradient inside the persistent BPTT kernel (template WG): OFF by default.  Measured on B200 (M3 layer 0,
// B = 8880, bf16x2, scripts/bptt_timeline.py): fused 5.28 ms vs 3.10 ms BPTT + 1.81 ms separate weight-gradient launch =
// 4.91 ms.  dZ stays on chip (3.0 GB of writes and 3.5 GB of reads less per step), but the 42 extra MN-major MMAs per
// timestep share the shared-memory port with the workers' transposes and operand stores, whose phase grows from 12.5 k to
// 20 k cycles per step (the workers, not the MMA thread, are the critical path of this kernel).  Kept behind the switch,
// parity-tested (tests/test_gpu_bench_shapes.py), as the starting point for a version with the operands in TMEM.
static int g_bwd_disable = 0, g_bwd_dbg = 0, g_bwd_nostack = 0, g_bwd_no_wg = 1;
extern "C" void fov_debug_seq_bwd_wgrad(int enable) { g_bwd_no_wg = !enable; }
extern "C" void fov_debug_seq_bwd_nostack(int on) { g_bwd_nostack = on; }
extern "C" void fov_debug_convlstm_persistent_bwd(int enable) { g_bwd_disable = !enable; }
extern "C" void fov_debug_seq_bwd_enable(int on) { g_bwd_dbg = on; }
extern "C" int fov_debug_seq_bwd_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_bwd_timeline, sizeof(unsigned long long) * 8);
}

bool tc_convlstm_seq_bwd_supported(const fov_convlstm_cfg* c, const TcConv& rT, const TcConv* kT) {
  if (g_bwd_disable) return false;
  BwdPlan pl;
  const bool ok = bwd_plan(c, rT, kT, &pl) == FOV_OK;
  fov_set_error("");
  return ok;
}

// does the persistent BPTT of this layer also produce its weight gradients (g_kernel / g_recurrent / g_bias)?
bool tc_convlstm_seq_bwd_fuses_wgrad(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const fov_convlstm_grads* gr,
                                     const TcConv& rT, const TcConv* kT) {
  // not with a fused dx (TMEM columns), not when an unfused dx convolution still needs dZ in HBM, not with an initial state
  if (g_bwd_disable || g_bwd_no_wg || kT || gr->dx || io->h0 || !gr->g_kernel || !gr->g_recurrent) return false;
  if ((uintptr_t)io->x % 8 != 0 || (uintptr_t)io->hseq % 16 != 0) return false;
  BwdPlan pl;
  const bool ok = bwd_plan(c, rT, kT, &pl, true) == FOV_OK && pl.wg;
  fov_set_error("");
  return ok;
}

// rT: the recurrent backward-data convolution of this layer (convlstm.cu rec_bwd_conv) with ws = its packed-weight
// workspace.  Runs the whole reverse time loop: gates (in: activated gates, out: dZ), optional dc0.  dh0 is not
// produced (callers that need it use the per-timestep path).  When tc_convlstm_seq_bwd_fuses_wgrad() the weight
// gradients are accumulated too and dZ is not written to HBM (nothing reads it afterwards).
// image groups (= CTAs) of the persistent BPTT kernel, 0 when the configuration does not take it
int tc_convlstm_seq_bwd_wave_groups(const fov_convlstm_cfg* c, const TcConv& rT, const TcConv* kT) {
  if (!tc_convlstm_seq_bwd_supported(c, rT, kT)) return 0;
  BwdPlan pl;
  const bool ok = bwd_plan(c, rT, kT, &pl, false) == FOV_OK;
  fov_set_error("");
  return ok ? (c->B + pl.G - 1) / pl.G : 0;
}

int tc_convlstm_seq_bwd(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const fov_convlstm_grads* gr,
                        const TcConv& rT, const TcConv* kT, cudaStream_t st) {
  const bool want_wg = tc_convlstm_seq_bwd_fuses_wgrad(c, io, gr, rT, kT);
  BwdPlan pl;
  int rc = bwd_plan(c, rT, kT, &pl, want_wg);
  if (rc) return rc;
  if ((rc = tc_conv_pack(rT, st))) return rc;
  if (kT && (rc = tc_conv_pack(*kT, st))) return rc;
  const TcStepPlan& sp = pl.sp;
  const TcStepSeg& sg = sp.seg[0];
  BwdParams p{};
  const int HW = c->H * c->W, F = c->F;
  p.B = c->B; p.T = c->T; p.H = c->H; p.W = c->W; p.HW = HW; p.Hp = sp.Hp; p.Wp = sp.Wp; p.PLh = sp.PLh; p.PLw = sp.PLw;
  p.HpWp = sp.Hp * sp.Wp; p.G = pl.G; p.rec_act = c->rec_act; p.dbg = g_bwd_dbg;
  p.F = F; p.Cin_p = sg.Cin_p; p.nch = sg.nch; p.row_bytes = sg.row_bytes; p.swz_mask = sg.swz_mask;
  p.term_bytes = sg.term_bytes; p.R = sg.R; p.minshift = sg.minshift; p.taps = sg.taps; p.kw = sg.kw;
  p.pad_h = sg.pad_h; p.pad_w = sg.pad_w; p.desc_hi = sg.desc_hi;
  p.K_total = sp.K_total; p.KB = sp.KB; p.BLOCK_N = pl.Fp + pl.Cp; p.stack = g_bwd_nostack ? 0 : pl.stack;
  p.w_bytes = (uint32_t)pl.w_bytes; p.kb_bytes = (uint32_t)(sp.NS * (pl.Fp + pl.Cp) * 128);
  p.kb1_bytes = (uint32_t)pl.Fp * 128u; p.kb2_bytes = (uint32_t)pl.Cp * 128u;
  p.Fp = pl.Fp; p.Cp = pl.Cp; p.Cin = c->Cin;
  if (kT) {
    p.wpk2 = reinterpret_cast<const uint8_t*>(((uintptr_t)kT->ws + 255) & ~(uintptr_t)255);
    p.dx = gr->dx; p.x_b = c->x_b_stride; p.x_t = c->x_t_stride; p.x_pix = c->x_pix_stride;
    p.dx_accumulate = gr->dx_accumulate;
    FOV_CHECK_ARG(gr->dx && (uintptr_t)gr->dx % 16 == 0 && c->x_pix_stride % 4 == 0 && c->x_b_stride % 4 == 0 &&
                      c->x_t_stride % 4 == 0,
                  "fused input gradient needs a 16-byte aligned dx");
  }
  p.chunk_bytes = pl.chunk_bytes; p.act_off = pl.act_off; p.dh_off = pl.dh_off; p.mma_off = pl.mma_off; p.data_bytes = pl.data_bytes;
  p.tmem_cols = pl.tmem_cols;
  if (pl.wg) {
    p.xin = io->x; p.xi_b = c->x_b_stride; p.xi_t = c->x_t_stride; p.xi_pix = c->x_pix_stride; p.xi_cin = c->Cin;
    p.hin = io->hseq; p.hi_b = c->h_b_stride; p.hi_t = c->h_t_stride; p.hi_pix = c->h_pix_stride;
    p.gK = gr->g_kernel; p.gR = gr->g_recurrent; p.gB = gr->g_bias;
    p.x_off = pl.x_off; p.h_off = pl.h_off; p.wtab_off = pl.wtab_off; p.x_term = pl.x_term; p.h_term = pl.h_term;
    p.x_cw = pl.x_cw; p.wg_col0 = (int)pl.wg_col0; p.write_dz = 0;
    auto mn_hi = [](int row_bytes) {           // MN-major descriptor high word: SBO = 8 rows, version 1, swizzle = row width
      const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
      return ((uint32_t)(8 * row_bytes) >> 4) | (1u << 14) | (layout << 29);
    };
    p.x_desc_hi = mn_hi(pl.x_cw * 2); p.h_desc_hi = mn_hi(F * 2);
    FOV_CHECK_ARG((long long)c->B * c->x_b_stride < (1LL << 31), "input too large for 32-bit offsets");
  }
  p.wpk = reinterpret_cast<const uint8_t*>(((uintptr_t)rT.ws + 255) & ~(uintptr_t)255);
  p.gates = io->gates; p.z_t = (long long)HW * 4 * F; p.z_b = p.z_t * c->T;
  p.cseq = io->cseq; p.c_t = (long long)HW * F; p.c_b = p.c_t * c->T;
  p.c0 = io->c0;
  p.dhseq = gr->dhseq; p.dh_b = c->h_b_stride; p.dh_t = c->h_t_stride; p.dh_pix = c->h_pix_stride;
  p.dhT = gr->dhT; p.dcT = gr->dcT; p.dc0 = gr->dc0;
  p.wait_flags = gr->wave_wait; p.set_flags = gr->wave_set;
  FOV_CHECK_ARG(!gr->wave_set || kT, "wavefront: wave_set needs the fused input gradient");
  FOV_CHECK_ARG(!(gr->wave_wait || gr->wave_set) || !pl.wg, "wavefront flags are not available with the in-kernel weight gradient");
  auto a16 = [](const void* q) { return (uintptr_t)q % 16 == 0; };
  FOV_CHECK_ARG(a16(io->gates) && a16(io->cseq) && a16(io->c0) && a16(gr->dhseq) && a16(gr->dhT) && a16(gr->dcT) &&
                    a16(gr->dc0) && c->h_pix_stride % 4 == 0 && c->h_b_stride % 4 == 0 && c->h_t_stride % 4 == 0,
                "persistent BPTT needs 16-byte aligned tensors");
  const int grid = (c->B + pl.G - 1) / pl.G;
  switch (sp.NS) {
    case 1: return launch_bwd_f<1>(F, p, pl, grid, st);
    case 2: return launch_bwd_f<2>(F, p, pl, grid, st);
    default: return launch_bwd_f<3>(F, p, pl, grid, st);
  }
}
