// Tensor-core (tcgen05 + TMEM) implicit-GEMM convolution family for sm_100a.
//
//   D[m][n] = sum_k A(m,k) * Wp[k][n]      m = pixel (image, y, x), n = output channel,
//                                          k = (segment, tap, channel) of the im2col row
//
// * A is never materialised in HBM: 128 producer threads gather the im2col rows of a
//   128-pixel tile straight from the fp32 NHWC activations (coalesced along channels),
//   split each value into 1..3 bf16 terms and store them into the 128-byte-swizzled
//   K-major shared-memory tiles tcgen05.mma reads.
// * The weights are repacked once per call (tc_pack_kernel) into the exact shared-memory
//   image of each (n-tile, k-block) B tile, so one TMA-engine bulk copy
//   (cp.async.bulk -> UBLKCP) per pipeline stage brings them in, completion on an mbarrier.
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N<=256, K=16 per instruction),
//   accumulating in TMEM in fp32; with NS bf16 terms per operand it issues the NS(NS+1)/2
//   cross products whose weight is above 2^-8NS (bf16x3 ~ fp32 accuracy on tensor cores).
// * Epilogues read the accumulator with tcgen05.ld and write through a per-warp
//   shared-memory transpose so every global access is a full 128-byte line:
//     TC_EPI_CONV: y = act(acc + bias + beta*y)
//     TC_EPI_LSTM: ConvLSTM2D gate algebra + cell update fused (keras ConvLSTM2D step,
//                  mycode/others_LSTM_span_whole.py:88-100, mycode/convlstm_seq2seq.py:100-126)
//
// Also here: the tensor-core weight-gradient kernel (MN-major operands, split over pixels).
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                    // bf16 per 128-byte swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * 128;    // 16 KB per bf16 term
constexpr int kProducers = 128;
constexpr int kThreads = 192;                  // 4 producer/epilogue warps + B-copy warp + MMA warp
constexpr int kMaxStages = 4;
constexpr int CONV_RS = 36;                    // conv epilogue staging row stride (floats): 32 cols + pad
enum { O_OUT = 0, O_H = 1, O_GATES = 2, O_CPREV = 3, O_DENSE = 4 };

struct DevSeg {
  const float* x;
  int pix_stride, Cin, Cin_p, kw, taps, dil_h, dil_w, pad_h, pad_w, k_begin, vec;
};

struct DevParams {
  DevSeg seg[2];
  int nseg;
  long long x_outer[2], x_inner[2];
  int H, W, HW, T_inner;
  long long M;
  int KB, Cout, BLOCK_N, stages, tmem_cols, stage_bytes, data_bytes;
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

// shared-memory bookkeeping placed after the data region (pipeline stages / epilogue staging)
struct Book {
  const float* rp[2][BLOCK_M];     // per row, per segment: address of the row's own pixel (channel 0)
  long long off_o[5][BLOCK_M];     // epilogue element offsets per row (image + pixel), see O_*
  int pyx[BLOCK_M];                // (y << 16) | x of the row's pixel; y = 0x4000 marks a row past M
  uint64_t full[kMaxStages], empty[kMaxStages], tmem_full;
  uint32_t tmem_ptr;
};

// element offset of image n in an (outer, inner) strided tensor
__device__ __forceinline__ long long img_off(int n, int T_inner, long long outer, long long inner) {
  const int no = n / T_inner;
  return (long long)no * outer + (long long)(n - no * T_inner) * inner;
}

// exp-based activations for the fused epilogue (abs error ~1e-7, far inside the bf16-split budget)
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float fast_rec(int rec, float x) {
  if (rec == FOV_REC_HARD_SIGMOID) return fov_hard_sigmoid(x);
  return __fdividef(1.0f, 1.0f + __expf(-x));
}

template <int NS, int EPI>
__global__ void __launch_bounds__(kThreads, 3) tc_conv_kernel(const DevParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  Book* bk = reinterpret_cast<Book*>(smem + p.data_bytes);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * BLOCK_M;
  const int n_tile = blockIdx.y;
  const int n0 = n_tile * p.BLOCK_N;
  const int S = p.stages;
  const uint32_t b_tile_bytes = (uint32_t)p.BLOCK_N * 128u;

  // ---------------- setup ----------------
  if (tid < BLOCK_M) {
    const long long m = m0 + tid;
    if (m < p.M) {
      const int n = (int)((unsigned)m / (unsigned)p.HW);     // M < 2^31 (checked on the host)
      const int pix = (int)m - n * p.HW;
      bk->pyx[tid] = ((pix / p.W) << 16) | (pix - (pix / p.W) * p.W);
      bk->rp[0][tid] = p.seg[0].x + img_off(n, p.T_inner, p.x_outer[0], p.x_inner[0]) +
                       (long long)pix * p.seg[0].pix_stride;
      bk->rp[1][tid] = p.nseg > 1 ? p.seg[1].x + img_off(n, p.T_inner, p.x_outer[1], p.x_inner[1]) +
                                        (long long)pix * p.seg[1].pix_stride
                                  : nullptr;
      if (EPI == TC_EPI_CONV) {
        bk->off_o[O_OUT][tid] = img_off(n, p.T_inner, p.y_outer, p.y_inner) + (long long)pix * p.y_pix_stride;
      } else {
        bk->off_o[O_OUT][tid] = img_off(n, p.T_inner, p.c_outer, p.c_inner) + (long long)pix * p.F;
        bk->off_o[O_H][tid] = img_off(n, p.T_inner, p.h_outer, p.h_inner) + (long long)pix * p.h_pix_stride;
        bk->off_o[O_GATES][tid] = img_off(n, p.T_inner, p.g_outer, p.g_inner) + (long long)pix * 4 * p.F;
        bk->off_o[O_CPREV][tid] = img_off(n, p.T_inner, p.cp_outer, p.cp_inner) + (long long)pix * p.F;
        bk->off_o[O_DENSE][tid] = ((long long)n * p.HW + pix) * p.F;
      }
    } else {
      bk->pyx[tid] = 0x4000 << 16;
      bk->rp[0][tid] = nullptr; bk->rp[1][tid] = nullptr;
    }
  }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bk->full[s]), kProducers + 1);
      mbar_init(smem_u32(&bk->empty[s]), 1);
    }
    mbar_init(smem_u32(&bk->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  if (warp < 4) {
    // ---------------- A producers: gather + split + swizzled store ----------------
    // thread (q, rsub) owns the 4 consecutive k values q*4.. of rows rsub, rsub+8, ... ; the loads of
    // the next half k-block are always in flight while the current half is converted and stored.
    const int q = tid & 15;          // float4 slot inside the 64-wide k slice
    const int rsub = tid >> 4;       // 0..7
    struct KDec { int si, dy, dx, nvalid, delta, vec; };
    auto decode = [&](int kb) {
      KDec d;
      const int k = kb * BLOCK_K + q * 4;
      d.si = (p.nseg > 1 && k >= p.seg[1].k_begin) ? 1 : 0;
      const DevSeg& sg = p.seg[d.si];
      const int kk = k - sg.k_begin;
      const int tap = kk / sg.Cin_p;
      const int ci = kk - tap * sg.Cin_p;
      int nvalid = (tap < sg.taps && sg.x != nullptr) ? (sg.Cin - ci) : 0;
      d.nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
      const int ty = tap / sg.kw, tx = tap - ty * sg.kw;
      d.dy = ty * sg.dil_h - sg.pad_h; d.dx = tx * sg.dil_w - sg.pad_w;
      d.delta = (d.dy * p.W + d.dx) * sg.pix_stride + ci;   // tap + channel offset from the row's own pixel
      d.vec = sg.vec;
      return d;
    };
    auto issue = [&](const KDec& d, int half, float4 (&v)[8]) {
      const float* const* rp = bk->rp[d.si];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int row = (half * 8 + jj) * 8 + rsub;
        const int pyx = bk->pyx[row];
        const unsigned yy = (unsigned)((pyx >> 16) + d.dy), xx = (unsigned)((pyx & 0xffff) + d.dx);
        v[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.nvalid > 0 && yy < (unsigned)p.H && xx < (unsigned)p.W)
          v[jj] = ldg_vec4(rp[row] + d.delta, d.nvalid, d.vec);
      }
    };
    // row = 8*(...) + rsub, so the swizzle phase (row & 7) is the thread constant rsub
    const uint32_t st_off = (uint32_t)rsub * 128u + ((uint32_t)((q >> 1) ^ rsub) << 4) + (uint32_t)(q & 1) * 8u;
    auto store = [&](uint8_t* a_stage, int half, const float4 (&v)[8]) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        uint2 pk[NS];
        split4<NS>(v[jj], pk);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          *reinterpret_cast<uint2*>(a_stage + s * A_TILE_BYTES + (half * 8 + jj) * 1024 + st_off) = pk[s];
      }
    };
    float4 va[8], vb[8];
    KDec dcur = decode(0);
    issue(dcur, 0, va);
    for (int kb = 0; kb < p.KB; ++kb) {
      const int stage = kb % S;
      const uint32_t phase = (uint32_t)(kb / S) & 1u;
      issue(dcur, 1, vb);
      mbar_wait(smem_u32(&bk->empty[stage]), phase ^ 1u);
      uint8_t* a_stage = smem + (size_t)stage * p.stage_bytes;
      store(a_stage, 0, va);
      if (kb + 1 < p.KB) {
        dcur = decode(kb + 1);
        issue(dcur, 0, va);
      }
      store(a_stage, 1, vb);
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bk->full[stage]));
    }
  } else if (warp == 4) {
    // ---------------- B producer: one bulk copy per stage ----------------
    if (lane == 0) {
      const uint8_t* src = p.wpk + (size_t)n_tile * p.KB * NS * b_tile_bytes;
      for (int kb = 0; kb < p.KB; ++kb) {
        const int stage = kb % S;
        const uint32_t phase = (uint32_t)(kb / S) & 1u;
        mbar_wait(smem_u32(&bk->empty[stage]), phase ^ 1u);
        const uint32_t bar = smem_u32(&bk->full[stage]);
        mbar_arrive_expect_tx(bar, NS * b_tile_bytes);
        bulk_g2s(base + (uint32_t)stage * p.stage_bytes + NS * A_TILE_BYTES, src + (size_t)kb * NS * b_tile_bytes,
                 NS * b_tile_bytes, bar);
      }
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(BLOCK_M, p.BLOCK_N, 0, 0);
      for (int kb = 0; kb < p.KB; ++kb) {
        const int stage = kb % S;
        const uint32_t phase = (uint32_t)(kb / S) & 1u;
        mbar_wait(smem_u32(&bk->full[stage]), phase);
        tc_fence_after();
        const uint32_t a0 = base + (uint32_t)stage * p.stage_bytes;
        const uint32_t b0 = a0 + NS * A_TILE_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < BLOCK_K / 16; ++k4) {
          // smallest cross terms first, the hi*hi product last
#pragma unroll
          for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
            for (int sa = 0; sa <= sum; ++sa) {
              const int sb = sum - sa;
              const uint64_t ad = smem_desc_sw128(a0 + sa * A_TILE_BYTES + k4 * 32, 16, 1024);
              const uint64_t bd = smem_desc_sw128(b0 + sb * b_tile_bytes + k4 * 32, 16, 1024);
              const uint32_t acc = (kb > 0 || k4 > 0 || sum != NS - 1 || sa > 0) ? 1u : 0u;
              umma_bf16(tmem_d, ad, bd, idesc, acc);
            }
          }
        }
        umma_commit(smem_u32(&bk->empty[stage]));
      }
      umma_commit(smem_u32(&bk->tmem_full));
    }
  }

  // ---------------- epilogue (warps 0-3; warp w owns TMEM lanes / tile rows 32w..32w+31) ----------------
  // all MMAs have completed when tmem_full fires, so the pipeline stages are free: the per-warp
  // staging rows alias them.  Every global access below is a coalesced float4 row segment.
  if (warp < 4) {
    mbar_wait(smem_u32(&bk->tmem_full), 0);
    tc_fence_after();
    const int r0 = warp * 32;
    const uint32_t t_row = tmem_d + ((uint32_t)r0 << 16);

    if (EPI == TC_EPI_CONV) {
      float* stg = reinterpret_cast<float*>(smem) + warp * (32 * CONV_RS);
      for (int c0 = 0; c0 < p.BLOCK_N; c0 += 32) {
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
            if (col_ok && (bk->pyx[row] >> 16) != 0x4000) {
              float* dst = p.y + bk->off_o[O_OUT][row] + col;
              float4 a = *reinterpret_cast<const float4*>(&stg[rr * CONV_RS + c4]);
              a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
              if (p.beta != 0.0f) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                a.x += p.beta * o.x; a.y += p.beta * o.y; a.z += p.beta * o.z; a.w += p.beta * o.w;
              }
              a.x = fov_act(p.act, a.x); a.y = fov_act(p.act, a.y); a.z = fov_act(p.act, a.z); a.w = fov_act(p.act, a.w);
              *reinterpret_cast<float4*>(dst) = a;
            }
          }
        } else {
          const int col = n0 + c0 + lane;
          const bool col_ok = (c0 + lane < p.BLOCK_N) && (col < p.Cout);
          const float bv = (col_ok && p.bias) ? __ldg(&p.bias[col]) : 0.0f;
          for (int rr = 0; rr < 32; ++rr) {
            const int row = r0 + rr;
            if ((bk->pyx[row] >> 16) == 0x4000) break;
            if (col_ok) {
              float* dst = p.y + bk->off_o[O_OUT][row] + col;
              float val = stg[rr * CONV_RS + lane] + bv;
              if (p.beta != 0.0f) val += p.beta * *dst;
              *dst = fov_act(p.act, val);
            }
          }
        }
        __syncwarp();
      }
    } else {
      // Fused ConvLSTM step.  Accumulator columns: [i | f | c~ | o], each F wide.  Channels are
      // processed CW at a time; staging row = [i f g o | c | h] (CW floats each) + 4 pad floats.
      const int F = p.F;
      const int CW = F < 16 ? F : 16;
      const int RS = 6 * CW + 4;
      float* stg = reinterpret_cast<float*>(smem) + warp * (32 * RS);
      const int lpr = CW >> 2;                  // float4 lanes per row segment
      const int rpi = 32 / lpr;                 // rows per warp access
      const int srow = lane / lpr, sc4 = (lane - srow * lpr) * 4;
      // coalesced copy of one CW-wide segment between the staging rows and a global tensor
      auto seg_store = [&](int stage_col, float* dst, int oidx, int dst_col, float* dense) {
        if (!dst) return;
        for (int it = 0; it < 32; it += rpi) {
          const int rr = it + srow, row = r0 + rr;
          if ((bk->pyx[row] >> 16) != 0x4000) {
            const float4 v = *reinterpret_cast<const float4*>(&stg[rr * RS + stage_col + sc4]);
            *reinterpret_cast<float4*>(dst + bk->off_o[oidx][row] + dst_col + sc4) = v;
            if (dense) *reinterpret_cast<float4*>(dense + bk->off_o[O_DENSE][row] + dst_col + sc4) = v;
          }
        }
      };
      for (int cpass = 0; cpass < F; cpass += CW) {
        if (p.c_prev) {
          for (int it = 0; it < 32; it += rpi) {
            const int rr = it + srow, row = r0 + rr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((bk->pyx[row] >> 16) != 0x4000)
              v = __ldg(reinterpret_cast<const float4*>(p.c_prev + bk->off_o[O_CPREV][row] + cpass + sc4));
            *reinterpret_cast<float4*>(&stg[rr * RS + 4 * CW + sc4]) = v;
          }
          __syncwarp();
        }
        // ---- gates from TMEM ----
        float g[4][16];
        if (CW == 16) {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) tmem_ld16(t_row + gi * F + cpass, g[gi]);
          tmem_ld_wait();
        } else {   // F == 8: the four 8-wide gate blocks are the 32 accumulator columns
          float v[32];
          tmem_ld16(t_row, v);
          tmem_ld16(t_row + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int gi = 0; gi < 4; ++gi)
#pragma unroll
            for (int j = 0; j < 8; ++j) g[gi][j] = v[gi * 8 + j];
        }
        float* myrow = stg + lane * RS;
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          if (j4 < CW) {
            float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.c_prev) cp = *reinterpret_cast<const float4*>(&myrow[4 * CW + j4]);
            const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
            float cn[4], hn[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = j4 + e, ch = cpass + j;
              const float ai = fast_rec(p.rec_act, g[0][j] + __ldg(&p.bias[ch]));
              const float af = fast_rec(p.rec_act, g[1][j] + __ldg(&p.bias[F + ch]));
              const float ag = fast_tanh(g[2][j] + __ldg(&p.bias[2 * F + ch]));
              const float ao = fast_rec(p.rec_act, g[3][j] + __ldg(&p.bias[3 * F + ch]));
              cn[e] = af * cpv[e] + ai * ag;
              hn[e] = ao * fast_tanh(cn[e]);
              g[0][j] = ai; g[1][j] = af; g[2][j] = ag; g[3][j] = ao;
            }
#pragma unroll
            for (int gi = 0; gi < 4; ++gi)
              *reinterpret_cast<float4*>(&myrow[gi * CW + j4]) =
                  make_float4(g[gi][j4], g[gi][j4 + 1], g[gi][j4 + 2], g[gi][j4 + 3]);
            *reinterpret_cast<float4*>(&myrow[4 * CW + j4]) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(&myrow[5 * CW + j4]) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          }
        }
        __syncwarp();
        seg_store(4 * CW, p.c_out, O_OUT, cpass, p.cT);
        seg_store(5 * CW, p.h_out, O_H, cpass, p.hT);
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) seg_store(gi * CW, p.gates_out, O_GATES, gi * F + cpass, nullptr);
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// Weight packing: Keras-layout fp32 weights -> bf16 terms in the swizzled smem image of every
// (n-tile, k-block) B tile.  One thread per 16-byte chunk.
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

struct Plan {
  int Cin_p[2], k_begin[2], taps[2];
  int K_total, KB, BLOCK_N, n_tiles, NS, stages, stage_bytes, tmem_cols, data_bytes;
  size_t smem_bytes, ws_bytes;
};

int make_plan(const TcConv& c, Plan* pl) {
  FOV_CHECK_ARG(c.nseg == 1 || c.nseg == 2, "nseg must be 1 or 2");
  FOV_CHECK_ARG(c.math >= 1 && c.math <= 3, "math must be 1..3 bf16 terms");
  FOV_CHECK_ARG(c.N_img > 0 && c.H > 0 && c.W > 0 && c.Cout > 0 && c.T_inner > 0, "bad shape");
  FOV_CHECK_ARG((long long)c.N_img * c.H * c.W < (1LL << 31) - BLOCK_M && c.H < 0x4000 && c.W < 0x10000,
                "too many pixels for 32-bit indexing");
  int k = 0;
  for (int s = 0; s < c.nseg; ++s) {
    const TcSeg& g = c.seg[s];
    FOV_CHECK_ARG(g.Cin > 0 && g.kh > 0 && g.kw > 0 && g.dil_h > 0 && g.dil_w > 0, "bad segment");
    pl->Cin_p[s] = (g.Cin + 7) / 8 * 8;
    pl->taps[s] = g.kh * g.kw;
    pl->k_begin[s] = k;
    k += pl->taps[s] * pl->Cin_p[s];
  }
  pl->K_total = k;
  pl->KB = (k + BLOCK_K - 1) / BLOCK_K;
  pl->NS = c.math;
  const int n_pad = (c.Cout + 15) / 16 * 16;
  if (c.epi == TC_EPI_LSTM) {
    FOV_CHECK_ARG(c.Cout % 4 == 0, "LSTM epilogue needs Cout = 4F");
    const int F = c.Cout / 4;
    FOV_CHECK_ARG(F == 8 || F == 16 || F == 32 || F == 64, "LSTM epilogue supports F in {8,16,32,64}");
    pl->n_tiles = 1;
    pl->BLOCK_N = c.Cout;
  } else {
    const int max_n = pl->NS >= 3 ? 128 : 256;   // keep >= 2 pipeline stages in shared memory
    pl->n_tiles = (n_pad + max_n - 1) / max_n;
    pl->BLOCK_N = ((n_pad + pl->n_tiles - 1) / pl->n_tiles + 15) / 16 * 16;
  }
  pl->tmem_cols = (int)tmem_cols_for(pl->BLOCK_N);
  pl->stage_bytes = pl->NS * (A_TILE_BYTES + pl->BLOCK_N * 128);
  // epilogue staging (aliases the pipeline stages)
  int staging;
  if (c.epi == TC_EPI_LSTM) {
    const int F = c.Cout / 4, CW = F < 16 ? F : 16;
    staging = 4 * 32 * (6 * CW + 4) * 4;
  } else {
    staging = 4 * 32 * CONV_RS * 4;
  }
  // Stage count.  The gather (not the MMA) bounds these kernels, so latency is hidden by co-resident
  // CTAs: aim at 4 CTAs per SM for narrow tiles, 3 up to N=128, 2 beyond (TMEM: CTAs x columns <= 512),
  // and give each CTA as many stages as its share of the shared memory holds.
  const int kUsable = 227 * 1024, book = (int)sizeof(Book) + 1024 + 1024;
  int target = pl->BLOCK_N <= 64 ? 4 : (pl->BLOCK_N <= 128 ? 3 : 2);
  int st = 0;
  for (; target >= 1 && st < 1; --target) {
    int share = kUsable / target - book;
    if (share < staging) continue;
    st = share / pl->stage_bytes;
  }
  if (st > 2 && pl->BLOCK_N <= 64) st = 2;
  if (st > kMaxStages) st = kMaxStages;
  if (st > pl->KB) st = pl->KB;
  if (st < 1) st = 1;
  pl->stages = st;
  pl->data_bytes = st * pl->stage_bytes > staging ? st * pl->stage_bytes : staging;
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
  static size_t configured = 0;
  if (pl.smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<NS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(227 * 1024));
    if (e != cudaSuccess) {
      fov_set_error("tc_conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured = 227 * 1024;
  }
  dim3 grid((unsigned)((dp.M + BLOCK_M - 1) / BLOCK_M), (unsigned)pl.n_tiles);
  tc_conv_kernel<NS, EPI><<<grid, kThreads, pl.smem_bytes, st>>>(dp);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

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
    pp.seg[s] = PackSeg{c.seg[s].w, c.seg[s].Cin, pl.Cin_p[s], pl.taps[s], pl.k_begin[s], c.seg[s].w_mode};
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
  dp.nseg = c.nseg;
  for (int s = 0; s < c.nseg; ++s) {
    const TcSeg& g = c.seg[s];
    // g.x == NULL: the segment reads as zeros (ConvLSTM step 0 with a zero initial state)
    dp.seg[s] = DevSeg{g.x, g.pix_stride, g.Cin, pl.Cin_p[s], g.kw, pl.taps[s], g.dil_h, g.dil_w,
                       g.pad_h, g.pad_w, pl.k_begin[s], pick_vec(g)};
    dp.x_outer[s] = g.img_outer; dp.x_inner[s] = g.img_inner;
  }
  dp.H = c.H; dp.W = c.W; dp.HW = c.H * c.W; dp.T_inner = c.T_inner;
  dp.M = (long long)c.N_img * c.H * c.W;
  dp.KB = pl.KB; dp.Cout = c.Cout; dp.BLOCK_N = pl.BLOCK_N; dp.stages = pl.stages; dp.tmem_cols = pl.tmem_cols;
  dp.stage_bytes = pl.stage_bytes; dp.data_bytes = pl.data_bytes;
  dp.wpk = reinterpret_cast<const uint8_t*>(((uintptr_t)c.ws + 255) & ~(uintptr_t)255);
  dp.bias = c.bias;
  if (c.epi == TC_EPI_CONV) {
    FOV_CHECK_ARG(c.y != nullptr, "NULL output");
    dp.y = c.y; dp.y_outer = c.y_outer; dp.y_inner = c.y_inner; dp.y_pix_stride = c.y_pix_stride;
    dp.act = c.act; dp.beta = c.beta;
    dp.vec_out = ((uintptr_t)c.y % 16 == 0) && (c.y_pix_stride % 4 == 0) && (c.y_outer % 4 == 0) &&
                 (c.y_inner % 4 == 0) && (c.Cout % 4 == 0) && (!c.bias || (uintptr_t)c.bias % 16 == 0);
  } else {
    FOV_CHECK_ARG(c.bias && c.c_out && c.h_out, "NULL LSTM epilogue pointer");
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
  return FOV_OK;
}
}  // namespace

extern "C" size_t fov_conv_tc_ws_bytes(const fov_conv_cfg* cfg, int math, int bwd_data) {
  TcConv c;
  if (conv_from_cfg(cfg, math, &c)) return 0;
  TcSeg& g = c.seg[0];
  g.kh = cfg->kh; g.kw = cfg->kw; g.dil_h = cfg->dil_h; g.dil_w = cfg->dil_w;
  g.Cin = bwd_data ? cfg->Cout : cfg->Cin;
  c.Cout = bwd_data ? cfg->Cin : cfg->Cout;
  return tc_conv_ws_bytes(c);
}

extern "C" int fov_conv2d_fwd_tc(const fov_conv_cfg* cfg, const float* x, const float* w, const float* bias,
                                 float* y, void* ws, int math, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(x && w && y && ws, "NULL pointer");
  TcSeg& g = c.seg[0];
  g.x = x; g.img_outer = cfg->x_img_stride; g.img_inner = 0; g.pix_stride = cfg->x_pix_stride;
  g.Cin = cfg->Cin; g.kh = cfg->kh; g.kw = cfg->kw; g.dil_h = cfg->dil_h; g.dil_w = cfg->dil_w;
  g.pad_h = cfg->pad_h; g.pad_w = cfg->pad_w; g.w = w; g.w_mode = 0;
  c.Cout = cfg->Cout; c.ws = ws;
  c.bias = bias; c.y = y; c.y_outer = cfg->y_img_stride; c.y_inner = 0; c.y_pix_stride = cfg->y_pix_stride;
  c.act = cfg->act; c.beta = cfg->beta;
  return tc_conv_run(c, (cudaStream_t)stream);
}

extern "C" int fov_conv2d_bwd_data_tc(const fov_conv_cfg* cfg, const float* dy, const float* w, float* dx,
                                      void* ws, int math, void* stream) {
  TcConv c;
  int rc = conv_from_cfg(cfg, math, &c);
  if (rc) return rc;
  FOV_CHECK_ARG(dy && w && dx && ws, "NULL pointer");
  TcSeg& g = c.seg[0];
  g.x = dy; g.img_outer = cfg->y_img_stride; g.img_inner = 0; g.pix_stride = cfg->y_pix_stride;
  g.Cin = cfg->Cout; g.kh = cfg->kh; g.kw = cfg->kw; g.dil_h = cfg->dil_h; g.dil_w = cfg->dil_w;
  g.pad_h = (cfg->kh - 1) * cfg->dil_h - cfg->pad_h; g.pad_w = (cfg->kw - 1) * cfg->dil_w - cfg->pad_w;
  g.w = w; g.w_mode = 1;
  c.Cout = cfg->Cin; c.ws = ws;
  c.bias = nullptr; c.y = dx; c.y_outer = cfg->x_img_stride; c.y_inner = 0; c.y_pix_stride = cfg->x_pix_stride;
  c.act = FOV_ACT_LINEAR; c.beta = cfg->beta;
  return tc_conv_run(c, (cudaStream_t)stream);
}
