// Persistent fc-LSTM encoder-decoder forward on tensor cores (tcgen05 + TMEM): ONE launch runs the T_enc encoder
// steps and the T_dec decoder steps (teacher forced or autoregressive) of 128 sequences per image group.
//
//   * a group = 128 sequences = the 128 rows (TMEM lanes) of one accumulator tile; a CTA runs NG groups with their own
//     TMEM columns and barriers so the gate algebra of one group overlaps the MMAs of the other;
//   * the gate weights [W ; U] of the running phase are converted ONCE per phase from their Keras layout
//     (in,4H) / (H,4H) into bf16 terms in K-major 128-byte-swizzled shared-memory tiles (B operand, N = 4H = 256);
//   * per step and group one tcgen05.mma chain  D[128 x 256] = [x_t | h_{t-1}] x [W ; U]  (K = 16 + 64, every
//     significant product of the bf16 terms), issued by one elected thread of the group once its 128 rows are staged
//     (no dedicated MMA warp: 8 warps per CTA keep the full 255-register budget for the 64-float cell state);
//   * the epilogue thread of a row reads its 256 gate pre-activations with tcgen05.ld, keeps the cell state of its
//     sequence (64 floats) in registers, writes h_t as bf16 terms straight into the swizzled A-operand rows of the next
//     step (h never goes through HBM), runs the Dense head (+ optional additive term) on the h it just produced and,
//     when decoding autoregressively, writes the head output as the next step's x operand: no per-step launch and no
//     host round trip (mycode/FoV_seq2seq.py:154-178 does 11 predict() calls per sample here);
//   * teacher-forced inputs x_{t+1} are prefetched into registers while the MMAs of step t run.
// In training mode the saved tensors the BPTT kernel needs (lstm_seq2seq.cu) are written by the row's thread.
//
// Replaces the Keras LSTM/Dense calls cited in include/fov360.h when cfg.math != 0.
#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kH = 64;                     // latent_dim
constexpr int kG = 256;                    // 4H = MMA N
constexpr int kRows = 128;                 // sequences per group = MMA M
constexpr int kXK = 16;                    // K extent of the x operand (in_dim <= 16)
constexpr uint32_t kBTerm = kG * 128;      // bytes of one bf16 term of a weight tile (256 rows x 128 B)
constexpr uint32_t kATerm = kRows * 128;   // bytes of one bf16 term of an operand tile (128 rows x 128 B)

struct TcPhase {
  const float *Wk, *Uk, *bk;               // Keras layout
  const float* x;                          // (B, x_T, in_dim)
  int in_dim, T, x_T;
  int ar, has_head, zero_init;
  const float* extra;                      // optional (B,T,out)
  float* y;                                // (B,T,out)
  fov_lstm_saved sv;
  // time-batched input projection x_t . W of every step of the phase (xproj_tc.cu), tiled [b/128][t][col/4][b%128][4];
  // when set, the per-step GEMM is the K = 64 recurrent part only and the epilogue adds the projection (any in_dim)
  const float* xproj;
};

struct LstmTcParams {
  TcPhase ph[2];
  int nph;
  int B, out_dim, head_act, training, dbg;
  const float *Wo, *bo;
  const float *h0, *c0;
  float *hT, *cT;
};

template <int OD>
struct LstmTcBook {
  float4 bias4[kH];                        // (b_i, b_f, b_c, b_o) of a hidden unit
  float Wo_s[kH * OD];                     // head kernel, rows padded to OD columns
  float bo_s[OD];
  float ypart[2 * kRows * OD];             // WPG = 8: head partial sums of the threads that own units 32..63
  uint64_t tmem_full[2];
  uint32_t tmem_ptr;
};

// diagnostics (fov_debug_lstm_tc_read): cycles CTA 0 / thread 0 spent [0] waiting for the accumulator, [1] in the gate
// algebra, [2] head + stores + operand writes, [3] whole step loop, [5] issuing the MMAs
__device__ unsigned long long g_lstm_tc_timeline[8];

__device__ __forceinline__ void bar_workers(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }
__device__ __forceinline__ void bar_group(int g, int n) { asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "r"(n) : "memory"); }

// tanh(x) = 1 - 2 / (1 + 2^(x * 2 log2 e)): FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA (abs error ~1e-7)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh5(float x) {
  return fmaf(-2.0f, rcp_approx(ex2_approx(x * 2.885390082f) + 1.0f), 1.0f);
}
__device__ __forceinline__ float head_act_fn(int act, float x) {
  if (act == FOV_ACT_TANH) return tanh5(x);
  if (act == FOV_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

// 16 floats -> global row segment with the widest store the row alignment allows (vec = 4, 2 or 1 floats)
__device__ __forceinline__ void store16(float* dst, const float (&v)[16], int vec) {
  if (vec == 4) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else if (vec == 2) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[j], v[j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j] = v[j];
  }
}

// one thread = one row here, so a warp's store touches 32 different rows: a 256-bit store (sm_100: STG.E.256) fills a whole
// 32-byte sector per lane where two 128-bit stores write two half sectors
__device__ __forceinline__ void st_global_v8(float* dst, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void store8(float* dst, const float (&v)[8], int vec) {
  if (vec == 4) {
    if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
      st_global_v8(dst, v);
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else if (vec == 2) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[j], v[j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = v[j];
  }
}
// Orders every later use of v[] after the preceding tcgen05.wait::ld (volatile asm statements keep their order; the
// loads that fill v[] were issued a whole pass earlier).
__device__ __forceinline__ void pin8(float (&v)[8]) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}

// WPG = warps per 128-sequence group: 4 (a thread owns the 64 hidden units of its sequence) or 8 (the two warps of a
// TMEM lane quarter own 32 units each: twice the warps per scheduler to hide the MUFU / tcgen05.ld latencies, half the
// cell state per thread; the head partial sums of the upper half travel through shared memory)
template <int NS, int NG, int REC, int OD, bool TRAIN, int WPG, bool XP>
__global__ void __launch_bounds__(32 * WPG * NG, 1) lstm_tc_fwd_kernel(const __grid_constant__ LstmTcParams P) {
  constexpr int NW = 32 * WPG * NG;        // threads
  constexpr int GT = 32 * WPG;             // threads per group
  constexpr int UPT = kH / (WPG / 4);      // hidden units per thread
  constexpr int NPASS = UPT / 8;
  constexpr bool PF = WPG == 4;            // double-buffered accumulator reads (register budget allows it at WPG = 4)
  constexpr uint32_t kWbOff = NS * kBTerm;                 // x-part weights (all terms in one tile, 32-byte chunks)
  constexpr uint32_t kActOff = kWbOff + kBTerm;
  constexpr uint32_t kGrpBytes = (NS + 1) * kATerm;        // h terms, then the x tile (terms as 32-byte chunks)
  using Book = LstmTcBook<OD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  Book* bk = reinterpret_cast<Book*>(smem + kActOff + NG * kGrpBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---------------- setup ----------------
  if (warp == 0) {
    if (lane == 0) {
      for (int g = 0; g < NG; ++g) mbar_init(smem_u32(&bk->tmem_full[g]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bk->tmem_ptr), NG * kG);
    tmem_relinquish();
  }
  for (int idx = tid; idx < kH * OD; idx += blockDim.x) {
    const int u = idx / OD, d = idx - u * OD;
    bk->Wo_s[idx] = (P.out_dim > 0 && d < P.out_dim) ? __ldg(&P.Wo[u * P.out_dim + d]) : 0.0f;
  }
  if (tid < OD) bk->bo_s[tid] = (P.out_dim > 0 && tid < P.out_dim) ? __ldg(&P.bo[tid]) : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  {
    // ---------------- thread = one sequence (accumulator row) ----------------
    const int g = warp / WPG, wg = warp % WPG, q = wg & 3, half = wg >> 2, r = q * 32 + lane;
    const int u0 = half * UPT;               // my first hidden unit
    const bool lead = half == 0;             // the thread of the row that handles x, the head tail and y
    const long long b = ((long long)blockIdx.x * NG + g) * kRows + r;
    const bool valid = b < P.B;
    uint8_t* hA = smem + kActOff + (uint32_t)g * kGrpBytes;
    uint8_t* xA = hA + NS * kATerm;
    const uint32_t rowoff = (uint32_t)r * 128u, rsw = (uint32_t)(r & 7);
    const uint32_t t_row = tmem_d + (uint32_t)(g * kG) + ((uint32_t)(q * 32) << 16);
    const bool dbg = P.dbg && blockIdx.x == 0 && tid == 0;
    const bool issuer = r == 0 && lead;
    long long tw = 0, ta = 0, tb = 0, tm = 0;
    const long long t_begin = clock64();
    const uint32_t idesc = idesc_bf16_f32(kRows, kG, 0, 0);
    // D[128 x 256] = [x_t | h_{t-1}] x [W ; U] for my group: called by ONE thread after the group barrier
    auto issue_mmas = [&](bool use_x) {
      tc_fence_after();
      const uint32_t hA_s = base + kActOff + (uint32_t)g * kGrpBytes, xA_s = hA_s + NS * kATerm;
      const uint32_t d = tmem_d + (uint32_t)(g * kG);
      uint32_t acc = 0;
      // x part: one k16 step, the terms are 32-byte chunks of the same tile (skipped when the projection is time-batched)
      if (use_x) {
#pragma unroll
        for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
          for (int sa = 0; sa <= sum; ++sa) {
            const int sb = sum - sa;
            umma_bf16(d, desc_at(kDescHi128, xA_s + sa * 32), desc_at(kDescHi128, base + kWbOff + sb * 32), idesc, acc);
            acc = 1;
          }
        }
      }
      // h part: four k16 steps
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
        for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
          for (int sa = 0; sa <= sum; ++sa) {
            const int sb = sum - sa;
            umma_bf16(d, desc_at(kDescHi128, hA_s + sa * kATerm + k4 * 32),
                      desc_at(kDescHi128, base + sb * kBTerm + k4 * 32), idesc, acc);
            acc = 1;
          }
        }
      }
      umma_commit(smem_u32(&bk->tmem_full[g]));
    };

    // 8 consecutive hidden units of my row -> bf16 terms, chunk c8 of the h operand tile
    auto h_store8 = [&](int c8, const float* hn) {
      uint2 lo[NS], hi[NS];
      split4<NS>(make_float4(hn[0], hn[1], hn[2], hn[3]), lo);
      split4<NS>(make_float4(hn[4], hn[5], hn[6], hn[7]), hi);
      const uint32_t so = rowoff + ((((uint32_t)c8) ^ rsw) << 4);
#pragma unroll
      for (int s = 0; s < NS; ++s)
        *reinterpret_cast<uint4*>(hA + s * kATerm + so) = make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
    };
    // the 16 input features of my row -> term s in the 32-byte chunk pair (2s, 2s+1) of the x tile
    auto x_store = [&](const float (&xn)[kXK]) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint2 lo[NS], hi[NS];
        split4<NS>(make_float4(xn[hf * 8 + 0], xn[hf * 8 + 1], xn[hf * 8 + 2], xn[hf * 8 + 3]), lo);
        split4<NS>(make_float4(xn[hf * 8 + 4], xn[hf * 8 + 5], xn[hf * 8 + 6], xn[hf * 8 + 7]), hi);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          *reinterpret_cast<uint4*>(xA + rowoff + ((((uint32_t)(2 * s + hf)) ^ rsw) << 4)) =
              make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
      }
    };

    float c[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) c[u] = 0.0f;

    int step = 0;
    for (int pi = 0; pi < P.nph; ++pi) {
      const TcPhase& ph = P.ph[pi];
      const bool xp = XP && ph.xproj != nullptr;          // input projection precomputed for every step of this phase
      const int in_dim = ph.in_dim, T = ph.T, K = fov_lstm_xh_stride(kH, in_dim);   // padded [h | x | 0] row width
      const bool save = TRAIN;                            // inference instantiations carry no saved-tensor stores
      const int xh_vec = 4;
      const int Kn = (pi + 1 < P.nph) ? fov_lstm_xh_stride(kH, P.ph[pi + 1].in_dim) : K;   // of the next phase
      const int xhn_vec = 4;
      // every MMA of the previous phase has completed once all workers got here (each group waited for its last commit)
      tc_fence_before();
      bar_workers(NW);
      // ---- gate weights of this phase -> bf16 terms, K-major swizzled rows (row n = gate column n) ----
      for (int n = tid; n < kG; n += NW) {
        const uint32_t nsw = (uint32_t)(n & 7), noff = (uint32_t)n * 128u;
#pragma unroll 2
        for (int c8 = 0; c8 < 8; ++c8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __ldg(&ph.Uk[(c8 * 8 + j) * kG + n]);
          uint2 lo[NS], hi[NS];
          split4<NS>(make_float4(v[0], v[1], v[2], v[3]), lo);
          split4<NS>(make_float4(v[4], v[5], v[6], v[7]), hi);
#pragma unroll
          for (int s = 0; s < NS; ++s)
            *reinterpret_cast<uint4*>(smem + s * kBTerm + noff + ((((uint32_t)c8) ^ nsw) << 4)) =
                make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (xp) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = hf * 8 + j;
            v[j] = k < in_dim ? __ldg(&ph.Wk[k * kG + n]) : 0.0f;
          }
          uint2 lo[NS], hi[NS];
          split4<NS>(make_float4(v[0], v[1], v[2], v[3]), lo);
          split4<NS>(make_float4(v[4], v[5], v[6], v[7]), hi);
#pragma unroll
          for (int s = 0; s < NS; ++s)
            *reinterpret_cast<uint4*>(smem + kWbOff + noff + ((((uint32_t)(2 * s + hf)) ^ nsw) << 4)) =
                make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
        }
      }
      // biases with the affine part of the recurrent activation folded in: hard_sigmoid(z + b) = sat(0.2 z + (0.2 b + 0.5)),
      // sigmoid(z + b) = 1 / (1 + 2^(-log2e z - log2e b))
      if (tid < kH) {
        const float bi = __ldg(&ph.bk[tid]), bf = __ldg(&ph.bk[kH + tid]), bc = __ldg(&ph.bk[2 * kH + tid]),
                    bo = __ldg(&ph.bk[3 * kH + tid]);
        if (REC == FOV_REC_HARD_SIGMOID)
          bk->bias4[tid] = make_float4(fmaf(0.2f, bi, 0.5f), fmaf(0.2f, bf, 0.5f), bc, fmaf(0.2f, bo, 0.5f));
        else
          bk->bias4[tid] = make_float4(-1.442695041f * bi, -1.442695041f * bf, bc, -1.442695041f * bo);
      }
      // ---- initial state of this phase ----
      if (pi == 0 || ph.zero_init) {
        const float* h0 = (pi == 0 && !ph.zero_init && valid) ? P.h0 : nullptr;
        const float* c0 = (pi == 0 && !ph.zero_init && valid) ? P.c0 : nullptr;
        float* xh0 = (save && valid && ph.sv.xh) ? ph.sv.xh + (size_t)b * T * K : nullptr;
#pragma unroll
        for (int p16 = 0; p16 < UPT / 16; ++p16) {
          float hv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            hv[j] = h0 ? __ldg(&h0[(size_t)b * kH + u0 + p16 * 16 + j]) : 0.0f;
            c[p16 * 16 + j] = c0 ? __ldg(&c0[(size_t)b * kH + u0 + p16 * 16 + j]) : 0.0f;
          }
          h_store8((u0 >> 3) + p16 * 2, hv);
          h_store8((u0 >> 3) + p16 * 2 + 1, hv + 8);
          if (xh0) store16(xh0 + u0 + p16 * 16, hv, xh_vec);
        }
      }
      // ---- x_0 ----
      if (lead && !xp) {
        float xn[kXK];
#pragma unroll
        for (int k = 0; k < kXK; ++k)
          xn[k] = (valid && k < in_dim) ? __ldg(&ph.x[(size_t)b * ph.x_T * in_dim + k]) : 0.0f;
        x_store(xn);
        if (save && valid && ph.sv.xh) {
          float* xr = ph.sv.xh + (size_t)b * T * K + kH;
#pragma unroll
          for (int k = 0; k < kXK + 3; ++k)
            if (kH + k < K) xr[k] = k < kXK ? xn[k < kXK ? k : 0] : 0.0f;     // xn is zero beyond in_dim: the row pad
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      bar_workers(NW);
      if (issuer) issue_mmas(!xp);
      // my row of the tiled projection workspace: column group cg of step t at Pr + (t * 64 + cg) * 512 floats
      const float* Pr = xp ? ph.xproj + (((size_t)(b >> 7) * T) * 64 * 128 + (size_t)(b & 127)) * 4 : nullptr;

      const bool next_phase_carries = (pi + 1 < P.nph) && !P.ph[pi + 1].zero_init;
      for (int t = 0; t < T; ++t) {
        const bool more = t + 1 < T;
        float xn[kXK];
#pragma unroll
        for (int k = 0; k < kXK; ++k) xn[k] = 0.0f;
        if (!ph.ar && more && lead && !xp) {
#pragma unroll
          for (int k = 0; k < kXK; ++k)
            xn[k] = (valid && k < in_dim) ? __ldg(&ph.x[((size_t)b * ph.x_T + t + 1) * in_dim + k]) : 0.0f;
        }
        // projection of this step, first pass: in flight while waiting for the accumulator
        float4 pq[XP ? 2 : 1][8];
        auto p_load = [&](int p8, float4 (&dst)[8]) {
          const float* src = Pr + ((size_t)t * 64 + (u0 >> 2) + p8 * 2) * 512;
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            dst[gi * 2] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gi * 16 * 512));
            dst[gi * 2 + 1] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(gi * 16 + 1) * 512));
          }
        };
        if (XP && xp) p_load(0, pq[0]);
        const size_t rowt = (size_t)b * T + t;
        float* gates_p = (save && valid && ph.sv.gates) ? ph.sv.gates + rowt * kG : nullptr;
        float* c_p = (save && valid && ph.sv.c) ? ph.sv.c + rowt * kH : nullptr;
        float* hseq_p = (valid && ph.sv.hseq) ? ph.sv.hseq + rowt * kH : nullptr;
        float* xh_next = nullptr;
        if (save && valid) {
          if (more) xh_next = ph.sv.xh ? ph.sv.xh + (rowt + 1) * K : nullptr;
          else if (next_phase_carries && P.ph[pi + 1].sv.xh)
            xh_next = P.ph[pi + 1].sv.xh + (size_t)b * P.ph[pi + 1].T * Kn;
        }
        const int xh_next_vec = more ? xh_vec : xhn_vec;
        float* hT_p = (!more && pi + 1 == P.nph && valid && P.hT) ? P.hT + (size_t)b * kH : nullptr;

        long long k0 = 0, k1 = 0, k2 = 0;
        if (dbg) k0 = clock64();
        mbar_wait(smem_u32(&bk->tmem_full[g]), (uint32_t)step & 1u);
        __syncwarp();
        tc_fence_after();
        if (dbg) { k1 = clock64(); tw += k1 - k0; }

        float yacc[OD];
#pragma unroll
        for (int d = 0; d < OD; ++d) yacc[d] = 0.0f;
        // 8 hidden units per pass; at WPG = 4 the accumulator columns of pass p+1 are in flight (tcgen05.ld) during the
        // math of pass p
        float gt[PF ? 2 : 1][4][8];
        const uint32_t t_col = t_row + (uint32_t)u0;
        if (PF) {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) tmem_ld8(t_col + gi * kH, gt[0][gi]);
          tmem_ld_wait();
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) pin8(gt[0][gi]);
        }
#pragma unroll
        for (int p8 = 0; p8 < NPASS; ++p8) {
          float (&ga)[4][8] = gt[PF ? (p8 & 1) : 0];
          if (PF) {
            if (p8 + 1 < NPASS) {
#pragma unroll
              for (int gi = 0; gi < 4; ++gi) tmem_ld8(t_col + gi * kH + (p8 + 1) * 8, gt[PF ? ((p8 + 1) & 1) : 0][gi]);
            }
          } else {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) tmem_ld8(t_col + gi * kH + p8 * 8, ga[gi]);
            tmem_ld_wait();
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) pin8(ga[gi]);
          }
          if (XP && xp) {
            if (p8 + 1 < NPASS) p_load(p8 + 1, pq[XP ? ((p8 + 1) & 1) : 0]);
            const float4 (&pp)[8] = pq[XP ? (p8 & 1) : 0];
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              ga[gi][0] += pp[gi * 2].x; ga[gi][1] += pp[gi * 2].y; ga[gi][2] += pp[gi * 2].z; ga[gi][3] += pp[gi * 2].w;
              ga[gi][4] += pp[gi * 2 + 1].x; ga[gi][5] += pp[gi * 2 + 1].y; ga[gi][6] += pp[gi * 2 + 1].z;
              ga[gi][7] += pp[gi * 2 + 1].w;
            }
          }
          float hn[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int u = p8 * 8 + j;
            const float4 bb = bk->bias4[u0 + u];
            float ai, af, ao;
            if (REC == FOV_REC_HARD_SIGMOID) {
              ai = __saturatef(fmaf(0.2f, ga[0][j], bb.x));
              af = __saturatef(fmaf(0.2f, ga[1][j], bb.y));
              ao = __saturatef(fmaf(0.2f, ga[3][j], bb.w));
            } else {
              ai = rcp_approx(1.0f + ex2_approx(fmaf(-1.442695041f, ga[0][j], bb.x)));
              af = rcp_approx(1.0f + ex2_approx(fmaf(-1.442695041f, ga[1][j], bb.y)));
              ao = rcp_approx(1.0f + ex2_approx(fmaf(-1.442695041f, ga[3][j], bb.w)));
            }
            const float ag = tanh5(ga[2][j] + bb.z);
            const float cn = fmaf(af, c[u], ai * ag);
            c[u] = cn;
            hn[j] = ao * tanh5(cn);
            ga[0][j] = ai; ga[1][j] = af; ga[2][j] = ag; ga[3][j] = ao;
          }
          h_store8((u0 >> 3) + p8, hn);
          if (ph.has_head) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4* wr = reinterpret_cast<const float4*>(&bk->Wo_s[(u0 + p8 * 8 + j) * OD]);
#pragma unroll
              for (int d4 = 0; d4 < OD / 4; ++d4) {
                const float4 w4 = wr[d4];
                yacc[d4 * 4 + 0] = fmaf(hn[j], w4.x, yacc[d4 * 4 + 0]);
                yacc[d4 * 4 + 1] = fmaf(hn[j], w4.y, yacc[d4 * 4 + 1]);
                yacc[d4 * 4 + 2] = fmaf(hn[j], w4.z, yacc[d4 * 4 + 2]);
                yacc[d4 * 4 + 3] = fmaf(hn[j], w4.w, yacc[d4 * 4 + 3]);
              }
            }
          }
          if (gates_p) {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) store8(gates_p + gi * kH + u0 + p8 * 8, ga[gi], 4);
          }
          if (c_p) {
            float cv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cv[j] = c[p8 * 8 + j];
            store8(c_p + u0 + p8 * 8, cv, 4);
          }
          if (hseq_p) store8(hseq_p + u0 + p8 * 8, hn, 4);
          if (xh_next) store8(xh_next + u0 + p8 * 8, hn, xh_next_vec);
          if (hT_p) store8(hT_p + u0 + p8 * 8, hn, 4);
          if (PF && p8 + 1 < NPASS) {
            tmem_ld_wait();
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) pin8(gt[PF ? ((p8 + 1) & 1) : 0][gi]);
          }
        }
        tc_fence_before();                 // my tcgen05.ld of this accumulator are complete before the next MMAs
        if (dbg) { k2 = clock64(); ta += k2 - k1; }

        if (ph.has_head) {
          if (WPG == 8) {
            // the thread that owns units 32..63 of the row hands its partial head sums to the lead thread
            float* yp = &bk->ypart[(g * kRows + r) * OD];
            if (!lead) {
#pragma unroll
              for (int d4 = 0; d4 < OD / 4; ++d4)
                reinterpret_cast<float4*>(yp)[d4] = make_float4(yacc[d4 * 4], yacc[d4 * 4 + 1], yacc[d4 * 4 + 2], yacc[d4 * 4 + 3]);
            }
            bar_group(g, GT);
            if (lead) {
#pragma unroll
              for (int d4 = 0; d4 < OD / 4; ++d4) {
                const float4 v = reinterpret_cast<const float4*>(yp)[d4];
                yacc[d4 * 4] += v.x; yacc[d4 * 4 + 1] += v.y; yacc[d4 * 4 + 2] += v.z; yacc[d4 * 4 + 3] += v.w;
              }
            }
          }
          if (lead) {
            const size_t o = rowt * (size_t)P.out_dim;
#pragma unroll
            for (int d = 0; d < OD; ++d) {
              float sum = yacc[d] + bk->bo_s[d];
              const bool live = d < P.out_dim;
              if (live && valid && ph.extra) sum += __ldg(&ph.extra[o + d]);
              const float yv = head_act_fn(P.head_act, sum);
              if (live && valid) ph.y[o + d] = yv;
              if (ph.ar && d < kXK) xn[d] = live ? yv : 0.0f;
            }
          }
        }
        if (more) {
          if (lead && !xp) {
            x_store(xn);
            if (save && valid && ph.sv.xh) {
              float* xr = ph.sv.xh + (rowt + 1) * K + kH;
#pragma unroll
              for (int k = 0; k < kXK + 3; ++k)
                if (kH + k < K) xr[k] = k < kXK ? xn[k < kXK ? k : 0] : 0.0f;
            }
          }
          fence_proxy_async_smem();
          bar_group(g, GT);
          long long k3 = 0;
          if (dbg) k3 = clock64();
          if (issuer) issue_mmas(!xp);
          if (dbg) tm += clock64() - k3;
        }
        ++step;
        if (dbg) tb += clock64() - k2;
      }
    }
    if (valid && P.cT) {
#pragma unroll
      for (int p8 = 0; p8 < NPASS; ++p8) {
        float cv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cv[j] = c[p8 * 8 + j];
        store8(P.cT + (size_t)b * kH + u0 + p8 * 8, cv, 4);
      }
    }
    if (dbg) {
      g_lstm_tc_timeline[0] = tw; g_lstm_tc_timeline[1] = ta; g_lstm_tc_timeline[2] = tb;
      g_lstm_tc_timeline[3] = clock64() - t_begin;
      g_lstm_tc_timeline[5] = tm;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, NG * kG);
}

template <int NS, int NG, int OD>
size_t tc_smem_bytes() {
  return (size_t)(NS + 1) * kBTerm + (size_t)NG * (NS + 1) * kATerm + sizeof(LstmTcBook<OD>) + 1024;
}

template <int NS, int NG, int REC, int OD, bool TRAIN, int WPG, bool XP = false>
int launch_tc_w(const LstmTcParams& P, cudaStream_t st) {
  const size_t smem = tc_smem_bytes<NS, NG, OD>();
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_fwd_kernel<NS, NG, REC, OD, TRAIN, WPG, XP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      fov_set_error("fov_lstm (tensor-core): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  const int per_cta = kRows * NG;
  const int grid = (P.B + per_cta - 1) / per_cta;
  lstm_tc_fwd_kernel<NS, NG, REC, OD, TRAIN, WPG, XP><<<grid, 32 * WPG * NG, smem, st>>>(P);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

int g_lstm_tc_wpg = 8;                       // diagnostics: 4 = one thread per sequence everywhere

template <int NS, int NG, int REC, int OD, bool TRAIN>
int launch_tc_t(const LstmTcParams& P, cudaStream_t st) {
  // a phase with a time-batched input projection: ONE group of 8 warps per CTA (32 hidden units per thread leave the
  // registers for the projection prefetch; without the x operand / x weights the CTA needs ~100 KB of shared memory)
  if (NG == 1 && (P.ph[0].xproj || (P.nph > 1 && P.ph[1].xproj)))
    return launch_tc_w<NS, NG, REC, OD, TRAIN, (NG == 1 ? 8 : 4), (NG == 1)>(P, st);
  // 8 warps per group when a CTA runs ONE group (batches below 128 x 148 sequences, default two-term arithmetic): the
  // step is then a serial MMA -> epilogue chain and the extra warps shorten it (B=4096 AR decode: 0.091 vs 0.112 ms).
  // With two groups per CTA 16 warps get 128 registers each, spill, and lose 5 % to the 4-warp form (measured).
  if (NS == 2 && NG == 1 && g_lstm_tc_wpg == 8)
    return launch_tc_w<NS, NG, REC, OD, TRAIN, (NS == 2 && NG == 1 ? 8 : 4)>(P, st);
  return launch_tc_w<NS, NG, REC, OD, TRAIN, 4>(P, st);
}

template <int NS, int NG, int REC, int OD>
int launch_tc(const LstmTcParams& P, cudaStream_t st) {
  return P.training ? launch_tc_t<NS, NG, REC, OD, true>(P, st) : launch_tc_t<NS, NG, REC, OD, false>(P, st);
}

template <int NS, int NG>
int launch_tc_ro(const LstmTcParams& P, int rec, cudaStream_t st) {
  const bool hs = rec == FOV_REC_HARD_SIGMOID;
  if (P.out_dim <= 8)
    return hs ? launch_tc<NS, NG, FOV_REC_HARD_SIGMOID, 8>(P, st) : launch_tc<NS, NG, FOV_REC_SIGMOID, 8>(P, st);
  return hs ? launch_tc<NS, NG, FOV_REC_HARD_SIGMOID, 16>(P, st) : launch_tc<NS, NG, FOV_REC_SIGMOID, 16>(P, st);
}

int g_lstm_tc_dbg = 0;

}  // namespace

extern "C" void fov_debug_lstm_tc_enable(int on) { g_lstm_tc_dbg = on; }
extern "C" void fov_debug_lstm_tc_wpg(int wpg) { g_lstm_tc_wpg = wpg; }
extern "C" int fov_debug_lstm_tc_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_lstm_tc_timeline, sizeof(unsigned long long) * 8);
}

// =====================================================================================================================
// Persistent fc-LSTM BPTT on tensor cores: ONE launch walks the decoder steps T_dec-1..0 and then the encoder steps.
//   * thread = one sequence (accumulator row); dc of its 64 hidden units lives in registers for the whole sequence;
//   * per step the thread turns the saved gates / cell states of its row into the gate pre-activation gradients dZ_t
//     (written to HBM for the time-batched weight gradients, and as bf16 terms into the swizzled A-operand rows);
//   * one tcgen05.mma chain  [dh_rec_{t-1} | dx_t] = dZ_t x [U | W]^T  (K = 4H = 256; the Keras layouts (H,4H) / (in,4H)
//     ARE the K-major B operand, rows = N); dh_rec is read back from TMEM by the next (earlier) step, dx_t - only needed
//     when decoding autoregressively - re-enters through the Dense head chain dy_{t-1} += dx_t.
// Same contract as the fp32 kernel of lstm_seq2seq.cu (saved tensors in, dz / dpre out).
// =====================================================================================================================
namespace {

struct BwdTcPhase {
  const float *Uk, *Wk;                    // recurrent (H,4H); kernel (in,4H), read in autoregressive mode only
  int in_dim, T;
  int ar, has_head, zero_carry;            // zero_carry: the phase starts from dh = dc = 0
  const float *dy, *y, *dhseq;             // (B,T,out) x2; optional (B,T,H)
  float *dpre, *dz;                        // (B,T,out); (B,T,4H)
  const float *gates, *c;                  // saved activated gates (B,T,4H), cell states (B,T,H)
  const float* c_init;                     // c_{-1}: base pointer + per-sequence stride, or NULL (zeros)
  long long c_init_stride;
};
struct LstmTcBwdParams {
  BwdTcPhase ph[2];
  int nph, B, out_dim, head_act;
  const float* Wo;
};
template <int OD>
struct LstmTcBwdBook {
  float Wo_s[kH * OD];
  uint64_t tmem_full;
  uint32_t tmem_ptr;
};

template <int REC>
__device__ __forceinline__ float rec_grad(float a) {
  if (REC == FOV_REC_HARD_SIGMOID) return (a > 0.0f && a < 1.0f) ? 0.2f : 0.0f;
  return a * (1.0f - a);
}
// 8 consecutive floats of the thread's own row: one 256-bit load (LDG.E.256, a full 32-byte sector) when aligned
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  if ((reinterpret_cast<uintptr_t>(p) & 31u) == 0) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

template <int NS, int REC, int OD, bool AR>
__global__ void __launch_bounds__(kRows, 1) lstm_tc_bwd_kernel(const __grid_constant__ LstmTcBwdParams P) {
  constexpr int NB = AR ? kH + kXK : kH;                 // B rows = accumulator columns: dh_rec (+ dx)
  constexpr uint32_t kBTile = NB * 128;                  // one 64-wide k tile of one term of [U | W]^T
  constexpr uint32_t kAOff = NS * 4 * kBTile;
  using Book = LstmTcBwdBook<OD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  Book* bk = reinterpret_cast<Book*>(smem + kAOff + NS * 4 * kATerm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&bk->tmem_full), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bk->tmem_ptr), 128);
    tmem_relinquish();
  }
  for (int idx = tid; idx < kH * OD; idx += kRows) {
    const int u = idx / OD, d = idx - u * OD;
    bk->Wo_s[idx] = (P.out_dim > 0 && d < P.out_dim) ? __ldg(&P.Wo[u * P.out_dim + d]) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  const int r = tid;
  const long long b = (long long)blockIdx.x * kRows + r;
  const bool valid = b < P.B;
  const uint32_t rowoff = (uint32_t)r * 128u, rsw = (uint32_t)(r & 7);
  const uint32_t t_row = tmem_d + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = idesc_bf16_f32(kRows, NB, 0, 0);

  float dc[kH];
#pragma unroll
  for (int u = 0; u < kH; ++u) dc[u] = 0.0f;
  bool pending = false, acc_valid = false;
  uint32_t waits = 0;

  for (int pi = 0; pi < P.nph; ++pi) {
    const BwdTcPhase& ph = P.ph[pi];
    const int T = ph.T;
    if (pending) {                                       // the weights may only change once the last MMA has read them
      mbar_wait(smem_u32(&bk->tmem_full), waits & 1u);
      ++waits; pending = false; acc_valid = true;
    }
    if (ph.zero_carry) {
      acc_valid = false;
#pragma unroll
      for (int u = 0; u < kH; ++u) dc[u] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    // ---- [U | W]^T: row n of the Keras tensor = B-operand row n, 256 k values = four 64-wide swizzled tiles ----
    for (int item = tid; item < NB * 4; item += kRows) {
      const int n = item >> 2, kt = item & 3;
      const float* src = nullptr;
      if (n < kH) src = ph.Uk + (size_t)n * kG + kt * 64;
      else if (AR && ph.ar && n - kH < ph.in_dim) src = ph.Wk + (size_t)(n - kH) * kG + kt * 64;
      const uint32_t nsw = (uint32_t)(n & 7), noff = (uint32_t)n * 128u;
#pragma unroll 2
      for (int c8 = 0; c8 < 8; ++c8) {
        float v[8];
        if (src) load8(src + c8 * 8, v);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.0f;
        }
        uint2 lo[NS], hi[NS];
        split4<NS>(make_float4(v[0], v[1], v[2], v[3]), lo);
        split4<NS>(make_float4(v[4], v[5], v[6], v[7]), hi);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          *reinterpret_cast<uint4*>(smem + (uint32_t)(s * 4 + kt) * kBTile + noff + ((((uint32_t)c8) ^ nsw) << 4)) =
              make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    const bool next_carries = (pi + 1 < P.nph) && !P.ph[pi + 1].zero_carry;
    for (int t = T - 1; t >= 0; --t) {
      if (pending) {
        mbar_wait(smem_u32(&bk->tmem_full), waits & 1u);
        ++waits; pending = false; acc_valid = true;
      }
      __syncwarp();
      tc_fence_after();
      const size_t rowt = (size_t)b * T + t;

      // ---- Dense head: dpre = (dy + dx_{t+1}) * act'(y) ----
      float dpre[OD];
#pragma unroll
      for (int d = 0; d < OD; ++d) dpre[d] = 0.0f;
      if (ph.has_head) {
        float dxv[kXK];
#pragma unroll
        for (int d = 0; d < kXK; ++d) dxv[d] = 0.0f;
        if (AR && ph.ar && acc_valid && t < T - 1) {
          tmem_ld16(t_row + kH, dxv);
          tmem_ld_wait();
        }
        const size_t o = rowt * (size_t)P.out_dim;
#pragma unroll
        for (int d = 0; d < OD; ++d) {
          const bool live = valid && d < P.out_dim;
          float gy = live ? __ldg(&ph.dy[o + d]) : 0.0f;
          if (d < kXK) gy += dxv[d];
          const float yv = live ? __ldg(&ph.y[o + d]) : 0.0f;
          const float dp = live ? gy * fov_act_grad(P.head_act, yv) : 0.0f;
          dpre[d] = dp;
          if (live) ph.dpre[o + d] = dp;
        }
      }

      const float* g_row = ph.gates + rowt * kG;
      const float* c_row = ph.c + rowt * kH;
      const float* cp_row = t > 0 ? c_row - kH : (ph.c_init ? ph.c_init + (size_t)b * ph.c_init_stride : nullptr);
      const float* dhs_row = ph.dhseq ? ph.dhseq + rowt * kH : nullptr;
      float* dz_row = ph.dz + rowt * kG;
      // the row's saved tensors of the NEXT (earlier) step -> L2 while this step computes: with four warps per SM the
      // eight load / compute passes of a step would otherwise each expose a full DRAM round trip
      if (valid && t > 0) {
        const char* pg = reinterpret_cast<const char*>(g_row - kG);
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(pg + i * 128));
        if (t > 1) {
          const char* pc = reinterpret_cast<const char*>(c_row - 2 * kH);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pc));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + 128));
        }
        if (dhs_row) {
          const char* pd = reinterpret_cast<const char*>(dhs_row - kH);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pd));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pd + 128));
        }
      }
#pragma unroll
      for (int p8 = 0; p8 < 8; ++p8) {
        float acc8[8], gi8[8], gf8[8], gg8[8], go8[8], ct8[8], cp8[8], dhs8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc8[j] = 0.0f; gi8[j] = 0.0f; gf8[j] = 0.0f; gg8[j] = 0.0f; go8[j] = 0.0f;
                                       ct8[j] = 0.0f; cp8[j] = 0.0f; dhs8[j] = 0.0f; }
        if (acc_valid) tmem_ld8(t_row + p8 * 8, acc8);
        if (valid) {
          load8(g_row + p8 * 8, gi8); load8(g_row + kH + p8 * 8, gf8);
          load8(g_row + 2 * kH + p8 * 8, gg8); load8(g_row + 3 * kH + p8 * 8, go8);
          load8(c_row + p8 * 8, ct8);
          if (cp_row) load8(cp_row + p8 * 8, cp8);
          if (dhs_row) load8(dhs_row + p8 * 8, dhs8);
        }
        if (acc_valid) {
          tmem_ld_wait();
          pin8(acc8);
        }
        float dz[4][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int u = p8 * 8 + j;
          float dh = acc8[j] + dhs8[j];
          if (ph.has_head) {
            const float4* wr = reinterpret_cast<const float4*>(&bk->Wo_s[u * OD]);
#pragma unroll
            for (int d4 = 0; d4 < OD / 4; ++d4) {
              const float4 w4 = wr[d4];
              dh = fmaf(dpre[d4 * 4], w4.x, dh); dh = fmaf(dpre[d4 * 4 + 1], w4.y, dh);
              dh = fmaf(dpre[d4 * 4 + 2], w4.z, dh); dh = fmaf(dpre[d4 * 4 + 3], w4.w, dh);
            }
          }
          const float tc_ = tanh5(ct8[j]);
          const float dog = dh * tc_;
          const float dct = fmaf(dh * go8[j], 1.0f - tc_ * tc_, dc[u]);
          dc[u] = valid ? dct * gf8[j] : 0.0f;
          dz[0][j] = dct * gg8[j] * rec_grad<REC>(gi8[j]);
          dz[1][j] = dct * cp8[j] * rec_grad<REC>(gf8[j]);
          dz[2][j] = dct * gi8[j] * (1.0f - gg8[j] * gg8[j]);
          dz[3][j] = dog * rec_grad<REC>(go8[j]);
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          if (valid) store8(dz_row + gi * kH + p8 * 8, dz[gi], 4);
          uint2 lo[NS], hi[NS];
          split4<NS>(make_float4(dz[gi][0], dz[gi][1], dz[gi][2], dz[gi][3]), lo);
          split4<NS>(make_float4(dz[gi][4], dz[gi][5], dz[gi][6], dz[gi][7]), hi);
#pragma unroll
          for (int s = 0; s < NS; ++s)
            *reinterpret_cast<uint4*>(smem + kAOff + (uint32_t)(s * 4 + gi) * kATerm + rowoff + ((((uint32_t)p8) ^ rsw) << 4)) =
                make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
        }
      }
      tc_fence_before();
      acc_valid = false;
      const bool need_mma = t > 0 || next_carries;
      if (need_mma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          uint32_t acc = 0;
#pragma unroll
          for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
              for (int sum = NS - 1; sum >= 0; --sum) {
#pragma unroll
                for (int sa = 0; sa <= sum; ++sa) {
                  const int sb = sum - sa;
                  umma_bf16(tmem_d, desc_at(kDescHi128, base + kAOff + (uint32_t)(sa * 4 + kt) * kATerm + k4 * 32),
                            desc_at(kDescHi128, base + (uint32_t)(sb * 4 + kt) * kBTile + k4 * 32), idesc, acc);
                  acc = 1;
                }
              }
            }
          }
          umma_commit(smem_u32(&bk->tmem_full));
        }
        pending = true;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

template <int NS, int REC, int OD, bool AR>
int launch_tc_bwd(const LstmTcBwdParams& P, cudaStream_t st) {
  constexpr int NB = AR ? kH + kXK : kH;
  const size_t smem = (size_t)NS * 4 * NB * 128 + (size_t)NS * 4 * kATerm + sizeof(LstmTcBwdBook<OD>) + 1024;
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_bwd_kernel<NS, REC, OD, AR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      fov_set_error("fov_lstm BPTT (tensor-core): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  lstm_tc_bwd_kernel<NS, REC, OD, AR><<<(P.B + kRows - 1) / kRows, kRows, smem, st>>>(P);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
template <int NS, int REC>
int launch_tc_bwd_oa(const LstmTcBwdParams& P, bool ar, cudaStream_t st) {
  if (P.out_dim <= 8) return ar ? launch_tc_bwd<NS, REC, 8, true>(P, st) : launch_tc_bwd<NS, REC, 8, false>(P, st);
  return ar ? launch_tc_bwd<NS, REC, 16, true>(P, st) : launch_tc_bwd<NS, REC, 16, false>(P, st);
}

}  // namespace

// teacher-forced phases never read the input kernel, so any input width works; autoregressive decoding needs in_dec <= 16
bool lstm_tc_bwd_supported(const fov_lstm_cfg* cfg) {
  if (cfg->H != kH || (cfg->math != FOV_MATH_BF16 && cfg->math != FOV_MATH_BF16X2)) return false;
  if (cfg->out_dim > 16) return false;
  if (cfg->T_dec > 0 && !cfg->teacher_forcing && cfg->in_dec > kXK) return false;
  return true;
}

int lstm_tc_bwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w, const fov_lstm_io* io, const fov_lstm_grads* g,
                cudaStream_t st) {
  LstmTcBwdParams P{};
  P.B = cfg->B; P.out_dim = cfg->T_dec > 0 ? cfg->out_dim : 0; P.head_act = cfg->head_act; P.Wo = w->head_kernel;
  const bool ar = cfg->T_dec > 0 && cfg->teacher_forcing == 0;
  int n = 0;
  if (cfg->T_dec > 0) {
    BwdTcPhase& ph = P.ph[n++];
    ph.Uk = w->dec_recurrent; ph.Wk = w->dec_kernel; ph.in_dim = cfg->in_dec; ph.T = cfg->T_dec;
    ph.ar = ar; ph.has_head = P.out_dim > 0; ph.zero_carry = 1;
    ph.dy = g->dy; ph.y = g->y; ph.dhseq = g->dhseq_dec; ph.dpre = g->dpre; ph.dz = g->dz_dec;
    ph.gates = io->dec.gates; ph.c = io->dec.c;
    if (cfg->dec_zero_init) { ph.c_init = nullptr; ph.c_init_stride = 0; }
    else if (cfg->T_enc > 0) { ph.c_init = io->enc.c + (size_t)(cfg->T_enc - 1) * kH; ph.c_init_stride = (long long)cfg->T_enc * kH; }
    else { ph.c_init = io->c0; ph.c_init_stride = kH; }
  }
  if (cfg->T_enc > 0) {
    BwdTcPhase& ph = P.ph[n++];
    ph.Uk = w->enc_recurrent; ph.Wk = w->enc_kernel; ph.in_dim = cfg->in_enc; ph.T = cfg->T_enc;
    ph.ar = 0; ph.has_head = 0; ph.zero_carry = (n == 1 || cfg->dec_zero_init) ? 1 : 0;
    ph.dy = nullptr; ph.y = nullptr; ph.dhseq = g->dhseq_enc; ph.dpre = nullptr; ph.dz = g->dz_enc;
    ph.gates = io->enc.gates; ph.c = io->enc.c; ph.c_init = io->c0; ph.c_init_stride = kH;
  }
  P.nph = n;
  auto a16 = [](const void* q) { return (uintptr_t)q % 16 == 0; };
  FOV_CHECK_ARG(a16(io->enc.gates) && a16(io->enc.c) && a16(io->dec.gates) && a16(io->dec.c) && a16(g->dz_enc) &&
                    a16(g->dz_dec) && a16(g->dhseq_enc) && a16(g->dhseq_dec) && a16(io->c0) && a16(w->enc_recurrent) && a16(w->dec_recurrent) &&
                    a16(w->dec_kernel),
                "tensor-core fc-LSTM BPTT needs 16-byte aligned tensors");
  const bool hs = cfg->rec_act == FOV_REC_HARD_SIGMOID;
  if (cfg->math == FOV_MATH_BF16)
    return hs ? launch_tc_bwd_oa<1, FOV_REC_HARD_SIGMOID>(P, ar, st) : launch_tc_bwd_oa<1, FOV_REC_SIGMOID>(P, ar, st);
  return hs ? launch_tc_bwd_oa<2, FOV_REC_HARD_SIGMOID>(P, ar, st) : launch_tc_bwd_oa<2, FOV_REC_SIGMOID>(P, ar, st);
}

extern int g_lstm_xproj_mode;    // lstm_seq2seq.cu (fov_debug_lstm_xproj): -1 never, 0 inputs wider than 16, 1 whenever allowed

namespace {
// a phase takes the time-batched projection when its inputs are known in advance (not autoregressive), the shape fits
// the TMA view and either the input is wider than the fused x operand (16) or the debug switch asks for it
bool phase_wants_xproj(const fov_lstm_cfg* cfg, bool dec) {
  const int in = dec ? cfg->in_dec : cfg->in_enc, T = dec ? cfg->T_dec : cfg->T_enc;
  if (T <= 0 || g_lstm_xproj_mode < 0 || (dec && !cfg->teacher_forcing)) return false;
  if (in <= kXK && g_lstm_xproj_mode == 0) return false;
  return lstm_xproj_supported(cfg->B, T, in, cfg->math, nullptr);
}
}  // namespace

bool lstm_tc_supported(const fov_lstm_cfg* cfg) {
  if (cfg->H != kH || cfg->math < FOV_MATH_BF16 || cfg->math > FOV_MATH_BF16X3) return false;
  if (cfg->T_enc > 0 && cfg->in_enc > kXK && !phase_wants_xproj(cfg, false)) return false;
  if (cfg->T_dec > 0 && cfg->in_dec > kXK && !phase_wants_xproj(cfg, true)) return false;
  if (cfg->out_dim > 16) return false;
  return true;
}

// floats of fov_lstm_io.ws the tensor-core forward wants for this configuration (0: none)
size_t lstm_tc_fwd_ws_floats(const fov_lstm_cfg* cfg) {
  if (!lstm_tc_supported(cfg)) return 0;
  size_t n = 0;
  if (phase_wants_xproj(cfg, false)) n += lstm_xproj_ws_floats(cfg->B, cfg->T_enc);
  if (phase_wants_xproj(cfg, true)) n += lstm_xproj_ws_floats(cfg->B, cfg->T_dec);
  return n;
}

int lstm_tc_fwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w, const fov_lstm_io* io, cudaStream_t st) {
  LstmTcParams P{};
  P.B = cfg->B; P.out_dim = cfg->T_dec > 0 ? cfg->out_dim : 0; P.head_act = cfg->head_act; P.training = cfg->training;
  P.dbg = g_lstm_tc_dbg;
  P.Wo = w->head_kernel; P.bo = w->head_bias; P.h0 = io->h0; P.c0 = io->c0; P.hT = io->hT; P.cT = io->cT;
  int n = 0;
  if (cfg->T_enc > 0) {
    TcPhase& ph = P.ph[n++];
    ph.Wk = w->enc_kernel; ph.Uk = w->enc_recurrent; ph.bk = w->enc_bias;
    ph.x = io->x_enc; ph.in_dim = cfg->in_enc; ph.T = cfg->T_enc; ph.x_T = cfg->T_enc;
    ph.ar = 0; ph.has_head = 0; ph.zero_init = 0; ph.extra = nullptr; ph.y = nullptr; ph.sv = io->enc;
    ph.xproj = nullptr;
  }
  if (cfg->T_dec > 0) {
    TcPhase& ph = P.ph[n++];
    ph.Wk = w->dec_kernel; ph.Uk = w->dec_recurrent; ph.bk = w->dec_bias;
    ph.x = io->x_dec; ph.in_dim = cfg->in_dec; ph.T = cfg->T_dec;
    ph.ar = cfg->teacher_forcing == 0; ph.x_T = ph.ar ? 1 : cfg->T_dec;
    ph.has_head = P.out_dim > 0; ph.zero_init = cfg->dec_zero_init ? 1 : 0;
    ph.extra = io->extra; ph.y = io->y; ph.sv = io->dec;
    ph.xproj = nullptr;
  }
  P.nph = n;
  // ---- time-batched input projections (xproj_tc.cu): P = x . W for every step of a phase, one TMA-fed launch each ----
  bool any_xp = false;
  {
    float* wsp = io->ws;
    for (int i = 0; i < n; ++i) {
      TcPhase& ph = P.ph[i];
      const bool dec = cfg->T_enc > 0 ? i == 1 : true;
      const bool want = phase_wants_xproj(cfg, dec) && wsp != nullptr && (uintptr_t)ph.x % 16 == 0;
      if (!want) {
        if (ph.in_dim > kXK) {
          fov_set_error("fov_lstm (tensor-core): a %d-wide input needs the time-batched projection: pass fov_lstm_io.ws "
                        "(fov_lstm_fwd_ws_bytes) and a 16-byte aligned input", ph.in_dim);
          return FOV_ERR_ARG;
        }
        continue;
      }
      float* xh = (cfg->training && ph.sv.xh) ? ph.sv.xh : nullptr;
      int rc = lstm_xproj_run(cfg->B, ph.T, ph.in_dim, cfg->math, ph.x, ph.Wk, wsp, xh, fov_lstm_xh_stride(kH, ph.in_dim), st);
      if (rc) return rc;
      ph.xproj = wsp;
      wsp += lstm_xproj_ws_floats(cfg->B, ph.T);
      any_xp = true;
    }
  }
  auto a16 = [](const void* q) { return (uintptr_t)q % 16 == 0; };
  FOV_CHECK_ARG(a16(io->enc.gates) && a16(io->enc.c) && a16(io->enc.hseq) && a16(io->dec.gates) && a16(io->dec.c) &&
                    a16(io->dec.hseq) && a16(io->enc.xh) && a16(io->dec.xh) && a16(io->hT) && a16(io->cT),
                "tensor-core fc-LSTM needs 16-byte aligned state / saved tensors");
  // two groups per CTA (MMA / epilogue overlap) once there are enough sequences to fill the SMs with one group each
  const bool two = !any_xp && cfg->math <= FOV_MATH_BF16X2 && cfg->B > kRows * fov_num_sms();
  switch (cfg->math) {
    case FOV_MATH_BF16: return two ? launch_tc_ro<1, 2>(P, cfg->rec_act, st) : launch_tc_ro<1, 1>(P, cfg->rec_act, st);
    case FOV_MATH_BF16X2: return two ? launch_tc_ro<2, 2>(P, cfg->rec_act, st) : launch_tc_ro<2, 1>(P, cfg->rec_act, st);
    default: return launch_tc_ro<3, 1>(P, cfg->rec_act, st);
  }
}
