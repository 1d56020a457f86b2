// Persistent fc-LSTM encoder-decoder, forward and BPTT (fp32 parity path).
//
// One CTA owns a tile of BT sequences for ALL T_enc + T_dec timesteps: the gate
// weights [U;W] live in shared memory (repacked so one LDS.128 yields the four
// gate weights of a hidden unit), the operand tile [h_{t-1}; x_t] is double
// buffered in shared memory, gates / cell state stay in registers, the Dense head
// (+ optional additive term) runs inside the loop so autoregressive decoding needs
// no per-step launch.  Thread (u, sg) owns hidden unit u for SPT sequences.
//
// Replaces the Keras LSTM/Dense calls cited in include/fov360.h.
#include "fov_common.cuh"
#include "fov_internal.h"

// diagnostics / A-B testing: -1 = never use the tensor-core forward, 0 = choose, 1 = whenever the shape allows
static int g_lstm_tc_mode = 0;
extern "C" void fov_debug_lstm_tc(int mode) { g_lstm_tc_mode = mode; }
// diagnostics / A-B: 0 = never use the 4 / 8-sequence tiles of the fp32 kernels at small batches
static int g_lstm_no_small = 0;
extern "C" void fov_debug_lstm_small_tiles(int on) { g_lstm_no_small = !on; }
// time-batched input projection in front of the tensor-core forward: -1 never, 0 choose (inputs wider than 16), 1 always
int g_lstm_xproj_mode = 0;
extern "C" void fov_debug_lstm_xproj(int mode) { g_lstm_xproj_mode = mode; }
// A/B switch for the tensor-core LSTM weight gradient (needs fov_lstm_grads.ws).  History: on unpadded 70-float [h|x]
// rows it took the unaligned gather path of wgrad_tc.cu and lost to the SIMT kernels (11.07 vs 10.96 ms per config-2
// step), hence the rows padded to a multiple of 4 floats and the single aligned launch per LSTM.
static int g_lstm_wgrad_tc = 1;
extern "C" void fov_debug_lstm_wgrad_tc(int on) { g_lstm_wgrad_tc = on; }
static int g_lstm_bptt_tc = 1;   // A/B switch
extern "C" void fov_debug_lstm_bptt_tc(int on) { g_lstm_bptt_tc = on; }

namespace {

constexpr int kTcTrainMinB = 6144;   // sequences from which the tensor-core training path (forward + BPTT) wins
constexpr int kH = 64;     // latent_dim
constexpr int kG = 256;    // 4H
constexpr int kNT = 256;   // threads per CTA
constexpr int kMaxOut = 16;

struct LstmKParams {
  fov_lstm_cfg cfg;
  fov_lstm_weights w;
  fov_lstm_io io;
  fov_lstm_grads g;
};

struct PhaseDesc {
  const float *Wk, *Uk, *bk;
  const float* x;
  int in_dim, T, x_T;
  bool ar, has_head;
  const float* extra;
  float* y;
  fov_lstm_saved sv;
};

template <int SPT, int REC>
__device__ __forceinline__ void lstm_fwd_phase(const PhaseDesc& ph, float* Wsm, float* Abuf, int KR,
                                               int& gstep, float (&c)[SPT], const float* Wo_s,
                                               const float* bo_s, int out_dim, int head_act,
                                               bool training, int b0, int nvalid) {
  constexpr int BT = 4 * SPT, BTS = BT + 4;
  const int tid = threadIdx.x, u = tid & 63, sg = tid >> 6, s0 = sg * SPT;
  const int in_dim = ph.in_dim, K = kH + in_dim, T = ph.T;

  __syncthreads();
  // gate weights -> smem, repacked [k][u][gate]
  for (int idx = tid; idx < K * kG; idx += kNT) {
    int k = idx >> 8, j = idx & 255;
    float v = (k < kH) ? ph.Uk[k * kG + j] : ph.Wk[(k - kH) * kG + j];
    Wsm[k * kG + (j & 63) * 4 + (j >> 6)] = v;
  }
  const float bi = ph.bk[u], bf = ph.bk[64 + u], bg = ph.bk[128 + u], bo = ph.bk[192 + u];
  {
    float* cur = Abuf + (gstep & 1) * KR * BTS;
    for (int idx = tid; idx < BT * in_dim; idx += kNT) {
      int s = idx / in_dim, k = idx - s * in_dim;
      float v = (s < nvalid) ? ph.x[((size_t)(b0 + s) * ph.x_T) * in_dim + k] : 0.0f;
      cur[(kH + k) * BTS + s] = v;
    }
  }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const float* cur = Abuf + (gstep & 1) * KR * BTS;
    float* nxt = Abuf + ((gstep + 1) & 1) * KR * BTS;
    if (!ph.ar && t + 1 < T) {
      for (int idx = tid; idx < BT * in_dim; idx += kNT) {
        int s = idx / in_dim, k = idx - s * in_dim;
        float* dst = &nxt[(kH + k) * BTS + s];
        if (s < nvalid) fov_cp_async4(dst, &ph.x[((size_t)(b0 + s) * T + t + 1) * in_dim + k]);
        else *dst = 0.0f;
      }
    }
    if (training && ph.sv.xh) {
      const int XS = fov_lstm_xh_stride(kH, in_dim);          // rows padded with zeros to a multiple of 4 floats
      for (int idx = tid; idx < BT * XS; idx += kNT) {
        int s = idx / XS, k = idx - s * XS;
        if (s < nvalid) ph.sv.xh[((size_t)(b0 + s) * T + t) * XS + k] = k < K ? cur[k * BTS + s] : 0.0f;
      }
    }
    float acc[4][SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) { acc[0][i] = bi; acc[1][i] = bf; acc[2][i] = bg; acc[3][i] = bo; }
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(&Wsm[k * kG + u * 4]);
      float a[SPT];
      if constexpr (SPT % 4 == 0) {
#pragma unroll
        for (int q = 0; q < SPT / 4; ++q) {
          float4 v = *reinterpret_cast<const float4*>(&cur[k * BTS + s0 + q * 4]);
          a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
        }
      } else {                                   // small-batch tiles (1 / 2 sequences per thread)
#pragma unroll
        for (int i = 0; i < SPT; ++i) a[i] = cur[k * BTS + s0 + i];
      }
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        acc[0][i] = fmaf(w.x, a[i], acc[0][i]);
        acc[1][i] = fmaf(w.y, a[i], acc[1][i]);
        acc[2][i] = fmaf(w.z, a[i], acc[2][i]);
        acc[3][i] = fmaf(w.w, a[i], acc[3][i]);
      }
    }
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = s0 + i;
      const float ig = fov_rec_act<REC>(acc[0][i]);
      const float fg = fov_rec_act<REC>(acc[1][i]);
      const float gg = tanhf(acc[2][i]);
      const float og = fov_rec_act<REC>(acc[3][i]);
      c[i] = fmaf(fg, c[i], ig * gg);
      const float h = og * tanhf(c[i]);
      nxt[u * BTS + s] = h;
      if (s < nvalid) {
        const size_t row = (size_t)(b0 + s) * T + t;
        if (training && ph.sv.gates) {
          float* gp = ph.sv.gates + row * kG + u;
          gp[0] = ig; gp[64] = fg; gp[128] = gg; gp[192] = og;
        }
        if (training && ph.sv.c) ph.sv.c[row * kH + u] = c[i];
        if (ph.sv.hseq) ph.sv.hseq[row * kH + u] = h;
      }
    }
    fov_cp_async_wait_all();
    __syncthreads();
    if (ph.has_head) {
      for (int idx = tid; idx < BT * out_dim; idx += kNT) {
        const int s = idx % BT, d = idx / BT;
        float sum = bo_s[d];
#pragma unroll 8
        for (int uu = 0; uu < kH; ++uu) sum = fmaf(nxt[uu * BTS + s], Wo_s[uu * out_dim + d], sum);
        const bool valid = s < nvalid;
        const size_t o = ((size_t)(b0 + s) * T + t) * out_dim + d;
        if (ph.extra && valid) sum += ph.extra[o];
        const float yv = fov_act(head_act, sum);
        if (valid) ph.y[o] = yv;
        if (ph.ar && t + 1 < T) nxt[(kH + d) * BTS + s] = valid ? yv : 0.0f;
      }
      if (ph.ar) __syncthreads();
    }
    ++gstep;
  }
}

template <int SPT, int REC>
__global__ void __launch_bounds__(kNT) lstm_seq2seq_fwd_kernel(const __grid_constant__ LstmKParams P) {
  constexpr int BT = 4 * SPT, BTS = BT + 4;
  extern __shared__ __align__(16) float smem[];
  const fov_lstm_cfg& cfg = P.cfg;
  const int in_e = cfg.T_enc > 0 ? cfg.in_enc : 0, in_d = cfg.T_dec > 0 ? cfg.in_dec : 0;
  const int KR = kH + (in_e > in_d ? in_e : in_d);
  float* Wsm = smem;
  float* Abuf = Wsm + KR * kG;
  float* Wo_s = Abuf + 2 * KR * BTS;
  float* bo_s = Wo_s + kH * kMaxOut;
  const int tid = threadIdx.x, u = tid & 63, sg = tid >> 6, s0 = sg * SPT;
  const int b0 = blockIdx.x * BT;
  const int nvalid = min(BT, cfg.B - b0);
  const int out_dim = cfg.out_dim;

  float c[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = s0 + i;
    const bool valid = s < nvalid;
    c[i] = (P.io.c0 && valid) ? P.io.c0[(size_t)(b0 + s) * kH + u] : 0.0f;
    Abuf[u * BTS + s] = (P.io.h0 && valid) ? P.io.h0[(size_t)(b0 + s) * kH + u] : 0.0f;
  }
  if (out_dim > 0) {
    for (int idx = tid; idx < kH * out_dim; idx += kNT) Wo_s[idx] = P.w.head_kernel[idx];
    if (tid < out_dim) bo_s[tid] = P.w.head_bias[tid];
  }
  int gstep = 0;
  if (cfg.T_enc > 0) {
    PhaseDesc ph;
    ph.Wk = P.w.enc_kernel; ph.Uk = P.w.enc_recurrent; ph.bk = P.w.enc_bias;
    ph.x = P.io.x_enc; ph.in_dim = cfg.in_enc; ph.T = cfg.T_enc; ph.x_T = cfg.T_enc;
    ph.ar = false; ph.has_head = false; ph.extra = nullptr; ph.y = nullptr; ph.sv = P.io.enc;
    lstm_fwd_phase<SPT, REC>(ph, Wsm, Abuf, KR, gstep, c, Wo_s, bo_s, out_dim, cfg.head_act,
                             cfg.training != 0, b0, nvalid);
  }
  if (cfg.T_dec > 0) {
    if (cfg.dec_zero_init) {
      __syncthreads();
      float* cur = Abuf + (gstep & 1) * KR * BTS;
#pragma unroll
      for (int i = 0; i < SPT; ++i) { c[i] = 0.0f; cur[u * BTS + s0 + i] = 0.0f; }
    }
    PhaseDesc ph;
    ph.Wk = P.w.dec_kernel; ph.Uk = P.w.dec_recurrent; ph.bk = P.w.dec_bias;
    ph.x = P.io.x_dec; ph.in_dim = cfg.in_dec; ph.T = cfg.T_dec;
    ph.ar = cfg.teacher_forcing == 0; ph.x_T = ph.ar ? 1 : cfg.T_dec;
    ph.has_head = out_dim > 0; ph.extra = P.io.extra; ph.y = P.io.y; ph.sv = P.io.dec;
    lstm_fwd_phase<SPT, REC>(ph, Wsm, Abuf, KR, gstep, c, Wo_s, bo_s, out_dim, cfg.head_act,
                             cfg.training != 0, b0, nvalid);
  }
  const float* fin = Abuf + (gstep & 1) * KR * BTS;
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = s0 + i;
    if (s < nvalid) {
      if (P.io.hT) P.io.hT[(size_t)(b0 + s) * kH + u] = fin[u * BTS + s];
      if (P.io.cT) P.io.cT[(size_t)(b0 + s) * kH + u] = c[i];
    }
  }
}

// --------------------------------------------------------------------------- //
// BPTT
// --------------------------------------------------------------------------- //

struct BwdPhaseDesc {
  const float *Wk, *Uk;        // kernel (in,4H) (only read in AR mode), recurrent (H,4H)
  int in_dim, T;
  bool ar, has_head;
  const float* dy;             // (B,T,out)
  const float* y;
  const float* dhseq;          // optional (B,T,H)
  float* dpre;                 // (B,T,out)
  float* dz;                   // (B,T,4H)
  fov_lstm_saved sv;
  const float* c_init;         // c_{-1}: base pointer, per-sample stride c_init_stride; may be NULL
  size_t c_init_stride;
};

template <int SPT, int REC>
__device__ __forceinline__ void lstm_bwd_phase(const BwdPhaseDesc& ph, float* UT, float* Wd_s, float* dzs,
                                               float* dpre_s, float* dxs, const float* Wo_s,
                                               int out_dim, int head_act, float (&dhr)[SPT],
                                               float (&dc)[SPT], int b0, int nvalid) {
  constexpr int BT = 4 * SPT, BTS = BT + 4;
  const int tid = threadIdx.x, u = tid & 63, sg = tid >> 6, s0 = sg * SPT;
  const int T = ph.T, in_dim = ph.in_dim;

  __syncthreads();
  for (int idx = tid; idx < kH * kG; idx += kNT) {       // UT[j][u] = U[u][j]
    int k = idx >> 8, j = idx & 255;
    UT[j * kH + k] = ph.Uk[idx];
  }
  if (ph.ar)
    for (int idx = tid; idx < in_dim * kG; idx += kNT) Wd_s[idx] = ph.Wk[idx];
  float wo[kMaxOut];
#pragma unroll
  for (int d = 0; d < kMaxOut; ++d) wo[d] = (ph.has_head && d < out_dim) ? Wo_s[u * out_dim + d] : 0.0f;
  __syncthreads();

  for (int t = T - 1; t >= 0; --t) {
    if (ph.has_head) {
      for (int idx = tid; idx < BT * out_dim; idx += kNT) {
        const int s = idx % BT, d = idx / BT;
        const bool valid = s < nvalid;
        const size_t o = ((size_t)(b0 + s) * T + t) * out_dim + d;
        float gy = valid ? ph.dy[o] : 0.0f;
        if (ph.ar && t < T - 1) gy += dxs[d * BT + s];
        const float yv = valid ? ph.y[o] : 0.0f;
        const float dp = gy * fov_act_grad(head_act, yv);
        dpre_s[d * BT + s] = dp;
        if (valid) ph.dpre[o] = dp;
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = s0 + i;
      const bool valid = s < nvalid;
      float dh = dhr[i];
      if (ph.has_head) {
#pragma unroll
        for (int d = 0; d < kMaxOut; ++d)
          if (d < out_dim) dh = fmaf(dpre_s[d * BT + s], wo[d], dh);
      }
      float dzi = 0.f, dzf = 0.f, dzg = 0.f, dzo = 0.f;
      if (valid) {
        const size_t row = (size_t)(b0 + s) * T + t;
        if (ph.dhseq) dh += ph.dhseq[row * kH + u];
        const float* gp = ph.sv.gates + row * kG + u;
        const float ig = gp[0], fg = gp[64], gg = gp[128], og = gp[192];
        const float ct = ph.sv.c[row * kH + u];
        float cp;
        if (t > 0) cp = ph.sv.c[(row - 1) * kH + u];
        else cp = ph.c_init ? ph.c_init[(size_t)(b0 + s) * ph.c_init_stride + u] : 0.0f;
        const float tc = tanhf(ct);
        const float dog = dh * tc;
        const float dct = fmaf(dh * og, 1.0f - tc * tc, dc[i]);
        dc[i] = dct * fg;
        dzi = dct * gg * fov_rec_act_grad<REC>(ig);
        dzf = dct * cp * fov_rec_act_grad<REC>(fg);
        dzg = dct * ig * (1.0f - gg * gg);
        dzo = dog * fov_rec_act_grad<REC>(og);
        float* zp = ph.dz + row * kG + u;
        zp[0] = dzi; zp[64] = dzf; zp[128] = dzg; zp[192] = dzo;
      } else {
        dc[i] = 0.0f;
      }
      dzs[(u)*BTS + s] = dzi;
      dzs[(64 + u) * BTS + s] = dzf;
      dzs[(128 + u) * BTS + s] = dzg;
      dzs[(192 + u) * BTS + s] = dzo;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SPT; ++i) dhr[i] = 0.0f;
#pragma unroll 4
    for (int j = 0; j < kG; ++j) {
      const float w = UT[j * kH + u];
      if constexpr (SPT % 4 == 0) {
#pragma unroll
        for (int q = 0; q < SPT / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&dzs[j * BTS + s0 + q * 4]);
          dhr[q * 4 + 0] = fmaf(w, v.x, dhr[q * 4 + 0]);
          dhr[q * 4 + 1] = fmaf(w, v.y, dhr[q * 4 + 1]);
          dhr[q * 4 + 2] = fmaf(w, v.z, dhr[q * 4 + 2]);
          dhr[q * 4 + 3] = fmaf(w, v.w, dhr[q * 4 + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < SPT; ++i) dhr[i] = fmaf(w, dzs[j * BTS + s0 + i], dhr[i]);
      }
    }
    if (ph.ar && t > 0) {
      for (int idx = tid; idx < BT * in_dim; idx += kNT) {
        const int s = idx % BT, d = idx / BT;
        float sum = 0.0f;
#pragma unroll 8
        for (int j = 0; j < kG; ++j) sum = fmaf(dzs[j * BTS + s], Wd_s[d * kG + j], sum);
        dxs[d * BT + s] = sum;
      }
    }
    __syncthreads();
  }
}

template <int SPT, int REC>
__global__ void __launch_bounds__(kNT) lstm_seq2seq_bwd_kernel(const __grid_constant__ LstmKParams P) {
  constexpr int BT = 4 * SPT, BTS = BT + 4;
  extern __shared__ __align__(16) float smem[];
  const fov_lstm_cfg& cfg = P.cfg;
  float* UT = smem;                         // 256*64
  float* dzs = UT + kG * kH;                // 256*BTS
  float* Wd_s = dzs + kG * BTS;             // kMaxOut*256
  float* Wo_s = Wd_s + kMaxOut * kG;        // 64*kMaxOut
  float* dpre_s = Wo_s + kH * kMaxOut;      // kMaxOut*BT
  float* dxs = dpre_s + kMaxOut * BT;       // kMaxOut*BT
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT;
  const int nvalid = min(BT, cfg.B - b0);
  const int out_dim = cfg.out_dim;
  if (out_dim > 0)
    for (int idx = tid; idx < kH * out_dim; idx += kNT) Wo_s[idx] = P.w.head_kernel[idx];

  float dhr[SPT], dc[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) { dhr[i] = 0.0f; dc[i] = 0.0f; }

  if (cfg.T_dec > 0) {
    BwdPhaseDesc ph;
    ph.Wk = P.w.dec_kernel; ph.Uk = P.w.dec_recurrent; ph.in_dim = cfg.in_dec; ph.T = cfg.T_dec;
    ph.ar = cfg.teacher_forcing == 0; ph.has_head = out_dim > 0;
    ph.dy = P.g.dy; ph.y = P.g.y; ph.dhseq = P.g.dhseq_dec; ph.dpre = P.g.dpre; ph.dz = P.g.dz_dec;
    ph.sv = P.io.dec;
    if (cfg.dec_zero_init) { ph.c_init = nullptr; ph.c_init_stride = 0; }
    else if (cfg.T_enc > 0) {
      ph.c_init = P.io.enc.c + (size_t)(cfg.T_enc - 1) * kH; ph.c_init_stride = (size_t)cfg.T_enc * kH;
    } else { ph.c_init = P.io.c0; ph.c_init_stride = kH; }
    lstm_bwd_phase<SPT, REC>(ph, UT, Wd_s, dzs, dpre_s, dxs, Wo_s, out_dim, cfg.head_act, dhr, dc,
                             b0, nvalid);
  }
  if (cfg.T_enc > 0) {
    if (cfg.dec_zero_init) {
#pragma unroll
      for (int i = 0; i < SPT; ++i) { dhr[i] = 0.0f; dc[i] = 0.0f; }
    }
    BwdPhaseDesc ph;
    ph.Wk = P.w.enc_kernel; ph.Uk = P.w.enc_recurrent; ph.in_dim = cfg.in_enc; ph.T = cfg.T_enc;
    ph.ar = false; ph.has_head = false;
    ph.dy = nullptr; ph.y = nullptr; ph.dhseq = P.g.dhseq_enc; ph.dpre = nullptr; ph.dz = P.g.dz_enc;
    ph.sv = P.io.enc; ph.c_init = P.io.c0; ph.c_init_stride = kH;
    lstm_bwd_phase<SPT, REC>(ph, UT, Wd_s, dzs, dpre_s, dxs, Wo_s, out_dim, cfg.head_act, dhr, dc,
                             b0, nvalid);
  }
}

// g_recurrent (H,4H) += ws rows [0,H); g_kernel (in,4H) += ws rows [H,H+in)   (ws = [h | x | 0]^T dZ, row-major, 4H wide)
__global__ void __launch_bounds__(256) lstm_wgrad_scatter_kernel(const float* __restrict__ ws, float* __restrict__ g_recurrent,
                                                                  float* __restrict__ g_kernel, int in_dim) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= (kH + in_dim) * kG) return;
  const float4 v = *reinterpret_cast<const float4*>(ws + i);
  float* dst = i < kH * kG ? g_recurrent + i : g_kernel + (i - kH * kG);
  float4 o = *reinterpret_cast<float4*>(dst);
  o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
  *reinterpret_cast<float4*>(dst) = o;
}

size_t fwd_smem_bytes(const fov_lstm_cfg& cfg, int spt) {
  const int bts = 4 * spt + 4;
  const int in_e = cfg.T_enc > 0 ? cfg.in_enc : 0, in_d = cfg.T_dec > 0 ? cfg.in_dec : 0;
  const int KR = kH + (in_e > in_d ? in_e : in_d);
  return sizeof(float) * ((size_t)KR * kG + 2 * (size_t)KR * bts + kH * kMaxOut + kMaxOut);
}
size_t bwd_smem_bytes(int spt) {
  const int bt = 4 * spt, bts = bt + 4;
  return sizeof(float) * ((size_t)kG * kH + (size_t)kG * bts + kMaxOut * kG + kH * kMaxOut +
                          2 * kMaxOut * bt);
}

int check_cfg(const fov_lstm_cfg* cfg) {
  FOV_CHECK_ARG(cfg != nullptr, "cfg is NULL");
  FOV_CHECK_ARG(cfg->B > 0, "B must be > 0");
  if (cfg->H != kH) { fov_set_error("fov_lstm: latent_dim %d unsupported (64 only)", cfg->H); return FOV_ERR_UNSUPPORTED; }
  FOV_CHECK_ARG(cfg->T_enc >= 0 && cfg->T_dec >= 0 && cfg->T_enc + cfg->T_dec > 0, "bad T");
  FOV_CHECK_ARG(cfg->out_dim >= 0 && cfg->out_dim <= kMaxOut, "out_dim must be in [0,16]");
  if (cfg->T_dec > 0 && !cfg->teacher_forcing) {
    FOV_CHECK_ARG(cfg->out_dim == cfg->in_dec, "autoregressive decode needs out_dim == in_dec");
    FOV_CHECK_ARG(cfg->in_dec <= kMaxOut, "autoregressive in_dec must be <= 16");
  }
  return FOV_OK;
}

template <typename K>
int launch(K kernel, const LstmKParams& P, int grid, size_t smem, cudaStream_t st) {
  if (smem > 227 * 1024) { fov_set_error("fov_lstm: needs %zu B of shared memory (> 227 KB)", smem); return FOV_ERR_UNSUPPORTED; }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { fov_set_error("fov_lstm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FOV_ERR_CUDA; }
  kernel<<<grid, kNT, smem, st>>>(P);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

extern "C" size_t fov_lstm_bwd_ws_floats(const fov_lstm_cfg* cfg) {
  if (!cfg) return 0;
  size_t n = 0;
  if (cfg->T_enc > 0) n += (size_t)fov_lstm_xh_stride(kH, cfg->in_enc) * kG;
  if (cfg->T_dec > 0) n += (size_t)fov_lstm_xh_stride(kH, cfg->in_dec) * kG;
  return n;
}

extern "C" size_t fov_lstm_fwd_ws_bytes(const fov_lstm_cfg* cfg) {
  if (!cfg || check_cfg(cfg) || cfg->math == FOV_MATH_FP32 || g_lstm_tc_mode < 0) return 0;
  return lstm_tc_fwd_ws_floats(cfg) * sizeof(float);
}

extern "C" int fov_lstm_seq2seq_fwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w,
                                    const fov_lstm_io* io, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  FOV_CHECK_ARG(w && io, "NULL weights/io");
  if (cfg->T_enc > 0) FOV_CHECK_ARG(w->enc_kernel && w->enc_recurrent && w->enc_bias && io->x_enc, "encoder pointers");
  if (cfg->T_dec > 0) FOV_CHECK_ARG(w->dec_kernel && w->dec_recurrent && w->dec_bias && io->x_dec, "decoder pointers");
  if (cfg->out_dim > 0 && cfg->T_dec > 0) FOV_CHECK_ARG(w->head_kernel && w->head_bias && io->y, "head pointers");
  if (cfg->training) {
    if (cfg->T_enc > 0) FOV_CHECK_ARG(io->enc.xh && io->enc.gates && io->enc.c, "training needs enc saved buffers");
    if (cfg->T_dec > 0) FOV_CHECK_ARG(io->dec.xh && io->dec.gates && io->dec.c && io->dec.hseq, "training needs dec saved buffers");
  }
  LstmKParams P{};
  P.cfg = *cfg; P.w = *w; P.io = *io;
  if (P.cfg.T_dec == 0) P.cfg.out_dim = 0;
  cudaStream_t st = (cudaStream_t)stream;
  // tensor-core forward (lstm_seq2seq_tc.cu): 128-sequence tiles.  Measured on B200, AR decode of the mu/var model:
  // B=512..2048 0.088 ms vs 0.096 ms fp32 (both latency bound), B=3072 0.088 vs 0.133, B=75776 0.29 vs 1.47 ms.  In
  // training mode the thread-per-row saved tensors travel as 256-bit stores / loads (one full sector per lane); whole
  // train step of the mu/var model, tensor-core forward + BPTT vs fp32 kernels (scripts/lstm_train_threshold.py):
  // B=4096 0.88 vs 0.64 ms, B=8192 0.98 vs 1.17, B=16384 1.12 vs 1.91, B=37888 2.04 vs 3.60.
  const bool wide = (cfg->T_enc > 0 && cfg->in_enc > 16) || (cfg->T_dec > 0 && cfg->in_dec > 16);
  const int tc_min_B = (cfg->training && !wide) ? kTcTrainMinB : 512;
  if (cfg->math != FOV_MATH_FP32 && g_lstm_tc_mode >= 0 && lstm_tc_supported(cfg) &&
      (g_lstm_tc_mode > 0 || cfg->B >= tc_min_B) &&
      (!wide || (io->ws != nullptr && (uintptr_t)io->x_enc % 16 == 0 && (uintptr_t)io->x_dec % 16 == 0)))
    return lstm_tc_fwd(&P.cfg, w, io, st);
  // 64 sequences per CTA when the operand tile fits, else 32
  // ... and 16 when even 32-sequence tiles would leave SMs idle (the recurrence is latency bound: smaller tiles =
  // more CTAs in flight and a shorter per-step chain)
  int spt = 16;
  if (fwd_smem_bytes(P.cfg, 16) > 200 * 1024 || cfg->B <= 32 * fov_num_sms()) spt = 8;
  if (cfg->B <= 48 * fov_num_sms()) spt = 4;
  // small batches (the reference trains at 32 / 64): 8 / 4 sequences per CTA, one CTA per SM - the per-step chain of a
  // thread shrinks with its sequences and the idle SMs take the rest
  if (cfg->B <= 8 * fov_num_sms() && !g_lstm_no_small) spt = 2;
  if (cfg->B <= 4 * fov_num_sms() && !g_lstm_no_small) spt = 1;
  const int bt = 4 * spt, grid = (cfg->B + bt - 1) / bt;
  const size_t smem = fwd_smem_bytes(P.cfg, spt);
  const bool hs = cfg->rec_act == FOV_REC_HARD_SIGMOID;
  if (spt == 1)
    return hs ? launch(lstm_seq2seq_fwd_kernel<1, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
              : launch(lstm_seq2seq_fwd_kernel<1, FOV_REC_SIGMOID>, P, grid, smem, st);
  if (spt == 2)
    return hs ? launch(lstm_seq2seq_fwd_kernel<2, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
              : launch(lstm_seq2seq_fwd_kernel<2, FOV_REC_SIGMOID>, P, grid, smem, st);
  if (spt == 4)
    return hs ? launch(lstm_seq2seq_fwd_kernel<4, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
              : launch(lstm_seq2seq_fwd_kernel<4, FOV_REC_SIGMOID>, P, grid, smem, st);
  if (spt == 16)
    return hs ? launch(lstm_seq2seq_fwd_kernel<16, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
              : launch(lstm_seq2seq_fwd_kernel<16, FOV_REC_SIGMOID>, P, grid, smem, st);
  return hs ? launch(lstm_seq2seq_fwd_kernel<8, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
            : launch(lstm_seq2seq_fwd_kernel<8, FOV_REC_SIGMOID>, P, grid, smem, st);
}

extern "C" int fov_lstm_seq2seq_bwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w,
                                    const fov_lstm_io* io, const fov_lstm_grads* g, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  FOV_CHECK_ARG(w && io && g, "NULL weights/io/grads");
  const bool has_head = cfg->out_dim > 0 && cfg->T_dec > 0;
  if (cfg->T_enc > 0) FOV_CHECK_ARG(io->enc.xh && io->enc.gates && io->enc.c && g->dz_enc, "encoder saved/dz buffers");
  if (cfg->T_dec > 0) FOV_CHECK_ARG(io->dec.xh && io->dec.gates && io->dec.c && io->dec.hseq && g->dz_dec, "decoder saved/dz buffers");
  if (has_head) FOV_CHECK_ARG(g->dy && g->y && g->dpre, "head gradient buffers");
  LstmKParams P{};
  P.cfg = *cfg; P.w = *w; P.io = *io; P.g = *g;
  if (P.cfg.T_dec == 0) P.cfg.out_dim = 0;
  cudaStream_t st = (cudaStream_t)stream;
  int spt = cfg->B <= 48 * fov_num_sms() ? 4 : 8;
  if (cfg->B <= 8 * fov_num_sms() && !g_lstm_no_small) spt = 2;      // small batches: see the forward dispatch
  if (cfg->B <= 4 * fov_num_sms() && !g_lstm_no_small) spt = 1;
  const int bt = 4 * spt, grid = (cfg->B + bt - 1) / bt;
  const size_t smem = bwd_smem_bytes(spt);
  const bool hsb = cfg->rec_act == FOV_REC_HARD_SIGMOID;
  // tensor-core BPTT (lstm_seq2seq_tc.cu): 128-sequence tiles, dh_rec = dZ x U^T on tcgen05, saved tensors by 256-bit
  // loads (crossover with the fp32 kernel near 6 k sequences, see the forward dispatch above)
  if (cfg->math != FOV_MATH_FP32 && g_lstm_tc_mode >= 0 && g_lstm_bptt_tc && lstm_tc_bwd_supported(cfg) &&
      (g_lstm_tc_mode > 0 || cfg->B >= kTcTrainMinB))
    rc = lstm_tc_bwd(&P.cfg, w, io, g, st);
  else if (spt == 1)
    rc = hsb ? launch(lstm_seq2seq_bwd_kernel<1, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
             : launch(lstm_seq2seq_bwd_kernel<1, FOV_REC_SIGMOID>, P, grid, smem, st);
  else if (spt == 2)
    rc = hsb ? launch(lstm_seq2seq_bwd_kernel<2, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
             : launch(lstm_seq2seq_bwd_kernel<2, FOV_REC_SIGMOID>, P, grid, smem, st);
  else if (spt == 4)
    rc = hsb ? launch(lstm_seq2seq_bwd_kernel<4, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
             : launch(lstm_seq2seq_bwd_kernel<4, FOV_REC_SIGMOID>, P, grid, smem, st);
  else
    rc = hsb ? launch(lstm_seq2seq_bwd_kernel<8, FOV_REC_HARD_SIGMOID>, P, grid, smem, st)
             : launch(lstm_seq2seq_bwd_kernel<8, FOV_REC_SIGMOID>, P, grid, smem, st);
  if (rc) return rc;

  // time-batched weight gradients: [dU; dW] = [h_{t-1} | x_t]^T dZ over all (b,t) rows
  auto wgrad = [&](const float* A, long long lda, int K, const float* dZ, int N, long long rows,
                   float* gw, float* gb, bool tc) -> int {
    if (!gw && !gb) return FOV_OK;
    fov_conv_cfg c{};
    c.N = (int)rows; c.H = 1; c.W = 1; c.Cin = K; c.Cout = N; c.kh = 1; c.kw = 1; c.dil_h = 1; c.dil_w = 1;
    c.x_img_stride = lda; c.x_pix_stride = (int)lda; c.y_img_stride = N; c.y_pix_stride = N;
    if (tc) return fov_conv2d_bwd_weight_tc(&c, A, dZ, gw, gb, cfg->math, stream);
    return fov_conv2d_bwd_weight(&c, A, dZ, gw, gb, stream);
  };
  // One LSTM: with a workspace and a tensor-core math mode, ONE tcgen05 launch computes the whole padded product
  // [h | x | 0]^T dZ (aligned rows: the fast path of wgrad_tc.cu) and the bias gradient, and a small kernel adds its
  // row blocks to the two weight-gradient tensors; otherwise two SIMT products + a column sum.
  auto lstm_wgrads = [&](const fov_lstm_saved& sv, int in_dim, int T, const float* dz, float* g_kernel,
                         float* g_recurrent, float* g_bias, float* wsp) -> int {
    const long long rows = (long long)cfg->B * T;
    const int XS = fov_lstm_xh_stride(kH, in_dim);
    if (wsp && cfg->math != FOV_MATH_FP32 && g_lstm_wgrad_tc && g_kernel && g_recurrent) {
      cudaError_t e = cudaMemsetAsync(wsp, 0, sizeof(float) * (size_t)XS * kG, st);
      if (e != cudaSuccess) { fov_set_error("fov_lstm_seq2seq_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); return FOV_ERR_CUDA; }
      int r = wgrad(sv.xh, XS, XS, dz, kG, rows, wsp, g_bias, true);
      if (r) return r;
      const int n = (kH + in_dim) * kG;
      lstm_wgrad_scatter_kernel<<<(n / 4 + 255) / 256, 256, 0, st>>>(wsp, g_recurrent, g_kernel, in_dim);
      FOV_CUDA_LAUNCH_CHECK();
      return FOV_OK;
    }
    int r = wgrad(sv.xh, XS, kH, dz, kG, rows, g_recurrent, g_bias, false);
    if (r) return r;
    return wgrad(sv.xh + kH, XS, in_dim, dz, kG, rows, g_kernel, nullptr, false);
  };
  // the weight gradients feed only the optimiser: on request they leave the backward chain for a side stream
  if (g->wgrad_stream && (cudaStream_t)g->wgrad_stream != st) {
    if (fov_fork_stream(st, (cudaStream_t)g->wgrad_stream)) {
      fov_set_error("fov_lstm_seq2seq_bwd: could not fork the weight-gradient stream");
      return FOV_ERR_CUDA;
    }
    st = (cudaStream_t)g->wgrad_stream;
    stream = g->wgrad_stream;
  }
  float* wsp = g->ws;
  if (cfg->T_enc > 0) {
    if ((rc = lstm_wgrads(io->enc, cfg->in_enc, cfg->T_enc, g->dz_enc, g->g_enc_kernel, g->g_enc_recurrent,
                          g->g_enc_bias, wsp)))
      return rc;
    if (wsp) wsp += (size_t)fov_lstm_xh_stride(kH, cfg->in_enc) * kG;
  }
  if (cfg->T_dec > 0) {
    if ((rc = lstm_wgrads(io->dec, cfg->in_dec, cfg->T_dec, g->dz_dec, g->g_dec_kernel, g->g_dec_recurrent,
                          g->g_dec_bias, wsp)))
      return rc;
    if (has_head) {
      const long long rows = (long long)cfg->B * cfg->T_dec;
      if ((rc = wgrad(io->dec.hseq, kH, kH, g->dpre, cfg->out_dim, rows, g->g_head_kernel, g->g_head_bias, false))) return rc;
    }
  }
  return FOV_OK;
}
