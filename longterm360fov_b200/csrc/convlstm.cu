// ConvLSTM2D layer over a whole sequence, forward and BPTT, built from the
// implicit-GEMM convolution kernels and the fused gate kernels.
//
// forward:  Zx = conv(x_all_t, K) + b      (one time-batched launch over B*T images)
//           for t: Z_t += conv(h_{t-1}, R);  gates/cell update in place -> h_t, c_t
// backward: for t reversed: dZ_t from (dh_t, dc_t, saved gates); dh_{t-1} = conv^T(dZ_t, R)
//           then time-batched dx = conv^T(dZ, K), gK += x^T dZ, gR += h_{t-1}^T dZ, gb += sum dZ
//
// Replaces keras ConvLSTM2D at mycode/others_LSTM_span_whole.py:88-100 and
// mycode/convlstm_seq2seq.py:100-126,146-165,213-218.
#include "fov_common.cuh"
#include "fov_internal.h"

namespace {

int check(const fov_convlstm_cfg* c) {
  FOV_CHECK_ARG(c != nullptr, "cfg is NULL");
  FOV_CHECK_ARG(c->B > 0 && c->T > 0 && c->H > 0 && c->W > 0 && c->Cin > 0 && c->F > 0, "bad shape");
  FOV_CHECK_ARG(c->kh > 0 && c->kw > 0 && c->dil_h > 0 && c->dil_w > 0, "bad kernel");
  return FOV_OK;
}

struct Geo {
  int HW;
  long long z_b, z_t;   // gates buffer strides (B,T,HW,4F)
  long long c_b, c_t;   // cseq strides (B,T,HW,F)
  long long hw_f;       // HW*F
};
Geo geo(const fov_convlstm_cfg* c) {
  Geo g;
  g.HW = c->H * c->W;
  g.z_t = (long long)g.HW * 4 * c->F;
  g.z_b = g.z_t * c->T;
  g.c_t = (long long)g.HW * c->F;
  g.c_b = g.c_t * c->T;
  g.hw_f = g.c_t;
  return g;
}

fov_conv_cfg input_conv_cfg(const fov_convlstm_cfg* c) {
  fov_conv_cfg k{};
  k.H = c->H; k.W = c->W; k.Cin = c->Cin; k.Cout = 4 * c->F;
  k.kh = c->kh; k.kw = c->kw; k.dil_h = c->dil_h; k.dil_w = c->dil_w;
  k.pad_h = ((c->kh - 1) * c->dil_h) / 2; k.pad_w = ((c->kw - 1) * c->dil_w) / 2;
  k.x_pix_stride = c->x_pix_stride; k.y_pix_stride = 4 * c->F;
  k.act = FOV_ACT_LINEAR; k.beta = 0.0f;
  return k;
}
fov_conv_cfg rec_conv_cfg(const fov_convlstm_cfg* c) {
  fov_conv_cfg k{};
  k.N = c->B; k.H = c->H; k.W = c->W; k.Cin = c->F; k.Cout = 4 * c->F;
  k.kh = c->kh; k.kw = c->kw; k.dil_h = 1; k.dil_w = 1;
  k.pad_h = (c->kh - 1) / 2; k.pad_w = (c->kw - 1) / 2;
  k.y_pix_stride = 4 * c->F;
  k.act = FOV_ACT_LINEAR; k.beta = 1.0f;
  return k;
}

// fused tensor-core step: z = [x_t taps | h_{t-1} taps] x [K ; R], gates + cell update in the epilogue
TcConv step_conv(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const Geo& g) {
  TcConv k{};
  k.nseg = 2;
  k.N_img = c->B; k.T_inner = 1; k.H = c->H; k.W = c->W;
  k.Cout = 4 * c->F; k.math = c->math; k.epi = TC_EPI_LSTM;
  TcSeg& a = k.seg[0];
  a.img_outer = c->x_b_stride; a.img_inner = 0; a.pix_stride = c->x_pix_stride; a.Cin = c->Cin;
  a.kh = c->kh; a.kw = c->kw; a.dil_h = c->dil_h; a.dil_w = c->dil_w;
  a.pad_h = ((c->kh - 1) * c->dil_h) / 2; a.pad_w = ((c->kw - 1) * c->dil_w) / 2;
  a.w = io->kernel; a.w_mode = 0;
  TcSeg& h = k.seg[1];
  h.Cin = c->F; h.kh = c->kh; h.kw = c->kw; h.dil_h = 1; h.dil_w = 1;
  h.pad_h = (c->kh - 1) / 2; h.pad_w = (c->kw - 1) / 2;
  h.w = io->recurrent; h.w_mode = 0;
  k.bias = io->bias; k.rec_act = c->rec_act;
  k.c_outer = g.c_b; k.h_outer = c->h_b_stride; k.h_pix_stride = c->h_pix_stride; k.g_outer = g.z_b;
  return k;
}

// dh_{t-1} = conv^T(dZ_t, R) on tensor cores: "x" = dZ_t (B,HW,4F), output dense (B,HW,F)
TcConv rec_bwd_conv(const fov_convlstm_cfg* c, const float* recurrent, const Geo& g) {
  TcConv k{};
  k.nseg = 1; k.N_img = c->B; k.T_inner = 1; k.H = c->H; k.W = c->W;
  k.Cout = c->F; k.math = c->math; k.epi = TC_EPI_CONV;
  TcSeg& a = k.seg[0];
  a.img_outer = g.z_b; a.img_inner = 0; a.pix_stride = 4 * c->F; a.Cin = 4 * c->F;
  a.kh = c->kh; a.kw = c->kw; a.dil_h = 1; a.dil_w = 1;
  a.pad_h = (c->kh - 1) - (c->kh - 1) / 2; a.pad_w = (c->kw - 1) - (c->kw - 1) / 2;
  a.w = recurrent; a.w_mode = 1;
  k.y_outer = g.hw_f; k.y_inner = 0; k.y_pix_stride = c->F; k.act = FOV_ACT_LINEAR; k.beta = 0.0f;
  return k;
}
// dx = conv^T(dZ, K) for every (b,t) image in one launch
TcConv in_bwd_conv(const fov_convlstm_cfg* c, const float* kernel, const Geo& g) {
  TcConv k{};
  k.nseg = 1; k.N_img = c->B * c->T; k.T_inner = c->T; k.H = c->H; k.W = c->W;
  k.Cout = c->Cin; k.math = c->math; k.epi = TC_EPI_CONV;
  TcSeg& a = k.seg[0];
  a.img_outer = g.z_b; a.img_inner = g.z_t; a.pix_stride = 4 * c->F; a.Cin = 4 * c->F;
  a.kh = c->kh; a.kw = c->kw; a.dil_h = c->dil_h; a.dil_w = c->dil_w;
  a.pad_h = (c->kh - 1) * c->dil_h - ((c->kh - 1) * c->dil_h) / 2;
  a.pad_w = (c->kw - 1) * c->dil_w - ((c->kw - 1) * c->dil_w) / 2;
  a.w = kernel; a.w_mode = 1;
  k.y_outer = c->x_b_stride; k.y_inner = c->x_t_stride; k.y_pix_stride = c->x_pix_stride;
  k.act = FOV_ACT_LINEAR;
  return k;
}

// The fused tensor-core step covers the filter counts of the reference models (8/16/32/64) with
// <= 64 input channels and 16-byte aligned hidden-state rows; other shapes run the fp32 CUDA-core kernels.
bool tc_step_ok(const fov_convlstm_cfg* c) {
  if (c->math == 0) return false;
  const int F = c->F;
  if (!(F == 8 || F == 16 || F == 32 || F == 64)) return false;
  if (!(c->h_pix_stride % 4 == 0 && c->h_b_stride % 4 == 0 && c->h_t_stride % 4 == 0)) return false;
  fov_convlstm_io io{};
  const Geo g = geo(c);
  const bool ok = tc_conv_supported(step_conv(c, &io, g)) && tc_conv_supported(rec_bwd_conv(c, nullptr, g)) &&
                  tc_conv_supported(in_bwd_conv(c, nullptr, g));
  fov_set_error("");
  return ok;
}

}  // namespace

extern "C" size_t fov_convlstm_fwd_ws_bytes(const fov_convlstm_cfg* cfg) {
  if (!cfg || check(cfg) || !tc_step_ok(cfg)) return 0;
  fov_convlstm_io io{};
  return tc_conv_ws_bytes(step_conv(cfg, &io, geo(cfg)));
}

extern "C" int fov_convlstm_fwd_persistent(const fov_convlstm_cfg* cfg) {
  if (!cfg || check(cfg) || !tc_step_ok(cfg)) { fov_set_error(""); return 0; }
  fov_convlstm_io io{};
  return tc_convlstm_seq_supported(cfg, step_conv(cfg, &io, geo(cfg))) ? 1 : 0;
}

extern "C" int fov_convlstm_wave_groups(const fov_convlstm_cfg* cfg, int backward, int with_dx) {
  if (!cfg || check(cfg) || !tc_step_ok(cfg)) { fov_set_error(""); return 0; }
  const Geo g = geo(cfg);
  fov_convlstm_io io{};
  if (!backward) return tc_convlstm_seq_wave_groups(cfg, step_conv(cfg, &io, g));
  TcConv rT = rec_bwd_conv(cfg, nullptr, g);
  TcConv kT = in_bwd_conv(cfg, nullptr, g);
  return tc_convlstm_seq_bwd_wave_groups(cfg, rT, with_dx ? &kT : nullptr);
}

static int convlstm_fwd_tc(const fov_convlstm_cfg* cfg, const fov_convlstm_io* io, cudaStream_t st) {
  FOV_CHECK_ARG(io->ws != nullptr, "math != 0 needs io->ws (fov_convlstm_fwd_ws_bytes)");
  const Geo g = geo(cfg);
  const int F = cfg->F;
  TcConv k = step_conv(cfg, io, g);
  k.ws = io->ws;
  // whole images per MMA tile: one persistent launch runs every timestep (convlstm_seq_tc.cu)
  if (tc_convlstm_seq_supported(cfg, k)) return tc_convlstm_seq_fwd(cfg, io, k, st);
  FOV_CHECK_ARG(!io->wave_wait && !io->wave_set, "wavefront flags need the persistent kernel (fov_convlstm_wave_groups)");
  for (int t = 0; t < cfg->T; ++t) {
    k.prepacked = t > 0 || cfg->ws_prepacked;
    k.seg[0].x = io->x + t * cfg->x_t_stride;
    if (t == 0) {
      k.seg[1].x = io->h0; k.seg[1].img_outer = g.hw_f; k.seg[1].pix_stride = F;
      k.c_prev = io->c0; k.cp_outer = g.hw_f;
    } else {
      k.seg[1].x = io->hseq + (t - 1) * cfg->h_t_stride; k.seg[1].img_outer = cfg->h_b_stride;
      k.seg[1].pix_stride = cfg->h_pix_stride;
      k.c_prev = io->cseq + (t - 1) * g.c_t; k.cp_outer = g.c_b;
    }
    k.c_out = io->cseq + t * g.c_t;
    k.h_out = io->hseq + t * cfg->h_t_stride;
    k.gates_out = cfg->training ? io->gates + t * g.z_t : nullptr;
    k.hT = (t == cfg->T - 1) ? io->hT : nullptr;
    k.cT = (t == cfg->T - 1) ? io->cT : nullptr;
    int rc = tc_conv_run(k, st);
    if (rc) return rc;
  }
  return FOV_OK;
}

extern "C" int fov_convlstm_fwd(const fov_convlstm_cfg* cfg, const fov_convlstm_io* io, void* stream) {
  int rc = check(cfg);
  if (rc) return rc;
  FOV_CHECK_ARG(io && io->x && io->kernel && io->recurrent && io->bias && io->hseq && io->gates && io->cseq,
                "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_step_ok(cfg)) return convlstm_fwd_tc(cfg, io, st);
  FOV_CHECK_ARG(!io->wave_wait && !io->wave_set, "wavefront flags need the persistent kernel (fov_convlstm_wave_groups)");
  const Geo g = geo(cfg);
  const int F = cfg->F;

  // ---- input projection for every timestep ----
  // (input dropout is expressed by the caller as a widened input: fov_dropout_expand / fov_gate_kernel_expand)
  {
    fov_conv_cfg k = input_conv_cfg(cfg);
    if (cfg->x_b_stride == (long long)cfg->T * cfg->x_t_stride || cfg->T == 1) {
      k.N = cfg->B * cfg->T; k.x_img_stride = cfg->T == 1 ? cfg->x_b_stride : cfg->x_t_stride;
      k.y_img_stride = g.z_t;
      if ((rc = fov_conv2d_fwd(&k, io->x, io->kernel, io->bias, io->gates, stream))) return rc;
    } else {
      k.N = cfg->B; k.x_img_stride = cfg->x_b_stride; k.y_img_stride = g.z_b;
      for (int t = 0; t < cfg->T; ++t)
        if ((rc = fov_conv2d_fwd(&k, io->x + t * cfg->x_t_stride, io->kernel, io->bias,
                                 io->gates + t * g.z_t, stream))) return rc;
    }
  }

  // ---- recurrence ----
  fov_conv_cfg r = rec_conv_cfg(cfg);
  for (int t = 0; t < cfg->T; ++t) {
    float* zt = io->gates + t * g.z_t;
    const float* hprev = nullptr;
    if (t == 0) {
      if (io->h0) { hprev = io->h0; r.x_img_stride = g.hw_f; r.x_pix_stride = F; }
    } else {
      hprev = io->hseq + (t - 1) * cfg->h_t_stride; r.x_img_stride = cfg->h_b_stride; r.x_pix_stride = cfg->h_pix_stride;
    }
    if (hprev) {
      r.y_img_stride = g.z_b;
      if ((rc = fov_conv2d_fwd(&r, hprev, io->recurrent, nullptr, zt, stream))) return rc;
    }
    GatesFwdArgs a{};
    a.npix = (long long)cfg->B * g.HW; a.HW = g.HW; a.F = F; a.rec = cfg->rec_act;
    a.z = zt; a.z_img = g.z_b;
    if (t == 0) { a.c_prev = io->c0; a.cp_img = g.hw_f; }
    else { a.c_prev = io->cseq + (t - 1) * g.c_t; a.cp_img = g.c_b; }
    a.c_out = io->cseq + t * g.c_t; a.c_img = g.c_b;
    a.h_out = io->hseq + t * cfg->h_t_stride; a.h_img = cfg->h_b_stride; a.h_pix = cfg->h_pix_stride;
    a.hT = (t == cfg->T - 1) ? io->hT : nullptr;
    a.cT = (t == cfg->T - 1) ? io->cT : nullptr;
    if ((rc = fov_launch_gates_fwd(a, st))) return rc;
  }
  return FOV_OK;
}

extern "C" size_t fov_convlstm_bwd_ws_floats(const fov_convlstm_cfg* c) {
  if (!c) return 0;
  const size_t bhwf = (size_t)c->B * c->H * c->W * c->F;
  const size_t taps = (size_t)c->kh * c->kw;
  // dh_rec + dc (ping-pong x2) + flipped recurrent + flipped kernel
  size_t n = 4 * bhwf + taps * c->F * 4 * c->F + taps * c->Cin * 4 * c->F + 64;
  if (tc_step_ok(c)) {   // packed bf16 weights of the two tensor-core backward-data convolutions
    const Geo g = geo(c);
    n += (tc_conv_ws_bytes(rec_bwd_conv(c, nullptr, g)) + tc_conv_ws_bytes(in_bwd_conv(c, nullptr, g))) / 4 + 128;
  }
  return n;
}

extern "C" int fov_convlstm_bwd(const fov_convlstm_cfg* cfg, const fov_convlstm_io* io,
                                const fov_convlstm_grads* gr, void* stream) {
  int rc = check(cfg);
  if (rc) return rc;
  FOV_CHECK_ARG(io && gr && io->x && io->kernel && io->recurrent && io->hseq && io->gates && io->cseq && gr->ws,
                "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const Geo g = geo(cfg);
  const int F = cfg->F;
  const size_t bhwf = (size_t)cfg->B * g.HW * F;
  float* dh_rec = gr->ws;
  float* dc_a = gr->ws + bhwf;
  float* dc_b = gr->ws + 2 * bhwf;
  float* Rt = gr->ws + 4 * bhwf;
  float* Kt = Rt + (size_t)cfg->kh * cfg->kw * F * 4 * F;

  fov_conv_cfg r = rec_conv_cfg(cfg);
  r.beta = 0.0f;
  // dh_rec is produced as the "x" side of the recurrent conv: dense (B,HW,F)
  r.x_img_stride = g.hw_f; r.x_pix_stride = F; r.y_img_stride = g.z_b;
  const bool tcm = tc_step_ok(cfg);
  TcConv rT = rec_bwd_conv(cfg, io->recurrent, g);
  TcConv kT = in_bwd_conv(cfg, io->kernel, g);
  bool seq = false, dx_fused = false;
  if (tcm) {
    float* pk = Kt + (size_t)cfg->kh * cfg->kw * cfg->Cin * 4 * F + 64;
    rT.ws = pk;
    kT.ws = pk + tc_conv_ws_bytes(rT) / 4 + 64;
    // whole images per MMA tile and no gradient w.r.t. an initial hidden state: the persistent kernel walks the
    // whole reverse time loop in one launch (convlstm_seq_bwd_tc.cu)
    if (!(io->h0 && gr->dh0)) {
      if (gr->dx && tc_convlstm_seq_bwd_supported(cfg, rT, &kT)) { seq = true; dx_fused = true; }
      else if (tc_convlstm_seq_bwd_supported(cfg, rT, nullptr)) seq = true;
    }
    bool wg_fused = false;
    FOV_CHECK_ARG(seq || (!gr->wave_wait && !gr->wave_set), "wavefront flags need the persistent BPTT kernel");
    FOV_CHECK_ARG(!gr->wave_set || dx_fused, "wavefront: wave_set needs the fused input gradient");
    if (seq) {
      wg_fused = tc_convlstm_seq_bwd_fuses_wgrad(cfg, io, gr, rT, dx_fused ? &kT : nullptr);
      if ((rc = tc_convlstm_seq_bwd(cfg, io, gr, rT, dx_fused ? &kT : nullptr, st))) return rc;
      if (wg_fused) return FOV_OK;               // dZ never left the SM; gK / gR / gb are done
    } else {
      if ((rc = tc_conv_pack(rT, st))) return rc;
      rT.prepacked = 1;
    }
  } else {
    FOV_CHECK_ARG(!gr->wave_wait && !gr->wave_set, "wavefront flags need the persistent BPTT kernel");
    if ((rc = fov_conv_flip_weights(&r, io->recurrent, Rt, st))) return rc;
  }

  const float* dc_in = gr->dcT;
  const float* dh_in = gr->dhT;
  for (int t = cfg->T - 1; t >= 0 && !seq; --t) {
    float* zt = io->gates + t * g.z_t;
    GatesBwdArgs a{};
    a.npix = (long long)cfg->B * g.HW; a.HW = g.HW; a.F = F; a.rec = cfg->rec_act;
    a.gates = zt; a.z_img = g.z_b;
    a.c_t = io->cseq + t * g.c_t; a.c_img = g.c_b;
    if (t == 0) { a.c_prev = io->c0; a.cp_img = g.hw_f; }
    else { a.c_prev = io->cseq + (t - 1) * g.c_t; a.cp_img = g.c_b; }
    if (gr->dhseq) { a.dh_ext = gr->dhseq + t * cfg->h_t_stride; a.dhe_img = cfg->h_b_stride; a.dhe_pix = cfg->h_pix_stride; }
    a.dh_rec = dh_in; a.dc_in = dc_in;
    float* dc_out = (t == 0 && gr->dc0) ? gr->dc0 : ((dc_in == dc_a) ? dc_b : dc_a);
    a.dc_out = dc_out;
    if ((rc = fov_launch_gates_bwd(a, st))) return rc;
    dc_in = dc_out;
    if (t > 0 || (io->h0 && gr->dh0)) {
      float* dst = (t == 0) ? gr->dh0 : dh_rec;
      if (tcm) {
        rT.seg[0].x = zt; rT.y = dst;
        if ((rc = tc_conv_run(rT, st))) return rc;
      } else {
        if ((rc = fov_conv_bwd_data_preflipped(&r, zt, Rt, dst, st))) return rc;
      }
      dh_in = dst;
    }
  }

  if (tcm) {
    // ---- tensor-core path: dx, gK (+ bias gradient) and gR, each one launch over every (b,t) image ----
    if (gr->dx && !dx_fused) {
      kT.seg[0].x = io->gates; kT.y = gr->dx; kT.beta = gr->dx_accumulate ? 1.0f : 0.0f;
      if ((rc = tc_conv_run(kT, st))) return rc;
    }
    // the weight gradients feed only the optimiser: on request they leave the backward chain for a side stream
    if (gr->wgrad_stream && (cudaStream_t)gr->wgrad_stream != st) {
      if (fov_fork_stream(st, (cudaStream_t)gr->wgrad_stream)) {
        fov_set_error("fov_convlstm_bwd: could not fork the weight-gradient stream");
        return FOV_ERR_CUDA;
      }
      st = (cudaStream_t)gr->wgrad_stream;
    }
    // fused weight gradient (gK, gR, gb in one launch, no gather) when the taps fit the TMEM columns
    if (gr->g_kernel && gr->g_recurrent) {
      TcWgradRows wr{};
      wr.nseg = 2; wr.dz = io->gates; wr.B = cfg->B; wr.T = cfg->T; wr.H = cfg->H; wr.W = cfg->W;
      wr.Cout = 4 * F; wr.math = cfg->math; wr.gbias = gr->g_bias;
      TcWgradRowsSeg& sx = wr.seg[0];
      sx.x = io->x; sx.b_stride = cfg->x_b_stride; sx.t_stride = cfg->x_t_stride; sx.pix_stride = cfg->x_pix_stride;
      sx.t_shift = 0; sx.Cin = cfg->Cin; sx.kh = cfg->kh; sx.kw = cfg->kw; sx.dil_h = cfg->dil_h; sx.dil_w = cfg->dil_w;
      sx.pad_h = ((cfg->kh - 1) * cfg->dil_h) / 2; sx.pad_w = ((cfg->kw - 1) * cfg->dil_w) / 2;
      sx.gw = gr->g_kernel;
      TcWgradRowsSeg& sh = wr.seg[1];
      sh.x = io->hseq; sh.b_stride = cfg->h_b_stride; sh.t_stride = cfg->h_t_stride; sh.pix_stride = cfg->h_pix_stride;
      sh.x0 = io->h0; sh.x0_b_stride = g.hw_f; sh.x0_pix_stride = F;
      sh.t_shift = -1; sh.Cin = F; sh.kh = cfg->kh; sh.kw = cfg->kw; sh.dil_h = 1; sh.dil_w = 1;
      sh.pad_h = (cfg->kh - 1) / 2; sh.pad_w = (cfg->kw - 1) / 2;
      sh.gw = gr->g_recurrent;
      if (tc_wgrad_rows_supported(wr)) return tc_wgrad_rows_run(wr, st);
    }
    TcWgrad w{};
    w.N_img = cfg->B * cfg->T; w.T_inner = cfg->T; w.H = cfg->H; w.W = cfg->W; w.math = cfg->math;
    w.x = io->x; w.x_outer = cfg->x_b_stride; w.x_inner = cfg->x_t_stride; w.x_pix_stride = cfg->x_pix_stride;
    w.Cin = cfg->Cin; w.kh = cfg->kh; w.kw = cfg->kw; w.dil_h = cfg->dil_h; w.dil_w = cfg->dil_w;
    w.pad_h = ((cfg->kh - 1) * cfg->dil_h) / 2; w.pad_w = ((cfg->kw - 1) * cfg->dil_w) / 2;
    w.dy = io->gates; w.dy_outer = g.z_b; w.dy_inner = g.z_t; w.dy_pix_stride = 4 * F; w.Cout = 4 * F;
    w.gw = gr->g_kernel; w.gbias = gr->g_bias;
    if (w.gw || w.gbias) {
      if (!w.gw) {
        fov_set_error("fov_convlstm_bwd: g_kernel is required on the tensor-core path");
        return FOV_ERR_ARG;
      }
      if ((rc = tc_wgrad_run(w, st))) return rc;
    }
    if (gr->g_recurrent) {
      w.dil_h = 1; w.dil_w = 1; w.pad_h = (cfg->kh - 1) / 2; w.pad_w = (cfg->kw - 1) / 2;
      w.Cin = F; w.gw = gr->g_recurrent; w.gbias = nullptr;
      if (cfg->T > 1) {   // pairs (h_{t-1}, dZ_t), t = 1..T-1
        w.N_img = cfg->B * (cfg->T - 1); w.T_inner = cfg->T - 1;
        w.x = io->hseq; w.x_outer = cfg->h_b_stride; w.x_inner = cfg->h_t_stride; w.x_pix_stride = cfg->h_pix_stride;
        w.dy = io->gates + g.z_t;
        if ((rc = tc_wgrad_run(w, st))) return rc;
      }
      if (io->h0) {       // pair (h0, dZ_0)
        w.N_img = cfg->B; w.T_inner = 1;
        w.x = io->h0; w.x_outer = g.hw_f; w.x_inner = 0; w.x_pix_stride = F;
        w.dy = io->gates;
        if ((rc = tc_wgrad_run(w, st))) return rc;
      }
    }
    return FOV_OK;
  }
  // ---- time-batched input-side gradients ----
  fov_conv_cfg k = input_conv_cfg(cfg);
  k.beta = gr->dx_accumulate ? 1.0f : 0.0f;
  k.N = cfg->B;
  const bool batched = (cfg->x_b_stride == (long long)cfg->T * cfg->x_t_stride) || cfg->T == 1;
  if (gr->dx) {
    if ((rc = fov_conv_flip_weights(&k, io->kernel, Kt, st))) return rc;
  }
  if (batched) {
    k.N = cfg->B * cfg->T; k.x_img_stride = cfg->T == 1 ? cfg->x_b_stride : cfg->x_t_stride; k.y_img_stride = g.z_t;
    if (gr->dx && (rc = fov_conv_bwd_data_preflipped(&k, io->gates, Kt, gr->dx, st))) return rc;
    if ((rc = fov_conv2d_bwd_weight(&k, io->x, io->gates, gr->g_kernel, gr->g_bias, stream))) return rc;
  } else {
    k.N = cfg->B; k.x_img_stride = cfg->x_b_stride; k.y_img_stride = g.z_b;
    for (int t = 0; t < cfg->T; ++t) {
      if (gr->dx && (rc = fov_conv_bwd_data_preflipped(&k, io->gates + t * g.z_t, Kt, gr->dx + t * cfg->x_t_stride, st))) return rc;
      if ((rc = fov_conv2d_bwd_weight(&k, io->x + t * cfg->x_t_stride, io->gates + t * g.z_t, gr->g_kernel,
                                      gr->g_bias, stream))) return rc;
    }
  }
  // ---- recurrent weight gradient: pairs (h_{t-1}, dZ_t) ----
  if (gr->g_recurrent) {
    fov_conv_cfg w = rec_conv_cfg(cfg);
    w.y_img_stride = g.z_b;
    for (int t = 0; t < cfg->T; ++t) {
      const float* hprev;
      if (t == 0) {
        if (!io->h0) continue;
        hprev = io->h0; w.x_img_stride = g.hw_f; w.x_pix_stride = F;
      } else {
        hprev = io->hseq + (t - 1) * cfg->h_t_stride; w.x_img_stride = cfg->h_b_stride; w.x_pix_stride = cfg->h_pix_stride;
      }
      if ((rc = fov_conv2d_bwd_weight(&w, hprev, io->gates + t * g.z_t, gr->g_recurrent, nullptr, stream))) return rc;
    }
  }
  return FOV_OK;
}
