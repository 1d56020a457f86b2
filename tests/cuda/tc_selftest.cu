// Stand-alone bring-up / regression harness for the tensor-core kernels of libfov360.so.
// Compares every tcgen05 entry point with the fp32 CUDA-core kernel of the same C ABI on the
// same random inputs (the fp32 kernels are the ones pinned to the oracle by tests/), prints the
// max-abs errors and a few timings.  No torch, no Python: starts in milliseconds on the GPU box.
//
//   build:  longterm360fov_b200/csrc/build.sh   (-> csrc/build/tc_selftest)
//   run:    longterm360fov_b200/csrc/build/tc_selftest [quick]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/fov360.h"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)
#define FK(x)                                                                  \
  do {                                                                         \
    int r_ = (x);                                                              \
    if (r_ != 0) {                                                             \
      printf("fov error %d (%s) at %s:%d\n", r_, fov_last_error(), __FILE__, __LINE__); \
      exit(3);                                                                 \
    }                                                                          \
  } while (0)

static unsigned long long g_seed = 0x9E3779B97F4A7C15ull;
static float frand() {   // uniform (-1,1)
  g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17;
  return (float)((g_seed >> 11) * (1.0 / 9007199254740992.0)) * 2.0f - 1.0f;
}
static float* dev_rand(size_t n, float scale) {
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = frand() * scale;
  float* d;
  CK(cudaMalloc(&d, n * sizeof(float) + 256));
  CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}
static float* dev_zero(size_t n) {
  float* d;
  CK(cudaMalloc(&d, n * sizeof(float) + 256));
  CK(cudaMemset(d, 0, n * sizeof(float) + 256));
  return d;
}
static float* dev_copy(const float* src, size_t n) {
  float* d;
  CK(cudaMalloc(&d, n * sizeof(float) + 256));
  CK(cudaMemcpy(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice));
  return d;
}
struct Err { double max_abs, max_ref; };
static Err compare(const float* a, const float* ref, size_t n) {
  std::vector<float> ha(n), hr(n);
  CK(cudaMemcpy(ha.data(), a, n * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), ref, n * sizeof(float), cudaMemcpyDeviceToHost));
  Err e{0.0, 0.0};
  for (size_t i = 0; i < n; ++i) {
    double d = fabs((double)ha[i] - (double)hr[i]);
    if (!(d == d)) d = 1e30;
    if (d > e.max_abs) e.max_abs = d;
    if (fabs((double)hr[i]) > e.max_ref) e.max_ref = fabs((double)hr[i]);
  }
  return e;
}
static int g_fail = 0;
static void report(const char* what, int math, Err e, double tol_rel) {
  const double rel = e.max_abs / (e.max_ref > 0 ? e.max_ref : 1.0);
  const bool ok = rel <= tol_rel;
  if (!ok) ++g_fail;
  printf("  %-34s math=%d max_abs=%.3e max_ref=%.3e rel=%.3e %s\n", what, math, e.max_abs, e.max_ref, rel,
         ok ? "ok" : "FAIL");
}
static const double kTol[4] = {0, 2e-2, 2e-4, 2e-4};

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a)); }
  float stop_ms() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

static fov_conv_cfg mkcfg(int N, int H, int W, int Cin, int Cout, int kh, int kw, int dh, int dw, int act, float beta) {
  fov_conv_cfg c{};
  c.N = N; c.H = H; c.W = W; c.Cin = Cin; c.Cout = Cout; c.kh = kh; c.kw = kw; c.dil_h = dh; c.dil_w = dw;
  c.pad_h = ((kh - 1) * dh) / 2; c.pad_w = ((kw - 1) * dw) / 2;
  c.x_img_stride = (long long)H * W * Cin; c.x_pix_stride = Cin;
  c.y_img_stride = (long long)H * W * Cout; c.y_pix_stride = Cout;
  c.act = act; c.beta = beta;
  return c;
}

static void test_conv(const char* name, fov_conv_cfg c, bool time_it) {
  printf("[conv] %s N=%d HxW=%dx%d Cin=%d Cout=%d k=%dx%d dil=%d,%d act=%d beta=%g\n", name, c.N, c.H, c.W, c.Cin,
         c.Cout, c.kh, c.kw, c.dil_h, c.dil_w, c.act, c.beta);
  const size_t nx = (size_t)c.N * c.H * c.W * c.Cin, ny = (size_t)c.N * c.H * c.W * c.Cout;
  const size_t nw = (size_t)c.kh * c.kw * c.Cin * c.Cout;
  float* x = dev_rand(nx, 1.0f);
  float* w = dev_rand(nw, 1.0f / sqrtf((float)(c.kh * c.kw * c.Cin)));
  float* b = dev_rand(c.Cout, 0.5f);
  float* y0 = dev_rand(ny, 1.0f);
  float* yref = dev_copy(y0, ny);
  FK(fov_conv2d_fwd(&c, x, w, b, yref, nullptr));
  // backward-data reference: dy random, dx0 random (beta)
  float* dy = dev_rand(ny, 1.0f);
  float* dx0 = dev_rand(nx, 1.0f);
  float* dxref = dev_copy(dx0, nx);
  float* wsf = dev_zero(nw);
  FK(fov_conv2d_bwd_data(&c, dy, w, dxref, wsf, nullptr));
  // weight / bias gradient reference (accumulating into random initial values)
  float* gw0 = dev_rand(nw, 1.0f);
  float* gb0 = dev_rand(c.Cout, 1.0f);
  float* gwref = dev_copy(gw0, nw);
  float* gbref = dev_copy(gb0, c.Cout);
  {
    fov_conv_cfg cl = c; cl.act = FOV_ACT_LINEAR;
    FK(fov_conv2d_bwd_weight(&cl, x, dy, gwref, gbref, nullptr));
  }
  CK(cudaDeviceSynchronize());
  // fp64 truth on a sample of outputs (host): true error of the fp32 kernel and of each tensor-core mode
  const int NSAMP = 512;
  std::vector<size_t> sidx(NSAMP);
  std::vector<double> struth(NSAMP);
  {
    std::vector<float> hx(nx), hw(nw), hb(c.Cout), hy0(ny);
    CK(cudaMemcpy(hx.data(), x, nx * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hw.data(), w, nw * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, c.Cout * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hy0.data(), y0, ny * 4, cudaMemcpyDeviceToHost));
    for (int s = 0; s < NSAMP; ++s) {
      const size_t idx = (size_t)((frand() * 0.5 + 0.5) * (double)(ny - 1));
      sidx[s] = idx;
      const int co = (int)(idx % c.Cout);
      size_t r = idx / c.Cout;
      const int xx0 = (int)(r % c.W); r /= c.W;
      const int yy0 = (int)(r % c.H);
      const int n = (int)(r / c.H);
      double acc = hb[co] + c.beta * hy0[idx];
      for (int ty = 0; ty < c.kh; ++ty)
        for (int tx = 0; tx < c.kw; ++tx) {
          const int yy = yy0 + ty * c.dil_h - c.pad_h, xx = xx0 + tx * c.dil_w - c.pad_w;
          if (yy < 0 || yy >= c.H || xx < 0 || xx >= c.W) continue;
          for (int ci = 0; ci < c.Cin; ++ci)
            acc += (double)hx[(((size_t)n * c.H + yy) * c.W + xx) * c.Cin + ci] *
                   (double)hw[((size_t)(ty * c.kw + tx) * c.Cin + ci) * c.Cout + co];
        }
      if (c.act == FOV_ACT_TANH) acc = tanh(acc);
      if (c.act == FOV_ACT_RELU) acc = acc > 0 ? acc : 0;
      struth[s] = acc;
    }
  }
  auto true_err = [&](const float* yd, const char* tag) {
    std::vector<float> hy(ny);
    CK(cudaMemcpy(hy.data(), yd, ny * 4, cudaMemcpyDeviceToHost));
    double e = 0;
    for (int s = 0; s < NSAMP; ++s) { const double d = fabs((double)hy[sidx[s]] - struth[s]); if (d > e) e = d; }
    printf("  true max-abs error vs fp64 (%d samples) %-12s %.3e\n", NSAMP, tag, e);
  };
  true_err(yref, "fp32 SIMT");
  Timer tm;
  for (int math = 1; math <= 3; ++math) {
    size_t wb = fov_conv_tc_ws_bytes(&c, math, 0), wb2 = fov_conv_tc_ws_bytes(&c, math, 1);
    void *ws, *ws2;
    CK(cudaMalloc(&ws, wb + 256)); CK(cudaMalloc(&ws2, wb2 + 256));
    float* y = dev_copy(y0, ny);
    FK(fov_conv2d_fwd_tc(&c, x, w, b, y, ws, math, nullptr));
    CK(cudaDeviceSynchronize());
    report("fwd", math, compare(y, yref, ny), kTol[math]);
    { char tag[32]; snprintf(tag, sizeof tag, "tc math=%d", math); true_err(y, tag); }
    float* dx = dev_copy(dx0, nx);
    FK(fov_conv2d_bwd_data_tc(&c, dy, w, dx, ws2, math, nullptr));
    CK(cudaDeviceSynchronize());
    report("bwd_data", math, compare(dx, dxref, nx), kTol[math]);
    {
      float* gw = dev_copy(gw0, nw);
      float* gb = dev_copy(gb0, c.Cout);
      FK(fov_conv2d_bwd_weight_tc(&c, x, dy, gw, gb, math, nullptr));
      CK(cudaDeviceSynchronize());
      report("bwd_weight", math, compare(gw, gwref, nw), kTol[math] * 2);
      report("bwd_bias", math, compare(gb, gbref, c.Cout), 1e-4);
      if (time_it) {
        for (int i = 0; i < 2; ++i) FK(fov_conv2d_bwd_weight_tc(&c, x, dy, gw, gb, math, nullptr));
        tm.start();
        for (int i = 0; i < 5; ++i) FK(fov_conv2d_bwd_weight_tc(&c, x, dy, gw, gb, math, nullptr));
        const float ms = tm.stop_ms() / 5;
        const double fl = 2.0 * c.N * c.H * c.W * (double)c.Cout * c.kh * c.kw * c.Cin;
        printf("  time wgrad_tc math=%d: %.3f ms  (%.1f TFLOP/s algorithmic)\n", math, ms, fl / ms * 1e-9);
      }
      cudaFree(gw); cudaFree(gb);
    }
    if (time_it) {
      fov_conv_cfg ct = c; ct.beta = 0.f;
      for (int i = 0; i < 3; ++i) FK(fov_conv2d_fwd_tc(&ct, x, w, b, y, ws, math, nullptr));
      tm.start();
      for (int i = 0; i < 10; ++i) FK(fov_conv2d_fwd_tc(&ct, x, w, b, y, ws, math, nullptr));
      const float ms = tm.stop_ms() / 10;
      const double fl = 2.0 * c.N * c.H * c.W * (double)c.Cout * c.kh * c.kw * c.Cin;
      printf("  time fwd_tc math=%d: %.3f ms  (%.1f TFLOP/s algorithmic)\n", math, ms, fl / ms * 1e-9);
    }
    cudaFree(ws); cudaFree(ws2); cudaFree(y); cudaFree(dx);
  }
  if (time_it) {
    fov_conv_cfg ct = c; ct.beta = 0.f;
    for (int i = 0; i < 2; ++i) FK(fov_conv2d_fwd(&ct, x, w, b, yref, nullptr));
    tm.start();
    for (int i = 0; i < 5; ++i) FK(fov_conv2d_fwd(&ct, x, w, b, yref, nullptr));
    const float ms = tm.stop_ms() / 5;
    const double fl = 2.0 * c.N * c.H * c.W * (double)c.Cout * c.kh * c.kw * c.Cin;
    printf("  time fwd fp32 SIMT : %.3f ms  (%.1f TFLOP/s)\n", ms, fl / ms * 1e-9);
    tm.start();
    for (int i = 0; i < 3; ++i) FK(fov_conv2d_bwd_weight(&ct, x, dy, gwref, gbref, nullptr));
    printf("  time wgrad fp32 SIMT : %.3f ms\n", tm.stop_ms() / 3);
  }
  cudaFree(gw0); cudaFree(gb0); cudaFree(gwref); cudaFree(gbref);
  cudaFree(x); cudaFree(w); cudaFree(b); cudaFree(y0); cudaFree(yref); cudaFree(dy); cudaFree(dx0); cudaFree(dxref);
  cudaFree(wsf);
}

// ConvLSTM layer: SIMT (math 0) vs fused tensor-core step (math 1..3), forward + BPTT
extern "C" void fov_debug_wgrad_rows_timeline(int on);
extern "C" void fov_debug_seq_bwd_enable(int on);
extern "C" int fov_debug_seq_bwd_read(unsigned long long* out);
extern "C" int fov_debug_wgrad_rows_read(unsigned long long* out);
static void test_convlstm(const char* name, int B, int T, int H, int W, int Cin, int F, int kh, int kw, bool with_state,
                          bool time_it) {
  printf("[convlstm] %s B=%d T=%d HxW=%dx%d Cin=%d F=%d k=%dx%d state=%d\n", name, B, T, H, W, Cin, F, kh, kw,
         (int)with_state);
  const size_t HW = (size_t)H * W;
  const size_t nx = (size_t)B * T * HW * Cin, nh = (size_t)B * T * HW * F, nz = nh * 4, ns = (size_t)B * HW * F;
  const size_t nK = (size_t)kh * kw * Cin * 4 * F, nR = (size_t)kh * kw * F * 4 * F;
  float* x = dev_rand(nx, 1.0f);
  float* K = dev_rand(nK, 1.0f / sqrtf((float)(kh * kw * Cin)));
  float* R = dev_rand(nR, 1.0f / sqrtf((float)(kh * kw * F)));
  float* bias = dev_rand(4 * F, 0.5f);
  float* h0 = with_state ? dev_rand(ns, 0.5f) : nullptr;
  float* c0 = with_state ? dev_rand(ns, 0.5f) : nullptr;
  float* dhseq = dev_rand(nh, 1.0f);
  float* dhT = dev_rand(ns, 1.0f);
  float* dcT = dev_rand(ns, 1.0f);

  struct Out { float *hseq, *gates, *cseq, *hT, *cT, *dx, *dh0, *dc0, *gK, *gR, *gb; };
  Out o[4];
  Timer tm;
  for (int math = 0; math <= 3; ++math) {
    fov_convlstm_cfg c{};
    c.B = B; c.T = T; c.H = H; c.W = W; c.Cin = Cin; c.F = F; c.kh = kh; c.kw = kw; c.dil_h = 1; c.dil_w = 1;
    c.rec_act = FOV_REC_HARD_SIGMOID;
    c.x_b_stride = (long long)T * HW * Cin; c.x_t_stride = (long long)HW * Cin; c.x_pix_stride = Cin;
    c.h_b_stride = (long long)T * HW * F; c.h_t_stride = (long long)HW * F; c.h_pix_stride = F;
    c.training = 1; c.math = math;
    Out& r = o[math];
    r.hseq = dev_zero(nh); r.gates = dev_zero(nz); r.cseq = dev_zero(nh); r.hT = dev_zero(ns); r.cT = dev_zero(ns);
    r.dx = dev_zero(nx); r.dh0 = dev_zero(ns); r.dc0 = dev_zero(ns); r.gK = dev_zero(nK); r.gR = dev_zero(nR);
    r.gb = dev_zero(4 * F);
    fov_convlstm_io io{};
    io.x = x; io.kernel = K; io.recurrent = R; io.bias = bias; io.h0 = h0; io.c0 = c0;
    io.hseq = r.hseq; io.gates = r.gates; io.cseq = r.cseq; io.hT = r.hT; io.cT = r.cT;
    const size_t fws = fov_convlstm_fwd_ws_bytes(&c);
    void* wsf = nullptr;
    if (fws) CK(cudaMalloc(&wsf, fws + 256));
    io.ws = (float*)wsf;
    FK(fov_convlstm_fwd(&c, &io, nullptr));
    CK(cudaDeviceSynchronize());
    if (time_it) {
      tm.start();
      for (int i = 0; i < 3; ++i) FK(fov_convlstm_fwd(&c, &io, nullptr));
      printf("  time convlstm fwd math=%d: %.3f ms\n", math, tm.stop_ms() / 3);
    }
    fov_convlstm_grads g{};
    g.dhseq = dhseq; g.dhT = dhT; g.dcT = dcT; g.dx = r.dx;
    g.dh0 = with_state ? r.dh0 : nullptr; g.dc0 = with_state ? r.dc0 : nullptr;
    g.g_kernel = r.gK; g.g_recurrent = r.gR; g.g_bias = r.gb;
    const size_t bws = fov_convlstm_bwd_ws_floats(&c);
    g.ws = dev_zero(bws);
    g.dx_accumulate = 0;
    // the BPTT overwrites the saved gates: keep a copy for the comparison
    float* gates_keep = dev_copy(r.gates, nz);
    if (time_it) { tm.start(); fov_debug_wgrad_rows_timeline(1); fov_debug_seq_bwd_enable(1); }
    FK(fov_convlstm_bwd(&c, &io, &g, nullptr));
    if (time_it) {
      printf("  time convlstm bwd math=%d: %.3f ms\n", math, tm.stop_ms());
      fov_debug_wgrad_rows_timeline(0);
      fov_debug_seq_bwd_enable(0);
      unsigned long long w[8];
      fov_debug_seq_bwd_read(w);
      if (math > 0)
        printf("    seq_bwd CTA0 cycles: worker wait-mma %llu, compute+io %llu, total %llu | mma: wait %llu, issue %llu\n",
               w[0], w[1], w[2], w[4], w[5]);
      fov_debug_wgrad_rows_read(w);
      if (math > 0)
        printf("    wgrad_rows CTA0 (%llu tiles) cycles: producer wait %llu, dZ stage %llu, rows %llu, total %llu | mma: wait %llu, "
               "issue %llu, total %llu\n", w[7], w[0], w[1], w[2], w[3], w[4], w[5], w[6]);
    }
    CK(cudaDeviceSynchronize());
    cudaFree(r.gates); r.gates = gates_keep;
    cudaFree(g.ws);
    if (wsf) cudaFree(wsf);
    if (math > 0) {
      const double tol = kTol[math] * 4;
      report("hseq", math, compare(r.hseq, o[0].hseq, nh), tol);
      report("cseq", math, compare(r.cseq, o[0].cseq, nh), tol);
      report("gates", math, compare(r.gates, o[0].gates, nz), tol);
      report("hT", math, compare(r.hT, o[0].hT, ns), tol);
      report("cT", math, compare(r.cT, o[0].cT, ns), tol);
      report("dx", math, compare(r.dx, o[0].dx, nx), tol * 4);
      report("g_kernel", math, compare(r.gK, o[0].gK, nK), tol * 4);
      report("g_recurrent", math, compare(r.gR, o[0].gR, nR), tol * 4);
      report("g_bias", math, compare(r.gb, o[0].gb, 4 * F), tol * 4);
      if (with_state) {
        report("dh0", math, compare(r.dh0, o[0].dh0, ns), tol * 4);
        report("dc0", math, compare(r.dc0, o[0].dc0, ns), tol * 4);
      }
    }
  }
  for (int m = 0; m <= 3; ++m) {
    Out& r = o[m];
    cudaFree(r.hseq); cudaFree(r.gates); cudaFree(r.cseq); cudaFree(r.hT); cudaFree(r.cT); cudaFree(r.dx);
    cudaFree(r.dh0); cudaFree(r.dc0); cudaFree(r.gK); cudaFree(r.gR); cudaFree(r.gb);
  }
  cudaFree(x); cudaFree(K); cudaFree(R); cudaFree(bias); cudaFree(dhseq); cudaFree(dhT); cudaFree(dcT);
  if (h0) { cudaFree(h0); cudaFree(c0); }
}

// Persistent (one launch for all T) vs per-timestep fused ConvLSTM forward at the same arithmetic: same MMA
// order and same fp32 epilogue, so the results must agree to rounding noise.  Strided layouts as in the stacked
// models: x is a channel slice of a wider buffer, h goes into a channel slice of the concat buffer.
extern "C" void fov_debug_convlstm_persistent(int enable);
extern "C" void fov_debug_seq_enable(int on);
extern "C" void fov_debug_wgrad_rows(int enable);
extern "C" void fov_debug_convlstm_persistent_bwd(int enable);

extern "C" int fov_debug_seq_read(unsigned long long* out);
static void test_convlstm_seq(const char* name, int B, int T, int H, int W, int Cin, int F, int kh, int kw,
                              bool with_state, int training, bool time_it) {
  printf("[convlstm persistent] %s B=%d T=%d HxW=%dx%d Cin=%d F=%d k=%dx%d state=%d training=%d\n", name, B, T, H, W,
         Cin, F, kh, kw, (int)with_state, training);
  const size_t HW = (size_t)H * W;
  const int XP = Cin + 8, HP = F + 24;            // pixel strides of the wider buffers
  const size_t nx = (size_t)B * T * HW * XP, nh = (size_t)B * T * HW * HP, nc = (size_t)B * T * HW * F, nz = nc * 4,
               ns = (size_t)B * HW * F;
  const size_t nK = (size_t)kh * kw * Cin * 4 * F, nR = (size_t)kh * kw * F * 4 * F;
  float* x = dev_rand(nx, 1.0f);
  float* K = dev_rand(nK, 1.0f / sqrtf((float)(kh * kw * Cin)));
  float* R = dev_rand(nR, 1.0f / sqrtf((float)(kh * kw * F)));
  float* bias = dev_rand(4 * F, 0.5f);
  float* h0 = with_state ? dev_rand(ns, 0.5f) : nullptr;
  float* c0 = with_state ? dev_rand(ns, 0.5f) : nullptr;
  Timer tm;
  for (int math = 1; math <= 3; ++math) {
    fov_convlstm_cfg c{};
    c.B = B; c.T = T; c.H = H; c.W = W; c.Cin = Cin; c.F = F; c.kh = kh; c.kw = kw; c.dil_h = 1; c.dil_w = 1;
    c.rec_act = FOV_REC_HARD_SIGMOID;
    c.x_b_stride = (long long)T * HW * XP; c.x_t_stride = (long long)HW * XP; c.x_pix_stride = XP;
    c.h_b_stride = (long long)T * HW * HP; c.h_t_stride = (long long)HW * HP; c.h_pix_stride = HP;
    c.training = training; c.math = math;
    float* out[2][5];
    for (int mode = 0; mode < 2; ++mode) {
      fov_debug_convlstm_persistent(mode);
      float* hseq = dev_zero(nh); float* gates = dev_zero(nz); float* cseq = dev_zero(nc);
      float* hT = dev_zero(ns); float* cT = dev_zero(ns);
      fov_convlstm_io io{};
      io.x = x + 4; io.kernel = K; io.recurrent = R; io.bias = bias; io.h0 = h0; io.c0 = c0;
      io.hseq = hseq + 8; io.gates = gates; io.cseq = cseq; io.hT = hT; io.cT = cT;
      const size_t fws = fov_convlstm_fwd_ws_bytes(&c);
      void* wsf = nullptr;
      CK(cudaMalloc(&wsf, fws + 256));
      io.ws = (float*)wsf;
      FK(fov_convlstm_fwd(&c, &io, nullptr));
      CK(cudaDeviceSynchronize());
      if (time_it) {
        tm.start();
        for (int i = 0; i < 3; ++i) FK(fov_convlstm_fwd(&c, &io, nullptr));
        printf("  time convlstm fwd math=%d %s: %.3f ms\n", math, mode ? "persistent" : "per-step  ", tm.stop_ms() / 3);
        if (mode) {
          fov_debug_seq_enable(1);
          FK(fov_convlstm_fwd(&c, &io, nullptr));
          CK(cudaDeviceSynchronize());
          fov_debug_seq_enable(0);
          unsigned long long w[16];
          fov_debug_seq_read(w);
          printf("    CTA0 cycles: worker wait-mma %llu, phase A %llu, phase B %llu, x-store %llu, total %llu | mma thread: "
                 "wait-operands %llu, issue %llu, total %llu\n", w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
          printf("      phase A split: tmem-ld wait %llu, gate math %llu, h store %llu, tmem st %llu\n", w[8], w[9], w[10], w[11]);
        }
      }
      cudaFree(wsf);
      out[mode][0] = hseq; out[mode][1] = gates; out[mode][2] = cseq; out[mode][3] = hT; out[mode][4] = cT;
    }
    const double tol = 2e-6;
    report("hseq  (persistent vs per-step)", math, compare(out[1][0], out[0][0], nh), tol);
    if (training) report("gates (persistent vs per-step)", math, compare(out[1][1], out[0][1], nz), tol);
    report("cseq  (persistent vs per-step)", math, compare(out[1][2], out[0][2], nc), tol);
    report("hT    (persistent vs per-step)", math, compare(out[1][3], out[0][3], ns), tol);
    report("cT    (persistent vs per-step)", math, compare(out[1][4], out[0][4], ns), tol);
    for (int mode = 0; mode < 2; ++mode)
      for (int i = 0; i < 5; ++i) cudaFree(out[mode][i]);
  }
  fov_debug_convlstm_persistent(1);
  cudaFree(x); cudaFree(K); cudaFree(R); cudaFree(bias);
  if (h0) { cudaFree(h0); cudaFree(c0); }
}

extern "C" void fov_debug_wgrad_enable(int on);
extern "C" int fov_debug_wgrad_read(unsigned long long* out);
extern "C" void fov_debug_timeline_enable(int on);
extern "C" int fov_debug_timeline_read(unsigned long long* out, int n_words);
static void print_timeline(const char* what) {
  std::vector<unsigned long long> t(256 * 8);
  fov_debug_timeline_read(t.data(), 256 * 8);
  double ph[4] = {0, 0, 0, 0};
  int n = 0;
  unsigned long long tmin = ~0ull, tmax = 0;
  for (int c = 0; c < 148; ++c) {
    const unsigned long long* r = &t[c * 8];
    if (r[4] <= r[0]) continue;
    ph[0] += (double)(r[1] - r[0]); ph[1] += (double)(r[2] - r[1]); ph[2] += (double)(r[3] - r[2]); ph[3] += (double)(r[4] - r[3]);
    ++n;
  }
  if (n)
    printf("  timeline %-22s (cycles, mean of %d CTAs): setup %.0f | stage activations %.0f | to accumulator ready %.0f | epilogue %.0f\n",
           what, n, ph[0] / n, ph[1] / n, ph[2] / n, ph[3] / n);
}

// a handful of bench-sized launches (math = bf16x2 only) for `ncu --set full`
static void profile_mode() {
  const int B = 4096, math = 2;
  {   // conv: ConvLSTM L0 backward-data shape (dZ_t 128 ch -> dh 32 ch, 1x5) and the forward recurrent conv
    fov_conv_cfg c = mkcfg(B, 1, 33, 32, 128, 1, 5, 1, 1, FOV_ACT_LINEAR, 0.f);
    const size_t nx = (size_t)B * 33 * 32, ny = (size_t)B * 33 * 128, nw = 5 * 32 * 128;
    float *x = dev_rand(nx, 1.f), *w = dev_rand(nw, 0.1f), *b = dev_rand(128, 0.1f), *y = dev_zero(ny), *dx = dev_zero(nx);
    float *gw = dev_zero(nw), *gb = dev_zero(128);
    void *ws, *ws2;
    CK(cudaMalloc(&ws, fov_conv_tc_ws_bytes(&c, math, 0) + 256));
    CK(cudaMalloc(&ws2, fov_conv_tc_ws_bytes(&c, math, 1) + 256));
    for (int i = 0; i < 2; ++i) {
      FK(fov_conv2d_fwd_tc(&c, x, w, b, y, ws, math, nullptr));
      FK(fov_conv2d_bwd_data_tc(&c, y, w, dx, ws2, math, nullptr));
      FK(fov_conv2d_bwd_weight_tc(&c, x, y, gw, gb, math, nullptr));
    }
    CK(cudaDeviceSynchronize());
    fov_debug_timeline_enable(1);
    FK(fov_conv2d_fwd_tc(&c, x, w, b, y, ws, math, nullptr));
    CK(cudaDeviceSynchronize());
    print_timeline("conv fwd 32->128");
    FK(fov_conv2d_bwd_data_tc(&c, y, w, dx, ws2, math, nullptr));
    CK(cudaDeviceSynchronize());
    print_timeline("conv bwd-data 128->32");
    fov_debug_timeline_enable(0);
    fov_debug_wgrad_enable(1);
    FK(fov_conv2d_bwd_weight_tc(&c, x, y, gw, gb, math, nullptr));
    CK(cudaDeviceSynchronize());
    fov_debug_wgrad_enable(0);
    unsigned long long wt[8];
    fov_debug_wgrad_read(wt);
    printf("  wgrad CTA0: %llu blocks; MMA thread cycles: wait-full %llu of %llu total\n", wt[3], wt[4], wt[5]);
  }
  {   // fused ConvLSTM step, layer 0 of config 2 (training: saves the activated gates)
    const int T = 2, Cin = 6, F = 32;
    const size_t HW = 33, nx = (size_t)B * T * HW * Cin, nh = (size_t)B * T * HW * F;
    float *x = dev_rand(nx, 1.f), *K = dev_rand(5 * Cin * 4 * F, 0.2f), *R = dev_rand(5 * F * 4 * F, 0.1f), *bias = dev_rand(4 * F, 0.1f);
    fov_convlstm_cfg c{};
    c.B = B; c.T = T; c.H = 1; c.W = 33; c.Cin = Cin; c.F = F; c.kh = 1; c.kw = 5; c.dil_h = 1; c.dil_w = 1;
    c.x_b_stride = (long long)T * HW * Cin; c.x_t_stride = (long long)HW * Cin; c.x_pix_stride = Cin;
    c.h_b_stride = (long long)T * HW * F; c.h_t_stride = (long long)HW * F; c.h_pix_stride = F;
    c.training = 1; c.math = math;
    fov_convlstm_io io{};
    io.x = x; io.kernel = K; io.recurrent = R; io.bias = bias;
    io.hseq = dev_zero(nh); io.gates = dev_zero(nh * 4); io.cseq = dev_zero(nh);
    void* wsf; CK(cudaMalloc(&wsf, fov_convlstm_fwd_ws_bytes(&c) + 256)); io.ws = (float*)wsf;
    for (int i = 0; i < 2; ++i) FK(fov_convlstm_fwd(&c, &io, nullptr));
    CK(cudaDeviceSynchronize());
    fov_debug_timeline_enable(1);
    FK(fov_convlstm_fwd(&c, &io, nullptr));
    CK(cudaDeviceSynchronize());
    print_timeline("fused LSTM step (t=1)");
    fov_debug_timeline_enable(0);
  }
  printf("profile mode done\n");
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  if (!fov_device_is_sm100()) { printf("not an sm_100 device\n"); return 1; }
  if (argc > 1 && !strcmp(argv[1], "profile")) { profile_mode(); return 0; }
  if (argc > 1 && !strcmp(argv[1], "wr")) {      // fused row-shift weight gradient: parity vs fp32, timing on/off
    test_convlstm("m3 L0", 6, 4, 1, 33, 6, 32, 1, 5, false, false);
    test_convlstm("m3 L1", 5, 3, 1, 33, 32, 16, 1, 5, true, false);
    test_convlstm("m3 L2", 5, 3, 1, 33, 16, 8, 1, 5, false, false);
    test_convlstm("3x3 on 5x4", 9, 3, 5, 4, 12, 16, 3, 3, true, false);
    test_convlstm("traj 1x30 Cin 3", 4, 6, 1, 30, 3, 32, 1, 5, true, false);
    for (int on = 0; on < 3; ++on) {
      fov_debug_wgrad_rows(on >= 1);
      fov_debug_convlstm_persistent_bwd(on >= 2);
      printf("=== fused row-shift wgrad %s, persistent BPTT %s ===\n", on >= 1 ? "ON" : "OFF", on >= 2 ? "ON" : "OFF");
      test_convlstm("T m3 L0 B=2048", 2048, 20, 1, 33, 6, 32, 1, 5, false, true);
      test_convlstm("T m3 L1 B=2048", 2048, 20, 1, 33, 32, 16, 1, 5, false, true);
      test_convlstm("T m3 L2 B=2048", 2048, 20, 1, 33, 16, 8, 1, 5, false, true);
    }
    printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK (%d failures)\n", g_fail);
    return g_fail ? 1 : 0;
  }
  if (argc > 1 && !strcmp(argv[1], "dense")) {    // phase timelines + timings of the wide dense layers of config 2 (B = 4096)
    const int math = 2;
    const int shapes[2][3] = {{81920, 1848, 198}, {40960, 1848, 256}};
    for (int si = 0; si < 2; ++si) {
      const int rows = shapes[si][0], Cin = shapes[si][1], Cout = shapes[si][2];
      fov_conv_cfg c = mkcfg(rows, 1, 1, Cin, Cout, 1, 1, 1, 1, FOV_ACT_LINEAR, 0.f);
      float *x = dev_rand((size_t)rows * Cin, 1.f), *w = dev_rand((size_t)Cin * Cout, 0.05f), *b = dev_rand(Cout, 0.1f);
      float *y = dev_zero((size_t)rows * Cout), *dx = dev_zero((size_t)rows * Cin), *gw = dev_zero((size_t)Cin * Cout), *gb = dev_zero(Cout);
      void *ws, *ws2;
      CK(cudaMalloc(&ws, fov_conv_tc_ws_bytes(&c, math, 0) + 256));
      CK(cudaMalloc(&ws2, fov_conv_tc_ws_bytes(&c, math, 1) + 256));
      printf("[dense] rows=%d %d -> %d\n", rows, Cin, Cout);
      Timer tm;
      for (int pass = 0; pass < 2; ++pass) {
        if (pass) tm.start();
        for (int i = 0; i < (pass ? 5 : 1); ++i) FK(fov_conv2d_fwd_tc(&c, x, w, b, y, ws, math, nullptr));
        if (pass) printf("  fwd      %.3f ms\n", tm.stop_ms() / 5);
        if (pass) tm.start();
        for (int i = 0; i < (pass ? 5 : 1); ++i) FK(fov_conv2d_bwd_data_tc(&c, y, w, dx, ws2, math, nullptr));
        if (pass) printf("  bwd-data %.3f ms\n", tm.stop_ms() / 5);
        if (pass) tm.start();
        for (int i = 0; i < (pass ? 5 : 1); ++i) FK(fov_conv2d_bwd_weight_tc(&c, x, y, gw, gb, math, nullptr));
        if (pass) printf("  wgrad    %.3f ms\n", tm.stop_ms() / 5);
      }
      CK(cudaDeviceSynchronize());
      fov_debug_timeline_enable(1);
      FK(fov_conv2d_fwd_tc(&c, x, w, b, y, ws, math, nullptr));
      CK(cudaDeviceSynchronize());
      print_timeline("dense fwd");
      FK(fov_conv2d_bwd_data_tc(&c, y, w, dx, ws2, math, nullptr));
      CK(cudaDeviceSynchronize());
      print_timeline("dense bwd-data");
      fov_debug_timeline_enable(0);
      cudaFree(x); cudaFree(w); cudaFree(b); cudaFree(y); cudaFree(dx); cudaFree(gw); cudaFree(gb); cudaFree(ws); cudaFree(ws2);
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "bw")) {       // one mid-size stacked-layer backward (for compute-sanitizer runs)
    const int B = argc > 2 ? atoi(argv[2]) : 300;
    if (argc > 3) fov_debug_seq_bwd_enable(atoi(argv[3]));
    test_convlstm("m3 L1 mid", B, 20, 1, 33, 32, 16, 1, 5, false, false);
    test_convlstm("m3 L2 mid", B, 20, 1, 33, 16, 8, 1, 5, false, false);
    printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK (%d failures)\n", g_fail);
    return g_fail ? 1 : 0;
  }
  if (argc > 1 && !strcmp(argv[1], "seq")) {
    test_convlstm_seq("m3 L0", 7, 5, 1, 33, 6, 32, 1, 5, false, 1, false);
    test_convlstm_seq("m3 L1", 50, 20, 1, 33, 32, 16, 1, 5, true, 1, false);
    test_convlstm_seq("m3 L2", 13, 4, 1, 33, 16, 8, 1, 5, false, 0, false);
    test_convlstm_seq("3x3 on 5x4", 9, 3, 5, 4, 12, 16, 3, 3, true, 1, false);
    test_convlstm_seq("traj 1x30 Cin 3", 4, 6, 1, 30, 3, 32, 1, 5, true, 0, false);
    test_convlstm_seq("T m3 L0 B=2048", 2048, 20, 1, 33, 6, 32, 1, 5, false, 1, true);
    test_convlstm_seq("T m3 L1 B=2048", 2048, 20, 1, 33, 32, 16, 1, 5, false, 1, true);
    test_convlstm_seq("T m3 L2 B=2048", 2048, 20, 1, 33, 16, 8, 1, 5, false, 1, true);
    printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK (%d failures)\n", g_fail);
    return g_fail ? 1 : 0;
  }
  // --- convolution family ---
  test_conv("dense 64->16", mkcfg(300, 1, 1, 64, 16, 1, 1, 1, 1, FOV_ACT_LINEAR, 0.f), false);
  test_conv("m3 L0 rec", mkcfg(64, 1, 33, 32, 128, 1, 5, 1, 1, FOV_ACT_LINEAR, 1.f), false);
  test_conv("m3 L0 in (Cin 6)", mkcfg(40, 1, 33, 6, 128, 1, 5, 1, 1, FOV_ACT_TANH, 0.f), false);
  test_conv("5x5 30->32 36x18", mkcfg(3, 36, 18, 30, 32, 5, 5, 1, 1, FOV_ACT_RELU, 0.f), false);
  test_conv("4x3 dil 2 odd", mkcfg(2, 9, 7, 5, 7, 4, 3, 2, 1, FOV_ACT_LINEAR, 1.f), false);
  test_conv("dense 1848->256", mkcfg(500, 1, 1, 1848, 256, 1, 1, 1, 1, FOV_ACT_LINEAR, 0.f), false);
  test_conv("dense 1848->198", mkcfg(200, 1, 1, 1848, 198, 1, 1, 1, 1, FOV_ACT_LINEAR, 0.f), false);
  test_conv("head 5x5 56->512", mkcfg(2, 36, 18, 56, 512, 5, 5, 1, 1, FOV_ACT_RELU, 0.f), false);
  test_conv("conv1d k7 56->40", mkcfg(5, 1, 30, 56, 40, 1, 7, 1, 1, FOV_ACT_RELU, 0.f), false);
  // --- fused ConvLSTM ---
  test_convlstm("m3 L0", 6, 4, 1, 33, 6, 32, 1, 5, false, false);
  test_convlstm("m3 L1", 5, 3, 1, 33, 32, 16, 1, 5, true, false);
  test_convlstm("m3 L2", 5, 3, 1, 33, 16, 8, 1, 5, false, false);
  test_convlstm("m4 L0", 2, 3, 36, 18, 30, 32, 5, 5, true, false);
  // --- persistent ConvLSTM (whole images per tile) vs the per-timestep launches ---
  test_convlstm_seq("m3 L0", 7, 5, 1, 33, 6, 32, 1, 5, false, 1, false);
  test_convlstm_seq("m3 L1", 50, 20, 1, 33, 32, 16, 1, 5, true, 1, false);
  test_convlstm_seq("m3 L2", 13, 4, 1, 33, 16, 8, 1, 5, false, 0, false);
  test_convlstm_seq("3x3 on 5x4", 9, 3, 5, 4, 12, 16, 3, 3, true, 1, false);
  test_convlstm_seq("traj 1x30 Cin 3", 4, 6, 1, 30, 3, 32, 1, 5, true, 0, false);
  if (!quick) {
    // --- timings at bench size (config 2, B=4096) ---
    test_conv("T m3 L0 rec B=4096", mkcfg(4096, 1, 33, 32, 128, 1, 5, 1, 1, FOV_ACT_LINEAR, 0.f), true);
    test_conv("T dense256 B=4096", mkcfg(40960, 1, 1, 1848, 256, 1, 1, 1, 1, FOV_ACT_LINEAR, 0.f), true);
    test_conv("T head 56->512 B=32", mkcfg(32, 36, 18, 56, 512, 5, 5, 1, 1, FOV_ACT_RELU, 0.f), true);
    test_conv("T head 512->1024 B=32", mkcfg(32, 36, 18, 512, 1024, 5, 5, 1, 1, FOV_ACT_RELU, 0.f), true);
    test_convlstm("T m3 L0 B=2048", 2048, 20, 1, 33, 6, 32, 1, 5, false, true);
    test_convlstm_seq("T m3 L0 B=2048", 2048, 20, 1, 33, 6, 32, 1, 5, false, 1, true);
    test_convlstm_seq("T m3 L1 B=2048", 2048, 20, 1, 33, 32, 16, 1, 5, false, 1, true);
    test_convlstm_seq("T m3 L2 B=2048", 2048, 20, 1, 33, 16, 8, 1, 5, false, 1, true);
  }
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK (%d failures)\n", g_fail);
  return g_fail ? 1 : 0;
}
