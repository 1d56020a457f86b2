// Internal (non-exported) launch helpers shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include "../../include/fov360.h"
#include "../../include/fov_debug.h"

struct GatesFwdArgs {
  long long npix;   // B*H*W
  int HW, F, rec;
  float* z;         long long z_img;              // pixel stride 4F
  const float* c_prev; long long cp_img;          // pixel stride F (may be NULL)
  float* c_out;     long long c_img;              // pixel stride F
  float* h_out;     long long h_img; int h_pix;
  float *hT, *cT;   // optional dense (B,HW,F)
};

struct GatesBwdArgs {
  long long npix;
  int HW, F, rec;
  float* gates;     long long z_img;              // in: activated gates, out: dz
  const float* c_t; long long c_img;
  const float* c_prev; long long cp_img;          // may be NULL
  const float* dh_ext; long long dhe_img; int dhe_pix;   // may be NULL
  const float* dh_rec;                            // dense (B,HW,F), may be NULL
  const float* dc_in;                             // dense, may be NULL
  float* dc_out;                                  // dense
};

int fov_launch_gates_fwd(const GatesFwdArgs& a, cudaStream_t st);
int fov_launch_gates_bwd(const GatesBwdArgs& a, cudaStream_t st);
int fov_launch_mul_mask(long long n_img, long long img_elems, int C, const float* x, long long x_img,
                        int x_pix, const float* mask, float* out, cudaStream_t st);

// conv helpers (conv_igemm.cu): flip/transposed weights once, reuse across timesteps
int fov_conv_flip_weights(const fov_conv_cfg* fwd_cfg, const float* w, float* wt, cudaStream_t st);
int fov_conv_bwd_data_preflipped(const fov_conv_cfg* fwd_cfg, const float* dy, const float* wt, float* dx,
                                 cudaStream_t st);

// Persistent fc-LSTM forward on tensor cores (lstm_seq2seq_tc.cu); same contract as fov_lstm_seq2seq_fwd
bool lstm_tc_supported(const fov_lstm_cfg* cfg);
int lstm_tc_fwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w, const fov_lstm_io* io, cudaStream_t st);
size_t lstm_tc_fwd_ws_floats(const fov_lstm_cfg* cfg);
// Time-batched input projection P = x . W on tensor cores, fed by TMA (xproj_tc.cu); P is the tiled workspace the
// tensor-core forward reads, xh (optional) receives the x part of the saved [h | x | 0] rows
size_t lstm_xproj_ws_floats(int B, int T);
bool lstm_xproj_supported(int B, int T, int in, int math, const float* x);
int lstm_xproj_run(int B, int T, int in, int math, const float* x, const float* W, float* P, float* xh, int K_xh,
                   cudaStream_t st);
// Persistent fc-LSTM BPTT on tensor cores: writes dz_enc / dz_dec / dpre like the fp32 kernel (no weight gradients)
bool lstm_tc_bwd_supported(const fov_lstm_cfg* cfg);
int lstm_tc_bwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w, const fov_lstm_io* io, const fov_lstm_grads* g,
                cudaStream_t st);

// ---- tensor-core (tcgen05) implicit-GEMM convolution family (conv_tc.cu) -------------------
// One K segment of the implicit-GEMM A operand: an NHWC activation tensor convolved with its
// own (kh,kw) taps.  Two segments = the fused ConvLSTM gate GEMM [x taps | h taps] x [K ; R].
struct TcSeg {
  const float* x;
  long long img_outer, img_inner;   // element offset of image n: (n / T_inner)*outer + (n % T_inner)*inner
  int pix_stride;
  int Cin;
  int kh, kw, dil_h, dil_w, pad_h, pad_w;
  const float* w;                   // Keras-layout weights of this segment
  int w_mode;                       // 0: w[(tap*Cin+ci)*Cout+n]   1 (backward-data): w[((taps-1-tap)*Cout+n)*Cin+ci]
};
enum { TC_EPI_CONV = 0, TC_EPI_LSTM = 1 };
struct TcConv {
  int nseg;
  TcSeg seg[2];
  int N_img, T_inner, H, W;
  int Cout;
  int math;                         // bf16 terms per operand: 1 (bf16), 2 (3 MMAs), 3 (6 MMAs, ~fp32)
  void* ws;                         // packed-weight workspace, tc_conv_ws_bytes() bytes
  int prepacked;                    // 1: ws already holds the packed weights of this exact problem
  int no_mt2;                       // two accumulator tiles per CTA: 1 never, 0 auto, -1 whenever it fits (fov_debug_conv_mt2)
  // TC_EPI_CONV epilogue: y = act(acc + bias + beta*y)
  int epi;
  const float* bias;
  float* y;
  long long y_outer, y_inner;
  int y_pix_stride;
  int act;
  float beta;
  // TC_EPI_LSTM epilogue (Cout = 4F, gate blocks i,f,c,o): c_t = f*c_prev + i*tanh(zc); h = o*tanh(c_t)
  int rec_act;
  const float* c_prev; long long cp_outer, cp_inner;   // pixel stride F; may be NULL (zero state)
  float* c_out;        long long c_outer, c_inner;     // pixel stride F
  float* h_out;        long long h_outer, h_inner; int h_pix_stride;
  float* gates_out;    long long g_outer, g_inner;     // optional activated gates (pixel stride 4F)
  float *hT, *cT;                                      // optional dense copies (N_img,HW,F)
};
// Geometry of a shifted-tap GEMM whose whole K extent stays resident in shared memory (<= 64 channels per
// segment, one n tile): what the persistent ConvLSTM kernels need to stage operands and issue MMAs.
struct TcStepSeg {
  int Cin, Cin_p, cp_log2, cw, lpr_log2, row_bytes, swz_mask, term_bytes, R, minshift;
  int taps, kw, dil_h, dil_w, pad_h, pad_w, k_begin;
  int nch;          // 64-channel chunks of the segment (> 1: k order is tap-major inside each chunk pair, see conv_tc.cu)
  unsigned desc_hi;
};
struct TcStepPlan {
  TcStepSeg seg[2];
  int mode_b;       // segment wider than 64 channels: regions hold nch chunks of R rows x 128 B
  int nseg, Hp, Wp, PLh, PLw, K_total, KB, BLOCK_N, NS;
  size_t w_bytes;   // packed weights: KB x NS x BLOCK_N x 128 B
};
int tc_conv_step_plan(const TcConv& c, TcStepPlan* out);
bool tc_conv_supported(const TcConv& c);   // shape fits the shifted-tap kernel (no error is recorded)
size_t tc_conv_ws_bytes(const TcConv& c);
int tc_conv_pack(const TcConv& c, cudaStream_t st);
int tc_conv_run(const TcConv& c, cudaStream_t st);

// Persistent ConvLSTM2D layer (convlstm_seq_tc.cu): every timestep in one launch when whole images fit a tile.
bool tc_convlstm_seq_supported(const fov_convlstm_cfg* c, const TcConv& step);
int tc_convlstm_seq_fwd(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const TcConv& step, cudaStream_t st);

// kT (optional): the input backward-data convolution; when it shares the k order of rT its result dx rides in the
// same MMAs (extra accumulator columns) and the separate time-batched dx launch is not needed.
bool tc_convlstm_seq_bwd_supported(const fov_convlstm_cfg* c, const TcConv& rT, const TcConv* kT);
int tc_convlstm_seq_bwd_wave_groups(const fov_convlstm_cfg* c, const TcConv& rT, const TcConv* kT);
int tc_convlstm_seq_wave_groups(const fov_convlstm_cfg* c, const TcConv& step);
int tc_convlstm_seq_bwd(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const fov_convlstm_grads* gr,
                        const TcConv& rT, const TcConv* kT, cudaStream_t st);
// true: that launch also accumulates g_kernel / g_recurrent / g_bias (no separate weight-gradient launch, dZ stays on chip)
bool tc_convlstm_seq_bwd_fuses_wgrad(const fov_convlstm_cfg* c, const fov_convlstm_io* io, const fov_convlstm_grads* gr,
                                     const TcConv& rT, const TcConv* kT);

// Tensor-core weight gradient: gw[(tap*Cin+ci)*Cout+n] += sum_pixels x(pixel+tap, ci) * dy(pixel, n),
// gbias[n] += sum_pixels dy(pixel, n).  Both operands are read in their natural NHWC layout and
// fed to tcgen05.mma as MN-major tiles; the pixel reduction is split across CTAs (atomic adds).
struct TcWgrad {
  const float* x;  long long x_outer, x_inner;   int x_pix_stride;  int Cin;
  int kh, kw, dil_h, dil_w, pad_h, pad_w;
  const float* dy; long long dy_outer, dy_inner; int dy_pix_stride; int Cout;
  int N_img, T_inner, H, W;
  int math;
  float* gw;       // may be NULL (bias gradient only)
  float* gbias;    // may be NULL
};
int tc_wgrad_run(const TcWgrad& c, cudaStream_t st);
// TMA-fed form for wide k x k convolutions (wgrad_planes_tc.cu): both operands converted once into bf16 planes
size_t tc_wgrad_planes_ws_bytes(const TcWgrad& c);
int tc_wgrad_planes_run(const TcWgrad& c, void* ws, cudaStream_t st);

// Fused ConvLSTM weight gradient (wgrad_rows_tc.cu): gK, gR and gb of a layer in one launch, taps as descriptor
// offsets (no gather).  dZ is the dense (B,T,HW,Cout) gate-gradient buffer; segment s pairs image (b, t + t_shift)
// of its tensor with dZ image (b,t); images with t + t_shift < 0 come from x0 (or read as zeros).
struct TcWgradRowsSeg {
  const float* x;  long long b_stride, t_stride; int pix_stride;
  const float* x0; long long x0_b_stride; int x0_pix_stride;
  int t_shift, Cin, kh, kw, dil_h, dil_w, pad_h, pad_w;
  float* gw;       // (kh,kw,Cin,Cout), accumulated; may be NULL
};
struct TcWgradRows {
  int nseg;
  TcWgradRowsSeg seg[2];
  const float* dz;
  int B, T, H, W, Cout, math;
  float* gbias;    // may be NULL
};
bool tc_wgrad_rows_supported(const TcWgradRows& c);
int tc_wgrad_rows_run(const TcWgradRows& c, cudaStream_t st);
