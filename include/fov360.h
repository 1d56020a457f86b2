/*
 * fov360.h — C ABI of libfov360.so, the B200 (sm_100a) kernels behind the
 * LongTerm360FoV sequence-prediction hot path.
 *
 * The reference (ChengeLi/LongTerm360FoV) has no FFI of its own: its hot path is
 * the set of Keras layer calls listed below, executed by TF1.  Each entry point
 * here replaces the Keras/TF op behind one of those call sites; the Python layer
 * in longterm360fov_b200/ binds them with ctypes and wraps them in
 * torch.autograd.Functions (INTEGRATION.md shows the script-side swap).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous float32 unless stated,
 *     16-byte aligned; the caller owns all buffers (inputs, outputs, saved
 *     activations, workspaces); the library allocates nothing and keeps no state except the
 *     NCCL communicator of fov_dp_init and the diagnostic switches of include/fov_debug.h.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it.
 *   - weight layouts are Keras': LSTM kernel (in,4H) / recurrent_kernel (H,4H) /
 *     bias (4H), gate blocks i,f,c,o; ConvLSTM2D kernel (kh,kw,Cin,4F) /
 *     recurrent_kernel (kh,kw,F,4F); Dense (in,out); Conv2D (kh,kw,Cin,Cout).
 *   - return value: 0 = ok, negative = error (fov_last_error() gives a
 *     thread-local message).  Nothing throws across the boundary.
 */
#ifndef FOV360_H
#define FOV360_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FOV_OK = 0, FOV_ERR_ARG = -1, FOV_ERR_UNSUPPORTED = -2, FOV_ERR_CUDA = -3 };
enum { FOV_ACT_LINEAR = 0, FOV_ACT_TANH = 1, FOV_ACT_RELU = 2 };
enum { FOV_REC_HARD_SIGMOID = 0, FOV_REC_SIGMOID = 1 };

const char* fov_last_error(void);
int fov_version(void);
/* kernel launches issued by the library since it was loaded (diagnostics). */
unsigned long long fov_launch_count(void);
/* 1 when the running device is sm_100 (B200). */
int fov_device_is_sm100(void);

/* ------------------------------------------------------------------------- *
 * Persistent fc-LSTM encoder-decoder.
 * Replaces: keras LSTM(64,return_state) + LSTM(64,return_sequences) + Dense at
 *   mycode/FoV_seq2seq.py:82-97 (teacher forcing), the host decode loop
 *   :154-178, the in-graph autoregressive loop
 *   mycode/FoV_seq2seq_no_teac_forc.py:88-129, the mu/var form
 *   mycode/FoV_seq2seq_mu_var.py:219-234,286-311 and the target branch of
 *   mycode/others_LSTM_span_whole.py:119-121,262-271.
 * One launch runs T_enc encoder steps then T_dec decoder steps; weights stay in
 * shared memory, gates/cell in registers, no per-step launch.
 * ------------------------------------------------------------------------- */
typedef struct {
  int B;              /* sequences */
  int T_enc, T_dec;   /* either may be 0 (encoder_model / decoder_model split) */
  int in_enc, in_dec; /* input widths */
  int H;              /* latent_dim; 64 supported */
  int out_dim;        /* head width (0 = no head) */
  int teacher_forcing;/* 1: x_dec is (B,T_dec,in_dec); 0: x_dec is (B,1,in_dec), y re-fed */
  int head_act;       /* FOV_ACT_* */
  int rec_act;        /* FOV_REC_* */
  int dec_zero_init;  /* 1: decoder starts from a zero state (FoV_seq2seq_no_teac_forc.py:29) */
  int training;       /* 1: write the saved-activation buffers */
  int math;           /* FOV_MATH_* (declared below): 0 = fp32 CUDA-core kernel; 1..3 = the forward runs on tensor
                         cores (tcgen05, that many bf16 terms per operand): inputs up to 16 features ride in the
                         per-step GEMM, wider ones (FoV_seq2seq's 90) go through the time-batched projection (io.ws) */
} fov_lstm_cfg;

typedef struct {
  const float *enc_kernel, *enc_recurrent, *enc_bias;   /* (in_enc,4H) (H,4H) (4H) */
  const float *dec_kernel, *dec_recurrent, *dec_bias;   /* (in_dec,4H) (H,4H) (4H) */
  const float *head_kernel, *head_bias;                 /* (H,out) (out) */
} fov_lstm_weights;

typedef struct {                 /* per LSTM, all (B,T,...) row-major */
  float *xh;                     /* (B,T,XS): [h_{t-1} | x_t | 0], XS = (H+in+3)/4*4 (rows padded to 16 bytes so the
                                    tensor-core weight-gradient GEMM reads them with aligned vector loads) */
  float *gates;                  /* (B,T,4H): activated i,f,g,o */
  float *c;                      /* (B,T,H) */
  float *hseq;                   /* (B,T,H)  (also the return_sequences output) */
} fov_lstm_saved;

typedef struct {
  const float *x_enc;            /* (B,T_enc,in_enc) */
  const float *x_dec;            /* (B,T_dec,in_dec) or (B,1,in_dec) */
  const float *extra;            /* optional (B,T_dec,out): added to the head pre-activation */
  const float *h0, *c0;          /* optional initial state (B,H) */
  float *y;                      /* (B,T_dec,out) */
  float *hT, *cT;                /* optional final state (B,H) */
  fov_lstm_saved enc, dec;       /* any pointer may be NULL when training == 0 */
  float *ws;                     /* optional workspace of fov_lstm_fwd_ws_bytes() bytes: holds the time-batched input
                                    projection x_t . W of every step (one TMA-fed tcgen05 GEMM per phase) that lets
                                    inputs wider than 16 features run the tensor-core recurrence */
} fov_lstm_io;

/* bytes of fov_lstm_io.ws this configuration can use (0: none needed) */
size_t fov_lstm_fwd_ws_bytes(const fov_lstm_cfg* cfg);

int fov_lstm_seq2seq_fwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w,
                         const fov_lstm_io* io, void* stream);

typedef struct {
  const float *dy;               /* (B,T_dec,out) gradient of the loss w.r.t. y */
  const float *dhseq_enc;        /* optional (B,T_enc,H): gradient w.r.t. enc.hseq */
  const float *y;                /* forward output (for tanh'/relu') */
  float *dz_enc, *dz_dec;        /* scratch+output (B,T,4H): gate pre-activation gradients */
  float *dpre;                   /* (B,T_dec,out): head pre-activation gradient == d(extra) */
  /* weight gradients, ACCUMULATED (+=) */
  float *g_enc_kernel, *g_enc_recurrent, *g_enc_bias;
  float *g_dec_kernel, *g_dec_recurrent, *g_dec_bias;
  float *g_head_kernel, *g_head_bias;
  float *ws;                     /* optional workspace of fov_lstm_bwd_ws_floats() floats: with it (and cfg.math != 0)
                                    each LSTM's [dU; dW; db] is ONE tcgen05 launch over the padded xh rows */
  const float *dhseq_dec;        /* optional (B,T_dec,H): gradient w.r.t. dec.hseq (stacked LSTMs: the layer above reads
                                    this layer's hidden sequence, mycode/Fov_seq2seq_2layers.py:232-272) */
  void *wgrad_stream;            /* optional cudaStream_t: the weight-gradient launches (nothing on the backward chain reads
                                    them) go to this stream after an event on `stream`; the caller joins it before it
                                    reads the gradients and keeps dz / saved tensors / ws alive until then.  NULL: `stream` */
} fov_lstm_grads;
size_t fov_lstm_bwd_ws_floats(const fov_lstm_cfg* cfg);

int fov_lstm_seq2seq_bwd(const fov_lstm_cfg* cfg, const fov_lstm_weights* w,
                         const fov_lstm_io* io, const fov_lstm_grads* g, void* stream);

/* ------------------------------------------------------------------------- *
 * Convolution / dense family (implicit GEMM, channels-last, stride 1).
 * Replaces: Dense (mycode/others_LSTM_span_whole.py:105-109,265-271), Conv2D /
 *   Conv1D heads (mycode/convlstm_seq2seq.py:175-189,230-258) and the gate
 *   convolutions inside ConvLSTM2D (mycode/others_LSTM_span_whole.py:88-100,
 *   mycode/convlstm_seq2seq.py:100-126,146-165).
 * A Dense layer is the 1x1 case: N=rows, H=W=1, Cin=in, Cout=out.
 * y[n,p,co] = act( sum_{tap,ci} x[n, p+tap*dil-pad, ci] * w[tap,ci,co] + bias[co]
 *                  + beta * y_old[n,p,co] )
 * ------------------------------------------------------------------------- */
typedef struct {
  int N, H, W, Cin, Cout;
  int kh, kw, dil_h, dil_w;
  int pad_h, pad_w;             /* low-side padding; TF 'same' = ((k-1)*dil)/2 */
  long long x_img_stride;       /* elements between consecutive images of x */
  int x_pix_stride;             /* elements between consecutive pixels of x (>= Cin) */
  long long y_img_stride;
  int y_pix_stride;
  int act;                      /* FOV_ACT_* */
  float beta;                   /* 0 or 1 */
} fov_conv_cfg;

int fov_conv2d_fwd(const fov_conv_cfg* cfg, const float* x, const float* w, const float* bias,
                   float* y, void* stream);
/* dx = conv^T(dy, w).  ws: workspace of kh*kw*Cin*Cout floats (flipped weights).
 * cfg describes the FORWARD conv (x = its input, y = its output); beta applies to dx. */
int fov_conv2d_bwd_data(const fov_conv_cfg* cfg, const float* dy, const float* w, float* dx,
                        float* ws, void* stream);
/* gw += x^T (*) dy ; gbias += colsum(dy)   (either may be NULL) */
int fov_conv2d_bwd_weight(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw,
                          float* gbias, void* stream);
/* y = act'(y) * dy elementwise helper for fused bias+activation layers:
 * dpre = dy * act'(y_out)  over `rows` x `cols` with row strides. */
int fov_act_bwd(int act, long long rows, int cols, const float* y, long long y_stride,
                const float* dy, long long dy_stride, float* dpre, long long dpre_stride,
                void* stream);

/* Tap-stacked narrow convolutions (Cout <= 32, e.g. the 1024 -> 30 head of mycode/convlstm_seq2seq.py:179-181): the
 * caller runs the kh x 1 convolution with Cout' = kw*Cp stacked output columns (tx, c) on fov_conv2d_fwd_tc and folds
 * the kw column taps here: y[r,w,c] = act(bias[c] + sum_tx yp[r, w+tx-pad_w, tx*Cp+c]); rows = N*H image rows.
 * fov_tapstack_expand is the transpose for the backward pass: dyp[r,w,tx*Cp+c] = dpre[r, w-tx+pad_w, c] (else 0). */
int fov_tapstack_reduce(long long rows, int W, int kw, int pad_w, int Cp, int Cout, const float* yp, const float* bias,
                        int act, float* y, void* stream);
int fov_tapstack_expand(long long rows, int W, int kw, int pad_w, int Cp, int Cout, const float* dpre, float* dyp,
                        void* stream);

/* ------------------------------------------------------------------------- *
 * Tensor-core (tcgen05 / TMEM) forms of the convolution family.
 * Same operands and semantics as fov_conv2d_fwd / _bwd_data / _bwd_weight; the
 * activations and weights stay fp32 in HBM and are split on the fly into `math`
 * bf16 terms per operand, accumulated in fp32 in tensor memory:
 *   FOV_MATH_BF16   1 term  (1 MMA  per k-step, ~8 mantissa bits per operand)
 *   FOV_MATH_BF16X2 2 terms (3 MMAs per k-step, ~16 bits)
 *   FOV_MATH_BF16X3 3 terms (6 MMAs per k-step, ~24 bits: fp32-grade results)
 * `ws` is a caller-owned workspace of fov_conv_tc_ws_bytes() bytes that receives
 * the repacked weights (the call repacks them every time; it keeps no state).
 * ------------------------------------------------------------------------- */
enum { FOV_MATH_FP32 = 0, FOV_MATH_BF16 = 1, FOV_MATH_BF16X2 = 2, FOV_MATH_BF16X3 = 3 };
size_t fov_conv_tc_ws_bytes(const fov_conv_cfg* cfg, int math, int bwd_data);
int fov_conv2d_fwd_tc(const fov_conv_cfg* cfg, const float* x, const float* w, const float* bias,
                      float* y, void* ws, int math, void* stream);
int fov_conv2d_bwd_data_tc(const fov_conv_cfg* cfg, const float* dy, const float* w, float* dx,
                           void* ws, int math, void* stream);
/* The same two calls with the weight repack hoisted out: fov_conv_tc_pack fills ws (fov_conv_tc_ws_bytes bytes) for
 * the forward (bwd_data = 0) or the backward-data (1) convolution of cfg; the *_packed calls only read it.  For callers
 * that run one layer many times between weight updates (the 10 decoder steps of mycode/convlstm_seq2seq.py:211-238). */
int fov_conv_tc_pack(const fov_conv_cfg* cfg, const float* w, void* ws, int math, int bwd_data, void* stream);
int fov_conv2d_fwd_tc_packed(const fov_conv_cfg* cfg, const float* x, const float* bias, float* y,
                             const void* ws, int math, void* stream);
int fov_conv2d_bwd_data_tc_packed(const fov_conv_cfg* cfg, const float* dy, float* dx, const void* ws,
                                  int math, void* stream);
/* gw += x^T (*) dy ; gbias += colsum(dy): pixel reduction on tensor cores, split across CTAs */
int fov_conv2d_bwd_weight_tc(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw,
                             float* gbias, int math, void* stream);
/* Same with a caller-owned workspace of fov_conv_wgrad_ws_bytes() bytes (0: this shape does not use one): wide
 * k x k convolutions (the 56 -> 512 -> 1024 -> 30 heads of mycode/convlstm_seq2seq.py:175-181) then convert both
 * operands ONCE into bf16 term planes in a zero-padded frame and run a TMA-fed (cp.async.bulk.tensor) tcgen05
 * GEMM over them; other shapes fall through to fov_conv2d_bwd_weight_tc. */
size_t fov_conv_wgrad_ws_bytes(const fov_conv_cfg* cfg, int math);
int fov_conv2d_bwd_weight_tc_ws(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw,
                                float* gbias, void* ws, int math, void* stream);

/* ------------------------------------------------------------------------- *
 * ConvLSTM2D layer over a whole sequence (return_sequences=True, return_state).
 * Replaces: keras ConvLSTM2D at mycode/others_LSTM_span_whole.py:88-100 and
 *   mycode/convlstm_seq2seq.py:100-126 (T=10/20) and the one-step decoder calls
 *   :213-218 (T=1 with given initial state).
 * x: (B,T,H,W,Cin) with strides; hseq written at (B,T,H,W,F) with strides so
 * three layers can write straight into the channel-concatenated buffer.
 * ------------------------------------------------------------------------- */
typedef struct {
  int B, T, H, W, Cin, F;
  int kh, kw, dil_h, dil_w;      /* dilation applies to the input conv only */
  int rec_act;
  long long x_b_stride, x_t_stride;  int x_pix_stride;
  long long h_b_stride, h_t_stride;  int h_pix_stride;   /* layout of hseq */
  int training;
  int math;                      /* FOV_MATH_*: 0 = fp32 CUDA-core kernels; 1..3 = tcgen05 kernels with that
                                    many bf16 terms per operand (the step becomes ONE fused launch:
                                    [x taps | h taps] x [K;R] GEMM + gate algebra + cell update) */
  int ws_prepacked;              /* 1: io.ws still holds this layer's packed weights from an earlier fov_convlstm_fwd
                                    call with the same weights and shapes (the 10 one-step decoder calls of
                                    mycode/convlstm_seq2seq.py:211-220): the repack launch is skipped */
  int wave_layers;               /* layer wavefront: number of stacked layers that will run concurrently (0 / 1: this
                                    layer has the GPU to itself).  The persistent kernels size their image groups so
                                    that the CTAs of all wave_layers layers fit on the SMs at once */
} fov_convlstm_cfg;

typedef struct {
  const float *x;
  const float *kernel, *recurrent, *bias;
  const float *h0, *c0;          /* optional (B,H,W,F) dense */
  float *hseq;                   /* strided output */
  float *gates;                  /* (B,T,H,W,4F): pre-activations then activated gates (saved) */
  float *cseq;                   /* (B,T,H,W,F) */
  float *hT, *cT;                /* optional dense (B,H,W,F) */
  float *ws;                     /* workspace of fov_convlstm_fwd_ws_bytes() bytes (math != 0) */
  /* Layer wavefront of stacked ConvLSTMs at small batches (optional, persistent tensor-core kernels only, see
   * fov_convlstm_wave_groups): the layers of a stack are launched on DIFFERENT streams and run concurrently; layer l
   * raises wave_set[group*T + t] once h_t of the group's images is in global memory and layer l+1 waits for
   * wave_wait[group*T + t] before it reads x_t.  int32 device arrays of fov_convlstm_wave_groups() * T zeros.  The
   * caller guarantees that all the stack's CTAs can be resident at once (sum of the groups <= number of SMs). */
  const int *wave_wait;
  int *wave_set;
} fov_convlstm_io;

size_t fov_convlstm_fwd_ws_bytes(const fov_convlstm_cfg* cfg);
/* image groups (= CTAs) of the persistent forward (backward = 0) / BPTT (backward = 1, with a fused input gradient when
 * with_dx) kernel for this configuration, 0 when the configuration does not take a persistent kernel with one group per
 * CTA (then the wavefront flags are not available).  The group of image b is b / (ceil(B / groups)). */
int fov_convlstm_wave_groups(const fov_convlstm_cfg* cfg, int backward, int with_dx);
/* 1 when fov_convlstm_fwd runs this configuration as ONE persistent launch over all timesteps (whole images per MMA
 * tile), 0 when it launches per timestep (then a caller may interleave the layers of a stack at launch level) */
int fov_convlstm_fwd_persistent(const fov_convlstm_cfg* cfg);
int fov_convlstm_fwd(const fov_convlstm_cfg* cfg, const fov_convlstm_io* io, void* stream);

typedef struct {
  const float *dhseq;            /* gradient w.r.t. hseq, same strides as hseq (may be NULL) */
  const float *dhT, *dcT;        /* optional gradient w.r.t. final state, dense */
  float *dx;                     /* optional, same strides as x (see dx_accumulate) */
  float *dh0, *dc0;              /* optional dense */
  float *g_kernel, *g_recurrent, *g_bias;   /* accumulated (+=) */
  float *ws;                     /* workspace floats: see fov_convlstm_bwd_ws_floats */
  int dx_accumulate;             /* 0: dx overwritten; 1: dx += (stacked layers add into the
                                    gradient of the layer below) */
  void *wgrad_stream;            /* optional cudaStream_t for the weight-gradient launches of the tensor-core path (see
                                    fov_lstm_grads.wgrad_stream).  NULL: `stream` */
  /* layer wavefront of the backward pass (see fov_convlstm_io.wave_wait): this layer's BPTT waits for wave_wait[group*T+t]
   * before it reads dhseq[t] (the layer above is still adding its dx into it) and raises wave_set[group*T+t] once its
   * own fused dx_t is in global memory.  wave_set needs the fused input gradient (dx != NULL on the persistent path). */
  const int *wave_wait;
  int *wave_set;
} fov_convlstm_grads;

size_t fov_convlstm_bwd_ws_floats(const fov_convlstm_cfg* cfg);
int fov_convlstm_bwd(const fov_convlstm_cfg* cfg, const fov_convlstm_io* io,
                     const fov_convlstm_grads* g, void* stream);

/* ------------------------------------------------------------------------- *
 * Pointwise / reduction kernels.
 * ------------------------------------------------------------------------- */
/* Softmax over the last (channel) axis: keras.layers.Softmax(axis=-1),
 * mycode/convlstm_seq2seq.py:237. */
int fov_softmax_fwd(long long rows, int C, const float* x, float* y, void* stream);
int fov_softmax_bwd(long long rows, int C, const float* y, const float* dy, float* dx, void* stream);

/* MSE (Keras 'mean_squared_error', mycode/FoV_seq2seq.py:103, mycode/cost.py:20-29):
 * loss[0] += weight * mean((y-t)^2);  dy = weight * 2 (y-t) / n  (dy may be NULL). */
int fov_mse_fwd_bwd(long long n, const float* y, const float* t, float weight, float* loss,
                    float* dy, void* stream);
/* Gaussian NLL, mycode/cost.py:138-187.  y (B,T,6), frames (B,T,90) interleaved xyz;
 * loss[0] += weight * mean_B(sum_{t,frame,axis} l) / running_length / 30. */
int fov_gauss_nll_fwd_bwd(int B, int T, int running_length, const float* y, const float* frames,
                          float weight, float* loss, float* dy, void* stream);
/* Keras categorical_crossentropy on probabilities (mycode/convlstm_heatmap.py:281). */
int fov_cce_fwd_bwd(long long rows, int C, const float* p, const float* t, float weight,
                    float* loss, float* dp, void* stream);

/* Keras-form Adam / RMSprop over one flat parameter buffer
 * ('Adam' mycode/FoV_seq2seq.py:103; 'RMSprop' mycode/convlstm_seq2seq.py:287).
 * g is first multiplied by grad_scale and, when grad_div is not NULL, divided by the DEVICE scalar
 * *grad_div (the summed sample count of a count-weighted data-parallel allreduce: no host sync). */
int fov_adam_step(long long n, float* p, const float* g, float* m, float* v, int t, float lr,
                  float beta1, float beta2, float eps, float grad_scale, const float* grad_div,
                  void* stream);
int fov_rmsprop_step(long long n, float* p, const float* g, float* a, float lr, float rho,
                     float eps, float grad_scale, const float* grad_div, void* stream);

/* Per-second mean / population variance featuriser (mycode/utility.py:483-517):
 * frames (rows,30,3) interleaved xyz -> (rows,6) = [mx,my,mz,vx,vy,vz]. */
int fov_mean_var_xyz(long long rows, const float* frames, float* out, void* stream);
/* Gaussian sample-and-refeed with explicit N(0,1) noise
 * (mycode/utility.py:73-80 mode 0; others_LSTM_span_whole.py:64-69 mode 1;
 *  convlstm_seq2seq.py:51-58 mode 2).  muvar (rows,6), noise (rows,30,3) -> (rows,30,3) */
int fov_gauss_resample(long long rows, int mode, const float* muvar, const float* noise,
                       float* out, void* stream);
/* Gradient of the re-sampler w.r.t. (mu, var) - K.random_normal(mean=mu, stddev=...) is differentiable in
 * its mean and stddev, so training graphs with cfg.sample_and_refeed back-propagate through the draw
 * (mycode/convlstm_seq2seq.py:259-272): dmuvar (rows,6) = [sum_f dout, sum_f dout*noise * d sd/d var]. */
int fov_gauss_resample_bwd(long long rows, int mode, const float* muvar, const float* noise,
                           const float* dout, float* dmuvar, void* stream);
/* Philox4x32-10 counter-based generator (the generator behind K.random_normal): element i of the stream is
 * word i%4 of philox(counter = offset + i/4, key = seed).  words (n) raw 32-bit outputs and/or normal (n)
 * Box-Muller N(0,1) draws; either may be NULL.  Results depend on (seed, offset) only - a rank draws the
 * rows it owns with offset = first_row * 90 / 4 and gets the numbers a single process would. */
int fov_philox_normal(long long n, unsigned long long seed, unsigned long long offset,
                      unsigned int* words, float* normal, void* stream);

/* ConvLSTM2D input dropout (keras ConvLSTM2D(dropout=0.3), mycode/others_LSTM_span_whole.py:89,
 * mycode/convlstm_seq2seq.py:100-126): one mask per gate, shape (B,H,W,Cin), constant over time, applied to the
 * layer input before the input convolution of that gate.  It is expressed with the unchanged ConvLSTM kernels by
 * widening the input: x4[b,t,p, g*Cin + c] = x[b,t,p,c] * mask[g,b,p,c] and a block kernel
 * K4[tap, g*Cin + c, n] = K[tap,c,n] if n is a column of gate g, else 0; the reduce calls fold the gradients back.
 * masks (4,B,HW,Cin) dense, already scaled by 1/(1-rate); x / dx (B,T,HW,Cin) with strides; x4 / dx4 dense. */
int fov_dropout_expand(int B, int T, int HW, int Cin, const float* x, long long x_b_stride, long long x_t_stride,
                       int x_pix_stride, const float* masks, float* x4, void* stream);
int fov_dropout_reduce(int B, int T, int HW, int Cin, const float* dx4, const float* masks, float* dx,
                       long long x_b_stride, long long x_t_stride, int x_pix_stride, int accumulate, void* stream);
int fov_gate_kernel_expand(int taps, int Cin, int F, const float* kernel, float* kernel4, void* stream);
/* g_kernel[tap,c,n] += g_kernel4[tap, gate(n)*Cin + c, n] */
int fov_gate_kernel_reduce(int taps, int Cin, int F, const float* g_kernel4, float* g_kernel, void* stream);

/* ------------------------------------------------------------------------- *
 * Sample builders just before the hot path.
 * ------------------------------------------------------------------------- */
/* reshape2second_stacks (mycode/utility.py:264-305): per-video seconds src (U,S,C) -> windows of L =
 * cfg.running_length seconds every `stride` seconds; future = the window L//stride rows later; future_input =
 * [last past second, future[:-1]].  purely_testing appends L//stride zero seconds (cfg.purelly_testing).
 * Outputs, n = fov_window_count(): collapse_user=1 (n*U, L, C) window-major; 0: (U, n, L, C). */
int fov_window_count(int S, int L, int stride, int purely_testing);
int fov_window_stacks(int U, int S, int C, int L, int stride, int purely_testing, int collapse_user,
                      const float* src, float* past, float* future, float* future_input, void* stream);
/* get_whole_span (mycode/others_LSTM_span_whole.py:403-419): x (N, half) rows (half = L * everything after the
 * time axis) -> out (N, 2*half): out[i] = [x[i] ; x[i+1]], the last row is all zero. */
int fov_whole_span(long long N, long long half, const float* x, float* out, void* stream);
/* get_data's target / others split (mycode/utility.py:389-430) applied to the windows of one video: every viewer is the
 * target once; for target t the others are viewers idx[t*K .. t*K+K) (the remaining viewers, padded with duplicates or
 * truncated to K = num_user - 1 by the caller's index list).  src (U, n, row) -> out (K, N_total, row) with
 * out[j][base_rows + t*n + w] = src[idx[t*K+j]][w]; out_j_stride = N_total * row floats.  idx is a DEVICE int32 array. */
int fov_pick_user_gather(int T, int K, int n, long long row, const int* idx, const float* src, float* out,
                         long long out_j_stride, long long base_rows, void* stream);
/* Training batches of the concat-state model from the per-second mean/var features mv (U,S,6) of one video (the data
 * preparation of mycode/others_LSTM_span_whole.py:403-419,640-668 + mycode/utility.py:389-430 composed): sequence
 * b = target viewer (b / n) x window (b % n), n = (S-20)/stride + 1, start second s0 = (b % n) * stride:
 *   enc (B,10,6) = mv[t, s0..s0+10); oth (B,20,1,K,6) = mv[idx[t*K+j], s0..s0+20); dec0 (B,1,6) = mv[t, s0+9];
 *   fut (B,10,6) = mv[t, s0+10..s0+20).  idx: DEVICE int32 (U,K) others of every target.  B <= U*n (truncation). */
int fov_m3_batches(int U, int S, int stride, int K, const int* idx, const float* mv, long long B, float* enc,
                   float* oth, float* dec0, float* fut, void* stream);
/* One-hot FoV-centre heatmaps: xyz (rows, frames, 3) -> out (rows, 360/bin, 180/bin, frames), one 1 per frame at
 * (theta bin, phi bin) with theta, phi of mycode/dataIO.py:77-82 and the binning of mycode/utility.py:533-539,
 * _create_one_hot :546-556; frames stacked as channels as mycode/data_generator_for_heatmap.py:32,65-67 feeds them.
 * Angles are evaluated in float64 so the bin indices equal NumPy's. */
int fov_onehot_heatmaps(long long rows, int frames, int bin_size, const float* xyz, float* out, void* stream);

/* Gaussian-FoV / head-direction tiles (mycode/data_generator_gaussian_FoV.py): the per-second heat maps of the
 * reference's Gaussian-FoV data generator, from frame centres.
 * fov_theta_phi_frames: xyz (frames,3) -> phi_theta (frames,2) float64 = [phi/pi, (theta+pi)/2/pi] (:21-55 on
 *   xyz2thetaphi, mycode/dataIO.py:77-82).
 * fov_gaussian_fov_tiles: phi_theta (maps*frames, 2) float64 centres in [0,1] -> out (maps, 18, 36, frames) float32;
 *   kind 0 = crop_FoV_from_equirect / get_gaussian_FoV (:57-127), kind 1 = blur_head_direction_equirect /
 *   get_head_direction (:163-232); every 10th row / column of the 180 x 360 map of each frame, divided by the maximum
 *   over the FULL-resolution maps of every frame of the call (as the reference normalises).  peak: one device float of
 *   scratch that returns that maximum.  Frames of a map are its channels (:130-138, :235-243).
 * fov_heatmap_sum: tiles (maps, cells, frames) -> out (maps, cells): frame channels summed per cell and each map scaled
 *   to sum 1 (heatmap_sum + normalize_to_distribution, :246-261), float32 in NumPy's summation order. */
int fov_theta_phi_frames(long long frames, const float* xyz, double* phi_theta, void* stream);
int fov_gaussian_fov_tiles(long long maps, int frames, int kind, const double* phi_theta, float* out, float* peak,
                           void* stream);
int fov_heatmap_sum(long long maps, int cells, int frames, const float* tiles, float* out, void* stream);

/* FoV hit rate (evaluation metric, mycode/baseline_knn_mean.py:48-93,123-168): pred / gt (rows,2) = (theta, phi)
 * centres in radians, spans in radians (the scripts use 120 x 120 degrees); out[i] = overlap of the two boxes over the
 * ground-truth box area after the +-pi wrap fix of boundary_cases. */
int fov_hit_rate(long long rows, const float* pred, const float* gt, float span_theta, float span_phi,
                 float gt_span_theta, float gt_span_phi, float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Data-parallel training (the reference is single-process): one process per GPU, one NCCL
 * communicator per process, one summed allreduce of the flat fp32 gradient bucket per step over NVLink.
 * The communicator handle is the only state the library keeps.  Rank 0 calls fov_dp_get_unique_id, the
 * caller ships the fov_dp_unique_id_bytes() bytes to every rank (any transport), every rank calls fov_dp_init.
 * libnccl.so.2 is bound at run time; without it these calls return FOV_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------- */
int fov_dp_unique_id_bytes(void);
int fov_dp_get_unique_id(void* id_out);
int fov_dp_init(const void* unique_id, int rank, int world);
int fov_dp_world(void);
int fov_dp_rank(void);
/* in-place sum over all ranks / copy from root, asynchronous on `stream` */
int fov_dp_allreduce(float* flat, size_t n, void* stream);
int fov_dp_broadcast(float* flat, size_t n, int root, void* stream);
int fov_dp_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* FOV360_H */
