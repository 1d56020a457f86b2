/*
 * fov_debug.h — diagnostic switches and in-kernel timelines of libfov360.so.
 *
 * NOT part of the drop-in boundary (include/fov360.h): nothing on the product path calls
 * these.  They exist for the A/B parity tests (tests/test_gpu_parity.py compares the
 * persistent kernels with the per-timestep launches they replace, and the tensor-core
 * fc-LSTM with the fp32 one) and for the bring-up harness tests/cuda/tc_selftest.cu.
 *
 * The switches are process-global, not thread safe, and default to "choose automatically";
 * they are the ONLY mutable state in the library besides the NCCL communicator of
 * fov_dp_init() and the launch counter.  No environment variable changes kernel selection.
 */
#ifndef FOV_DEBUG_H
#define FOV_DEBUG_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- kernel selection (A/B tests) ---- */
/* fc-LSTM forward: -1 = never the tcgen05 kernel, 0 = choose by shape (default), 1 = whenever the shape allows */
void fov_debug_lstm_tc(int mode);
/* fc-LSTM weight gradients as one tcgen05 launch per LSTM (default 1) or the SIMT kernels (0) */
void fov_debug_lstm_wgrad_tc(int on);
/* fc-LSTM BPTT on tcgen05 when the shape allows (default 1) or the fp32 kernel (0) */
void fov_debug_lstm_bptt_tc(int on);
/* time-batched input projection (xproj GEMM fed by TMA) in front of the tcgen05 fc-LSTM forward:
 * -1 = never, 0 = choose by shape (default: inputs wider than 16 features), 1 = whenever the shape allows */
void fov_debug_lstm_xproj(int mode);
/* persistent ConvLSTM forward / BPTT / fused weight gradient (default 1) or the per-timestep launches (0) */
void fov_debug_convlstm_persistent(int enable);
void fov_debug_convlstm_persistent_bwd(int enable);
void fov_debug_wgrad_rows(int enable);
/* ConvLSTM weight gradient accumulated inside the persistent BPTT kernel (1) or by the separate fused launch (0, the
 * default: measured faster, see convlstm_seq_bwd_tc.cu) */
void fov_debug_seq_bwd_wgrad(int enable);
/* persistent BPTT without the stacked-N operand layout (bring-up) */
void fov_debug_seq_bwd_nostack(int on);
/* worker warps per image group of the persistent ConvLSTM / fc-LSTM forward (0 = default) */
void fov_debug_seq_wpg(int wpg);
/* persistent ConvLSTM kernels at small batches: 1 (default) = as few images per CTA as still fills the SMs, one group
 * per CTA; 0 = always full 128-row groups */
void fov_debug_seq_spread(int on);
/* fp32 fc-LSTM kernels at small batches: 1 (default) = 4 / 8 sequences per CTA while the batch leaves SMs idle */
void fov_debug_lstm_small_tiles(int on);
/* fused ConvLSTM weight gradient: 1 = one tile per CTA even with few tiles (default 0: at least 4 tiles per CTA) */
void fov_debug_wgrad_rows_full_grid(int on);
void fov_debug_lstm_tc_wpg(int wpg);
/* wide single-term (bf16) convolutions with two 128-row accumulator tiles per CTA: 0 never, 1 when the grid still
 * fills the machine (default), 2 whenever the tile fits */
void fov_debug_conv_mt2(int mode);
/* TMA-fed weight gradient of wide k x k convolutions (default 1) or the general gather kernel (0) */
void fov_debug_wgrad_planes(int enable);
/* general weight gradient: force the 128-wide M tile / the narrow path (bring-up) */
void fov_debug_wgrad_single_m(int on);
void fov_debug_wgrad_narrow(int on);

/* ---- in-kernel clock64 timelines: enable, run the kernel, read the counters of CTA 0 ---- */
void fov_debug_timeline_enable(int on);                       /* tc_conv_kernel: 8 words per CTA, 256 CTAs */
int fov_debug_timeline_read(unsigned long long* out, int n_words);
void fov_debug_seq_enable(int on);                            /* convlstm_seq_fwd_kernel: 16 words */
int fov_debug_seq_read(unsigned long long* out);
void fov_debug_seq_bwd_enable(int on);                        /* convlstm_seq_bwd_kernel: 8 words */
int fov_debug_seq_bwd_read(unsigned long long* out);
void fov_debug_wgrad_rows_timeline(int on);                   /* tc_wgrad_rows_kernel: 8 words */
int fov_debug_wgrad_rows_read(unsigned long long* out);
void fov_debug_wgrad_enable(int on);                          /* tc_wgrad_kernel: 8 words */
int fov_debug_wgrad_read(unsigned long long* out);
void fov_debug_lstm_tc_enable(int on);                        /* lstm_tc_fwd_kernel: 8 words */
int fov_debug_lstm_tc_read(unsigned long long* out);

#ifdef __cplusplus
}
#endif
#endif /* FOV_DEBUG_H */
