// Shared device/host helpers for libfov360 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fov360.h"

void fov_set_error(const char* fmt, ...);
// number of kernel launches issued by this library since load (diagnostics only)
extern unsigned long long g_fov_launches;

#define FOV_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      fov_set_error("%s: %s", __func__, msg);                      \
      return FOV_ERR_ARG;                                          \
    }                                                              \
  } while (0)

#define FOV_CUDA_LAUNCH_CHECK()                                              \
  do {                                                                       \
    ++g_fov_launches;                                                        \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      fov_set_error("%s: CUDA error %s", __func__, cudaGetErrorString(e__)); \
      return FOV_ERR_CUDA;                                                   \
    }                                                                        \
  } while (0)

__device__ __forceinline__ float fov_hard_sigmoid(float x) {
  return fminf(fmaxf(0.2f * x + 0.5f, 0.0f), 1.0f);
}
__device__ __forceinline__ float fov_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int REC>
__device__ __forceinline__ float fov_rec_act(float x) {
  return REC == FOV_REC_HARD_SIGMOID ? fov_hard_sigmoid(x) : fov_sigmoid(x);
}
// derivative of the recurrent activation expressed through its OUTPUT a
template <int REC>
__device__ __forceinline__ float fov_rec_act_grad(float a) {
  if (REC == FOV_REC_HARD_SIGMOID) return (a > 0.0f && a < 1.0f) ? 0.2f : 0.0f;
  return a * (1.0f - a);
}
__device__ __forceinline__ float fov_rec_act_rt(int rec, float x) {
  return rec == FOV_REC_HARD_SIGMOID ? fov_hard_sigmoid(x) : fov_sigmoid(x);
}
__device__ __forceinline__ float fov_rec_act_grad_rt(int rec, float a) {
  if (rec == FOV_REC_HARD_SIGMOID) return (a > 0.0f && a < 1.0f) ? 0.2f : 0.0f;
  return a * (1.0f - a);
}
__device__ __forceinline__ float fov_act(int act, float x) {
  if (act == FOV_ACT_TANH) return tanhf(x);
  if (act == FOV_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float fov_act_grad(int act, float y) {
  if (act == FOV_ACT_TANH) return 1.0f - y * y;
  if (act == FOV_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  return 1.0f;
}

__device__ __forceinline__ void fov_cp_async4(float* smem_dst, const float* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void fov_cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void fov_cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// row width of the saved [h_{t-1} | x_t | 0] tensor of an fc-LSTM (include/fov360.h fov_lstm_saved.xh)
__host__ __device__ inline int fov_lstm_xh_stride(int H, int in_dim) { return (H + in_dim + 3) / 4 * 4; }

// Per-device one-time setup flag (kernel attributes are per device: a second GPU in the same process needs its own
// cudaFuncSetAttribute call).  Racing first calls from two host threads repeat the idempotent setup, nothing more.
struct FovPerDevice {
  static constexpr int kMaxDev = 64;
  bool flag[kMaxDev] = {};
  static int cur() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
  }
  bool done() const {
    const int d = cur();
    return d >= 0 && d < kMaxDev && flag[d];
  }
  void mark() {
    const int d = cur();
    if (d >= 0 && d < kMaxDev) flag[d] = true;
  }
};

static inline int fov_num_sms() {
  static int n[FovPerDevice::kMaxDev] = {};
  const int dev = FovPerDevice::cur();
  if (dev < 0 || dev >= FovPerDevice::kMaxDev) return 148;
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// Fork helper for launches that nothing downstream on `main` reads (weight gradients): `side` waits for everything
// enqueued on `main` so far.  Graph-capture safe (event record / wait is how a captured fork is expressed).
static inline int fov_fork_stream(cudaStream_t main, cudaStream_t side) {
  cudaEvent_t ev;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return -1;
  int rc = 0;
  if (cudaEventRecord(ev, main) != cudaSuccess || cudaStreamWaitEvent(side, ev, 0) != cudaSuccess) rc = -1;
  cudaEventDestroy(ev);          // released once the recorded work has completed
  return rc;
}
