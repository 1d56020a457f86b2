// Sample builders that sit just before the hot path: sliding-window stacks, whole-span
// pairing and one-hot FoV-centre heatmaps, as HBM-bound gather / scatter kernels.  The reference runs these as Python
// double loops over NumPy arrays (mycode/utility.py:264-305, mycode/others_LSTM_span_whole.py:403-419,
// mycode/utility.py:519-571); here one launch builds every window of a video from the per-second tensor that already
// lives in HBM, so the host feed disappears from the step.
#include "fov_common.cuh"

namespace {

// One thread per VEC consecutive output elements; the three outputs share one layout, so every store is coalesced and
// every load is a contiguous piece of a source row.
template <int VEC>
__global__ void __launch_bounds__(256) window_stacks_kernel(int U, int S, int C, int L, int stride, int shift, int n,
                                                            int collapse, const float* __restrict__ src,
                                                            float* __restrict__ past, float* __restrict__ fut,
                                                            float* __restrict__ futin, long long total_vec) {
  const int CV = C / VEC;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total_vec;
       e += (long long)gridDim.x * blockDim.x) {
    const int kv = (int)(e % CV);
    long long r = e / CV;
    const int j = (int)(r % L);
    r /= L;
    int i, u;
    if (collapse) { u = (int)(r % U); i = (int)(r / U); }
    else { i = (int)(r % n); u = (int)(r / n); }
    const int sp = stride * i + j;                                  // second of the past window
    const int sf = stride * (i + shift) + j;                        // second of the future window
    const int si = j == 0 ? stride * i + L - 1 : sf - 1;            // decoder input: last past second, then future shifted
    const float* row = src + (long long)u * S * C + (long long)kv * VEC;
    float vp[VEC], vf[VEC], vi[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) { vp[q] = 0.0f; vf[q] = 0.0f; vi[q] = 0.0f; }
    // seconds >= S only exist as the zero tail appended in purely-testing mode
    if (VEC == 2) {
      if (sp < S) { const float2 a = __ldg(reinterpret_cast<const float2*>(row + (long long)sp * C)); vp[0] = a.x; vp[VEC - 1] = a.y; }
      if (sf < S) { const float2 a = __ldg(reinterpret_cast<const float2*>(row + (long long)sf * C)); vf[0] = a.x; vf[VEC - 1] = a.y; }
      if (si < S) { const float2 a = __ldg(reinterpret_cast<const float2*>(row + (long long)si * C)); vi[0] = a.x; vi[VEC - 1] = a.y; }
      reinterpret_cast<float2*>(past)[e] = make_float2(vp[0], vp[VEC - 1]);
      reinterpret_cast<float2*>(fut)[e] = make_float2(vf[0], vf[VEC - 1]);
      reinterpret_cast<float2*>(futin)[e] = make_float2(vi[0], vi[VEC - 1]);
    } else {
      if (sp < S) vp[0] = __ldg(row + (long long)sp * C);
      if (sf < S) vf[0] = __ldg(row + (long long)sf * C);
      if (si < S) vi[0] = __ldg(row + (long long)si * C);
      past[e] = vp[0]; fut[e] = vf[0]; futin[e] = vi[0];
    }
  }
}

// out[i] = [x[i] ; x[i+1]] for i < N-1, zeros for the last row (float4 or scalar elements)
template <typename T>
__global__ void __launch_bounds__(256) whole_span_kernel(long long N, long long half, const T* __restrict__ x,
                                                          T* __restrict__ out, long long total) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / (2 * half), k = e - i * 2 * half;
    // [x[i] ; x[i+1]] is the contiguous range x[i*half .. (i+2)*half)
    T v{};
    if (i + 1 < N) v = x[i * half + k];
    out[e] = v;
  }
}

// One CTA per (second) row: bins of its frames in shared memory, then the whole (360/bin, 180/bin, frames) block is
// written with coalesced float4 stores (zeros and ones in one pass: no memset + scatter).
__global__ void __launch_bounds__(256) onehot_heatmap_kernel(long long rows, int F, int nth, int nph, double bin_deg,
                                                              const float* __restrict__ xyz, float* __restrict__ out) {
  extern __shared__ int bin_s[];
  const long long row = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float* p = xyz + (row * F + f) * 3;
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    const double kPi = 3.141592653589793;
    // mycode/dataIO.py:77-82 (np.mod: result carries the sign of the divisor)
    double a = atan2(y, x);
    double m = fmod(a, 2.0 * kPi);
    if (m < 0.0) m += 2.0 * kPi;
    const double theta = m - kPi;
    double b = atan2(z, sqrt(x * x + y * y)) + kPi / 2.0;
    double phi = fmod(b, kPi);
    if (phi < 0.0) phi += kPi;
    // mycode/utility.py:533-539
    int ti = (int)floor((theta + kPi) / kPi * 180.0 / bin_deg);
    if (ti == nth) ti -= 1;
    int pi = (int)floor(phi / kPi * 180.0 / bin_deg);
    if (pi == nph) pi -= 1;
    bin_s[f] = ti * nph + pi;
  }
  __syncthreads();
  const int cells = nth * nph;
  const long long per_row = (long long)cells * F;
  float* o = out + row * per_row;
  if ((per_row & 3) == 0) {
    for (int e4 = threadIdx.x; e4 < per_row / 4; e4 += blockDim.x) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = e4 * 4 + q, cell = e / F, f = e - cell * F;
        v[q] = bin_s[f] == cell ? 1.0f : 0.0f;
      }
      reinterpret_cast<float4*>(o)[e4] = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (int e = threadIdx.x; e < per_row; e += blockDim.x) {
      const int cell = e / F, f = e - cell * F;
      o[e] = bin_s[f] == cell ? 1.0f : 0.0f;
    }
  }
}

// FoV hit rate of one (predicted, ground-truth) centre pair: overlap of the two (theta, phi) boxes over the ground-truth
// box area, after the +-pi wrap fix of boundary_cases (mycode/baseline_knn_mean.py:48-93).  float64 like the reference.
__global__ void __launch_bounds__(256) hit_rate_kernel(long long rows, const float* __restrict__ pred,
                                                        const float* __restrict__ gt, double span_t, double span_p,
                                                        double gspan_t, double gspan_p, float* __restrict__ out) {
  const double kPi = 3.141592653589793;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    double ct = (double)pred[2 * i], cp = (double)pred[2 * i + 1];
    double gtt = (double)gt[2 * i], gtp = (double)gt[2 * i + 1];
    if (gtt > 2 / 3.0 * kPi && ct < -2 / 3.0 * kPi) ct += 2 * kPi;
    if (gtt < -2 / 3.0 * kPi && ct > 2 / 3.0 * kPi) gtt += 2 * kPi;
    const double b0 = ct - span_t / 2, b1 = cp - span_p / 2, b2 = ct + span_t / 2, b3 = cp + span_p / 2;
    const double q0 = gtt - gspan_t / 2, q1 = gtp - gspan_p / 2, q2 = gtt + gspan_t / 2, q3 = gtp + gspan_p / 2;
    const double area = (q2 - q0) * (q3 - q1);
    double ov = 0.0;
    const double iw = fmin(b2, q2) - fmax(b0, q0);
    if (iw > 0) {
      const double ih = fmin(b3, q3) - fmax(b1, q1);
      if (ih > 0) ov = iw * ih / area;
    }
    out[i] = (float)ov;
  }
}

int grid_for(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  const long long cap = (long long)fov_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int fov_window_count(int S, int L, int stride, int purely_testing) {
  if (L <= 0 || stride <= 0 || S < 2 * L) return 0;
  const int shift = L / stride;
  const int Sp = S + (purely_testing ? shift : 0);
  const int nrows = (Sp - L) / stride + 1;
  return nrows - shift > 0 ? nrows - shift : 0;
}

extern "C" int fov_window_stacks(int U, int S, int C, int L, int stride, int purely_testing, int collapse_user,
                                 const float* src, float* past, float* future, float* future_input, void* stream) {
  FOV_CHECK_ARG(U > 0 && S > 0 && C > 0 && L > 0 && stride > 0, "bad shape");
  FOV_CHECK_ARG(S >= 2 * L, "needs at least 2 x running_length seconds (mycode/utility.py:271)");
  FOV_CHECK_ARG(L / stride >= 1, "stride must not exceed running_length (shift = running_length // stride)");
  FOV_CHECK_ARG(src && past && future && future_input, "NULL pointer");
  const int shift = L / stride;
  const int n = fov_window_count(S, L, stride, purely_testing);
  FOV_CHECK_ARG(n > 0, "no complete window");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * U * L * C;
  auto a8 = [](const void* p) { return (uintptr_t)p % 8 == 0; };
  if (C % 2 == 0 && a8(src) && a8(past) && a8(future) && a8(future_input)) {
    const long long tv = total / 2;
    window_stacks_kernel<2><<<grid_for(tv, 256), 256, 0, st>>>(U, S, C, L, stride, shift, n, collapse_user, src, past,
                                                               future, future_input, tv);
  } else {
    window_stacks_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(U, S, C, L, stride, shift, n, collapse_user, src, past,
                                                                  future, future_input, total);
  }
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_whole_span(long long N, long long half, const float* x, float* out, void* stream) {
  FOV_CHECK_ARG(N > 0 && half > 0 && x && out, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  auto a16 = [](const void* p) { return (uintptr_t)p % 16 == 0; };
  if (half % 4 == 0 && a16(x) && a16(out)) {
    const long long total = N * 2 * (half / 4);
    whole_span_kernel<float4><<<grid_for(total, 256), 256, 0, st>>>(N, half / 4, reinterpret_cast<const float4*>(x),
                                                                   reinterpret_cast<float4*>(out), total);
  } else {
    const long long total = N * 2 * half;
    whole_span_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(N, half, x, out, total);
  }
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

// get_data's target / others split (mycode/utility.py:389-430) on windowed tensors: every viewer of a video is the target
// once; for target t the others are the viewers idx[t][0..K) (the rest of the video, padded with duplicates or truncated
// to num_user - 1).  src (U, n, row) windows of one video; out[j][base + t*n + w][:] = src[idx[t*K + j]][w][:].
template <typename V>
__global__ void __launch_bounds__(256) pick_user_gather_kernel(int n, long long rowv, int K, const int* __restrict__ idx,
                                                               const V* __restrict__ src, V* __restrict__ out,
                                                               long long out_j_stride_v, long long base_rows,
                                                               long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % rowv;
    long long r = i / rowv;
    const int w = (int)(r % n); r /= n;
    const int j = (int)(r % K);
    const long long t = r / K;
    const int u = __ldg(&idx[t * K + j]);
    out[(long long)j * out_j_stride_v + (base_rows + t * n + w) * rowv + e] = __ldg(&src[((long long)u * n + w) * rowv + e]);
  }
}

extern "C" int fov_pick_user_gather(int T, int K, int n, long long row, const int* idx, const float* src, float* out,
                                    long long out_j_stride, long long base_rows, void* stream) {
  FOV_CHECK_ARG(T > 0 && K > 0 && n > 0 && row > 0 && idx && src && out && out_j_stride >= 0 && base_rows >= 0,
                "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  auto a16 = [](const void* p) { return (uintptr_t)p % 16 == 0; };
  if (row % 4 == 0 && out_j_stride % 4 == 0 && a16(src) && a16(out)) {
    const long long total = (long long)T * K * n * (row / 4);
    pick_user_gather_kernel<float4><<<grid_for(total, 256), 256, 0, st>>>(
        n, row / 4, K, idx, reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(out), out_j_stride / 4,
        base_rows, total);
  } else {
    const long long total = (long long)T * K * n * row;
    pick_user_gather_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(n, row, K, idx, src, out, out_j_stride, base_rows,
                                                                          total);
  }
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

// Training batches of the concat-state model straight from the per-second mean/var features of one video
// (mycode/others_LSTM_span_whole.py:403-419,640-668 with mycode/utility.py:389-430,483-517 composed): sequence
// b = (target viewer t = b / n, window w = b % n) starts at second s0 = w * stride and gets
//   enc  (B,10,6)        = mv[t, s0 .. s0+10)                     (target past; also the reconstruction target)
//   oth  (B,20,1,K,6)    = mv[idx[t*K+j], s0 .. s0+20)            (others' whole span; also the others target)
//   dec0 (B,1,6)         = mv[t, s0+9]                            (last observed second)
//   fut  (B,10,6)        = mv[t, s0+10 .. s0+20)                  (target future)
// One thread per output float, outputs coalesced; the feature table (U*S*6 floats) stays in L2.
__global__ void __launch_bounds__(256) m3_batch_kernel(int S, int n, int stride, int K, const int* __restrict__ idx,
                                                       const float* __restrict__ mv, long long B, float* __restrict__ enc,
                                                       float* __restrict__ oth, float* __restrict__ dec0,
                                                       float* __restrict__ fut) {
  const int per_oth = 20 * K * 6, per = per_oth + 126;
  const long long total = B * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per;
    const int e = (int)(i - b * per);
    const int t = (int)(b / n), s0 = (int)(b % n) * stride;
    if (e < per_oth) {
      const int c = e % 6, j = (e / 6) % K, s = e / (6 * K);
      const int u = __ldg(&idx[(long long)t * K + j]);
      oth[b * per_oth + e] = __ldg(&mv[((long long)u * S + s0 + s) * 6 + c]);
    } else if (e < per_oth + 60) {
      const int q = e - per_oth;
      enc[b * 60 + q] = __ldg(&mv[((long long)t * S + s0) * 6 + q]);
    } else if (e < per_oth + 120) {
      const int q = e - per_oth - 60;
      fut[b * 60 + q] = __ldg(&mv[((long long)t * S + s0 + 10) * 6 + q]);
    } else {
      const int q = e - per_oth - 120;
      dec0[b * 6 + q] = __ldg(&mv[((long long)t * S + s0 + 9) * 6 + q]);
    }
  }
}

extern "C" int fov_m3_batches(int U, int S, int stride, int K, const int* idx, const float* mv, long long B, float* enc,
                              float* oth, float* dec0, float* fut, void* stream) {
  FOV_CHECK_ARG(U > 0 && S >= 20 && stride > 0 && K > 0 && idx && mv && enc && oth && dec0 && fut, "bad arguments");
  const int n = (S - 20) / stride + 1;                       // windows whose 20 seconds exist
  FOV_CHECK_ARG(B > 0 && B <= (long long)U * n, "B exceeds viewers x windows");
  const long long total = B * (20LL * K * 6 + 126);
  m3_batch_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(S, n, stride, K, idx, mv, B, enc, oth, dec0, fut);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_onehot_heatmaps(long long rows, int frames, int bin_size, const float* xyz, float* out, void* stream) {
  FOV_CHECK_ARG(rows > 0 && rows < (1LL << 31) && frames > 0 && frames <= 4096, "bad shape");
  FOV_CHECK_ARG(bin_size > 0 && 360 % bin_size == 0 && 180 % bin_size == 0, "bin_size must divide 180");
  FOV_CHECK_ARG(xyz && out, "NULL pointer");
  FOV_CHECK_ARG((uintptr_t)out % 16 == 0, "out must be 16-byte aligned");
  const int nth = 360 / bin_size, nph = 180 / bin_size;
  onehot_heatmap_kernel<<<(unsigned)rows, 256, frames * sizeof(int), (cudaStream_t)stream>>>(rows, frames, nth, nph,
                                                                                           (double)bin_size, xyz, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_hit_rate(long long rows, const float* pred, const float* gt, float span_theta, float span_phi,
                            float gt_span_theta, float gt_span_phi, float* out, void* stream) {
  FOV_CHECK_ARG(rows > 0 && pred && gt && out, "bad arguments");
  FOV_CHECK_ARG(gt_span_theta > 0 && gt_span_phi > 0, "ground-truth span must be positive");
  hit_rate_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(rows, pred, gt, (double)span_theta, (double)span_phi,
                                                                        (double)gt_span_theta, (double)gt_span_phi, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
