// Pointwise / reduction kernels of the hot path: ConvLSTM gate update and its
// backward, channel softmax, the three losses fused with their gradients, the
// Keras-form Adam / RMSprop steps over one flat parameter buffer, the per-second
// mean/variance featuriser and the Gaussian re-sampler.  All HBM-bound.
#include "fov_common.cuh"
#include "fov_internal.h"

namespace {

inline int grid_for(long long n, int block = 256, int waves = 8) {
  long long b = (n + block - 1) / block;
  const long long cap = (long long)waves * fov_num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ float block_reduce_sum(float v) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}

// ---------------------------- ConvLSTM gates ------------------------------ //

__global__ void __launch_bounds__(256) convlstm_gates_fwd_kernel(GatesFwdArgs a) {
  const long long total = a.npix * a.F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx / a.F;
    const int f = (int)(idx - pix * a.F);
    const long long n = pix / a.HW;
    const long long p = pix - n * a.HW;
    float* zp = a.z + n * a.z_img + p * 4 * a.F + f;
    const float ig = fov_rec_act_rt(a.rec, zp[0]);
    const float fg = fov_rec_act_rt(a.rec, zp[a.F]);
    const float gg = tanhf(zp[2 * a.F]);
    const float og = fov_rec_act_rt(a.rec, zp[3 * a.F]);
    const float cp = a.c_prev ? a.c_prev[n * a.cp_img + p * a.F + f] : 0.0f;
    const float c = fmaf(fg, cp, ig * gg);
    const float h = og * tanhf(c);
    zp[0] = ig; zp[a.F] = fg; zp[2 * a.F] = gg; zp[3 * a.F] = og;
    a.c_out[n * a.c_img + p * a.F + f] = c;
    a.h_out[n * a.h_img + p * a.h_pix + f] = h;
    if (a.hT) a.hT[pix * a.F + f] = h;
    if (a.cT) a.cT[pix * a.F + f] = c;
  }
}

__global__ void __launch_bounds__(256) convlstm_gates_bwd_kernel(GatesBwdArgs a) {
  const long long total = a.npix * a.F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx / a.F;
    const int f = (int)(idx - pix * a.F);
    const long long n = pix / a.HW;
    const long long p = pix - n * a.HW;
    float* zp = a.gates + n * a.z_img + p * 4 * a.F + f;
    const float ig = zp[0], fg = zp[a.F], gg = zp[2 * a.F], og = zp[3 * a.F];
    const float ct = a.c_t[n * a.c_img + p * a.F + f];
    const float cp = a.c_prev ? a.c_prev[n * a.cp_img + p * a.F + f] : 0.0f;
    float dh = a.dh_rec ? a.dh_rec[pix * a.F + f] : 0.0f;
    if (a.dh_ext) dh += a.dh_ext[n * a.dhe_img + p * a.dhe_pix + f];
    const float dcin = a.dc_in ? a.dc_in[pix * a.F + f] : 0.0f;
    const float tc = tanhf(ct);
    const float dog = dh * tc;
    const float dct = fmaf(dh * og, 1.0f - tc * tc, dcin);
    a.dc_out[pix * a.F + f] = dct * fg;
    zp[0] = dct * gg * fov_rec_act_grad_rt(a.rec, ig);
    zp[a.F] = dct * cp * fov_rec_act_grad_rt(a.rec, fg);
    zp[2 * a.F] = dct * ig * (1.0f - gg * gg);
    zp[3 * a.F] = dog * fov_rec_act_grad_rt(a.rec, og);
  }
}

__global__ void mul_mask_kernel(long long n_img, long long img_elems, int C, const float* __restrict__ x,
                                long long x_img, int x_pix, const float* __restrict__ mask,
                                float* __restrict__ out) {
  const long long total = n_img * img_elems;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / img_elems;
    const long long r = idx - n * img_elems;
    const long long p = r / C;
    const int c = (int)(r - p * C);
    out[idx] = x[n * x_img + p * x_pix + c] * mask[idx];
  }
}

// ------------------------------ softmax ----------------------------------- //

__global__ void softmax_fwd_kernel(long long rows, int C, const float* __restrict__ x, float* __restrict__ y) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float* xr = x + r * C;
    float* yr = y + r * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, xr[c]);
    float s = 0.0f;
    for (int c = 0; c < C; ++c) { const float e = expf(xr[c] - m); yr[c] = e; s += e; }
    const float inv = 1.0f / s;
    for (int c = 0; c < C; ++c) yr[c] *= inv;
  }
}

__global__ void softmax_bwd_kernel(long long rows, int C, const float* __restrict__ y,
                                   const float* __restrict__ dy, float* __restrict__ dx) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float* yr = y + r * C;
    const float* gr = dy + r * C;
    float dot = 0.0f;
    for (int c = 0; c < C; ++c) dot = fmaf(yr[c], gr[c], dot);
    for (int c = 0; c < C; ++c) dx[r * C + c] = yr[c] * (gr[c] - dot);
  }
}

// ------------------------------- losses ----------------------------------- //

__global__ void __launch_bounds__(256) mse_kernel(long long n, const float* __restrict__ y,
                                                  const float* __restrict__ t, float wscale,
                                                  float* __restrict__ loss, float* __restrict__ dy) {
  float acc = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float d = y[i] - t[i];
    acc = fmaf(d, d, acc);
    if (dy) dy[i] = 2.0f * wscale * d;
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * wscale);
}

__global__ void __launch_bounds__(256) gauss_nll_kernel(long long rows, const float* __restrict__ y,
                                                        const float* __restrict__ frames, float scale,
                                                        float* __restrict__ loss, float* __restrict__ dy) {
  // one thread per (b,t,axis)
  const float eps = 1e-7f;
  float acc = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * 3;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 3;
    const int a = (int)(i - r * 3);
    const float u = y[r * 6 + a], vr = y[r * 6 + 3 + a];
    const float av = fabsf(vr);
    const float v = fminf(fmaxf(av, 1e-4f), 2.0f);
    const float ve = v + eps, lg = logf(ve), inv = 1.0f / ve;
    float du = 0.0f, dv = 0.0f;
    const float* fr = frames + r * 90 + a;
    for (int f = 0; f < 30; ++f) {
      const float d = fr[3 * f] - u;
      const float l = lg + d * d * inv;
      acc += fminf(fmaxf(l, -2000.0f), 2000.0f);
      if (l >= -2000.0f && l <= 2000.0f) {
        du += -2.0f * d * inv;
        dv += inv - d * d * inv * inv;
      }
    }
    if (dy) {
      const float dvr = (av >= 1e-4f && av <= 2.0f) ? (vr > 0.0f ? dv : (vr < 0.0f ? -dv : 0.0f)) : 0.0f;
      dy[r * 6 + a] = du * scale;
      dy[r * 6 + 3 + a] = dvr * scale;
    }
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * scale);
}

__global__ void __launch_bounds__(256) cce_kernel(long long rows, int C, const float* __restrict__ p,
                                                  const float* __restrict__ t, float scale,
                                                  float* __restrict__ loss, float* __restrict__ dp) {
  const float eps = 1e-7f;
  float acc = 0.0f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float* pr = p + r * C;
    const float* tr = t + r * C;
    float S = 0.0f;
    for (int c = 0; c < C; ++c) S += pr[c];
    const float invS = 1.0f / S;
    float gq = 0.0f;   // sum_c g_c q_c
    for (int c = 0; c < C; ++c) {
      const float q = pr[c] * invS;
      const float qc = fminf(fmaxf(q, eps), 1.0f - eps);
      acc -= tr[c] * logf(qc);
      if (q >= eps && q <= 1.0f - eps) gq += -tr[c];       // g_c*q_c = -t_c/q_c*q_c
    }
    if (dp) {
      for (int c = 0; c < C; ++c) {
        const float q = pr[c] * invS;
        const float g = (q >= eps && q <= 1.0f - eps) ? -tr[c] / q : 0.0f;
        dp[r * C + c] = scale * invS * (g - gq);
      }
    }
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * scale);
}

// ----------------------------- optimisers --------------------------------- //

__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr_t,
                                                   float b1, float b2, float eps, float gs,
                                                   const float* __restrict__ gdiv) {
  if (gdiv) gs = gs / __ldg(gdiv);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void __launch_bounds__(256) rmsprop_kernel(long long n, float* __restrict__ p,
                                                      const float* __restrict__ g, float* __restrict__ a,
                                                      float lr, float rho, float eps, float gs,
                                                      const float* __restrict__ gdiv) {
  if (gdiv) gs = gs / __ldg(gdiv);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float ai = rho * a[i] + (1.0f - rho) * gi * gi;
    a[i] = ai;
    p[i] = p[i] - lr * gi / (sqrtf(ai) + eps);
  }
}

// ------------------------- featuriser / re-sampler ------------------------ //

__global__ void mean_var_kernel(long long rows, const float* __restrict__ frames, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * 3;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 3;
    const int a = (int)(i - r * 3);
    const float* fr = frames + r * 90 + a;
    float s = 0.0f;
    for (int f = 0; f < 30; ++f) s += fr[3 * f];
    const float mean = s * (1.0f / 30.0f);
    float q = 0.0f;
    for (int f = 0; f < 30; ++f) { const float d = fr[3 * f] - mean; q = fmaf(d, d, q); }
    out[r * 6 + a] = mean;
    out[r * 6 + 3 + a] = q * (1.0f / 30.0f);
  }
}

__global__ void gauss_resample_kernel(long long rows, int mode, const float* __restrict__ muvar,
                                      const float* __restrict__ noise, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * 90;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 90;
    const int a = (int)(i % 3);
    const float mu = muvar[r * 6 + a];
    float var = muvar[r * 6 + 3 + a];
    float sd;
    if (mode == 0) { if (var < 0.0f) var = 1e-3f; sd = sqrtf(var); }
    else if (mode == 1) sd = sqrtf(var);
    else sd = var;
    out[i] = fmaf(sd, noise[i], mu);
  }
}

// d(mu,var) of out = mu + sd(var) * noise: one thread per (row, axis), 30 frames each
__global__ void gauss_resample_bwd_kernel(long long rows, int mode, const float* __restrict__ muvar,
                                          const float* __restrict__ noise, const float* __restrict__ dout,
                                          float* __restrict__ dmuvar) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * 3;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 3;
    const int a = (int)(i - r * 3);
    const float* d = dout + r * 90 + a;
    const float* z = noise + r * 90 + a;
    float dm = 0.0f, ds = 0.0f;
    for (int f = 0; f < 30; ++f) { dm += d[3 * f]; ds = fmaf(d[3 * f], z[3 * f], ds); }
    const float var = muvar[r * 6 + 3 + a];
    float dv;
    if (mode == 0) dv = var < 0.0f ? 0.0f : ds * 0.5f * rsqrtf(var);        // floored branch is constant
    else if (mode == 1) dv = ds * 0.5f * rsqrtf(var);
    else dv = ds;
    dmuvar[r * 6 + a] = dm;
    dmuvar[r * 6 + 3 + a] = dv;
  }
}

// Philox4x32-10 (Salmon et al. 2011; the generator TF's K.random_normal uses, mycode/convlstm_seq2seq.py:57):
// counter = (idx_lo, idx_hi, 0, 0), key = (seed_lo, seed_hi); element 4*idx + j of the stream is word j.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
// raw words (bit-exact against the oracle's integer restatement) and Box-Muller normals:
// u1 = (w0 + 1) * 2^-32 in (0,1], u2 = w1 * 2^-32; z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = ... sin(...)
__global__ void philox_kernel(long long n, unsigned long long seed, unsigned long long offset,
                              uint32_t* __restrict__ words, float* __restrict__ normal) {
  const long long nblk = (n + 3) / 4;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    const unsigned long long ctr = offset + (unsigned long long)b;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)ww[2 * h] + 1.0f) * 2.3283064365386963e-10f;
      const float u2 = (float)ww[2 * h + 1] * 2.3283064365386963e-10f;
      const float rad = sqrtf(-2.0f * logf(fminf(u1, 1.0f)));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      z[2 * h] = rad * cs; z[2 * h + 1] = rad * sn;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = 4 * b + j;
      if (i < n) {
        if (words) words[i] = ww[j];
        if (normal) normal[i] = z[j];
      }
    }
  }
}

}  // namespace

// internal launchers (used by convlstm.cu)
int fov_launch_gates_fwd(const GatesFwdArgs& a, cudaStream_t st) {
  convlstm_gates_fwd_kernel<<<grid_for(a.npix * a.F), 256, 0, st>>>(a);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
int fov_launch_gates_bwd(const GatesBwdArgs& a, cudaStream_t st) {
  convlstm_gates_bwd_kernel<<<grid_for(a.npix * a.F), 256, 0, st>>>(a);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
int fov_launch_mul_mask(long long n_img, long long img_elems, int C, const float* x, long long x_img, int x_pix,
                        const float* mask, float* out, cudaStream_t st) {
  mul_mask_kernel<<<grid_for(n_img * img_elems), 256, 0, st>>>(n_img, img_elems, C, x, x_img, x_pix, mask, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_softmax_fwd(long long rows, int C, const float* x, float* y, void* stream) {
  FOV_CHECK_ARG(rows >= 0 && C > 0 && x && y, "bad args");
  if (rows == 0) return FOV_OK;
  softmax_fwd_kernel<<<grid_for(rows, 128), 128, 0, (cudaStream_t)stream>>>(rows, C, x, y);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
extern "C" int fov_softmax_bwd(long long rows, int C, const float* y, const float* dy, float* dx, void* stream) {
  FOV_CHECK_ARG(rows >= 0 && C > 0 && y && dy && dx, "bad args");
  if (rows == 0) return FOV_OK;
  softmax_bwd_kernel<<<grid_for(rows, 128), 128, 0, (cudaStream_t)stream>>>(rows, C, y, dy, dx);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_mse_fwd_bwd(long long n, const float* y, const float* t, float weight, float* loss,
                               float* dy, void* stream) {
  FOV_CHECK_ARG(n > 0 && y && t && loss, "bad args");
  mse_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(n, y, t, weight / (float)n, loss, dy);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_gauss_nll_fwd_bwd(int B, int T, int running_length, const float* y, const float* frames,
                                     float weight, float* loss, float* dy, void* stream) {
  FOV_CHECK_ARG(B > 0 && T > 0 && running_length > 0 && y && frames && loss, "bad args");
  const long long rows = (long long)B * T;
  const float scale = weight / ((float)B * (float)running_length * 30.0f);
  gauss_nll_kernel<<<grid_for(rows * 3, 256, 4), 256, 0, (cudaStream_t)stream>>>(rows, y, frames, scale, loss, dy);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_cce_fwd_bwd(long long rows, int C, const float* p, const float* t, float weight, float* loss,
                               float* dp, void* stream) {
  FOV_CHECK_ARG(rows > 0 && C > 0 && p && t && loss, "bad args");
  cce_kernel<<<grid_for(rows, 256, 4), 256, 0, (cudaStream_t)stream>>>(rows, C, p, t, weight / (float)rows, loss, dp);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_adam_step(long long n, float* p, const float* g, float* m, float* v, int t, float lr,
                             float beta1, float beta2, float eps, float grad_scale, const float* grad_div,
                             void* stream) {
  FOV_CHECK_ARG(n > 0 && p && g && m && v && t >= 1, "bad args");
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, p, g, m, v, (float)lr_t, beta1, beta2, eps, grad_scale,
                                                          grad_div);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_rmsprop_step(long long n, float* p, const float* g, float* a, float lr, float rho, float eps,
                                float grad_scale, const float* grad_div, void* stream) {
  FOV_CHECK_ARG(n > 0 && p && g && a, "bad args");
  rmsprop_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, p, g, a, lr, rho, eps, grad_scale, grad_div);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_mean_var_xyz(long long rows, const float* frames, float* out, void* stream) {
  FOV_CHECK_ARG(rows > 0 && frames && out, "bad args");
  mean_var_kernel<<<grid_for(rows * 3), 256, 0, (cudaStream_t)stream>>>(rows, frames, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_gauss_resample(long long rows, int mode, const float* muvar, const float* noise, float* out,
                                  void* stream) {
  FOV_CHECK_ARG(rows > 0 && mode >= 0 && mode <= 2 && muvar && noise && out, "bad args");
  gauss_resample_kernel<<<grid_for(rows * 90), 256, 0, (cudaStream_t)stream>>>(rows, mode, muvar, noise, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
extern "C" int fov_gauss_resample_bwd(long long rows, int mode, const float* muvar, const float* noise,
                                      const float* dout, float* dmuvar, void* stream) {
  FOV_CHECK_ARG(rows > 0 && mode >= 0 && mode <= 2 && muvar && noise && dout && dmuvar, "bad args");
  gauss_resample_bwd_kernel<<<grid_for(rows * 3), 256, 0, (cudaStream_t)stream>>>(rows, mode, muvar, noise, dout, dmuvar);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_philox_normal(long long n, unsigned long long seed, unsigned long long offset, unsigned int* words,
                                 float* normal, void* stream) {
  FOV_CHECK_ARG(n > 0 && (words || normal), "bad args");
  philox_kernel<<<grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(n, seed, offset, words, normal);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

// ---------------------------------------------------------------------------------------------
// ConvLSTM2D input dropout as a widened input (include/fov360.h): expand / reduce helpers
// ---------------------------------------------------------------------------------------------
namespace {
// x4[((b*T + t)*HW + p), g*Cin + c] = x[b,t,p,c] * masks[g,b,p,c]
__global__ void dropout_expand_kernel(long long n, int B, int T, int HW, int Cin, const float* __restrict__ x,
                                      long long xb, long long xt, int xp, const float* __restrict__ masks,
                                      float* __restrict__ x4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    long long r = i / Cin;
    const int g = (int)(r & 3); r >>= 2;
    const int p = (int)(r % HW); r /= HW;
    const int t = (int)(r % T);
    const long long b = r / T;
    const float m = __ldg(&masks[(((long long)g * B + b) * HW + p) * Cin + c]);
    x4[i] = __ldg(&x[b * xb + (long long)t * xt + (long long)p * xp + c]) * m;
  }
}
// dx[b,t,p,c] (+)= sum_g dx4[(b,t,p), g*Cin + c] * masks[g,b,p,c]
__global__ void dropout_reduce_kernel(long long n, int B, int T, int HW, int Cin, const float* __restrict__ dx4,
                                      const float* __restrict__ masks, float* __restrict__ dx, long long xb, long long xt,
                                      int xp, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    long long r = i / Cin;
    const int p = (int)(r % HW); r /= HW;
    const int t = (int)(r % T);
    const long long b = r / T;
    const float* src = dx4 + (((b * T + t) * HW + p) * 4) * (long long)Cin + c;
    float acc = 0.0f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      acc = fmaf(__ldg(&src[(long long)g * Cin]), __ldg(&masks[(((long long)g * B + b) * HW + p) * Cin + c]), acc);
    float* dst = dx + b * xb + (long long)t * xt + (long long)p * xp + c;
    *dst = accumulate ? *dst + acc : acc;
  }
}
// K4[tap, g*Cin + c, n] = K[tap,c,n] when n / F == g, else 0
__global__ void gate_kernel_expand_kernel(long long n, int Cin, int F, const float* __restrict__ k, float* __restrict__ k4) {
  const int N4 = 4 * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % N4);
    long long r = i / N4;
    const int gc = (int)(r % (4 * Cin));
    const long long tap = r / (4 * Cin);
    const int g = gc / Cin, c = gc - g * Cin;
    k4[i] = (col / F == g) ? __ldg(&k[(tap * Cin + c) * N4 + col]) : 0.0f;
  }
}
// gK[tap,c,n] += gK4[tap, (n / F)*Cin + c, n]
__global__ void gate_kernel_reduce_kernel(long long n, int Cin, int F, const float* __restrict__ g4, float* __restrict__ gk) {
  const int N4 = 4 * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % N4);
    long long r = i / N4;
    const int c = (int)(r % Cin);
    const long long tap = r / Cin;
    gk[i] += __ldg(&g4[((tap * 4 + col / F) * Cin + c) * N4 + col]);
  }
}
}  // namespace

extern "C" int fov_dropout_expand(int B, int T, int HW, int Cin, const float* x, long long x_b_stride,
                                  long long x_t_stride, int x_pix_stride, const float* masks, float* x4, void* stream) {
  FOV_CHECK_ARG(B > 0 && T > 0 && HW > 0 && Cin > 0 && x && masks && x4, "bad args");
  const long long n = (long long)B * T * HW * 4 * Cin;
  dropout_expand_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, B, T, HW, Cin, x, x_b_stride, x_t_stride,
                                                                     x_pix_stride, masks, x4);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
extern "C" int fov_dropout_reduce(int B, int T, int HW, int Cin, const float* dx4, const float* masks, float* dx,
                                  long long x_b_stride, long long x_t_stride, int x_pix_stride, int accumulate,
                                  void* stream) {
  FOV_CHECK_ARG(B > 0 && T > 0 && HW > 0 && Cin > 0 && dx4 && masks && dx, "bad args");
  const long long n = (long long)B * T * HW * Cin;
  dropout_reduce_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, B, T, HW, Cin, dx4, masks, dx, x_b_stride,
                                                                     x_t_stride, x_pix_stride, accumulate);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
extern "C" int fov_gate_kernel_expand(int taps, int Cin, int F, const float* kernel, float* kernel4, void* stream) {
  FOV_CHECK_ARG(taps > 0 && Cin > 0 && F > 0 && kernel && kernel4, "bad args");
  const long long n = (long long)taps * 4 * Cin * 4 * F;
  gate_kernel_expand_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, Cin, F, kernel, kernel4);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
extern "C" int fov_gate_kernel_reduce(int taps, int Cin, int F, const float* g_kernel4, float* g_kernel, void* stream) {
  FOV_CHECK_ARG(taps > 0 && Cin > 0 && F > 0 && g_kernel4 && g_kernel, "bad args");
  const long long n = (long long)taps * Cin * 4 * F;
  gate_kernel_reduce_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, Cin, F, g_kernel4, g_kernel);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
