// Implicit-GEMM convolution family, fp32 (parity path): forward (+bias, activation,
// accumulate), backward-data (forward with flipped/transposed weights) and
// backward-weight (split over pixels, atomic accumulate).  Channels-last, stride 1.
// A Dense layer is the 1x1 / H=W=1 case.  The flattened K index is
// k = tap*Cin + ci, which is exactly the row index of the Keras kernel
// (kh,kw,Cin,Cout) viewed as a [K][Cout] matrix.
//
// Replaces the Keras Dense / Conv1D / Conv2D / ConvLSTM2D gate convolutions cited in
// include/fov360.h.
#include "fov_common.cuh"
#include "fov_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int BK = 8;

struct ConvK {
  int N, H, W, Cin, Cout, kh, kw, dil_h, dil_w, pad_h, pad_w;
  long long x_img_stride, y_img_stride;
  int x_pix_stride, y_pix_stride;
  int act;
  float beta;
  long long M;   // N*H*W
  int K;         // kh*kw*Cin
};

// Forward: C[m][n] = sum_k A(m,k) * Wt[k][n]
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(kThreads) conv_fwd_kernel(const ConvK p, const float* __restrict__ x,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            float* __restrict__ y) {
  static_assert((BM / TM) * (BN / TN) == kThreads, "tile/thread mismatch");
  constexpr int A_PER = BM * BK / kThreads;   // elements of A per thread per k-tile
  constexpr int B_PER = BN * BK / kThreads;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HW = p.H * p.W;

  // A loader: thread handles rows ar[i] with k offset ak[i]; all elements of one thread
  // share the row when A_PER <= BK (BM>=... ) -- we map idx = tid + i*kThreads, kk = idx % BK.
  int a_row[A_PER], a_kk[A_PER];
  long long a_base[A_PER];
  int a_y[A_PER], a_x[A_PER];
  bool a_ok[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    const int idx = tid + i * kThreads;
    a_kk[i] = idx % BK;
    a_row[i] = idx / BK;
    const long long m = m0 + a_row[i];
    a_ok[i] = m < p.M;
    const long long mm = a_ok[i] ? m : 0;
    const int n = (int)(mm / HW);
    const int rem = (int)(mm - (long long)n * HW);
    a_y[i] = rem / p.W;
    a_x[i] = rem - a_y[i] * p.W;
    a_base[i] = (long long)n * p.x_img_stride;
  }
  int b_kk[B_PER], b_n[B_PER];
#pragma unroll
  for (int i = 0; i < B_PER; ++i) {
    const int idx = tid + i * kThreads;
    b_n[i] = idx % BN;
    b_kk[i] = idx / BN;
  }

  float ra[A_PER], rb[B_PER];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int k = k0 + a_kk[i];
      float v = 0.0f;
      if (a_ok[i] && k < p.K) {
        const int tap = k / p.Cin, ci = k - tap * p.Cin;
        const int ty = tap / p.kw, tx = tap - ty * p.kw;
        const int yy = a_y[i] + ty * p.dil_h - p.pad_h;
        const int xx = a_x[i] + tx * p.dil_w - p.pad_w;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
          v = __ldg(&x[a_base[i] + (long long)(yy * p.W + xx) * p.x_pix_stride + ci]);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int k = k0 + b_kk[i], n = n0 + b_n[i];
      rb[i] = (k < p.K && n < p.Cout) ? __ldg(&w[(long long)k * p.Cout + n]) : 0.0f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) As[buf][a_kk[i]][a_row[i]] = ra[i];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) Bs[buf][b_kk[i]][b_n[i]] = rb[i];
  };

  constexpr int TX = BN / TN;          // threads along n
  const int tx = tid % TX, ty = tid / TX;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const int ntiles = (p.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][(i / 4) * (BM / (TM / 4)) + ty * 4]);
        av[i] = v.x; av[i + 1] = v.y; av[i + 2] = v.z; av[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][(j / 4) * (BN / (TN / 4)) + tx * 4]);
        bv[j] = v.x; bv[j + 1] = v.y; bv[j + 2] = v.z; bv[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < ntiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4);
    const long long m = m0 + row;
    if (m >= p.M) continue;
    const int n = (int)(m / HW);
    const int pix = (int)(m - (long long)n * HW);
    float* yrow = y + (long long)n * p.y_img_stride + (long long)pix * p.y_pix_stride;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + (j / 4) * (BN / (TN / 4)) + tx * 4 + (j % 4);
      if (col >= p.Cout) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(&bias[col]);
      if (p.beta != 0.0f) v += p.beta * yrow[col];
      yrow[col] = fov_act(p.act, v);
    }
  }
}

// Backward-weight: gw[k][n] += sum_m A(m,k) * dY[m][n], reduction split over blockIdx.z.
constexpr int WK = 64, WN = 64, WM = 16;
__global__ void __launch_bounds__(kThreads) conv_wgrad_kernel(const ConvK p, const float* __restrict__ x,
                                                              const float* __restrict__ dy,
                                                              float* __restrict__ gw, long long m_per_split) {
  __shared__ __align__(16) float As[WM][WK + 4];
  __shared__ __align__(16) float Ds[WM][WN + 4];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * WK, n0 = blockIdx.y * WN;
  const long long mbeg = (long long)blockIdx.z * m_per_split;
  long long mend = mbeg + m_per_split;
  if (mend > p.M) mend = p.M;
  const int HW = p.H * p.W;

  // loader mapping: 16 rows x 64 cols = 1024 elements, 4 per thread
  const int l_mm = tid / 16;            // 0..15
  const int l_c4 = (tid % 16) * 4;      // 0..60
  int a_ci[4], a_dy[4], a_dx[4];
  bool a_kok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + l_c4 + i;
    a_kok[i] = k < p.K;
    const int kk = a_kok[i] ? k : 0;
    const int tap = kk / p.Cin;
    a_ci[i] = kk - tap * p.Cin;
    const int ty = tap / p.kw, tx = tap - ty * p.kw;
    a_dy[i] = ty * p.dil_h - p.pad_h;
    a_dx[i] = tx * p.dil_w - p.pad_w;
  }
  const int tx = tid % 16, ty = tid / 16;   // 16x16 threads, 4x4 micro tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (long long mt = mbeg; mt < mend; mt += WM) {
    const long long m = mt + l_mm;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      const int n = (int)(m / HW);
      const int rem = (int)(m - (long long)n * HW);
      const int yy0 = rem / p.W, xx0 = rem - yy0 * p.W;
      const float* xb = x + (long long)n * p.x_img_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yy = yy0 + a_dy[i], xx = xx0 + a_dx[i];
        if (a_kok[i] && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
          av[i] = __ldg(&xb[(long long)(yy * p.W + xx) * p.x_pix_stride + a_ci[i]]);
      }
      const float* db = dy + (long long)n * p.y_img_stride + (long long)rem * p.y_pix_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = n0 + l_c4 + i;
        if (col < p.Cout) dv[i] = __ldg(&db[col]);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[l_mm][l_c4]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Ds[l_mm][l_c4]) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      const float4 a = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
      const float4 d = *reinterpret_cast<const float4*>(&Ds[mm][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], dd[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= p.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < p.Cout) atomicAdd(&gw[(long long)k * p.Cout + col], acc[i][j]);
    }
  }
}

// gbias[c] += sum over rows of dy[row][c]
__global__ void __launch_bounds__(256) colsum_kernel(long long M, int C, int HW, long long img_stride,
                                                     int pix_stride, const float* __restrict__ dy,
                                                     float* __restrict__ gb, long long rows_per_block) {
  // blockDim = (32, 8): x over columns, y over rows
  __shared__ float red[8][33];
  const int c = blockIdx.y * 32 + threadIdx.x;
  const long long rbeg = (long long)blockIdx.x * rows_per_block;
  long long rend = rbeg + rows_per_block;
  if (rend > M) rend = M;
  float s = 0.0f;
  if (c < C) {
    for (long long r = rbeg + threadIdx.y; r < rend; r += 8) {
      const long long n = r / HW;
      const long long pix = r - n * HW;
      s += dy[n * img_stride + pix * pix_stride + c];
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(&gb[c], t);
  }
}

// Narrow dense matrices (C <= 32 columns, rows contiguous): the flat array is walked with a stride that is a multiple of
// C, so a thread always sees the same column; all 256 threads of a CTA carry data whatever C is.
__global__ void __launch_bounds__(256) colsum_narrow_kernel(long long total, int C, const float* __restrict__ dy,
                                                            float* __restrict__ gb) {
  __shared__ float part[32];
  if (threadIdx.x < 32) part[threadIdx.x] = 0.0f;
  __syncthreads();
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const long long step = nthr / C * C;                 // <= nthr: the threads e0 < step cover every element once
  const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.0f;
  if (e0 < step) {
    for (long long e = e0; e < total; e += step) s += __ldg(dy + e);
    atomicAdd(&part[(int)(e0 % C)], s);
  }
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(&gb[threadIdx.x], part[threadIdx.x]);
}

// wt[(kh-1-i, kw-1-j)][co][ci] = w[(i,j)][ci][co]
__global__ void flip_transpose_kernel(int taps_h, int taps_w, int Cin, int Cout, const float* __restrict__ w,
                                      float* __restrict__ wt) {
  const long long total = (long long)taps_h * taps_w * Cin * Cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin);
    long long r = idx / Cin;
    const int co = (int)(r % Cout);
    r /= Cout;
    const int tj = (int)(r % taps_w), ti = (int)(r / taps_w);
    const int si = taps_h - 1 - ti, sj = taps_w - 1 - tj;
    wt[idx] = w[(((long long)si * taps_w + sj) * Cin + ci) * Cout + co];
  }
}

__global__ void act_bwd_kernel(int act, long long rows, int cols, const float* __restrict__ y, long long ys,
                               const float* __restrict__ dy, long long dys, float* __restrict__ dpre,
                               long long ds) {
  const long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = (int)(idx - r * cols);
    dpre[r * ds + c] = dy[r * dys + c] * fov_act_grad(act, y[r * ys + c]);
  }
}

int make_k(const fov_conv_cfg* c, ConvK* k) {
  FOV_CHECK_ARG(c != nullptr, "cfg is NULL");
  FOV_CHECK_ARG(c->N > 0 && c->H > 0 && c->W > 0 && c->Cin > 0 && c->Cout > 0, "bad shape");
  FOV_CHECK_ARG(c->kh > 0 && c->kw > 0 && c->dil_h > 0 && c->dil_w > 0, "bad kernel/dilation");
  k->N = c->N; k->H = c->H; k->W = c->W; k->Cin = c->Cin; k->Cout = c->Cout;
  k->kh = c->kh; k->kw = c->kw; k->dil_h = c->dil_h; k->dil_w = c->dil_w;
  k->pad_h = c->pad_h; k->pad_w = c->pad_w;
  k->x_img_stride = c->x_img_stride; k->y_img_stride = c->y_img_stride;
  k->x_pix_stride = c->x_pix_stride; k->y_pix_stride = c->y_pix_stride;
  k->act = c->act; k->beta = c->beta;
  k->M = (long long)c->N * c->H * c->W;
  k->K = c->kh * c->kw * c->Cin;
  return FOV_OK;
}

int launch_fwd(const ConvK& k, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  if (k.Cout <= 32) {
    dim3 grid((unsigned)((k.M + 255) / 256), (k.Cout + 31) / 32);
    conv_fwd_kernel<256, 32, 8, 4><<<grid, kThreads, 0, st>>>(k, x, w, bias, y);
  } else if (k.Cout <= 64) {
    dim3 grid((unsigned)((k.M + 127) / 128), (k.Cout + 63) / 64);
    conv_fwd_kernel<128, 64, 8, 4><<<grid, kThreads, 0, st>>>(k, x, w, bias, y);
  } else {
    dim3 grid((unsigned)((k.M + 127) / 128), (k.Cout + 127) / 128);
    conv_fwd_kernel<128, 128, 8, 8><<<grid, kThreads, 0, st>>>(k, x, w, bias, y);
  }
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

extern "C" int fov_conv2d_fwd(const fov_conv_cfg* cfg, const float* x, const float* w, const float* bias,
                              float* y, void* stream) {
  ConvK k;
  int rc = make_k(cfg, &k);
  if (rc) return rc;
  FOV_CHECK_ARG(x && w && y, "NULL pointer");
  return launch_fwd(k, x, w, bias, y, (cudaStream_t)stream);
}

int fov_conv_flip_weights(const fov_conv_cfg* cfg, const float* w, float* wt, cudaStream_t st) {
  ConvK f;
  int rc = make_k(cfg, &f);
  if (rc) return rc;
  FOV_CHECK_ARG(w && wt, "NULL pointer");
  const long long total = (long long)f.kh * f.kw * f.Cin * f.Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * fov_num_sms()) blocks = 4 * fov_num_sms();
  flip_transpose_kernel<<<blocks, 256, 0, st>>>(f.kh, f.kw, f.Cin, f.Cout, w, wt);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

int fov_conv_bwd_data_preflipped(const fov_conv_cfg* cfg, const float* dy, const float* wt, float* dx,
                                 cudaStream_t st) {
  ConvK f;
  int rc = make_k(cfg, &f);
  if (rc) return rc;
  FOV_CHECK_ARG(dy && wt && dx, "NULL pointer");
  ConvK b = f;
  b.Cin = f.Cout; b.Cout = f.Cin;
  b.pad_h = (f.kh - 1) * f.dil_h - f.pad_h;
  b.pad_w = (f.kw - 1) * f.dil_w - f.pad_w;
  b.x_img_stride = f.y_img_stride; b.x_pix_stride = f.y_pix_stride;
  b.y_img_stride = f.x_img_stride; b.y_pix_stride = f.x_pix_stride;
  b.act = FOV_ACT_LINEAR;
  b.K = f.kh * f.kw * f.Cout;
  return launch_fwd(b, dy, wt, nullptr, dx, st);
}

extern "C" int fov_conv2d_bwd_data(const fov_conv_cfg* cfg, const float* dy, const float* w, float* dx,
                                   float* ws, void* stream) {
  FOV_CHECK_ARG(ws != nullptr, "NULL workspace");
  int rc = fov_conv_flip_weights(cfg, w, ws, (cudaStream_t)stream);
  if (rc) return rc;
  return fov_conv_bwd_data_preflipped(cfg, dy, ws, dx, (cudaStream_t)stream);
}

extern "C" int fov_conv2d_bwd_weight(const fov_conv_cfg* cfg, const float* x, const float* dy, float* gw,
                                     float* gbias, void* stream) {
  ConvK k;
  int rc = make_k(cfg, &k);
  if (rc) return rc;
  FOV_CHECK_ARG(dy != nullptr, "NULL dy");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = k.H * k.W;
  if (gw) {
    FOV_CHECK_ARG(x != nullptr, "NULL x");
    const int gx = (k.K + WK - 1) / WK, gy = (k.Cout + WN - 1) / WN;
    // split the pixel reduction so the grid covers ~4 waves of the SMs
    long long want = (4LL * fov_num_sms() + gx * gy - 1) / (gx * gy);
    long long max_splits = (k.M + 255) / 256;
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    long long per = (k.M + want - 1) / want;
    per = (per + WM - 1) / WM * WM;
    const int gz = (int)((k.M + per - 1) / per);
    dim3 grid(gx, gy, gz);
    conv_wgrad_kernel<<<grid, kThreads, 0, st>>>(k, x, dy, gw, per);
    FOV_CUDA_LAUNCH_CHECK();
  }
  if (gbias && k.Cout <= 32 && k.y_pix_stride == k.Cout && k.y_img_stride == (long long)HW * k.Cout) {
    const long long total = k.M * k.Cout;
    long long blocks = (total + 256 * 16 - 1) / (256 * 16);
    if (blocks > 2 * fov_num_sms()) blocks = 2 * fov_num_sms();
    if (blocks < 1) blocks = 1;
    colsum_narrow_kernel<<<(unsigned)blocks, 256, 0, st>>>(total, k.Cout, dy, gbias);
    FOV_CUDA_LAUNCH_CHECK();
  } else if (gbias) {
    long long blocks = (k.M + 1023) / 1024;
    if (blocks > 2 * fov_num_sms()) blocks = 2 * fov_num_sms();
    if (blocks < 1) blocks = 1;
    const long long per = (k.M + blocks - 1) / blocks;
    dim3 grid((unsigned)((k.M + per - 1) / per), (k.Cout + 31) / 32);
    colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(k.M, k.Cout, HW, k.y_img_stride, k.y_pix_stride, dy, gbias, per);
    FOV_CUDA_LAUNCH_CHECK();
  }
  return FOV_OK;
}

extern "C" int fov_act_bwd(int act, long long rows, int cols, const float* y, long long y_stride,
                           const float* dy, long long dy_stride, float* dpre, long long dpre_stride,
                           void* stream) {
  FOV_CHECK_ARG(rows >= 0 && cols > 0 && y && dy && dpre, "bad args");
  if (rows == 0) return FOV_OK;
  long long blocks = (rows * cols + 255) / 256;
  if (blocks > 8LL * fov_num_sms()) blocks = 8LL * fov_num_sms();
  act_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(act, rows, cols, y, y_stride, dy, dy_stride, dpre,
                                                                dpre_stride);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

// ---------------------------------------------------------------------------------------------
// Tap-stacked narrow convolutions (the 1024 -> 30 head of convlstm_seq2seq, mycode/convlstm_seq2seq.py:179-181, and
// the 1024 -> 3 Conv1D head, :187-189).  With Cout <= 32 a kh x kw convolution issues kh*kw tcgen05.mma of N = 32 per
// k-step, each at the ~45-cycle floor of a small-N MMA.  The host instead runs the kh x 1 convolution
//   Y'[n,h,w,(tx,c)] = sum_{ty,ci} x[n,h+ty-ph,w,ci] * K[ty,tx,ci,c]          (Cout' = kw*Cp columns: kw x fewer MMAs)
// on the same shifted-tap kernel and these two kernels fold / unfold the kw column taps:
//   y[n,h,w,c] = act(b[c] + sum_tx Y'[n,h,w+tx-pw,(tx,c)])            dY'[n,h,w',(tx,c)] = dpre[n,h,w'-tx+pw,c]
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) tapstack_reduce_kernel(long long rows, int W, int kw, int pad_w, int Cp, int Cout,
                                                              const float* __restrict__ yp, const float* __restrict__ bias,
                                                              int act, float* __restrict__ y) {
  const long long total = rows * W * Cout;
  const int CW = kw * Cp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cout);
    const long long pw = e / Cout;
    const int w = (int)(pw % W);
    const long long r = pw / W;
    float acc = bias ? __ldg(&bias[c]) : 0.0f;
    for (int tx = 0; tx < kw; ++tx) {
      const int ws = w + tx - pad_w;
      if (ws >= 0 && ws < W) acc += __ldg(&yp[(r * W + ws) * CW + tx * Cp + c]);
    }
    y[e] = fov_act(act, acc);
  }
}
__global__ void __launch_bounds__(256) tapstack_expand_kernel(long long rows, int W, int kw, int pad_w, int Cp, int Cout,
                                                              const float* __restrict__ dpre, float* __restrict__ dyp) {
  const int CW = kw * Cp;
  const long long total = rows * W * CW;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(e % CW);
    const long long pw = e / CW;
    const int w = (int)(pw % W);
    const long long r = pw / W;
    const int tx = col / Cp, c = col - tx * Cp;
    const int wd = w - tx + pad_w;
    float v = 0.0f;
    if (c < Cout && wd >= 0 && wd < W) v = __ldg(&dpre[(r * W + wd) * Cout + c]);
    dyp[e] = v;
  }
}
}  // namespace

extern "C" int fov_tapstack_reduce(long long rows, int W, int kw, int pad_w, int Cp, int Cout, const float* yp,
                                   const float* bias, int act, float* y, void* stream) {
  FOV_CHECK_ARG(rows > 0 && W > 0 && kw > 0 && pad_w >= 0 && pad_w < kw && Cp >= Cout && Cout > 0 && yp && y, "bad args");
  long long blocks = (rows * W * Cout + 255) / 256;
  if (blocks > 8LL * fov_num_sms()) blocks = 8LL * fov_num_sms();
  tapstack_reduce_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rows, W, kw, pad_w, Cp, Cout, yp, bias, act, y);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_tapstack_expand(long long rows, int W, int kw, int pad_w, int Cp, int Cout, const float* dpre,
                                   float* dyp, void* stream) {
  FOV_CHECK_ARG(rows > 0 && W > 0 && kw > 0 && pad_w >= 0 && pad_w < kw && Cp >= Cout && Cout > 0 && dpre && dyp, "bad args");
  long long blocks = (rows * W * kw * Cp + 255) / 256;
  if (blocks > 8LL * fov_num_sms()) blocks = 8LL * fov_num_sms();
  tapstack_expand_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rows, W, kw, pad_w, Cp, Cout, dpre, dyp);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
