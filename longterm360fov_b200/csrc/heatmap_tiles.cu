// Gaussian-FoV and head-direction tiles: the (18, 36, fps) per-second heat maps that feed the
// heatmap ConvLSTM in the reference's Gaussian-FoV data generator, built on the device from frame centres.
//
// Reference: mycode/data_generator_gaussian_FoV.py
//   crop_FoV_from_equirect :57-118 / get_gaussian_FoV :121-127            (kind 0)
//   blur_head_direction_equirect :163-224 / get_head_direction :226-232   (kind 1)
//   get_theta_phi_array(_per_user) :21-55, *_giventhetaphi :130-138 / :235-243, heatmap_sum :246-261
// The reference paints a 180 x 360 float64 map per frame in Python loops, normalises ALL maps of the call by their
// common maximum, and keeps every 10th row / column.  Here:
//   1. frame_peak_kernel: the maximum of the full-resolution maps (one CTA per frame walks only the visited rows;
//      float32 of each pixel exactly as `astype(float32)` sees it; atomicMax on the bit pattern of a non-negative float);
//   2. tiles_kernel: one CTA per second writes its (18, 36, F) block with coalesced stores, evaluating only the 648
//      kept pixels of each frame, already divided by the peak.
// Everything that depends on the image row only (the FoV half width, the sigma of the last visited row) is computed
// on the HOST with the libm calls CPython makes (cos, pow), so the integer truncations match the reference bit for
// bit; the device evaluates exp() of exactly the same float64 argument and rounds once to float32.
#include <math.h>

#include "fov_common.cuh"

namespace {

constexpr int kH = 180, kW = 360, kTH = 18, kTW = 36;

struct TileTabs {
  short lon[kH];     // half width of the visited column span at image row r
  double s2[256];    // sigma ** 2 as a function of the LAST visited row
  int kind;          // 0 = FoV crop, 1 = head direction
  int nrows;         // rows visited per frame (90 / 50)
};

struct FrameP {
  int xi, zi, zoo, wrapped, r_first;
  double s2;
};

__device__ __forceinline__ FrameP frame_params(const TileTabs& tb, double x, double z) {
  FrameP p;
  p.xi = (int)(x * (double)kH);                       // int(): truncation
  p.zi = (int)(z * (double)kW);
  int last;
  if (tb.kind == 0) {
    p.r_first = (p.xi + 136) % kH;                    // row = int(xi - 45.0 + 180); row += 1; row %= 180
    last = (p.r_first + tb.nrows - 1) % kH;
  } else {
    int row = (int)((double)p.xi - 25.0);             // int(xi - blur_h / 2)
    if (row <= 0) row = 0;
    p.r_first = row + 1;
    last = p.r_first + tb.nrows - 1;
  }
  p.s2 = tb.s2[last < 0 ? 0 : (last > 255 ? 255 : last)];
  p.zoo = p.zi;
  if (p.zi + 108 > kW) p.zoo = p.zi - kW;             // margin = int(360 * 0.3)
  if (p.zi - 108 < 0) p.zoo = p.zi + kW;
  p.wrapped = p.zoo != p.zi;
  return p;
}

__device__ __forceinline__ bool row_visited(const TileTabs& tb, const FrameP& p, int r) {
  if (tb.kind == 0) return (r - p.r_first + kH) % kH < tb.nrows;
  return r >= p.r_first && r < p.r_first + tb.nrows;
}

// the three slice assignments of one row (:82-93 / :192-203), as a predicate on the column
__device__ __forceinline__ bool col_painted(int zi, int lon, int c) {
  int zlow = zi - lon, zhigh = zi + lon;
  bool in = false;
  if (zlow < 0) { in = c >= zlow + kW; zlow = 0; }
  if (zhigh > kW) { in = in || c < zhigh % kW; zhigh = kW - 1; }
  return in || (c >= zlow && c < zhigh);
}

// (G + G1) of a painted pixel, float64 like NumPy, then astype(float32)
__device__ __forceinline__ float pixel_value(const FrameP& p, int r, int c) {
  const int dy = r - p.xi, dx = c - p.zi, dx1 = c - p.zoo;
  const double g = exp(__dmul_rn(-(double)(dx * dx + dy * dy) / 2.0, p.s2));
  double g1;
  if (p.wrapped) g1 = exp(__dmul_rn(-(double)(dx1 * dx1 + dy * dy) / 2.0, p.s2));
  else g1 = __dmul_rn(0.00001, g);
  return (float)__dadd_rn(g, g1);
}

__global__ void __launch_bounds__(128) frame_peak_kernel(TileTabs tb, long long n, const double* __restrict__ pt,
                                                         int* __restrict__ peak_bits) {
  const long long k = blockIdx.x;
  if (k >= n) return;
  const FrameP p = frame_params(tb, pt[2 * k], pt[2 * k + 1]);
  float m = 0.0f;
  for (int i = 0; i < tb.nrows; ++i) {
    const int r = tb.kind == 0 ? (p.r_first + i) % kH : p.r_first + i;
    if (r >= kH) break;
    const int lon = tb.lon[r];
    for (int c = threadIdx.x; c < kW; c += blockDim.x)
      if (col_painted(p.zi, lon, c)) m = fmaxf(m, pixel_value(p, r, c));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(peak_bits, __float_as_int(m));
}

__global__ void __launch_bounds__(256) tiles_kernel(TileTabs tb, long long maps, int F, const double* __restrict__ pt,
                                                    const float* __restrict__ peak, float* __restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  FrameP* fp = reinterpret_cast<FrameP*>(smem_raw);
  const long long m = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x) fp[f] = frame_params(tb, pt[2 * (m * F + f)], pt[2 * (m * F + f) + 1]);
  __syncthreads();
  const float pk = *peak;
  float* o = out + m * (long long)(kTH * kTW) * F;
  for (int e = threadIdx.x; e < kTH * kTW * F; e += blockDim.x) {
    const int cell = e / F, f = e - cell * F;
    const int r = (cell / kTW) * 10, c = (cell % kTW) * 10;
    const FrameP& p = fp[f];
    float v = 0.0f;
    if (row_visited(tb, p, r) && col_painted(p.zi, tb.lon[r], c)) v = pixel_value(p, r, c);
    o[e] = __fdiv_rn(v, pk);
  }
}

// [phi / pi, (theta + pi) / 2 / pi] of every frame (get_theta_phi_array_per_user :48-53 after xyz2thetaphi)
__global__ void __launch_bounds__(256) theta_phi_frames_kernel(long long n, const float* __restrict__ xyz,
                                                               double* __restrict__ pt) {
  const double kPi = 3.141592653589793;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = (double)xyz[3 * i], y = (double)xyz[3 * i + 1], z = (double)xyz[3 * i + 2];
    double mth = fmod(atan2(y, x), 2.0 * kPi);
    if (mth < 0.0) mth += 2.0 * kPi;
    const double theta = mth - kPi;
    double phi = fmod(atan2(z, sqrt(x * x + y * y)) + kPi / 2.0, kPi);
    if (phi < 0.0) phi += kPi;
    pt[2 * i] = phi / kPi;
    pt[2 * i + 1] = (theta + kPi) / 2.0 / kPi;
  }
}

// NumPy's float32 pairwise summation of a contiguous run (the order np.sum uses), n <= 128 leaf
__device__ float np_sum_leaf(const float* a, int n) {
  if (n < 8) {
    float r = 0.0f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}
__device__ float np_sum(const float* a, int n) {
  if (n <= 128) return np_sum_leaf(a, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(np_sum(a, n2), np_sum(a + n2, n - n2));
}

// heatmap_sum + normalize_to_distribution (:246-261): per map, sum the F frame channels of every cell, then divide
// by the sum of the map; both sums in NumPy's order so the float32 results are the reference's
__global__ void __launch_bounds__(256) heatmap_sum_kernel(int cells, int F, const float* __restrict__ tiles,
                                                          float* __restrict__ out) {
  extern __shared__ float cell_s[];      // cells + 1
  const long long m = blockIdx.x;
  const float* t = tiles + m * (long long)cells * F;
  for (int c = threadIdx.x; c < cells; c += blockDim.x) cell_s[c] = np_sum_leaf(t + (long long)c * F, F);
  __syncthreads();
  if (threadIdx.x == 0) cell_s[cells] = np_sum(cell_s, cells);
  __syncthreads();
  const float tot = cell_s[cells];
  for (int c = threadIdx.x; c < cells; c += blockDim.x) out[m * cells + c] = __fdiv_rn(cell_s[c], tot);
}

int make_tabs(int kind, TileTabs* tb) {
  tb->kind = kind;
  const double img_h = kH, img_w = kW;
  const double shrink = img_h / 256.0;
  // loop counts as the while loops count them (:71 `iterative<img_h/2`, :177 `iterative<blur_h`)
  const double limit = kind == 0 ? img_h / 2.0 : 5.0 / 18.0 * img_h;
  int n_it = 0;
  while ((double)n_it < limit) ++n_it;
  tb->nrows = n_it;
  const double half_span = kind == 0 ? (double)(int)(img_w / 6.0) : 0.5 * (5.0 / 36.0) * img_w;
  for (int r = 0; r < kH; ++r) {
    const double rowx = (double)r / img_h;
    int lon = (int)(half_span / (cos(M_PI * fabs(rowx - 0.5)) + 0.00001));
    if (kind == 0) {
      if ((double)lon > img_w / 2.0 - 1.0) lon = (int)(img_w / 2.0 - 1.0);
    } else if (lon >= kH - 1) {
      lon = kH - 1;
    }
    tb->lon[r] = (short)lon;
  }
  for (int r = 0; r < 256; ++r) {
    const double rowx = (double)r / img_h;
    double sigma = (kind == 0 ? 0.01 : 0.05) + 0.008 * (M_PI * fabs(rowx - 0.5));
    sigma = sigma / shrink;
    tb->s2[r] = pow(sigma, 2.0);          // CPython's float ** 2
  }
  return 0;
}

}  // namespace

extern "C" int fov_theta_phi_frames(long long frames, const float* xyz, double* phi_theta, void* stream) {
  FOV_CHECK_ARG(frames > 0 && xyz && phi_theta, "bad arguments");
  long long blocks = (frames + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  theta_phi_frames_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(frames, xyz, phi_theta);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_gaussian_fov_tiles(long long maps, int frames, int kind, const double* phi_theta, float* out,
                                      float* peak, void* stream) {
  FOV_CHECK_ARG(maps > 0 && maps * frames < (1LL << 31) && frames > 0 && frames <= 1024, "bad shape");
  FOV_CHECK_ARG(kind == 0 || kind == 1, "kind must be 0 (FoV crop) or 1 (head direction)");
  FOV_CHECK_ARG(phi_theta && out && peak, "NULL pointer");
  TileTabs tb;
  make_tabs(kind, &tb);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(peak, 0, sizeof(float), st) != cudaSuccess) {
    fov_set_error("%s: cudaMemsetAsync failed", __func__);
    return FOV_ERR_CUDA;
  }
  frame_peak_kernel<<<(unsigned)(maps * frames), 128, 0, st>>>(tb, maps * frames, phi_theta, reinterpret_cast<int*>(peak));
  FOV_CUDA_LAUNCH_CHECK();
  tiles_kernel<<<(unsigned)maps, 256, frames * sizeof(FrameP), st>>>(tb, maps, frames, phi_theta, peak, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

extern "C" int fov_heatmap_sum(long long maps, int cells, int frames, const float* tiles, float* out, void* stream) {
  FOV_CHECK_ARG(maps > 0 && maps < (1LL << 31) && cells > 0 && cells <= 8192 && frames > 0 && frames <= 128, "bad shape");
  FOV_CHECK_ARG(tiles && out, "NULL pointer");
  heatmap_sum_kernel<<<(unsigned)maps, 256, (cells + 1) * sizeof(float), (cudaStream_t)stream>>>(cells, frames, tiles, out);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}
