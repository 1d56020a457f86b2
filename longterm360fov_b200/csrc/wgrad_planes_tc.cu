// Weight gradient of the wide convolution heads on tensor cores, fed by TMA from bf16 planes.
//
//   gw[ty][tx][ci][n] += sum_{img,y,x} x[img, y+ty-pad_h, x+tx-pad_w, ci] * dy[img, y, x, n]
//
// for the 5x5 Conv2D heads of convlstm_seq2seq (56 -> 512 -> 1024 -> 30 on 36x18 maps,
// mycode/convlstm_seq2seq.py:175-181): the 512 -> 1024 layer alone is 31 % of a training step.  The general
// weight-gradient kernel (wgrad_tc.cu) gathers fp32 operands per (tap, channel) k tile, i.e. re-reads every x element
// 100 times and every dY element 50 times from L2 and converts it each time: it sits at the L2 bandwidth cap
// (17 GB of L2 reads per call, ~19 % of the tensor peak).  Here
//   * both operands are converted ONCE per call into bf16 planes [term][frame position][channel] (hi / lo terms of
//     the fp32 split) laid out in a zero-padded linear frame: pixel (img,y,x) sits at row img*Hp*Wp + y*Wp + x with
//     Hp = H + pad, Wp = W + pad rows / columns of zeros SHARED between neighbouring rows / images, so a kernel tap
//     (ty,tx) is the constant row offset (ty-pad_h)*Wp + (tx-pad_w);
//   * the pixel reduction is the MMA K index, both operands MN-major exactly as the planes lie in HBM, and whole
//     tiles arrive by cp.async.bulk.tensor (SASS UTMALDG) with the 128-byte swizzle: no thread touches operand data;
//   * the kw taps of a kernel row are overlapping N groups of ONE MMA (descriptor LBO = one row, verified by
//     tests/cuda/tc_mnshift_probe.cu): a CTA owns (kernel row ty, 64 input channels, 128 output channels) and
//     accumulates D[128 n][kw x 64 ci] in TMEM over every frame position; per 64-position stage it needs one
//     64 x 128 dY tile and one (64 + kw - 1) x 64 x tile: 25 KB per 700 tensor cycles instead of 64 KB per 512;
//   * partial results of the pixel splits are added with coalesced red.global.add.
// The bias gradient rides on the dY plane conversion (per-thread column sums, one red.global.add per channel).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <type_traits>

#include "fov_common.cuh"
#include "fov_internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kStagePos = 64;                 // frame positions per pipeline stage (4 MMA K steps)
constexpr int kXRows = 72;                    // x tile rows: 64 + kw - 1 <= 72
constexpr int kThr = 128;
constexpr uint32_t kAGroupBytes = kStagePos * 128;   // one 64-channel group of a dY tile
constexpr uint32_t kXBytes = kXRows * 128;
constexpr int kMaxStages = 8;
constexpr int kMaxColsumC = 2048;            // widest dY the fused bias gradient handles (channels, padded)

struct WpParams {
  int Cin, Cout, kh, kw, pad_h, pad_w, Wp;
  int n_cg, n_nb, nbw, a_groups;
  int stages_total, stages_per_split, splits, depth;
  uint32_t a_bytes, stage_bytes;
  float* gw;
};

struct WpBook {
  uint4 mtab[4 * 6 * 2];                      // MMAs of one stage: 4 K steps x term pairs (<= 6) x tap chunks (<= 2)
  uint64_t full[kMaxStages], empty[kMaxStages], done;
  uint32_t tmem_ptr;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

template <int NS>
__global__ void __launch_bounds__(kThr, 1) wgrad_planes_kernel(const __grid_constant__ CUtensorMap xmap,
                                                               const __grid_constant__ CUtensorMap ymap, const WpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  WpBook* bk = reinterpret_cast<WpBook*>(smem + (size_t)p.depth * p.stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // work item: (kernel row, 64-channel group of x, output-channel block), pixel split fastest.  (Measured: the other
  // order - CTAs of a wave sharing one position range - is 12-200 % SLOWER: a hundred SMs asking the same L2 lines at the
  // same time serialise on a few L2 slices; spreading the wave over the planes keeps every slice busy.)
  const int item = (int)blockIdx.x / p.splits, split = (int)blockIdx.x - item * p.splits;
  const int nb = item % p.n_nb, cg = (item / p.n_nb) % p.n_cg, ty = item / (p.n_nb * p.n_cg);
  const int n0 = nb * p.nbw, c0 = cg * 64;
  const int st_beg = split * p.stages_per_split;
  int st_end = st_beg + p.stages_per_split;
  if (st_end > p.stages_total) st_end = p.stages_total;
  const int nst = st_end - st_beg;
  const uint32_t tmem_cols = tmem_cols_for(p.kw * 64);

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&ymap) : "memory");
    for (int s = 0; s < p.depth; ++s) {
      mbar_init(smem_u32(&bk->full[s]), 1);
      mbar_init(smem_u32(&bk->empty[s]), 1);
    }
    mbar_init(smem_u32(&bk->done), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bk->tmem_ptr), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bk->tmem_ptr;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    const int xrow_off = (ty - p.pad_h) * p.Wp - p.pad_w;        // frame row of tap (ty, 0) relative to the dY row
    for (int it = 0; it < nst; ++it) {
      const int slot = it % p.depth;
      mbar_wait(smem_u32(&bk->empty[slot]), ((uint32_t)(it / p.depth) & 1u) ^ 1u);
      const uint32_t bar = smem_u32(&bk->full[slot]);
      mbar_arrive_expect_tx(bar, p.stage_bytes);
      const uint32_t s0 = base + (uint32_t)slot * p.stage_bytes;
      const int pos = (st_beg + it) * kStagePos;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        for (int g = 0; g < p.a_groups; ++g)
          tma_load_3d(s0 + (uint32_t)s * p.a_bytes + (uint32_t)g * kAGroupBytes, &ymap, n0 + g * 64, pos, s, bar);
        tma_load_3d(s0 + NS * p.a_bytes + (uint32_t)s * kXBytes, &xmap, c0, pos + xrow_off, s, bar);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    // Every MMA of a stage fully resolved once for ring slot 0 (descriptor low words, instruction descriptor,
    // accumulator column): the issue loop is one 16-byte table read, two adds and the tcgen05.mma.  A lone thread
    // that computes its descriptors issues an MMA every 100-180 cycles; the tensor pipe needs 64-128 here.
    const uint32_t a_lbo = p.a_groups > 1 ? kAGroupBytes : 0u;    // 0: one 64-channel group, lanes 64..127 duplicate it
    const int n_chunk = (p.kw + 3) / 4;                           // taps of a kernel row in <= 4-tap MMAs (N <= 256),
    const int per_chunk = (p.kw + n_chunk - 1) / n_chunk;         // split evenly: 5 taps = 3 + 2
    uint4* mtab = bk->mtab;
    int nm = 0;
    for (int k16 = 0; k16 < kStagePos / 16; ++k16)
      for (int sum = NS - 1; sum >= 0; --sum)
        for (int sa = 0; sa <= sum; ++sa) {
          const int sb = sum - sa;
          const uint64_t ad = smem_desc_sw128(base + (uint32_t)sa * p.a_bytes + (uint32_t)k16 * 2048u, a_lbo, 1024);
          for (int ch = 0; ch < n_chunk; ++ch) {
            const int tap0 = ch * per_chunk, ntap = (p.kw - tap0) < per_chunk ? (p.kw - tap0) : per_chunk;
            const uint32_t baddr = base + NS * p.a_bytes + (uint32_t)sb * kXBytes + (uint32_t)(k16 * 16 + tap0) * 128u;
            const uint64_t bd = desc_at_lbo(kDescHi128, baddr, 128u);
            mtab[nm++] = make_uint4((uint32_t)ad, (uint32_t)bd, idesc_bf16_f32(128, ntap * 64, 1, 1), (uint32_t)(tap0 * 64));
          }
        }
    const uint32_t a_hi = (uint32_t)(smem_desc_sw128(0, a_lbo, 1024) >> 32);
    const int first_n = n_chunk;                                  // the first MMA of every accumulator column range
    for (int it = 0; it < nst; ++it) {
      const int slot = it % p.depth;
      mbar_wait(smem_u32(&bk->full[slot]), (uint32_t)(it / p.depth) & 1u);
      tc_fence_after();
      const uint32_t soff = ((uint32_t)slot * p.stage_bytes) >> 4;   // start-address field is (address >> 4), 14 bits
#pragma unroll 4
      for (int i = 0; i < nm; ++i) {
        const uint4 m = mtab[i];
        umma_bf16(tmem_d + m.w, ((uint64_t)a_hi << 32) | (uint64_t)(m.x + soff), ((uint64_t)kDescHi128 << 32) | (uint64_t)(m.y + soff),
                  m.z, (it == 0 && i < first_n) ? 0u : 1u);
      }
      umma_commit(smem_u32(&bk->empty[slot]));
    }
    umma_commit(smem_u32(&bk->done));
  }
  __syncwarp();

  // ---------------- epilogue: D[n][tx*64 + ci] -> gw[(ty*kw+tx)][c0+ci][n0+n], lanes = consecutive n ----------------
  if (nst > 0) {
    mbar_wait(smem_u32(&bk->done), 0);
    tc_fence_after();
    const int nl = warp * 32 + lane, n = n0 + nl;
    const bool n_ok = nl < p.nbw && n < p.Cout;
    const uint32_t t_row = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < p.kw * 64; col += 16) {
      const int tx = col >> 6, ci0 = c0 + (col & 63);
      if (ci0 >= p.Cin) continue;                                  // warp-uniform
      float v[16];
      tmem_ld16(t_row + (uint32_t)col, v);
      tmem_ld_wait();
      if (n_ok) {
        float* dst = p.gw + ((size_t)(ty * p.kw + tx) * p.Cin + ci0) * p.Cout + n;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (ci0 + j < p.Cin) atomicAdd(dst + (size_t)j * p.Cout, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, tmem_cols);
}

// fp32 NHWC tensor -> bf16 term planes [NS][P][Cp] in the zero-padded frame (pads and channels >= C are zeros)
// colsum (optional, C floats): += column sums of src (the bias gradient rides on the dY conversion).  The grid stride is
// a multiple of Cp / 8, so a thread keeps its 8-channel group for every position it visits.
template <int NS>
__global__ void __launch_bounds__(256) planes_kernel(const float* __restrict__ src, long long img_stride, int pix_stride,
                                                     int C, int Cp, int H, int W, int Hp, int Wp, long long P, int vec4,
                                                     __nv_bfloat16* __restrict__ out, float* __restrict__ colsum) {
  const int c8n = Cp / 8;
  const long long total = P * c8n;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pos = i / c8n;
    const long long img = pos / ((long long)Hp * Wp);
    const int r = (int)(pos - img * Hp * Wp), y = r / Wp, x = r - y * Wp;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    if (y < H && x < W) {
      const float* s = src + img * img_stride + (long long)(y * W + x) * pix_stride + c8 * 8;
      if (vec4 && c8 * 8 + 8 <= C) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c8 * 8 + j < C) v[j] = __ldg(s + j);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
    uint2 lo[NS], hi[NS];
    split4<NS>(make_float4(v[0], v[1], v[2], v[3]), lo);
    split4<NS>(make_float4(v[4], v[5], v[6], v[7]), hi);
#pragma unroll
    for (int s = 0; s < NS; ++s)
      *reinterpret_cast<uint4*>(out + ((size_t)s * P + pos) * Cp + c8 * 8) = make_uint4(lo[s].x, lo[s].y, hi[s].x, hi[s].y);
  }
  if (colsum) {                                     // block-level sums in shared memory, one red.global.add per channel
    __shared__ float csum[kMaxColsumC];
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) csum[c] = 0.0f;
    __syncthreads();
    if (i0 < total) {
      const int c8 = (int)(i0 % c8n);
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&csum[c8 * 8 + j], acc[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(colsum + c, csum[c]);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)sym;
  }
  return fn;
}

struct WpPlan {
  WpParams p;
  int Hp, Wp, Cin_p, Cout_p, NS;
  long long P;
  size_t x_plane_bytes, y_plane_bytes, smem_bytes;
};

int wp_plan(const TcWgrad& c, WpPlan* out) {
  if (c.math < 1 || c.math > 3 || !c.gw) return FOV_ERR_UNSUPPORTED;
  if (c.T_inner != 1 || c.dil_h != 1 || c.dil_w != 1 || c.kw < 2 || c.kw > 8 || c.kh < 1 || c.kh > 8) return FOV_ERR_UNSUPPORTED;
  if (c.kw + kStagePos - 1 > kXRows) return FOV_ERR_UNSUPPORTED;
  if (c.Cin < 32 || (long long)c.N_img * c.H * c.W < 4096) return FOV_ERR_UNSUPPORTED;    // small problems: wgrad_tc.cu
  WpPlan pl{};
  const int ph_hi = (c.kh - 1) - c.pad_h, pw_hi = (c.kw - 1) - c.pad_w;
  if (c.pad_h < 0 || c.pad_w < 0 || ph_hi < 0 || pw_hi < 0) return FOV_ERR_UNSUPPORTED;
  // shared zero rows / columns: wide enough for the larger of the low and high reach of the kernel
  pl.Hp = c.H + (c.pad_h > ph_hi ? c.pad_h : ph_hi);
  pl.Wp = c.W + (c.pad_w > pw_hi ? c.pad_w : pw_hi);
  pl.P = (long long)c.N_img * pl.Hp * pl.Wp;
  if (pl.P >= (1LL << 31) - 4096) return FOV_ERR_UNSUPPORTED;
  pl.Cin_p = (c.Cin + 63) / 64 * 64;
  pl.Cout_p = (c.Cout + 63) / 64 * 64;
  pl.NS = c.math;
  WpParams& p = pl.p;
  p.Cin = c.Cin; p.Cout = c.Cout; p.kh = c.kh; p.kw = c.kw; p.pad_h = c.pad_h; p.pad_w = c.pad_w; p.Wp = pl.Wp;
  p.n_cg = pl.Cin_p / 64;
  p.nbw = pl.Cout_p >= 128 ? 128 : 64;
  p.a_groups = p.nbw / 64;
  p.n_nb = (pl.Cout_p + p.nbw - 1) / p.nbw;
  p.a_bytes = (uint32_t)p.a_groups * kAGroupBytes;
  p.stage_bytes = (uint32_t)pl.NS * (p.a_bytes + kXBytes);
  int depth = (int)((200 * 1024) / p.stage_bytes);
  if (depth > kMaxStages) depth = kMaxStages;
  if (depth < 2) return FOV_ERR_UNSUPPORTED;
  p.depth = depth;
  p.stages_total = (int)((pl.P + kStagePos - 1) / kStagePos);
  // pixel splits: fill the SMs in whole waves without multiplying the red.global.add traffic more than needed
  const int items = p.kh * p.n_cg * p.n_nb, sms = fov_num_sms();
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 16; ++s) {
    const int per = (p.stages_total + s - 1) / s;
    if (per < 8 && s > 1) break;
    const double waves = (double)(((long long)items * s + sms - 1) / sms);
    const double cost = waves * (per + 30.0);      // ~30 stage times per CTA for its 128 x kw*64 red.global.add epilogue
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  p.splits = best;
  p.stages_per_split = (p.stages_total + best - 1) / best;
  pl.x_plane_bytes = (size_t)pl.NS * pl.P * pl.Cin_p * 2;
  pl.y_plane_bytes = (size_t)pl.NS * pl.P * pl.Cout_p * 2;
  pl.smem_bytes = (size_t)depth * p.stage_bytes + sizeof(WpBook) + 1024;
  *out = pl;
  return FOV_OK;
}

int encode_plane(CUtensorMap* map, void* base, int Cp, long long P, int NS, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { fov_set_error("wgrad_planes: cuTensorMapEncodeTiled is not available"); return FOV_ERR_UNSUPPORTED; }
  const cuuint64_t gdim[3] = {(cuuint64_t)Cp, (cuuint64_t)P, (cuuint64_t)NS};
  const cuuint64_t gstr[2] = {(cuuint64_t)Cp * 2, (cuuint64_t)P * Cp * 2};
  const cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult cr = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { fov_set_error("wgrad_planes: cuTensorMapEncodeTiled failed (%d)", (int)cr); return FOV_ERR_CUDA; }
  return FOV_OK;
}

template <int NS>
int wp_launch(const CUtensorMap& xm, const CUtensorMap& ym, const WpPlan& pl, cudaStream_t st) {
  static FovPerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_planes_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      fov_set_error("wgrad_planes: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FOV_ERR_CUDA;
    }
    configured.mark();
  }
  const int items = pl.p.kh * pl.p.n_cg * pl.p.n_nb;
  wgrad_planes_kernel<NS><<<items * pl.p.splits, kThr, pl.smem_bytes, st>>>(xm, ym, pl.p);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

template <int NS>
int planes_launch(const float* src, long long img_stride, int pix_stride, int C, int Cp, int H, int W, const WpPlan& pl,
                  __nv_bfloat16* out, float* colsum, cudaStream_t st) {
  const long long total = pl.P * (Cp / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * fov_num_sms()) blocks = 8LL * fov_num_sms();
  // grid stride = blocks * 256 threads must be a multiple of Cp / 8 (fixed channel group per thread: column sums)
  const int c8n = Cp / 8;
  if (colsum && blocks * 256 % c8n != 0) {
    long long lcm_blocks = c8n;                       // c8n * 256 threads is always a multiple of c8n
    blocks = (blocks + lcm_blocks - 1) / lcm_blocks * lcm_blocks;
  }
  const int vec4 = ((uintptr_t)src % 16 == 0) && (img_stride % 4 == 0) && (pix_stride % 4 == 0);
  planes_kernel<NS><<<(int)blocks, 256, 0, st>>>(src, img_stride, pix_stride, C, Cp, H, W, pl.Hp, pl.Wp, pl.P, vec4, out,
                                                 colsum);
  FOV_CUDA_LAUNCH_CHECK();
  return FOV_OK;
}

}  // namespace

static int g_wp_disable = 0;
extern "C" void fov_debug_wgrad_planes(int enable) { g_wp_disable = !enable; }

// bytes of workspace the TMA-fed weight gradient wants for this problem (0: the shape takes the general kernel)
size_t tc_wgrad_planes_ws_bytes(const TcWgrad& c) {
  WpPlan pl;
  if (g_wp_disable || wp_plan(c, &pl) != FOV_OK || !get_encode()) return 0;
  return pl.x_plane_bytes + pl.y_plane_bytes + 2048;
}

// gw += x^T (*) dy over the planes; c.gbias (optional) += column sums of dy
int tc_wgrad_planes_run(const TcWgrad& c, void* ws, cudaStream_t st) {
  WpPlan pl;
  int rc = wp_plan(c, &pl);
  if (rc) { fov_set_error("wgrad_planes: unsupported shape"); return rc; }
  FOV_CHECK_ARG(ws && c.x && c.dy && c.gw, "NULL pointer");
  pl.p.gw = c.gw;
  uint8_t* w0 = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(w0);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(w0 + ((pl.x_plane_bytes + 1023) & ~(size_t)1023));
  auto run = [&](auto tag) -> int {
    constexpr int NS = decltype(tag)::value;
    int r = planes_launch<NS>(c.x, c.x_outer, c.x_pix_stride, c.Cin, pl.Cin_p, c.H, c.W, pl, xp, nullptr, st);
    if (r) return r;
    if ((r = planes_launch<NS>(c.dy, c.dy_outer, c.dy_pix_stride, c.Cout, pl.Cout_p, c.H, c.W, pl, yp,
                               pl.Cout_p <= kMaxColsumC ? c.gbias : nullptr, st))) return r;
    CUtensorMap xm, ym;
    if ((r = encode_plane(&xm, xp, pl.Cin_p, pl.P, NS, kXRows))) return r;
    if ((r = encode_plane(&ym, yp, pl.Cout_p, pl.P, NS, kStagePos))) return r;
    return wp_launch<NS>(xm, ym, pl, st);
  };
  switch (c.math) {
    case 1: return run(std::integral_constant<int, 1>{});
    case 2: return run(std::integral_constant<int, 2>{});
    default: return run(std::integral_constant<int, 3>{});
  }
}
