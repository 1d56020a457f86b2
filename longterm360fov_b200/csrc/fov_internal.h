// Internal (non-exported) launch helpers shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include "../../include/fov360.h"

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
