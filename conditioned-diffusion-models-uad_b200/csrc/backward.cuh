// Bandwidth-bound kernels of the UNet BACKWARD pass (training step, DDPM_2D.py:114-138 -> loss.backward()):
// GroupNorm(+FiLM +SiLU) backward, resampling backward, bias / stem / head gradients and the small fp32 linears of
// the embedding path.  Activations and activation gradients are NHWC 16-bit; parameter gradients are fp32.
#pragma once
#include "common.h"
#include "elementwise.cuh"

namespace cddpm {

// Backward of y = act(GN32(x) * (1 + scale) + shift) (the forward is launch_gn_apply without resampling):
//   out0 | out1 = dL/dx (+ add0 + add1), split at the concat boundary of x; dgamma, dbeta, dfilm accumulate.
struct GnBwdArgs {
  CatView x;
  int B = 0, H = 0, W = 0;
  const double* stats0 = nullptr;  // the forward's (sum, sumsq) buckets, [B][c0/4][2] (and [B][c1/4][2])
  const double* stats1 = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  const float* film = nullptr;  // forward FiLM output [B][film_stride] (scale at film_off + c, shift at film_off + C + c)
  int film_stride = 0;
  int film_off = 0;
  int silu = 1;
  const void* dy = nullptr;    // [B,H,W,C] gradient of the normalised, activated tensor
  const void* add0 = nullptr;  // optional [B,H,W,C] gradients that reach x along other paths (skip connections)
  const void* add1 = nullptr;
  float* sums = nullptr;       // workspace [B][C][2] fp32, zeroed by the caller
  void* out0 = nullptr;        // [B,H,W,c0]
  void* out1 = nullptr;        // [B,H,W,c1] (concat views only)
  float* bsum0 = nullptr;      // optional [c0] += column sums of out0 (the bias gradient of out0's producer)
  float* bsum1 = nullptr;
  float* dgamma = nullptr;     // [C] +=
  float* dbeta = nullptr;      // [C] +=
  float* dfilm = nullptr;      // optional [B][film_stride]: dscale at film_off + c, dshift at film_off + C + c
  int fmt = 1;
};
int launch_gn_bwd(const GnBwdArgs& a, cudaStream_t stream);

// Backward of the ResBlock resampling of launch_gn_apply: `dy` is [B,Ho,Wo,C], `dx` is [B,H,W,C] (H, W = the
// un-resampled size).  mode kResampleUp2 (forward nearest x2): dx = sum of the 2x2 block; kResampleDown2 (forward
// 2x2 average): dx = dy / 4 broadcast.
int launch_resample_bwd(const void* dy, void* dx, int B, int H, int W, int C, int mode, int fmt, cudaStream_t stream);

// out[c] += sum over rows of x[row][c]  (x is [rows][C] 16-bit; bias gradient of a convolution).
int launch_col_sum(const void* x, int64_t rows, int C, float* out, int fmt, cudaStream_t stream);
// out[0] += sum of x[0..n)
int launch_sum_f32(const float* x, int64_t n, float* out, cudaStream_t stream);

// out[c][tap] += sum_{n,y,x} act[n,y,x,c] * img[n, y + sgn*dy(tap), x + sgn*dx(tap)]   (zero outside the image)
// sgn = +1: weight gradient of the 1 -> C stem convolution (act = dL/d stem output, img = the network input);
// sgn = -1: weight gradient of the C -> 1 head convolution (act = the head's input, img = dL/d output).
int launch_wgrad_1ch(const void* act, const float* img, float* out, int B, int H, int W, int C, int sgn, int fmt,
                     cudaStream_t stream);
// Data gradient of the C -> 1 head convolution: dact[n,y,x,c] = sum_tap dout[n, y - dy(tap), x - dx(tap)] * w[c][tap].
int launch_head_bwd_data(const float* dout, const float* w, void* dact, int B, int H, int W, int C, int fmt,
                         cudaStream_t stream);

// nn.Linear backward on fp32 rows.  W is [O][I]: fp32 (w16 == nullptr) or the 16-bit K-major panel.
//   dx[b][i] = (sum_o dy[b][o] * W[o][i]) * (z ? silu'(z[b][i]) : 1)
int launch_linear_bwd_input(const float* dy, int dy_stride, const float* w32, const void* w16, int fmt, float* dx,
                            int dx_stride, const float* z, int z_stride, int B, int I, int O, cudaStream_t stream);
//   dw[o][i] = sum_b dy[b][o] * act(x[b][i]);  db[o] = sum_b dy[b][o]     (act: 0 none, 1 SiLU)
int launch_linear_bwd_weight(const float* dy, int dy_stride, const float* x, int x_stride, int act_x, float* dw,
                             float* db, int B, int I, int O, cudaStream_t stream);
// Same over the row-wise concatenation of several Linear layers (the FiLM projection): rows [32 t, 32 t + 32) of the
// weight gradient go to grads + tile_w_off[t] (row-major [..][I]), their bias gradients to grads + tile_b_off[t].
int launch_linear_bwd_weight_tiled(const float* dy, int dy_stride, const float* x, int x_stride, int act_x,
                                   float* grads, const int64_t* tile_w_off, const int64_t* tile_b_off, int B, int I,
                                   int O, cudaStream_t stream);
// dx[i] *= silu'(z[i])
int launch_mul_silu_grad(float* dx, const float* z, int64_t n, cudaStream_t stream);
// torch.optim.Adam(lr, betas, eps) over many tensors in ONE launch (DDPM_2D.configure_optimizers, DDPM_2D.py:305-306):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// p/g/m/v: device arrays of per-tensor device pointers (a NULL gradient skips the tensor), numel per tensor;
// block b works on elements [block_off[b], block_off[b] + 4096) of tensor block_tensor[b].
int launch_adam_step(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel,
                     const int* block_tensor, const int64_t* block_off, int total_blocks, float lr, float beta1,
                     float beta2, float eps, float bc1, float bc2, cudaStream_t stream);
// dst[i] = src[i] for i < n (fp32), plain device copy helper that is graph-capturable
int launch_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t stream);

}  // namespace cddpm
