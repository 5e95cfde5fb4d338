// Backward of the implicit-GEMM convolutions (training step, DDPM_2D.py:114-138 -> loss.backward()).
//
// Data gradient: a stride-1 "same" convolution's data gradient is the same convolution with the kernel flipped and
// its channel roles exchanged, so it runs on the FORWARD kernels (conv_igemm2.cu) against a second packed panel
// (launch_pack_conv_weight_T).  Weight gradient: its own tcgen05 kernel (conv_wgrad.cu).
#pragma once
#include "common.h"
#include "conv_igemm.cuh"

namespace cddpm {

// Packed panel of the data-gradient convolution of `w_oihw` [Cout][Cin_total][k][k]:
//   wpacked_t[ci - cin_off][koff + tap' * Cout + co] = w[co][ci][k*k - 1 - tap']      (ci in [cin_off, cin_off + C_s))
// i.e. a convolution with C_s output channels over a Cout-channel source with `ksize*ksize` taps.
int launch_pack_conv_weight_T(const float* w_oihw, int Cout, int Cin_total, int ksize, int cin_off, int C_s,
                              void* wpacked_t, int Ktot, int koff, int ab_format, cudaStream_t stream);

// Weight gradient of out = sum_s conv(src_s): for every source s, tap and channel
//   dw[co][koff_s + tap * C_s + ci] += sum_{n,y,x} dy[n,y,x,co] * src_s[n, y + dy(tap), x + dx(tap), ci]
// in the K order of the packed forward panel ([Cout][Ktot] fp32, accumulated with atomics: zero it first).
// Sources marked skip (e.g. the identity block of a ResBlock's second convolution) produce nothing.
struct WgradDesc {
  int num_src = 0;
  const void* src[kConvMaxSrc] = {nullptr, nullptr, nullptr};
  int src_c[kConvMaxSrc] = {0, 0, 0};
  int src_taps[kConvMaxSrc] = {0, 0, 0};
  int src_skip[kConvMaxSrc] = {0, 0, 0};
  const void* dy = nullptr;  // [B,H,W,Cout] 16-bit NHWC
  int B = 0, H = 0, W = 0, Cout = 0;
  float* dw = nullptr;       // [Cout][Ktot] fp32
  int ab_format = 1;
};
bool wgrad_supported(const WgradDesc& d);
int build_wgrad(const WgradDesc& d, std::shared_ptr<void>* holder);
int launch_wgrad(const std::shared_ptr<void>& holder, cudaStream_t stream);
int64_t wgrad_flops(const WgradDesc& d);

// grad_oihw[co][cin_off + ci][tap] (=|+=) dw_packed[co][koff + tap * C_s + ci]  (inverse of launch_pack_conv_weight)
int launch_unpack_conv_grad(const float* dw_packed, int Cout, int Cin_total, int ksize, int cin_off, int C_s,
                            float* grad_oihw, int Ktot, int koff, cudaStream_t stream);

}  // namespace cddpm
