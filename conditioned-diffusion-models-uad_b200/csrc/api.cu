// extern "C" surface of libcddpm_b200 (declared in include/cddpm_b200.h).  Thin argument checking + dispatch.
#include "../../include/cddpm_b200.h"

#include "common.h"
#include "conv_igemm.cuh"

namespace cddpm {
const char* last_error_cstr();
}

using namespace cddpm;

extern "C" {

const char* cddpm_last_error(void) { return last_error_cstr(); }
const char* cddpm_version(void) { return "cddpm_b200 0.1 (sm_100a)"; }

int cddpm_pack_conv_weight(const float* w_oihw, int cout, int cin_total, int ksize, int cin_off, int c_s,
                           void* wpacked, int ktot, int koff, int fmt, void* stream) {
  if (!w_oihw || !wpacked) return fail(kInvalidArgument, "pack_conv_weight: null pointer");
  if (ksize != 1 && ksize != 3) return fail(kInvalidArgument, "pack_conv_weight: ksize must be 1 or 3");
  if (cin_off < 0 || cin_off + c_s > cin_total || koff < 0 || koff + ksize * ksize * c_s > ktot)
    return fail(kInvalidArgument, "pack_conv_weight: slice out of range");
  return launch_pack_conv_weight(w_oihw, cout, cin_total, ksize, cin_off, c_s, wpacked, ktot, koff, fmt,
                                 static_cast<cudaStream_t>(stream));
}

int cddpm_conv_igemm(int num_src, const void* const* src, const int* src_c, const int* src_taps, int B, int H,
                     int W, int cout, const void* wpacked, const float* bias, const void* residual, void* out,
                     int out_f32, int fmt, void* stream) {
  if (!src || !src_c || !src_taps || !wpacked || !out) return fail(kInvalidArgument, "conv_igemm: null pointer");
  if (num_src < 1 || num_src > kConvMaxSrc) return fail(kInvalidArgument, "conv_igemm: num_src must be 1..3");
  ConvDesc d;
  d.num_src = num_src;
  for (int s = 0; s < num_src; ++s) {
    d.src[s] = src[s];
    d.src_c[s] = src_c[s];
    d.src_taps[s] = src_taps[s];
  }
  d.B = B;
  d.H = H;
  d.W = W;
  d.Cout = cout;
  d.wpacked = wpacked;
  d.bias = bias;
  d.residual = residual;
  d.out = out;
  d.out_is_f32 = out_f32;
  d.ab_format = fmt;
  ConvIgemmParams p;
  CDDPM_TRY(build_conv_params(d, &p));
  return launch_conv_igemm(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
