/* libcddpm_b200 — C ABI of the B200-native cDDPM reconstruction + anomaly-scoring path.
 *
 * The reference (raymondfdavey/Conditioned-Diffusion-Models-UAD) has no native code and no FFI: its only stable
 * seam is the Python module surface.  This header is therefore the boundary our Python drop-ins
 * (src/models/DDPM_2D.py, src/models/modules/{OpenAI_Unet,cond_DDPM,DDPM_encoder}.py, src/utils/utils_eval.py)
 * bind through ctypes.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the library never takes ownership;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is stream-ordered;
 *   - return value: 0 = ok, non-zero = error (see cddpm_last_error());
 *   - activations are NHWC 16-bit (bf16 by default) inside the engine; public tensors keep the reference layouts
 *     (NCHW fp32 images, [B,C] fp32 vectors, [H,W,D] fp32 volumes).
 */
#ifndef CDDPM_B200_H_
#define CDDPM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDDPM_OK 0
#define CDDPM_INVALID_ARGUMENT 1
#define CDDPM_CUDA_ERROR 2
#define CDDPM_UNSUPPORTED 3
#define CDDPM_NOT_READY 4

#define CDDPM_FMT_F16 0
#define CDDPM_FMT_BF16 1

/* Message of the last failing call on this thread. */
const char* cddpm_last_error(void);
/* Library version string, e.g. "cddpm_b200 0.1 (sm_100a)". */
const char* cddpm_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution (nn.Conv2d 3x3 pad 1 / 1x1, stride 1): OpenAI_Unet.py:231,257,268 (ResBlock in/out/skip convs),
 * :367,375 (attention qkv / proj_out as 1x1), util.py:218 (conv_nd).
 * ---------------------------------------------------------------------------------------------------------- */

/* Re-layout input-channel slice [cin_off, cin_off+c_s) of an OIHW fp32 weight [cout][cin_total][k][k] into the
 * packed 16-bit matrix wpacked[cout][ktot] at column offset koff, column order (tap, ci). */
int cddpm_pack_conv_weight(const float* w_oihw, int cout, int cin_total, int ksize, int cin_off, int c_s,
                           void* wpacked, int ktot, int koff, int fmt, void* stream);

/* out[B,H,W,cout] = bias + residual + sum over sources s of conv(src[s] ([B,H,W,src_c[s]], 16-bit NHWC), taps 9|1).
 * K order of wpacked = sources in order, each (tap, ci).  tcgen05 implicit GEMM; H, W multiples of 8,
 * src_c multiples of 64, cout multiple of 32.  bias/residual may be NULL.  out is 16-bit, or fp32 if out_f32. */
int cddpm_conv_igemm(int num_src, const void* const* src, const int* src_c, const int* src_taps, int B, int H,
                     int W, int cout, const void* wpacked, const float* bias, const void* residual, void* out,
                     int out_f32, int fmt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDDPM_B200_H_ */
