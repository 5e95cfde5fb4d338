// DDPM arithmetic around the UNet (see diffusion.cu for the reference call sites).
#pragma once
#include "common.h"

namespace cddpm {

// x_t = sqrt_ac[t_b] * x0 + sqrt_1mac[t_b] * noise, x0 = normalize ? 2*img-1 : img.  t: [B] or [1] (t_shared).
int launch_q_sample(const float* img, const void* noise, int noise_f16, float* out, const float* sqrt_ac,
                    const float* sqrt_1mac, const int64_t* t, int t_shared, int B, int HW, int normalize,
                    cudaStream_t stream);

// x_{t-1} = coef1[t]*clamp(x0_hat,-1,1) + coef2[t]*x_t + (t>0 ? exp(logvar[t]/2)*noise : 0); noise may be NULL.
// clip_denoised == 0 skips the clamp (p_mean_variance, cond_DDPM.py:422-430).
// x0_hat = model_out (pred_x0) or sqrt_recip_ac[t]*x_t - sqrt_recipm1_ac[t]*model_out (pred_noise).
// final_unnormalize additionally maps the result through (v+1)/2 (last step of p_sample_loop).
int launch_posterior_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_prev,
                          const float* coef1, const float* coef2, const float* logvar, const float* sqrt_recip_ac,
                          const float* sqrt_recipm1_ac, int64_t t, int B, int HW, int pred_noise, int clip_denoised,
                          int final_unnormalize, cudaStream_t stream);

// One DDIM update (cond_DDPM.py:487-511) with host-computed step scalars; noise may be NULL (time_next == 0).
int launch_ddim_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_next,
                     float sqrt_recip, float sqrt_recipm1, float sqrt_alpha_next, float c, float sigma, int B, int HW,
                     int pred_noise, int clip_denoised, int final_unnormalize, cudaStream_t stream);

// reco = reco*beta + alpha * ((model_out+1)/2  |  (x_t - sqrt_1mac[t]*model_out + 1)/2);
// loss[b] = mean_i |model_out - target| (or squared) * p2w[t_b], target = 2*img-1 | noise.
int launch_recon_finish(const float* model_out, const float* img, const float* x_t, const void* noise, int noise_f16,
                        float* reco, float reco_alpha, float reco_beta, float* loss, const float* sqrt_1mac,
                        const float* p2w, const int64_t* t, int t_shared, int B, int HW, int pred_noise, int l2,
                        cudaStream_t stream);

// dout = grad_loss[0] * d/d model_out of mean_b(loss[b]) with loss[b] as in launch_recon_finish (training step).
int launch_loss_backward(const float* model_out, const float* img, const void* noise, int noise_f16, const float* p2w,
                         const int64_t* t, const float* grad_loss, float* dout, int B, int HW, int pred_noise, int l2,
                         cudaStream_t stream);

}  // namespace cddpm
