// Fused, vectorised elementwise kernels of the DDPM arithmetic around the UNet (GaussianDiffusion):
//   q_sample            src/models/modules/cond_DDPM.py:548-554 (+ normalize_to_neg_one_to_one :75, :653)
//   p_losses tail       :636-645  (per-sample L1/L2 loss, reco = unnormalize(model_out) or the pred_noise variant)
//   p_sample            :432-444 with model_predictions :400-420 and q_posterior :391-398
// Images are [B,1,H,W] fp32 (HW contiguous); schedule buffers are the module's fp32 [T] device arrays.
#include "diffusion.cuh"

#include <cuda_fp16.h>

namespace cddpm {

namespace {

__device__ __forceinline__ float load_noise(const void* noise, int is_f16, size_t i) {
  if (is_f16) return __half2float(reinterpret_cast<const __half*>(noise)[i]);
  return reinterpret_cast<const float*>(noise)[i];
}

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ img, const void* __restrict__ noise,
                                                       int noise_f16, float* __restrict__ out,
                                                       const float* __restrict__ sqrt_ac,
                                                       const float* __restrict__ sqrt_1mac,
                                                       const int64_t* __restrict__ t, int t_shared, int HW,
                                                       int normalize) {
  const int b = blockIdx.y;
  const int64_t tb = t[t_shared ? 0 : b];
  const float a = sqrt_ac[tb], s = sqrt_1mac[tb];
  const size_t base = static_cast<size_t>(b) * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float x0 = img[base + i];
    if (normalize) x0 = x0 * 2.0f - 1.0f;
    // same association as the reference: (a * x0) + (s * noise), each product rounded to fp32 first
    out[base + i] = __fadd_rn(__fmul_rn(a, x0), __fmul_rn(s, load_noise(noise, noise_f16, base + i)));
  }
}

__global__ void __launch_bounds__(256) posterior_step_kernel(
    const float* __restrict__ model_out, const float* __restrict__ x_t, const void* __restrict__ noise, int noise_f16,
    float* __restrict__ x_prev, const float* __restrict__ coef1, const float* __restrict__ coef2,
    const float* __restrict__ logvar, const float* __restrict__ sqrt_recip_ac,
    const float* __restrict__ sqrt_recipm1_ac, int64_t t, int HW, int pred_noise, int clip_denoised,
    int final_unnormalize) {
  const int b = blockIdx.y;
  const float c1 = coef1[t], c2 = coef2[t];
  const float sigma = (t > 0 && noise != nullptr) ? expf(0.5f * logvar[t]) : 0.f;
  const float sr = pred_noise ? sqrt_recip_ac[t] : 0.f, srm1 = pred_noise ? sqrt_recipm1_ac[t] : 0.f;
  const size_t base = static_cast<size_t>(b) * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float xt = x_t[base + i];
    float x0 = model_out[base + i];
    if (pred_noise) x0 = __fsub_rn(__fmul_rn(sr, xt), __fmul_rn(srm1, x0));
    if (clip_denoised) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    float v = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt));
    if (sigma != 0.f) v = __fadd_rn(v, __fmul_rn(sigma, load_noise(noise, noise_f16, base + i)));
    if (final_unnormalize) v = (v + 1.0f) * 0.5f;
    x_prev[base + i] = v;
  }
}

// One DDIM update (cond_DDPM.py:487-511).  The step's scalars are computed by the host in fp32 exactly as the reference
// computes them from its 0-dim tensors; the association below is the reference's: model_predictions (:400-420, no
// clipping inside), x_start.clamp_ (:495-496, AFTER pred_noise was derived from the unclamped prediction), then
// ((x_start * sqrt(alpha_next)) + (c * pred_noise)) + (sigma * noise).
__global__ void __launch_bounds__(256) ddim_step_kernel(const float* __restrict__ model_out,
                                                        const float* __restrict__ x_t, const void* __restrict__ noise,
                                                        int noise_f16, float* __restrict__ x_next, float sqrt_recip,
                                                        float sqrt_recipm1, float sqrt_alpha_next, float c, float sigma,
                                                        int HW, int pred_noise, int clip_denoised,
                                                        int final_unnormalize) {
  const int b = blockIdx.y;
  const size_t base = static_cast<size_t>(b) * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float xt = x_t[base + i];
    const float o = model_out[base + i];
    float x0, pn;
    if (pred_noise) {
      pn = o;
      x0 = __fsub_rn(__fmul_rn(sqrt_recip, xt), __fmul_rn(sqrt_recipm1, o));
    } else {
      pn = __fdiv_rn(__fsub_rn(__fmul_rn(sqrt_recip, xt), o), sqrt_recipm1);
      x0 = o;
    }
    if (clip_denoised) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    float v = __fadd_rn(__fmul_rn(x0, sqrt_alpha_next), __fmul_rn(c, pn));
    // the reference adds sigma * noise even when noise is the Python float 0. (a zero tensor): v + 0 == v
    if (noise != nullptr) v = __fadd_rn(v, __fmul_rn(sigma, load_noise(noise, noise_f16, base + i)));
    if (final_unnormalize) v = (v + 1.0f) * 0.5f;
    x_next[base + i] = v;
  }
}

// One CTA per sample: reco and the per-sample mean loss (deterministic tree reduction).
__global__ void __launch_bounds__(256) recon_finish_kernel(
    const float* __restrict__ model_out, const float* __restrict__ img, const float* __restrict__ x_t,
    const void* __restrict__ noise, int noise_f16, float* __restrict__ reco, float reco_alpha, float reco_beta,
    float* __restrict__ loss, const float* __restrict__ sqrt_1mac, const float* __restrict__ p2w,
    const int64_t* __restrict__ t, int t_shared, int HW, int pred_noise, int l2) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  const int64_t tb = t[t_shared ? 0 : b];
  const float s = sqrt_1mac[tb];
  const size_t base = static_cast<size_t>(b) * HW;
  float acc = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float o = model_out[base + i];
    float target, r;
    if (pred_noise) {
      target = load_noise(noise, noise_f16, base + i);
      r = (__fsub_rn(x_t[base + i], __fmul_rn(s, o)) + 1.0f) * 0.5f;
    } else {
      target = img[base + i] * 2.0f - 1.0f;
      r = (o + 1.0f) * 0.5f;
    }
    const float d = o - target;
    acc += l2 ? d * d : fabsf(d);
    if (reco != nullptr) {
      const float prev = (reco_beta != 0.f) ? reco[base + i] * reco_beta : 0.f;
      reco[base + i] = prev + reco_alpha * r;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss != nullptr) loss[b] = red[0] / static_cast<float>(HW) * p2w[tb];
}

// d loss / d model_out for loss = mean_b(mean_i |o - target| (or squared) * p2w[t_b]), times the upstream scalar.
__global__ void __launch_bounds__(256) loss_backward_kernel(
    const float* __restrict__ model_out, const float* __restrict__ img, const void* __restrict__ noise, int noise_f16,
    const float* __restrict__ p2w, const int64_t* __restrict__ t, const float* __restrict__ grad_loss,
    float* __restrict__ dout, int B, int HW, int pred_noise, int l2) {
  const int b = blockIdx.y;
  const float scale = grad_loss[0] * p2w[t[b]] / (static_cast<float>(HW) * static_cast<float>(B));
  const size_t base = static_cast<size_t>(b) * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float o = model_out[base + i];
    const float target = pred_noise ? load_noise(noise, noise_f16, base + i) : img[base + i] * 2.0f - 1.0f;
    const float d = o - target;
    dout[base + i] = scale * (l2 ? 2.0f * d : (d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.0f)));
  }
}

}  // namespace

int launch_loss_backward(const float* model_out, const float* img, const void* noise, int noise_f16, const float* p2w,
                         const int64_t* t, const float* grad_loss, float* dout, int B, int HW, int pred_noise, int l2,
                         cudaStream_t stream) {
  if (!model_out || !p2w || !t || !grad_loss || !dout) return fail(kInvalidArgument, "loss_backward: null pointer");
  if (pred_noise ? !noise : !img) return fail(kInvalidArgument, "loss_backward: missing target");
  dim3 grid((HW + 1023) / 1024, B);
  loss_backward_kernel<<<grid, 256, 0, stream>>>(model_out, img, noise, noise_f16, p2w, t, grad_loss, dout, B, HW,
                                                 pred_noise, l2);
  return check_launch("loss_backward_kernel");
}

int launch_q_sample(const float* img, const void* noise, int noise_f16, float* out, const float* sqrt_ac,
                    const float* sqrt_1mac, const int64_t* t, int t_shared, int B, int HW, int normalize,
                    cudaStream_t stream) {
  if (!img || !noise || !out || !sqrt_ac || !sqrt_1mac || !t) return fail(kInvalidArgument, "q_sample: null pointer");
  dim3 grid((HW + 1023) / 1024, B);
  q_sample_kernel<<<grid, 256, 0, stream>>>(img, noise, noise_f16, out, sqrt_ac, sqrt_1mac, t, t_shared, HW, normalize);
  return check_launch("q_sample_kernel");
}

int launch_posterior_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_prev,
                          const float* coef1, const float* coef2, const float* logvar, const float* sqrt_recip_ac,
                          const float* sqrt_recipm1_ac, int64_t t, int B, int HW, int pred_noise, int clip_denoised,
                          int final_unnormalize, cudaStream_t stream) {
  if (!model_out || !x_t || !x_prev || !coef1 || !coef2 || !logvar)
    return fail(kInvalidArgument, "posterior_step: null pointer");
  if (pred_noise && (!sqrt_recip_ac || !sqrt_recipm1_ac))
    return fail(kInvalidArgument, "posterior_step: pred_noise needs sqrt_recip(m1)_alphas_cumprod");
  dim3 grid((HW + 1023) / 1024, B);
  posterior_step_kernel<<<grid, 256, 0, stream>>>(model_out, x_t, noise, noise_f16, x_prev, coef1, coef2, logvar,
                                                  sqrt_recip_ac, sqrt_recipm1_ac, t, HW, pred_noise, clip_denoised,
                                                  final_unnormalize);
  return check_launch("posterior_step_kernel");
}

int launch_ddim_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_next,
                     float sqrt_recip, float sqrt_recipm1, float sqrt_alpha_next, float c, float sigma, int B, int HW,
                     int pred_noise, int clip_denoised, int final_unnormalize, cudaStream_t stream) {
  if (!model_out || !x_t || !x_next) return fail(kInvalidArgument, "ddim_step: null pointer");
  if (B < 1 || HW < 1) return fail(kInvalidArgument, "ddim_step: empty batch");
  dim3 grid((HW + 1023) / 1024, B);
  ddim_step_kernel<<<grid, 256, 0, stream>>>(model_out, x_t, noise, noise_f16, x_next, sqrt_recip, sqrt_recipm1,
                                             sqrt_alpha_next, c, sigma, HW, pred_noise, clip_denoised, final_unnormalize);
  return check_launch("ddim_step_kernel");
}

int launch_recon_finish(const float* model_out, const float* img, const float* x_t, const void* noise, int noise_f16,
                        float* reco, float reco_alpha, float reco_beta, float* loss, const float* sqrt_1mac,
                        const float* p2w, const int64_t* t, int t_shared, int B, int HW, int pred_noise, int l2,
                        cudaStream_t stream) {
  if (!model_out || !img || !t || !sqrt_1mac || !p2w) return fail(kInvalidArgument, "recon_finish: null pointer");
  if (pred_noise && (!x_t || !noise)) return fail(kInvalidArgument, "recon_finish: pred_noise needs x_t and noise");
  recon_finish_kernel<<<B, 256, 0, stream>>>(model_out, img, x_t, noise, noise_f16, reco, reco_alpha, reco_beta, loss,
                                             sqrt_1mac, p2w, t, t_shared, HW, pred_noise, l2);
  return check_launch("recon_finish_kernel");
}

}  // namespace cddpm
