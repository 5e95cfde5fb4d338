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
/* Stream-ordered device-to-device copy (lets hosts without a CUDA runtime binding read engine-owned buffers). */
int cddpm_memcpy_d2d(void* dst, const void* src, int64_t nbytes, void* stream);

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
/* The same convolution (16-bit output) whose epilogue also ADDS the GroupNorm partial sums of its fp32 result into
 * gn_stats[B][cout/4][2] (float64: sum and sum of squares per image and 4-channel bucket; zeroed by the caller) - the
 * statistics util.py:214-216 (GroupNorm32) needs for the normalisation that follows every ResBlock convolution
 * (OpenAI_Unet.py:284-338), produced without a pass over the tensor.  Geometries of the macro-tile kernel only
 * (H, W multiples of 8, src_c multiples of 64, cout multiple of 128); others return an error. */
int cddpm_conv_igemm_stats(int num_src, const void* const* src, const int* src_c, const int* src_taps, int B, int H,
                           int W, int cout, const void* wpacked, const float* bias, const void* residual, void* out,
                           int fmt, double* gn_stats, void* stream);

/* Backward of the convolutions (training step: DDPM_2D.py:114-138 -> loss.backward(), torch autograd of nn.Conv2d).
 * Data gradient = the forward kernel over the transposed / flipped panel built here:
 *   wpacked_t[ci - cin_off][koff + tap' * cout + co] = w[co][ci][k*k - 1 - tap']. */
int cddpm_pack_conv_weight_t(const float* w_oihw, int cout, int cin_total, int ksize, int cin_off, int c_s,
                             void* wpacked_t, int ktot, int koff, int fmt, void* stream);
/* Weight gradient (tcgen05, both operands MN-major straight from NHWC): for every source s with skip[s] == 0
 *   dw[co][koff_s + tap * c_s + ci] += sum_{n,y,x} dy[n,y,x,co] * src_s[n, y + dy(tap), x + dx(tap), ci]
 * dw is [cout][sum_s taps_s * c_s] fp32 in the K order of the packed forward panel; the caller zeroes it.
 * cout and every src_c multiples of 128; H, W multiples of 8. */
int cddpm_conv_wgrad(int num_src, const void* const* src, const int* src_c, const int* src_taps, const int* skip,
                     const void* dy, int B, int H, int W, int cout, float* dw, int fmt, void* stream);
/* grad_oihw[co][cin_off + ci][tap] = dw_packed[co][koff + tap * c_s + ci] (inverse of cddpm_pack_conv_weight). */
int cddpm_unpack_conv_grad(const float* dw_packed, int cout, int cin_total, int ksize, int cin_off, int c_s,
                           float* grad_oihw, int ktot, int koff, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Bandwidth-bound UNet pieces, exposed one by one for the parity tests.
 * ---------------------------------------------------------------------------------------------------------- */

/* GroupNorm32 (util.py:214-216) over an NHWC 16-bit tensor that may be the channel concat of two tensors
 * (th.cat([h, hs.pop()], 1), OpenAI_Unet.py:948), fused with the FiLM modulation h*(1+scale)+shift
 * (OpenAI_Unet.py:325-331; film = [B][film_stride] fp32 with scale at film_off+c, shift at film_off+C+c, or NULL),
 * SiLU (silu != 0) and the ResBlock resampling (mode 0 none, 1 nearest x2 (Upsample :115-129), 2 avg-pool 2
 * (Downsample :172-179)).  raw_out (optional) receives the un-normalised input resampled the same way (x_upd).
 * workspace: at least cddpm_gn_workspace_floats(B, H*W) floats. */
int64_t cddpm_gn_workspace_floats(int B, int HW);
int cddpm_groupnorm_film_silu(const void* x0, int c0, const void* x1, int c1, int B, int H, int W,
                              const float* gamma, const float* beta, const float* film, int film_stride,
                              int film_off, int silu, int mode, void* out, void* raw_out, float* workspace, int fmt,
                              void* stream);

/* nn.Linear on fp32 rows with optional SiLU on the input and/or output (time_embed / label_emb / emb_layers,
 * OpenAI_Unet.py:583-602, :245-251). */
int cddpm_linear(const float* in, int in_stride, const float* w, const float* bias, float* out, int out_stride,
                 int B, int I, int O, int silu_in, int silu_out, void* stream);
/* timestep_embedding (util.py:151-171): emb[B][dim] = [cos(t f_i) | sin(t f_i)]. */
int cddpm_timestep_embedding(const int64_t* t, float* emb, int B, int dim, void* stream);

/* QKVAttention.forward (OpenAI_Unet.py:457-476) on the NHWC output of the qkv 1x1 conv: qkv [B,L,3C] -> [B,L,C]. */
int cddpm_attention(const void* qkv, void* out, int B, int L, int C, int fmt, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * UNet engine: UNetModel.__init__/forward (OpenAI_Unet.py:513-797, :823-1006) behind the reference's state_dict keys.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct cddpm_unet_config {
  int image_h, image_w;
  int in_channels, model_channels, out_channels;
  int num_res_blocks;
  int n_mult;
  int channel_mult[8];
  int n_attn_res;
  int attention_resolutions[8];
  int num_classes;       /* 0 = unconditioned (num_classes=None) */
  int num_head_channels; /* 64 */
  int fmt;               /* CDDPM_FMT_BF16 | CDDPM_FMT_F16: activation / tensor-core operand type */
} cddpm_unet_config;

typedef struct cddpm_unet cddpm_unet_t;

int cddpm_unet_create(const cddpm_unet_config* cfg, cddpm_unet_t** out);
void cddpm_unet_destroy(cddpm_unet_t* h);
/* Parameters in the reference's registration order; names are the reference's state_dict keys under diffusion.model. */
int cddpm_unet_param_count(const cddpm_unet_t* h);
int cddpm_unet_param_info(const cddpm_unet_t* h, int index, const char** name, int64_t* numel);
/* Copy / re-layout one fp32 parameter (device pointer, reference shape, contiguous) into the engine. */
int cddpm_unet_set_param(cddpm_unet_t* h, const char* name, const float* value, int64_t numel, void* stream);
/* Bulk form of cddpm_unet_set_param: values[i] (device pointer, fp32, reference shape) for parameter i in
 * cddpm_unet_param_info order; NULL entries are left unchanged.  One call per optimizer step. */
int cddpm_unet_set_params(cddpm_unet_t* h, const float* const* values, int count, void* stream);
/* model(x, timesteps, cond): x [B,1,H,W] fp32, t [B] int64, cond [B,num_classes] fp32 or NULL -> out [B,1,H,W] fp32. */
int cddpm_unet_forward(cddpm_unet_t* h, const float* x, const int64_t* t, const float* cond, float* out, int B,
                       void* stream);
/* Output buffer (NHWC 16-bit, valid until the next forward) of a layer of the last forward, by module path,
 * e.g. "input_blocks.3.0", "middle_block.1", "output_blocks.7.1"; "<res>/in_conv", "<attn>/qkv", "<attn>/attn". */
int cddpm_unet_tap(const cddpm_unet_t* h, const char* layer, void** ptr, int* C, int* H, int* W);
/* FiLM projections of the last forward: [B][stride] fp32 (all emb_layers outputs, concatenated in module order). */
int cddpm_unet_film(const cddpm_unet_t* h, const float** ptr, int* stride);
/* Algorithmic tensor-core FLOPs of the implicit-GEMM convolutions per sample, and kernel launches per forward. */
int64_t cddpm_unet_conv_flops(const cddpm_unet_t* h);
int cddpm_unet_launches(const cddpm_unet_t* h);
/* Measurement aid: bracket every convolution launch of the NEXT forward with CUDA events on its stream, then read
 * the summed device time (ms) and the launch count (this synchronises on the recorded events). */
int cddpm_unet_profile_arm(cddpm_unet_t* h);
int cddpm_unet_profile_read(cddpm_unet_t* h, double* conv_ms, int* conv_launches);

/* Training step (DDPM_2D.training_step, DDPM_2D.py:114-138: loss.backward() through UNetModel.forward).
 * Backward of the LAST forward of this handle (bf16 engines, model_channels 128): dout = dL/d out [B,1,H,W] fp32.
 * Every parameter gradient is written (not accumulated) into the flat fp32 buffer `grads` of
 * cddpm_unet_grad_total() floats: parameter i (cddpm_unet_param_info order) at cddpm_unet_grad_offset(i), in the
 * reference's parameter layout.  dcond (optional) receives dL/d cond [B,num_classes]. */
int64_t cddpm_unet_grad_total(const cddpm_unet_t* h);
int cddpm_unet_grad_offset(const cddpm_unet_t* h, int index, int64_t* offset);
int cddpm_unet_backward(cddpm_unet_t* h, const float* dout, float* grads, float* dcond, int B, void* stream);
/* training != 0: the engine keeps every intermediate cddpm_unet_backward reads and never plans an inference-only
 * fusion.  The one such fusion today is experimental and off unless CDDPM_FUSE_GN=1: the out_layers GroupNorm + FiLM +
 * SiLU of a ResBlock (OpenAI_Unet.py:287-296) finished inside the epilogue of the convolution that produces its input,
 * whose raw result is then never stored (a cross-CTA rendezvous per image: such forwards need the whole GPU and must not
 * overlap another spinning kernel); cddpm_unet_backward refuses a plan that contains it.  Switching modes drops the
 * current plan (the next forward re-plans). */
int cddpm_unet_set_training(cddpm_unet_t* h, int training);
/* Algorithmic tensor-core FLOPs of the backward per sample, and its kernel-launching plan steps (after the first
 * backward of a batch size). */
int64_t cddpm_unet_bwd_flops(const cddpm_unet_t* h);
int cddpm_unet_bwd_launches(const cddpm_unet_t* h);
/* torch.optim.Adam(lr, betas, eps) - DDPM_2D.configure_optimizers (DDPM_2D.py:305-306) - over many fp32 tensors in
 * ONE launch: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps), with the
 * bias corrections bc1 = 1 - b1^t, bc2 = 1 - b2^t passed by the host.  p, g, m, v: DEVICE arrays of per-tensor device
 * pointers (a NULL gradient skips that tensor); numel: device array; block b updates elements
 * [block_off[b], block_off[b] + 4096) of tensor block_tensor[b]. */
int cddpm_adam_step(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel,
                    const int* block_tensor, const int64_t* block_off, int total_blocks, float lr, float beta1,
                    float beta2, float eps, float bc1, float bc2, void* stream);
/* Backward of cddpm_attention (bf16): dqkv [B,L,3C] from dout [B,L,C]; scratch = cddpm_attention_bwd_scratch_bytes. */
int64_t cddpm_attention_bwd_scratch_bytes(int B, int L, int C);
int cddpm_attention_bwd(const void* qkv, const void* dout, void* dqkv, void* scratch, int B, int L, int C, int fmt,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Condition encoder: SparK_2D_encoder.forward (spark/Spark_2D.py:285-290) = timm ResNet-50 (v1.5, in_chans=1,
 * num_classes=cond_dim) in eval mode, forward(x, pyramid=0) (spark/resnet.py:13-46).  Parameter names are the
 * timm / torchvision keys under `encoder.encoder.` (conv1.weight, bn1.running_mean, layer3.0.downsample.1.bias, ...;
 * num_batches_tracked is not a parameter here).  BatchNorm is folded into the convolutions.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct cddpm_encoder cddpm_encoder_t;
int cddpm_encoder_create(int image_h, int image_w, int cond_dim, int fmt, cddpm_encoder_t** out);
void cddpm_encoder_destroy(cddpm_encoder_t* h);
int cddpm_encoder_param_count(const cddpm_encoder_t* h);
int cddpm_encoder_param_info(const cddpm_encoder_t* h, int index, const char** name, int64_t* numel);
int cddpm_encoder_set_param(cddpm_encoder_t* h, const char* name, const float* value, int64_t numel, void* stream);
/* x [B,1,H,W] fp32 -> c [B,cond_dim] fp32 */
int cddpm_encoder_forward(cddpm_encoder_t* h, const float* x, float* c, int B, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Condition encoder in TRAINING mode (DDPM_2D.training_step -> self(input), DDPM_2D.py:101-122, with the module in
 * train(): batch-statistics BatchNorm, running statistics updated, timm DropPath on the residual branches; and
 * loss.backward() through it).  Entries are the state_dict keys of the eval engine above (parameters and running
 * statistics, no num_batches_tracked), passed as a table of the caller's fp32 device tensors on every forward: the
 * optimizer owns the values, the engine re-packs its bf16 panels per step and updates the running statistics IN PLACE
 * (momentum 0.1, unbiased variance).  Convolutions: tcgen05 GEMMs (forward, data gradient, weight gradient), bf16
 * operands, fp32 accumulation; BatchNorm statistics / normalisation / summed gradients fp32 (sums in fp64).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct cddpm_encoder_train cddpm_encoder_train_t;
int cddpm_encoder_train_create(int image_h, int image_w, int cond_dim, cddpm_encoder_train_t** out);
void cddpm_encoder_train_destroy(cddpm_encoder_train_t* h);
int cddpm_encoder_train_entry_count(const cddpm_encoder_train_t* h);
int cddpm_encoder_train_entry_info(const cddpm_encoder_train_t* h, int index, const char** name, int64_t* numel,
                                   int* is_param);
/* Parameter gradients land in ONE flat fp32 buffer of grad_total floats; entry i at grad_offset (-1: a buffer). */
int64_t cddpm_encoder_train_grad_total(const cddpm_encoder_train_t* h);
int cddpm_encoder_train_grad_offset(const cddpm_encoder_train_t* h, int index, int64_t* offset);
int cddpm_encoder_train_num_blocks(const cddpm_encoder_train_t* h);       /* bottleneck blocks (16) */
int cddpm_encoder_train_launches(const cddpm_encoder_train_t* h, int backward);
/* x [B,1,H,W] fp32 -> out [B,cond_dim] fp32, B >= 2.  drop_scale: NULL or [num_blocks][B] per-sample scale of each
 * block's residual branch (timm DropPath: 0 or 1 / keep_prob). */
int cddpm_encoder_train_forward(cddpm_encoder_train_t* h, const float* const* values, int count, const float* x,
                                const float* drop_scale, float* out, int B, void* stream);
/* The weight-gradient GEMM of the training encoder on its own: dw[co][k] += sum_m dy[m][co] * x[m][k] for row-major
 * bf16 matrices dy [M][Cout], x [M][K] (the im2col'ed input of a convolution); dw fp32 [Cout][K] in the packed panel
 * order, accumulated with atomics (zero it first).  tcgen05, both operands MN-major. */
int cddpm_flat_wgrad(const void* dy, const void* x, int M, int Cout, int K, float* dw, void* stream);
/* Backward of the LAST forward: dout [B,cond_dim] -> grads (overwritten). */
int cddpm_encoder_train_backward(cddpm_encoder_train_t* h, const float* dout, float* grads, int B, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * gen_noise (src/utils/generate_noise.py:8-52): OpenSimplex-2D fractal field (octaves 6, persistence 0.8,
 * frequency 64 in the reference), bit-identical to the reference's float64 numba code.  perm_host is the 256-entry
 * permutation of generate_noise.py:214-232 in HOST memory.  out_f16 [B,1,H,W] (same field for every b) and/or
 * out_f32 [H,W]; either may be NULL.
 * ---------------------------------------------------------------------------------------------------------- */
int cddpm_simplex_noise(const uint8_t* perm_host, void* out_f16, float* out_f32, int B, int H, int W, int octaves,
                        double persistence, double frequency, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * DDPM arithmetic around the UNet: GaussianDiffusion (src/models/modules/cond_DDPM.py).
 * Images are [B,1,H,W] fp32 (HW = H*W contiguous per sample); noise is fp16 (gen_noise(...).half(),
 * generate_noise.py:12) when noise_f16 != 0, else fp32; schedule arrays are the module's fp32 [T] buffers.
 * ---------------------------------------------------------------------------------------------------------- */

/* q_sample (cond_DDPM.py:548-554), optionally fused with normalize_to_neg_one_to_one (:75, :653).
 * t is [B] int64, or a single shared index when t_shared != 0 (p_sample_loop's torch.tensor([T]), :452). */
int cddpm_q_sample(const float* img, const void* noise, int noise_f16, float* out, const float* sqrt_alphas_cumprod,
                   const float* sqrt_one_minus_alphas_cumprod, const int64_t* t, int t_shared, int B, int HW,
                   int normalize, void* stream);

/* p_sample (cond_DDPM.py:432-444): model_predictions (:400-420) + the clamp of p_mean_variance (:422-430, only when
 * clip_denoised != 0) + q_posterior (:391-398) + sigma_t * noise for t > 0.  noise may be NULL (t == 0).  pred_noise
 * selects the objective.  final_unnormalize applies unnormalize_to_zero_to_one (:78, :463) to the last step's result. */
int cddpm_posterior_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_prev,
                         const float* posterior_mean_coef1, const float* posterior_mean_coef2,
                         const float* posterior_log_variance_clipped, const float* sqrt_recip_alphas_cumprod,
                         const float* sqrt_recipm1_alphas_cumprod, int64_t t, int B, int HW, int pred_noise,
                         int clip_denoised, int final_unnormalize, void* stream);

/* One update of ddim_sample (cond_DDPM.py:487-511): pred_noise / x_start from model_predictions (:400-420; pred_noise
 * is derived from the UNclamped prediction), x_start.clamp_ when clip_denoised, then
 * x_next = x_start * sqrt_alpha_next + c * pred_noise + sigma * noise.  The five step scalars are fp32 values the host
 * computes exactly as the reference does from its schedule buffers (sqrt_recip(m1)_alphas_cumprod[time],
 * alphas_cumprod_prev[time / time_next], ddim_sampling_eta); noise may be NULL (time_next == 0). */
int cddpm_ddim_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_next,
                    float sqrt_recip_alphas_cumprod_t, float sqrt_recipm1_alphas_cumprod_t, float sqrt_alpha_next,
                    float c, float sigma, int B, int HW, int pred_noise, int clip_denoised, int final_unnormalize,
                    void* stream);

/* Tail of p_losses (cond_DDPM.py:636-645): reco = reco*reco_beta + reco_alpha * unnormalize(model_out) (pred_x0) or
 * unnormalize(x_t - sqrt(1-acp)*model_out) (pred_noise, as the reference computes it); loss[b] = mean |out-target|
 * (l2: squared) * p2_loss_weight[t_b].  reco / loss may be NULL.  The alpha/beta pair folds the test-time ensemble
 * mean (DDPM_2D.py:225-238) into the same pass. */
int cddpm_recon_finish(const float* model_out, const float* img, const float* x_t, const void* noise, int noise_f16,
                       float* reco, float reco_alpha, float reco_beta, float* loss,
                       const float* sqrt_one_minus_alphas_cumprod, const float* p2_loss_weight, const int64_t* t,
                       int t_shared, int B, int HW, int pred_noise, int l2, void* stream);

/* Backward of the loss of cddpm_recon_finish averaged over the batch (p_losses, cond_DDPM.py:636-645, as differentiated
 * by DDPM_2D.training_step): dout[b,i] = grad_loss[0] * p2_loss_weight[t_b] / (B*HW) * sign(model_out - target)
 * (l2: 2 (model_out - target)); target = 2 img - 1 (pred_x0) or the noise (pred_noise).  grad_loss: device scalar. */
int cddpm_loss_backward(const float* model_out, const float* img, const void* noise, int noise_f16,
                        const float* p2_loss_weight, const int64_t* t, const float* grad_loss, float* dout, int B,
                        int HW, int pred_noise, int l2, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Anomaly-scoring tail: _test_step of src/utils/utils_eval.py:18-194.
 * A volume is addressed logically as (y, x, d) = [H, W, D] like the reference's squeezed tensors.  Inputs carry
 * element strides (sy, sx, sd) so the dataloader layout [H,W,D] and the UNet output layout [D,1,H,W] are both read
 * in place; work buffers written here ("[D,H,W] buffers") are slice-major with x fastest.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct cddpm_vol_view {
  const float* ptr;
  int64_t sy, sx, sd;
} cddpm_vol_view;

/* diff = |orig - reco| (utils_eval.py:31), multiplied by the brain mask eroded per axial slice with a 3x3 cross,
 * `iterations` times, zero border (apply_brainmask_volume :447-460; the reference passes W // 25; iterations < 1
 * means scipy's "until stable").  erode == 0 skips the mask.  sums (7 doubles, may be NULL) receive
 * sum|d|, sum d^2 over all voxels / seg>0 / seg==0 and count(seg>0) for the l1/l2 errors (:36-49); run-to-run identical
 * (fixed-order fp64 per block, 2^-32 fixed-point integer atomics across blocks; a block partial >= 2^20 or not finite makes the sum read NaN). */
int cddpm_residual_erode(const cddpm_vol_view* orig, const cddpm_vol_view* reco, const cddpm_vol_view* seg,
                         const cddpm_vol_view* mask, int H, int W, int D, int iterations, int erode,
                         float* diff_masked_dhw, double* sums, void* stream);
/* Full-resolution evaluation (cfg.resizedEvaluation == False, utils_eval.py:24-25):
 * F.interpolate(final_volume, size=new_size, mode="trilinear", align_corners=True) of the logical [H,W,D] view into a
 * contiguous [Ho,Wo,Do] fp32 buffer (d fastest - the layout of the reference's squeezed tensor). */
int cddpm_trilinear_resize(const cddpm_vol_view* src, int H, int W, int D, float* dst_hwd, int Ho, int Wo, int Do,
                           void* stream);
/* One row of log_images' output grid (utils_eval.py:586-628, saveOutputImages): panels [4][H][W] fp32 on the device =
 * original, reconstruction, difference, segmentation of ONE axial slice; ranges [4][2] = (vmin, vmax) of each panel's
 * colour normalisation.  Each panel is drawn as torch.rot90(., 3) ('gray'; the difference with 'inferno') and the four
 * are laid side by side into rgb [W][4 H][3] uint8 (device).  Encoding / writing the PNG is host work done off the
 * critical path by the caller. */
int cddpm_compose_grid(const float* panels, const float* ranges, int H, int W, uint8_t* rgb, void* stream);
/* scipy.ndimage.median_filter(vol, (k,k,k)) with mode='reflect' (apply_3d_median_filter :462-464), k in {1,3,5}. */
int cddpm_median3d(const float* in_dhw, float* out_dhw, int H, int W, int D, int k, void* stream);
/* np.max of a buffer (val_range top of find_best_val, :86); out_max: one float on the device. */
int cddpm_max(const float* x, int64_t n, float* out_max, void* stream);
/* Dice counts for find_best_val (:508-545): counts[0] += #(seg>0); counts[1+2i] += #(x > q[i]);
 * counts[2+2i] += #(x > q[i] and seg>0), i < nq <= 4.  Accumulates (caller zeroes), so several volumes / ranks can
 * share the counters of the global threshold search (_test_end :262-271).  q_host is in HOST memory. */
int cddpm_threshold_counts(const float* x_dhw, const cddpm_vol_view* seg, int H, int W, int D, const float* q_host,
                           int nq, uint64_t* counts, void* stream);
/* out[i] = x[i] > thr (uint8), the "diffs_thresholded" volume (:95-98). */
int cddpm_threshold_mask(const float* x, int64_t n, float thr, uint8_t* out, void* stream);
/* Per image row y (the reference's per-"slice" loops run over axis 0, :138-144, :160-174):
 * rows[y][0..3] = #(x>thr), #(seg>0), #(both), #(mask>0); rowsum[y] = sum of x over mask>0. */
int cddpm_row_stats(const float* x_dhw, const cddpm_vol_view* seg, const cddpm_vol_view* mask, int H, int W, int D,
                    float thr, uint64_t* rows, double* rowsum, void* stream);
/* ROC-AUC and average precision with sklearn's tie semantics (compute_roc / compute_prc, :548-557):
 * result[0] = AUC, result[1] = AP (doubles on the device). */
int64_t cddpm_ranking_workspace_bytes(int64_t n);
int cddpm_ranking_metrics(const float* x_dhw, const cddpm_vol_view* seg, int H, int W, int D, void* workspace,
                          int64_t workspace_bytes, double* result, void* stream);

/* find_best_val (utils_eval.py:508-539, called per volume at :86-91 with val_range = (0, max) and max_steps = 10) as ONE
 * launch: the workspace must hold the result of cddpm_ranking_metrics for the same volume (scores sorted in descending
 * order + running positive counts), so a threshold's Dice counts are two look-ups and the bisection decisions are
 * taken on the device - no pass over the volume, no host round trip per step.  Range arithmetic in float32 with every
 * operation rounded separately, Dice in float64 from integer counts (the reference under NumPy >= 2).
 * result[0] = best Dice, result[1] = its threshold (a float32 value), result[2] = max(x) - 3 doubles on the device. */
int cddpm_dice_bisect(const void* ranking_workspace, int64_t n, int max_steps, double* result, void* stream);

/* filter_3d_connected_components (utils_eval.py:489-503; called at :100-102): skimage label(connectivity=3) and
 * regionprops.filled_area <= 7.  skimage fills holes with a full 3x3x3 element, so filled_area == area whenever
 * area <= 25: the filter drops every 26-connected component of at most max_size (reference: 7; <= 15) voxels.
 * mask_dhw / out_dhw: uint8 [D,H,W] buffers (0 / 1), not in place. */
int cddpm_filter_small_components(const uint8_t* mask_dhw, uint8_t* out_dhw, int H, int W, int D, int max_size,
                                  void* stream);
/* counts[0] += #(pred and seg>0), counts[1] += #(pred and not seg>0), counts[2] += #(not pred and seg>0): the
 * confusion_matrix / dice / tpr / fpr / precision / recall inputs of :105-129.  Accumulates (caller zeroes). */
int cddpm_confusion_counts(const uint8_t* pred_dhw, const cddpm_vol_view* seg, int H, int W, int D, uint64_t* counts,
                           void* stream);
/* monai.metrics.compute_hausdorff_distance(pred, seg, include_background=False, 'euclidean', percentile=None,
 * directed=False) (utils_eval.py:134) in integer-exact form.  result (4 x int64, device): [0] max over pred surface
 * voxels of the squared distance to the seg surface, [1] the reverse, [2] / [3] number of surface voxels of pred /
 * seg>0 (surface = mask minus its 6-neighbour erosion, zero border).  [0]/[1] are -1 when their own surface is
 * empty and >= 2^28 when the other one is.  Distance = sqrt(max([0],[1])) in float64 on the host; monai's corner
 * cases (nan when both are empty, inf when one is) follow from [2],[3]. */
int64_t cddpm_hausdorff_workspace_bytes(int H, int W, int D);
int cddpm_hausdorff(const uint8_t* pred_dhw, const cddpm_vol_view* seg, int H, int W, int D, void* workspace,
                    int64_t workspace_bytes, int64_t* result, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Volume preprocessing on the device: the once-per-volume transforms of the reference's datamodule,
 * src/datamodules/create_dataset.py:196-218 (get_transform: tio.CropOrPad -> tio.RescaleIntensity -> tio.Resample).
 * torchio / SimpleITK are third-party packages that are neither vendored in the reference nor installed here
 * (requirements: torchio==0.18.84, SimpleITK==2.2.0): CropOrPad / RescaleIntensity restate torchio's published source on
 * top of NumPy semantics that ARE pinned here (np.percentile / np.clip bit for bit); the B-spline path restates the
 * cubic B-spline decomposition + interpolation (pinned against scipy.ndimage, ITK's own truncated initialisation is
 * not reproducible here: parity unpinned).  Volumes are contiguous [H][W][D] fp32, one channel.
 * ---------------------------------------------------------------------------------------------------------- */
/* tio.CropOrPad((h, w, d), padding_mode=pad_value): centre crop / pad per axis; the odd voxel of a difference goes to
 * the start (torchio.CropOrPad._get_six_bounds_parameters). */
int cddpm_crop_or_pad(const float* in, int H, int W, int D, float* out, int h, int w, int d, float pad_value,
                      void* stream);
/* tio.RescaleIntensity((out_min, out_max), percentiles=(perc_low, perc_high), masking_method='mask') in place:
 * cut-offs = np.percentile(vol[mask > 0], (perc_low, perc_high)) (float64, 'linear'), np.clip, then
 * (x - min) / (max - min) * (out_max - out_min) + out_min in float32 with min / max of the CLIPPED array.  An empty mask
 * or a zero range leaves the volume as torchio leaves it.  cutoffs: optional 2 doubles on the device. */
int64_t cddpm_rescale_workspace_bytes(int64_t n);
int cddpm_rescale_intensity(float* vol, const float* mask, int64_t n, double perc_low, double perc_high, float out_min,
                            float out_max, void* workspace, int64_t workspace_bytes, double* cutoffs, void* stream);
/* tio.Resample(factor, image_interpolation='bspline') of a unit-spacing volume: output extent ceil(N / f) per axis
 * (cddpm_resample_size), sample i at continuous index 0.5 (f - 1) + f i, 0 outside the buffer.  bspline != 0: cubic
 * B-spline over float64 coefficients with mirror boundaries (workspace: cddpm_resample_workspace_bytes); bspline == 0:
 * nearest neighbour, halves rounded up (tio.LabelMap entries; no workspace needed). */
int cddpm_resample_size(int source, double factor);
int64_t cddpm_resample_workspace_bytes(int H, int W, int D);
int cddpm_resample(const float* in, int H, int W, int D, double fy, double fx, double fz, int bspline, float* out,
                   void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDDPM_B200_H_ */
