// extern "C" surface of libcddpm_b200 (declared in include/cddpm_b200.h).  Thin argument checking + dispatch.
#include "../../include/cddpm_b200.h"

#include "common.h"
#include "attention.cuh"
#include "backward.cuh"
#include "conv_bwd.cuh"
#include "conv_igemm.cuh"
#include "diffusion.cuh"
#include "elementwise.cuh"
#include "resnet_engine.cuh"
#include "resnet_train.cuh"
#include "simplex.cuh"
#include "preprocess.cuh"
#include "tail.cuh"
#include "unet_engine.cuh"

namespace cddpm {
const char* last_error_cstr();
}

using namespace cddpm;

extern "C" {

const char* cddpm_last_error(void) { return last_error_cstr(); }
const char* cddpm_version(void) { return "cddpm_b200 0.1 (sm_100a)"; }
int cddpm_memcpy_d2d(void* dst, const void* src, int64_t nbytes, void* stream) {
  if (!dst || !src || nbytes < 0) return fail(kInvalidArgument, "memcpy_d2d: bad arguments");
  return check_cuda(cudaMemcpyAsync(dst, src, static_cast<size_t>(nbytes), cudaMemcpyDeviceToDevice,
                                    static_cast<cudaStream_t>(stream)), "cudaMemcpyAsync");
}

int cddpm_pack_conv_weight(const float* w_oihw, int cout, int cin_total, int ksize, int cin_off, int c_s,
                           void* wpacked, int ktot, int koff, int fmt, void* stream) {
  if (!w_oihw || !wpacked) return fail(kInvalidArgument, "pack_conv_weight: null pointer");
  if (ksize != 1 && ksize != 3) return fail(kInvalidArgument, "pack_conv_weight: ksize must be 1 or 3");
  if (cin_off < 0 || cin_off + c_s > cin_total || koff < 0 || koff + ksize * ksize * c_s > ktot)
    return fail(kInvalidArgument, "pack_conv_weight: slice out of range");
  return launch_pack_conv_weight(w_oihw, cout, cin_total, ksize, cin_off, c_s, wpacked, ktot, koff, fmt,
                                 static_cast<cudaStream_t>(stream));
}

static int conv_igemm_impl(int num_src, const void* const* src, const int* src_c, const int* src_taps, int B, int H,
                           int W, int cout, const void* wpacked, const float* bias, const void* residual, void* out,
                           int out_f32, int fmt, double* gn_stats, void* stream) {
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
  d.gn_stats = gn_stats;
  if (gn_stats == nullptr) {
    // measurement aid (tools/bench_conv.py): CDDPM_CONV_BENCH_STATS=1 makes the epilogue also emit its GroupNorm
    // (sum, sumsq) partials, into a scratch table, the way every convolution planned by the UNet engine does
    static const bool bench_stats = [] {
      const char* e = getenv("CDDPM_CONV_BENCH_STATS");
      return e != nullptr && e[0] == '1';
    }();
    static double* scratch = nullptr;
    const size_t need = static_cast<size_t>(B) * (cout / 4 + 1) * 2;
    if (bench_stats && !out_f32 && need <= (1u << 20)) {
      if (scratch == nullptr) CDDPM_CUDA(cudaMalloc(&scratch, (1u << 20) * sizeof(double)));
      d.gn_stats = scratch;
    }
  }
  if (conv2_enabled() && conv2_supported(d)) {  // macro-tile kernel where the geometry allows it
    std::shared_ptr<void> holder;
    CDDPM_TRY(build_conv2(d, &holder));
    return launch_conv2(holder, static_cast<cudaStream_t>(stream));
  }
  if (gn_stats != nullptr)
    return fail(kUnsupported, "conv_igemm_stats: this geometry runs on the first-generation kernel, whose statistics "
                              "are per-box partial rows (planned by the engine), not the [B][cout/4][2] table");
  ConvIgemmParams p;
  CDDPM_TRY(build_conv_params(d, &p));
  return launch_conv_igemm(p, static_cast<cudaStream_t>(stream));
}

int cddpm_conv_igemm(int num_src, const void* const* src, const int* src_c, const int* src_taps, int B, int H,
                     int W, int cout, const void* wpacked, const float* bias, const void* residual, void* out,
                     int out_f32, int fmt, void* stream) {
  return conv_igemm_impl(num_src, src, src_c, src_taps, B, H, W, cout, wpacked, bias, residual, out, out_f32, fmt,
                         nullptr, stream);
}

int cddpm_conv_igemm_stats(int num_src, const void* const* src, const int* src_c, const int* src_taps, int B, int H,
                           int W, int cout, const void* wpacked, const float* bias, const void* residual, void* out,
                           int fmt, double* gn_stats, void* stream) {
  if (!gn_stats) return fail(kInvalidArgument, "conv_igemm_stats: gn_stats is required");
  return conv_igemm_impl(num_src, src, src_c, src_taps, B, H, W, cout, wpacked, bias, residual, out, 0, fmt, gn_stats,
                         stream);
}

int cddpm_pack_conv_weight_t(const float* w_oihw, int cout, int cin_total, int ksize, int cin_off, int c_s,
                             void* wpacked_t, int ktot, int koff, int fmt, void* stream) {
  if (!w_oihw || !wpacked_t) return fail(kInvalidArgument, "pack_conv_weight_t: null pointer");
  if (ksize != 1 && ksize != 3) return fail(kInvalidArgument, "pack_conv_weight_t: ksize must be 1 or 3");
  if (cin_off < 0 || cin_off + c_s > cin_total || koff < 0 || koff + ksize * ksize * cout > ktot)
    return fail(kInvalidArgument, "pack_conv_weight_t: slice out of range");
  return launch_pack_conv_weight_T(w_oihw, cout, cin_total, ksize, cin_off, c_s, wpacked_t, ktot, koff, fmt,
                                   static_cast<cudaStream_t>(stream));
}

int cddpm_conv_wgrad(int num_src, const void* const* src, const int* src_c, const int* src_taps, const int* skip,
                     const void* dy, int B, int H, int W, int cout, float* dw, int fmt, void* stream) {
  if (!src || !src_c || !src_taps || !dy || !dw) return fail(kInvalidArgument, "conv_wgrad: null pointer");
  if (num_src < 1 || num_src > kConvMaxSrc) return fail(kInvalidArgument, "conv_wgrad: num_src must be 1..3");
  WgradDesc d;
  d.num_src = num_src;
  for (int s = 0; s < num_src; ++s) {
    d.src[s] = src[s];
    d.src_c[s] = src_c[s];
    d.src_taps[s] = src_taps[s];
    d.src_skip[s] = skip ? skip[s] : 0;
  }
  d.dy = dy;
  d.B = B;
  d.H = H;
  d.W = W;
  d.Cout = cout;
  d.dw = dw;
  d.ab_format = fmt;
  std::shared_ptr<void> holder;
  CDDPM_TRY(build_wgrad(d, &holder));
  return launch_wgrad(holder, static_cast<cudaStream_t>(stream));
}

int cddpm_unpack_conv_grad(const float* dw_packed, int cout, int cin_total, int ksize, int cin_off, int c_s,
                           float* grad_oihw, int ktot, int koff, void* stream) {
  if (!dw_packed || !grad_oihw) return fail(kInvalidArgument, "unpack_conv_grad: null pointer");
  return launch_unpack_conv_grad(dw_packed, cout, cin_total, ksize, cin_off, c_s, grad_oihw, ktot, koff,
                                 static_cast<cudaStream_t>(stream));
}

int64_t cddpm_gn_workspace_floats(int B, int HW) {
  return static_cast<int64_t>(B) * gn_num_chunks(B, HW) * kGnGroups * 2;
}

int cddpm_groupnorm_film_silu(const void* x0, int c0, const void* x1, int c1, int B, int H, int W,
                              const float* gamma, const float* beta, const float* film, int film_stride,
                              int film_off, int silu, int mode, void* out, void* raw_out, float* workspace, int fmt,
                              void* stream) {
  if (!workspace) return fail(kInvalidArgument, "groupnorm: workspace is required");
  CatView v;
  v.p0 = x0;
  v.c0 = c0;
  v.p1 = x1;
  v.c1 = c1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CDDPM_TRY(launch_gn_stats(v, B, H * W, workspace, fmt, s));
  GnApplyArgs g;
  g.x = v;
  g.B = B;
  g.H = H;
  g.W = W;
  g.partial = workspace;
  g.gamma = gamma;
  g.beta = beta;
  g.film = film;
  g.film_stride = film_stride;
  g.film_off = film_off;
  g.silu = silu;
  g.mode = mode;
  g.out = out;
  g.raw_out = raw_out;
  g.fmt = fmt;
  return launch_gn_apply(g, s);
}

int cddpm_linear(const float* in, int in_stride, const float* w, const float* bias, float* out, int out_stride,
                 int B, int I, int O, int silu_in, int silu_out, void* stream) {
  return launch_linear_ex(in, in_stride, w, bias, out, out_stride, B, I, O, silu_in, silu_out,
                          static_cast<cudaStream_t>(stream));
}

int cddpm_timestep_embedding(const int64_t* t, float* emb, int B, int dim, void* stream) {
  if (!t || !emb) return fail(kInvalidArgument, "timestep_embedding: null pointer");
  return launch_timestep_embedding(t, emb, B, dim, static_cast<cudaStream_t>(stream));
}

int cddpm_attention(const void* qkv, void* out, int B, int L, int C, int fmt, void* stream) {
  return launch_attention(qkv, out, B, L, C, fmt, static_cast<cudaStream_t>(stream));
}

int cddpm_q_sample(const float* img, const void* noise, int noise_f16, float* out, const float* sqrt_alphas_cumprod,
                   const float* sqrt_one_minus_alphas_cumprod, const int64_t* t, int t_shared, int B, int HW,
                   int normalize, void* stream) {
  return launch_q_sample(img, noise, noise_f16, out, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, t, t_shared,
                         B, HW, normalize, static_cast<cudaStream_t>(stream));
}

int cddpm_posterior_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_prev,
                         const float* posterior_mean_coef1, const float* posterior_mean_coef2,
                         const float* posterior_log_variance_clipped, const float* sqrt_recip_alphas_cumprod,
                         const float* sqrt_recipm1_alphas_cumprod, int64_t t, int B, int HW, int pred_noise,
                         int clip_denoised, int final_unnormalize, void* stream) {
  return launch_posterior_step(model_out, x_t, noise, noise_f16, x_prev, posterior_mean_coef1, posterior_mean_coef2,
                               posterior_log_variance_clipped, sqrt_recip_alphas_cumprod,
                               sqrt_recipm1_alphas_cumprod, t, B, HW, pred_noise, clip_denoised, final_unnormalize,
                               static_cast<cudaStream_t>(stream));
}

int cddpm_ddim_step(const float* model_out, const float* x_t, const void* noise, int noise_f16, float* x_next,
                    float sqrt_recip_alphas_cumprod_t, float sqrt_recipm1_alphas_cumprod_t, float sqrt_alpha_next,
                    float c, float sigma, int B, int HW, int pred_noise, int clip_denoised, int final_unnormalize,
                    void* stream) {
  return launch_ddim_step(model_out, x_t, noise, noise_f16, x_next, sqrt_recip_alphas_cumprod_t,
                          sqrt_recipm1_alphas_cumprod_t, sqrt_alpha_next, c, sigma, B, HW, pred_noise, clip_denoised,
                          final_unnormalize, static_cast<cudaStream_t>(stream));
}

int cddpm_loss_backward(const float* model_out, const float* img, const void* noise, int noise_f16,
                        const float* p2_loss_weight, const int64_t* t, const float* grad_loss, float* dout, int B,
                        int HW, int pred_noise, int l2, void* stream) {
  return launch_loss_backward(model_out, img, noise, noise_f16, p2_loss_weight, t, grad_loss, dout, B, HW, pred_noise,
                              l2, static_cast<cudaStream_t>(stream));
}

int cddpm_recon_finish(const float* model_out, const float* img, const float* x_t, const void* noise, int noise_f16,
                       float* reco, float reco_alpha, float reco_beta, float* loss,
                       const float* sqrt_one_minus_alphas_cumprod, const float* p2_loss_weight, const int64_t* t,
                       int t_shared, int B, int HW, int pred_noise, int l2, void* stream) {
  return launch_recon_finish(model_out, img, x_t, noise, noise_f16, reco, reco_alpha, reco_beta, loss,
                             sqrt_one_minus_alphas_cumprod, p2_loss_weight, t, t_shared, B, HW, pred_noise, l2,
                             static_cast<cudaStream_t>(stream));
}

int cddpm_simplex_noise(const uint8_t* perm_host, void* out_f16, float* out_f32, int B, int H, int W, int octaves,
                        double persistence, double frequency, void* stream) {
  return launch_simplex_noise(perm_host, out_f16, out_f32, B, H, W, octaves, persistence, frequency,
                              static_cast<cudaStream_t>(stream));
}

struct cddpm_encoder {
  ResNetEngine engine;
};
int cddpm_encoder_create(int image_h, int image_w, int cond_dim, int fmt, cddpm_encoder_t** out) {
  if (!out) return fail(kInvalidArgument, "encoder_create: null pointer");
  cddpm_encoder* h = new cddpm_encoder();
  int st = h->engine.init(image_h, image_w, cond_dim, fmt);
  if (st != kOk) {
    delete h;
    return st;
  }
  *out = h;
  return kOk;
}
void cddpm_encoder_destroy(cddpm_encoder_t* h) { delete h; }
int cddpm_encoder_param_count(const cddpm_encoder_t* h) { return h ? h->engine.param_count() : 0; }
int cddpm_encoder_param_info(const cddpm_encoder_t* h, int index, const char** name, int64_t* numel) {
  if (!h || !name || !numel) return fail(kInvalidArgument, "encoder_param_info: null pointer");
  return h->engine.param_info(index, name, numel);
}
int cddpm_encoder_set_param(cddpm_encoder_t* h, const char* name, const float* value, int64_t numel, void* stream) {
  if (!h || !name) return fail(kInvalidArgument, "encoder_set_param: null pointer");
  return h->engine.set_param(name, value, numel, static_cast<cudaStream_t>(stream));
}
int cddpm_encoder_forward(cddpm_encoder_t* h, const float* x, float* c, int B, void* stream) {
  if (!h) return fail(kInvalidArgument, "encoder_forward: null handle");
  return h->engine.forward(x, c, B, static_cast<cudaStream_t>(stream));
}

struct cddpm_encoder_train {
  ResNetTrainEngine engine;
};
int cddpm_encoder_train_create(int image_h, int image_w, int cond_dim, cddpm_encoder_train_t** out) {
  if (!out) return fail(kInvalidArgument, "encoder_train_create: null pointer");
  cddpm_encoder_train* h = new cddpm_encoder_train();
  int st = h->engine.init(image_h, image_w, cond_dim);
  if (st != kOk) {
    delete h;
    return st;
  }
  *out = h;
  return kOk;
}
void cddpm_encoder_train_destroy(cddpm_encoder_train_t* h) { delete h; }
int cddpm_encoder_train_entry_count(const cddpm_encoder_train_t* h) { return h ? h->engine.entry_count() : 0; }
int cddpm_encoder_train_entry_info(const cddpm_encoder_train_t* h, int index, const char** name, int64_t* numel,
                                   int* is_param) {
  if (!h || !name || !numel || !is_param) return fail(kInvalidArgument, "encoder_train_entry_info: null pointer");
  return h->engine.entry_info(index, name, numel, is_param);
}
int64_t cddpm_encoder_train_grad_total(const cddpm_encoder_train_t* h) { return h ? h->engine.grad_total() : 0; }
int cddpm_encoder_train_grad_offset(const cddpm_encoder_train_t* h, int index, int64_t* offset) {
  if (!h || !offset) return fail(kInvalidArgument, "encoder_train_grad_offset: null pointer");
  return h->engine.grad_offset(index, offset);
}
int cddpm_encoder_train_num_blocks(const cddpm_encoder_train_t* h) { return h ? h->engine.num_blocks() : 0; }
int cddpm_encoder_train_launches(const cddpm_encoder_train_t* h, int backward) {
  if (!h) return 0;
  return backward ? h->engine.launches_backward() : h->engine.launches_forward();
}
int cddpm_encoder_train_forward(cddpm_encoder_train_t* h, const float* const* values, int count, const float* x,
                                const float* drop_scale, float* out, int B, void* stream) {
  if (!h) return fail(kInvalidArgument, "encoder_train_forward: null handle");
  return h->engine.forward(values, count, x, drop_scale, out, B, static_cast<cudaStream_t>(stream));
}
int cddpm_flat_wgrad(const void* dy, const void* x, int M, int Cout, int K, float* dw, void* stream) {
  return launch_flat_wgrad_bf16(dy, x, M, Cout, K, dw, static_cast<cudaStream_t>(stream));
}
int cddpm_encoder_train_backward(cddpm_encoder_train_t* h, const float* dout, float* grads, int B, void* stream) {
  if (!h) return fail(kInvalidArgument, "encoder_train_backward: null handle");
  return h->engine.backward(dout, grads, B, static_cast<cudaStream_t>(stream));
}

static VolView to_view(const cddpm_vol_view* v) {
  VolView o;
  if (v) {
    o.p = v->ptr;
    o.sy = v->sy;
    o.sx = v->sx;
    o.sd = v->sd;
  }
  return o;
}

int cddpm_residual_erode(const cddpm_vol_view* orig, const cddpm_vol_view* reco, const cddpm_vol_view* seg,
                         const cddpm_vol_view* mask, int H, int W, int D, int iterations, int erode,
                         float* diff_masked_dhw, double* sums, void* stream) {
  if (!orig || !reco) return fail(kInvalidArgument, "residual_erode: null view");
  return launch_residual_erode(to_view(orig), to_view(reco), to_view(seg), to_view(mask), H, W, D, iterations, erode,
                               diff_masked_dhw, sums, static_cast<cudaStream_t>(stream));
}
int cddpm_trilinear_resize(const cddpm_vol_view* src, int H, int W, int D, float* dst_hwd, int Ho, int Wo, int Do,
                           void* stream) {
  if (!src) return fail(kInvalidArgument, "trilinear_resize: null view");
  return launch_trilinear_resize(to_view(src), H, W, D, dst_hwd, Ho, Wo, Do, static_cast<cudaStream_t>(stream));
}
int cddpm_compose_grid(const float* panels, const float* ranges, int H, int W, uint8_t* rgb, void* stream) {
  return launch_compose_grid(panels, ranges, H, W, rgb, static_cast<cudaStream_t>(stream));
}
int cddpm_median3d(const float* in_dhw, float* out_dhw, int H, int W, int D, int k, void* stream) {
  return launch_median3d(in_dhw, out_dhw, H, W, D, k, static_cast<cudaStream_t>(stream));
}
int cddpm_max(const float* x, int64_t n, float* out_max, void* stream) {
  return launch_max(x, n, out_max, static_cast<cudaStream_t>(stream));
}
int cddpm_threshold_counts(const float* x_dhw, const cddpm_vol_view* seg, int H, int W, int D, const float* q_host,
                           int nq, uint64_t* counts, void* stream) {
  if (!seg) return fail(kInvalidArgument, "threshold_counts: null view");
  return launch_threshold_counts(x_dhw, to_view(seg), H, W, D, q_host, nq,
                                 reinterpret_cast<unsigned long long*>(counts), static_cast<cudaStream_t>(stream));
}
int cddpm_threshold_mask(const float* x, int64_t n, float thr, uint8_t* out, void* stream) {
  return launch_threshold_mask(x, n, thr, out, static_cast<cudaStream_t>(stream));
}
int cddpm_row_stats(const float* x_dhw, const cddpm_vol_view* seg, const cddpm_vol_view* mask, int H, int W, int D,
                    float thr, uint64_t* rows, double* rowsum, void* stream) {
  if (!mask) return fail(kInvalidArgument, "row_stats: null view");
  return launch_row_stats(x_dhw, to_view(seg), to_view(mask), H, W, D, thr,
                          reinterpret_cast<unsigned long long*>(rows), rowsum, static_cast<cudaStream_t>(stream));
}
int64_t cddpm_ranking_workspace_bytes(int64_t n) { return static_cast<int64_t>(ranking_workspace_bytes(n)); }
int cddpm_ranking_metrics(const float* x_dhw, const cddpm_vol_view* seg, int H, int W, int D, void* workspace,
                          int64_t workspace_bytes, double* result, void* stream) {
  if (!seg) return fail(kInvalidArgument, "ranking_metrics: null view");
  return launch_ranking_metrics(x_dhw, to_view(seg), H, W, D, workspace, static_cast<size_t>(workspace_bytes), result,
                                static_cast<cudaStream_t>(stream));
}
int cddpm_dice_bisect(const void* ranking_workspace, int64_t n, int max_steps, double* result, void* stream) {
  return launch_dice_bisect(ranking_workspace, n, max_steps, result, static_cast<cudaStream_t>(stream));
}
int cddpm_filter_small_components(const uint8_t* mask_dhw, uint8_t* out_dhw, int H, int W, int D, int max_size,
                                  void* stream) {
  return launch_filter_small_components(mask_dhw, out_dhw, H, W, D, max_size, static_cast<cudaStream_t>(stream));
}
int cddpm_confusion_counts(const uint8_t* pred_dhw, const cddpm_vol_view* seg, int H, int W, int D, uint64_t* counts,
                           void* stream) {
  if (!seg) return fail(kInvalidArgument, "confusion_counts: null view");
  return launch_confusion_counts(pred_dhw, to_view(seg), H, W, D, reinterpret_cast<unsigned long long*>(counts),
                                 static_cast<cudaStream_t>(stream));
}
int64_t cddpm_hausdorff_workspace_bytes(int H, int W, int D) {
  return static_cast<int64_t>(hausdorff_workspace_bytes(H, W, D));
}
int cddpm_hausdorff(const uint8_t* pred_dhw, const cddpm_vol_view* seg, int H, int W, int D, void* workspace,
                    int64_t workspace_bytes, int64_t* result, void* stream) {
  if (!seg) return fail(kInvalidArgument, "hausdorff: null view");
  return launch_hausdorff(pred_dhw, to_view(seg), H, W, D, workspace, static_cast<size_t>(workspace_bytes),
                          reinterpret_cast<long long*>(result), static_cast<cudaStream_t>(stream));
}

struct cddpm_unet {
  UNetEngine engine;
};

int cddpm_unet_create(const cddpm_unet_config* cfg, cddpm_unet_t** out) {
  if (!cfg || !out) return fail(kInvalidArgument, "unet_create: null pointer");
  cddpm_unet* h = new cddpm_unet();
  int st = h->engine.init(*cfg);
  if (st != kOk) {
    delete h;
    return st;
  }
  *out = h;
  return kOk;
}
void cddpm_unet_destroy(cddpm_unet_t* h) { delete h; }
int cddpm_unet_param_count(const cddpm_unet_t* h) { return h ? h->engine.param_count() : 0; }
int cddpm_unet_param_info(const cddpm_unet_t* h, int index, const char** name, int64_t* numel) {
  if (!h || !name || !numel) return fail(kInvalidArgument, "unet_param_info: null pointer");
  return h->engine.param_info(index, name, numel);
}
int cddpm_unet_set_param(cddpm_unet_t* h, const char* name, const float* value, int64_t numel, void* stream) {
  if (!h || !name) return fail(kInvalidArgument, "unet_set_param: null pointer");
  return h->engine.set_param(name, value, numel, static_cast<cudaStream_t>(stream));
}
int cddpm_unet_forward(cddpm_unet_t* h, const float* x, const int64_t* t, const float* cond, float* out, int B,
                       void* stream) {
  if (!h) return fail(kInvalidArgument, "unet_forward: null handle");
  return h->engine.forward(x, t, cond, out, B, static_cast<cudaStream_t>(stream));
}
int cddpm_unet_tap(const cddpm_unet_t* h, const char* layer, void** ptr, int* C, int* H, int* W) {
  if (!h || !layer || !ptr || !C || !H || !W) return fail(kInvalidArgument, "unet_tap: null pointer");
  return h->engine.tap(layer, ptr, C, H, W);
}
int cddpm_unet_film(const cddpm_unet_t* h, const float** ptr, int* stride) {
  if (!h || !ptr || !stride) return fail(kInvalidArgument, "unet_film: null pointer");
  return h->engine.film(ptr, stride);
}
int64_t cddpm_unet_conv_flops(const cddpm_unet_t* h) { return h ? h->engine.conv_flops_per_sample() : 0; }
int cddpm_unet_launches(const cddpm_unet_t* h) { return h ? h->engine.launches_per_forward() : 0; }
int cddpm_unet_profile_arm(cddpm_unet_t* h) {
  if (!h) return fail(kInvalidArgument, "unet_profile_arm: null handle");
  return h->engine.profile_arm();
}
int cddpm_unet_profile_read(cddpm_unet_t* h, double* conv_ms, int* conv_launches) {
  if (!h) return fail(kInvalidArgument, "unet_profile_read: null handle");
  return h->engine.profile_read(conv_ms, conv_launches);
}

int cddpm_unet_set_params(cddpm_unet_t* h, const float* const* values, int count, void* stream) {
  if (!h || !values) return fail(kInvalidArgument, "unet_set_params: null pointer");
  if (count != h->engine.param_count()) return fail(kInvalidArgument, "unet_set_params: wrong parameter count");
  bool all = true;
  for (int i = 0; i < count; ++i) all = all && values[i] != nullptr;
  if (all) return h->engine.set_params_all(values, count, static_cast<cudaStream_t>(stream));
  for (int i = 0; i < count; ++i) {
    if (values[i] == nullptr) continue;
    const char* name = nullptr;
    int64_t numel = 0;
    CDDPM_TRY(h->engine.param_info(i, &name, &numel));
    CDDPM_TRY(h->engine.set_param(name, values[i], numel, static_cast<cudaStream_t>(stream)));
  }
  return kOk;
}
int64_t cddpm_unet_grad_total(const cddpm_unet_t* h) { return h ? h->engine.grad_total() : 0; }
int cddpm_unet_grad_offset(const cddpm_unet_t* h, int index, int64_t* offset) {
  if (!h || !offset) return fail(kInvalidArgument, "unet_grad_offset: null pointer");
  return h->engine.grad_offset(index, offset);
}
int cddpm_unet_set_training(cddpm_unet_t* h, int training) {
  if (!h) return fail(kInvalidArgument, "unet_set_training: null handle");
  return h->engine.set_training(training != 0);
}
int cddpm_unet_backward(cddpm_unet_t* h, const float* dout, float* grads, float* dcond, int B, void* stream) {
  if (!h) return fail(kInvalidArgument, "unet_backward: null handle");
  return h->engine.backward(dout, grads, dcond, B, static_cast<cudaStream_t>(stream));
}
int64_t cddpm_unet_bwd_flops(const cddpm_unet_t* h) { return h ? h->engine.bwd_flops_per_sample() : 0; }
int cddpm_unet_bwd_launches(const cddpm_unet_t* h) { return h ? h->engine.bwd_launches() : 0; }
int cddpm_adam_step(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel,
                    const int* block_tensor, const int64_t* block_off, int total_blocks, float lr, float beta1,
                    float beta2, float eps, float bc1, float bc2, void* stream) {
  return launch_adam_step(p, g, m, v, numel, block_tensor, block_off, total_blocks, lr, beta1, beta2, eps, bc1, bc2,
                          static_cast<cudaStream_t>(stream));
}
int64_t cddpm_attention_bwd_scratch_bytes(int B, int L, int C) { return attention_bwd_scratch_elems(B, L, C) * 2; }
int cddpm_attention_bwd(const void* qkv, const void* dout, void* dqkv, void* scratch, int B, int L, int C, int fmt,
                        void* stream) {
  return launch_attention_bwd(qkv, dout, dqkv, scratch, B, L, C, fmt, static_cast<cudaStream_t>(stream));
}

int cddpm_crop_or_pad(const float* in, int H, int W, int D, float* out, int h, int w, int d, float pad_value,
                      void* stream) {
  return launch_crop_or_pad(in, H, W, D, out, h, w, d, pad_value, static_cast<cudaStream_t>(stream));
}
int64_t cddpm_rescale_workspace_bytes(int64_t n) {
  if (n < 1 || n > (1ll << 30)) return 0;
  return static_cast<int64_t>(rescale_workspace_bytes(n));
}
int cddpm_rescale_intensity(float* vol, const float* mask, int64_t n, double perc_low, double perc_high, float out_min,
                            float out_max, void* workspace, int64_t workspace_bytes, double* cutoffs, void* stream) {
  return launch_rescale_intensity(vol, mask, n, perc_low, perc_high, out_min, out_max, workspace,
                                  static_cast<size_t>(workspace_bytes), cutoffs, static_cast<cudaStream_t>(stream));
}
int cddpm_resample_size(int source, double factor) {
  int t = 0;
  resample_size(source, factor, &t);
  return t;
}
int64_t cddpm_resample_workspace_bytes(int H, int W, int D) {
  return static_cast<int64_t>(resample_workspace_bytes(H, W, D));
}
int cddpm_resample(const float* in, int H, int W, int D, double fy, double fx, double fz, int bspline, float* out,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  return launch_resample(in, H, W, D, fy, fx, fz, bspline, out, workspace, static_cast<size_t>(workspace_bytes),
                         static_cast<cudaStream_t>(stream));
}

}  // extern "C"
