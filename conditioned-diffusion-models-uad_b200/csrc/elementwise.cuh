// Bandwidth-bound kernels of the UNet forward: GroupNorm statistics / apply (+FiLM +SiLU +resample), the
// embedding MLPs, and the 1-channel stem / head convolutions.  All activations are NHWC 16-bit.
#pragma once
#include "common.h"

namespace cddpm {

constexpr int kGnGroups = 32;
constexpr int kGnMaxChunks = 64;

enum ResampleMode : int { kResampleNone = 0, kResampleUp2 = 1, kResampleDown2 = 2 };

// Two-source (channel-concat aware) NHWC tensor view: channels [0,c0) come from p0, [c0,c0+c1) from p1.
struct CatView {
  const void* p0 = nullptr;
  const void* p1 = nullptr;
  int c0 = 0;
  int c1 = 0;
  int C() const { return c0 + c1; }
};

// Number of statistics chunks per image launch_gn_stats() uses for (B, HW); partial needs B*chunks*32*2 floats.
int gn_num_chunks(int B, int HW);

// partial[b][chunk][32][2] = (sum, sum of squares) over the chunk's pixels and the group's channels.
// GroupNorm32 statistics: src/models/LDM/modules/diffusionmodules/util.py:214-216 (fp32, eps 1e-5).
int launch_gn_stats(const CatView& x, int B, int HW, float* partial, int fmt, cudaStream_t stream);
// stats[B][C/4][2] += (sum, sumsq) per image and 4-channel bucket (double atomics; zero it first).
int launch_gn_stats4(const void* x, int C, int B, int HW, double* stats, int fmt, cudaStream_t stream);

struct GnApplyArgs {
  CatView x;
  int B = 0, H = 0, W = 0;    // input spatial size
  const float* partial = nullptr;  // from launch_gn_stats over the same view, OR
  const double* stats0 = nullptr;  // [B][c0/4][2] (sum, sumsq) per 4-channel bucket of the first source (conv epilogue /
  const double* stats1 = nullptr;  // launch_gn_stats4), and of the second source when the view is a concat
  const float* gamma = nullptr;    // [C]
  const float* beta = nullptr;     // [C]
  const float* film = nullptr;     // optional [B][film_stride]: scale at [film_off + c], shift at [film_off + C + c]
  int film_stride = 0;
  int film_off = 0;
  int silu = 1;
  int mode = kResampleNone;
  void* out = nullptr;      // [B,H',W',C] normalised (+FiLM, +SiLU), resampled
  void* raw_out = nullptr;  // optional [B,H',W',C]: the un-normalised input resampled the same way (ResBlock x_upd)
  int fmt = 1;
};
// y = act(GN(x) * (1 + scale) + shift), then nearest-up x2 / avg-pool 2 (OpenAI_Unet.py:287-296, :325-331).
int launch_gn_apply(const GnApplyArgs& a, cudaStream_t stream);

// out[b][o] = bias[o] + sum_i act(in[b][i]) * W[o][i]   (fp32; act: 0 none, 1 SiLU).  nn.Linear call sites:
// time_embed / label_emb (OpenAI_Unet.py:583-602) and every ResBlock.emb_layers (:245-251).
int launch_linear(const float* in, int in_stride, const float* W, const float* bias, float* out, int out_stride,
                  int B, int I, int O, int act_in, cudaStream_t stream);
// Same with an optional SiLU on the output as well (lets the embedding MLP keep only activated values).
int launch_linear_ex(const float* in, int in_stride, const float* W, const float* bias, float* out, int out_stride,
                     int B, int I, int O, int act_in, int act_out, cudaStream_t stream);
// ... and an optional 16-bit copy of the result (the tensor-core operand of the FiLM projection); `out` may be NULL.
int launch_linear_16(const float* in, int in_stride, const float* W, const float* bias, float* out, int out_stride,
                     int B, int I, int O, int act_in, int act_out, void* out16, int out16_stride, int fmt,
                     cudaStream_t stream);

// emb[b][0:half] = cos(t_b * f_i), emb[b][half:] = sin(t_b * f_i), f_i = exp(-ln(1e4) * i / half) (util.py:151-171).
int launch_timestep_embedding(const int64_t* t, float* emb, int B, int dim, cudaStream_t stream);
// ... with an optional 16-bit copy (tensor-core operand); emb may be NULL.
int launch_timestep_embedding16(const int64_t* t, float* emb, void* emb16, int fmt, int B, int dim, cudaStream_t stream);
int launch_to16(const float* in, void* out, int n, int fmt, cudaStream_t stream);
// out16[b][col_off + o] = act(sum_k in16[b][k] w16[o][k] + bias[o]) for small batches (embedding MLPs); fp32 accumulation
int launch_linear16(const void* in, const void* w16, const float* bias, void* out, int out_stride, int col_off, int B,
                    int I, int O, int fmt, int silu, cudaStream_t stream);

// Stem: x fp32 [B,1,H,W] -> NHWC 16-bit [B,H,W,Cout], 3x3 pad 1 (OpenAI_Unet.py:609).
int launch_conv_in(const float* x, const float* w, const float* bias, void* out, double* stats, int B, int H, int W,
                   int Cout, int fmt, cudaStream_t stream);
// Head: NHWC 16-bit [B,H,W,C] -> fp32 [B,1,H,W], 3x3 pad 1, one output channel (OpenAI_Unet.py:796).
int launch_conv_out(const void* x, const float* w, const float* bias, float* out, int B, int H, int W, int C,
                    int fmt, cudaStream_t stream);

// Head with its GroupNorm + SiLU folded in: x is the RAW input of out.0, stats its [B][C/4][2] (sum, sumsq) buckets.
int launch_conv_out_gn(const void* x, const double* stats, const float* gamma, const float* beta, const float* w,
                       const float* bias, float* out, int B, int H, int W, int C, int fmt, cudaStream_t stream);

int launch_vec_add(const float* a, const float* b, float* out, int n, cudaStream_t stream);

}  // namespace cddpm
