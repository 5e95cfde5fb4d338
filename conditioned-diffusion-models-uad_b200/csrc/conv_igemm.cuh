// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Computes, for stride-1 "same" convolutions over NHWC 16-bit activations,
//     out[n,y,x,co] = bias[co] + residual[n,y,x,co] + sum_s sum_tap sum_ci A_s[n, y+dy(tap), x+dx(tap), ci] * Wp[co, k(s,tap,ci)]
// where up to three A sources (e.g. the two halves of a channel concat, or a 3x3 branch plus a fused 1x1 skip
// branch) share one accumulator.  GEMM view: M = B*H*W pixels, N = Cout, K = sum_s taps_s * C_s.
// This is the kernel behind nn.Conv2d 3x3/1x1 in the reference UNet (OpenAI_Unet.py:231,257,268,367,375).
#pragma once
#include <memory>

#include "common.h"

namespace cddpm {

constexpr int kConvMaxSrc = 3;
constexpr int kConvTileM = 128;   // UMMA M (one TMEM lane per output pixel)
constexpr int kConvBlockK = 64;   // 64 x 16-bit = one 128-byte swizzle row per pixel per K step
constexpr int kConvThreads = 256; // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-7 epilogue

struct ConvIgemmParams {
  CUtensorMap tmap_a[kConvMaxSrc];  // 4-D maps {C, W, H, B} over each NHWC source, box {64, box_w, box_h, 1}
  CUtensorMap tmap_b;               // 2-D map {Ktot, Cout} over the packed weights, box {64, n_tile}
  CUtensorMap tmap_b_half;          // same matrix, box {64, n_tile / 2}: each CTA of a pair stages half of the rows
  int num_src;
  int src_c[kConvMaxSrc];     // channels of each source (multiple of 64)
  int src_taps[kConvMaxSrc];  // 9 (3x3, pad 1) or 1 (1x1)
  int B, H, W, Cout;
  int n_tile;          // UMMA N: multiple of 32, <= 256, divides Cout
  int box_w, box_h;    // spatial box; box_w*box_h is 128 or 64
  int boxes_per_tile;  // 128 / (box_w*box_h)
  int tiles_w, tiles_h;
  int total_boxes;  // B * tiles_w * tiles_h
  int num_m_tiles, num_n_tiles;
  int num_stages;
  int ab_format;  // 0 = f16, 1 = bf16
  int tmem_cols;  // power of two >= 2*n_tile
  int out_is_f32;
  const float* bias;     // [Cout] or nullptr
  const bf16* residual;  // [B,H,W,Cout] or nullptr (same 16-bit type as the activations)
  void* out;             // [B,H,W,Cout] 16-bit, or fp32 when out_is_f32
  double* gn_stats;      // optional [B][Cout/4][2]: (sum, sum of squares) of the produced tensor per image and
                         // 4-channel bucket, accumulated with atomics (the caller zeroes it) — feeds the next GroupNorm
  int flat;              // plain GEMM: A is a row-major [M][C] matrix (2-D tensor map), one source, taps == 1
  int M;                 // rows of the flat problem
  int relu;              // activation after bias and residual: 0 none, 1 ReLU, 2 SiLU
  int out_stride;        // elements between consecutive output rows (pixels); Cout unless writing into a wider matrix
  int out_col_off;       // first output column within the row
  int pair;              // launched as 2-CTA clusters issuing tcgen05.mma.cta_group::2 (M = 256)
};
// true unless the environment sets CDDPM_CONV_PAIR=0 (A/B switch for measurements)
bool pair_enabled();

// Describes one convolution launch in host terms; build_conv_params() turns it into ConvIgemmParams.
struct ConvDesc {
  int num_src = 0;
  const void* src[kConvMaxSrc] = {nullptr, nullptr, nullptr};
  int src_c[kConvMaxSrc] = {0, 0, 0};
  int src_taps[kConvMaxSrc] = {0, 0, 0};
  int B = 0, H = 0, W = 0, Cout = 0;
  const void* wpacked = nullptr;  // [Cout][Ktot] 16-bit
  const float* bias = nullptr;
  const void* residual = nullptr;
  void* out = nullptr;
  int out_is_f32 = 0;
  int ab_format = 1;
  double* gn_stats = nullptr;
  int flat_rows = 0;  // > 0: plain GEMM over src[0] = [flat_rows][src_c[0]] (B/H/W ignored), e.g. nn.Linear or an
                      // im2col'ed convolution
  int relu = 0;         // 0 none, 1 ReLU, 2 SiLU
  int out_stride = 0;   // 0 = Cout
  int out_col_off = 0;
  int identity_k = 0;   // K columns that only add a tensor through an identity weight block (not counted as FLOPs)
  // Inference only, conv_igemm2 only: finish the GroupNorm (+FiLM) + SiLU that FOLLOWS this convolution inside its
  // epilogue.  `out` then receives act(GN(conv + bias) * (1 + scale) + shift); the raw result is never written.
  // Needs gn_stats (the statistics rendezvous) and gn_counters ([B][Cout/128] 64-bit, zeroed with the statistics).
  const float* gn_gamma = nullptr;
  const float* gn_beta = nullptr;
  const float* gn_film = nullptr;  // optional [B][gn_film_stride]: scale at [gn_film_off + c], shift at [+ Cout + c]
  int gn_film_stride = 0;
  int gn_film_off = 0;
  unsigned long long* gn_counters = nullptr;
};

// Second-generation kernel (conv_igemm2.cu): 16x16-pixel macro tiles x 128 channels; one staged 18x18 halo tile per
// 64-channel chunk serves all nine taps through row-shifted UMMA descriptors.  Applies when W % 16 == 0, H % 16 == 0,
// Cout % 128 == 0 and the output is a plain 16-bit NHWC tensor.
bool conv2_supported(const ConvDesc& d);
int build_conv2(const ConvDesc& d, std::shared_ptr<void>* holder);
int launch_conv2(const std::shared_ptr<void>& holder, cudaStream_t stream);
// true unless the environment sets CDDPM_CONV_V2=0 (A/B switch for measurements)
bool conv2_enabled();

int conv_ktot(const ConvDesc& d);
int build_conv_params(const ConvDesc& d, ConvIgemmParams* p);
int launch_conv_igemm(const ConvIgemmParams& p, cudaStream_t stream, int max_ctas = 0);
// Number of per-box statistic rows build_conv_params() will produce for this geometry.
int conv_num_boxes(int B, int H, int W);

// Re-layout an OIHW fp32 convolution weight (a slice of its input channels) into the packed K-major 16-bit matrix.
int launch_pack_conv_weight(const float* w_oihw, int Cout, int Cin_total, int ksize, int cin_off, int C_s,
                            void* wpacked, int Ktot, int koff, int ab_format, cudaStream_t stream);

}  // namespace cddpm
