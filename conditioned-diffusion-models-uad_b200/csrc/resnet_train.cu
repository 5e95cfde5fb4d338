// Training-mode condition encoder (see resnet_train.cuh): kernels + the planned forward / backward launch lists.
#include "resnet_train.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "backward.cuh"
#include "elementwise.cuh"
#include "ptx.cuh"

namespace cddpm {

namespace {

constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;

__device__ __forceinline__ uint16_t to_bf16(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float from_bf16(uint16_t u) { return __uint_as_float(static_cast<uint32_t>(u) << 16); }

// rows per block of the BatchNorm column reductions: about four blocks per SM, at least eight rows
int bn_rows_per_block(int M) { return std::max(8, (M + 591) / 592); }

int grid_for(size_t n, int per_block = 256) {
  return static_cast<int>(std::min<size_t>((n + per_block - 1) / per_block, 148 * 16));
}

// ---------------------------------------------------------------------------------------------- parameters
// w fp32 [cout][cin][taps] -> panel[cout][Kp] (k = tap * cin + ci, zero for k >= cin * taps) and panel_t[Kp][cout], bf16.
// Every convolution of the encoder in ONE launch: the job table travels as a kernel parameter, a block finds its job
// from the block ranges.
struct PackJob {
  const float* w;
  uint16_t *panel, *panel_t;
  int cout, cin, taps, Kp, block0;
};
constexpr int kMaxPackJobs = 56;
// input channels per pack tile: span * taps <= 288 elements per output channel row of the shared tile
__host__ __device__ inline int pack_ci_span(int cin, int taps) { return taps > 1 ? 32 : (cin < 256 ? cin : 256); }
struct PackTable {
  PackJob job[kMaxPackJobs];
  int njobs;
};
// compile-time taps and span: every index split below is a shift or a multiply-shift
template <int kTaps, int kSpan>
__device__ __forceinline__ void pack_tile(const PackJob& jb, int local, uint16_t (*tile)[32 * 9 + 2]) {
  constexpr int kRun = kSpan * kTaps;
  const int K = jb.cin * kTaps;
  const int ci_tiles = jb.cin / kSpan;
  const int co0 = (local / ci_tiles) * 32, ci0 = (local % ci_tiles) * kSpan;
  for (int idx = threadIdx.x; idx < 32 * kRun; idx += 256) {
    const int co_l = idx / kRun, r = idx - co_l * kRun;
    const int ci_l = r / kTaps, tap = r - ci_l * kTaps;
    tile[co_l][tap * kSpan + ci_l] = to_bf16(__ldg(jb.w + (static_cast<size_t>(co0 + co_l) * jb.cin + ci0) * kTaps + r));
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * kRun; idx += 256) {
    const int co_l = idx / kRun, r = idx - co_l * kRun;
    const int tap = r / kSpan, ci_l = r - tap * kSpan;
    jb.panel[static_cast<size_t>(co0 + co_l) * K + tap * jb.cin + ci0 + ci_l] = tile[co_l][r];
  }
  for (int idx = threadIdx.x; idx < 32 * kRun; idx += 256) {
    const int r = idx >> 5, co_l = idx & 31;
    const int tap = r / kSpan, ci_l = r - tap * kSpan;
    jb.panel_t[static_cast<size_t>(tap * jb.cin + ci0 + ci_l) * jb.cout + co0 + co_l] = tile[co_l][r];
  }
}

__global__ void __launch_bounds__(256) pack_panels_kernel(const __grid_constant__ PackTable t) {
  __shared__ uint16_t tile[32][32 * 9 + 2];
  int j = 0;
  while (j + 1 < t.njobs && static_cast<int>(blockIdx.x) >= t.job[j + 1].block0) ++j;
  const PackJob jb = t.job[j];
  const int local = static_cast<int>(blockIdx.x) - jb.block0;
  const int K = jb.cin * jb.taps;
  if (jb.cin % 64 != 0 || (jb.taps != 1 && jb.taps != 9)) {  // the stem (1 x 49 taps): a few thousand elements, element-wise
    const int nblk = (j + 1 < t.njobs ? t.job[j + 1].block0 : static_cast<int>(gridDim.x)) - jb.block0;
    const size_t total = static_cast<size_t>(jb.cout) * jb.Kp;
    for (size_t i = local * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(nblk) * blockDim.x) {
      const int k = static_cast<int>(i % jb.Kp);
      const int co = static_cast<int>(i / jb.Kp);
      uint16_t v = 0;
      if (k < K) {
        const int tap = k / jb.cin, ci = k - tap * jb.cin;
        v = to_bf16(__ldg(jb.w + (static_cast<size_t>(co) * jb.cin + ci) * jb.taps + tap));
      }
      jb.panel[i] = v;
      jb.panel_t[static_cast<size_t>(k) * jb.cout + co] = v;
    }
    return;
  }
  // a block owns 32 output channels x `span` input channels x every tap (span = 32 for 3x3, up to 256 for 1x1): the
  // reads, the panel rows and the transposed panel rows are all contiguous runs
  const int span = pack_ci_span(jb.cin, jb.taps);
  if (jb.taps == 9) {
    pack_tile<9, 32>(jb, local, tile);
  } else if (span == 256) {
    pack_tile<1, 256>(jb, local, tile);
  } else if (span == 128) {
    pack_tile<1, 128>(jb, local, tile);
  } else {
    pack_tile<1, 64>(jb, local, tile);
  }
}

// grad[co][ci][tap] = dwp[co][tap * cin + ci]
__global__ void unpack_grad_kernel(const float* __restrict__ dwp, int cout, int cin, int taps, float* __restrict__ grad) {
  const uint32_t total = static_cast<uint32_t>(cout) * cin * taps;
  const int K = cin * taps;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % taps);
    const int ci = static_cast<int>((i / taps) % cin);
    const int co = static_cast<int>(i / (static_cast<uint32_t>(taps) * cin));
    grad[i] = dwp[static_cast<uint32_t>(co) * K + tap * cin + ci];
  }
}

// ---------------------------------------------------------------------------------------------- forward kernels
// Element indices are 32-bit throughout (plan() refuses batches whose largest tensor would not fit): these passes move
// 8-16 bytes per thread, and the 64-bit divisions of a size_t index split made them instruction-bound (im2col: 72 %
// issue-slot utilisation at 1.4 TB/s).
// col[(n,oy,ox)][tap][c] = in[n][oy*stride+ky-pad][ox*stride+kx-pad][c] (zero outside); 8 channels per thread.
__global__ void im2col_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ col, int B, int H, int W, int C,
                              int k, int stride, int pad, int Ho, int Wo) {
  const int cv = C >> 3;
  const uint32_t total = static_cast<uint32_t>(B) * Ho * Wo * k * k * cv;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    const int tap = static_cast<int>((i / cv) % (k * k));
    const uint32_t pix = i / (static_cast<uint32_t>(cv) * k * k);
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>((pix / Wo) % Ho);
    const uint32_t n = pix / (static_cast<uint32_t>(Wo) * Ho);
    const int iy = oy * stride + tap / k - pad, ix = ox * stride + tap % k - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      val = __ldg(reinterpret_cast<const uint4*>(in + ((n * H + iy) * W + ix) * C + v * 8));
    *reinterpret_cast<uint4*>(col + (pix * k * k + tap) * C + v * 8) = val;
  }
}

// Stem 7x7 stride 2 pad 3 over one input channel as a GEMM operand: xcol[(n,oy,ox)][tap] = bf16(x[n][2 oy + ky - 3]
// [2 ox + kx - 3]) for tap < 49, zero for the padding taps 49..63.  One thread per (pixel, group of 8 taps).  The same
// matrix is the x operand of the stem's weight gradient.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, uint16_t* __restrict__ xcol, int B,
                                                          int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const uint32_t total = static_cast<uint32_t>(B) * Ho * Wo * 8;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i & 7);
    const uint32_t pix = i >> 3;
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>((pix / Wo) % Ho);
    const uint32_t n = pix / (static_cast<uint32_t>(Wo) * Ho);
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int tap = g * 8 + j;
      const int iy = oy * 2 + tap / 7 - 3, ix = ox * 2 + tap % 7 - 3;
      uint16_t v = 0;
      if (tap < 49 && iy >= 0 && iy < H && ix >= 0 && ix < W) v = to_bf16(__ldg(x + (n * H + iy) * W + ix));
      if (j & 1)
        pk[j >> 1] |= static_cast<uint32_t>(v) << 16;
      else
        pk[j >> 1] = v;
    }
    *reinterpret_cast<uint4*>(xcol + pix * 64 + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}
// grad[co][tap] = dwp[co][tap] for the 49 real taps of the 64-wide stem panel
__global__ void stem_unpack_kernel(const float* __restrict__ dwp, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 49) grad[i] = dwp[(i / 49) * 64 + i % 49];
}

// Column sums of a row-major [M][C] matrix without atomics: block g owns `rows_per_block` rows and writes its partial
// sums to partial[g][0..C) and partial[g][C..2C) (fp32; combined in float64 by the finalize kernels).  Adding ~590
// blocks' partials into one [C][2] table with atomics serialised on the few cache lines of that table (30 us for a
// 64-channel layer); the partial rows cost one coalesced store per block.
// Thread = (4 channels, row lane); the row lanes of a column group are combined in shared memory.
template <typename F>
__device__ __forceinline__ void column_partials(int M, int C, int rows_per_block, float* __restrict__ partial, F&& term) {
  __shared__ float red[256][8];
  const int cv = C >> 2;  // float4 columns; 256 / cv row lanes share a column group (one lane when C > 1024)
  const int nl = cv > 256 ? 1 : max(1, 256 / cv);
  const int m0 = blockIdx.x * rows_per_block;
  const int m1 = min(m0 + rows_per_block, M);
  float* prow = partial + static_cast<uint32_t>(blockIdx.x) * 2 * C;
  for (int c4 = threadIdx.x % cv; c4 < cv; c4 += (cv > 256 ? 256 : cv)) {
    const int lane = cv > 256 ? 0 : threadIdx.x / cv;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane < nl)
      for (int m = m0 + lane; m < m1; m += nl) term(m, c4 * 4, s, q);
    if (nl == 1) {
      *reinterpret_cast<float4*>(prow + c4 * 4) = make_float4(s[0], s[1], s[2], s[3]);
      *reinterpret_cast<float4*>(prow + C + c4 * 4) = make_float4(q[0], q[1], q[2], q[3]);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[threadIdx.x][j] = s[j];
      red[threadIdx.x][4 + j] = q[j];
    }
    __syncthreads();
    if (lane == 0) {
      float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int l = 0; l < nl; ++l)
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += red[l * cv + c4][j];
      *reinterpret_cast<float4*>(prow + c4 * 4) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(prow + C + c4 * 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
    __syncthreads();
  }
}

// float64 sum over the G partial rows of columns c and C + c.  Block = 32 channels x 32 row lanes (1024 threads): a lane
// adds at most ~19 rows, four loads in flight, so the kernel is a few L2 round trips long instead of G / 8 of them.
constexpr int kRedLanes = 32;
__device__ __forceinline__ bool reduce_partials(const float* __restrict__ partial, int G, int C, double* s0, double* s1) {
  __shared__ double red2[kRedLanes][32][2];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double a = 0.0, b = 0.0;
  if (c < C) {
    const float* p = partial + c;
    const uint32_t row = static_cast<uint32_t>(2) * C;
    int g = lane;
    for (; g + 3 * kRedLanes < G; g += 4 * kRedLanes) {
      float va[4], vb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        va[j] = p[(g + j * kRedLanes) * row];
        vb[j] = p[(g + j * kRedLanes) * row + C];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a += static_cast<double>(va[j]);
        b += static_cast<double>(vb[j]);
      }
    }
    for (; g < G; g += kRedLanes) {
      a += static_cast<double>(p[g * row]);
      b += static_cast<double>(p[g * row + C]);
    }
  }
  red2[lane][cl][0] = a;
  red2[lane][cl][1] = b;
  __syncthreads();
  if (lane >= 4) return false;
  // lanes 0..3 each fold eight rows, lane 0 folds the four results
  for (int l = lane + 4; l < kRedLanes; l += 4) {
    a += red2[l][cl][0];
    b += red2[l][cl][1];
  }
  red2[lane][cl][0] = a;
  red2[lane][cl][1] = b;
  __syncwarp();
  // the four lanes of a channel are four different warps: a block-level barrier among the 128 remaining threads
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (lane != 0 || c >= C) return false;
  for (int l = 1; l < 4; ++l) {
    a += red2[l][cl][0];
    b += red2[l][cl][1];
  }
  *s0 = a;
  *s1 = b;
  return true;
}

// partial[g] = (sum_m y[m][c], sum_m y[m][c]^2) over block g's rows
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ y, int M, int C, int rows_per_block,
                                                       float* __restrict__ partial) {
  column_partials(M, C, rows_per_block, partial, [&](int m, int c, float* s, float* q) {
    const float4 v = *reinterpret_cast<const float4*>(y + static_cast<uint32_t>(m) * C + c);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
  });
}

// Batch statistics -> (mean, rstd) for the normalisation (biased variance) and the running statistics update of
// nn.BatchNorm2d in train() mode: running = (1 - momentum) running + momentum * (mean | unbiased variance).
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int G, int M, int C,
                                                          float* __restrict__ mean, float* __restrict__ rstd,
                                                          float* __restrict__ running_mean,
                                                          float* __restrict__ running_var) {
  double s0, s1;
  if (!reduce_partials(partial, G, C, &s0, &s1)) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const double mu = s0 / M;
  double var = s1 / M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(mu);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
  const double unbiased = M > 1 ? var * (static_cast<double>(M) / (M - 1)) : var;
  running_mean[c] = (1.0f - kBnMomentum) * running_mean[c] + kBnMomentum * static_cast<float>(mu);
  running_var[c] = (1.0f - kBnMomentum) * running_var[c] + kBnMomentum * static_cast<float>(unbiased);
}

// a = relu?((gamma (y - mean) rstd + beta) * scale[n] + identity), bf16; 4 channels per thread
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const uint16_t* __restrict__ identity,
                                                       const float* __restrict__ scale, int rows_per_sample, int relu,
                                                       uint32_t M, int C, uint16_t* __restrict__ a) {
  const int cv = C >> 2;
  const uint32_t total = M * cv;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 4;
    const uint32_t m = i / cv;
    const float4 v = *reinterpret_cast<const float4*>(y + m * C + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    float o[4] = {(v.x - mu.x) * rs.x * ga.x + be.x, (v.y - mu.y) * rs.y * ga.y + be.y,
                  (v.z - mu.z) * rs.z * ga.z + be.z, (v.w - mu.w) * rs.w * ga.w + be.w};
    if (scale != nullptr) {
      const float s = scale[m / rows_per_sample];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] *= s;
    }
    if (identity != nullptr) {
      const uint2 id = *reinterpret_cast<const uint2*>(identity + m * C + c);
      o[0] += from_bf16(static_cast<uint16_t>(id.x & 0xFFFF));
      o[1] += from_bf16(static_cast<uint16_t>(id.x >> 16));
      o[2] += from_bf16(static_cast<uint16_t>(id.y & 0xFFFF));
      o[3] += from_bf16(static_cast<uint16_t>(id.y >> 16));
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    uint2 pk;
    pk.x = static_cast<uint32_t>(to_bf16(o[0])) | (static_cast<uint32_t>(to_bf16(o[1])) << 16);
    pk.y = static_cast<uint32_t>(to_bf16(o[2])) | (static_cast<uint32_t>(to_bf16(o[3])) << 16);
    *reinterpret_cast<uint2*>(a + m * C + c) = pk;
  }
}

// max_pool2d(3, 2, 1); arg[i] = window position (ky * 3 + kx) of the FIRST maximum in row-major scan order (torch's
// argmax), kept for the backward pass.
__global__ void maxpool3s2_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, uint8_t* __restrict__ arg,
                                  int B, int H, int W, int C) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const uint32_t total = static_cast<uint32_t>(B) * Ho * Wo * C;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int ox = static_cast<int>((i / C) % Wo), oy = static_cast<int>((i / (static_cast<uint32_t>(C) * Wo)) % Ho);
    const uint32_t n = i / (static_cast<uint32_t>(C) * Wo * Ho);
    float m = -INFINITY;
    int best = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = oy * 2 + ky - 1, ix = ox * 2 + kx - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const float v = from_bf16(in[((n * H + iy) * W + ix) * C + c]);
        if (v > m) {
          m = v;
          best = ky * 3 + kx;
        }
      }
    out[i] = to_bf16(m);
    arg[i] = static_cast<uint8_t>(best);
  }
}

__global__ void avgpool_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int B, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, n = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += from_bf16(in[(static_cast<uint32_t>(n) * HW + p) * C + c]);
  out[i] = s / static_cast<float>(HW);
}

// ---------------------------------------------------------------------------------------------- backward kernels
// A gradient that reaches a [B,H,W,C] tensor along up to two paths.
struct Src {
  const float* p;
  int mode;  // 0 none, 1 same shape, 2 source is the stride-2 subsampled tensor [B,H/2,W/2,C], 3 [B][C] / (H W)
};
__device__ __forceinline__ float4 src4_at(const Src& s, uint32_t n, int y, int x, int c, int H, int W, int C) {
  if (s.mode == 1) return *reinterpret_cast<const float4*>(s.p + ((n * H + y) * W + x) * C + c);
  if (s.mode == 2) {
    if ((y | x) & 1) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(s.p + ((n * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * C + c);
  }
  if (s.mode == 3) {
    const float4 v = *reinterpret_cast<const float4*>(s.p + n * C + c);
    const float r = 1.0f / static_cast<float>(H * W);
    return make_float4(v.x * r, v.y * r, v.z * r, v.w * r);
  }
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

// E = (g0 + g1) * [o > 0]; 4 channels per thread
__global__ void __launch_bounds__(256) mask_relu_kernel(Src g0, Src g1, const uint16_t* __restrict__ o, int B, int H,
                                                        int W, int C, float* __restrict__ E) {
  const int cv = C >> 2;
  const uint32_t total = static_cast<uint32_t>(B) * H * W * cv;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 4;
    const int x = static_cast<int>((i / cv) % W);
    const int y = static_cast<int>((i / (static_cast<uint32_t>(cv) * W)) % H);
    const uint32_t n = i / (static_cast<uint32_t>(cv) * W * H);
    const float4 a = src4_at(g0, n, y, x, c, H, W, C), b = src4_at(g1, n, y, x, c, H, W, C);
    const uint32_t idx = ((n * H + y) * W + x) * C + c;
    const uint2 ov = *reinterpret_cast<const uint2*>(o + idx);
    float4 e;
    e.x = from_bf16(static_cast<uint16_t>(ov.x & 0xFFFF)) > 0.f ? a.x + b.x : 0.f;
    e.y = from_bf16(static_cast<uint16_t>(ov.x >> 16)) > 0.f ? a.y + b.y : 0.f;
    e.z = from_bf16(static_cast<uint16_t>(ov.y & 0xFFFF)) > 0.f ? a.z + b.z : 0.f;
    e.w = from_bf16(static_cast<uint16_t>(ov.y >> 16)) > 0.f ? a.w + b.w : 0.f;
    *reinterpret_cast<float4*>(E + idx) = e;
  }
}

// dz = up * (relu ? a > 0 : 1) * (scale ? scale[n] : 1);  partial[g] = (sum dz, sum dz xhat), xhat = (y - mean) rstd,
// over block g's rows (the tiling of bn_stats_kernel).
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ up, const uint16_t* __restrict__ a,
                                                            int relu, const float* __restrict__ scale,
                                                            int rows_per_sample, const float* __restrict__ y,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            int M, int C, int rows_per_block,
                                                            float* __restrict__ partial) {
  column_partials(M, C, rows_per_block, partial, [&](int m, int c, float* s1, float* s2) {
    const uint32_t idx = static_cast<uint32_t>(m) * C + c;
    float4 g = *reinterpret_cast<const float4*>(up + idx);
    if (relu) {
      const uint2 av = *reinterpret_cast<const uint2*>(a + idx);
      if (!(from_bf16(static_cast<uint16_t>(av.x & 0xFFFF)) > 0.f)) g.x = 0.f;
      if (!(from_bf16(static_cast<uint16_t>(av.x >> 16)) > 0.f)) g.y = 0.f;
      if (!(from_bf16(static_cast<uint16_t>(av.y & 0xFFFF)) > 0.f)) g.z = 0.f;
      if (!(from_bf16(static_cast<uint16_t>(av.y >> 16)) > 0.f)) g.w = 0.f;
    }
    if (scale != nullptr) {
      const float sc = scale[m / rows_per_sample];
      g.x *= sc; g.y *= sc; g.z *= sc; g.w *= sc;
    }
    const float4 v = *reinterpret_cast<const float4*>(y + idx);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    s1[0] += g.x; s1[1] += g.y; s1[2] += g.z; s1[3] += g.w;
    s2[0] = fmaf(g.x, (v.x - mu.x) * rs.x, s2[0]);
    s2[1] = fmaf(g.y, (v.y - mu.y) * rs.y, s2[1]);
    s2[2] = fmaf(g.z, (v.z - mu.z) * rs.z, s2[2]);
    s2[3] = fmaf(g.w, (v.w - mu.w) * rs.w, s2[3]);
  });
}

// (S1, S2) = sum of the partial rows: the gradients of beta and gamma, and the fp32 copy bn_bwd_apply_kernel reads
// (sums[c] = S1, sums[C + c] = S2).
__global__ void __launch_bounds__(1024) bn_bwd_params_kernel(const float* __restrict__ partial, int G, int C,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ sums) {
  double s0, s1;
  if (!reduce_partials(partial, G, C, &s0, &s1)) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  dbeta[c] = static_cast<float>(s0);
  dgamma[c] = static_cast<float>(s1);
  sums[c] = static_cast<float>(s0);
  sums[C + c] = static_cast<float>(s1);
}

// dy = gamma rstd (dz - S1 / M - xhat S2 / M), bf16; 4 channels per thread
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ up, const uint16_t* __restrict__ a,
                                                           int relu, const float* __restrict__ scale,
                                                           int rows_per_sample, const float* __restrict__ y,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ sums, uint32_t M, int C,
                                                           uint16_t* __restrict__ dy) {
  const int cv = C >> 2;
  const uint32_t total = M * cv;
  const float invM = 1.0f / static_cast<float>(M);
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 4;
    const uint32_t m = i / cv;
    const uint32_t idx = m * C + c;
    float4 g = *reinterpret_cast<const float4*>(up + idx);
    if (relu) {
      const uint2 av = *reinterpret_cast<const uint2*>(a + idx);
      if (!(from_bf16(static_cast<uint16_t>(av.x & 0xFFFF)) > 0.f)) g.x = 0.f;
      if (!(from_bf16(static_cast<uint16_t>(av.x >> 16)) > 0.f)) g.y = 0.f;
      if (!(from_bf16(static_cast<uint16_t>(av.y & 0xFFFF)) > 0.f)) g.z = 0.f;
      if (!(from_bf16(static_cast<uint16_t>(av.y >> 16)) > 0.f)) g.w = 0.f;
    }
    if (scale != nullptr) {
      const float sc = scale[m / rows_per_sample];
      g.x *= sc; g.y *= sc; g.z *= sc; g.w *= sc;
    }
    const float4 v = *reinterpret_cast<const float4*>(y + idx);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 t1 = *reinterpret_cast<const float4*>(sums + c);
    const float4 t2 = *reinterpret_cast<const float4*>(sums + C + c);
    const float dz[4] = {g.x, g.y, g.z, g.w};
    const float S1[4] = {t1.x, t1.y, t1.z, t1.w};
    const float S2[4] = {t2.x, t2.y, t2.z, t2.w};
    const float yy[4] = {v.x, v.y, v.z, v.w};
    const float mm[4] = {mu.x, mu.y, mu.z, mu.w};
    const float rr[4] = {rs.x, rs.y, rs.z, rs.w};
    const float gg[4] = {ga.x, ga.y, ga.z, ga.w};
    uint16_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xhat = (yy[j] - mm[j]) * rr[j];
      o[j] = to_bf16(gg[j] * rr[j] * (dz[j] - S1[j] * invM - xhat * S2[j] * invM));
    }
    uint2 pk;
    pk.x = static_cast<uint32_t>(o[0]) | (static_cast<uint32_t>(o[1]) << 16);
    pk.y = static_cast<uint32_t>(o[2]) | (static_cast<uint32_t>(o[3]) << 16);
    *reinterpret_cast<uint2*>(dy + idx) = pk;
  }
}

// dxin[n,iy,ix,c] = sum over the taps (ky,kx) and output pixels (oy,ox) with oy*stride + ky - 1 == iy (same for x) of
// dcol[(n,oy,ox)][(ky*3+kx) * C + c]   (3x3, pad 1; dcol bf16, sum in fp32); 4 channels per thread
__global__ void __launch_bounds__(256) col2im3_kernel(const uint16_t* __restrict__ dcol, int B, int H, int W, int C,
                                                      int stride, int Ho, int Wo, float* __restrict__ dxin) {
  const int cv = C >> 2;
  const uint32_t total = static_cast<uint32_t>(B) * H * W * cv;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 4;
    const int ix = static_cast<int>((i / cv) % W);
    const int iy = static_cast<int>((i / (static_cast<uint32_t>(cv) * W)) % H);
    const uint32_t n = i / (static_cast<uint32_t>(cv) * W * H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + 1 - ky;
      if (ty < 0 || ty % stride != 0) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + 1 - kx;
        if (tx < 0 || tx % stride != 0) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const uint2 v = *reinterpret_cast<const uint2*>(dcol + ((n * Ho + oy) * Wo + ox) * (static_cast<uint32_t>(9) * C) +
                                                        (ky * 3 + kx) * C + c);
        acc[0] += from_bf16(static_cast<uint16_t>(v.x & 0xFFFF));
        acc[1] += from_bf16(static_cast<uint16_t>(v.x >> 16));
        acc[2] += from_bf16(static_cast<uint16_t>(v.y & 0xFFFF));
        acc[3] += from_bf16(static_cast<uint16_t>(v.y >> 16));
      }
    }
    *reinterpret_cast<float4*>(dxin + (((n * H + iy) * W + ix) * C + c)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// Backward of max_pool2d(3, 2, 1) as a gather over the stored argmax: an input pixel collects the gradient of the (up
// to four) windows whose first maximum it is.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(Src g0, Src g1, const uint8_t* __restrict__ arg, int B, int H,
                                                          int W, int C, float* __restrict__ gin) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int cv = C >> 2;  // 4 channels per thread
  const uint32_t total = static_cast<uint32_t>(B) * H * W * cv;
  for (uint32_t i = blockIdx.x * static_cast<uint32_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint32_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * 4;
    const int ix = static_cast<int>((i / cv) % W);
    const int iy = static_cast<int>((i / (static_cast<uint32_t>(cv) * W)) % H);
    const uint32_t n = i / (static_cast<uint32_t>(cv) * W * H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int oy = max(0, iy / 2); oy <= min(Ho - 1, (iy + 1) / 2); ++oy)
      for (int ox = max(0, ix / 2); ox <= min(Wo - 1, (ix + 1) / 2); ++ox) {
        const uint32_t pos = static_cast<uint32_t>((iy - (oy * 2 - 1)) * 3 + ix - (ox * 2 - 1));  // inside window (oy, ox)
        const uint32_t a4 = *reinterpret_cast<const uint32_t*>(arg + ((n * Ho + oy) * Wo + ox) * C + c);
        if (((a4 & 0xFF) != pos) && (((a4 >> 8) & 0xFF) != pos) && (((a4 >> 16) & 0xFF) != pos) && ((a4 >> 24) != pos))
          continue;
        const float4 u = src4_at(g0, n, oy, ox, c, Ho, Wo, C), v = src4_at(g1, n, oy, ox, c, Ho, Wo, C);
        if ((a4 & 0xFF) == pos) acc[0] += u.x + v.x;
        if (((a4 >> 8) & 0xFF) == pos) acc[1] += u.y + v.y;
        if (((a4 >> 16) & 0xFF) == pos) acc[2] += u.z + v.z;
        if ((a4 >> 24) == pos) acc[3] += u.w + v.w;
      }
    *reinterpret_cast<float4*>(gin + i * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---------------------------------------------------------------------------------------------- weight gradient GEMM
// dw[co][k] += sum_m dy[m][co] * x[m][k] over this CTA's slice of rows m.  Both operands are row-major [M][.] bf16
// matrices, i.e. MN-major UMMA operands (the contraction index m is the ROW): TMA stages [128 rows x 64 columns] boxes
// as 128-byte swizzled rows and one MMA (M = 128 co, N = 128 k, K = 16) consumes sixteen rows - the descriptor form of
// conv_wgrad.cu / attention_bwd_tc.cu.  Three stages; thread 0 produces and issues, all 128 threads drain the accumulator
// (one co row each) with red.global.add.v4.f32.
constexpr int kWgRows = 128;                         // rows of m per stage
constexpr int kWgBox = kWgRows * 128;                // bytes of one [128 x 64] bf16 box
constexpr int kWgStage = 4 * kWgBox;                 // dy x 2 column chunks + x x 2 column chunks
constexpr int kWgStages = 3;
constexpr int kWgSmem = kWgStages * kWgStage + 1024 + 128;

struct FlatWgradParams {
  CUtensorMap tmap_dy;  // {Cout, M}, box {64, 128}
  CUtensorMap tmap_x;   // {K, M}, box {64, 128}
  float* dw;
  int M, Cout, K, chunks_per_split, nchunks;
};

__device__ __forceinline__ uint64_t desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) flat_wgrad_tc_kernel(const __grid_constant__ FlatWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStage);
  uint64_t* full = bars;                // [kWgStages]
  uint64_t* empty = bars + kWgStages;   // [kWgStages]
  uint64_t* done = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);
  const int warp = threadIdx.x >> 5;
  const int k0 = blockIdx.x * 128, co0 = blockIdx.y * 128;
  const int c_begin = blockIdx.z * p.chunks_per_split;
  const int c_end = min(c_begin + p.chunks_per_split, p.nchunks);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_dy);
    tma_prefetch_desc(&p.tmap_x);
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (c_begin >= c_end) {  // an empty slice (the split count does not divide the chunk count)
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc(tmem_base, 128);
    }
    return;
  }
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 128, 1u) | (1u << 15) | (1u << 16);  // bf16, A and B MN-major
    const uint32_t base = smem_u32(smem);
    auto load = [&](int c, int st) {
      uint8_t* s = smem + st * kWgStage;
      mbar_arrive_expect_tx(&full[st], static_cast<uint32_t>(kWgStage));
      tma_load_2d(s, &p.tmap_dy, &full[st], co0, c * kWgRows);
      tma_load_2d(s + kWgBox, &p.tmap_dy, &full[st], co0 + 64, c * kWgRows);
      tma_load_2d(s + 2 * kWgBox, &p.tmap_x, &full[st], k0, c * kWgRows);
      tma_load_2d(s + 3 * kWgBox, &p.tmap_x, &full[st], k0 + 64, c * kWgRows);
    };
    auto mma = [&](int st, bool first) {
      const uint32_t a_addr = base + st * kWgStage, b_addr = a_addr + 2 * kWgBox;
#pragma unroll
      for (int j = 0; j < kWgRows / 16; ++j) {
        const uint32_t ko = static_cast<uint32_t>(j) * 16u * 128u;
        umma_f16_ss(tmem_base, desc_mn128(a_addr + ko, kWgBox), desc_mn128(b_addr + ko, kWgBox), idesc,
                    (first && j == 0) ? 0u : 1u);
      }
    };
    // chunk i of this slice lives in stage i % kWgStages; its k-th reuse waits for the MMAs of use k - 1
    const int n = c_end - c_begin;
    int loaded = 0;
    auto try_load = [&](int i) {
      const int st = i % kWgStages, k = i / kWgStages;
      if (k > 0) mbar_wait(&empty[st], static_cast<uint32_t>((k - 1) & 1));
      load(c_begin + i, st);
    };
    for (; loaded < n && loaded < kWgStages - 1; ++loaded) try_load(loaded);
    for (int i = 0; i < n; ++i) {
      if (loaded < n) {
        try_load(loaded);
        ++loaded;
      }
      const int st = i % kWgStages;
      mbar_wait(&full[st], static_cast<uint32_t>((i / kWgStages) & 1));
      tc_fence_after();
      mma(st, i == 0);
      umma_commit(&empty[st]);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int co = co0 + static_cast<int>(threadIdx.x);
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + lane_off + c0, v);
    tmem_ld_wait();
    if (co < p.Cout) {
      float* row = p.dw + static_cast<size_t>(co) * p.K + k0 + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)  // K is a multiple of 8: a group of four is inside the row or outside it
        if (k0 + c0 + j < p.K)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "f"(__uint_as_float(v[j])),
                       "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                       : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

int launch_flat_wgrad(const uint16_t* dy, const uint16_t* x, int M, int Cout, int K, float* dw, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(flat_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
    attr_set = true;
  }
  FlatWgradParams p;
  memset(&p, 0, sizeof(p));
  p.dw = dw;
  p.M = M;
  p.Cout = Cout;
  p.K = K;
  p.nchunks = (M + kWgRows - 1) / kWgRows;
  const uint32_t box[2] = {64u, static_cast<uint32_t>(kWgRows)};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Cout) * 2};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_dy, dy, 2, dims, strides, box));
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_x, x, 2, dims, strides, box));
  }
  const int tiles = ((K + 127) / 128) * ((Cout + 127) / 128);
  int splits = (2 * 148 + tiles - 1) / tiles;
  splits = std::max(1, std::min(splits, p.nchunks));
  p.chunks_per_split = (p.nchunks + splits - 1) / splits;
  splits = (p.nchunks + p.chunks_per_split - 1) / p.chunks_per_split;
  dim3 grid((K + 127) / 128, (Cout + 127) / 128, splits);
  flat_wgrad_tc_kernel<<<grid, 128, kWgSmem, stream>>>(p);
  return check_launch("flat_wgrad_tc_kernel");
}

}  // namespace

int launch_flat_wgrad_bf16(const void* dy, const void* x, int M, int Cout, int K, float* dw, cudaStream_t stream) {
  if (!dy || !x || !dw) return fail(kInvalidArgument, "flat_wgrad: null pointer");
  if (M < 1 || Cout < 8 || K < 8 || Cout % 8 != 0 || K % 8 != 0)
    return fail(kInvalidArgument, "flat_wgrad: Cout and K must be multiples of 8");
  return launch_flat_wgrad(reinterpret_cast<const uint16_t*>(dy), reinterpret_cast<const uint16_t*>(x), M, Cout, K, dw,
                           stream);
}

// ================================================================================================ engine
ResNetTrainEngine::~ResNetTrainEngine() {
  free_acts();
  for (void* p : owned_) cudaFree(p);
  if (cap_stream_ != nullptr) cudaStreamDestroy(cap_stream_);
  if (side_stream_ != nullptr) cudaStreamDestroy(side_stream_);
  if (ev_fork_ != nullptr) cudaEventDestroy(ev_fork_);
  if (ev_join_ != nullptr) cudaEventDestroy(ev_join_);
}

void ResNetTrainEngine::drop_graphs() {
  for (auto* slots : {&fwd_graphs_, &bwd_graphs_}) {
    for (GraphSlot& g : *slots)
      if (g.exec != nullptr) cudaGraphExecDestroy(g.exec);
    slots->clear();
  }
}

static bool enc_graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("CDDPM_ENC_GRAPH");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

// Runs a launch list on `stream`: eagerly for a key seen for the first time, as a captured graph from then on.  At most
// four graphs per list (least recently used is replaced), so a caller that hands in new buffers on every step costs
// one capture per step at worst and nothing in correctness.
// The ops of a list in order on `stream`; ops flagged in `side` go to a second stream that forks off `stream` in front
// of each of them and is joined at the end.  The weight-gradient GEMMs of the backward list hang off the data-gradient
// chain (they read dy and the forward activations, write their own slice of the gradients): on the side branch their
// 12-15 us launch floors run beside the chain instead of inside it.  CDDPM_ENC_SIDE=0 keeps one stream.
int ResNetTrainEngine::run_ops(std::vector<std::function<int(cudaStream_t)>>& ops, const std::vector<uint8_t>* side,
                               cudaStream_t stream) {
  static const bool side_on = [] {
    const char* e = getenv("CDDPM_ENC_SIDE");
    return !(e != nullptr && e[0] == '0');
  }();
  if (side == nullptr || !side_on || side->size() != ops.size()) {
    for (auto& op : ops) CDDPM_TRY(op(stream));
    return kOk;
  }
  if (side_stream_ == nullptr) {
    CDDPM_CUDA(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
    CDDPM_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    CDDPM_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
  }
  bool forked = false;
  int st = kOk;
  for (size_t i = 0; i < ops.size() && st == kOk; ++i) {
    if ((*side)[i]) {
      CDDPM_CUDA(cudaEventRecord(ev_fork_, stream));
      CDDPM_CUDA(cudaStreamWaitEvent(side_stream_, ev_fork_, 0));
      st = ops[i](side_stream_);
      forked = true;
    } else {
      st = ops[i](stream);
    }
  }
  if (forked) {  // always join, also on an error: a capture must not end with a dangling branch
    const cudaError_t e1 = cudaEventRecord(ev_join_, side_stream_);
    const cudaError_t e2 = cudaStreamWaitEvent(stream, ev_join_, 0);
    if (st == kOk) {
      CDDPM_CUDA(e1);
      CDDPM_CUDA(e2);
    }
  }
  return st;
}

int ResNetTrainEngine::run_list(std::vector<std::function<int(cudaStream_t)>>& ops, std::vector<GraphSlot>& slots,
                                const std::vector<const void*>& key, cudaStream_t stream,
                                const std::vector<uint8_t>* side) {
  auto eager = [&]() -> int { return run_ops(ops, side, stream); };
  // a caller whose buffers never repeat would pay a capture on every second call: once captures clearly outnumber
  // replays, stay eager
  if (!enc_graphs_enabled() || (graph_captures_ >= 8 && graph_replays_ < 2 * graph_captures_)) return eager();
  GraphSlot* slot = nullptr;
  for (GraphSlot& g : slots)
    if (g.key == key) slot = &g;
  if (slot == nullptr) {
    if (slots.size() < 4) {
      slots.emplace_back();
      slot = &slots.back();
    } else {
      slot = &slots[0];
      for (GraphSlot& g : slots)
        if (g.last_use < slot->last_use) slot = &g;
      if (slot->exec != nullptr) cudaGraphExecDestroy(slot->exec);
      *slot = GraphSlot();
    }
    slot->key = key;
  }
  slot->last_use = ++use_clock_;
  if (slot->exec == nullptr) {
    if (slot->seen++ == 0) return eager();
    if (cap_stream_ == nullptr) CDDPM_CUDA(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
    CDDPM_CUDA(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal));
    const int st = run_ops(ops, side, cap_stream_);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cap_stream_, &graph);
    if (st != kOk || ce != cudaSuccess) {
      if (graph != nullptr) cudaGraphDestroy(graph);
      return st != kOk ? st : check_cuda(ce, "cudaStreamEndCapture (encoder)");
    }
    const cudaError_t ie = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    CDDPM_TRY(check_cuda(ie, "cudaGraphInstantiate (encoder)"));
    ++graph_captures_;
  } else {
    ++graph_replays_;
  }
  CDDPM_CUDA(cudaGraphLaunch(slot->exec, stream));
  return kOk;
}

void ResNetTrainEngine::free_acts() {
  drop_graphs();
  for (void* p : act_owned_) cudaFree(p);
  act_owned_.clear();
  fwd_ops_.clear();
  bwd_ops_.clear();
  bwd_side_.clear();
  planned_B_ = 0;
  forward_done_ = false;
}

template <typename T>
int ResNetTrainEngine::dalloc(T** p, size_t n, std::vector<void*>* pool) {
  void* q = nullptr;
  CDDPM_CUDA(cudaMalloc(&q, n * sizeof(T) + 1024));
  pool->push_back(q);
  *p = reinterpret_cast<T*>(q);
  return kOk;
}

int ResNetTrainEngine::add_entry(const std::string& name, int64_t numel, int is_param) {
  Entry e;
  e.name = name;
  e.numel = numel;
  e.is_param = is_param;
  if (is_param) {
    e.goff = grad_total_;
    grad_total_ += (numel + 3) & ~int64_t(3);  // 16-byte aligned slots
  }
  entries_.push_back(e);
  return static_cast<int>(entries_.size()) - 1;
}

int ResNetTrainEngine::add_unit(const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride,
                                int pad) {
  Unit u;
  u.cin = cin;
  u.cout = cout;
  u.k = k;
  u.stride = stride;
  u.pad = pad;
  u.e_w = add_entry(conv + ".weight", static_cast<int64_t>(cout) * cin * k * k, 1);
  u.e_gamma = add_entry(bn + ".weight", cout, 1);
  u.e_beta = add_entry(bn + ".bias", cout, 1);
  u.e_mean = add_entry(bn + ".running_mean", cout, 0);
  u.e_var = add_entry(bn + ".running_var", cout, 0);
  const size_t K = cin == 1 ? 64 : static_cast<size_t>(cin) * k * k;  // the stem's 49 taps are padded to 64
  CDDPM_TRY(dalloc(&u.panel, static_cast<size_t>(cout) * K, &owned_));
  CDDPM_TRY(dalloc(&u.panel_t, static_cast<size_t>(cout) * K, &owned_));
  units_.push_back(u);
  return kOk;
}

int ResNetTrainEngine::init(int image_h, int image_w, int cond_dim) {
  H_ = image_h;
  W_ = image_w;
  cond_dim_ = cond_dim;
  if (image_h % 32 != 0 || image_w % 32 != 0) return fail(kUnsupported, "encoder: image size must be a multiple of 32");
  CDDPM_TRY(add_unit("conv1", "bn1", 1, 64, 7, 2, 3));  // unit 0: the stem (direct kernels)
  const int layers[4] = {3, 4, 6, 3};
  const int widths[4] = {64, 128, 256, 512};
  int cin = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < layers[li]; ++bi) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      const int w = widths[li];
      const int stride = (bi == 0 && li > 0) ? 2 : 1;
      Block b;
      CDDPM_TRY(add_unit(p + ".conv1", p + ".bn1", cin, w, 1, 1, 0));
      b.c1 = static_cast<int>(units_.size()) - 1;
      CDDPM_TRY(add_unit(p + ".conv2", p + ".bn2", w, w, 3, stride, 1));
      b.c2 = static_cast<int>(units_.size()) - 1;
      CDDPM_TRY(add_unit(p + ".conv3", p + ".bn3", w, 4 * w, 1, 1, 0));
      b.c3 = static_cast<int>(units_.size()) - 1;
      if (bi == 0) {
        CDDPM_TRY(add_unit(p + ".downsample.0", p + ".downsample.1", cin, 4 * w, 1, stride, 0));
        b.down = static_cast<int>(units_.size()) - 1;
      }
      blocks_.push_back(b);
      cin = 4 * w;
    }
  }
  e_fc_w_ = add_entry("fc.weight", static_cast<int64_t>(cond_dim) * 2048, 1);
  e_fc_b_ = add_entry("fc.bias", cond_dim, 1);
  return kOk;
}

int ResNetTrainEngine::entry_info(int i, const char** name, int64_t* numel, int* is_param) const {
  if (i < 0 || i >= entry_count()) return fail(kInvalidArgument, "entry index out of range");
  *name = entries_[i].name.c_str();
  *numel = entries_[i].numel;
  *is_param = entries_[i].is_param;
  return kOk;
}

int ResNetTrainEngine::grad_offset(int i, int64_t* off) const {
  if (i < 0 || i >= entry_count()) return fail(kInvalidArgument, "entry index out of range");
  *off = entries_[i].goff;
  return kOk;
}

int ResNetTrainEngine::push_gemm(std::vector<std::function<int(cudaStream_t)>>* ops, const void* a, int rows, int K,
                                 const void* panel, int N, void* out, bool out_f32) {
  ConvDesc d;
  d.num_src = 1;
  d.src[0] = a;
  d.src_c[0] = K;
  d.src_taps[0] = 1;
  d.flat_rows = rows;
  d.Cout = N;
  d.wpacked = panel;
  d.out = out;
  d.out_is_f32 = out_f32 ? 1 : 0;
  d.ab_format = 1;  // bf16
  auto p = std::make_shared<ConvIgemmParams>();
  CDDPM_TRY(build_conv_params(d, p.get()));
  // (splitting K over idle SMs for the deep layers' few-tile GEMMs was measured: 23 -> 12 us for those launches, but the
  // zero-fill launches it needs cancel the gain and the atomic accumulation makes the forward irreproducible)
  ops->push_back([p](cudaStream_t s) { return launch_conv_igemm(*p, s); });
  return kOk;
}

// conv (im2col + GEMM) -> batch statistics -> finalize (+ running statistics); the apply step is pushed by the caller
int ResNetTrainEngine::plan_unit_forward(Unit& u, const uint16_t* in, int Hin, int Win, int B) {
  u.in = in;
  u.Hin = Hin;
  u.Win = Win;
  u.Hout = (Hin + 2 * u.pad - u.k) / u.stride + 1;
  u.Wout = (Win + 2 * u.pad - u.k) / u.stride + 1;
  const int M = B * u.Hout * u.Wout;
  const int K = u.cin * u.k * u.k;
  const uint16_t* operand = in;
  if (u.k > 1 || u.stride > 1) {
    CDDPM_TRY(dalloc(&u.col, static_cast<size_t>(M) * K, &act_owned_));
    uint16_t* col = u.col;
    const int cin = u.cin, k = u.k, stride = u.stride, pad = u.pad, ho = u.Hout, wo = u.Wout;
    fwd_ops_.push_back([=](cudaStream_t s) {
      const size_t total = static_cast<size_t>(B) * ho * wo * k * k * (cin / 8);
      im2col_kernel<<<grid_for(total), 256, 0, s>>>(in, col, B, Hin, Win, cin, k, stride, pad, ho, wo);
      return check_launch("im2col_kernel");
    });
    operand = u.col;
  }
  CDDPM_TRY(dalloc(&u.y, static_cast<size_t>(M) * u.cout, &act_owned_));
  CDDPM_TRY(dalloc(&u.mean, static_cast<size_t>(u.cout), &act_owned_));
  CDDPM_TRY(dalloc(&u.rstd, static_cast<size_t>(u.cout), &act_owned_));
  CDDPM_TRY(dalloc(&u.a, static_cast<size_t>(M) * u.cout, &act_owned_));
  CDDPM_TRY(push_gemm(&fwd_ops_, operand, M, K, u.panel, u.cout, u.y, true));
  push_stats(u, M);
  return kOk;
}

// batch statistics of u.y [M][cout] -> mean, rstd, running statistics
void ResNetTrainEngine::push_stats(const Unit& u, int M) {
  float* y = u.y;
  float *mean = u.mean, *rstd = u.rstd;
  const int C = u.cout, e_mean = u.e_mean, e_var = u.e_var;
  ResNetTrainEngine* self = this;
  fwd_ops_.push_back([=](cudaStream_t s) {
    const int rpb = bn_rows_per_block(M), G = (M + rpb - 1) / rpb;
    bn_stats_kernel<<<G, 256, 0, s>>>(y, M, C, rpb, self->partial_);
    CDDPM_TRY(check_launch("bn_stats_kernel"));
    bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, s>>>(self->partial_, G, M, C, mean, rstd,
                                                     const_cast<float*>(self->values_[e_mean]),
                                                     const_cast<float*>(self->values_[e_var]));
    return check_launch("bn_finalize_kernel");
  });
}

// BatchNorm backward (the upstream gradient and its options are bound by the caller through `up`), weight gradient,
// data gradient.  Pushes: reduce, params, apply, wgrad, unpack, [dgrad, col2im].
int ResNetTrainEngine::plan_unit_backward(Unit& u, int B, bool need_dx) {
  const int M = B * u.Hout * u.Wout;
  const int K = u.cin * u.k * u.k;
  if (u.k > 1) CDDPM_TRY(dalloc(&u.dwp, static_cast<size_t>(u.cout) * K, &act_owned_));
  {
    const uint16_t* dy = u.dy;
    const uint16_t* operand = (u.col != nullptr) ? u.col : u.in;
    float* dwp = u.dwp;
    const int cout = u.cout, cin = u.cin, taps = u.k * u.k, e_w = u.e_w;
    ResNetTrainEngine* self = this;
    bwd_ops_.push_back([=](cudaStream_t s) -> int {
      float* g = self->cur_grads_ + self->entries_[e_w].goff;
      // 1x1: the panel order [cout][cin] IS the parameter layout, the GEMM accumulates into the gradient itself
      float* acc = taps == 1 ? g : dwp;
      CDDPM_CUDA(cudaMemsetAsync(acc, 0, static_cast<size_t>(cout) * K * sizeof(float), s));
      CDDPM_TRY(launch_flat_wgrad(dy, operand, M, cout, K, acc, s));
      if (taps == 1) return kOk;
      unpack_grad_kernel<<<grid_for(static_cast<size_t>(cout) * K), 256, 0, s>>>(dwp, cout, cin, taps, g);
      return check_launch("unpack_grad_kernel");
    });
    bwd_side_.resize(bwd_ops_.size(), 0);
    bwd_side_.back() = 1;  // off the data-gradient chain: runs on the side branch (run_ops)
  }
  if (!need_dx) return kOk;
  if (u.k == 3) {
    // the im2col-space gradient [M][9 cin] is the largest tensor of the backward pass: bf16 (it is summed over the nine
    // taps in fp32 right away)
    uint16_t* dcol16 = nullptr;
    CDDPM_TRY(dalloc(&dcol16, static_cast<size_t>(M) * K, &act_owned_));
    CDDPM_TRY(push_gemm(&bwd_ops_, u.dy, M, u.cout, u.panel_t, K, dcol16, false));
    const size_t n_in = static_cast<size_t>(B) * u.Hin * u.Win * u.cin;
    CDDPM_TRY(dalloc(&u.dxin, n_in, &act_owned_));
    const uint16_t* dcol = dcol16;
    float* dxin = u.dxin;
    const int H = u.Hin, W = u.Win, C = u.cin, stride = u.stride, ho = u.Hout, wo = u.Wout;
    bwd_ops_.push_back([=](cudaStream_t s) {
      col2im3_kernel<<<grid_for(n_in / 4), 256, 0, s>>>(dcol, B, H, W, C, stride, ho, wo, dxin);
      return check_launch("col2im3_kernel");
    });
  } else {
    CDDPM_TRY(dalloc(&u.dx, static_cast<size_t>(M) * K, &act_owned_));
    CDDPM_TRY(push_gemm(&bwd_ops_, u.dy, M, u.cout, u.panel_t, K, u.dx, true));
    u.dxin = u.dx;  // 1x1: the GEMM output is the data gradient (at the output resolution when strided)
  }
  return kOk;
}

// BatchNorm backward of unit u given the upstream gradient `up` [M][cout]: partial sums -> (dgamma, dbeta, S1, S2) ->
// dy (bf16)
int ResNetTrainEngine::launch_bn_backward(const Unit& u, const float* up, int relu, const float* scale,
                                          int rows_per_sample, int M, cudaStream_t s) {
  const int rpb = bn_rows_per_block(M), G = (M + rpb - 1) / rpb;
  bn_bwd_reduce_kernel<<<G, 256, 0, s>>>(up, u.a, relu, scale, rows_per_sample, u.y, u.mean, u.rstd, M, u.cout, rpb,
                                         partial_);
  CDDPM_TRY(check_launch("bn_bwd_reduce_kernel"));
  float* g = cur_grads_;
  bn_bwd_params_kernel<<<(u.cout + 31) / 32, 1024, 0, s>>>(partial_, G, u.cout, g + entries_[u.e_gamma].goff,
                                                          g + entries_[u.e_beta].goff, u.bsum);
  CDDPM_TRY(check_launch("bn_bwd_params_kernel"));
  bn_bwd_apply_kernel<<<grid_for(static_cast<size_t>(M) * (u.cout / 4)), 256, 0, s>>>(
      up, u.a, relu, scale, rows_per_sample, u.y, u.mean, u.rstd, values_[u.e_gamma], u.bsum, static_cast<size_t>(M),
      u.cout, u.dy);
  return check_launch("bn_bwd_apply_kernel");
}

int ResNetTrainEngine::plan(int B) {
  free_acts();
  // 32-bit element indices in the elementwise kernels: the largest tensor is layer1's im2col matrix [B H/4 W/4][576]
  if (static_cast<double>(B) * (H_ / 4) * (W_ / 4) * 576.0 >= 4.0e9 || static_cast<double>(B) * H_ * W_ * 16.0 >= 4.0e9)
    return fail(kUnsupported, "encoder_train: batch too large for one call (32-bit tensor indices)");
  ResNetTrainEngine* self = this;
  // ---- parameter panels (every forward: the optimizer changed the weights), one launch for all convolutions
  if (units_.size() > static_cast<size_t>(kMaxPackJobs)) return fail(kUnsupported, "encoder: pack table too small");
  CDDPM_TRY(dalloc(&partial_, static_cast<size_t>(600) * 2 * 2048, &act_owned_));
  fwd_ops_.push_back([=](cudaStream_t s) {
    PackTable t;
    memset(&t, 0, sizeof(t));
    int block = 0;
    for (size_t ui = 0; ui < self->units_.size(); ++ui) {
      const Unit& u = self->units_[ui];
      PackJob& j = t.job[t.njobs++];
      j.w = self->values_[u.e_w];
      j.panel = u.panel;
      j.panel_t = u.panel_t;
      j.cout = u.cout;
      j.cin = u.cin;
      j.taps = u.k * u.k;
      j.Kp = u.cin == 1 ? 64 : u.cin * u.k * u.k;
      j.block0 = block;
      block += (u.cin % 64 == 0 && (j.taps == 1 || j.taps == 9)) ? (u.cout / 32) * (u.cin / pack_ci_span(u.cin, j.taps))  // 32 x span (x taps) tiles
                               : static_cast<int>((static_cast<size_t>(j.cout) * j.Kp + 1023) / 1024);
    }
    pack_panels_kernel<<<block, 256, 0, s>>>(t);
    return check_launch("pack_panels_kernel");
  });
  // ---- stem: conv 7x7 s2 (raw) -> BN (batch statistics) + ReLU -> maxpool 3x3 s2
  Unit& st = units_[0];
  st.Hin = H_;
  st.Win = W_;
  st.Hout = H_ / 2;
  st.Wout = W_ / 2;
  const int Ms = B * st.Hout * st.Wout;
  CDDPM_TRY(dalloc(&st.y, static_cast<size_t>(Ms) * 64, &act_owned_));
  CDDPM_TRY(dalloc(&st.mean, 64, &act_owned_));
  CDDPM_TRY(dalloc(&st.rstd, 64, &act_owned_));
  CDDPM_TRY(dalloc(&st.a, static_cast<size_t>(Ms) * 64, &act_owned_));
  CDDPM_TRY(dalloc(&stem_col_, static_cast<size_t>(Ms) * 64, &act_owned_));
  {
    uint16_t* xcol = stem_col_;
    const int Hin = H_, Win = W_;
    fwd_ops_.push_back([=](cudaStream_t s) {
      stem_im2col_kernel<<<grid_for(static_cast<size_t>(Ms) * 8), 256, 0, s>>>(self->cur_x_, xcol, B, Hin, Win);
      return check_launch("stem_im2col_kernel");
    });
  }
  CDDPM_TRY(push_gemm(&fwd_ops_, stem_col_, Ms, 64, st.panel, 64, st.y, true));
  push_stats(st, Ms);
  {
    const Unit u0 = st;
    fwd_ops_.push_back([=](cudaStream_t s) {
      bn_apply_kernel<<<grid_for(static_cast<size_t>(Ms) * 16), 256, 0, s>>>(
          u0.y, u0.mean, u0.rstd, self->values_[u0.e_gamma], self->values_[u0.e_beta], nullptr, nullptr, 1, 1,
          static_cast<size_t>(Ms), 64, u0.a);
      return check_launch("bn_apply_kernel");
    });
  }
  int H = (st.Hout + 2 - 3) / 2 + 1, W = (st.Wout + 2 - 3) / 2 + 1;
  CDDPM_TRY(dalloc(&pool_a_, static_cast<size_t>(B) * H * W * 64, &act_owned_));
  CDDPM_TRY(dalloc(&pool_arg_, static_cast<size_t>(B) * H * W * 64, &act_owned_));
  {
    const uint16_t* src = st.a;
    uint16_t* dst = pool_a_;
    uint8_t* parg = pool_arg_;
    const int h = st.Hout, w = st.Wout;
    fwd_ops_.push_back([=](cudaStream_t s) {
      const size_t total = static_cast<size_t>(B) * H * W * 64;
      maxpool3s2_kernel<<<grid_for(total), 256, 0, s>>>(src, dst, parg, B, h, w, 64);
      return check_launch("maxpool3s2_kernel");
    });
  }
  // ---- bottleneck blocks
  const uint16_t* x = pool_a_;
  int C = 64;
  for (size_t bi = 0; bi < blocks_.size(); ++bi) {
    Block& b = blocks_[bi];
    Unit& c1 = units_[b.c1];
    Unit& c2 = units_[b.c2];
    Unit& c3 = units_[b.c3];
    auto apply = [&](const Unit& u, const uint16_t* identity, bool scaled, int relu) {
      const Unit uu = u;
      const int M = B * u.Hout * u.Wout, rows = u.Hout * u.Wout;
      const int blk = static_cast<int>(bi);
      fwd_ops_.push_back([=](cudaStream_t s) {
        const float* scale = (scaled && self->cur_drop_ != nullptr) ? self->cur_drop_ + static_cast<size_t>(blk) * B : nullptr;
        bn_apply_kernel<<<grid_for(static_cast<size_t>(M) * (uu.cout / 4)), 256, 0, s>>>(
            uu.y, uu.mean, uu.rstd, self->values_[uu.e_gamma], self->values_[uu.e_beta], identity, scale, rows, relu,
            static_cast<size_t>(M), uu.cout, uu.a);
        return check_launch("bn_apply_kernel");
      });
    };
    CDDPM_TRY(plan_unit_forward(c1, x, H, W, B));
    apply(c1, nullptr, false, 1);
    CDDPM_TRY(plan_unit_forward(c2, c1.a, H, W, B));
    apply(c2, nullptr, false, 1);
    const uint16_t* identity = x;
    if (b.down >= 0) {
      Unit& cd = units_[b.down];
      CDDPM_TRY(plan_unit_forward(cd, x, H, W, B));
      apply(cd, nullptr, false, 0);
      identity = cd.a;
    }
    CDDPM_TRY(plan_unit_forward(c3, c2.a, c2.Hout, c2.Wout, B));
    apply(c3, identity, true, 1);  // out = relu(bn3(.) * drop_scale + identity)
    x = c3.a;
    H = c3.Hout;
    W = c3.Wout;
    C = c3.cout;
  }
  // ---- head: global average pool + fc
  CDDPM_TRY(dalloc(&pooled_, static_cast<size_t>(B) * C, &act_owned_));
  CDDPM_TRY(dalloc(&dpooled_, static_cast<size_t>(B) * C, &act_owned_));
  {
    const uint16_t* src = x;
    float* pooled = pooled_;
    const int hw = H * W, cc = C, cd = cond_dim_;
    fwd_ops_.push_back([=](cudaStream_t s) {
      avgpool_kernel<<<(B * cc + 255) / 256, 256, 0, s>>>(src, pooled, B, hw, cc);
      CDDPM_TRY(check_launch("avgpool_kernel"));
      return launch_linear(pooled, cc, self->values_[self->e_fc_w_], self->values_[self->e_fc_b_], self->cur_out_, cd, B,
                           cc, cd, 0, s);
    });
  }

  // ================================================================ backward plan (reverse order)
  {
    float *pooled = pooled_, *dpooled = dpooled_;
    const int cc = C, cd = cond_dim_;
    bwd_ops_.push_back([=](cudaStream_t s) {
      float* g = self->cur_grads_;
      CDDPM_TRY(launch_linear_bwd_weight(self->cur_dout_, cd, pooled, cc, 0, g + self->entries_[self->e_fc_w_].goff,
                                         g + self->entries_[self->e_fc_b_].goff, B, cc, cd, s));
      return launch_linear_bwd_input(self->cur_dout_, cd, self->values_[self->e_fc_w_], nullptr, 1, dpooled, cc, nullptr,
                                     0, B, cc, cd, s);
    });
  }
  GradSrc g0{dpooled_, 3}, g1{nullptr, 0};
  for (int bi = static_cast<int>(blocks_.size()) - 1; bi >= 0; --bi) {
    Block& b = blocks_[bi];
    Unit& c1 = units_[b.c1];
    Unit& c2 = units_[b.c2];
    Unit& c3 = units_[b.c3];
    const int Ho = c3.Hout, Wo = c3.Wout, C4 = c3.cout;
    const int Mo = B * Ho * Wo;
    CDDPM_TRY(dalloc(&b.E, static_cast<size_t>(Mo) * C4, &act_owned_));
    {
      const Src s0{g0.p, g0.mode}, s1{g1.p, g1.mode};
      const uint16_t* o = c3.a;
      float* E = b.E;
      bwd_ops_.push_back([=](cudaStream_t s) {
        mask_relu_kernel<<<grid_for(static_cast<size_t>(Mo) * (C4 / 4)), 256, 0, s>>>(s0, s1, o, B, Ho, Wo, C4, E);
        return check_launch("mask_relu_kernel");
      });
    }
    // BatchNorm backward of a unit given its upstream gradient
    auto bn_bwd = [&](Unit& u, const float* up, int relu, bool scaled) -> int {
      const int M = B * u.Hout * u.Wout, rows = u.Hout * u.Wout;
      CDDPM_TRY(dalloc(&u.dy, static_cast<size_t>(M) * u.cout, &act_owned_));
      CDDPM_TRY(dalloc(&u.bsum, static_cast<size_t>(u.cout) * 2, &act_owned_));
      const Unit uu = u;
      const int blk = bi;
      bwd_ops_.push_back([=](cudaStream_t s) {
        const float* scale = (scaled && self->cur_drop_ != nullptr) ? self->cur_drop_ + static_cast<size_t>(blk) * B : nullptr;
        return self->launch_bn_backward(uu, up, relu, scale, rows, M, s);
      });
      return kOk;
    };
    CDDPM_TRY(bn_bwd(c3, b.E, 0, true));
    CDDPM_TRY(plan_unit_backward(c3, B, true));  // dx: gradient of a2 [Mo][w]
    CDDPM_TRY(bn_bwd(c2, c3.dxin, 1, false));
    CDDPM_TRY(plan_unit_backward(c2, B, true));  // dxin: gradient of a1 [M_in][w] after col2im
    CDDPM_TRY(bn_bwd(c1, c2.dxin, 1, false));
    CDDPM_TRY(plan_unit_backward(c1, B, true));  // dxin: gradient of the block input [M_in][cin]
    GradSrc n0{c1.dxin, 1}, n1;
    if (b.down >= 0) {
      Unit& cd = units_[b.down];
      CDDPM_TRY(bn_bwd(cd, b.E, 0, false));
      CDDPM_TRY(plan_unit_backward(cd, B, true));  // [Mo][cin] at the output resolution
      n1 = GradSrc{cd.dxin, cd.stride == 2 ? 2 : 1};
    } else {
      n1 = GradSrc{b.E, 1};
    }
    g0 = n0;
    g1 = n1;
  }
  // ---- maxpool backward -> stem BatchNorm backward -> stem weight gradient
  {
    Unit& st0 = units_[0];
    const int Hs = st0.Hout, Ws = st0.Wout;
    CDDPM_TRY(dalloc(&stem_g_, static_cast<size_t>(Ms) * 64, &act_owned_));
    CDDPM_TRY(dalloc(&st0.dy, static_cast<size_t>(Ms) * 64, &act_owned_));
    CDDPM_TRY(dalloc(&st0.bsum, 128, &act_owned_));
    CDDPM_TRY(dalloc(&stem_dwp_, 64 * 64, &act_owned_));
    const Src s0{g0.p, g0.mode}, s1{g1.p, g1.mode};
    const Unit u0 = st0;
    float* sg = stem_g_;
    float* sdw = stem_dwp_;
    const uint16_t* xcol = stem_col_;
    const uint8_t* parg = pool_arg_;
    bwd_ops_.push_back([=](cudaStream_t s) {
      maxpool_bwd_kernel<<<grid_for(static_cast<size_t>(Ms) * 16), 256, 0, s>>>(s0, s1, parg, B, Hs, Ws, 64, sg);
      CDDPM_TRY(check_launch("maxpool_bwd_kernel"));
      CDDPM_TRY(self->launch_bn_backward(u0, sg, 1, nullptr, 1, Ms, s));
      // weight gradient: the tensor-core GEMM over the forward's im2col matrix (taps padded to 64)
      CDDPM_CUDA(cudaMemsetAsync(sdw, 0, 64 * 64 * sizeof(float), s));
      CDDPM_TRY(launch_flat_wgrad(u0.dy, xcol, Ms, 64, 64, sdw, s));
      stem_unpack_kernel<<<(64 * 49 + 255) / 256, 256, 0, s>>>(sdw, self->cur_grads_ + self->entries_[u0.e_w].goff);
      return check_launch("stem_unpack_kernel");
    });
  }
  planned_B_ = B;
  return kOk;
}

int ResNetTrainEngine::forward(const float* const* values, int count, const float* x, const float* drop_scale,
                               float* out, int B, cudaStream_t stream) {
  if (!values || !x || !out) return fail(kInvalidArgument, "encoder_train_forward: null pointer");
  if (count != entry_count()) return fail(kInvalidArgument, "encoder_train_forward: wrong number of entries");
  if (B < 1) return fail(kInvalidArgument, "encoder_train_forward: empty batch");
  values_.assign(values, values + count);
  for (const float* v : values_)
    if (v == nullptr) return fail(kInvalidArgument, "encoder_train_forward: null entry pointer");
  if (B != planned_B_) {
    CDDPM_CUDA(cudaDeviceSynchronize());
    int st = plan(B);
    if (st != kOk) {
      free_acts();
      return st;
    }
  }
  cur_x_ = x;
  cur_drop_ = drop_scale;
  cur_out_ = out;
  std::vector<const void*> key = {x, drop_scale, out};
  key.insert(key.end(), values_.begin(), values_.end());
  CDDPM_TRY(run_list(fwd_ops_, fwd_graphs_, key, stream));
  forward_done_ = true;
  return kOk;
}

int ResNetTrainEngine::backward(const float* dout, float* grads, int B, cudaStream_t stream) {
  if (!dout || !grads) return fail(kInvalidArgument, "encoder_train_backward: null pointer");
  if (B != planned_B_ || !forward_done_) return fail(kNotReady, "encoder_train_backward: run the forward of this batch first");
  cur_dout_ = dout;
  cur_grads_ = grads;
  // the backward list also binds the forward's drop scales and the parameters (gamma, fc weight)
  std::vector<const void*> key = {dout, grads, cur_drop_};
  key.insert(key.end(), values_.begin(), values_.end());
  bwd_side_.resize(bwd_ops_.size(), 0);
  CDDPM_TRY(run_list(bwd_ops_, bwd_graphs_, key, stream, &bwd_side_));
  forward_done_ = false;
  return kOk;
}

}  // namespace cddpm
