// Bandwidth-bound kernels of the UNet forward (see elementwise.cuh).  16-byte vector accesses along the channel
// (innermost NHWC) dimension, fp32 math, one rounding to the 16-bit activation type on store.
#include "elementwise.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

namespace cddpm {

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, int fmt, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 t;
    if (fmt == 1) {
      t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
    } else {
      t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
    }
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int fmt) {
  uint4 o;
  o.x = pack2(f[0], f[1], fmt);
  o.y = pack2(f[2], f[3], fmt);
  o.z = pack2(f[4], f[5], fmt);
  o.w = pack2(f[6], f[7], fmt);
  return o;
}
// SiLU as x * (0.5 + 0.5 * tanh(x / 2)): ONE special-function op per value (tanh.approx.f32, relative error 2^-11,
// the same size as the 16-bit rounding of the stored activation) instead of two for exp + reciprocal.  At 16 SFU
// results per clock per SM the exp form alone costs 17 us on a 96x96x128x32 tensor whose HBM time is 23 us.
__device__ __forceinline__ float silu_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}

// ---------------------------------------------------------------------------------------------- GroupNorm stats
__global__ void __launch_bounds__(256) gn_stats_kernel(const uint16_t* __restrict__ p0, const uint16_t* __restrict__ p1,
                                                       int c0, int c1, int HW, int P, float* __restrict__ partial,
                                                       int fmt) {
  extern __shared__ float sh[];  // [C] sums, [C] squares
  const int C = c0 + c1;
  const int nvec = C >> 3;
  const int lanes = blockDim.x / nvec;
  const int chunk = blockIdx.x, b = blockIdx.y, nchunks = gridDim.x;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  if (pl < lanes) {
    const int cb = v << 3;
    const uint16_t* src;
    int cs, cbs;
    if (cb < c0) {
      src = p0; cs = c0; cbs = cb;
    } else {
      src = p1; cs = c1; cbs = cb - c0;
    }
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const size_t base = static_cast<size_t>(b) * HW;
    for (int p = chunk * P + pl; p < (chunk + 1) * P; p += lanes) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (base + p) * cs + cbs));
      float f[8];
      unpack8(u, fmt, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[cb + j], s[j]);
      atomicAdd(&sh[C + cb + j], q[j]);
    }
  }
  __syncthreads();
  if (threadIdx.x < kGnGroups) {
    const int cpg = C / kGnGroups;
    float s = 0.f, q = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
      s += sh[c];
      q += sh[C + c];
    }
    float* o = partial + ((static_cast<size_t>(b) * nchunks + chunk) * kGnGroups + threadIdx.x) * 2;
    o[0] = s;
    o[1] = q;
  }
}

// ---------------------------------------------------------------------------------------------- GroupNorm apply
struct GnApplyDev {
  const uint16_t* p0;
  const uint16_t* p1;
  int c0, c1, H, W, Ho, Wo, Pout, nchunks_stats, P_stats;
  const float* partial;
  const double* stats0;
  const double* stats1;
  const float* gamma;
  const float* beta;
  const float* film;
  int film_stride, film_off, silu, mode, fmt;
  uint16_t* out;
  uint16_t* raw_out;
  int reverse;  // walk images and pixel chunks back to front (see launch_gn_apply)
  int stream_loads;  // read x with the evict-first (ld.global.cs) policy
};

// (sum, sum of squares) per image and 4-channel bucket of an NHWC tensor, accumulated into stats[B][C/4][2] (double
// atomics; the caller zeroes it).  Used for tensors whose producer is not the tcgen05 convolution (the stem).
__global__ void __launch_bounds__(256) gn_stats4_kernel(const uint16_t* __restrict__ x, int C, int HW, int P,
                                                        double* __restrict__ stats, int fmt) {
  const int nvec = C >> 3;
  const int lanes = blockDim.x / nvec;
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  if (pl >= lanes) return;
  float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
  const size_t base = static_cast<size_t>(b) * HW;
  for (int p = chunk * P + pl; p < (chunk + 1) * P; p += lanes) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (base + p) * C + v * 8));
    float f[8];
    unpack8(u, fmt, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j >> 2] += f[j];
      q[j >> 2] = fmaf(f[j], f[j], q[j >> 2]);
    }
  }
  double* o = stats + (static_cast<size_t>(b) * (C >> 2) + v * 2) * 2;
  atomicAdd(o + 0, static_cast<double>(s[0]));
  atomicAdd(o + 1, static_cast<double>(q[0]));
  atomicAdd(o + 2, static_cast<double>(s[1]));
  atomicAdd(o + 3, static_cast<double>(q[1]));
}

// kPlain: no resampling and no raw copy - the shape of 50 of the 56 launches of a forward; the specialisation keeps
// the kernel at <= 64 registers, i.e. four resident blocks per SM (the general body needs 102: two blocks, and too few
// loads in flight to cover HBM latency - measured 3.5 TB/s).
// kMode 0: plain; 1: nearest x2 upsampling (or a raw copy without resampling); 2: 2x2 average pooling.
template <int kMode>
__global__ void __launch_bounds__(256, kMode == 0 ? 4 : 3) gn_apply_kernel(const GnApplyDev a) {
  constexpr bool kPlain = kMode == 0;
  extern __shared__ float sh[];  // A[C], Bc[C], mean[32], rstd[32], FiLM scale[C], shift[C]
  const int C = a.c0 + a.c1;
  float* sA = sh;
  float* sB = sh + C;
  float* sMean = sh + 2 * C;
  float* sRstd = sMean + kGnGroups;
  const int b = a.reverse ? static_cast<int>(gridDim.y - 1 - blockIdx.y) : static_cast<int>(blockIdx.y);
  const int chunk = a.reverse ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int cpg = C / kGnGroups;
  // wait first: a pass that was itself launched programmatically must not release its own dependent (the next
  // convolution, whose CTAs take whole SMs) before it has begun to run
  pdl_wait();  // the statistics and x come from the preceding convolution
  pdl_trigger();

  // ---- pixel / channel assignment, and the FIRST batch of loads before the coefficient prologue: the prologue is two
  // dependent L2 round trips + a block barrier each (~1 us), as long as the whole data phase of a 64-pixel block
  const int nvec = C >> 3;
  const int lanes = blockDim.x / nvec;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const bool worker = pl < lanes;
  const int cb = v << 3;
  const uint16_t* src;
  int cs, cbs;
  if (cb < a.c0) {
    src = a.p0; cs = a.c0; cbs = cb;
  } else {
    src = a.p1; cs = a.c1; cbs = cb - a.c0;
  }
  const size_t in_base = static_cast<size_t>(b) * a.H * a.W;
  const size_t out_base = static_cast<size_t>(b) * a.Ho * a.Wo;
  const int HWo = a.Ho * a.Wo;
  const int p_end = min((chunk + 1) * a.Pout, HWo);
  constexpr int kU = 4;  // four pixels per thread and batch: four 16-byte requests in flight
  uint4 u[kU];
  auto issue = [&](int op0) {
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int opk = op0 + k * lanes;
      if (opk < p_end) {
        if (kPlain) {
          const uint4* lp = reinterpret_cast<const uint4*>(src + (in_base + opk) * cs + cbs);
          u[k] = a.stream_loads ? __ldcs(lp) : __ldg(lp);
        } else {
          const int oyk = opk / a.Wo, oxk = opk - oyk * a.Wo;
          const int iy = (a.mode == kResampleUp2) ? (oyk >> 1) : oyk;
          const int ix = (a.mode == kResampleUp2) ? (oxk >> 1) : oxk;
          u[k] = __ldg(reinterpret_cast<const uint4*>(src + (in_base + static_cast<size_t>(iy) * a.W + ix) * cs + cbs));
        }
      }
    }
  };
  int op = chunk * a.Pout + pl;
  if (kMode != 2 && worker) issue(op);

  // Coefficient prologue.  The group statistics are finalised in float64 by 32 threads (one per group: float64 division
  // and square root are slow enough on this part that letting every channel redo them cost 2.5 us per launch), while
  // ALL threads fetch gamma / beta / FiLM into shared memory: both sets of loads are in flight together and the second
  // phase runs out of shared memory.
  float* sSc = sRstd + kGnGroups;  // FiLM scale / shift staging: [C], [C]
  float* sSh = sSc + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sA[c] = a.gamma[c];
    sB[c] = a.beta[c];
    if (a.film != nullptr) {
      const float* f = a.film + static_cast<size_t>(b) * a.film_stride + a.film_off;
      sSc[c] = f[c];
      sSh[c] = f[C + c];
    }
  }
  if (threadIdx.x < kGnGroups) {
    double s = 0.0, q = 0.0;
    if (a.stats0 != nullptr) {
      // statistics emitted by the producers' epilogues at 4-channel granularity, one array per concat member
      const int nb0 = a.c0 >> 2;
      for (int j = threadIdx.x * (cpg >> 2); j < (threadIdx.x + 1) * (cpg >> 2); ++j) {
        const double* sp = (j < nb0) ? a.stats0 + (static_cast<size_t>(b) * nb0 + j) * 2
                                     : a.stats1 + (static_cast<size_t>(b) * (a.c1 >> 2) + (j - nb0)) * 2;
        s += sp[0];
        q += sp[1];
      }
    } else {
      const float* pp = a.partial + (static_cast<size_t>(b) * a.nchunks_stats * kGnGroups + threadIdx.x) * 2;
      for (int k = 0; k < a.nchunks_stats; ++k) {
        s += static_cast<double>(pp[static_cast<size_t>(k) * kGnGroups * 2]);
        q += static_cast<double>(pp[static_cast<size_t>(k) * kGnGroups * 2 + 1]);
      }
    }
    const double n = static_cast<double>(a.H) * a.W * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    sMean[threadIdx.x] = static_cast<float>(mean);
    sRstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // each thread rewrites only the entries it staged itself
    const int g = c / cpg;
    float A = sRstd[g] * sA[c];
    float Bc = sB[c] - sMean[g] * A;
    if (a.film != nullptr) {
      const float sc = 1.0f + sSc[c];
      A *= sc;
      Bc = Bc * sc + sSh[c];
    }
    sA[c] = A;
    sB[c] = Bc;
  }
  __syncthreads();

  if (!worker) return;
  float A[8], Bc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    A[j] = sA[cb + j];
    Bc[j] = sB[cb + j];
  }
  if (kMode != 2) {
    for (;;) {
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int opk = op + k * lanes;
        if (opk < p_end) {
          float y[8], r[8];
          unpack8(u[k], a.fmt, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float t = fmaf(r[j], A[j], Bc[j]);
            if (a.silu) t = silu_f(t);
            y[j] = t;
          }
          *reinterpret_cast<uint4*>(a.out + (out_base + opk) * C + cb) = pack8(y, a.fmt);
          if (!kPlain && a.raw_out != nullptr)
            *reinterpret_cast<uint4*>(a.raw_out + (out_base + opk) * C + cb) = pack8(r, a.fmt);
        }
      }
      op += kU * lanes;
      if (op >= p_end) break;
      issue(op);
    }
    return;
  }
  // 2x2 average pooling of the activated tensor (and of the raw input): all four source pixels of an output pixel first
  for (; op < p_end; op += lanes) {
    const int oy = op / a.Wo, ox = op - oy * a.Wo;
    uint4 q[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int iy = 2 * oy + (d >> 1), ix = 2 * ox + (d & 1);
      q[d] = __ldg(reinterpret_cast<const uint4*>(src + (in_base + static_cast<size_t>(iy) * a.W + ix) * cs + cbs));
    }
    float y[8], r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = r[j] = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float f[8];
      unpack8(q[d], a.fmt, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(f[j], A[j], Bc[j]);
        if (a.silu) t = silu_f(t);
        y[j] += t;
        r[j] += f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y[j] *= 0.25f;
      r[j] *= 0.25f;
    }
    *reinterpret_cast<uint4*>(a.out + (out_base + op) * C + cb) = pack8(y, a.fmt);
    if (a.raw_out != nullptr) *reinterpret_cast<uint4*>(a.raw_out + (out_base + op) * C + cb) = pack8(r, a.fmt);
  }
}

// ---------------------------------------------------------------------------------------------- small linears
constexpr int kLinBT = 8;
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, int in_stride,
                                                     const float* __restrict__ W, const float* __restrict__ bias,
                                                     float* __restrict__ out, int out_stride, int B, int I, int O,
                                                     int act_in, int act_out, uint16_t* __restrict__ out16,
                                                     int out16_stride, int fmt) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  for (int o = warp_global; o < O; o += total_warps) {
    const float* w = W + static_cast<size_t>(o) * I;
    for (int b0 = 0; b0 < B; b0 += kLinBT) {
      float acc[kLinBT];
#pragma unroll
      for (int j = 0; j < kLinBT; ++j) acc[j] = 0.f;
      for (int i = lane; i < I; i += 32) {
        const float wv = __ldg(w + i);
#pragma unroll
        for (int j = 0; j < kLinBT; ++j) {
          if (b0 + j < B) {
            float xv = __ldg(in + static_cast<size_t>(b0 + j) * in_stride + i);
            if (act_in) xv = xv / (1.0f + expf(-xv));
            acc[j] = fmaf(wv, xv, acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kLinBT; ++j) {
        float v = acc[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0 && b0 + j < B) {
          v += (bias != nullptr) ? bias[o] : 0.f;
          if (act_out) v = v / (1.0f + expf(-v));
          if (out != nullptr) out[static_cast<size_t>(b0 + j) * out_stride + o] = v;
          if (out16 != nullptr) {
            uint16_t bits;
            if (fmt == 1) {
              __nv_bfloat16 h = __float2bfloat16_rn(v);
              bits = *reinterpret_cast<uint16_t*>(&h);
            } else {
              __half h = __float2half_rn(v);
              bits = *reinterpret_cast<uint16_t*>(&h);
            }
            out16[static_cast<size_t>(b0 + j) * out16_stride + o] = bits;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ uint16_t f2h16(float v, int fmt) {
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ emb, int B, int dim,
                                          uint16_t* __restrict__ emb16, int fmt) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx - b * half;
  const float freq = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
  const float arg = static_cast<float>(t[b]) * freq;
  const float c = cosf(arg), s = sinf(arg);
  if (emb != nullptr) {
    emb[static_cast<size_t>(b) * dim + i] = c;
    emb[static_cast<size_t>(b) * dim + half + i] = s;
    if ((dim & 1) && i == 0) emb[static_cast<size_t>(b) * dim + dim - 1] = 0.f;
  }
  if (emb16 != nullptr) {
    emb16[static_cast<size_t>(b) * dim + i] = f2h16(c, fmt);
    emb16[static_cast<size_t>(b) * dim + half + i] = f2h16(s, fmt);
  }
}

// Wide-K variant (the encoder head 2048 -> 128 took 830 us in linear_kernel: 128 warps, each walking its whole weight
// row once per 8 batch rows with one dependent load per step).  A block owns 4 outputs x 8 batch rows; warp w owns batch
// row w, its lanes stride the K axis with float4 loads (4 in flight per operand row), so the grid is O/4 x B/8 blocks
// and the 4 weight rows of a block are shared by its 8 warps through L1.  fp32 throughout, like linear_kernel.
constexpr int kLinWO = 4;
__global__ void __launch_bounds__(256) linear_wide_kernel(const float* __restrict__ in, int in_stride,
                                                          const float* __restrict__ W, const float* __restrict__ bias,
                                                          float* __restrict__ out, int out_stride, int B, int I, int O,
                                                          int act_in, int act_out, uint16_t* __restrict__ out16,
                                                          int out16_stride, int fmt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o0 = blockIdx.x * kLinWO;
  const int b = blockIdx.y * 8 + warp;
  if (b >= B) return;
  const float4* x4 = reinterpret_cast<const float4*>(in + static_cast<size_t>(b) * in_stride);
  const float4* w4[kLinWO];
#pragma unroll
  for (int j = 0; j < kLinWO; ++j)
    w4[j] = reinterpret_cast<const float4*>(W + static_cast<size_t>(o0 + j < O ? o0 + j : O - 1) * I);
  float acc[kLinWO];
#pragma unroll
  for (int j = 0; j < kLinWO; ++j) acc[j] = 0.f;
  const int n4 = I >> 2;
#pragma unroll 4
  for (int i = lane; i < n4; i += 32) {
    float4 xv = __ldg(x4 + i);
    if (act_in) {
      xv.x = xv.x / (1.0f + expf(-xv.x));
      xv.y = xv.y / (1.0f + expf(-xv.y));
      xv.z = xv.z / (1.0f + expf(-xv.z));
      xv.w = xv.w / (1.0f + expf(-xv.w));
    }
#pragma unroll
    for (int j = 0; j < kLinWO; ++j) {
      const float4 wv = __ldg(w4[j] + i);
      acc[j] = fmaf(wv.x, xv.x, acc[j]);
      acc[j] = fmaf(wv.y, xv.y, acc[j]);
      acc[j] = fmaf(wv.z, xv.z, acc[j]);
      acc[j] = fmaf(wv.w, xv.w, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kLinWO; ++j) {
    float v = acc[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int o = o0 + j;
    if (lane == 0 && o < O) {
      v += (bias != nullptr) ? bias[o] : 0.f;
      if (act_out) v = v / (1.0f + expf(-v));
      if (out != nullptr) out[static_cast<size_t>(b) * out_stride + o] = v;
      if (out16 != nullptr) out16[static_cast<size_t>(b) * out16_stride + o] = f2h16(v, fmt);
    }
  }
}

// 16-bit small-batch linear with SiLU (the four embedding MLP layers: M = batch <= 64 rows, K <= 1024).  These ran as
// flat tcgen05 GEMMs of 4 CTAs (25 us each: barrier / TMEM / tensor-map prologue and a serial K loop) whose 200 KB of
// shared memory kept them from slipping in beside the stem kernel on the forked graph branches; one warp per output
// channel with 16-byte loads needs no shared memory and ~32 registers, and finishes in a few microseconds.
__device__ __forceinline__ float dot8_16(const uint4& a, const uint4& b, int fmt) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 fa, fb;
    if (fmt == 1) {
      fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
      fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[e]));
    } else {
      fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
      fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[e]));
    }
    acc = fmaf(fa.x, fb.x, acc);
    acc = fmaf(fa.y, fb.y, acc);
  }
  return acc;
}
constexpr int kLin16BT = 4;
__global__ void __launch_bounds__(256) linear16_kernel(const uint16_t* __restrict__ in, const uint16_t* __restrict__ w,
                                                       const float* __restrict__ bias, uint16_t* __restrict__ out,
                                                       int out_stride, int col_off, int B, int I, int O, int fmt,
                                                       int silu) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.x * 8 + warp;
  if (o >= O) return;
  const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<size_t>(o) * I);
  const int n8 = I >> 3;
  const float bv = bias != nullptr ? __ldg(bias + o) : 0.f;
  for (int b0 = 0; b0 < B; b0 += kLin16BT) {
    float acc[kLin16BT];
#pragma unroll
    for (int j = 0; j < kLin16BT; ++j) acc[j] = 0.f;
    for (int i = lane; i < n8; i += 32) {
      const uint4 wv = __ldg(wr + i);
#pragma unroll
      for (int j = 0; j < kLin16BT; ++j) {
        if (b0 + j < B) {
          const uint4 xv = *(reinterpret_cast<const uint4*>(in + static_cast<size_t>(b0 + j) * I) + i);
          acc[j] += dot8_16(wv, xv, fmt);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kLin16BT; ++j) {
      float v = acc[j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == j && b0 + j < B) {
        v += bv;
        if (silu) v = v / (1.0f + __expf(-v));
        out[static_cast<size_t>(b0 + j) * out_stride + col_off + o] = f2h16(v, fmt);
      }
    }
  }
}

__global__ void to16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int n, int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f2h16(in[i], fmt);
}

// ---------------------------------------------------------------------------------------------- stem / head convs
// Both are HBM-bound (75 MB written / read at B=32) and far too thin for the tensor cores (K = 9, N = 1); what they
// must not be is instruction-bound.  A warp owns a run of 32 consecutive pixels of one image; lane l owns channels
// [l*NPER, (l+1)*NPER) (NPER = C/32), so every access to the NHWC tensor is one contiguous 2*C-byte row per pixel and
// the lane's 9*NPER weights live in registers.  The 3x3 window slides along the run: three new loads per pixel, the
// column slots rotate at compile time (the loop over the run is fully unrolled), and the window is re-primed where
// a run wraps onto the next image row.

// Stem 1 -> C: the inputs of a pixel are warp-uniform broadcast loads.  Also emits the GroupNorm (sum, sumsq)
// buckets of the tensor it writes (on the rounded values, exactly what a separate pass over the output would see).
template <int NPER>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, uint16_t* __restrict__ out,
                                                      double* __restrict__ stats, int H, int W, int fmt) {
  constexpr int C = 32 * NPER;
  constexpr int NB = NPER >= 4 ? NPER / 4 : 1;  // 4-channel statistic buckets (partly) owned by a lane
  __shared__ float red[8][32][2 * NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, HW = H * W;
  pdl_trigger();
  float wr[9][NPER], br[NPER];
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    br[j] = __ldg(bias + lane * NPER + j);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[tap][j] = __ldg(w + (lane * NPER + j) * 9 + tap);
  }
  float sum[NB], sq[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) sum[k] = sq[k] = 0.f;
  pdl_wait();
  const float* xb = x + static_cast<size_t>(b) * HW;
  const int p0 = (blockIdx.x * 8 + warp) * 32;
  int yh = p0 / W, xw = p0 - yh * W;
  float win[3][3];  // [row dy+1][column slot]
  auto load_col = [&](int slot, int xx) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = yh + r - 1;
      win[r][slot] = (iy >= 0 && iy < H && xx >= 0 && xx < W) ? __ldg(xb + iy * W + xx) : 0.f;
    }
  };
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int p = p0 + j;
    if (p < HW) {
      // column x + d lives in slot (j + d + 1) % 3
      if (j == 0 || xw == 0) {
        load_col(j % 3, xw - 1);
        load_col((j + 1) % 3, xw);
      }
      load_col((j + 2) % 3, xw + 1);
      float acc[NPER];
#pragma unroll
      for (int c = 0; c < NPER; ++c) acc[c] = br[c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float v = win[r][(j + d) % 3];
#pragma unroll
          for (int c = 0; c < NPER; ++c) acc[c] = fmaf(v, wr[r * 3 + d][c], acc[c]);
        }
      uint32_t pk[NPER / 2];
#pragma unroll
      for (int c = 0; c < NPER; c += 2) {
        pk[c / 2] = pack2(acc[c], acc[c + 1], fmt);
        float2 rr;
        if (fmt == 1) {
          rr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[c / 2]));
        } else {
          rr = __half22float2(*reinterpret_cast<const __half2*>(&pk[c / 2]));
        }
        constexpr int kLast = NB - 1;
        const int k = c / 4 < kLast ? c / 4 : kLast;
        sum[k] += rr.x + rr.y;
        sq[k] = fmaf(rr.x, rr.x, fmaf(rr.y, rr.y, sq[k]));
      }
      uint16_t* op = out + (static_cast<size_t>(b) * HW + p) * C + lane * NPER;
      if (NPER == 2) {
        *reinterpret_cast<uint32_t*>(op) = pk[0];
      } else if (NPER == 4) {
        *reinterpret_cast<uint2*>(op) = make_uint2(pk[0], pk[1 % (NPER / 2)]);
      } else {
        *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1 % (NPER / 2)], pk[2 % (NPER / 2)], pk[3 % (NPER / 2)]);
      }
    }
    if (++xw == W) {
      xw = 0;
      ++yh;
    }
  }
  if (stats == nullptr) return;
  if (NPER == 2) {  // two lanes share a bucket
    sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], 1);
    sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 1);
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    red[warp][lane][2 * k] = sum[k];
    red[warp][lane][2 * k + 1] = sq[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 2 * NB; i += blockDim.x) {
    const int l = i / (2 * NB), kq = i - l * (2 * NB);
    if (NPER == 2 && (l & 1)) continue;
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += red[wq][l][kq];
    const int bucket = NPER == 2 ? (l >> 1) : l * NB + (kq >> 1);
    atomicAdd(&stats[(static_cast<size_t>(b) * (C / 4) + bucket) * 2 + (kq & 1)], static_cast<double>(t));
  }
}

// Stem, second generation (the kernel above needed 67 us for 75 MB of output: every lane wrote 8 bytes per pixel and the
// window loads were serialised behind each other).  A block owns 256 consecutive pixels of one image; thread =
// (pixel lane, group of 8 output channels): its 72 weights + 8 biases stay in registers, the few input rows the block
// touches are staged once in shared memory (zero-padded columns and rows), and every pixel costs 9 broadcast shared
// loads, 72 FMAs and ONE 16-byte store - a warp writes whole 256-byte pixel rows.  Same accumulation order as the
// first kernel (bias, then taps row-major), same statistics of the ROUNDED values: results are bit-identical.
template <int CG>  // channel groups of 8: C = 8 * CG
__global__ void __launch_bounds__(256, 2) conv_in2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, uint16_t* __restrict__ out,
                                                          double* __restrict__ stats, int H, int W, int fmt) {
  constexpr int C = 8 * CG;
  constexpr int kLanes = 256 / CG;   // pixels in flight per block
  constexpr int kPix = 256;          // pixels per block
  extern __shared__ float rows[];    // [nrows + 2][W + 2], zero border
  __shared__ float red[8][CG][4];
  const int b = blockIdx.y, HW = H * W;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int p_first = blockIdx.x * kPix;
  const int p_last = min(p_first + kPix, HW) - 1;
  const int y_first = p_first / W, y_last = p_last / W;
  const int nrows = y_last - y_first + 3;  // rows y_first-1 .. y_last+1
  const int pitch = W + 2;
  pdl_trigger();
  float wr[9][8], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = __ldg(bias + cg * 8 + j);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[tap][j] = __ldg(w + (cg * 8 + j) * 9 + tap);
  }
  pdl_wait();
  const float* xb = x + static_cast<size_t>(b) * HW;
  for (int i = threadIdx.x; i < nrows * pitch; i += blockDim.x) {
    const int r = i / pitch, cx = i - r * pitch;
    const int iy = y_first - 1 + r, ix = cx - 1;
    rows[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xb + iy * W + ix) : 0.f;
  }
  __syncthreads();
  float sum[2] = {0.f, 0.f}, sq[2] = {0.f, 0.f};
  for (int p = p_first + pl; p <= p_last; p += kLanes) {
    const int y = p / W, xx = p - y * W;
    const float* rp = rows + (y - y_first) * pitch + xx;  // window origin (y-1, x-1) in the padded rows
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = br[c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float v = rp[r * pitch + d];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(v, wr[r * 3 + d][c], acc[c]);
      }
    uint32_t pk[4];
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      pk[c / 2] = pack2(acc[c], acc[c + 1], fmt);
      float2 rr;
      if (fmt == 1) {
        rr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[c / 2]));
      } else {
        rr = __half22float2(*reinterpret_cast<const __half2*>(&pk[c / 2]));
      }
      sum[c / 4] += rr.x + rr.y;
      sq[c / 4] = fmaf(rr.x, rr.x, fmaf(rr.y, rr.y, sq[c / 4]));
    }
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * HW + p) * C + cg * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (stats == nullptr) return;
  // lanes of a warp that share a channel group differ in the bits above log2(CG)
  float v4[4] = {sum[0], sq[0], sum[1], sq[1]};
#pragma unroll
  for (int m = CG; m < 32; m <<= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) v4[k] += __shfl_xor_sync(0xffffffffu, v4[k], m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < CG || CG >= 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red[warp][lane % CG][k] = v4[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CG * 4; i += blockDim.x) {
    const int g = i >> 2, k = i & 3;
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += red[wq][g][k];
    // k: 0 sum / 1 sumsq of bucket 2g, 2 sum / 3 sumsq of bucket 2g + 1
    atomicAdd(&stats[(static_cast<size_t>(b) * (C / 4) + g * 2 + (k >> 1)) * 2 + (k & 1)], static_cast<double>(t));
  }
}

// Head C -> 1.  A block owns a 32 x 8 pixel tile: it first stages the 34 x 10 halo of pixel rows in shared memory
// with 16-byte cp.async copies (21 in flight per thread - the kernel was latency-bound when every warp fetched its own
// window from global memory, 171 us for 75 MB), then warp r slides the 3x3 window along tile row r out of shared
// memory: one partial sum per pixel and lane (its NPER channels x 9 taps), and a 31-shuffle butterfly over the run
// leaves lane l with the total of pixel l.
constexpr int kHeadTW = 32, kHeadTH = 8;
// kNorm: the head's GroupNorm + SiLU (out.0 / out.1, OpenAI_Unet.py:790-797) applied to the staged tile in shared memory
// - the separate gn_apply pass over the last 75 MB tensor of the forward (read + write) disappears.  Coefficients and
// rounding are those of gn_apply_kernel (fp64 finalize of the producer's 4-channel buckets, one rounding to the 16-bit
// activation type), so the result is bit-identical to the two-pass path.  Pixels outside the image stay zero: the
// convolution pads the NORMALISED tensor.
template <int NPER, bool kNorm = false>
__global__ void __launch_bounds__(256, 2) conv_out_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out,
                                                       int H, int W, int tiles_w, int fmt,
                                                       const double* __restrict__ stats = nullptr,
                                                       const float* __restrict__ gamma = nullptr,
                                                       const float* __restrict__ beta = nullptr) {
  constexpr int C = 32 * NPER;
  __shared__ float sA[kNorm ? C : 1], sB[kNorm ? C : 1], sMean[kGnGroups], sRstd[kGnGroups];
  constexpr int kRow = 2 * C;                    // bytes per pixel
  constexpr int kHW = kHeadTW + 2, kHH = kHeadTH + 2;
  constexpr int kChunks = kRow / 16;             // 16-byte chunks per pixel
  extern __shared__ __align__(16) uint8_t tile[];  // [kHH][kHW][kRow]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int ty = blockIdx.x / tiles_w, tx = blockIdx.x - ty * tiles_w;
  const int x0 = tx * kHeadTW, y0 = ty * kHeadTH;
  const uint16_t* xb = x + static_cast<size_t>(b) * H * W * C;
  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < kHH * kHW * kChunks; i += blockDim.x) {
    const int px = i / kChunks, ch = i - px * kChunks;
    const int hy = px / kHW, hx = px - hy * kHW;
    const int iy = y0 + hy - 1, ix = x0 + hx - 1;
    const uint32_t dst = tile_s + px * kRow + ch * 16;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const uint16_t* src = xb + (static_cast<size_t>(iy) * W + ix) * C + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
  float wr[9][NPER];
#pragma unroll
  for (int j = 0; j < NPER; ++j)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[tap][j] = __ldg(w + (lane * NPER + j) * 9 + tap);
  if (kNorm) {  // coefficients while the tile copies are in flight
    constexpr int cpg = C / kGnGroups;
    if (threadIdx.x < kGnGroups) {
      double s = 0.0, q = 0.0;
      for (int j = threadIdx.x * (cpg >> 2); j < (threadIdx.x + 1) * (cpg >> 2); ++j) {
        const double* sp = stats + (static_cast<size_t>(b) * (C >> 2) + j) * 2;
        s += sp[0];
        q += sp[1];
      }
      const double n = static_cast<double>(H) * W * cpg;
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      sMean[threadIdx.x] = static_cast<float>(mean);
      sRstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int g = c / cpg;
      const float A = sRstd[g] * gamma[c];
      sA[c] = A;
      sB[c] = beta[c] - sMean[g] * A;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (kNorm) {
    for (int i = threadIdx.x; i < kHH * kHW * kChunks; i += blockDim.x) {
      const int px = i / kChunks, ch = i - px * kChunks;
      const int hy = px / kHW, hx = px - hy * kHW;
      const int iy = y0 + hy - 1, ix = x0 + hx - 1;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        uint4* tp = reinterpret_cast<uint4*>(tile + px * kRow + ch * 16);
        float f[8], y[8];
        unpack8(*tp, fmt, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = silu_f(fmaf(f[j], sA[ch * 8 + j], sB[ch * 8 + j]));
        *tp = pack8(y, fmt);
      }
    }
    __syncthreads();
  }

  // warp `warp` = tile row; window column slot of halo column hx is hx % 3
  float win[3][3][NPER];
  auto load_col = [&](int hx) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const uint8_t* src = tile + ((warp + r) * kHW + hx) * kRow + lane * NPER * 2;
      uint32_t u[NPER / 2];
      if (NPER == 2) {
        u[0] = *reinterpret_cast<const uint32_t*>(src);
      } else if (NPER == 4) {
        const uint2 t = *reinterpret_cast<const uint2*>(src);
        u[0] = t.x;
        u[1 % (NPER / 2)] = t.y;
      } else {
        const uint4 t = *reinterpret_cast<const uint4*>(src);
        u[0] = t.x;
        u[1 % (NPER / 2)] = t.y;
        u[2 % (NPER / 2)] = t.z;
        u[3 % (NPER / 2)] = t.w;
      }
#pragma unroll
      for (int c = 0; c < NPER; c += 2) {
        float2 f;
        if (fmt == 1) {
          f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[c / 2]));
        } else {
          f = __half22float2(*reinterpret_cast<const __half2*>(&u[c / 2]));
        }
        win[r][hx % 3][c] = f.x;
        win[r][hx % 3][c + 1] = f.y;
      }
    }
  };
  float acc[32];
  load_col(0);
  load_col(1);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    load_col(j + 2);
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int c = 0; c < NPER; ++c) a = fmaf(win[r][(j + d) % 3][c], wr[r * 3 + d][c], a);
    acc[j] = a;
  }
#pragma unroll
  for (int wd = 16; wd >= 1; wd >>= 1) {
    const bool hi = (lane & wd) != 0;
#pragma unroll
    for (int k = 0; k < wd; ++k) {
      const float keep = hi ? acc[k + wd] : acc[k];
      const float send = hi ? acc[k] : acc[k + wd];
      acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, wd);
    }
  }
  const int oy = y0 + warp, ox = x0 + lane;
  if (oy < H && ox < W) out[(static_cast<size_t>(b) * H + oy) * W + ox] = acc[0] + __ldg(bias);
}

__global__ void vec_add_kernel(const float* a, const float* b, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

}  // namespace

// number of statistics chunks per image: smallest divisor k of HW/64 with B*k >= 256 (else the largest <= 64)
static int gn_pick_chunks(int B, int HW) {
  const int units = HW / 64;
  int best = 1;
  for (int k = 1; k <= units && k <= kGnMaxChunks; ++k) {
    if (units % k != 0) continue;
    best = k;
    if (B * k >= 256) break;
  }
  return best;
}

int gn_num_chunks(int B, int HW) { return gn_pick_chunks(B, HW); }

static int check_view(const CatView& x, const char* who) {
  if (x.p0 == nullptr || x.c0 <= 0) return fail(kInvalidArgument, std::string(who) + ": missing source");
  if (x.c0 % 8 != 0 || x.c1 % 8 != 0) return fail(kUnsupported, std::string(who) + ": channel counts must be multiples of 8");
  const int C = x.c0 + x.c1;
  if (C % kGnGroups != 0) return fail(kUnsupported, std::string(who) + ": channels must be a multiple of 32");
  if (C / 8 > 256) return fail(kUnsupported, std::string(who) + ": at most 2048 channels");
  if (x.c1 > 0 && x.p1 == nullptr) return fail(kInvalidArgument, std::string(who) + ": missing second source");
  return kOk;
}

int launch_gn_stats(const CatView& x, int B, int HW, float* partial, int fmt, cudaStream_t stream) {
  CDDPM_TRY(check_view(x, "gn_stats"));
  if (HW % 64 != 0) return fail(kUnsupported, "gn_stats: H*W must be a multiple of 64");
  const int C = x.C();
  const int k = gn_pick_chunks(B, HW);
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  dim3 grid(k, B);
  gn_stats_kernel<<<grid, threads, 2 * C * sizeof(float), stream>>>(
      reinterpret_cast<const uint16_t*>(x.p0), reinterpret_cast<const uint16_t*>(x.p1), x.c0, x.c1, HW, HW / k,
      partial, fmt);
  return check_launch("gn_stats_kernel");
}

int launch_gn_stats4(const void* x, int C, int B, int HW, double* stats, int fmt, cudaStream_t stream) {
  if (!x || !stats) return fail(kInvalidArgument, "gn_stats4: null pointer");
  if (C % 8 != 0 || C / 8 > 256 || HW % 64 != 0) return fail(kUnsupported, "gn_stats4: unsupported shape");
  const int k = gn_pick_chunks(B, HW);
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  dim3 grid(k, B);
  gn_stats4_kernel<<<grid, threads, 0, stream>>>(reinterpret_cast<const uint16_t*>(x), C, HW, HW / k, stats, fmt);
  return check_launch("gn_stats4_kernel");
}

int launch_gn_apply(const GnApplyArgs& a, cudaStream_t stream) {
  CDDPM_TRY(check_view(a.x, "gn_apply"));
  if ((!a.partial && !a.stats0) || !a.gamma || !a.beta || !a.out) return fail(kInvalidArgument, "gn_apply: null pointer");
  if (a.stats0 && a.x.c1 > 0 && !a.stats1) return fail(kInvalidArgument, "gn_apply: missing statistics of the second source");
  if (a.stats0 && ((a.x.C() / kGnGroups) % 4 != 0 || a.x.c0 % 4 != 0))
    return fail(kUnsupported, "gn_apply: 4-channel statistics need group size and concat split to be multiples of 4");
  const int HW = a.H * a.W;
  if (HW % 64 != 0) return fail(kUnsupported, "gn_apply: H*W must be a multiple of 64");
  if (a.mode == kResampleDown2 && ((a.H | a.W) & 1)) return fail(kInvalidArgument, "gn_apply: odd size for avg-pool");
  GnApplyDev d;
  d.p0 = reinterpret_cast<const uint16_t*>(a.x.p0);
  d.p1 = reinterpret_cast<const uint16_t*>(a.x.p1);
  d.c0 = a.x.c0;
  d.c1 = a.x.c1;
  d.H = a.H;
  d.W = a.W;
  d.Ho = a.mode == kResampleUp2 ? a.H * 2 : (a.mode == kResampleDown2 ? a.H / 2 : a.H);
  d.Wo = a.mode == kResampleUp2 ? a.W * 2 : (a.mode == kResampleDown2 ? a.W / 2 : a.W);
  d.nchunks_stats = gn_pick_chunks(a.B, HW);
  d.P_stats = HW / d.nchunks_stats;
  d.partial = a.partial;
  d.stats0 = a.stats0;
  d.stats1 = a.stats1;
  d.gamma = a.gamma;
  d.beta = a.beta;
  d.film = a.film;
  d.film_stride = a.film_stride;
  d.film_off = a.film_off;
  d.silu = a.silu;
  d.mode = a.mode;
  d.fmt = a.fmt;
  d.out = reinterpret_cast<uint16_t*>(a.out);
  d.raw_out = reinterpret_cast<uint16_t*>(a.raw_out);
  // Back-to-front traversal: the producing convolution wrote image B-1 last, so at B=32 (75-151 MB per tensor against
  // 126 MB of L2) its tail is what is still cached; reading it first, and thereby writing image 0 LAST - the image the
  // consuming convolution reads first - turns part of both HBM passes into L2 hits.  CDDPM_GN_REVERSE=0/1.
  static const int reverse = [] {
    const char* e = getenv("CDDPM_GN_REVERSE");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  d.reverse = reverse;
  static const int hints = [] {
    const char* e = getenv("CDDPM_L2_HINTS");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  d.stream_loads = hints;
  const int C = a.x.C();
  const int HWo = d.Ho * d.Wo;
  // Output pixels per block: 256 when that already fills the GPU (4 x 148 blocks), else 64 - and for small batches,
  // where even that leaves SMs without a block (B = 1 at 24 x 24: 9 blocks), down to 16 until there are >= 256 blocks.
  // Measured (profiles/r02_c5_ab.log): B = 1 forward 1.247 -> 1.161 ms; at B = 32 finer blocks only add coefficient
  // prologues (5.385 -> 5.424 ms), so large batches keep the round-1 rule.
  d.Pout = (static_cast<long long>(HWo / 256) * a.B >= 592) ? 256 : 64;
  static const int min_blocks = [] {
    const char* e = getenv("CDDPM_GN_MIN_BLOCKS");  // A/B switch for measurements
    return (e != nullptr && atoi(e) > 0) ? atoi(e) : 256;
  }();
  while (d.Pout > 16 && static_cast<long long>((HWo + d.Pout - 1) / d.Pout) * a.B < min_blocks) d.Pout >>= 1;
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  dim3 grid((HWo + d.Pout - 1) / d.Pout, a.B);
  const size_t shmem = (4 * C + 2 * kGnGroups) * sizeof(float);
  cudaError_t e;
  if (a.mode == kResampleNone && a.raw_out == nullptr) {
    e = launch_k(gn_apply_kernel<0>, grid, dim3(threads), shmem, stream, d);
  } else if (a.mode == kResampleDown2) {
    e = launch_k(gn_apply_kernel<2>, grid, dim3(threads), shmem, stream, d);
  } else {
    e = launch_k(gn_apply_kernel<1>, grid, dim3(threads), shmem, stream, d);
  }
  return check_cuda(e, "gn_apply_kernel");
}

int launch_linear_ex(const float* in, int in_stride, const float* W, const float* bias, float* out, int out_stride,
                     int B, int I, int O, int act_in, int act_out, cudaStream_t stream) {
  return launch_linear_16(in, in_stride, W, bias, out, out_stride, B, I, O, act_in, act_out, nullptr, 0, 0, stream);
}

int launch_linear_16(const float* in, int in_stride, const float* W, const float* bias, float* out, int out_stride,
                     int B, int I, int O, int act_in, int act_out, void* out16, int out16_stride, int fmt,
                     cudaStream_t stream) {
  if (!in || !W || (!out && !out16)) return fail(kInvalidArgument, "linear: null pointer");
  if (I >= 256 && I % 4 == 0 && in_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15) == 0 && B >= 1 && O >= 1) {
    dim3 grid((O + kLinWO - 1) / kLinWO, (B + 7) / 8);
    linear_wide_kernel<<<grid, 256, 0, stream>>>(in, in_stride, W, bias, out, out_stride, B, I, O, act_in, act_out,
                                                 reinterpret_cast<uint16_t*>(out16), out16_stride, fmt);
    return check_launch("linear_wide_kernel");
  }
  int blocks = (O + 7) / 8;
  if (blocks > 8192) blocks = 8192;
  if (blocks < 1) blocks = 1;
  linear_kernel<<<blocks, 256, 0, stream>>>(in, in_stride, W, bias, out, out_stride, B, I, O, act_in, act_out,
                                            reinterpret_cast<uint16_t*>(out16), out16_stride, fmt);
  return check_launch("linear_kernel");
}

int launch_linear(const float* in, int in_stride, const float* W, const float* bias, float* out, int out_stride,
                  int B, int I, int O, int act_in, cudaStream_t stream) {
  return launch_linear_ex(in, in_stride, W, bias, out, out_stride, B, I, O, act_in, 0, stream);
}

int launch_timestep_embedding(const int64_t* t, float* emb, int B, int dim, cudaStream_t stream) {
  return launch_timestep_embedding16(t, emb, nullptr, 0, B, dim, stream);
}

int launch_timestep_embedding16(const int64_t* t, float* emb, void* emb16, int fmt, int B, int dim, cudaStream_t stream) {
  const int n = B * (dim / 2);
  timestep_embedding_kernel<<<(n + 255) / 256, 256, 0, stream>>>(t, emb, B, dim, reinterpret_cast<uint16_t*>(emb16), fmt);
  return check_launch("timestep_embedding_kernel");
}

int launch_linear16(const void* in, const void* w16, const float* bias, void* out, int out_stride, int col_off, int B,
                    int I, int O, int fmt, int silu, cudaStream_t stream) {
  if (!in || !w16 || !out) return fail(kInvalidArgument, "linear16: null pointer");
  if (I % 8 != 0 || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w16)) & 15) != 0)
    return fail(kInvalidArgument, "linear16: K must be a multiple of 8 and the operands 16-byte aligned");
  linear16_kernel<<<(O + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(in),
                                                   reinterpret_cast<const uint16_t*>(w16), bias,
                                                   reinterpret_cast<uint16_t*>(out), out_stride, col_off, B, I, O, fmt, silu);
  return check_launch("linear16_kernel");
}

int launch_to16(const float* in, void* out, int n, int fmt, cudaStream_t stream) {
  if (!in || !out) return fail(kInvalidArgument, "to16: null pointer");
  to16_kernel<<<(n + 255) / 256, 256, 0, stream>>>(in, reinterpret_cast<uint16_t*>(out), n, fmt);
  return check_launch("to16_kernel");
}

int launch_conv_in(const float* x, const float* w, const float* bias, void* out, double* stats, int B, int H, int W,
                   int Cout, int fmt, cudaStream_t stream) {
  dim3 grid((H * W + 255) / 256, B);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  static const bool v2 = [] {
    const char* e = getenv("CDDPM_STEM_V2");  // A/B switch: 0 = first-generation stem kernel
    return !(e != nullptr && e[0] == '0');
  }();
  if (v2 && W >= 8) {
    // rows touched by 256 consecutive pixels: at most 256 / W + 2, plus the two halo rows
    const int max_rows = 256 / W + 2 + 2;
    const size_t smem = static_cast<size_t>(max_rows) * (W + 2) * sizeof(float);
    if (smem <= 40 * 1024) {
      switch (Cout) {
        case 64: return check_cuda(launch_k(conv_in2_kernel<8>, grid, dim3(256), smem, stream, x, w, bias, o, stats, H, W, fmt), "conv_in2_kernel");
        case 128: return check_cuda(launch_k(conv_in2_kernel<16>, grid, dim3(256), smem, stream, x, w, bias, o, stats, H, W, fmt), "conv_in2_kernel");
        case 256: return check_cuda(launch_k(conv_in2_kernel<32>, grid, dim3(256), smem, stream, x, w, bias, o, stats, H, W, fmt), "conv_in2_kernel");
        default: return fail(kUnsupported, "conv_in: model_channels must be 64, 128 or 256");
      }
    }
  }
  switch (Cout) {
    case 64: return check_cuda(launch_k(conv_in_kernel<2>, grid, dim3(256), 0, stream, x, w, bias, o, stats, H, W, fmt), "conv_in_kernel");
    case 128: return check_cuda(launch_k(conv_in_kernel<4>, grid, dim3(256), 0, stream, x, w, bias, o, stats, H, W, fmt), "conv_in_kernel");
    case 256: return check_cuda(launch_k(conv_in_kernel<8>, grid, dim3(256), 0, stream, x, w, bias, o, stats, H, W, fmt), "conv_in_kernel");
    default: return fail(kUnsupported, "conv_in: model_channels must be 64, 128 or 256");
  }
}

int launch_conv_out_gn(const void* x, const double* stats, const float* gamma, const float* beta, const float* w,
                       const float* bias, float* out, int B, int H, int W, int C, int fmt, cudaStream_t stream) {
  if (!x || !stats || !gamma || !beta || !w || !bias || !out) return fail(kInvalidArgument, "conv_out_gn: null pointer");
  const int tiles_w = (W + kHeadTW - 1) / kHeadTW, tiles_h = (H + kHeadTH - 1) / kHeadTH;
  dim3 grid(tiles_w * tiles_h, B);
  const int smem = (kHeadTW + 2) * (kHeadTH + 2) * 2 * C;
  const uint16_t* xi = reinterpret_cast<const uint16_t*>(x);
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(conv_out_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 340 * 256));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_out_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 340 * 512));
    attr_set = true;
  }
  switch (C) {  // 4-channel statistic buckets need C / 32 to be a multiple of 4
    case 128: return check_cuda(launch_k(conv_out_kernel<4, true>, grid, dim3(256), smem, stream, xi, w, bias, out, H, W, tiles_w, fmt, stats, gamma, beta), "conv_out_kernel");
    case 256: return check_cuda(launch_k(conv_out_kernel<8, true>, grid, dim3(256), smem, stream, xi, w, bias, out, H, W, tiles_w, fmt, stats, gamma, beta), "conv_out_kernel");
    default: return fail(kUnsupported, "conv_out_gn: model_channels must be 128 or 256");
  }
}

int launch_conv_out(const void* x, const float* w, const float* bias, float* out, int B, int H, int W, int C,
                    int fmt, cudaStream_t stream) {
  const int tiles_w = (W + kHeadTW - 1) / kHeadTW, tiles_h = (H + kHeadTH - 1) / kHeadTH;
  dim3 grid(tiles_w * tiles_h, B);
  const int smem = (kHeadTW + 2) * (kHeadTH + 2) * 2 * C;
  const uint16_t* xi = reinterpret_cast<const uint16_t*>(x);
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(conv_out_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 340 * 128));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_out_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 340 * 256));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_out_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 340 * 512));
    attr_set = true;
  }
  switch (C) {
    case 64: return check_cuda(launch_k(conv_out_kernel<2, false>, grid, dim3(256), smem, stream, xi, w, bias, out, H, W, tiles_w, fmt, static_cast<const double*>(nullptr), static_cast<const float*>(nullptr), static_cast<const float*>(nullptr)), "conv_out_kernel");
    case 128: return check_cuda(launch_k(conv_out_kernel<4, false>, grid, dim3(256), smem, stream, xi, w, bias, out, H, W, tiles_w, fmt, static_cast<const double*>(nullptr), static_cast<const float*>(nullptr), static_cast<const float*>(nullptr)), "conv_out_kernel");
    case 256: return check_cuda(launch_k(conv_out_kernel<8, false>, grid, dim3(256), smem, stream, xi, w, bias, out, H, W, tiles_w, fmt, static_cast<const double*>(nullptr), static_cast<const float*>(nullptr), static_cast<const float*>(nullptr)), "conv_out_kernel");
    default: return fail(kUnsupported, "conv_out: model_channels must be 64, 128 or 256");
  }
}

int launch_vec_add(const float* a, const float* b, float* out, int n, cudaStream_t stream) {
  if (job_recorder() != nullptr) {
    ParamJob j = {kJobVecAdd, 0, 0, 0, 0, 0, 0, 0, 0, a, b, out, static_cast<long long>(n)};
    job_record(j);
    return kOk;
  }
  vec_add_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, out, n);
  return check_launch("vec_add_kernel");
}

}  // namespace cddpm
