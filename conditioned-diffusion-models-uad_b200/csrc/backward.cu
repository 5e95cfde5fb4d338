// Bandwidth-bound kernels of the UNet backward pass (see backward.cuh).  Same conventions as elementwise.cu:
// 16-byte vector accesses along the channel dimension, fp32 math, one rounding to the 16-bit type on store.
#include "backward.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

#include "ptx.cuh"

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
__device__ __forceinline__ float sigmoid_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
// d/dz (z * sigmoid(z))
__device__ __forceinline__ float silu_grad(float z) {
  const float s = sigmoid_f(z);
  return s * fmaf(z, 1.0f - s, 1.0f);
}
__device__ __forceinline__ float silu_precise(float z) { return z / (1.0f + __expf(-z)); }
__device__ __forceinline__ float silu_grad_precise(float z) {
  const float s = 1.0f / (1.0f + __expf(-z));
  return s * (1.0f + z * (1.0f - s));
}

struct GnBwdDev {
  const uint16_t* p0;
  const uint16_t* p1;
  int c0, c1;
  int B, HW;
  const double* stats0;
  const double* stats1;
  const float* gamma;
  const float* beta;
  const float* film;
  int film_stride, film_off;
  int silu;
  const uint16_t* dy;
  const uint16_t* add0;
  const uint16_t* add1;
  float* sums;
  uint16_t* out0;
  uint16_t* out1;
  float* bsum0;
  float* bsum1;
  float* dgamma;
  float* dbeta;
  float* dfilm;
  int fmt;
  int P;  // pixels per block
};

constexpr int kGnU = 4;   // pixels (16-byte loads per tensor) in flight per thread: apply (3-4 tensors)
constexpr int kGnUr = 6;  // ... reduce (2 tensors)

// Shared prologue.  The forward computed z = gp * xhat + bp with xhat = x * rs - mr (rs = rstd of the channel's group,
// mr = mean * rstd), gp = gamma * (1 + scale), bp = beta * (1 + scale) + shift.  Per channel we keep (gp, bp); the
// group constants are shared by each aligned run of four channels (the group size is a multiple of four), which keeps
// the kernels' coefficient registers at 16 + 2 * (2 or 4) per thread instead of 32 / 48.
// sh layout: gp[C], bp[C]; sG: rs[32], mr[32]
__device__ __forceinline__ void gn_bwd_prologue(const GnBwdDev& a, int b, float* sh, float* sRs, float* sMr) {
  const int C = a.c0 + a.c1;
  const int cpg = C / kGnGroups;
  if (threadIdx.x < kGnGroups) {
    double s = 0.0, q = 0.0;
    const int nb0 = a.c0 >> 2;
    for (int j = threadIdx.x * (cpg >> 2); j < (threadIdx.x + 1) * (cpg >> 2); ++j) {
      const double* sp = (j < nb0) ? a.stats0 + (static_cast<size_t>(b) * nb0 + j) * 2
                                   : a.stats1 + (static_cast<size_t>(b) * (a.c1 >> 2) + (j - nb0)) * 2;
      s += sp[0];
      q += sp[1];
    }
    const double n = static_cast<double>(a.HW) * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + 1e-5);
    sRs[threadIdx.x] = static_cast<float>(rstd);
    sMr[threadIdx.x] = static_cast<float>(mean * rstd);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float gp = a.gamma[c];
    float bp = a.beta[c];
    if (a.film != nullptr) {
      const float* f = a.film + static_cast<size_t>(b) * a.film_stride + a.film_off;
      const float sc = 1.0f + f[c];
      gp *= sc;
      bp = bp * sc + f[C + c];
    }
    sh[c] = gp;
    sh[C + c] = bp;
  }
  __syncthreads();
}

// sums[b][c] = (sum_p g, sum_p g * xhat), g = dy * act'(z)
__global__ void __launch_bounds__(256, 2) gn_bwd_reduce_kernel(const GnBwdDev a) {
  extern __shared__ float sh[];
  __shared__ float sRs[kGnGroups], sMr[kGnGroups];
  const int C = a.c0 + a.c1;
  const int cpg = C / kGnGroups;
  const int b = blockIdx.y;
  gn_bwd_prologue(a, b, sh, sRs, sMr);
  const int nvec = C >> 3;
  const int lanes = blockDim.x / nvec;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int cb = v << 3;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (pl < lanes) {
    const uint16_t* src;
    int cs, cbs;
    if (cb < a.c0) {
      src = a.p0; cs = a.c0; cbs = cb;
    } else {
      src = a.p1; cs = a.c1; cbs = cb - a.c0;
    }
    float gp[8], bp[8], rs[2], mr[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gp[j] = sh[cb + j];
      bp[j] = sh[C + cb + j];
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      rs[hf] = sRs[(cb + 4 * hf) / cpg];
      mr[hf] = sMr[(cb + 4 * hf) / cpg];
    }
    const size_t base = static_cast<size_t>(b) * a.HW;
    const int p_end = min((static_cast<int>(blockIdx.x) + 1) * a.P, a.HW);
    for (int p = blockIdx.x * a.P + pl; p < p_end; p += kGnUr * lanes) {
      uint4 ux[kGnUr], ud[kGnUr];
#pragma unroll
      for (int k = 0; k < kGnUr; ++k) {
        const int pk = p + k * lanes;
        if (pk < p_end) {
          ux[k] = __ldg(reinterpret_cast<const uint4*>(src + (base + pk) * cs + cbs));
          ud[k] = __ldg(reinterpret_cast<const uint4*>(a.dy + (base + pk) * C + cb));
        }
      }
#pragma unroll
      for (int k = 0; k < kGnUr; ++k) {
        if (p + k * lanes < p_end) {
          float x[8], d[8];
          unpack8(ux[k], a.fmt, x);
          unpack8(ud[k], a.fmt, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(x[j], rs[j >> 2], -mr[j >> 2]);
            float g = d[j];
            if (a.silu) g *= silu_grad(fmaf(gp[j], xh, bp[j]));
            s1[j] += g;
            s2[j] = fmaf(g, xh, s2[j]);
          }
        }
      }
    }
  }
  __syncthreads();  // the prologue's coefficients are dead: reuse the buffer for the block reduction
  float* red = sh;  // [lanes][C][2]
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[(pl * C + cb + j) * 2] = s1[j];
      red[(pl * C + cb + j) * 2 + 1] = s2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * 2 * C + i];
    atomicAdd(&a.sums[static_cast<size_t>(b) * 2 * C + i], t);
  }
}

// dx = rstd * (gamma' g - (P1 + xhat P2) / N) + add0 + add1
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(const GnBwdDev a) {
  extern __shared__ float sh[];  // gp, bp (per channel), later the column-sum scratch
  __shared__ float sRs[kGnGroups], sMr[kGnGroups], sK1[kGnGroups], sK2[kGnGroups];
  const int C = a.c0 + a.c1;
  const int cpg = C / kGnGroups;
  const int b = blockIdx.y;
  gn_bwd_prologue(a, b, sh, sRs, sMr);
  const float* S = a.sums + static_cast<size_t>(b) * 2 * C;
  if (threadIdx.x < kGnGroups) {
    // group sums of gamma' * (S1, S2)
    float p1 = 0.f, p2 = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
      p1 = fmaf(sh[c], S[2 * c], p1);
      p2 = fmaf(sh[c], S[2 * c + 1], p2);
    }
    const float inv_n = 1.0f / (static_cast<float>(a.HW) * cpg);
    sK1[threadIdx.x] = sRs[threadIdx.x] * p1 * inv_n;
    sK2[threadIdx.x] = sRs[threadIdx.x] * p2 * inv_n;
  }
  if (blockIdx.x == 0) {
    // parameter gradients: one block per image owns them
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float s1 = S[2 * c], s2 = S[2 * c + 1];
      float sc = 1.0f;
      if (a.film != nullptr) {
        sc = 1.0f + a.film[static_cast<size_t>(b) * a.film_stride + a.film_off + c];
        if (a.dfilm != nullptr) {
          float* df = a.dfilm + static_cast<size_t>(b) * a.film_stride + a.film_off;
          df[c] = fmaf(a.gamma[c], s2, a.beta[c] * s1);  // d scale = sum g * (gamma xhat + beta)
          df[C + c] = s1;                                 // d shift
        }
      }
      atomicAdd(&a.dgamma[c], sc * s2);
      atomicAdd(&a.dbeta[c], sc * s1);
    }
  }
  __syncthreads();

  const int nvec = C >> 3;
  const int lanes = blockDim.x / nvec;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int cb = v << 3;
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = 0.f;
  const bool first = cb < a.c0;
  if (pl < lanes) {
    const uint16_t* src = first ? a.p0 : a.p1;
    const int cs = first ? a.c0 : a.c1;
    const int cbs = first ? cb : cb - a.c0;
    uint16_t* dst = first ? a.out0 : a.out1;
    float gp[8], bp[8], rs[2], mr[2], k1[2], k2[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gp[j] = sh[cb + j];
      bp[j] = sh[C + cb + j];
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int g = (cb + 4 * hf) / cpg;
      rs[hf] = sRs[g];
      mr[hf] = sMr[g];
      k1[hf] = sK1[g];
      k2[hf] = sK2[g];
    }
    const size_t base = static_cast<size_t>(b) * a.HW;
    const int p_end = min((static_cast<int>(blockIdx.x) + 1) * a.P, a.HW);
    for (int p = blockIdx.x * a.P + pl; p < p_end; p += kGnU * lanes) {
      uint4 ux[kGnU], ud[kGnU], u0[kGnU], u1[kGnU];
#pragma unroll
      for (int k = 0; k < kGnU; ++k) {
        const int pk = p + k * lanes;
        if (pk < p_end) {
          ux[k] = __ldg(reinterpret_cast<const uint4*>(src + (base + pk) * cs + cbs));
          ud[k] = __ldg(reinterpret_cast<const uint4*>(a.dy + (base + pk) * C + cb));
          if (a.add0 != nullptr) u0[k] = __ldg(reinterpret_cast<const uint4*>(a.add0 + (base + pk) * C + cb));
          if (a.add1 != nullptr) u1[k] = __ldg(reinterpret_cast<const uint4*>(a.add1 + (base + pk) * C + cb));
        }
      }
#pragma unroll
      for (int k = 0; k < kGnU; ++k) {
        const int pk = p + k * lanes;
        if (pk < p_end) {
          float x[8], d[8], o[8];
          unpack8(ux[k], a.fmt, x);
          unpack8(ud[k], a.fmt, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(x[j], rs[j >> 2], -mr[j >> 2]);
            float g = d[j];
            if (a.silu) g *= silu_grad(fmaf(gp[j], xh, bp[j]));
            o[j] = fmaf(g * gp[j], rs[j >> 2], -k1[j >> 2]) - xh * k2[j >> 2];
          }
          if (a.add0 != nullptr) {
            float t[8];
            unpack8(u0[k], a.fmt, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += t[j];
          }
          if (a.add1 != nullptr) {
            float t[8];
            unpack8(u1[k], a.fmt, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += t[j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) bs[j] += o[j];
          *reinterpret_cast<uint4*>(dst + (base + pk) * cs + cbs) = pack8(o, a.fmt);
        }
      }
    }
  }
  if (a.bsum0 == nullptr && a.bsum1 == nullptr) return;
  __syncthreads();
  float* red = sh;  // [lanes][C]
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl * C + cb + j] = bs[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * C + c];
    if (c < a.c0) {
      if (a.bsum0 != nullptr) atomicAdd(&a.bsum0[c], t);
    } else {
      if (a.bsum1 != nullptr) atomicAdd(&a.bsum1[c - a.c0], t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Ring versions of the two GroupNorm backward kernels.  The register-fed kernels above are latency-bound (ncu: issue
// slots 45 % busy, DRAM 44 %): a warp issues no loads while it chews through the SiLU' / normalisation math of the
// pixels it holds.  Here one producer thread streams whole pixel tiles (x | dy | skip-path gradients: contiguous NHWC
// row ranges) into a shared-memory ring with cp.async.bulk + mbarriers, so the bytes in flight (6 stages x 16-32 KB per
// SM) no longer depend on registers or on what the 16 consumer warps are doing.
// ------------------------------------------------------------------------------------------------------------
constexpr int kRingConsumers = 512;
constexpr int kRingThreads = kRingConsumers + 32;
constexpr int kRingStages = 6;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc), "r"(bytes), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
               : "memory");
}

template <bool kApply>
__global__ void __launch_bounds__(kRingThreads, 1) gn_bwd_ring_kernel(const GnBwdDev a, int TP, int stage_bytes) {
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ float sRs[kGnGroups], sMr[kGnGroups], sK1[kGnGroups], sK2[kGnGroups];
  __shared__ __align__(8) uint64_t full[kRingStages], empty[kRingStages];
  const int C = a.c0 + a.c1;
  const int cpg = C / kGnGroups;
  const int b = blockIdx.y;
  uint8_t* ring = ring_raw;
  float* sh = reinterpret_cast<float*>(ring_raw + static_cast<size_t>(kRingStages) * stage_bytes);  // gp[C], bp[C]
  const int p_begin = blockIdx.x * a.P;
  const int p_end = min(p_begin + a.P, a.HW);
  const int ntiles = (p_end - p_begin + TP - 1) / TP;
  const size_t base = static_cast<size_t>(b) * a.HW;
  const int off_x1 = TP * a.c0 * 2;
  const int off_dy = TP * C * 2;
  const int off_a0 = 2 * TP * C * 2;
  const int off_a1 = off_a0 + (a.add0 != nullptr ? TP * C * 2 : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == kRingConsumers / 32 && lane == 0;
  auto issue_tile = [&](int t, int st) {
    const int ps = p_begin + t * TP;
    const int np = min(TP, p_end - ps);
    uint32_t bytes = static_cast<uint32_t>(np) * C * 2 * 2;
    if (kApply && a.add0 != nullptr) bytes += static_cast<uint32_t>(np) * C * 2;
    if (kApply && a.add1 != nullptr) bytes += static_cast<uint32_t>(np) * C * 2;
    mbar_arrive_expect_tx(&full[st], bytes);
    uint8_t* dst = ring + static_cast<size_t>(st) * stage_bytes;
    bulk_load(dst, a.p0 + (base + ps) * a.c0, static_cast<uint32_t>(np) * a.c0 * 2, &full[st]);
    if (a.c1 > 0) bulk_load(dst + off_x1, a.p1 + (base + ps) * a.c1, static_cast<uint32_t>(np) * a.c1 * 2, &full[st]);
    bulk_load(dst + off_dy, a.dy + (base + ps) * C, static_cast<uint32_t>(np) * C * 2, &full[st]);
    if (kApply && a.add0 != nullptr)
      bulk_load(dst + off_a0, a.add0 + (base + ps) * C, static_cast<uint32_t>(np) * C * 2, &full[st]);
    if (kApply && a.add1 != nullptr)
      bulk_load(dst + off_a1, a.add1 + (base + ps) * C, static_cast<uint32_t>(np) * C * 2, &full[st]);
  };
  if (producer) {
    for (int i = 0; i < kRingStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kRingConsumers / 32);
    }
    fence_mbar_init();
    // the first ring-full of tiles does not depend on the prologue: start it now
    for (int t = 0; t < ntiles && t < kRingStages; ++t) issue_tile(t, t);
  }
  gn_bwd_prologue(a, b, sh, sRs, sMr);
  const float* S = a.sums + static_cast<size_t>(b) * 2 * C;
  if (kApply) {
    if (threadIdx.x < kGnGroups) {
      float p1 = 0.f, p2 = 0.f;
      for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
        p1 = fmaf(sh[c], S[2 * c], p1);
        p2 = fmaf(sh[c], S[2 * c + 1], p2);
      }
      const float inv_n = 1.0f / (static_cast<float>(a.HW) * cpg);
      sK1[threadIdx.x] = sRs[threadIdx.x] * p1 * inv_n;
      sK2[threadIdx.x] = sRs[threadIdx.x] * p2 * inv_n;
    }
    if (blockIdx.x == 0) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float s1 = S[2 * c], s2 = S[2 * c + 1];
        float sc = 1.0f;
        if (a.film != nullptr) {
          sc = 1.0f + a.film[static_cast<size_t>(b) * a.film_stride + a.film_off + c];
          if (a.dfilm != nullptr) {
            float* df = a.dfilm + static_cast<size_t>(b) * a.film_stride + a.film_off;
            df[c] = fmaf(a.gamma[c], s2, a.beta[c] * s1);
            df[C + c] = s1;
          }
        }
        atomicAdd(&a.dgamma[c], sc * s2);
        atomicAdd(&a.dbeta[c], sc * s1);
      }
    }
    __syncthreads();
  }

  const int nvec = C >> 3;
  const int lanes = kRingConsumers / nvec;

  if (warp == kRingConsumers / 32) {
    // ------------------------------------------------------------------ producer: the rest of the tiles
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 1;  // second pass over the ring
      for (int t = kRingStages; t < ntiles; ++t) {
        mbar_wait(&empty[st], ph ^ 1);
        issue_tile(t, st);
        if (++st == kRingStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int cb = v << 3;
  const bool active = pl < lanes;
  const bool first = cb < a.c0;
  const int cs = first ? a.c0 : a.c1;
  const int cbs = first ? cb : cb - a.c0;
  const int x_off = (first ? 0 : off_x1) + (pl * cs + cbs) * 2;
  const int d_off = (pl * C + cb) * 2;
  float gp[8], bp[8], rs[2], mr[2], k1[2], k2[2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gp[j] = sh[cb + j];
    bp[j] = sh[C + cb + j];
  }
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int g = (cb + 4 * hf) / cpg;
    rs[hf] = sRs[g];
    mr[hf] = sMr[g];
    k1[hf] = kApply ? sK1[g] : 0.f;
    k2[hf] = kApply ? sK2[g] : 0.f;
  }
  float acc0[8], acc1[8];  // reduce: sum g, sum g * xhat; apply: column sums of dx (acc0 only)
#pragma unroll
  for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
  uint16_t* dst = first ? a.out0 : a.out1;
  int st = 0;
  uint32_t ph = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int pix = p_begin + t * TP + pl;
    const bool valid = active && pix < p_end;
    mbar_wait(&full[st], ph);
    const uint8_t* sp = ring + static_cast<size_t>(st) * stage_bytes;
    uint4 ux, ud, u0, u1;
    if (valid) {
      ux = *reinterpret_cast<const uint4*>(sp + x_off);
      ud = *reinterpret_cast<const uint4*>(sp + off_dy + d_off);
      if (kApply && a.add0 != nullptr) u0 = *reinterpret_cast<const uint4*>(sp + off_a0 + d_off);
      if (kApply && a.add1 != nullptr) u1 = *reinterpret_cast<const uint4*>(sp + off_a1 + d_off);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // the warp holds its part of the stage in registers
    if (++st == kRingStages) {
      st = 0;
      ph ^= 1;
    }
    if (!valid) continue;
    float x[8], d[8];
    unpack8(ux, a.fmt, x);
    unpack8(ud, a.fmt, d);
    if (!kApply) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(x[j], rs[j >> 2], -mr[j >> 2]);
        float g = d[j];
        if (a.silu) g *= silu_grad(fmaf(gp[j], xh, bp[j]));
        acc0[j] += g;
        acc1[j] = fmaf(g, xh, acc1[j]);
      }
    } else {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(x[j], rs[j >> 2], -mr[j >> 2]);
        float g = d[j];
        if (a.silu) g *= silu_grad(fmaf(gp[j], xh, bp[j]));
        o[j] = fmaf(g * gp[j], rs[j >> 2], -k1[j >> 2]) - xh * k2[j >> 2];
      }
      if (a.add0 != nullptr) {
        float tt[8];
        unpack8(u0, a.fmt, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += tt[j];
      }
      if (a.add1 != nullptr) {
        float tt[8];
        unpack8(u1, a.fmt, tt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += tt[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc0[j] += o[j];
      *reinterpret_cast<uint4*>(dst + (base + pix) * cs + cbs) = pack8(o, a.fmt);
    }
  }
  if (kApply && a.bsum0 == nullptr && a.bsum1 == nullptr) return;
  // block reduction over the pixel lanes through the (now idle) ring
  asm volatile("bar.sync 1, %0;" ::"n"(kRingConsumers) : "memory");
  float* red = reinterpret_cast<float*>(ring);
  const int nacc = kApply ? 1 : 2;
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[(pl * C + cb + j) * nacc] = acc0[j];
      if (!kApply) red[(pl * C + cb + j) * nacc + 1] = acc1[j];
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kRingConsumers) : "memory");
  for (int i = threadIdx.x; i < nacc * C; i += kRingConsumers) {
    float tsum = 0.f;
    for (int l = 0; l < lanes; ++l) tsum += red[l * nacc * C + i];
    if (!kApply) {
      atomicAdd(&a.sums[static_cast<size_t>(b) * 2 * C + i], tsum);
    } else if (i < a.c0) {
      if (a.bsum0 != nullptr) atomicAdd(&a.bsum0[i], tsum);
    } else {
      if (a.bsum1 != nullptr) atomicAdd(&a.bsum1[i - a.c0], tsum);
    }
  }
}

__global__ void __launch_bounds__(256) resample_bwd_kernel(const uint16_t* __restrict__ dy, uint16_t* __restrict__ dx,
                                                           int B, int H, int W, int C, int mode, int fmt) {
  const int nvec = C >> 3;
  const size_t total = static_cast<size_t>(B) * H * W * nvec;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    size_t r = i / nvec;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    float o[8];
    if (mode == kResampleUp2) {
      const int Ho = 2 * H, Wo = 2 * W;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(b) * Ho + oy) * Wo + ox) * C + v * 8));
        float f[8];
        unpack8(u, fmt, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += f[j];
      }
    } else {
      const int Ho = H / 2, Wo = W / 2;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(b) * Ho + (y >> 1)) * Wo + (x >> 1)) * C + v * 8));
      unpack8(u, fmt, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
    }
    *reinterpret_cast<uint4*>(dx + ((static_cast<size_t>(b) * H + y) * W + x) * C + v * 8) = pack8(o, fmt);
  }
}

__global__ void __launch_bounds__(256) col_sum_kernel(const uint16_t* __restrict__ x, int64_t rows, int C, int P,
                                                      float* __restrict__ out, int fmt) {
  extern __shared__ float sh[];
  const int nvec = C >> 3;
  const int lanes = blockDim.x / nvec;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (pl < lanes) {
    const int64_t r_end = min(static_cast<int64_t>(blockIdx.x + 1) * P, rows);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * P + pl; r < r_end; r += lanes) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * C + v * 8)), fmt, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[pl * C + v * 8 + j] = s[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += sh[l * C + c];
    atomicAdd(&out[c], t);
  }
}

__global__ void __launch_bounds__(256) sum_f32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += x[i];
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// One thread owns 8 channels x 9 taps of partial sums over its pixels; pairs of lanes, then the 8 warps, are reduced
// before the atomics.  C == 128 (nvec == 16): lane l and l ^ 16 share a channel vector.
__global__ void __launch_bounds__(256) wgrad_1ch_kernel(const uint16_t* __restrict__ act, const float* __restrict__ img,
                                                        float* __restrict__ out, int B, int H, int W, int C, int sgn,
                                                        int P, int fmt) {
  __shared__ float red[8][16][72];
  const int nvec = C >> 3;  // 16
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int lanes = blockDim.x / nvec;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  // 32-bit pixel indices (the launcher checks B H W < 2^31): the 64-bit index splits were most of this kernel's
  // instructions
  const int total = B * H * W;
  const int p_end = min((static_cast<int>(blockIdx.x) + 1) * P, total);
  constexpr int kU = 4;  // pixels in flight per thread
  for (int p0 = static_cast<int>(blockIdx.x) * P + pl; p0 < p_end; p0 += kU * lanes) {
    uint4 u[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int p = p0 + k * lanes;
      if (p < p_end) u[k] = __ldg(reinterpret_cast<const uint4*>(act + static_cast<size_t>(p) * C + v * 8));
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int p = p0 + k * lanes;
      if (p >= p_end) continue;
      const int row = p / W;
      const int x = p - row * W;
      const int img_n = row / H;
      const int y = row - img_n * H;
      const int nb = img_n * H * W;
      float f[8];
      unpack8(u[k], fmt, f);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + sgn * (t / 3 - 1), xx = x + sgn * (t % 3 - 1);
        const float im = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + nb + yy * W + xx) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(f[j], im, acc[t][j]);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = acc[t][j] + __shfl_xor_sync(0xffffffffu, acc[t][j], 16);
      if (lane < 16) red[warp][lane][j * 9 + t] = s;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * 72; i += blockDim.x) {
    const int vv = i / 72, k = i % 72;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][vv][k];
    atomicAdd(&out[(vv * 8 + k / 9) * 9 + k % 9], s);
  }
}

// One thread owns one 8-channel vector (fixed for its whole grid-stride loop) and keeps its 72 weights in registers.
__global__ void __launch_bounds__(256) head_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                            uint16_t* __restrict__ dact, int B, int H, int W, int C,
                                                            int fmt) {
  const int nvec = C >> 3;  // divides the block size and therefore the grid stride
  const int v = threadIdx.x % nvec;
  float wr[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = __ldg(w + (v * 8 + j) * 9 + t);
  const int npix = B * H * W;  // < 2^31 (checked by the launcher): 32-bit index splits
  const int lanes = blockDim.x / nvec;
  for (int p = static_cast<int>(blockIdx.x) * lanes + threadIdx.x / nvec; p < npix;
       p += static_cast<int>(gridDim.x) * lanes) {
    const int row = p / W;
    const int x = p - row * W;
    const int img_n = row / H;
    const int y = row - img_n * H;
    const int nb = img_n * H * W;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
      const float d = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(dout + nb + yy * W + xx) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(d, wr[j][t], o[j]);
    }
    *reinterpret_cast<uint4*>(dact + static_cast<size_t>(p) * C + v * 8) = pack8(o, fmt);
  }
}

constexpr int kLbBT = 8;
// thread = one input column i, kLbBT batch rows; dy rows are staged through shared memory in chunks of 128 outputs
__global__ void __launch_bounds__(128) linear_bwd_input_kernel(const float* __restrict__ dy, int dy_stride,
                                                               const float* __restrict__ w32,
                                                               const uint16_t* __restrict__ w16, int fmt,
                                                               float* __restrict__ dx, int dx_stride,
                                                               const float* __restrict__ z, int z_stride, int B, int I,
                                                               int O) {
  __shared__ float sdy[kLbBT][128];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * kLbBT;
  float acc[kLbBT];
#pragma unroll
  for (int k = 0; k < kLbBT; ++k) acc[k] = 0.f;
  for (int o0 = 0; o0 < O; o0 += 128) {
    __syncthreads();
    for (int k = 0; k < kLbBT; ++k) {
      const int o = o0 + threadIdx.x;
      sdy[k][threadIdx.x] = (b0 + k < B && o < O) ? dy[static_cast<size_t>(b0 + k) * dy_stride + o] : 0.f;
    }
    __syncthreads();
    if (i < I) {
      const int on = min(128, O - o0);
      for (int oo = 0; oo < on; ++oo) {
        float wv;
        if (w16 != nullptr) {
          const uint16_t bits = w16[static_cast<size_t>(o0 + oo) * I + i];
          wv = fmt == 1 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&bits))
                        : __half2float(*reinterpret_cast<const __half*>(&bits));
        } else {
          wv = w32[static_cast<size_t>(o0 + oo) * I + i];
        }
#pragma unroll
        for (int k = 0; k < kLbBT; ++k) acc[k] = fmaf(sdy[k][oo], wv, acc[k]);
      }
    }
  }
  if (i < I) {
#pragma unroll
    for (int k = 0; k < kLbBT; ++k) {
      if (b0 + k < B) {
        float r = acc[k];
        if (z != nullptr) r *= silu_grad_precise(z[static_cast<size_t>(b0 + k) * z_stride + i]);
        dx[static_cast<size_t>(b0 + k) * dx_stride + i] = r;
      }
    }
  }
}

// Block = 128 input columns x 32 output rows; the batch rows of dy and act(x) for the tile are staged in shared
// memory once, then every thread (one column) accumulates its 32 outputs with float4 broadcast reads of dy.
constexpr int kLwO = 32;
constexpr int kLwBMax = 64;
__global__ void __launch_bounds__(128) linear_bwd_weight_kernel(const float* __restrict__ dy, int dy_stride,
                                                                const float* __restrict__ x, int x_stride, int act_x,
                                                                float* dw, float* db, int B,
                                                                int I, int O, const int64_t* __restrict__ tile_w_off,
                                                                const int64_t* __restrict__ tile_b_off) {
  __shared__ __align__(16) float sdy[kLwBMax][kLwO];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int o0 = blockIdx.y * kLwO;
  if (tile_w_off != nullptr) {
    // rows [o0, o0 + 32) of the concatenated matrix live at their own offsets of the gradient buffer
    dw += tile_w_off[blockIdx.y] - static_cast<int64_t>(o0) * I;
    db += tile_b_off[blockIdx.y] - o0;
  }
  float acc[kLwO];
#pragma unroll
  for (int k = 0; k < kLwO; ++k) acc[k] = 0.f;
  for (int b0 = 0; b0 < B; b0 += kLwBMax) {
    const int nb = min(kLwBMax, B - b0);
    __syncthreads();
    for (int e = threadIdx.x; e < nb * kLwO; e += blockDim.x) {
      const int bb = e / kLwO, oo = e % kLwO;
      sdy[bb][oo] = (o0 + oo < O) ? dy[static_cast<size_t>(b0 + bb) * dy_stride + o0 + oo] : 0.f;
    }
    __syncthreads();
    if (i < I) {
      for (int bb = 0; bb < nb; ++bb) {
        float xv = x[static_cast<size_t>(b0 + bb) * x_stride + i];
        if (act_x) xv = silu_precise(xv);
#pragma unroll
        for (int k = 0; k < kLwO; k += 4) {
          const float4 d = *reinterpret_cast<const float4*>(&sdy[bb][k]);
          acc[k] = fmaf(d.x, xv, acc[k]);
          acc[k + 1] = fmaf(d.y, xv, acc[k + 1]);
          acc[k + 2] = fmaf(d.z, xv, acc[k + 2]);
          acc[k + 3] = fmaf(d.w, xv, acc[k + 3]);
        }
      }
    }
    if (db != nullptr && blockIdx.x == 0 && threadIdx.x < kLwO && o0 + threadIdx.x < O) {
      float t = b0 == 0 ? 0.f : db[o0 + threadIdx.x];
      for (int bb = 0; bb < nb; ++bb) t += sdy[bb][threadIdx.x];
      db[o0 + threadIdx.x] = t;
    }
  }
  if (i < I) {
#pragma unroll
    for (int k = 0; k < kLwO; ++k)
      if (o0 + k < O) dw[static_cast<size_t>(o0 + k) * I + i] = acc[k];
  }
}

__global__ void mul_silu_grad_kernel(float* __restrict__ dx, const float* __restrict__ z, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dx[i] *= silu_grad_precise(z[i]);
}

constexpr int kAdamChunk = 4096;
__global__ void __launch_bounds__(256) adam_multi_kernel(float* const* __restrict__ pp, const float* const* __restrict__ gp,
                                                         float* const* __restrict__ mp, float* const* __restrict__ vp,
                                                         const int64_t* __restrict__ numel,
                                                         const int* __restrict__ block_tensor,
                                                         const int64_t* __restrict__ block_off, float step_size,
                                                         float beta1, float beta2, float eps, float inv_sqrt_bc2) {
  const int t = block_tensor[blockIdx.x];
  const float* g = gp[t];
  if (g == nullptr) return;
  float* p = pp[t];
  float* m = mp[t];
  float* v = vp[t];
  const int64_t n = numel[t];
  const int64_t lo = block_off[blockIdx.x];
  const int64_t hi = min(lo + kAdamChunk, n);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](float& pe, float ge, float& me, float& ve) {
    me = beta1 * me + (1.0f - beta1) * ge;
    ve = beta2 * ve + (1.0f - beta2) * ge * ge;
    pe -= step_size * me / (sqrtf(ve) * inv_sqrt_bc2 + eps);
  };
  if (vec) {
    const int64_t hi4 = lo + ((hi - lo) & ~int64_t(3));
    for (int64_t i = lo + 4 * threadIdx.x; i < hi4; i += 4 * blockDim.x) {
      float4 p4 = *reinterpret_cast<float4*>(p + i);
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 m4 = *reinterpret_cast<float4*>(m + i);
      float4 v4 = *reinterpret_cast<float4*>(v + i);
      upd(p4.x, g4.x, m4.x, v4.x);
      upd(p4.y, g4.y, m4.y, v4.y);
      upd(p4.z, g4.z, m4.z, v4.z);
      upd(p4.w, g4.w, m4.w, v4.w);
      *reinterpret_cast<float4*>(p + i) = p4;
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
    }
    for (int64_t i = hi4 + threadIdx.x; i < hi; i += blockDim.x) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) upd(p[i], g[i], m[i], v[i]);
  }
}

__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

// Chunks per image for the GroupNorm backward kernels: the grid (chunks x B blocks, 2 resident per SM) should fill
// whole waves - at 10 chunks x 64 images the second wave ran 44 % full and cost a quarter of the kernel's time.
int pick_blocks(int B, int HW, int lanes) {
  const int slots = 2 * device_sm_count();
  int max_chunks = HW / (24 * lanes);
  if (max_chunks > 64) max_chunks = 64;
  if (max_chunks < 1) max_chunks = 1;
  int best = 1;
  double best_eff = 0.0;
  for (int ch = 1; ch <= max_chunks; ++ch) {
    const int blocks = ch * B;
    const int waves = (blocks + slots - 1) / slots;
    const double eff = static_cast<double>(blocks) / (static_cast<double>(waves) * slots);
    if (eff > best_eff + 0.02 || (eff >= best_eff && ch <= 2 * best)) {
      best_eff = eff > best_eff ? eff : best_eff;
      best = ch;
    }
  }
  return best;
}

bool gn_ring_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDDPM_GN_BWD_RING");  // A/B switch for measurements: 0 = register-fed kernels everywhere
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

int launch_gn_bwd(const GnBwdArgs& g, cudaStream_t stream) {
  const int C = g.x.C();
  if (C % 32 != 0 || C / 8 > 256) return fail(kUnsupported, "gn_bwd: unsupported channel count");
  if ((C / kGnGroups) % 4 != 0) return fail(kUnsupported, "gn_bwd: group size must be a multiple of 4");
  if (g.x.c0 % 8 != 0 || g.x.c1 % 8 != 0) return fail(kUnsupported, "gn_bwd: concat members must be multiples of 8");
  if (!g.stats0 || (g.x.c1 > 0 && !g.stats1) || !g.dy || !g.sums || !g.out0 || (g.x.c1 > 0 && !g.out1) ||
      !g.dgamma || !g.dbeta || !g.gamma || !g.beta)
    return fail(kInvalidArgument, "gn_bwd: null pointer");
  GnBwdDev a;
  a.p0 = reinterpret_cast<const uint16_t*>(g.x.p0);
  a.p1 = reinterpret_cast<const uint16_t*>(g.x.p1);
  a.c0 = g.x.c0;
  a.c1 = g.x.c1;
  a.B = g.B;
  a.HW = g.H * g.W;
  a.stats0 = g.stats0;
  a.stats1 = g.stats1;
  a.gamma = g.gamma;
  a.beta = g.beta;
  a.film = g.film;
  a.film_stride = g.film_stride;
  a.film_off = g.film_off;
  a.silu = g.silu;
  a.dy = reinterpret_cast<const uint16_t*>(g.dy);
  a.add0 = reinterpret_cast<const uint16_t*>(g.add0);
  a.add1 = reinterpret_cast<const uint16_t*>(g.add1);
  a.sums = g.sums;
  a.out0 = reinterpret_cast<uint16_t*>(g.out0);
  a.out1 = reinterpret_cast<uint16_t*>(g.out1);
  a.bsum0 = g.bsum0;
  a.bsum1 = g.bsum1;
  a.dgamma = g.dgamma;
  a.dbeta = g.dbeta;
  a.dfilm = g.dfilm;
  a.fmt = g.fmt;
  const int nvec = C / 8;
  if (gn_ring_enabled() && a.HW >= 2304 && kRingConsumers / nvec >= 1) {
    const int rl = kRingConsumers / nvec;  // pixels per tile: one per consumer thread
    const int ntens = 2 + (g.add0 != nullptr ? 1 : 0) + (g.add1 != nullptr ? 1 : 0);
    const int stage_apply = rl * C * 2 * ntens;
    const int stage_reduce = rl * C * 2 * 2;
    const int slots = device_sm_count();
    int max_chunks = a.HW / (16 * rl);
    if (max_chunks < 1) max_chunks = 1;
    if (max_chunks > 64) max_chunks = 64;
    auto fill = [&](int ch) {
      const int blocks = ch * g.B;
      const int waves = (blocks + slots - 1) / slots;
      return static_cast<double>(blocks) / (static_cast<double>(waves) * slots);
    };
    double best_eff = 0.0;
    for (int ch = 1; ch <= max_chunks; ++ch) best_eff = fill(ch) > best_eff ? fill(ch) : best_eff;
    int best = 1;
    for (int ch = 1; ch <= max_chunks; ++ch)
      if (fill(ch) >= best_eff - 0.03) {  // fewest chunks (least prologue work) within 3 % of the best wave fill
        best = ch;
        break;
      }
    a.P = (a.HW + best - 1) / best;
    const dim3 rgrid((a.HW + a.P - 1) / a.P, g.B);
    static bool attr_set = false;
    if (!attr_set) {
      CDDPM_CUDA(cudaFuncSetAttribute(gn_bwd_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      CDDPM_CUDA(cudaFuncSetAttribute(gn_bwd_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr_set = true;
    }
    const size_t sh_r = static_cast<size_t>(kRingStages) * stage_reduce + static_cast<size_t>(2) * C * sizeof(float);
    const size_t sh_a = static_cast<size_t>(kRingStages) * stage_apply + static_cast<size_t>(2) * C * sizeof(float);
    if (sh_a <= 220 * 1024) {
      gn_bwd_ring_kernel<false><<<rgrid, kRingThreads, sh_r, stream>>>(a, rl, stage_reduce);
      CDDPM_TRY(check_launch("gn_bwd_ring_kernel<reduce>"));
      gn_bwd_ring_kernel<true><<<rgrid, kRingThreads, sh_a, stream>>>(a, rl, stage_apply);
      return check_launch("gn_bwd_ring_kernel<apply>");
    }
  }
  const int threads = 256;
  const int lanes = threads / nvec;
  if (lanes < 1) return fail(kUnsupported, "gn_bwd: too many channels");
  const int chunks = pick_blocks(g.B, a.HW, lanes);
  a.P = (a.HW + chunks - 1) / chunks;
  const dim3 grid((a.HW + a.P - 1) / a.P, g.B);
  size_t sh_reduce = static_cast<size_t>(lanes) * C * 2 * sizeof(float);
  if (sh_reduce < static_cast<size_t>(2) * C * sizeof(float)) sh_reduce = static_cast<size_t>(2) * C * sizeof(float);
  gn_bwd_reduce_kernel<<<grid, threads, sh_reduce, stream>>>(a);
  CDDPM_TRY(check_launch("gn_bwd_reduce_kernel"));
  size_t sh_apply = static_cast<size_t>(lanes) * C * sizeof(float);
  if (sh_apply < static_cast<size_t>(2) * C * sizeof(float)) sh_apply = static_cast<size_t>(2) * C * sizeof(float);
  gn_bwd_apply_kernel<<<grid, threads, sh_apply, stream>>>(a);
  return check_launch("gn_bwd_apply_kernel");
}

int launch_resample_bwd(const void* dy, void* dx, int B, int H, int W, int C, int mode, int fmt, cudaStream_t stream) {
  if (mode != kResampleUp2 && mode != kResampleDown2) return fail(kInvalidArgument, "resample_bwd: bad mode");
  if (C % 8 != 0 || (mode == kResampleDown2 && (H % 2 != 0 || W % 2 != 0)))
    return fail(kUnsupported, "resample_bwd: unsupported geometry");
  const size_t total = static_cast<size_t>(B) * H * W * (C / 8);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 8 * device_sm_count()) blocks = 8 * device_sm_count();
  resample_bwd_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(dy),
                                                  reinterpret_cast<uint16_t*>(dx), B, H, W, C, mode, fmt);
  return check_launch("resample_bwd_kernel");
}

int launch_col_sum(const void* x, int64_t rows, int C, float* out, int fmt, cudaStream_t stream) {
  if (C % 8 != 0 || C / 8 > 256) return fail(kUnsupported, "col_sum: unsupported channel count");
  const int lanes = 256 / (C / 8);
  int blocks = 4 * device_sm_count();
  const int64_t max_blocks = (rows + lanes - 1) / lanes;
  if (blocks > max_blocks) blocks = static_cast<int>(max_blocks);
  if (blocks < 1) blocks = 1;
  const int P = static_cast<int>((rows + blocks - 1) / blocks);
  col_sum_kernel<<<blocks, 256, static_cast<size_t>(lanes) * C * sizeof(float), stream>>>(
      reinterpret_cast<const uint16_t*>(x), rows, C, P, out, fmt);
  return check_launch("col_sum_kernel");
}

int launch_sum_f32(const float* x, int64_t n, float* out, cudaStream_t stream) {
  int blocks = static_cast<int>((n + 1023) / 1024);
  if (blocks > 2 * device_sm_count()) blocks = 2 * device_sm_count();
  if (blocks < 1) blocks = 1;
  sum_f32_kernel<<<blocks, 256, 0, stream>>>(x, n, out);
  return check_launch("sum_f32_kernel");
}

int launch_wgrad_1ch(const void* act, const float* img, float* out, int B, int H, int W, int C, int sgn, int fmt,
                     cudaStream_t stream) {
  if (C != 128) return fail(kUnsupported, "wgrad_1ch: the stem / head gradient kernel expects 128 channels");
  const int64_t total = static_cast<int64_t>(B) * H * W;
  if (total >= (1ll << 31) - (1 << 20)) return fail(kUnsupported, "wgrad_1ch: more than 2^31 pixels in one call");
  int blocks = 2 * device_sm_count();
  if (blocks > (total + 15) / 16) blocks = static_cast<int>((total + 15) / 16);
  const int P = static_cast<int>((total + blocks - 1) / blocks);
  wgrad_1ch_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(act), img, out, B, H, W, C, sgn, P, fmt);
  return check_launch("wgrad_1ch_kernel");
}

int launch_head_bwd_data(const float* dout, const float* w, void* dact, int B, int H, int W, int C, int fmt,
                         cudaStream_t stream) {
  if (C % 8 != 0 || 256 % (C / 8) != 0) return fail(kUnsupported, "head_bwd_data: unsupported channel count");
  if (static_cast<int64_t>(B) * H * W >= (1ll << 31) - (1 << 24))
    return fail(kUnsupported, "head_bwd_data: more than 2^31 pixels in one call");
  const int64_t total = static_cast<int64_t>(B) * H * W * (C / 8);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 8 * device_sm_count()) blocks = 8 * device_sm_count();
  head_bwd_data_kernel<<<blocks, 256, 0, stream>>>(
      dout, w, reinterpret_cast<uint16_t*>(dact), B, H, W, C, fmt);
  return check_launch("head_bwd_data_kernel");
}

int launch_linear_bwd_input(const float* dy, int dy_stride, const float* w32, const void* w16, int fmt, float* dx,
                            int dx_stride, const float* z, int z_stride, int B, int I, int O, cudaStream_t stream) {
  const dim3 grid((I + 127) / 128, (B + kLbBT - 1) / kLbBT);
  linear_bwd_input_kernel<<<grid, 128, 0, stream>>>(dy, dy_stride, w32, reinterpret_cast<const uint16_t*>(w16), fmt, dx,
                                                    dx_stride, z, z_stride, B, I, O);
  return check_launch("linear_bwd_input_kernel");
}

int launch_linear_bwd_weight(const float* dy, int dy_stride, const float* x, int x_stride, int act_x, float* dw,
                             float* db, int B, int I, int O, cudaStream_t stream) {
  const dim3 grid((I + 127) / 128, (O + kLwO - 1) / kLwO);
  linear_bwd_weight_kernel<<<grid, 128, 0, stream>>>(dy, dy_stride, x, x_stride, act_x, dw, db, B, I, O, nullptr,
                                                     nullptr);
  return check_launch("linear_bwd_weight_kernel");
}

int launch_linear_bwd_weight_tiled(const float* dy, int dy_stride, const float* x, int x_stride, int act_x,
                                   float* grads, const int64_t* tile_w_off, const int64_t* tile_b_off, int B, int I,
                                   int O, cudaStream_t stream) {
  if (O % kLwO != 0) return fail(kInvalidArgument, "linear_bwd_weight_tiled: rows must come in tiles of 32");
  const dim3 grid((I + 127) / 128, O / kLwO);
  linear_bwd_weight_kernel<<<grid, 128, 0, stream>>>(dy, dy_stride, x, x_stride, act_x, grads, grads, B, I, O,
                                                     tile_w_off, tile_b_off);
  return check_launch("linear_bwd_weight_kernel");
}

int launch_mul_silu_grad(float* dx, const float* z, int64_t n, cudaStream_t stream) {
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  mul_silu_grad_kernel<<<blocks, 256, 0, stream>>>(dx, z, n);
  return check_launch("mul_silu_grad_kernel");
}

int launch_adam_step(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel,
                     const int* block_tensor, const int64_t* block_off, int total_blocks, float lr, float beta1,
                     float beta2, float eps, float bc1, float bc2, cudaStream_t stream) {
  if (!p || !g || !m || !v || !numel || !block_tensor || !block_off) return fail(kInvalidArgument, "adam_step: null pointer");
  if (total_blocks <= 0) return kOk;
  adam_multi_kernel<<<total_blocks, 256, 0, stream>>>(p, g, m, v, numel, block_tensor, block_off, lr / bc1, beta1, beta2,
                                                      eps, 1.0f / sqrtf(bc2));
  return check_launch("adam_multi_kernel");
}

int launch_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t stream) {
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  copy_f32_kernel<<<blocks, 256, 0, stream>>>(src, dst, n);
  return check_launch("copy_f32_kernel");
}

}  // namespace cddpm
