// Self-attention core, head dim 64.  Two kernels: a tcgen05 one for the shapes the UNet uses (see attention_tc_kernel
// below) and a general SIMT one, in which one thread owns one query row (q and the output accumulator live in
// registers, fp32), keys/values stream through shared memory in chunks that every thread reads with broadcast loads,
// and the softmax is the usual online (running max / running sum) form evaluated per chunk.
#include "attention.cuh"

#include <cuda_fp16.h>
#include <math_constants.h>
#include <string.h>

#include "ptx.cuh"

namespace cddpm {

namespace {

constexpr int kHeadDim = 64;
constexpr int kQueriesPerCta = 192;
constexpr int kKeyChunk = 32;

__device__ __forceinline__ float2 cvt2(uint32_t u, int fmt) {
  if (fmt == 1) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ uint32_t pk2(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kQueriesPerCta) attention_kernel(const uint16_t* __restrict__ qkv,
                                                                    uint16_t* __restrict__ out, int L, int C,
                                                                    int fmt) {
  __shared__ __align__(16) float sK[kKeyChunk][kHeadDim];
  __shared__ __align__(16) float sV[kKeyChunk][kHeadDim];
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * kQueriesPerCta + threadIdx.x;
  const bool active = t < L;
  const size_t row_stride = static_cast<size_t>(3) * C;
  const uint16_t* base = qkv + static_cast<size_t>(b) * L * row_stride;
  const float scale = rsqrtf(static_cast<float>(kHeadDim));  // (d^-1/4)^2

  float q[kHeadDim], acc[kHeadDim];
#pragma unroll
  for (int c = 0; c < kHeadDim; ++c) {
    q[c] = 0.f;
    acc[c] = 0.f;
  }
  if (active) {
    const uint4* qp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * row_stride + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 8; ++i) {
      const uint4 u = __ldg(qp + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = cvt2(w[e], fmt);
        q[i * 8 + e * 2] = f.x * scale;
        q[i * 8 + e * 2 + 1] = f.y * scale;
      }
    }
  }
  float m = -CUDART_INF_F, l = 0.f;

  for (int k0 = 0; k0 < L; k0 += kKeyChunk) {
    __syncthreads();
    // stage K and V rows [k0, k0+32) of this head as fp32
    for (int i = threadIdx.x; i < kKeyChunk * (kHeadDim / 2) * 2; i += blockDim.x) {
      const int which = i / (kKeyChunk * (kHeadDim / 2));
      const int r = (i / (kHeadDim / 2)) % kKeyChunk;
      const int c2 = i % (kHeadDim / 2);
      const int s = k0 + r;
      float2 f = make_float2(0.f, 0.f);
      if (s < L) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(s) * row_stride +
                                                                  (which + 1) * C + h * kHeadDim) + c2);
        f = cvt2(u, fmt);
      }
      float* dst = which == 0 ? &sK[r][c2 * 2] : &sV[r][c2 * 2];
      dst[0] = f.x;
      dst[1] = f.y;
    }
    __syncthreads();
    float sc[kKeyChunk];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < kKeyChunk; ++j) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < kHeadDim; c += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&sK[j][c]);
        d = fmaf(q[c], kv.x, d);
        d = fmaf(q[c + 1], kv.y, d);
        d = fmaf(q[c + 2], kv.z, d);
        d = fmaf(q[c + 3], kv.w, d);
      }
      if (k0 + j >= L) d = -CUDART_INF_F;
      sc[j] = d;
      mx = fmaxf(mx, d);
    }
    const float m_new = fmaxf(m, mx);
    const float corr = __expf(m - m_new);  // m = -inf on the first chunk -> 0
    l *= corr;
#pragma unroll
    for (int c = 0; c < kHeadDim; ++c) acc[c] *= corr;
#pragma unroll
    for (int j = 0; j < kKeyChunk; ++j) {
      const float p = __expf(sc[j] - m_new);
      l += p;
#pragma unroll
      for (int c = 0; c < kHeadDim; c += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&sV[j][c]);
        acc[c] = fmaf(p, vv.x, acc[c]);
        acc[c + 1] = fmaf(p, vv.y, acc[c + 1]);
        acc[c + 2] = fmaf(p, vv.z, acc[c + 2]);
        acc[c + 3] = fmaf(p, vv.w, acc[c + 3]);
      }
    }
    m = m_new;
  }
  if (active) {
    const float inv = 1.0f / l;
    uint4* op = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * L + t) * C + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 8; ++i) {
      uint4 o;
      o.x = pk2(acc[i * 8 + 0] * inv, acc[i * 8 + 1] * inv, fmt);
      o.y = pk2(acc[i * 8 + 2] * inv, acc[i * 8 + 3] * inv, fmt);
      o.z = pk2(acc[i * 8 + 4] * inv, acc[i * 8 + 5] * inv, fmt);
      o.w = pk2(acc[i * 8 + 6] * inv, acc[i * 8 + 7] * inv, fmt);
      op[i] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core path (L a multiple of 64, K and V of one head resident in shared memory, i.e. L <= 576).
// One CTA per (128-query tile, head, image).  Q [128 x 64], K [L x 64] and V [L x 64] arrive by TMA as 128-byte
// swizzled rows.  S = Q K^T is a K-major x K-major tcgen05.mma into TMEM, one chunk of kc keys (<= 192 columns) per
// accumulator, two accumulators in flight; every thread owns one query row of the accumulator.
//   pass 1: row maximum over all chunks;
//   pass 2: S again (the MMAs are ~free: the whole op is 0.2 % of the UNet's FLOPs), p = exp2((s - max) * c) rounded
//           to the 16-bit type into a K-major shared-memory tile, O += P V with V as an MN-major B operand (its rows
//           are keys; tools/ubench_mma.cu checks that descriptor form).  Knowing the maximum up front, O is never
//           rescaled.
// The division by the row sum happens on the way out.
// ------------------------------------------------------------------------------------------------------------
constexpr int kTcThreads = 128;
constexpr int kTcQ = 128;         // queries per CTA (UMMA M)
constexpr int kTcMaxL = 576;
constexpr int kTcMaxChunk = 192;  // keys per S accumulator (UMMA N)
constexpr int kTcSmem = kTcQ * 128 + 2 * kTcMaxL * 128 + kTcQ * kTcMaxChunk * 2 + 1024 + 64;

struct AttnTcParams {
  CUtensorMap tmap_q;   // {3C, B*L}, box {64, 128}
  CUtensorMap tmap_kv;  // {3C, B*L}, box {64, kc}
  uint16_t* out;
  int L, C, kc, fmt;
};

__global__ void __launch_bounds__(kTcThreads, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_sm = smem;
  uint8_t* k_sm = q_sm + kTcQ * 128;
  uint8_t* v_sm = k_sm + kTcMaxL * 128;
  uint8_t* p_sm = v_sm + kTcMaxL * 128;  // kc/64 atoms of [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_sm + kTcQ * kTcMaxChunk * 2);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;  // [2]
  uint64_t* bar_pv = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kTcQ;
  const int L = p.L, C = p.C, kc = p.kc, fmt = p.fmt;
  pdl_trigger();
  const int nchunks = L / kc;
  const int row = threadIdx.x;  // query row of this thread = TMEM lane

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_kv);
    mbar_init(bar_load, 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(bar_pv, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s[2] = {tmem_base, tmem_base + kTcMaxChunk};
  const uint32_t tmem_o = tmem_base + 2 * kTcMaxChunk;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;

  const uint32_t idesc_s = umma_idesc_f16(kTcQ, static_cast<uint32_t>(kc), static_cast<uint32_t>(fmt));
  const uint32_t idesc_o = umma_idesc_f16(kTcQ, kHeadDim, static_cast<uint32_t>(fmt)) | (1u << 16);  // B = V is MN-major
  const uint32_t q_addr = smem_u32(q_sm), k_addr = smem_u32(k_sm), v_addr = smem_u32(v_sm), p_addr = smem_u32(p_sm);

  auto issue_s = [&](int c, int buf) {  // S[buf] = Q K_c^T, then signal bar_s[buf]
#pragma unroll
    for (int kk = 0; kk < kHeadDim / 16; ++kk)
      umma_f16_ss(tmem_s[buf], umma_desc_k128(q_addr + kk * 32), umma_desc_k128(k_addr + c * kc * 128 + kk * 32),
                  idesc_s, kk != 0 ? 1u : 0u);
    umma_commit(&bar_s[buf]);
  };

  pdl_wait();  // qkv is the predecessor's output (prologue above ran under its tail)
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(kTcQ * 128 + 2 * L * 128));
    tma_load_2d(q_sm, &p.tmap_q, bar_load, h * kHeadDim, b * L + q0);
    for (int c = 0; c < nchunks; ++c) {
      tma_load_2d(k_sm + c * kc * 128, &p.tmap_kv, bar_load, C + h * kHeadDim, b * L + c * kc);
      tma_load_2d(v_sm + c * kc * 128, &p.tmap_kv, bar_load, 2 * C + h * kHeadDim, b * L + c * kc);
    }
  }
  mbar_wait(bar_load, 0);
  tc_fence_after();

  uint32_t uses[2] = {0, 0};  // completed phases of bar_s[buf]
  // ---------------------------------------------------------------- pass 1: row maximum
  if (threadIdx.x == 0) {
    issue_s(0, 0);
    if (nchunks > 1) issue_s(1, 1);
  }
  float m = -CUDART_INF_F;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    mbar_wait(&bar_s[buf], uses[buf] & 1);
    ++uses[buf];
    tc_fence_after();
    for (int j0 = 0; j0 < kc; j0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_s[buf] + lane_off + j0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
    }
    tc_fence_before();
    __syncthreads();  // every row of S[buf] has been read
    if (threadIdx.x == 0 && c + 2 < nchunks) {
      tc_fence_after();
      issue_s(c + 2, buf);
    }
  }
  // ---------------------------------------------------------------- pass 2: P = exp(S - max), O += P V
  if (threadIdx.x == 0) {
    tc_fence_after();
    issue_s(0, 0);
    if (nchunks > 1) issue_s(1, 1);
  }
  // softmax((q . k) / sqrt(64)): exp2 with the scale and log2(e) folded into one constant
  const float cexp = 0.125f * 1.4426950408889634f;
  const float mc = m * cexp;
  float l = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    mbar_wait(&bar_s[buf], uses[buf] & 1);
    ++uses[buf];
    if (c > 0) mbar_wait(bar_pv, (c - 1) & 1);  // P V of the previous chunk is done: the P tile is free again
    tc_fence_after();
    for (int j0 = 0; j0 < kc; j0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_s[buf] + lane_off + j0, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = exp2f(fmaf(__uint_as_float(v[j]), cexp, -mc));
        const float p1 = exp2f(fmaf(__uint_as_float(v[j + 1]), cexp, -mc));
        l += p0 + p1;
        pk[j >> 1] = pk2(p0, p1, fmt);
      }
      // keys j0 .. j0+31 of this row: four 16-byte chunks in atom j0 / 64, K-major with the 128-byte swizzle
      const uint32_t atom = p_addr + (j0 >> 6) * (kTcQ * 128) + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = ((j0 & 63) >> 3) + q;
        const uint32_t addr = atom + ((chunk ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                     : "memory");
      }
    }
    fence_proxy_async_smem();  // P was written through the generic proxy; the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      for (int j = 0; j < kc / 16; ++j) {
        umma_f16_ss(tmem_o, umma_desc_k128(p_addr + (j >> 2) * (kTcQ * 128) + (j & 3) * 32),
                    umma_desc_k128(v_addr + (c * kc + j * 16) * 128), idesc_o, (c | j) != 0 ? 1u : 0u);
      }
      umma_commit(bar_pv);
      if (c + 2 < nchunks) issue_s(c + 2, buf);
    }
  }
  mbar_wait(bar_pv, (nchunks - 1) & 1);
  tc_fence_after();
  // ---------------------------------------------------------------- O / l -> global
  const int t = q0 + row;
  const float inv = 1.0f / l;
#pragma unroll
  for (int c0 = 0; c0 < kHeadDim; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_o + lane_off + c0, v);
    tmem_ld_wait();
    if (t < L) {
      uint4* op = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * L + t) * C + h * kHeadDim + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pk2(__uint_as_float(v[i * 8 + 0]) * inv, __uint_as_float(v[i * 8 + 1]) * inv, fmt);
        o.y = pk2(__uint_as_float(v[i * 8 + 2]) * inv, __uint_as_float(v[i * 8 + 3]) * inv, fmt);
        o.z = pk2(__uint_as_float(v[i * 8 + 4]) * inv, __uint_as_float(v[i * 8 + 5]) * inv, fmt);
        o.w = pk2(__uint_as_float(v[i * 8 + 6]) * inv, __uint_as_float(v[i * 8 + 7]) * inv, fmt);
        op[i] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_attention_tc(const void* qkv, void* out, int B, int L, int C, int fmt, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    attr_set = true;
  }
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<uint16_t*>(out);
  p.L = L;
  p.C = C;
  p.fmt = fmt;
  p.kc = (L % 192 == 0) ? 192 : ((L % 128 == 0) ? 128 : 64);
  const uint64_t dims[2] = {static_cast<uint64_t>(3) * C, static_cast<uint64_t>(B) * L};
  const uint64_t strides[1] = {static_cast<uint64_t>(3) * C * 2};
  const uint32_t box_q[2] = {static_cast<uint32_t>(kHeadDim), static_cast<uint32_t>(kTcQ)};
  const uint32_t box_kv[2] = {static_cast<uint32_t>(kHeadDim), static_cast<uint32_t>(p.kc)};
  CDDPM_TRY(encode_tmap_16bit(&p.tmap_q, qkv, 2, dims, strides, box_q));
  CDDPM_TRY(encode_tmap_16bit(&p.tmap_kv, qkv, 2, dims, strides, box_kv));
  dim3 grid((L + kTcQ - 1) / kTcQ, C / kHeadDim, B);
  return check_cuda(launch_k(attention_tc_kernel, grid, dim3(kTcThreads), kTcSmem, stream, p), "attention_tc_kernel");
}

}  // namespace

int launch_attention(const void* qkv, void* out, int B, int L, int C, int fmt, cudaStream_t stream) {
  if (!qkv || !out) return fail(kInvalidArgument, "attention: null pointer");
  if (C % kHeadDim != 0) return fail(kUnsupported, "attention: channels must be a multiple of the head dim 64");
  if (L % 64 == 0 && L <= kTcMaxL) return launch_attention_tc(qkv, out, B, L, C, fmt, stream);
  dim3 grid((L + kQueriesPerCta - 1) / kQueriesPerCta, C / kHeadDim, B);
  attention_kernel<<<grid, kQueriesPerCta, 0, stream>>>(reinterpret_cast<const uint16_t*>(qkv),
                                                        reinterpret_cast<uint16_t*>(out), L, C, fmt);
  return check_launch("attention_kernel");
}

}  // namespace cddpm
