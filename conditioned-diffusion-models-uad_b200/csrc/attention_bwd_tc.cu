// Backward of the self-attention core (QKVAttention, OpenAI_Unet.py:457-476) on tcgen05 tensor cores.
//
// Same two-kernel split and the same scratch as the first (mma.sync) version in attention_bwd.cu - probabilities P and
// scaled score gradients dS materialised once as 16-bit [B*heads][L][L] - but every product is a tcgen05.mma with its
// accumulator in TMEM, operands staged by TMA, one thread per accumulator row (the structure of the forward kernel,
// attention.cu):
//   kernel Q  (128 queries, head, image): K, V [L x 64], Q, dO [128 x 64] resident in shared memory.
//       pass 1  S = Q K^T and dP = dO V^T per chunk of kc keys (two K-major x K-major MMAs into two TMEM accumulators);
//               online row maximum m, row sum l and sum(e^(s - m) dP): D = rowsum(P * dP) without a third pass;
//       pass 2  S and dP again (the MMAs are ~free), p = e^(s - m) / l, dS = p (dP - D) / 8 -> scratch, and dS also
//               into a K-major shared-memory tile: dQ += dS K with K as an MN-major B operand (its rows are keys).
//   kernel KV (128 keys, head, image):  dV = P^T dO and dK = dS^T Q.  The contraction runs over QUERIES, and in the
//       row-major scratch / activations both operands have their M / N index contiguous: both are MN-major operands
//       straight from the TMA-staged tiles (the descriptor form of conv_wgrad.cu), query chunks of kq rows.
// L must be a multiple of 64 and <= 576 (the UNet's 24 x 24 attention level); other shapes keep the mma.sync kernels.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "attention.cuh"
#include "ptx.cuh"

namespace cddpm {

namespace {

constexpr int kD = 64;          // head dim
constexpr int kRows = 128;      // accumulator rows per CTA (UMMA M): queries (kernel Q) or keys (kernel KV)
constexpr int kMaxL = 576;
constexpr int kMaxChunk = 192;  // keys per S / dP accumulator; queries per staged chunk in kernel KV
constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t pk2f(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// MN-major operand (conv_wgrad.cu): rows of 128 bytes (64 M/N elements) per K index, 8-row groups of 1024 bytes (SBO),
// the next 64 M/N elements `lbo` bytes further.
__device__ __forceinline__ uint64_t desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct BwdQParams {
  CUtensorMap tmap_q;   // qkv {3C, B*L}, box {64, 128}
  CUtensorMap tmap_kv;  // qkv {3C, B*L}, box {64, kc}
  CUtensorMap tmap_do;  // dout {C, B*L}, box {64, 128}
  uint16_t* dqkv;
  uint16_t* P;
  uint16_t* dS;
  int L, C, kc, fmt;
};

constexpr int kQSmem = 2 * kRows * 128 + 2 * kMaxL * 128 + kRows * kMaxChunk * 2 + 1024 + 64;

__global__ void __launch_bounds__(kThreads, 1) attn_bwd_q_tc_kernel(const __grid_constant__ BwdQParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_sm = smem;
  uint8_t* do_sm = q_sm + kRows * 128;
  uint8_t* k_sm = do_sm + kRows * 128;
  uint8_t* v_sm = k_sm + kMaxL * 128;
  uint8_t* ds_sm = v_sm + kMaxL * 128;  // kc / 64 atoms of [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_sm + kRows * kMaxChunk * 2);
  uint64_t* bar_load = bars;
  uint64_t* bar_mma = bars + 1;
  uint64_t* bar_dq = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kRows;
  const int L = p.L, C = p.C, kc = p.kc, fmt = p.fmt;
  const int nchunks = L / kc;
  const int heads = C / kD;
  const int row = threadIdx.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_kv);
    tma_prefetch_desc(&p.tmap_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_dq, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + kMaxChunk, tmem_dq = tmem_base + 2 * kMaxChunk;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t idesc_s = umma_idesc_f16(kRows, static_cast<uint32_t>(kc), static_cast<uint32_t>(fmt));
  const uint32_t idesc_dq = umma_idesc_f16(kRows, kD, static_cast<uint32_t>(fmt)) | (1u << 16);  // B = K is MN-major
  const uint32_t q_addr = smem_u32(q_sm), do_addr = smem_u32(do_sm), k_addr = smem_u32(k_sm), v_addr = smem_u32(v_sm),
                 ds_addr = smem_u32(ds_sm);

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(2 * kRows * 128 + 2 * L * 128));
    tma_load_2d(q_sm, &p.tmap_q, bar_load, h * kD, b * L + q0);
    tma_load_2d(do_sm, &p.tmap_do, bar_load, h * kD, b * L + q0);
    for (int c = 0; c < nchunks; ++c) {
      tma_load_2d(k_sm + c * kc * 128, &p.tmap_kv, bar_load, C + h * kD, b * L + c * kc);
      tma_load_2d(v_sm + c * kc * 128, &p.tmap_kv, bar_load, 2 * C + h * kD, b * L + c * kc);
    }
  }
  mbar_wait(bar_load, 0);
  tc_fence_after();

  auto issue_s_dp = [&](int c) {  // S = Q K_c^T, dP = dO V_c^T, then signal bar_mma
#pragma unroll
    for (int kk = 0; kk < kD / 16; ++kk)
      umma_f16_ss(tmem_s, umma_desc_k128(q_addr + kk * 32), umma_desc_k128(k_addr + c * kc * 128 + kk * 32), idesc_s,
                  kk != 0 ? 1u : 0u);
#pragma unroll
    for (int kk = 0; kk < kD / 16; ++kk)
      umma_f16_ss(tmem_dp, umma_desc_k128(do_addr + kk * 32), umma_desc_k128(v_addr + c * kc * 128 + kk * 32), idesc_s,
                  kk != 0 ? 1u : 0u);
    umma_commit(bar_mma);
  };

  // softmax((q . k) / sqrt(64)): exp2 with the scale and log2(e) folded into one constant
  const float cexp = 0.125f * 1.4426950408889634f;
  uint32_t mma_uses = 0;
  // ---------------------------------------------------------------- pass 1: m, l, D (online)
  float m = -CUDART_INF_F, l = 0.f, acc = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_s_dp(c);
    }
    mbar_wait(bar_mma, mma_uses & 1);
    ++mma_uses;
    tc_fence_after();
    for (int j0 = 0; j0 < kc; j0 += 32) {
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tmem_s + lane_off + j0, sv);
      tmem_ld_32x32(tmem_dp + lane_off + j0, dv);
      tmem_ld_wait();
      float mx = m;
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]) * cexp);
      const float corr = exp2f(m - mx);  // 0 on the first piece (m = -inf)
      l *= corr;
      acc *= corr;
      m = mx;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = exp2f(fmaf(__uint_as_float(sv[j]), cexp, -m));
        l += e;
        acc = fmaf(e, __uint_as_float(dv[j]), acc);
      }
    }
    tc_fence_before();
    __syncthreads();  // every row of both accumulators has been read
  }
  const float inv_l = 1.0f / l;
  const float Dr = acc * inv_l;
  // ---------------------------------------------------------------- pass 2: P, dS, dQ += dS K
  const int t = q0 + row;
  const size_t bh = static_cast<size_t>(b) * heads + h;
  uint16_t* Prow = p.P + (bh * L + t) * static_cast<size_t>(L);
  uint16_t* dSrow = p.dS + (bh * L + t) * static_cast<size_t>(L);
  for (int c = 0; c < nchunks; ++c) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_s_dp(c);
    }
    mbar_wait(bar_mma, mma_uses & 1);
    ++mma_uses;
    if (c > 0) mbar_wait(bar_dq, (c - 1) & 1);  // dS K of the previous chunk is done: the dS tile is free again
    tc_fence_after();
    for (int j0 = 0; j0 < kc; j0 += 32) {
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tmem_s + lane_off + j0, sv);
      tmem_ld_32x32(tmem_dp + lane_off + j0, dv);
      tmem_ld_wait();
      uint32_t pp[16], dd[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = exp2f(fmaf(__uint_as_float(sv[j]), cexp, -m)) * inv_l;
        const float p1 = exp2f(fmaf(__uint_as_float(sv[j + 1]), cexp, -m)) * inv_l;
        const float d0 = p0 * (__uint_as_float(dv[j]) - Dr) * 0.125f;
        const float d1 = p1 * (__uint_as_float(dv[j + 1]) - Dr) * 0.125f;
        pp[j >> 1] = pk2f(p0, p1, fmt);
        dd[j >> 1] = pk2f(d0, d1, fmt);
      }
      if (t < L) {
        uint4* pg = reinterpret_cast<uint4*>(Prow + c * kc + j0);
        uint4* dg = reinterpret_cast<uint4*>(dSrow + c * kc + j0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          pg[q] = make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
          dg[q] = make_uint4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]);
        }
      }
      // keys j0 .. j0+31 of this row into the K-major dS tile: four 16-byte chunks in atom j0 / 64, 128-byte swizzle
      const uint32_t atom = ds_addr + (j0 >> 6) * (kRows * 128) + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = ((j0 & 63) >> 3) + q;
        const uint32_t addr = atom + ((chunk ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(dd[4 * q]), "r"(dd[4 * q + 1]),
                     "r"(dd[4 * q + 2]), "r"(dd[4 * q + 3])
                     : "memory");
      }
    }
    fence_proxy_async_smem();  // dS was written through the generic proxy; the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      for (int j = 0; j < kc / 16; ++j)
        umma_f16_ss(tmem_dq, umma_desc_k128(ds_addr + (j >> 2) * (kRows * 128) + (j & 3) * 32),
                    umma_desc_k128(k_addr + (c * kc + j * 16) * 128), idesc_dq, (c | j) != 0 ? 1u : 0u);
      umma_commit(bar_dq);
    }
  }
  mbar_wait(bar_dq, (nchunks - 1) & 1);
  tc_fence_after();
  // ---------------------------------------------------------------- dQ -> global
#pragma unroll
  for (int c0 = 0; c0 < kD; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_dq + lane_off + c0, v);
    tmem_ld_wait();
    if (t < L) {
      uint4* op = reinterpret_cast<uint4*>(p.dqkv + (static_cast<size_t>(b) * L + t) * 3 * C + h * kD + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pk2f(__uint_as_float(v[i * 8 + 0]), __uint_as_float(v[i * 8 + 1]), fmt);
        o.y = pk2f(__uint_as_float(v[i * 8 + 2]), __uint_as_float(v[i * 8 + 3]), fmt);
        o.z = pk2f(__uint_as_float(v[i * 8 + 4]), __uint_as_float(v[i * 8 + 5]), fmt);
        o.w = pk2f(__uint_as_float(v[i * 8 + 6]), __uint_as_float(v[i * 8 + 7]), fmt);
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

struct BwdKVParams {
  CUtensorMap tmap_p;   // P  {L, B*heads*L}, box {64, kq}
  CUtensorMap tmap_ds;  // dS {L, B*heads*L}, box {64, kq}
  CUtensorMap tmap_q;   // qkv {3C, B*L}, box {64, kq}
  CUtensorMap tmap_do;  // dout {C, B*L}, box {64, kq}
  uint16_t* dqkv;
  int L, C, kq, fmt;
};

constexpr int kKVSmem = (2 * 2 + 2) * kMaxChunk * 128 + 1024 + 64;

__global__ void __launch_bounds__(kThreads, 1) attn_bwd_kv_tc_kernel(const __grid_constant__ BwdKVParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kq = p.kq;
  uint8_t* p_sm = smem;                       // two 64-key chunks of [kq rows x 128 B]
  uint8_t* ds_sm = p_sm + 2 * kMaxChunk * 128;
  uint8_t* do_sm = ds_sm + 2 * kMaxChunk * 128;  // [kq rows x 128 B]
  uint8_t* q_sm = do_sm + kMaxChunk * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_sm + kMaxChunk * 128);
  uint64_t* bar_load = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * kRows;
  const int L = p.L, C = p.C, fmt = p.fmt;
  const int heads = C / kD;
  const int nchunks = L / kq;
  const int row = threadIdx.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_p);
    tma_prefetch_desc(&p.tmap_ds);
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_dv = tmem_base, tmem_dk = tmem_base + kD;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  // M = 128 keys, N = 64 head channels, both operands MN-major (A: bit 15, B: bit 16)
  const uint32_t idesc = umma_idesc_f16(kRows, kD, static_cast<uint32_t>(fmt)) | (1u << 15) | (1u << 16);
  const uint32_t p_addr = smem_u32(p_sm), ds_addr = smem_u32(ds_sm), do_addr = smem_u32(do_sm), q_addr = smem_u32(q_sm);
  const uint32_t lbo = static_cast<uint32_t>(kq) * 128u;
  const int bh_row0 = (b * heads + h) * L;

  for (int qc = 0; qc < nchunks; ++qc) {
    if (threadIdx.x == 0) {
      // (the MMAs of the previous chunk have read the staged tiles: every thread waited for them below)
      mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(6 * kq * 128));
      tma_load_2d(p_sm, &p.tmap_p, bar_load, k0, bh_row0 + qc * kq);
      tma_load_2d(p_sm + kq * 128, &p.tmap_p, bar_load, k0 + 64, bh_row0 + qc * kq);
      tma_load_2d(ds_sm, &p.tmap_ds, bar_load, k0, bh_row0 + qc * kq);
      tma_load_2d(ds_sm + kq * 128, &p.tmap_ds, bar_load, k0 + 64, bh_row0 + qc * kq);
      tma_load_2d(do_sm, &p.tmap_do, bar_load, h * kD, b * L + qc * kq);
      tma_load_2d(q_sm, &p.tmap_q, bar_load, h * kD, b * L + qc * kq);
      mbar_wait(bar_load, qc & 1);
      tc_fence_after();
      for (int j = 0; j < kq / 16; ++j) {
        const uint32_t ko = static_cast<uint32_t>(j) * 16u * 128u;  // sixteen query rows further
        const uint32_t accf = (qc | j) != 0 ? 1u : 0u;
        umma_f16_ss(tmem_dv, desc_mn128(p_addr + ko, lbo), desc_mn128(do_addr + ko, lbo), idesc, accf);
        umma_f16_ss(tmem_dk, desc_mn128(ds_addr + ko, lbo), desc_mn128(q_addr + ko, lbo), idesc, accf);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, qc & 1);  // every thread follows the phases one by one
  }
  tc_fence_after();
  const int key = k0 + row;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int c0 = 0; c0 < kD; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32((which == 0 ? tmem_dv : tmem_dk) + lane_off + c0, v);
      tmem_ld_wait();
      if (key < L) {
        uint4* op = reinterpret_cast<uint4*>(p.dqkv + (static_cast<size_t>(b) * L + key) * 3 * C +
                                             (which == 0 ? 2 * C : C) + h * kD + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pk2f(__uint_as_float(v[i * 8 + 0]), __uint_as_float(v[i * 8 + 1]), fmt);
          o.y = pk2f(__uint_as_float(v[i * 8 + 2]), __uint_as_float(v[i * 8 + 3]), fmt);
          o.z = pk2f(__uint_as_float(v[i * 8 + 4]), __uint_as_float(v[i * 8 + 5]), fmt);
          o.w = pk2f(__uint_as_float(v[i * 8 + 6]), __uint_as_float(v[i * 8 + 7]), fmt);
          op[i] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace

bool attention_bwd_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDDPM_ATTN_BWD_TC");  // A/B switch for measurements: 0 = the mma.sync kernels
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int launch_attention_bwd_tc(const void* qkv, const void* dout, void* dqkv, void* scratch, int B, int L, int C, int fmt,
                            cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(attn_bwd_q_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQSmem));
    CDDPM_CUDA(cudaFuncSetAttribute(attn_bwd_kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKVSmem));
    attr_set = true;
  }
  const int heads = C / kD;
  const int kc = (L % 192 == 0) ? 192 : ((L % 128 == 0) ? 128 : 64);
  uint16_t* Pbuf = reinterpret_cast<uint16_t*>(scratch);
  uint16_t* dSbuf = Pbuf + static_cast<size_t>(B) * heads * L * L;
  {
    BwdQParams p;
    memset(&p, 0, sizeof(p));
    p.dqkv = reinterpret_cast<uint16_t*>(dqkv);
    p.P = Pbuf;
    p.dS = dSbuf;
    p.L = L;
    p.C = C;
    p.kc = kc;
    p.fmt = fmt;
    const uint64_t dims[2] = {static_cast<uint64_t>(3) * C, static_cast<uint64_t>(B) * L};
    const uint64_t strides[1] = {static_cast<uint64_t>(3) * C * 2};
    const uint32_t box_q[2] = {static_cast<uint32_t>(kD), static_cast<uint32_t>(kRows)};
    const uint32_t box_kv[2] = {static_cast<uint32_t>(kD), static_cast<uint32_t>(kc)};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_q, qkv, 2, dims, strides, box_q));
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_kv, qkv, 2, dims, strides, box_kv));
    const uint64_t ddims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(B) * L};
    const uint64_t dstrides[1] = {static_cast<uint64_t>(C) * 2};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_do, dout, 2, ddims, dstrides, box_q));
    dim3 grid((L + kRows - 1) / kRows, heads, B);
    attn_bwd_q_tc_kernel<<<grid, kThreads, kQSmem, stream>>>(p);
    CDDPM_TRY(check_launch("attn_bwd_q_tc_kernel"));
  }
  {
    BwdKVParams p;
    memset(&p, 0, sizeof(p));
    p.dqkv = reinterpret_cast<uint16_t*>(dqkv);
    p.L = L;
    p.C = C;
    p.kq = kc;
    p.fmt = fmt;
    const uint64_t sdims[2] = {static_cast<uint64_t>(L), static_cast<uint64_t>(B) * heads * L};
    const uint64_t sstrides[1] = {static_cast<uint64_t>(L) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kD), static_cast<uint32_t>(kc)};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_p, Pbuf, 2, sdims, sstrides, box));
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_ds, dSbuf, 2, sdims, sstrides, box));
    const uint64_t dims[2] = {static_cast<uint64_t>(3) * C, static_cast<uint64_t>(B) * L};
    const uint64_t strides[1] = {static_cast<uint64_t>(3) * C * 2};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_q, qkv, 2, dims, strides, box));
    const uint64_t ddims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(B) * L};
    const uint64_t dstrides[1] = {static_cast<uint64_t>(C) * 2};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_do, dout, 2, ddims, dstrides, box));
    dim3 grid((L + kRows - 1) / kRows, heads, B);
    attn_bwd_kv_tc_kernel<<<grid, kThreads, kKVSmem, stream>>>(p);
    CDDPM_TRY(check_launch("attn_bwd_kv_tc_kernel"));
  }
  return kOk;
}

}  // namespace cddpm
