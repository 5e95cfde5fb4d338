// Backward of the self-attention core (QKVAttention, OpenAI_Unet.py:457-476), bf16.
//
// FIRST version, kept as the A/B reference (CDDPM_ATTN_BWD_TC=0); the product path is attention_bwd_tc.cu (tcgen05).
// 0.2 % of the training step's FLOPs (one 24x24 attention block), so this version trades speed for simplicity:
// warp-level mma.sync tensor cores through nvcuda::wmma, probabilities P and score gradients dS materialised once in
// an L2-friendly bf16 scratch ([B*heads][L][L]).  A tcgen05 version is listed under "next" in DESIGN.md.
//   kernel 1 (per 32 queries, head, image): S = Q K^T, dP = dO V^T, P = softmax(S / 8), D = rowsum(P * dP),
//                                           dS = P * (dP - D) / 8 -> scratch; dQ = dS K
//   kernel 2 (per 64 keys, head, image):    dV = P^T dO, dK = dS^T Q
#include "attention.cuh"

#include <cuda_bf16.h>
#include <mma.h>

namespace cddpm {

namespace {

using namespace nvcuda;

constexpr int kD = 64;  // head dim
constexpr int kQB = 32; // queries per CTA (kernel 1)
constexpr int kKB = 64; // keys per CTA (kernel 2)

typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> FragARow;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> FragACol;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> FragBRow;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> FragBCol;
typedef wmma::fragment<wmma::accumulator, 16, 16, 16, float> FragC;

__global__ void __launch_bounds__(256, 1) attn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            __nv_bfloat16* __restrict__ dqkv,
                                                            __nv_bfloat16* __restrict__ Pbuf,
                                                            __nv_bfloat16* __restrict__ dSbuf, int L, int C) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int LP = L + 8;
  float* S = reinterpret_cast<float*>(smem_raw);           // [32][LP]
  float* dP = S + kQB * LP;                                // [32][LP]
  __nv_bfloat16* dS16 = reinterpret_cast<__nv_bfloat16*>(dP + kQB * LP);  // [32][LP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQB, h = blockIdx.y, b = blockIdx.z;
  const int heads = C / kD;
  const int ld = 3 * C;
  const __nv_bfloat16* Q = qkv + static_cast<size_t>(b) * L * ld + h * kD;
  const __nv_bfloat16* K = Q + C;
  const __nv_bfloat16* V = Q + 2 * C;
  const __nv_bfloat16* dO = dout + static_cast<size_t>(b) * L * C + h * kD;
  const size_t bh = static_cast<size_t>(b) * heads + h;

  // ---- S = Q K^T and dP = dO V^T, 2 x (L/16) tiles each
  const int ntn = L / 16;
  for (int t = warp; t < 2 * ntn; t += 8) {
    const int mi = t / ntn, nj = t - mi * ntn;
    FragC cs, cp;
    wmma::fill_fragment(cs, 0.f);
    wmma::fill_fragment(cp, 0.f);
#pragma unroll
    for (int kk = 0; kk < kD / 16; ++kk) {
      FragARow a;
      FragBCol bk;
      wmma::load_matrix_sync(a, Q + static_cast<size_t>(q0 + mi * 16) * ld + kk * 16, ld);
      wmma::load_matrix_sync(bk, K + static_cast<size_t>(nj * 16) * ld + kk * 16, ld);
      wmma::mma_sync(cs, a, bk, cs);
      wmma::load_matrix_sync(a, dO + static_cast<size_t>(q0 + mi * 16) * C + kk * 16, C);
      wmma::load_matrix_sync(bk, V + static_cast<size_t>(nj * 16) * ld + kk * 16, ld);
      wmma::mma_sync(cp, a, bk, cp);
    }
    wmma::store_matrix_sync(S + mi * 16 * LP + nj * 16, cs, LP, wmma::mem_row_major);
    wmma::store_matrix_sync(dP + mi * 16 * LP + nj * 16, cp, LP, wmma::mem_row_major);
  }
  __syncthreads();

  // ---- softmax rows and dS (4 rows per warp)
  const float scale = 0.125f;  // (64^-1/4)^2
  for (int r = warp * 4; r < warp * 4 + 4; ++r) {
    float* s = S + r * LP;
    const float* dp = dP + r * LP;
    float m = -3.0e38f;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, s[j] * scale);
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, w));
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = __expf(s[j] * scale - m);
      s[j] = e;
      sum += e;
    }
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, w);
    const float inv = 1.0f / sum;
    float dsum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float p = s[j] * inv;
      s[j] = p;
      dsum = fmaf(p, dp[j], dsum);
    }
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, w);
    __nv_bfloat16* prow = Pbuf + (bh * L + q0 + r) * L;
    __nv_bfloat16* drow = dSbuf + (bh * L + q0 + r) * L;
    for (int j = lane; j < L; j += 32) {
      const float p = s[j];
      const __nv_bfloat16 d16 = __float2bfloat16_rn(p * (dp[j] - dsum) * scale);
      prow[j] = __float2bfloat16_rn(p);
      drow[j] = d16;
      dS16[r * LP + j] = d16;
    }
  }
  __syncthreads();

  // ---- dQ = dS K : 2 x 4 tiles, one per warp
  {
    const int mi = warp >> 2, nj = warp & 3;
    FragC c;
    wmma::fill_fragment(c, 0.f);
    for (int kk = 0; kk < L / 16; ++kk) {
      FragARow a;
      FragBRow bk;
      wmma::load_matrix_sync(a, dS16 + mi * 16 * LP + kk * 16, LP);
      wmma::load_matrix_sync(bk, K + static_cast<size_t>(kk * 16) * ld + nj * 16, ld);
      wmma::mma_sync(c, a, bk, c);
    }
    float* scratch = S + warp * 256;  // S is dead
    wmma::store_matrix_sync(scratch, c, 16, wmma::mem_row_major);
    __syncwarp();
    __nv_bfloat16* dq = dqkv + static_cast<size_t>(b) * L * ld + h * kD;
    for (int i = lane; i < 256; i += 32) {
      const int rr = i >> 4, cc = i & 15;
      dq[static_cast<size_t>(q0 + mi * 16 + rr) * ld + nj * 16 + cc] = __float2bfloat16_rn(scratch[i]);
    }
  }
}

__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                          const __nv_bfloat16* __restrict__ dout,
                                                          __nv_bfloat16* __restrict__ dqkv,
                                                          const __nv_bfloat16* __restrict__ Pbuf,
                                                          const __nv_bfloat16* __restrict__ dSbuf, int L, int C) {
  __shared__ __align__(32) float scratch[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kKB, h = blockIdx.y, b = blockIdx.z;
  const int heads = C / kD;
  const int ld = 3 * C;
  const __nv_bfloat16* Q = qkv + static_cast<size_t>(b) * L * ld + h * kD;
  const __nv_bfloat16* dO = dout + static_cast<size_t>(b) * L * C + h * kD;
  const size_t bh = static_cast<size_t>(b) * heads + h;
  const __nv_bfloat16* P = Pbuf + bh * L * L;
  const __nv_bfloat16* dS = dSbuf + bh * L * L;
  // 32 output tiles: [dV | dK] x 4 key tiles x 4 dim tiles; warp w owns tiles w, w + 8, w + 16, w + 24
  for (int t = warp; t < 32; t += 8) {
    const int which = t >> 4;  // 0: dV = P^T dO, 1: dK = dS^T Q
    const int mi = (t >> 2) & 3, nj = t & 3;
    const __nv_bfloat16* A = which == 0 ? P : dS;
    const __nv_bfloat16* Bm = which == 0 ? dO : Q;
    const int ldb = which == 0 ? C : ld;
    FragC c;
    wmma::fill_fragment(c, 0.f);
    for (int kk = 0; kk < L / 16; ++kk) {
      FragACol a;  // A^T[key][q] = A[q][key]
      FragBRow bk;
      wmma::load_matrix_sync(a, A + static_cast<size_t>(kk * 16) * L + k0 + mi * 16, L);
      wmma::load_matrix_sync(bk, Bm + static_cast<size_t>(kk * 16) * ldb + nj * 16, ldb);
      wmma::mma_sync(c, a, bk, c);
    }
    wmma::store_matrix_sync(scratch[warp], c, 16, wmma::mem_row_major);
    __syncwarp();
    // v gradient at channel offset 2C, k gradient at C
    __nv_bfloat16* dst = dqkv + static_cast<size_t>(b) * L * ld + (which == 0 ? 2 * C : C) + h * kD;
    for (int i = lane; i < 256; i += 32) {
      const int rr = i >> 4, cc = i & 15;
      dst[static_cast<size_t>(k0 + mi * 16 + rr) * ld + nj * 16 + cc] = __float2bfloat16_rn(scratch[warp][i]);
    }
    __syncwarp();
  }
}

}  // namespace

int64_t attention_bwd_scratch_elems(int B, int L, int C) { return 2ll * B * (C / kD) * L * L; }

int launch_attention_bwd(const void* qkv, const void* dout, void* dqkv, void* scratch, int B, int L, int C, int fmt,
                         cudaStream_t stream) {
  if (!qkv || !dout || !dqkv || !scratch) return fail(kInvalidArgument, "attention_bwd: null pointer");
  if (fmt != 1) return fail(kUnsupported, "attention_bwd: the training path is bf16");
  if (C % kD != 0 || L % 64 != 0 || L > 576) return fail(kUnsupported, "attention_bwd: needs L % 64 == 0, L <= 576");
  if (attention_bwd_tc_enabled()) return launch_attention_bwd_tc(qkv, dout, dqkv, scratch, B, L, C, fmt, stream);
  const int heads = C / kD;
  __nv_bfloat16* Pbuf = reinterpret_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* dSbuf = Pbuf + static_cast<size_t>(B) * heads * L * L;
  const int LP = L + 8;
  const size_t smem = static_cast<size_t>(kQB) * LP * (4 + 4 + 2);
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kQB * (576 + 8) * 10));
    attr_set = true;
  }
  attn_bwd_q_kernel<<<dim3(L / kQB, heads, B), 256, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout),
      reinterpret_cast<__nv_bfloat16*>(dqkv), Pbuf, dSbuf, L, C);
  CDDPM_TRY(check_launch("attn_bwd_q_kernel"));
  attn_bwd_kv_kernel<<<dim3(L / kKB, heads, B), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout),
      reinterpret_cast<__nv_bfloat16*>(dqkv), Pbuf, dSbuf, L, C);
  return check_launch("attn_bwd_kv_kernel");
}

}  // namespace cddpm
