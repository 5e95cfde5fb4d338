// Micro-measurements behind the convolution kernel's design (development aid, not part of the library):
//   1. how many cycles one thread needs to issue tcgen05.mma + tcgen05.commit, per code pattern and MMA shape;
//   2. whether a 128B-swizzled K-major operand may be read from a start address that is offset by whole 128-byte
//      rows (what a 3x3 convolution needs to reuse ONE staged halo tile for all nine taps).
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conditioned-diffusion-models-uad_b200/csrc \
//        tools/ubench_mma.cu -o gpurun_out/ubench_mma && gpurun_out/ubench_mma
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace cddpm;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ bool elect_one_unused() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. issue cost with warp-uniform operands (shuffled TMEM base, descriptor = base + 2*k, elected lane).
//    kWarps issuing warps, each with its own accumulators and barriers; kAccs accumulators alternated per MMA;
//    kCommit: one tcgen05.commit per group; kPair: cta_group::2 (M = 256) issued by the leader CTA of a 2-CTA cluster.
// ---------------------------------------------------------------------------------------------------------------
template <bool kCommit, int kAccs, int kWarps, bool kPair, int kPerGroup>
__global__ void __launch_bounds__(128, 1) issue_kernel(int n_tile, int groups, int stages, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[2][8];
  __shared__ uint64_t done_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = 16384 + n_tile * 128;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 2; ++w) {
      for (int i = 0; i < 8; ++i) mbar_init(&bars[w][i], 1);
      mbar_init(&done_bar[w], 1);
    }
    fence_mbar_init();
  }
  if (kPair) cluster_sync_all();
  if (warp == 0) {
    if (kPair) tmem_alloc_pair(&tmem_slot, 512); else tmem_alloc(&tmem_slot, 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = umma_idesc_f16(kPair ? 256 : 128, static_cast<uint32_t>(n_tile), 0);
  if (warp >= 1 && warp <= kWarps && rank == 0) {
    const int w = warp - 1;
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0) + w * 256;
    const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const bool elected = elect_one_sync();
    const long long t0 = clock64();
    int stage = 0;
    for (int g = 0; g < groups; ++g) {
      const uint64_t a_desc = umma_desc_k128(smem_base + stage * stage_bytes);
      const uint64_t b_desc = umma_desc_k128(smem_base + stage * stage_bytes + 16384);
      if (elected) {
#pragma unroll
        for (int k = 0; k < kPerGroup; ++k) {
          const uint32_t d = tmem_base + (kAccs > 1 ? (k % kAccs) * 128 : 0);
          if (kPair)
            umma_f16_ss_pair(d, a_desc + 2 * (k & 3), b_desc + 2 * (k & 3), idesc, g != 0 ? 1u : 0u);
          else
            umma_f16_ss(d, a_desc + 2 * (k & 3), b_desc + 2 * (k & 3), idesc, g != 0 ? 1u : 0u);
        }
        if (kCommit) {
          if (kPair) umma_commit_pair(&bars[w][stage]); else umma_commit(&bars[w][stage]);
        }
      }
      if (++stage == stages) stage = 0;
    }
    const long long t1 = clock64();
    if (elected) {
      if (kPair) umma_commit_pair(&done_bar[w]); else umma_commit(&done_bar[w]);
    }
    mbar_wait(&done_bar[w], 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out[2 * w + 0] = static_cast<unsigned long long>(t1 - t0);
      out[2 * w + 1] = static_cast<unsigned long long>(t2 - t0);
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (kPair) tmem_dealloc_pair(tmem_slot, 512); else tmem_dealloc(tmem_slot, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. row-shifted / strided operand reads.  A holds 400 rows x 64 halves in the TMA SWIZZLE_128B layout,
//    A[r][c] = (r % 32)*64 + c; B is the 64 x 64 identity.  With a group stride of `group_rows` rows (SBO =
//    group_rows * 128 B) MMA row m = 8g + j must read A row shift + g*group_rows + j, so D[m][n] = A[that row][n].
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) shift_kernel(int shift, int group_rows, int use_base_offset, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* a_sm = smem;              // 400 rows x 128 B
  uint8_t* b_sm = smem + 400 * 128;  // 64 rows x 128 B
  for (int i = threadIdx.x; i < 400 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c / 8) ^ (r % 8)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__half*>(a_sm + off) = __float2half(static_cast<float>((r % 32) * 64 + c));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c / 8) ^ (r % 8)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__half*>(b_sm + off) = __float2half(r == c ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_f16(128, 64, 0);
    const uint32_t a_addr = smem_u32(a_sm) + shift * 128;
    const uint32_t b_addr = smem_u32(b_sm);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = umma_desc_k128_sbo(a_addr + k * 32, static_cast<uint32_t>(group_rows) * 128u);
      if (use_base_offset) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      umma_f16_ss(tmem_base, ad, umma_desc_k128(b_addr + k * 32), idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------------------------------------------------------
// 3. MN-major B operand (what P x V of attention needs: V is stored [key][head_dim], i.e. N contiguous).
//    A = P [128 x 64 keys] K-major SW128; B = V [64 keys][64 dims] as TMA SWIZZLE_128B would stage it (one 128-byte row
//    per key); D = P x V.  The instruction descriptor's B-major bit (16) is set; (lbo, sbo) are swept by the host.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mnmajor_kernel(int lbo, int sbo, int kstep_bytes, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* a_sm = smem;              // P: 128 rows x 128 B
  uint8_t* b_sm = smem + 128 * 128;  // V: 64 keys x 128 B
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c / 8) ^ (r % 8)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__half*>(a_sm + off) = __float2half(static_cast<float>((r + c) % 3));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int k = i / 64, d = i % 64;
    const int off = k * 128 + (((d / 8) ^ (k % 8)) * 16) + (d % 8) * 2;
    *reinterpret_cast<__half*>(b_sm + off) = __float2half(static_cast<float>((k * 7 + d) % 5));
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_f16(128, 64, 0) | (1u << 16);
    for (int k = 0; k < 4; ++k) {
      uint64_t bd = 0;
      const uint32_t b_addr = smem_u32(b_sm) + k * kstep_bytes;
      bd |= static_cast<uint64_t>((b_addr & 0x3FFFF) >> 4);
      bd |= static_cast<uint64_t>(lbo >> 4) << 16;
      bd |= static_cast<uint64_t>(sbo >> 4) << 32;
      bd |= static_cast<uint64_t>(1) << 46;
      bd |= static_cast<uint64_t>(2) << 61;
      umma_f16_ss(tmem_base, umma_desc_k128(smem_u32(a_sm) + k * 32), bd, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

template <bool kCommit, int kAccs, int kWarps, bool kPair, int kPerGroup>
static void run_issue(int n_tile, int grid, unsigned long long* d_out) {
  const int groups = 2000, stages = 4;
  const int smem = stages * (16384 + n_tile * 128) + 1024;
  auto kern = issue_kernel<kCommit, kAccs, kWarps, kPair, kPerGroup>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, n_tile, groups, stages, d_out));
  CK(cudaDeviceSynchronize());
  unsigned long long h[4];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  const double exec = 128.0 * n_tile * 16 / 4096.0;  // per SM; a pair MMA keeps both SMs busy this long
  printf("N=%3d grid=%3d pair=%d warps=%d accs=%d mmas/group=%d commit=%d : issue %.1f cyc/group (%.1f/mma)  drained %.1f"
         " cyc/group  [tensor time %.0f cyc/group/warp]\n",
         n_tile, grid, int(kPair), kWarps, kAccs, kPerGroup, int(kCommit), double(h[0]) / groups,
         double(h[0]) / groups / kPerGroup, double(h[1]) / groups, exec * kPerGroup);
}

int main() {
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  for (int n : {32, 64, 128, 192, 256}) {
    run_issue<false, 1, 1, false, 4>(n, 148, d_out);
    run_issue<true, 1, 1, false, 4>(n, 148, d_out);
  }
  for (int n : {64, 128}) {
    run_issue<false, 2, 1, false, 4>(n, 148, d_out);
    run_issue<true, 2, 1, false, 4>(n, 148, d_out);
    run_issue<false, 1, 2, false, 4>(n, 148, d_out);
    run_issue<true, 1, 2, false, 4>(n, 148, d_out);
    run_issue<true, 1, 1, false, 8>(n, 148, d_out);
    run_issue<true, 1, 1, false, 2>(n, 148, d_out);
    run_issue<true, 1, 1, false, 1>(n, 148, d_out);
  }
  for (int n : {64, 128, 256}) {
    run_issue<false, 1, 1, true, 4>(n, 148, d_out);
    run_issue<true, 1, 1, true, 4>(n, 148, d_out);
    run_issue<true, 2, 1, true, 4>(n, 148, d_out);
  }

  float* d_f;
  CK(cudaMalloc(&d_f, 128 * 64 * sizeof(float)));
  std::vector<float> h(128 * 64);
  const int smem = 400 * 128 + 64 * 128 + 1024;
  CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int ubo = cfg == 2 ? 1 : 0;
    const int group_rows = cfg == 1 ? 18 : 8;
    for (int shift : {0, 1, 2, 3, 7, 8, 9, 17, 18, 19, 37, 98}) {
      if (shift + 15 * group_rows + 8 > 400) continue;
      shift_kernel<<<1, 128, smem>>>(shift, group_rows, ubo, d_f);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_f, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          const float want = static_cast<float>(((shift + (m / 8) * group_rows + m % 8) % 32) * 64 + n);
          if (h[m * 64 + n] != want) ++bad;
        }
      printf("shift %3d rows, group stride %2d rows, base_offset field %s: %s (%d of 8192 wrong)\n", shift, group_rows,
             ubo ? "set" : "zero", bad ? "MISMATCH" : "exact", bad);
    }
  }
  {
    const int smem2 = 128 * 128 + 64 * 128 + 1024;
    const int combos[][3] = {{0, 1024, 2048}, {1024, 1024, 2048}, {128, 1024, 2048}, {2048, 1024, 2048},
                             {1024, 2048, 2048}, {0, 2048, 2048}, {1024, 128, 2048}, {8192, 1024, 2048}};
    for (const auto& c : combos) {
      mnmajor_kernel<<<1, 128, smem2>>>(c[0], c[1], c[2], d_f);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_f, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float want = 0.f;
          for (int k = 0; k < 64; ++k) want += static_cast<float>((m + k) % 3) * static_cast<float>((k * 7 + n) % 5);
          if (h[m * 64 + n] != want) ++bad;
        }
      printf("MN-major B: lbo %5d sbo %5d k-step %5d B: %s (%d of 8192 wrong)\n", c[0], c[1], c[2],
             bad ? "MISMATCH" : "exact", bad);
    }
  }
  return 0;
}
