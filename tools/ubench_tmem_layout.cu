// Where do the rows of a tcgen05.mma accumulator land in tensor memory?  (development aid, not part of the library)
//
// The convolution kernel uses cta_group::2 with M = 256: each CTA of the pair owns 128 rows, row i in TMEM lane i.
// A 64-pixel M tile (what the 24 x 24 level needs: nine 8 x 8 tiles per image) would be issued as
// cta_group::2 with M = 128 (64 rows per CTA) or as cta_group::1 with M = 64.  This program issues one such MMA with
// A[m][0] = m + 1, B[n][0] = 1 (every accumulator element of row m becomes m + 1) and a second one with A[m][0] = 1,
// B[n][0] = n + 1 (element (m, n) becomes n + 1), then dumps all 128 lanes x N columns of each CTA's tensor memory.
//
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conditioned-diffusion-models-uad_b200/csrc \
//        tools/ubench_tmem_layout.cu -o gpurun_out/ubench_tmem_layout && gpurun_out/ubench_tmem_layout
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace cddpm;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

constexpr int kN = 64;  // accumulator columns

// element (row r, k) of a K-major bf16 tile stored as 128-byte rows with the 128-byte swizzle
__device__ __forceinline__ void put(uint8_t* tile, int r, int k, float v) {
  const int byte = k * 2;
  const int chunk = (byte >> 4) ^ (r & 7);
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(tile + r * 128 + chunk * 16 + (byte & 15)) = h;
}

// kPair: cta_group::2 with M = m_total over the two CTAs; else cta_group::1 with M = m_total in one CTA.
// mode 0: accumulator = global row index + 1; mode 1: accumulator = column index + 1.
template <bool kPair>
__global__ void __launch_bounds__(128, 1) layout_kernel(int m_total, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tile = smem;                // up to 128 rows x 128 B
  uint8_t* b_tile = smem + 128 * 128;    // up to 64 rows x 128 B
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int rows_here = kPair ? m_total / 2 : m_total;
  const int n_here = kPair ? kN / 2 : kN;
  for (int i = threadIdx.x; i < (128 * 128 + 64 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < rows_here; r += blockDim.x)
    put(a_tile, r, 0, mode == 0 ? static_cast<float>(rank * rows_here + r + 1) : 1.0f);
  for (int n = threadIdx.x; n < n_here; n += blockDim.x)
    put(b_tile, n, 0, mode == 0 ? 1.0f : static_cast<float>(rank * n_here + n + 1));
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  // generic-proxy writes of the operand tiles -> visible to the tensor core's async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (kPair) cluster_sync_all();
  if (warp == 1) {
    if (kPair) {
      tmem_alloc_pair(&tmem_slot, kN);
    } else {
      tmem_alloc(&tmem_slot, kN);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // zero the whole accumulator window first (lanes the MMA does not write keep the marker -1)
  {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(-1.0f);
    for (int c0 = 0; c0 < kN; c0 += 32) {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
          "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
          "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
          "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(m_total), kN, 1u);
    const uint64_t ad = umma_desc_k128(smem_u32(a_tile)), bd = umma_desc_k128(smem_u32(b_tile));
    if (kPair) {
      umma_f16_ss_pair(tmem_base, ad, bd, idesc, 0u);
      umma_commit_pair(&done);
    } else {
      umma_f16_ss(tmem_base, ad, bd, idesc, 0u);
      umma_commit(&done);
    }
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    float* o = out + (static_cast<size_t>(rank) * 128 + threadIdx.x) * kN + c0;
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) {
      tmem_dealloc_pair(tmem_base, kN);
    } else {
      tmem_dealloc(tmem_base, kN);
    }
  }
}

template <bool kPair>
static void run(int m_total, const char* what) {
  const int ctas = kPair ? 2 : 1;
  float* d = nullptr;
  CK(cudaMalloc(&d, sizeof(float) * ctas * 128 * kN));
  std::vector<float> h(static_cast<size_t>(ctas) * 128 * kN);
  const int smem = 128 * 128 + 64 * 128 + 1024;
  CK(cudaFuncSetAttribute(layout_kernel<kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(d, 0, sizeof(float) * ctas * 128 * kN));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, layout_kernel<kPair>, m_total, mode, d));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d, sizeof(float) * h.size(), cudaMemcpyDeviceToHost));
    printf("%s, %s:\n", what, mode == 0 ? "value = row + 1 (column 0 and column N-1 of every lane)" : "value = column + 1 (lanes 0, 16, 32, 64, 96)");
    for (int c = 0; c < ctas; ++c) {
      if (mode == 0) {
        printf("  cta %d lanes 0..127, column 0:", c);
        for (int l = 0; l < 128; ++l) printf(" %g", h[(static_cast<size_t>(c) * 128 + l) * kN]);
        printf("\n  cta %d lanes 0..127, column %d:", c, kN - 1);
        for (int l = 0; l < 128; ++l) printf(" %g", h[(static_cast<size_t>(c) * 128 + l) * kN + kN - 1]);
        printf("\n");
      } else {
        for (int lane : {0, 16, 32, 64, 96}) {
          printf("  cta %d lane %d, columns 0..%d:", c, lane, kN - 1);
          for (int j = 0; j < kN; ++j) printf(" %g", h[(static_cast<size_t>(c) * 128 + lane) * kN + j]);
          printf("\n");
        }
      }
    }
  }
  CK(cudaFree(d));
}

int main() {
  run<true>(256, "cta_group::2, M = 256 (the convolution kernel's shape)");
  run<true>(128, "cta_group::2, M = 128 (64 rows per CTA)");
  run<false>(128, "cta_group::1, M = 128");
  run<false>(64, "cta_group::1, M = 64");
  return 0;
}
