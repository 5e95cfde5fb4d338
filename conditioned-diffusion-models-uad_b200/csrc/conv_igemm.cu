// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.  See conv_igemm.cuh for the contract.
//
// One persistent CTA per SM walks output tiles of 128 pixels x n_tile channels.
//   warp 0  (one lane)  TMA producer: per K step, one 4-D box load per spatial box of the tile (the 3x3 tap is a
//                       coordinate shift; out-of-image coordinates are zero-filled by TMA = the conv padding) plus
//                       one 2-D load of the weight slab, into a ring of 128B-swizzled stages.
//   warp 1  (one lane)  MMA issuer: 4 x tcgen05.mma (K=16 each) per stage into a TMEM accumulator; tcgen05.commit
//                       releases the stage and, after the last K step, publishes the accumulator.
//   warp 2              TMEM allocation / deallocation.
//   warps 4-7           epilogue: tcgen05.ld the accumulator (one TMEM lane = one pixel per thread), add bias and
//                       residual, round to the 16-bit activation type and store NHWC.  Two accumulator stages let
//                       the epilogue of tile i overlap the main loop of tile i+1.
#include "conv_igemm.cuh"

#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"

namespace cddpm {

namespace {

constexpr int kAStageBytes = kConvTileM * kConvBlockK * 2;  // 16 KiB
constexpr int kSmemBudget = 227 * 1024;
constexpr int kStatW = 128;                      // (n_tile / 32) chunks x 16 values, n_tile <= 256
constexpr int kStatBytes = 2 * 4 * kStatW * 4;   // two accumulator stages x four epilogue warps

struct TapShift {
  int dy, dx;
};

__device__ __forceinline__ uint32_t pack_16bit(float a, float b, int ab_format) {
  if (ab_format == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
__device__ __forceinline__ float2 unpack_16bit(uint32_t u, int ab_format) {
  if (ab_format == 1) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}

// kPair: launched as 2-CTA clusters; the pair issues tcgen05.mma.cta_group::2 with M = 256 (each CTA owns one
// 128-pixel M tile and stages half of the weight slab), which halves both the shared-memory operand reads per MMA and
// the weight bytes each SM pulls from L2.  Without it the SS-mode MMA is limited by shared-memory bandwidth.
template <bool kPair>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_stages = p.num_stages;
  const int b_stage_bytes = (kPair ? p.n_tile / 2 : p.n_tile) * kConvBlockK * 2;
  const int stage_bytes = kAStageBytes + b_stage_bytes;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  // work decomposition: a "work item" is one M tile (kPair: two adjacent M tiles, one per CTA) x one N tile
  const int work_stride = kPair ? gridDim.x / 2 : gridDim.x;
  const int work_first = kPair ? blockIdx.x / 2 : blockIdx.x;
  const int num_work = (kPair ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tfull_bar = empty_bar + num_stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // [2 accumulator stages][4 epilogue warps][kStatW] staging of GroupNorm partial sums
  float* stat_sh = reinterpret_cast<float*>(smem + num_stages * stage_bytes + 256);

  int num_ksteps = 0;
  for (int s = 0; s < p.num_src; ++s) num_ksteps += p.src_taps[s] * (p.src_c[s] / kConvBlockK);
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int box_px = p.box_w * p.box_h;
  const int boxes_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.tmap_a[s]);
    tma_prefetch_desc(&p.tmap_b);
  }
  if (kPair) cluster_sync_relaxed();  // both CTAs are resident before TMEM is allocated for the pair
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < num_stages; ++i) {
      mbar_init(&full_bar[i], kPair ? 2 : 1);  // pair: the leader's barrier collects both producers
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kPair ? 256 : 128);  // pair: both epilogues release the leader's accumulator stage
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) {
      tmem_alloc_pair(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    } else {
      tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    }
  }
  tc_fence_before();
  if (kPair) {
    __syncthreads();
    cluster_sync_relaxed();
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = work_first; work < num_work; work += work_stride) {
        const int m_tile = kPair ? 2 * (work / p.num_n_tiles) + static_cast<int>(cta_rank) : work / p.num_n_tiles;
        const int n_idx = work % p.num_n_tiles;
        int kstep = 0;
        for (int s = 0; s < p.num_src; ++s) {
          const int chunks = p.src_c[s] / kConvBlockK;
          const int taps = p.src_taps[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
            const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
            for (int ch = 0; ch < chunks; ++ch, ++kstep) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = smem + stage * stage_bytes;
              uint8_t* b_dst = a_dst + kAStageBytes;
              if (kPair) {
                // both producers arrive on the LEADER's barrier, which expects the bytes of both CTAs
                if (leader) {
                  mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(2 * stage_bytes));
                } else {
                  mbar_arrive_cluster(&full_bar[stage], 0);
                }
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
              }
              if (p.flat) {
                if (kPair) {
                  tma_load_2d_pair(a_dst, &p.tmap_a[s], &full_bar[stage], ch * kConvBlockK, m_tile * kConvTileM);
                } else {
                  tma_load_2d(a_dst, &p.tmap_a[s], &full_bar[stage], ch * kConvBlockK, m_tile * kConvTileM);
                }
              } else {
                for (int b = 0; b < p.boxes_per_tile; ++b) {
                  const int box = m_tile * p.boxes_per_tile + b;
                  const int n = box / boxes_per_img;
                  const int r = box - n * boxes_per_img;
                  const int ty = r / p.tiles_w;
                  const int tx = r - ty * p.tiles_w;
                  if (kPair) {
                    tma_load_4d_pair(a_dst + b * box_px * 128, &p.tmap_a[s], &full_bar[stage], ch * kConvBlockK,
                                     tx * p.box_w + dx, ty * p.box_h + dy, n);
                  } else {
                    tma_load_4d(a_dst + b * box_px * 128, &p.tmap_a[s], &full_bar[stage], ch * kConvBlockK,
                                tx * p.box_w + dx, ty * p.box_h + dy, n);
                  }
                }
              }
              if (kPair) {
                // this CTA's half of the weight rows of the N tile
                tma_load_2d_pair(b_dst, &p.tmap_b_half, &full_bar[stage], kstep * kConvBlockK,
                                 n_idx * p.n_tile + static_cast<int>(cta_rank) * (p.n_tile / 2));
              } else {
                tma_load_2d(b_dst, &p.tmap_b, &full_bar[stage], kstep * kConvBlockK, n_idx * p.n_tile);
              }
              if (++stage == num_stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (pair: leader CTA only)
    // All MMA operands are warp-uniform (shuffled bases, descriptors advanced by constants) and the issuing lane is
    // elected: the compiler then emits back-to-back UTCHMMA from uniform registers, ~45 cycles of issue per MMA
    // instead of ~160 with per-thread operands under `lane == 0` (tools/ubench_mma.cu).
    const uint32_t idesc = umma_idesc_f16(kPair ? 2 * kConvTileM : kConvTileM, static_cast<uint32_t>(p.n_tile),
                                          static_cast<uint32_t>(p.ab_format));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const bool elected = elect_one_sync();
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int work = work_first; work < num_work; work += work_stride, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + static_cast<uint32_t>(acc * p.n_tile);
      uint32_t accum = 0;
      for (int kstep = 0; kstep < num_ksteps; ++kstep) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elected) {
          const uint64_t ad = umma_desc_k128(smem_u + stage * stage_bytes);
          const uint64_t bd = umma_desc_k128(smem_u + stage * stage_bytes + kAStageBytes);
#pragma unroll
          for (int k = 0; k < kConvBlockK / 16; ++k) {
            if (kPair) {
              umma_f16_ss_pair(tmem_d, ad + 2 * k, bd + 2 * k, idesc, k == 0 ? accum : 1u);
            } else {
              umma_f16_ss(tmem_d, ad + 2 * k, bd + 2 * k, idesc, k == 0 ? accum : 1u);
            }
          }
          if (kPair) {
            umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs
            if (kstep == num_ksteps - 1) umma_commit_pair(&tfull_bar[acc]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kstep == num_ksteps - 1) umma_commit(&tfull_bar[acc]);
          }
        }
        accum = 1;
        if (++stage == num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    const int fmt = p.ab_format;
    int iter = 0;
    for (int work = work_first; work < num_work; work += work_stride, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int m_tile = kPair ? 2 * (work / p.num_n_tiles) + static_cast<int>(cta_rank) : work / p.num_n_tiles;
      const int n_idx = work % p.num_n_tiles;

      // pixel owned by this thread
      const int b = row / box_px;
      const int rr = row - b * box_px;
      const int box = m_tile * p.boxes_per_tile + b;
      const int n = box / boxes_per_img;
      const int r = box - n * boxes_per_img;
      const int ty = r / p.tiles_w;
      const int tx = r - ty * p.tiles_w;
      const int y = ty * p.box_h + rr / p.box_w;
      const int x = tx * p.box_w + rr % p.box_w;
      bool valid = (box < p.total_boxes) && (y < p.H) && (x < p.W);
      size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
      if (p.flat) {
        pix = static_cast<size_t>(m_tile) * kConvTileM + row;
        valid = pix < static_cast<size_t>(p.M);
      }
      const size_t roff = pix * p.Cout + static_cast<size_t>(n_idx) * p.n_tile;              // residual [.., Cout]
      const size_t off = pix * p.out_stride + p.out_col_off + static_cast<size_t>(n_idx) * p.n_tile;  // output row

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * p.n_tile);
      const int nchunks = p.n_tile / 32;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (c == nchunks - 1) {
          // every TMEM read of this accumulator stage has completed: hand it back to the MMA warp
          tc_fence_before();
          if (kPair) {
            mbar_arrive_cluster(&tempty_bar[acc], 0);  // the leader's MMA warp owns the accumulator hand-off
          } else {
            mbar_arrive(&tempty_bar[acc]);
          }
        }
        if (valid) {
          const int co = n_idx * p.n_tile + c * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + co + j));
              f[j] += bv.x;
              f[j + 1] += bv.y;
              f[j + 2] += bv.z;
              f[j + 3] += bv.w;
            }
          }
          if (p.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + roff + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 rv = __ldg(rp + q);
              const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = unpack_16bit(w[e], fmt);
                f[q * 8 + e * 2] += t.x;
                f[q * 8 + e * 2 + 1] += t.y;
              }
            }
          }
          if (p.relu == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (p.relu == 2) {  // SiLU (embedding MLPs)
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = f[j] / (1.0f + __expf(-f[j]));
          }
          if (p.gn_stats != nullptr) {
            // GroupNorm statistics of the tensor being produced, at 4-channel granularity: per thread 8 buckets of
            // (sum, sum of squares) over its pixel, then a butterfly over the warp's 32 pixels (16 shuffles) that leaves
            // value (lane >> 1) & 15 in each lane; the four epilogue warps meet in shared memory below.
            float v16[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float s = 0.f, q = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                s += f[4 * k + j];
                q = fmaf(f[4 * k + j], f[4 * k + j], q);
              }
              v16[k] = s;
              v16[8 + k] = q;
            }
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1) {
              const int m = w * 2;
              const bool hi = (lane & m) != 0;
#pragma unroll
              for (int k = 0; k < w; ++k) {
                const float keep = hi ? v16[k + w] : v16[k];
                const float send = hi ? v16[k] : v16[k + w];
                v16[k] = keep + __shfl_xor_sync(0xffffffffu, send, m);
              }
            }
            const float tot = v16[0] + __shfl_xor_sync(0xffffffffu, v16[0], 1);
            if ((lane & 1) == 0) stat_sh[(acc * 4 + quarter) * kStatW + c * 16 + (lane >> 1)] = tot;
          }
          if (p.out_is_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) op[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_16bit(f[q * 8 + 0], f[q * 8 + 1], fmt);
              o.y = pack_16bit(f[q * 8 + 2], f[q * 8 + 3], fmt);
              o.z = pack_16bit(f[q * 8 + 4], f[q * 8 + 5], fmt);
              o.w = pack_16bit(f[q * 8 + 6], f[q * 8 + 7], fmt);
              op[q] = o;
            }
          }
        }
      }
      if (p.gn_stats != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
        const int nvals = nchunks * 16;
        if (row < nvals) {
          const int within = row & 15;
          const int is_q = within >> 3;
          const int bucket = (row >> 4) * 8 + (within & 7);
          const float* sp = stat_sh + acc * 4 * kStatW + row;
          const int gb = n_idx * (p.n_tile >> 2) + bucket;
          const int nb4 = p.Cout >> 2;
          const int wpb = 4 / p.boxes_per_tile;  // warps per spatial box (4 or 2)
          for (int b2 = 0; b2 < p.boxes_per_tile; ++b2) {
            const int bx = m_tile * p.boxes_per_tile + b2;
            if (bx >= p.total_boxes) continue;
            const int nimg = bx / boxes_per_img;
            float t = 0.f;
            for (int w = 0; w < wpb; ++w) t += sp[(b2 * wpb + w) * kStatW];
            atomicAdd(&p.gn_stats[(static_cast<size_t>(nimg) * nb4 + gb) * 2 + is_q], static_cast<double>(t));
          }
        }
      }
    }
  }

  tc_fence_before();
  if (kPair) {
    __syncthreads();
    cluster_sync_relaxed();  // the peer may still be reading operands / TMEM that belong to the pair
  } else {
    __syncthreads();
  }
  if (warp == 2) {
    tc_fence_after();
    if (kPair) {
      tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    } else {
      tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    }
  }
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin_total, int ksize, int cin_off,
                                        int C_s, uint16_t* __restrict__ out, int Ktot, int koff, int ab_format) {
  const int taps = ksize * ksize;
  const size_t total = static_cast<size_t>(Cout) * taps * C_s;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % C_s);
    const int tap = static_cast<int>((i / C_s) % taps);
    const int co = static_cast<int>(i / (static_cast<size_t>(C_s) * taps));
    const float v = w[(static_cast<size_t>(co) * Cin_total + cin_off + ci) * taps + tap];
    uint16_t bits;
    if (ab_format == 1) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    } else {
      __half h = __float2half_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    }
    out[static_cast<size_t>(co) * Ktot + koff + tap * C_s + ci] = bits;
  }
}

}  // namespace

int conv_ktot(const ConvDesc& d) {
  int k = 0;
  for (int s = 0; s < d.num_src; ++s) k += d.src_taps[s] * d.src_c[s];
  return k;
}

static bool pick_box(int H, int W, int* bw, int* bh) {
  if (W % 16 == 0 && H % 8 == 0) {
    *bw = 16;
    *bh = 8;
    return true;
  }
  if (W % 8 == 0 && H % 8 == 0) {
    *bw = 8;
    *bh = 8;
    return true;
  }
  return false;
}

int conv_num_boxes(int B, int H, int W) {
  int bw, bh;
  if (!pick_box(H, W, &bw, &bh)) return 0;
  return B * (W / bw) * (H / bh);
}

int build_conv_params(const ConvDesc& d, ConvIgemmParams* p) {
  if (d.num_src < 1 || d.num_src > kConvMaxSrc) return fail(kInvalidArgument, "conv: num_src must be 1..3");
  memset(p, 0, sizeof(*p));
  const bool flat = d.flat_rows > 0;
  int bw = 16, bh = 8;
  p->num_src = d.num_src;
  p->Cout = d.Cout;
  p->relu = d.relu;
  p->out_stride = d.out_stride > 0 ? d.out_stride : d.Cout;
  p->out_col_off = d.out_col_off;
  if (flat) {
    if (d.num_src != 1 || d.src_taps[0] != 1) return fail(kInvalidArgument, "conv: flat GEMM takes one 1-tap source");
    p->flat = 1;
    p->M = d.flat_rows;
    p->B = p->H = p->W = 1;
    p->box_w = bw;
    p->box_h = bh;
    p->boxes_per_tile = 1;
    p->tiles_w = p->tiles_h = 1;
    p->num_m_tiles = (d.flat_rows + kConvTileM - 1) / kConvTileM;
    p->total_boxes = p->num_m_tiles;
  } else {
    if (d.B < 1 || d.H < 1 || d.W < 1) return fail(kInvalidArgument, "conv: empty problem");
    if (!pick_box(d.H, d.W, &bw, &bh))
      return fail(kUnsupported, "conv: H and W must be multiples of 8 for the tcgen05 implicit-GEMM path");
    p->B = d.B;
    p->H = d.H;
    p->W = d.W;
    p->box_w = bw;
    p->box_h = bh;
    p->boxes_per_tile = kConvTileM / (bw * bh);
    p->tiles_w = d.W / bw;
    p->tiles_h = d.H / bh;
    p->total_boxes = d.B * p->tiles_w * p->tiles_h;
    p->num_m_tiles = (p->total_boxes + p->boxes_per_tile - 1) / p->boxes_per_tile;
  }
  // N tile: the largest of 256/192/128/64/32 dividing Cout
  int n_tile = 0;
  const int cands[] = {256, 192, 128, 96, 64, 32};
  for (int c : cands) {
    if (d.Cout % c == 0) {
      n_tile = c;
      break;
    }
  }
  if (n_tile == 0) return fail(kUnsupported, "conv: Cout must be a multiple of 32");
  p->n_tile = n_tile;
  p->num_n_tiles = d.Cout / n_tile;
  int cols = 32;
  while (cols < 2 * n_tile) cols *= 2;
  p->tmem_cols = cols;
  p->pair = pair_enabled() ? 1 : 0;
  const int stage_bytes = kAStageBytes + (p->pair ? n_tile / 2 : n_tile) * kConvBlockK * 2;
  int stages = (kSmemBudget - 2048 - kStatBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(kUnsupported, "conv: not enough shared memory for a 2-stage pipeline");
  p->num_stages = stages;
  p->ab_format = d.ab_format;
  p->out_is_f32 = d.out_is_f32;
  p->bias = d.bias;
  p->residual = reinterpret_cast<const bf16*>(d.residual);
  p->out = d.out;
  p->gn_stats = flat ? nullptr : d.gn_stats;
  if (p->gn_stats != nullptr && (d.out_is_f32 || d.Cout % 32 != 0)) return fail(kUnsupported, "conv: gn_stats needs a 16-bit output");

  int ktot = 0;
  for (int s = 0; s < d.num_src; ++s) {
    if (d.src_c[s] % kConvBlockK != 0) return fail(kUnsupported, "conv: source channels must be a multiple of 64");
    if (d.src_taps[s] != 1 && d.src_taps[s] != 9) return fail(kInvalidArgument, "conv: taps must be 1 or 9");
    p->src_c[s] = d.src_c[s];
    p->src_taps[s] = d.src_taps[s];
    const uint64_t C = static_cast<uint64_t>(d.src_c[s]);
    if (flat) {
      const uint64_t dims[2] = {C, static_cast<uint64_t>(d.flat_rows)};
      const uint64_t strides[1] = {C * 2};
      const uint32_t box[2] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(kConvTileM)};
      CDDPM_TRY(encode_tmap_16bit(&p->tmap_a[s], d.src[s], 2, dims, strides, box));
    } else {
      const uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
      const uint64_t strides[3] = {C * 2, C * 2 * d.W, C * 2 * d.W * d.H};
      const uint32_t box[4] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), 1u};
      CDDPM_TRY(encode_tmap_16bit(&p->tmap_a[s], d.src[s], 4, dims, strides, box));
    }
    ktot += d.src_taps[s] * d.src_c[s];
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(d.Cout)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ktot) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(n_tile)};
    CDDPM_TRY(encode_tmap_16bit(&p->tmap_b, d.wpacked, 2, dims, strides, box));
    const uint32_t box_half[2] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(n_tile / 2)};
    CDDPM_TRY(encode_tmap_16bit(&p->tmap_b_half, d.wpacked, 2, dims, strides, box_half));
  }
  return kOk;
}

bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDDPM_CONV_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int launch_conv_igemm(const ConvIgemmParams& p, cudaStream_t stream, int max_ctas) {
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_set = true;
  }
  const int stage_bytes = kAStageBytes + (p.pair ? p.n_tile / 2 : p.n_tile) * kConvBlockK * 2;
  const int smem = p.num_stages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + kStatBytes;
  int sms = device_sm_count();
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  if (!p.pair) {
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = num_tiles < sms ? num_tiles : sms;
    conv_igemm_kernel<false><<<grid, kConvThreads, smem, stream>>>(p);
    return check_launch("conv_igemm_kernel");
  }
  // CTA pairs: a cluster of two CTAs per work item (two adjacent M tiles x one N tile), persistent over SM pairs
  const int num_work = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  const int pairs = num_work < sms / 2 ? num_work : sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CDDPM_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true>, p));
  return check_launch("conv_igemm_kernel<pair>");
}

int launch_pack_conv_weight(const float* w_oihw, int Cout, int Cin_total, int ksize, int cin_off, int C_s,
                            void* wpacked, int Ktot, int koff, int ab_format, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(Cout) * ksize * ksize * C_s;
  if (job_recorder() != nullptr) {
    ParamJob j = {kJobPack, Cout, Cin_total, ksize, cin_off, C_s, Ktot, koff, ab_format, w_oihw, nullptr, wpacked,
                  static_cast<long long>(total)};
    job_record(j);
    return kOk;
  }
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  pack_conv_weight_kernel<<<blocks, 256, 0, stream>>>(w_oihw, Cout, Cin_total, ksize, cin_off, C_s,
                                                      reinterpret_cast<uint16_t*>(wpacked), Ktot, koff, ab_format);
  return check_launch("pack_conv_weight_kernel");
}

}  // namespace cddpm
