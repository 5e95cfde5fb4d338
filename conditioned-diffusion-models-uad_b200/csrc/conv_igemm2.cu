// Second-generation tcgen05 implicit-GEMM convolution: one staged halo tile serves all nine taps, CTA pairs share
// the weight slab.
//
// Measured on B200 (profiles/r01_*, tools/ubench_mma.cu): the first kernel (conv_igemm.cu) streams a 16 KB activation
// tile per tap and 64-channel chunk - every activation byte crosses L2->SM nine times - and sits on the chip-wide
// L2->SM limit (~6300 B/cycle, ~43 B/cycle/SM) at 35-50 % tensor-pipe utilisation.  Two measured facts of the UMMA
// shared-memory descriptor remove that traffic:
//   * the 128-byte swizzle is applied to the absolute shared-memory address of each operand row, so the descriptor's
//     start address may point at ANY 128-byte row of a staged tile, and
//   * the stride between 8-row groups (SBO) may be any multiple of 128 bytes.
// So a CTA stages, per 64-channel chunk, ONE TMA box of 18 x 18 pixels (a 16 x 16 output macro tile plus its 3x3 halo,
// zero-filled outside the image by TMA) and issues the nine taps as MMAs whose A descriptors differ only in their
// start row: M tile m (8 pixels wide, 16 high) and tap (dy, dx) read rows starting at halo pixel (dy, dx + 8 m), one
// 8-pixel group per image row, group stride = the halo row pitch (18 x 128 B).  Activation traffic drops from 9x to
// 1.27x.
//
// A single-CTA N = 128 MMA reads 8 KB of operands per 64 tensor cycles - all of the SM's 128 B/cycle shared-memory
// bandwidth, so every TMA write stalls the tensor pipe (measured 56-66 % active).  The kernel therefore runs as 2-CTA
// clusters issuing tcgen05.mma.cta_group::2 (M = 256: one M tile of each CTA; N = 128 with each CTA staging 64 of the
// weight rows): 6 KB of operand reads per 64 cycles and half the weight bytes per SM.
//
// Per CTA and 64-channel chunk (256 px x 128 ch x 9 taps = 4608 tensor cycles): 41.5 KB activations + 74 KB weights,
// ~25 B/cycle/SM from L2.  Weight slabs travel in stages of three taps (one kernel row), so the issuing thread waits
// on one barrier and commits once per 24 MMAs.
//
// Where the remaining time goes (ncu --set full at HEAD, 128->128 @96x96, 65 us, tensor pipe 81 % active): the
// tensor core's operand reads are l1tex__data_pipe_tc_wavefronts_mem_shared = 53 % of that pipe's peak (the LSU adds
// 9 %, the TMA writes show as l1tex__data_bank_writes = 10 %) - shared memory is the second-busiest unit, not a
// saturated one (an earlier version of this comment claimed ~95 % from a byte budget; the counter does not support
// it).  What separates this launch from the 95 %-active 256->256 ones (222 us) is fixed cost: launch, TMEM allocation,
// pipeline fill and the last tiles' epilogues are ~8 us of a 65 us kernel, and at K = 1152 the epilogue paces the
// MMA stream.  (Folding the preceding GroupNorm+SiLU into the staged tiles was built and measured: an extra LSU pass
// over every tile in shared memory cost more conv time than the HBM pass it saves.)
// Warp roles: warp 0 weight producer, warp 3 halo-tile producer (its own thread, so tile loads run ahead by the
// depth of the tile ring), warp 1 MMA issuer (leader CTA only, warp-uniform operands), warp 2 TMEM allocator,
// warps 4-11 epilogue (one warpgroup per M tile: bias, residual, activation, GroupNorm partial sums, 16-bit NHWC
// store); two TMEM accumulator stages of 2 x 128 columns.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "conv_igemm.cuh"
#include "ptx.cuh"

namespace cddpm {

namespace {

constexpr int kThreads2 = 384;
constexpr int kTileH = 16;                // rows of a macro tile (= the 16 row groups of an M tile)
constexpr int kHaloH = kTileH + 2;
constexpr int kNTile = 128;               // output channels per work item
constexpr int kRowBytes = 128;            // 64 channels x 16 bit
constexpr int kBTapBytes = (kNTile / 2) * kRowBytes;    // 8192: this CTA's 64 weight rows of one tap
constexpr int kTapsPerStage = 3;
constexpr int kBSlotBytes = kTapsPerStage * kBTapBytes; // 24576
constexpr int kStagesB = 4;
constexpr int kStatW2 = 64;               // (128 / 32) chunks x 16 values per warp
constexpr int kTmemCols = 512;            // 2 stages x (up to) 2 M tiles x 128 columns

// Geometry for kMT M tiles (each 8 pixels wide x 16 high) side by side per CTA: 2 where the image width is a multiple
// of 16, 1 for the 24 x 24 level (rows past the image bottom are computed on zero padding and masked on the way out).
//
// kIL (image-interleaved tiles, kMT = 1 only; the 24 x 24 level): an M tile is 8 image rows of TWO consecutive images,
// 8 pixels wide - its 16 row groups alternate between the images (group 2 j = row j of image 2 n, group 2 j + 1 = row j
// of image 2 n + 1).  One TMA box over the tensor seen as {C, W, B, H} (image index before the row index) lands in
// shared memory in exactly that order, so the group stride stays the halo row pitch and tap (dy, dx) still is a start
// address: (2 dy x halo width + dx) rows.  A 24-row image then is three full tiles instead of one full and one
// half-empty 16-row tile: 4.5 instead of 6 M tiles per image (144 instead of 192 work items at B = 32: two rounds on
// 74 CTA pairs instead of three).  With kMT = 2 a CTA owns two such tiles (consecutive in the tile list, each its own
// box) and every weight stage serves both (CDDPM_CONV_IL=2; measured no faster than one tile per CTA, see build_conv2).
template <int kMT, bool kIL = false>
struct Geo {
  static constexpr int kTileW = kIL ? 8 : 8 * kMT;                 // macro tile (output pixels) of one TMA box
  static constexpr int kHaloW = kTileW + 2;
  static constexpr int kHaloPitch = kHaloW * kRowBytes;            // bytes between the 8-pixel row groups of an M tile
  static constexpr int kImgs = kIL ? 2 : 1;                        // images whose rows alternate inside a tile
  static constexpr int kTileRows = kTileH / kImgs;                 // image rows of a tile
  static constexpr int kBoxes = kIL ? kMT : 1;                     // TMA boxes per chunk (interleaved: one per M tile)
  static constexpr int kBoxBytes = kHaloW * (kTileRows + 2) * kImgs * kRowBytes;  // one staged halo tile (41472 B for kMT = 2)
  static constexpr int kBareBytes = kTileW * kTileH * kRowBytes;   // the same without halo (1x1 sources)
  static constexpr int kBoxPitch = (kBoxBytes + 1023) & ~1023;     // boxes stay 1 KB aligned (swizzle atom)
  static constexpr int kASlotBytes = kBoxes * kBoxPitch;
  // descriptor units (16 B) between the M tiles of a CTA: 8 pixels of the shared halo tile, or the next box
  static constexpr int kMStep = kIL ? kBoxPitch / 16 : 8 * (kRowBytes / 16);
  static constexpr int kStagesA = (kIL && kMT == 2) ? 2 : 3;       // 2 x 51 KB + the weight ring fill the 227 KB
  static constexpr int kEpiWarps = 4 * kMT;
  static constexpr int kStatW = kStatW2 * kImgs;                   // per warp: (images x) 4 chunks x 16 values
  static constexpr int kStatBytes = 2 * kEpiWarps * kStatW * 4;
  static constexpr int kCoefBytes = 2 * kNTile * 4;               // fused GroupNorm finish: A[128], B[128]
  static constexpr int kSmemBytes = kStagesA * kASlotBytes + kStagesB * kBSlotBytes + 256 + kStatBytes + kCoefBytes + 1024;
};

constexpr int kMaxChunks = 48;  // 64-channel chunks of one work item over all sources (K <= 3072 per source set)
struct Conv2Params {
  CUtensorMap tmap_a[kConvMaxSrc];  // {C, W, H, B}; box {64, 18, 18, 1} (3x3 sources) or {64, 16, 16, 1} (1x1 sources);
                                    // image-interleaved tiles: {C, W, B, H}; box {64, 10, 2, 10} or {64, 8, 2, 8}
  CUtensorMap tmap_b;               // {Ktot, Cout}; box {64, 64}
  int num_src;
  int src_c[kConvMaxSrc];
  int src_taps[kConvMaxSrc];
  int src_koff[kConvMaxSrc];        // first K column of each source in the packed weight panel
  // The order in which the 64-channel chunks of a work item are staged and multiplied: all three roles (weight
  // producer, tile producer, MMA issuer) walk this list.  1x1 chunks (8 MMAs of tensor time for a whole tile of
  // traffic) are spread between the 3x3 chunks so the tile ring never drains on them.
  int num_chunks;
  uint8_t chunk_src[kMaxChunks];
  uint8_t chunk_ch[kMaxChunks];
  int box1;  // 1x1 sources stage the bare 16-row tile (no halo)
  int B, H, W, Cout;
  int tiles_w, tiles_h;  // macro tiles per image (per image PAIR with interleaved tiles)
  int il;                // image-interleaved tiles (kernel template argument kIL)
  int num_il_tiles;      // interleaved M tiles of the launch (num_m_tiles counts CTA tiles of `mt` of them)
  int num_m_tiles, num_n_tiles;
  int mt;  // M tiles per CTA (kernel template argument)
  int ab_format, relu;
  const float* bias;
  const uint16_t* residual;
  uint16_t* out;
  double* gn_stats;
  // fused GroupNorm finish (ConvDesc::gn_gamma): see the epilogue
  const float* gn_gamma;
  const float* gn_beta;
  const float* gn_film;
  int gn_film_stride, gn_film_off;
  unsigned long long* gn_counters;
  int a_evict_first;  // L2 hint for the activation tiles (CDDPM_L2_HINTS)
};

__device__ __forceinline__ uint32_t pack16(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int fmt) {
  if (fmt == 1) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

__device__ __forceinline__ float silu_tanh(float x) {  // the GroupNorm kernel's SiLU (elementwise.cu: one SFU op)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One step of the 4 x 4 transpose of 16-byte items inside every group of four lanes: lanes whose `hi` bit is set trade
// item a for their partner's, the others item b (partner = lane ^ mask).
__device__ __forceinline__ void quad_exchange(uint4& a, uint4& b, bool hi, int mask) {
  const uint4 send = hi ? a : b;
  uint4 recv;
  recv.x = __shfl_xor_sync(0xffffffffu, send.x, mask);
  recv.y = __shfl_xor_sync(0xffffffffu, send.y, mask);
  recv.z = __shfl_xor_sync(0xffffffffu, send.z, mask);
  recv.w = __shfl_xor_sync(0xffffffffu, send.w, mask);
  if (hi) {
    a = recv;
  } else {
    b = recv;
  }
}

// Per 4-channel bucket: sum and sum of squares of this pixel's 32 channels, then a butterfly over the warp's 32 pixels
// (16 shuffles leave value (lane >> 1) & 15 in every lane); lanes write the warp's 16 totals of chunk c.
//
// kIL: lanes 8-15 and 24-31 of a warp hold pixels of the SECOND image of the tile's pair, so the butterfly runs over
// lane bits 4, 2, 1, 0 only (16 lanes per image, 16 values -> one finished value per lane, index (lane >> 4) * 8 +
// (lane & 7)) and lane bit 3 selects the image's 64-value block of stat_dst.
template <bool kIL = false>
__device__ __forceinline__ void emit_stats(const float (&f)[32], bool in_img, int lane, float* stat_dst) {
  float v16[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v16[k] = 0.f;
    v16[8 + k] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float fv = in_img ? f[4 * k + j] : 0.f;
      v16[k] += fv;
      v16[8 + k] = fmaf(fv, fv, v16[8 + k]);
    }
  }
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
    const int msk = kIL ? (w == 8 ? 16 : w) : w * 2;
    const bool hi = (lane & msk) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = hi ? v16[k + w] : v16[k];
      const float send = hi ? v16[k] : v16[k + w];
      v16[k] = keep + __shfl_xor_sync(0xffffffffu, send, msk);
    }
  }
  if constexpr (kIL) {
    stat_dst[((lane >> 3) & 1) * kStatW2 + ((lane >> 4) << 3) + (lane & 7)] = v16[0];
  } else {
    const float tot = v16[0] + __shfl_xor_sync(0xffffffffu, v16[0], 1);
    if ((lane & 1) == 0) stat_dst[lane >> 1] = tot;
  }
}

// kFuse: the experimental GroupNorm finish inside the epilogue (ConvDesc::gn_gamma) is its own instantiation, so the
// default kernel does not carry its registers.
template <int kMT, bool kFuse = false, bool kIL = false>
__global__ void __launch_bounds__(kThreads2, 1) conv_igemm2_kernel(const __grid_constant__ Conv2Params p) {
  static_assert(!kIL || !kFuse, "image-interleaved tiles: no fused GroupNorm finish");
  using G = Geo<kMT, kIL>;
  constexpr int kTileRows = G::kTileRows;
  constexpr int kRowStep = G::kImgs * G::kHaloW * (kRowBytes / 16);  // descriptor units between image rows of the halo tile
  constexpr int kStatW = G::kStatW;
  constexpr int kMTiles = kMT;
  constexpr int kTileW = G::kTileW;
  constexpr int kHaloPitch = G::kHaloPitch;
  constexpr int kASlotBytes = G::kASlotBytes;
  constexpr int kEpiWarps = G::kEpiWarps;
  constexpr int kStagesA = G::kStagesA;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + kStagesA * kASlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + kStagesB * kBSlotBytes);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + kStagesA;
  uint64_t* full_b = empty_a + kStagesA;
  uint64_t* empty_b = full_b + kStagesB;
  uint64_t* tfull = empty_b + kStagesB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_sh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  // a work item is two adjacent macro tiles (one per CTA of the pair) x one 128-channel N tile
  const int num_work = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  const int work_first = blockIdx.x / 2;
  const int work_stride = gridDim.x / 2;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  // Programmatic dependent launch: let the next kernel's CTAs be scheduled as ours retire; everything up to the first
  // read of the predecessor's output (barrier init, TMEM allocation, descriptor prefetch, the weight ring) runs under
  // the predecessor's tail.  The waits sit in the halo-tile producer and in the epilogue warps.
  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.tmap_a[s]);
    tma_prefetch_desc(&p.tmap_b);
  }
  cluster_sync_relaxed();  // both CTAs are resident before TMEM is allocated for the pair
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStagesA; ++i) {
      mbar_init(&full_a[i], 2);  // the leader's barrier collects both producers
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < kStagesB; ++i) {
      mbar_init(&full_b[i], 2);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * kEpiWarps * 32);  // both epilogues release the leader's accumulator stage
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_relaxed();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer (one per CTA)
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      for (int work = work_first; work < num_work; work += work_stride) {
        const int n_idx = work % p.num_n_tiles;
        const int b_row = n_idx * kNTile + static_cast<int>(cta_rank) * (kNTile / 2);
        for (int ci = 0; ci < p.num_chunks; ++ci) {
          const int s = p.chunk_src[ci], ch = p.chunk_ch[ci];
          const int C = p.src_c[s];
          const int ntaps = p.src_taps[s];
          const int koff = p.src_koff[s];
          for (int tap0 = 0; tap0 < ntaps; tap0 += kTapsPerStage) {
            const int nt = ntaps - tap0 < kTapsPerStage ? ntaps - tap0 : kTapsPerStage;
            mbar_wait(&empty_b[sb], pb ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&full_b[sb], static_cast<uint32_t>(2 * nt * kBTapBytes));
            } else {
              mbar_arrive_cluster(&full_b[sb], 0);
            }
            for (int t = 0; t < nt; ++t)
              tma_load_2d_pair(b_ring + sb * kBSlotBytes + t * kBTapBytes, &p.tmap_b, &full_b[sb],
                               koff + (tap0 + t) * C + ch * kConvBlockK, b_row);
            if (++sb == kStagesB) {
              sb = 0;
              pb ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ halo-tile producer (one per CTA)
    // Its own thread, so tile loads run ahead by the depth of the tile ring instead of trailing the weight ring.
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      pdl_wait();  // the activations are the predecessor's output
      const uint64_t pol = l2_policy_evict_first();
      const bool hint = p.a_evict_first != 0;
      for (int work = work_first; work < num_work; work += work_stride) {
        const int m_tile = 2 * (work / p.num_n_tiles) + static_cast<int>(cta_rank);
        // an odd trailing tile has no partner: its coordinates fall outside the batch and TMA fills zeros
        // (interleaved tiles: box b is tile kMT * m_tile + b of the list; `n` counts image PAIRS there)
        int bn[G::kBoxes], by[G::kBoxes], bx[G::kBoxes];
#pragma unroll
        for (int b = 0; b < G::kBoxes; ++b) {
          const int t = kIL ? m_tile * kMT + b : m_tile;
          bn[b] = t / tiles_per_img;
          const int r = t - bn[b] * tiles_per_img;
          by[b] = r / p.tiles_w;
          bx[b] = r - by[b] * p.tiles_w;
        }
        for (int ci = 0; ci < p.num_chunks; ++ci) {
          const int s = p.chunk_src[ci], ch = p.chunk_ch[ci];
          // a 1x1 source needs no halo: its box is the bare tile (32 KB instead of 41.5 KB at kMT = 2)
          const bool bare = p.box1 != 0 && p.src_taps[s] == 1;
          const int halo = bare ? 0 : 1;
          mbar_wait(&empty_a[sa], pa ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&full_a[sa], 2 * G::kBoxes * (bare ? G::kBareBytes : G::kBoxBytes));
          } else {
            mbar_arrive_cluster(&full_a[sa], 0);
          }
#pragma unroll
          for (int b = 0; b < G::kBoxes; ++b) {
            // coordinates in the order of the tensor map's dimensions: {C, W, H, B}, interleaved tiles {C, W, B, H}
            // (an image index past the batch is zero-filled like any halo)
            const int cy = by[b] * kTileRows - halo;
            const int c2 = kIL ? 2 * bn[b] : cy;
            const int c3 = kIL ? cy : bn[b];
            uint8_t* dst = a_ring + sa * kASlotBytes + b * G::kBoxPitch;
            if (hint) {
              tma_load_4d_pair_hint(dst, &p.tmap_a[s], &full_a[sa], ch * kConvBlockK, bx[b] * kTileW - halo, c2, c3, pol);
            } else {
              tma_load_4d_pair(dst, &p.tmap_a[s], &full_a[sa], ch * kConvBlockK, bx[b] * kTileW - halo, c2, c3);
            }
          }
          if (++sa == kStagesA) {
            sa = 0;
            pa ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // Every operand of the MMAs is warp-uniform (shuffled bases, descriptors advanced by constants) and the issuing
    // lane is elected, so the compiler emits back-to-back UTCHMMA from uniform registers: ~45 cycles of issue per MMA
    // against 64 cycles of tensor time (tools/ubench_mma.cu; per-thread operands cost ~160 cycles per MMA).
    const uint32_t idesc = umma_idesc_f16(256, kNTile, static_cast<uint32_t>(p.ab_format));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_ring_u = __shfl_sync(0xffffffffu, smem_u32(a_ring), 0);
    const uint32_t b_ring_u = __shfl_sync(0xffffffffu, smem_u32(b_ring), 0);
    const bool elected = elect_one_sync();
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int iter = 0;
    for (int work = work_first; work < num_work; work += work_stride, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + static_cast<uint32_t>(acc * kMTiles * kNTile);
      uint32_t accum = 0;
      {
        for (int ci = 0; ci < p.num_chunks; ++ci) {
          const int s = p.chunk_src[ci];
          const bool c3 = p.src_taps[s] == 9;
          const bool bare = !c3 && p.box1 != 0;
          mbar_wait(&full_a[sa], pa);
          const uint64_t a_desc = umma_desc_k128_sbo(a_ring_u + sa * kASlotBytes, bare ? kTileW * kRowBytes : kHaloPitch);
          const int nrows = c3 ? 3 : 1;
          for (int row = 0; row < nrows; ++row) {
            mbar_wait(&full_b[sb], pb);
            tc_fence_after();
            const bool last_row = row == nrows - 1;
            const bool last_step = last_row && ci == p.num_chunks - 1;
            if (elected) {
              const uint64_t bd = umma_desc_k128(b_ring_u + sb * kBSlotBytes);
              if (c3) {
                // kernel row `row`: taps (row, 0..2) read the halo tile from pixel (row, dx + 8 m)
                const uint64_t ad = a_desc + static_cast<uint64_t>(row * kRowStep);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                  for (int m = 0; m < kMTiles; ++m) {
#pragma unroll
                    for (int kk = 0; kk < kConvBlockK / 16; ++kk) {
                      umma_f16_ss_pair(tmem_d + m * kNTile, ad + (dx * (kRowBytes / 16) + m * G::kMStep + kk * 2),
                                       bd + (dx * (kBTapBytes / 16) + kk * 2), idesc,
                                       (dx == 0 && kk == 0) ? accum : 1u);
                    }
                  }
                }
              } else {
                // 1x1 source: the centre of the halo tile, or the bare tile from its first pixel
                const uint64_t ad = a_desc + static_cast<uint64_t>(bare ? 0 : kRowStep + kRowBytes / 16);
#pragma unroll
                for (int m = 0; m < kMTiles; ++m) {
#pragma unroll
                  for (int kk = 0; kk < kConvBlockK / 16; ++kk) {
                    umma_f16_ss_pair(tmem_d + m * kNTile, ad + (m * G::kMStep + kk * 2), bd + kk * 2, idesc,
                                     kk == 0 ? accum : 1u);
                  }
                }
              }
              umma_commit_pair(&empty_b[sb]);
              if (last_row) umma_commit_pair(&empty_a[sa]);
              if (last_step) umma_commit_pair(&tfull[acc]);
            }
            accum = 1;
            if (++sb == kStagesB) {
              sb = 0;
              pb ^= 1;
            }
          }
          if (++sa == kStagesA) {
            sa = 0;
            pa ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue (one warpgroup per M tile)
    const int quarter = warp & 3;     // TMEM lane quarter this warp may read
    const int m = (warp - 4) >> 2;    // M tile of this warpgroup
    const int ew = warp - 4;
    const int row = quarter * 32 + lane;  // pixel within the M tile: (row / 8, row % 8)
    const int fmt = p.ab_format;
    const int nb4 = p.Cout >> 2;
    const int epi_tid = threadIdx.x - 128;
    int iter = 0;
    pdl_wait();  // residual reads, statistics atomics and the output stores are ordered after the predecessor
    for (int work = work_first; work < num_work; work += work_stride, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      // the tile of this warpgroup: the CTA's macro tile, or (interleaved) its own entry of the tile list
      const int m_tile = (2 * (work / p.num_n_tiles) + static_cast<int>(cta_rank)) * (kIL ? kMT : 1) + (kIL ? m : 0);
      const int n_idx = work % p.num_n_tiles;
      const bool valid = m_tile < (kIL ? p.num_il_tiles : p.num_m_tiles);
      const int n = m_tile / tiles_per_img;
      const int r = m_tile - n * tiles_per_img;
      const int ty = r / p.tiles_w;
      const int tx = r - ty * p.tiles_w;
      const int x = tx * kTileW + (kIL ? 0 : m * 8) + (row & 7);
      // row group g = row >> 3: image row g of the tile, or (interleaved tiles) row g >> 1 of image 2 n + (g & 1)
      const int y = kIL ? ty * kTileRows + (row >> 4) : ty * kTileH + (row >> 3);
      const int img = kIL ? 2 * n + ((row >> 3) & 1) : n;
      const size_t off0 = ((static_cast<size_t>(img) * p.H + y) * p.W + x) * p.Cout + n_idx * kNTile;
      // the bottom tile of a 24-row image hangs over its edge; the last pair of an odd batch has no second image
      const bool in_img = valid && (kIL ? img < p.B : y < p.H);
      const bool has_res = p.residual != nullptr && in_img;

      if constexpr (kFuse) {
        // ---------------------------------------------------------------- fused GroupNorm (+FiLM) + SiLU finish
        // The GroupNorm that follows this convolution needs whole-image statistics.  Phase 1 reads the accumulator,
        // publishes this tile's partial sums (double atomics) and counts the tile in a per-(image, N tile) counter;
        // the tile then waits until all tiles of its image are counted - they are all in flight: the grid is
        // persistent, one CTA per SM, work items in image order - and phase 2 reads the accumulator AGAIN (it is still
        // in TMEM: the stage is only released after the second read), normalises and stores the activated tensor.
        // The raw convolution result never reaches memory and the separate GroupNorm pass disappears.
        float* coef_sh = stat_sh + 2 * kEpiWarps * kStatW2;  // A[128], B[128]
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>((acc * kMTiles + m) * kNTile);
        if (valid) {
#pragma unroll 1
          for (int c = 0; c < kNTile / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n_idx * kNTile + c * 32 + j))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
              f[j] = __uint_as_float(v[j]) + bv.x;
              f[j + 1] = __uint_as_float(v[j + 1]) + bv.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + bv.z;
              f[j + 3] = __uint_as_float(v[j + 3]) + bv.w;
            }
            emit_stats(f, in_img, lane, stat_sh + (acc * kEpiWarps + ew) * kStatW2 + c * 16);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (valid) {
          if (epi_tid < kStatW2) {
            const int within = epi_tid & 15;
            const int is_q = within >> 3;
            const int bucket = (epi_tid >> 4) * 8 + (within & 7);
            const float* sp = stat_sh + acc * kEpiWarps * kStatW2 + epi_tid;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kEpiWarps; ++w) t += sp[w * kStatW2];
            atomicAdd(&p.gn_stats[(static_cast<size_t>(n) * nb4 + n_idx * (kNTile >> 2) + bucket) * 2 + is_q],
                      static_cast<double>(t));
            __threadfence();
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (valid) {
          bool timed_out = false;
          if (epi_tid == 0) {
            unsigned long long* ctr = p.gn_counters + static_cast<size_t>(n) * p.num_n_tiles + n_idx;
            __threadfence();
            atomicAdd(ctr, 1ull);
            const unsigned long long want = static_cast<unsigned long long>(tiles_per_img);
            int spins = 0;
            while (ld_acquire_u64(ctr) < want) {
              __nanosleep(64);
              if (++spins > (1 << 21)) {  // ~0.2 s: a scheduling assumption broke; poison the output instead of hanging
                timed_out = true;
                break;
              }
            }
            coef_sh[0] = timed_out ? __int_as_float(0x7fc00000) : 0.f;  // read back below by every thread
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          const float poison = coef_sh[0];
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (epi_tid < kNTile) {
            const int cg = n_idx * kNTile + epi_tid;
            const int cpg = p.Cout >> 5;  // GroupNorm32: 32 groups
            const int g = cg / cpg;
            double s = 0.0, q = 0.0;
            for (int j = g * (cpg >> 2); j < (g + 1) * (cpg >> 2); ++j) {
              const double* sp = p.gn_stats + (static_cast<size_t>(n) * nb4 + j) * 2;
              s += __ldcg(sp);
              q += __ldcg(sp + 1);
            }
            const double cnt = static_cast<double>(p.H) * p.W * cpg;
            const double mean = s / cnt;
            double var = q / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            float A = static_cast<float>(1.0 / sqrt(var + 1e-5)) * __ldg(p.gn_gamma + cg);
            float Bc = __ldg(p.gn_beta + cg) - static_cast<float>(mean) * A;
            if (p.gn_film != nullptr) {
              const float* fp = p.gn_film + static_cast<size_t>(n) * p.gn_film_stride + p.gn_film_off;
              const float sc = 1.0f + fp[cg];
              const float shf = fp[p.Cout + cg];
              A *= sc;
              Bc = Bc * sc + shf;
            }
            coef_sh[epi_tid] = A + poison;
            coef_sh[kNTile + epi_tid] = Bc;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        }
#pragma unroll 1
        for (int c = 0; c < kNTile / 32; ++c) {
          uint32_t v[32];
          if (valid) {
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait();
          }
          if (c == kNTile / 32 - 1) {
            tc_fence_before();
            mbar_arrive_cluster(&tempty[acc], 0);
          }
          if (in_img) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n_idx * kNTile + c * 32 + j))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 av = *reinterpret_cast<const float4*>(coef_sh + c * 32 + j);
              const float4 cv = *reinterpret_cast<const float4*>(coef_sh + kNTile + c * 32 + j);
              f[j] = silu_tanh(fmaf(__uint_as_float(v[j]) + bv.x, av.x, cv.x));
              f[j + 1] = silu_tanh(fmaf(__uint_as_float(v[j + 1]) + bv.y, av.y, cv.y));
              f[j + 2] = silu_tanh(fmaf(__uint_as_float(v[j + 2]) + bv.z, av.z, cv.z));
              f[j + 3] = silu_tanh(fmaf(__uint_as_float(v[j + 3]) + bv.w, av.w, cv.w));
            }
            uint4* op = reinterpret_cast<uint4*>(p.out + off0 + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack16(f[q * 8 + 0], f[q * 8 + 1], fmt);
              o.y = pack16(f[q * 8 + 2], f[q * 8 + 3], fmt);
              o.z = pack16(f[q * 8 + 4], f[q * 8 + 5], fmt);
              o.w = pack16(f[q * 8 + 6], f[q * 8 + 7], fmt);
              op[q] = o;
            }
          }
        }
        continue;
      }

      uint4 rnext[4];
      if (has_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off0);
#pragma unroll
        for (int q = 0; q < 4; ++q) rnext[q] = __ldg(rp + q);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>((acc * kMTiles + m) * kNTile);
#pragma unroll 1
      for (int c = 0; c < kNTile / 32; ++c) {
        const int co = n_idx * kNTile + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        uint4 rcur[4];
        if (has_res) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
          if (c + 1 < kNTile / 32) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off0 + (c + 1) * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) rnext[q] = __ldg(rp + q);
          }
        }
        tmem_ld_wait();
        if (c == kNTile / 32 - 1) {
          // every TMEM read of this accumulator stage has completed: hand it back to the leader's MMA warp
          tc_fence_before();
          mbar_arrive_cluster(&tempty[acc], 0);
        }
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
        if (has_res) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w[4] = {rcur[q].x, rcur[q].y, rcur[q].z, rcur[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 t = unpack16(w[e], fmt);
              f[q * 8 + e * 2] += t.x;
              f[q * 8 + e * 2 + 1] += t.y;
            }
          }
        }
        if (p.relu == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.relu == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] / (1.0f + __expf(-f[j]));
        }
        {
          // A lane holds 64 contiguous bytes of ITS pixel; stored as is, every 16-byte store instruction of the warp
          // touches 32 different lines (measured: the scattered stores made the epilogue, not the MMAs, the per-item
          // critical path at K = 1152).  Transposing the four 16-byte items inside each group of four lanes first lets
          // the group write one pixel's 64 bytes per instruction: 8 lines per store instead of 32.
          uint4 o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[q].x = pack16(f[q * 8 + 0], f[q * 8 + 1], fmt);
            o[q].y = pack16(f[q * 8 + 2], f[q * 8 + 3], fmt);
            o[q].z = pack16(f[q * 8 + 4], f[q * 8 + 5], fmt);
            o[q].w = pack16(f[q * 8 + 6], f[q * 8 + 7], fmt);
          }
          const int e = lane & 3;
          quad_exchange(o[0], o[1], (e & 1) != 0, 1);
          quad_exchange(o[2], o[3], (e & 1) != 0, 1);
          quad_exchange(o[0], o[2], (e & 2) != 0, 2);
          quad_exchange(o[1], o[3], (e & 2) != 0, 2);
          if (in_img) {  // uniform over the group: its four pixels share an image row
            // item k now is bytes [16 e, 16 e + 16) of the pixel of lane (lane & ~3) + k
            uint16_t* ob = p.out + (off0 - static_cast<size_t>(e) * p.Cout) + c * 32 + e * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(ob + static_cast<size_t>(k) * p.Cout) = o[k];
          }
        }
        if (p.gn_stats != nullptr)
          emit_stats<kIL>(f, in_img, lane, stat_sh + (acc * kEpiWarps + ew) * kStatW + c * 16);
      }
      if (p.gn_stats != nullptr) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        // every warpgroup sums the four warps of ITS M tile, one thread per (image of the tile, 32-channel chunk,
        // value): a tile's partial sums then do not depend on how many tiles the CTA holds, so a small batch planned
        // with one tile per CTA reproduces the statistics of a large one planned with two
        const int st = epi_tid & 127;
        const int stat_img = kIL ? 2 * n + (st >> 6) : n;
        if (st < kStatW && valid && stat_img < p.B) {
          const int e = st & (kStatW2 - 1);
          const int within = e & 15;
          const int is_q = within >> 3;
          const int bucket = (e >> 4) * 8 + (within & 7);
          const float* sp = stat_sh + (acc * kEpiWarps + 4 * m) * kStatW + st;
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) t += sp[w * kStatW];
          atomicAdd(&p.gn_stats[(static_cast<size_t>(stat_img) * nb4 + n_idx * (kNTile >> 2) + bucket) * 2 + is_q],
                    static_cast<double>(t));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_relaxed();  // the peer may still be reading operands / TMEM that belong to the pair
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace

bool conv2_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDDPM_CONV_V2");  // A/B switch for measurements: 0 = first-generation kernel everywhere
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool conv2_supported(const ConvDesc& d) {
  if (d.flat_rows > 0 || d.out_is_f32 || d.out_stride > 0 || d.out_col_off != 0) return false;
  if (d.W % 8 != 0 || d.H % 8 != 0) return false;
  if (d.Cout % kNTile != 0) return false;
  for (int s = 0; s < d.num_src; ++s)
    if (d.src_c[s] % kConvBlockK != 0 || (d.src_taps[s] != 1 && d.src_taps[s] != 9)) return false;
  return true;
}

struct Conv2Launch {
  Conv2Params p;
};

int build_conv2(const ConvDesc& d, std::shared_ptr<void>* holder) {
  auto L = std::make_shared<Conv2Launch>();
  Conv2Params& p = L->p;
  memset(&p, 0, sizeof(p));
  p.num_src = d.num_src;
  p.B = d.B;
  p.H = d.H;
  p.W = d.W;
  p.Cout = d.Cout;
  p.mt = (d.W % 16 == 0 && d.H % 16 == 0) ? 2 : 1;
  {
    // Small batches: when the 16 x 16 macro tiles leave more than half of the CTA pairs without a work item (B = 1:
    // 18 items at 96 x 96, 10 at 48 x 48), plan 8 x 16 tiles instead - twice the CTAs, half the MMA chain each.
    // CDDPM_CONV_SPLIT_SMALL=0 keeps the macro tiles.
    static const int split_small = [] {
      const char* e = getenv("CDDPM_CONV_SPLIT_SMALL");
      return (e != nullptr && e[0] == '0') ? 0 : 1;
    }();
    const int tiles2 = d.B * (d.W / 16) * ((d.H + kTileH - 1) / kTileH);
    if (split_small != 0 && p.mt == 2 && d.gn_gamma == nullptr &&
        ((tiles2 + 1) / 2) * (d.Cout / kNTile) <= device_sm_count() / 4)
      p.mt = 1;
  }
  {
    // Image-interleaved tiles for the geometries whose height is not a multiple of 16 (the 24 x 24 level).  A/B switch
    // for measurements: CDDPM_CONV_IL=0 keeps the 16-row tiles with a masked bottom half.
    // 1 (default) = one interleaved tile per CTA; 2 = two per CTA (each its own TMA box, every weight stage serving both)
    // once the one-tile form needs a second round.  Measured (profiles/r02_c17_il2_conv.log): 2 is slower at B = 32 / 64
    // (22.1 -> 22.8 us, 31.8 -> 34.3 us for 256->256) and 0.5-1.8 % faster at B = 150 - the one-tile rounds are not
    // limited by the weight stream (6.1-6.6 us per round either way against 4.85 us of MMA time); kept as a switch.
    static const int il = [] {
      const char* e = getenv("CDDPM_CONV_IL");
      return (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }();
    p.il = (il != 0 && p.mt == 1 && d.H % 16 != 0 && d.gn_gamma == nullptr) ? 1 : 0;
    if (p.il != 0) {
      p.tiles_w = d.W / 8;
      p.tiles_h = d.H / (kTileH / 2);
      p.num_il_tiles = ((d.B + 1) / 2) * p.tiles_w * p.tiles_h;
      const int items1 = ((p.num_il_tiles + 1) / 2) * (d.Cout / kNTile);
      p.mt = (il == 2 && items1 > device_sm_count() / 2) ? 2 : 1;
      p.num_m_tiles = (p.num_il_tiles + p.mt - 1) / p.mt;
    }
  }
  if (p.il == 0) {
    p.tiles_w = d.W / (8 * p.mt);
    p.tiles_h = (d.H + kTileH - 1) / kTileH;
    p.num_m_tiles = d.B * p.tiles_w * p.tiles_h;
  }
  p.num_n_tiles = d.Cout / kNTile;
  p.ab_format = d.ab_format;
  p.relu = d.relu;
  p.bias = d.bias;
  p.residual = reinterpret_cast<const uint16_t*>(d.residual);
  p.out = reinterpret_cast<uint16_t*>(d.out);
  p.gn_stats = d.gn_stats;
  {
    // The activation operand is read once per forward (concurrently by the N tiles and the halo neighbours): marking
    // its L2 lines evict-first leaves the cache to the OUTPUT, which the following GroupNorm pass reads back to front.
    static const int hints = [] {
      const char* e = getenv("CDDPM_L2_HINTS");
      return (e != nullptr && e[0] == '0') ? 0 : 1;
    }();
    p.a_evict_first = hints;
  }
  if (d.gn_gamma != nullptr) {
    if (d.gn_stats == nullptr || d.gn_counters == nullptr || d.gn_beta == nullptr || d.residual != nullptr || d.relu != 0 ||
        d.Cout % 128 != 0 || (d.Cout / 32) % 4 != 0)
      return fail(kInvalidArgument, "conv2: fused GroupNorm finish needs statistics, counters, no residual / activation");
    p.gn_gamma = d.gn_gamma;
    p.gn_beta = d.gn_beta;
    p.gn_film = d.gn_film;
    p.gn_film_stride = d.gn_film_stride;
    p.gn_film_off = d.gn_film_off;
    p.gn_counters = d.gn_counters;
  }
  static const int interleave = [] {
    const char* e = getenv("CDDPM_SKIP_INTERLEAVE");  // A/B switch: 0 = sources in order, 1x1 sources stage the halo box
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  p.box1 = interleave;
  int ktot = 0;
  int n9 = 0, n1 = 0;
  for (int s = 0; s < d.num_src; ++s) {
    p.src_c[s] = d.src_c[s];
    p.src_taps[s] = d.src_taps[s];
    p.src_koff[s] = ktot;
    const bool bare = interleave && d.src_taps[s] == 1;
    const uint64_t C = static_cast<uint64_t>(d.src_c[s]);
    if (p.il != 0) {
      // the tensor as {C, W, B, H}: a box of two images x (8 + halo) rows is written image-fastest, i.e. with the rows
      // of the two images alternating - the order the interleaved M tile reads them in
      const uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.B), static_cast<uint64_t>(d.H)};
      const uint64_t strides[3] = {C * 2, C * 2 * d.W * d.H, C * 2 * d.W};
      const uint32_t box[4] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(8 + (bare ? 0 : 2)), 2u,
                               static_cast<uint32_t>(kTileH / 2 + (bare ? 0 : 2))};
      CDDPM_TRY(encode_tmap_16bit(&p.tmap_a[s], d.src[s], 4, dims, strides, box));
    } else {
      const uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
      const uint64_t strides[3] = {C * 2, C * 2 * d.W, C * 2 * d.W * d.H};
      const uint32_t box[4] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(8 * p.mt + (bare ? 0 : 2)),
                               static_cast<uint32_t>(bare ? kTileH : kHaloH), 1u};
      CDDPM_TRY(encode_tmap_16bit(&p.tmap_a[s], d.src[s], 4, dims, strides, box));
    }
    ktot += d.src_taps[s] * d.src_c[s];
    (d.src_taps[s] == 9 ? n9 : n1) += d.src_c[s] / kConvBlockK;
  }
  if (n9 + n1 > kMaxChunks) return fail(kUnsupported, "conv2: too many 64-channel chunks per work item");
  {
    // chunk schedule: sources in order; with interleaving the 1x1 chunks are dealt evenly between the 3x3 chunks
    // (Bresenham), each 3x3 chunk first so the accumulator's first MMA belongs to the main source
    int cs[kMaxChunks], cc[kMaxChunks], n = 0;
    auto emit = [&](int taps) {
      for (int s = 0; s < d.num_src; ++s)
        if (d.src_taps[s] == taps)
          for (int ch = 0; ch < d.src_c[s] / kConvBlockK; ++ch) {
            cs[n] = s;
            cc[n] = ch;
            ++n;
          }
    };
    if (!interleave || n9 == 0 || n1 == 0) {
      for (int s = 0; s < d.num_src; ++s)
        for (int ch = 0; ch < d.src_c[s] / kConvBlockK; ++ch) {
          cs[n] = s;
          cc[n] = ch;
          ++n;
        }
      for (int i = 0; i < n; ++i) {
        p.chunk_src[i] = static_cast<uint8_t>(cs[i]);
        p.chunk_ch[i] = static_cast<uint8_t>(cc[i]);
      }
    } else {
      emit(9);
      const int m9 = n;
      emit(1);
      int i9 = 0, i1 = 0, o = 0;
      while (i9 < n9 || i1 < n1) {
        // keep i1 / n1 <= i9 / n9: a 1x1 chunk goes out when it does not run ahead of the 3x3 stream
        if (i9 < n9 && (i1 >= n1 || static_cast<long long>(i1) * n9 >= static_cast<long long>(i9) * n1)) {
          p.chunk_src[o] = static_cast<uint8_t>(cs[i9]);
          p.chunk_ch[o] = static_cast<uint8_t>(cc[i9]);
          ++i9;
        } else {
          p.chunk_src[o] = static_cast<uint8_t>(cs[m9 + i1]);
          p.chunk_ch[o] = static_cast<uint8_t>(cc[m9 + i1]);
          ++i1;
        }
        ++o;
      }
    }
    p.num_chunks = n9 + n1;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(d.Cout)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ktot) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kConvBlockK), static_cast<uint32_t>(kNTile / 2)};
    CDDPM_TRY(encode_tmap_16bit(&p.tmap_b, d.wpacked, 2, dims, strides, box));
  }
  *holder = L;
  return kOk;
}

int launch_conv2(const std::shared_ptr<void>& holder, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Geo<1>::kSmemBytes));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Geo<2>::kSmemBytes));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Geo<1>::kSmemBytes));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Geo<2>::kSmemBytes));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Geo<1, true>::kSmemBytes));
    CDDPM_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Geo<2, true>::kSmemBytes));
    attr_set = true;
  }
  const Conv2Launch* L = reinterpret_cast<const Conv2Launch*>(holder.get());
  const int num_work = ((L->p.num_m_tiles + 1) / 2) * L->p.num_n_tiles;
  int pairs = device_sm_count() / 2;
  if (L->p.gn_gamma != nullptr) {
    // Fused GroupNorm finish: a tile waits for all tiles of its image.  Work items are dealt round-robin, so an image
    // whose items straddle two rounds would hold its early tiles' accumulator stage for a whole extra round.  Use a
    // pair count that is a multiple of the items of one image (of two images when the tile count is odd and a pair
    // spans an image boundary): every image then lives inside one round.  (72 of 74 pairs for the three UNet levels.)
    const int tiles_per_img = L->p.tiles_w * L->p.tiles_h;
    const int group = (tiles_per_img % 2 == 0 ? tiles_per_img / 2 : tiles_per_img) * L->p.num_n_tiles;
    if (group <= pairs) pairs = pairs / group * group;
  }
  if (num_work < pairs) pairs = num_work;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads2);
  if (L->p.il != 0) {
    cfg.dynamicSmemBytes = L->p.mt == 2 ? Geo<2, true>::kSmemBytes : Geo<1, true>::kSmemBytes;
  } else {
    cfg.dynamicSmemBytes = L->p.mt == 2 ? Geo<2>::kSmemBytes : Geo<1>::kSmemBytes;
  }
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_take_next() && pdl_mode() != 2) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  const bool fuse = L->p.gn_gamma != nullptr;
  if (L->p.il != 0) {
    if (L->p.mt == 2) {
      CDDPM_CUDA(cudaLaunchKernelEx(&cfg, (conv_igemm2_kernel<2, false, true>), L->p));
    } else {
      CDDPM_CUDA(cudaLaunchKernelEx(&cfg, (conv_igemm2_kernel<1, false, true>), L->p));
    }
  } else if (L->p.mt == 2) {
    if (fuse) {
      CDDPM_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm2_kernel<2, true>, L->p));
    } else {
      CDDPM_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm2_kernel<2>, L->p));
    }
  } else {
    if (fuse) {
      CDDPM_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm2_kernel<1, true>, L->p));
    } else {
      CDDPM_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm2_kernel<1>, L->p));
    }
  }
  return check_launch("conv_igemm2_kernel");
}

}  // namespace cddpm
