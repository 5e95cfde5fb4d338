// Weight gradient of the 3x3 / 1x1 convolutions on tcgen05 (sm_100a).
//
// GEMM view: dW[co][tap][ci] = sum over pixels of dY[pixel][co] * X[pixel + tap][ci] - the contraction runs over
// PIXELS, and in NHWC both operands have their M / N index (the channel) contiguous, so both are MN-major UMMA
// operands read straight from the tiles TMA stages (no transposed copies in HBM):
//   A = dY tile: 8 x 16 output pixels, two 64-channel chunks of 128 rows x 128 B (chunk stride = LBO),
//   B = X tile of ONE kernel row dy: rows y + dy - 1 of the 18-pixel-wide halo, two 64-channel chunks of 144 rows.
// One MMA (M = 128 output channels, N = 128 input channels, K = 16) covers one image row of the tile: its A rows are
// pixels (y, 0..15), its B rows are halo pixels (y, dx .. dx + 15) - sixteen consecutive 128-byte rows starting at an
// arbitrary row, which the descriptor allows because the 128-byte swizzle is a function of the absolute shared-memory
// address (tools/ubench_mma.cu).  The three taps of the kernel row accumulate into three 128-column TMEM accumulators
// that live for the whole pixel range of the CTA.
//
// Work decomposition: unit = (128 output channels) x (128 input channels of one source) x (kernel row); the pixel
// tiles of a unit are split over as many CTAs as fill the GPU, each adding its partial result into the fp32 gradient
// panel with red.global.add (the caller zeroes it).  Per tile a CTA stages 68 KB for 1536 tensor cycles
// (~44 B/cycle, the L2->SM limit of the chip is ~43 B/cycle/SM): the kernel sits on the L2->SM roofline; pairing CTAs
// (cta_group::2, shared X tile) is the next step.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include <stdlib.h>
#include <string.h>

#include <memory>

#include "conv_bwd.cuh"
#include "ptx.cuh"

namespace cddpm {

namespace {

constexpr int kWgThreads = 256;
constexpr int kWgStages = 3;
constexpr int kWgTileW = 16;
constexpr int kWgTileH = 8;
constexpr int kWgABytes = kWgTileW * kWgTileH * 128;          // 16384: one 64-channel chunk of the dY tile
constexpr int kWgBBytes3 = (kWgTileW + 2) * kWgTileH * 128;   // 18432: one 64-channel chunk of a halo row band
// Image-interleaved tiles (widths that are not a multiple of 16, i.e. the 24 x 24 level): 8 pixels x 8 rows of TWO
// consecutive images through one TMA box over the tensor seen as {C, W, B, H} - a row of the staged tile is 8 pixels
// of image 2 n followed by 8 pixels of image 2 n + 1, the 16 contraction rows of one MMA.  The dY tile has the bytes
// and the layout of the plain 16-pixel tile; the halo band is 2 x (8 + 2) pixels wide, so the two 8-row groups of an
// MMA's B operand are 10 rows apart (SBO 1280).  A 24-pixel row then is three full tiles of pixels per image pair
// instead of two tiles per image, the second half empty: 9 instead of 12 tiles per pair of 24 x 24 images.
constexpr int kWgBBytes3IL = 2 * (8 + 2) * kWgTileH * 128;    // 20480
constexpr int kWgStageBytes = 2 * kWgABytes + 2 * kWgBBytes3IL; // 73728
constexpr int kWgSmem = kWgStages * kWgStageBytes + 256 + 1024;
constexpr int kWgMaxUnits = 48;
constexpr int kWgTmemCols = 512;

struct WgradParams {
  CUtensorMap tmap_dy;              // {Cout, W, H, B}; box {64, 16, 8, 1}
  CUtensorMap tmap_x[kConvMaxSrc];  // {C_s, W, H, B}; box {64, 18, 8, 1} (3x3) or {64, 16, 8, 1} (1x1)
  int B, H, W, Cout, ktot;
  int tiles_w, tiles_h, num_tiles;
  int il;  // image-interleaved tiles: maps are {C, W, B, H}, tiles count per image PAIR
  int units_per_co, splits;
  int ab_format;
  float* dw;
  int src_c[kConvMaxSrc], src_taps[kConvMaxSrc], src_koff[kConvMaxSrc];
  uint8_t u_src[kWgMaxUnits], u_ci[kWgMaxUnits], u_dy[kWgMaxUnits];
};

// MN-major operand: rows of 128 bytes (64 channels) per K index, 8-row groups of 1024 bytes (SBO), 64-channel chunks
// `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// fire-and-forget vector reduction (sm_90+): four consecutive floats per instruction
__device__ __forceinline__ void red_add_v4(float* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(__uint_as_float(a)),
               "f"(__uint_as_float(b)), "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kWgStages * kWgStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kWgStages;
  uint64_t* tfull = empty + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = static_cast<int>(blockIdx.x) % p.splits;
  const int unit_g = static_cast<int>(blockIdx.x) / p.splits;
  const int co_chunk = unit_g / p.units_per_co;
  const int u = unit_g - co_chunk * p.units_per_co;
  const int s = p.u_src[u];
  const int ci_chunk = p.u_ci[u];
  const int kdy = p.u_dy[u];
  const bool c3 = p.src_taps[s] == 9;
  const int ntaps = c3 ? 3 : 1;
  const int t0 = static_cast<int>(static_cast<int64_t>(p.num_tiles) * split / p.splits);
  const int t1 = static_cast<int>(static_cast<int64_t>(p.num_tiles) * (split + 1) / p.splits);
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_dy);
    tma_prefetch_desc(&p.tmap_x[s]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kWgTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (t1 > t0) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (lane == 0) {
        int st = 0;
        uint32_t ph = 0;
        const uint32_t b_bytes = c3 ? (p.il != 0 ? kWgBBytes3IL : kWgBBytes3) : kWgABytes;
        const int tw = p.il != 0 ? 8 : kWgTileW;
        for (int t = t0; t < t1; ++t) {
          const int n = t / tiles_per_img;
          const int r = t - n * tiles_per_img;
          const int ty = r / p.tiles_w;
          const int tx = r - ty * p.tiles_w;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&full[st], 2 * kWgABytes + 2 * b_bytes);
          uint8_t* a = ring + st * kWgStageBytes;
          uint8_t* b = a + 2 * kWgABytes;
          const int bx = c3 ? tx * tw - 1 : tx * tw;
          const int by = c3 ? ty * kWgTileH + kdy - 1 : ty * kWgTileH;
          // coordinates in the order of the maps' dimensions: {C, W, H, B}, interleaved {C, W, B, H} (n = image pair)
          const int a2 = p.il != 0 ? 2 * n : ty * kWgTileH, a3 = p.il != 0 ? ty * kWgTileH : n;
          const int b2 = p.il != 0 ? 2 * n : by, b3 = p.il != 0 ? by : n;
          tma_load_4d(a, &p.tmap_dy, &full[st], co_chunk * 128, tx * tw, a2, a3);
          tma_load_4d(a + kWgABytes, &p.tmap_dy, &full[st], co_chunk * 128 + 64, tx * tw, a2, a3);
          tma_load_4d(b, &p.tmap_x[s], &full[st], ci_chunk * 128, bx, b2, b3);
          tma_load_4d(b + b_bytes, &p.tmap_x[s], &full[st], ci_chunk * 128 + 64, bx, b2, b3);
          if (++st == kWgStages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (warp-uniform operands)
      const uint32_t idesc = umma_idesc_f16(128, 128, static_cast<uint32_t>(p.ab_format)) | (1u << 15) | (1u << 16);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t ring_u = __shfl_sync(0xffffffffu, smem_u32(ring), 0);
      const bool elected = elect_one_sync();
      const uint32_t b_lbo = c3 ? (p.il != 0 ? kWgBBytes3IL : kWgBBytes3) : kWgABytes;
      const uint32_t b_pitch = c3 ? (p.il != 0 ? 2 * (8 + 2) * 128 : (kWgTileW + 2) * 128) : kWgTileW * 128;
      const uint32_t b_sbo = (c3 && p.il != 0) ? (8 + 2) * 128 : 1024;  // the second image's pixels follow the first's halo
      int st = 0;
      uint32_t ph = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        if (elected) {
          const uint32_t a_addr = ring_u + st * kWgStageBytes;
          const uint32_t b_addr = a_addr + 2 * kWgABytes;
          const uint64_t ad0 = umma_desc_mn128(a_addr, kWgABytes);
          const uint64_t bd0 = umma_desc_mn128(b_addr, b_lbo, b_sbo);
#pragma unroll
          for (int y = 0; y < kWgTileH; ++y) {
            const uint64_t ad = ad0 + static_cast<uint64_t>(y * (kWgTileW * 128 / 16));
            const uint64_t bdy = bd0 + static_cast<uint64_t>(y * (b_pitch / 16));
            const uint32_t acc = (t != t0 || y != 0) ? 1u : 0u;
            if (c3) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
                umma_f16_ss(tmem_u + dx * 128, ad, bdy + static_cast<uint64_t>(dx * (128 / 16)), idesc, acc);
            } else {
              umma_f16_ss(tmem_u, ad, bdy, idesc, acc);
            }
          }
          umma_commit(&empty[st]);
          if (t == t1 - 1) umma_commit(tfull);
        }
        __syncwarp();
        if (++st == kWgStages) {
          st = 0;
          ph ^= 1;
        }
      }
    } else if (warp >= 4) {
      // ---------------------------------------------------------------- epilogue: TMEM -> red.global.add.f32
      const int quarter = warp & 3;
      const int co = co_chunk * 128 + quarter * 32 + lane;
      mbar_wait(tfull, 0);
      tc_fence_after();
      const int C = p.src_c[s];
      float* row = p.dw + static_cast<size_t>(co) * p.ktot + p.src_koff[s] + ci_chunk * 128;
      for (int dx = 0; dx < ntaps; ++dx) {
        const int tap = c3 ? kdy * 3 + dx : 0;
        float* dst = row + static_cast<size_t>(tap) * C;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(dx * 128);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) red_add_v4(dst + c * 32 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

__global__ void pack_conv_weight_t_kernel(const float* __restrict__ w, int Cout, int Cin_total, int ksize, int cin_off,
                                          int C_s, uint16_t* __restrict__ out, int Ktot, int koff, int ab_format) {
  const int taps = ksize * ksize;
  const size_t total = static_cast<size_t>(C_s) * taps * Cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    const int tap = static_cast<int>((i / Cout) % taps);
    const int ci = static_cast<int>(i / (static_cast<size_t>(Cout) * taps));
    const float v = w[(static_cast<size_t>(co) * Cin_total + cin_off + ci) * taps + (taps - 1 - tap)];
    uint16_t bits;
    if (ab_format == 1) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    } else {
      __half h = __float2half_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&h);
    }
    out[static_cast<size_t>(ci) * Ktot + koff + tap * Cout + co] = bits;
  }
}

__global__ void unpack_conv_grad_kernel(const float* __restrict__ dw, int Cout, int Cin_total, int ksize, int cin_off,
                                        int C_s, float* __restrict__ grad, int Ktot, int koff) {
  const int taps = ksize * ksize;
  const size_t total = static_cast<size_t>(Cout) * taps * C_s;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % C_s);
    const int tap = static_cast<int>((i / C_s) % taps);
    const int co = static_cast<int>(i / (static_cast<size_t>(C_s) * taps));
    grad[(static_cast<size_t>(co) * Cin_total + cin_off + ci) * taps + tap] =
        dw[static_cast<size_t>(co) * Ktot + koff + tap * C_s + ci];
  }
}

struct WgradLaunch {
  WgradParams p;
  int grid = 0;
};

}  // namespace

bool wgrad_supported(const WgradDesc& d) {
  if (d.num_src < 1 || d.num_src > kConvMaxSrc) return false;
  if (d.Cout % 128 != 0 || d.W % 8 != 0 || d.H % kWgTileH != 0) return false;
  for (int s = 0; s < d.num_src; ++s) {
    if (d.src_taps[s] != 1 && d.src_taps[s] != 9) return false;
    if (!d.src_skip[s] && d.src_c[s] % 128 != 0) return false;
  }
  return true;
}

int64_t wgrad_flops(const WgradDesc& d) {
  int64_t k = 0;
  for (int s = 0; s < d.num_src; ++s)
    if (!d.src_skip[s]) k += static_cast<int64_t>(d.src_taps[s]) * d.src_c[s];
  return 2ll * d.B * d.H * d.W * d.Cout * k;
}

int build_wgrad(const WgradDesc& d, std::shared_ptr<void>* holder) {
  if (!wgrad_supported(d))
    return fail(kUnsupported, "conv wgrad needs channel counts that are multiples of 128 and H, W multiples of 8");
  if (!d.dy || !d.dw) return fail(kInvalidArgument, "conv wgrad: null pointer");
  auto L = std::make_shared<WgradLaunch>();
  WgradParams& p = L->p;
  memset(&p, 0, sizeof(p));
  p.B = d.B;
  p.H = d.H;
  p.W = d.W;
  p.Cout = d.Cout;
  p.ab_format = d.ab_format;
  p.dw = d.dw;
  {
    static const int il = [] {
      const char* e = getenv("CDDPM_WGRAD_IL");  // A/B switch: 0 = 16-pixel tiles with a zero-filled overhang
      return (e != nullptr && e[0] == '0') ? 0 : 1;
    }();
    p.il = (il != 0 && d.W % kWgTileW != 0) ? 1 : 0;
  }
  p.tiles_h = d.H / kWgTileH;
  if (p.il != 0) {
    p.tiles_w = d.W / 8;
    p.num_tiles = ((d.B + 1) / 2) * p.tiles_w * p.tiles_h;
  } else {
    p.tiles_w = (d.W + kWgTileW - 1) / kWgTileW;
    p.num_tiles = d.B * p.tiles_w * p.tiles_h;
  }
  // a tensor [B][H][W][C] as a tiled map: {C, W, H, B} with a box of `bw` pixels x 8 rows, or (interleaved tiles)
  // {C, W, B, H} with a box of `bw` pixels x 2 images x 8 rows
  const int il_flag = p.il;
  auto encode = [il_flag, &d](CUtensorMap* map, const void* base, uint64_t C, uint32_t bw) {
    if (il_flag != 0) {
      const uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.B), static_cast<uint64_t>(d.H)};
      const uint64_t strides[3] = {C * 2, C * 2 * d.W * d.H, C * 2 * d.W};
      const uint32_t box[4] = {64u, bw, 2u, static_cast<uint32_t>(kWgTileH)};
      return encode_tmap_16bit(map, base, 4, dims, strides, box);
    }
    const uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
    const uint64_t strides[3] = {C * 2, C * 2 * d.W, C * 2 * d.W * d.H};
    const uint32_t box[4] = {64u, bw, static_cast<uint32_t>(kWgTileH), 1u};
    return encode_tmap_16bit(map, base, 4, dims, strides, box);
  };
  const uint32_t tile_w = p.il != 0 ? 8u : static_cast<uint32_t>(kWgTileW);
  CDDPM_TRY(encode(&p.tmap_dy, d.dy, static_cast<uint64_t>(d.Cout), tile_w));
  int koff = 0, units = 0;
  for (int s = 0; s < d.num_src; ++s) {
    p.src_c[s] = d.src_c[s];
    p.src_taps[s] = d.src_taps[s];
    p.src_koff[s] = koff;
    koff += d.src_taps[s] * d.src_c[s];
    if (d.src_skip[s]) continue;
    if (!d.src[s]) return fail(kInvalidArgument, "conv wgrad: null source");
    CDDPM_TRY(encode(&p.tmap_x[s], d.src[s], static_cast<uint64_t>(d.src_c[s]),
                     d.src_taps[s] == 9 ? tile_w + 2 : tile_w));
    for (int ci = 0; ci < d.src_c[s] / 128; ++ci)
      for (int dy = 0; dy < (d.src_taps[s] == 9 ? 3 : 1); ++dy) {
        if (units >= kWgMaxUnits) return fail(kUnsupported, "conv wgrad: too many work units");
        p.u_src[units] = static_cast<uint8_t>(s);
        p.u_ci[units] = static_cast<uint8_t>(ci);
        p.u_dy[units] = static_cast<uint8_t>(dy);
        ++units;
      }
  }
  p.ktot = koff;
  p.units_per_co = units;
  if (units == 0) return fail(kInvalidArgument, "conv wgrad: nothing to compute");
  const int total_units = units * (d.Cout / 128);
  int splits = device_sm_count() / total_units;
  if (splits < 1) splits = 1;
  if (splits > p.num_tiles) splits = p.num_tiles;
  p.splits = splits;
  L->grid = total_units * splits;
  *holder = L;
  return kOk;
}

int launch_wgrad(const std::shared_ptr<void>& holder, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CDDPM_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
    attr_set = true;
  }
  const WgradLaunch* L = reinterpret_cast<const WgradLaunch*>(holder.get());
  conv_wgrad_kernel<<<L->grid, kWgThreads, kWgSmem, stream>>>(L->p);
  return check_launch("conv_wgrad_kernel");
}

int launch_pack_conv_weight_T(const float* w_oihw, int Cout, int Cin_total, int ksize, int cin_off, int C_s,
                              void* wpacked_t, int Ktot, int koff, int ab_format, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(Cout) * ksize * ksize * C_s;
  if (job_recorder() != nullptr) {
    ParamJob j = {kJobPackT, Cout, Cin_total, ksize, cin_off, C_s, Ktot, koff, ab_format, w_oihw, nullptr, wpacked_t,
                  static_cast<long long>(total)};
    job_record(j);
    return kOk;
  }
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  pack_conv_weight_t_kernel<<<blocks, 256, 0, stream>>>(w_oihw, Cout, Cin_total, ksize, cin_off, C_s,
                                                        reinterpret_cast<uint16_t*>(wpacked_t), Ktot, koff, ab_format);
  return check_launch("pack_conv_weight_t_kernel");
}

int launch_unpack_conv_grad(const float* dw_packed, int Cout, int Cin_total, int ksize, int cin_off, int C_s,
                            float* grad_oihw, int Ktot, int koff, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(Cout) * ksize * ksize * C_s;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  unpack_conv_grad_kernel<<<blocks, 256, 0, stream>>>(dw_packed, Cout, Cin_total, ksize, cin_off, C_s, grad_oihw, Ktot,
                                                      koff);
  return check_launch("unpack_conv_grad_kernel");
}

}  // namespace cddpm
