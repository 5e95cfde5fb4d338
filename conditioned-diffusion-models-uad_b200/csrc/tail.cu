// Anomaly-scoring tail kernels (see tail.cuh).  The volumes are ~0.5 M voxels: these kernels are latency- and
// launch-bound, not bandwidth-bound; the win over the reference is removing ~1 s of host scipy/sklearn per volume.
#include "tail.cuh"

#include <stdlib.h>

#include <algorithm>
#include <cub/cub.cuh>

namespace cddpm {

namespace {

struct View {
  const float* p;
  long long sy, sx, sd;
  __device__ __forceinline__ float at(int y, int x, int d) const { return p[y * sy + x * sx + d * sd]; }
};
View dev_view(const VolView& v) { return View{v.p, v.sy, v.sx, v.sd}; }

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}
__device__ __forceinline__ unsigned long long block_sum_u(unsigned long long v, unsigned long long* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  unsigned long long r = 0;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------- trilinear resize
// F.interpolate(vol[None, None], size=(Ho, Wo, Do), mode="trilinear", align_corners=True) (utils_eval.py:24-25): source
// coordinate = i * (in - 1) / (out - 1) in fp32, lower index = floor, upper = lower + (lower < in - 1), weights
// (1 - l, l); the three axes are folded outermost-first like ATen's upsample kernel:
// wy0 (wx0 (wd0 a + wd1 b) + wx1 (...)) + wy1 (...).  dst is [Ho][Wo][Do] contiguous (d fastest), the layout of the
// reference's squeezed tensor.
struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_coord(int o, int in, int out) {
  Lerp r;
  const float scale = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(o);
  r.i0 = min(static_cast<int>(src), in - 1);
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.w1 = src - static_cast<float>(r.i0);
  r.w0 = 1.0f - r.w1;
  return r;
}
__global__ void __launch_bounds__(256) trilinear_kernel(View src, int H, int W, int D, float* __restrict__ dst, int Ho,
                                                        int Wo, int Do) {
  const long long n = static_cast<long long>(Ho) * Wo * Do;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(idx % Do);
    const int x = static_cast<int>((idx / Do) % Wo);
    const int y = static_cast<int>(idx / (static_cast<long long>(Do) * Wo));
    const Lerp ly = lerp_coord(y, H, Ho), lx = lerp_coord(x, W, Wo), ld = lerp_coord(d, D, Do);
    auto line = [&](int yy, int xx) {
      return __fadd_rn(__fmul_rn(ld.w0, src.at(yy, xx, ld.i0)), __fmul_rn(ld.w1, src.at(yy, xx, ld.i1)));
    };
    auto plane = [&](int yy) { return __fadd_rn(__fmul_rn(lx.w0, line(yy, lx.i0)), __fmul_rn(lx.w1, line(yy, lx.i1))); };
    dst[idx] = __fadd_rn(__fmul_rn(ly.w0, plane(ly.i0)), __fmul_rn(ly.w1, plane(ly.i1)));
  }
}

// ---------------------------------------------------------------------------------------------- residual + erosion
__global__ void __launch_bounds__(256) residual_erode_kernel(View orig, View reco, View seg, View mask, int H, int W,
                                                             int D, int iterations, int erode,
                                                             float* __restrict__ out, double* __restrict__ sums) {
  __shared__ double sh[8];
  const long long n = static_cast<long long>(H) * W * D;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int d = static_cast<int>(idx / (static_cast<long long>(W) * H));
    const float o = orig.at(y, x, d), r = reco.at(y, x, d);
    const float a = fabsf(o - r);
    const double ad = static_cast<double>(a), sq = static_cast<double>(r - o) * static_cast<double>(r - o);
    acc[0] += ad;
    acc[1] += sq;
    if (seg.p != nullptr && seg.at(y, x, d) > 0.f) {
      acc[2] += ad;
      acc[3] += sq;
      acc[6] += 1.0;
    } else {
      acc[4] += ad;
      acc[5] += sq;
    }
    float v = a;
    if (erode) {
      bool keep = iterations >= 1;
      if (keep) {
        for (int dy = -iterations; dy <= iterations && keep; ++dy) {
          const int span = iterations - abs(dy);
          const int yy = y + dy;
          for (int dx = -span; dx <= span; ++dx) {
            const int xx = x + dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W || !(mask.at(yy, xx, d) > 0.f)) {
              keep = false;
              break;
            }
          }
        }
      }
      v = keep ? a : 0.f;
    }
    out[idx] = v;
  }
  if (sums != nullptr) {
    // Run-to-run identical sums (a volume must score the same on whichever rank of a sharded sweep it lands, to the last
    // bit - found by the 8-GPU sweep check: double atomics across blocks moved l1recoErrorAll by one ulp).  The per-thread
    // and per-block sums are fixed-order fp64; across blocks the partials are added as 2^-32 fixed point in 64-bit
    // INTEGER atomics (order-free; rounding <= 1.2e-10 per block), and the block that draws the last
    // ticket (bits 48+ of slot 6, whose low bits carry the exact lesion count) converts the slots to doubles in place.
    // A block partial that is not finite or >= 2^20 flags its slot (bits 40-45 of slot 6): that sum is reported as NaN.
    double t[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) t[k] = block_sum(acc[k], sh);
    if (threadIdx.x == 0) {
      unsigned long long* isums = reinterpret_cast<unsigned long long*>(sums);
      constexpr double kScale = 4294967296.0;
      unsigned long long flags = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (!(fabs(t[k]) < 1048576.0)) {  // <= 1184 blocks x 2^20 x 2^32 stays inside 63 bits
          flags |= 1ull << (40 + k);
        } else {
          atomicAdd(&isums[k], static_cast<unsigned long long>(__double2ll_rn(t[k] * kScale)));
        }
      }
      if (flags) atomicOr(&isums[6], flags);
      __threadfence();
      const unsigned long long old = atomicAdd(&isums[6], static_cast<unsigned long long>(t[6]) + (1ull << 48));
      if ((old >> 48) == gridDim.x - 1) {
        __threadfence();
        const unsigned long long s6 = atomicAdd(&isums[6], 0ull);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const long long v = static_cast<long long>(atomicAdd(&isums[k], 0ull));
          sums[k] = ((s6 >> (40 + k)) & 1ull) ? __longlong_as_double(0x7ff8000000000000ll)
                                              : static_cast<double>(v) / kScale;
        }
        sums[6] = static_cast<double>(s6 & ((1ull << 40) - 1));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- 3-D median
__device__ __forceinline__ uint32_t sort_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ int reflect(int i, int n) {
  // scipy 'reflect' == numpy 'symmetric': (d c b a | a b c d | d c b a)
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - i - 1;
  return min(max(i, 0), n - 1);
}

template <int K>
__global__ void __launch_bounds__(256) median3d_kernel(const float* __restrict__ in, float* __restrict__ out, int H,
                                                       int W, int D) {
  constexpr int R = K / 2;
  constexpr int TX = 16, TY = 4, TD = 4;  // 256 output voxels per CTA
  constexpr int SX = TX + 2 * R, SY = TY + 2 * R, SD = TD + 2 * R;
  constexpr int NV = K * K * K;
  constexpr int RANK = NV / 2;
  __shared__ uint32_t tile[SD][SY][SX];
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, d0 = blockIdx.z * TD;
  for (int i = threadIdx.x; i < SD * SY * SX; i += blockDim.x) {
    const int tx = i % SX, ty = (i / SX) % SY, td = i / (SX * SY);
    const int gx = reflect(x0 + tx - R, W), gy = reflect(y0 + ty - R, H), gd = reflect(d0 + td - R, D);
    tile[td][ty][tx] = sort_key(in[(static_cast<size_t>(gd) * H + gy) * W + gx]);
  }
  __syncthreads();
  const int lx = threadIdx.x % TX, ly = (threadIdx.x / TX) % TY, ld = threadIdx.x / (TX * TY);
  const int x = x0 + lx, y = y0 + ly, d = d0 + ld;
  if (x >= W || y >= H || d >= D) return;
  uint32_t key[NV];
  int zeros = 0;
  const uint32_t zero_key = 0x80000000u;
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const uint32_t kv = tile[ld + a][ly + b][lx + c];
        key[(a * K + b) * K + c] = kv;
        zeros += (kv == zero_key);
      }
  const size_t oidx = (static_cast<size_t>(d) * H + y) * W + x;
  // the residual volume is >= 0 and mostly exactly 0 outside the eroded brain mask: then the median is 0 as soon as
  // more than half of the window is 0 and nothing is negative
  if (zeros > RANK) {
    int neg = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) neg += (key[i] < zero_key);
    if (neg == 0) {
      out[oidx] = 0.f;
      return;
    }
  }
  // exact rank selection by radix descent over the 32 key bits
  uint32_t result = 0;
  int k = RANK;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t hi_mask = (bit == 31) ? 0u : (0xFFFFFFFFu << (bit + 1));
    int cnt0 = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) cnt0 += (((key[i] & hi_mask) == result) && (((key[i] >> bit) & 1u) == 0u));
    if (k >= cnt0) {
      k -= cnt0;
      result |= (1u << bit);
    }
  }
  out[oidx] = key_value(result);
}

// ---------------------------------------------------------------------------------------------- small reductions
__global__ void max_key_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out) {
  unsigned int m = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    m = max(m, sort_key(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
__global__ void max_decode_kernel(unsigned int* inout) {
  const unsigned int k = *inout;
  *reinterpret_cast<float*>(inout) = key_value(k);
}

struct QList {
  float q[4];
  int nq;
};

__global__ void __launch_bounds__(256) threshold_counts_kernel(const float* __restrict__ x, View seg, int H, int W,
                                                               int D, QList ql, unsigned long long* __restrict__ counts) {
  __shared__ unsigned long long sh[8];
  const long long n = static_cast<long long>(H) * W * D;
  unsigned long long c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int d = static_cast<int>(idx / (static_cast<long long>(W) * H));
    const bool g = seg.at(y, xx, d) > 0.f;
    const float v = x[idx];
    c[0] += g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < ql.nq) {
        const bool p = v > ql.q[i];
        c[1 + 2 * i] += p;
        c[2 + 2 * i] += (p && g);
      }
    }
  }
  for (int k = 0; k < 1 + 2 * ql.nq; ++k) {
    const unsigned long long t = block_sum_u(c[k], sh);
    if (threadIdx.x == 0 && t) atomicAdd(&counts[k], t);
  }
}

__global__ void threshold_mask_kernel(const float* __restrict__ x, long long n, float thr, unsigned char* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = x[i] > thr ? 1 : 0;
}

__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ x, View seg, View mask, int H, int W,
                                                        int D, float thr, unsigned long long* __restrict__ rows,
                                                        double* __restrict__ rowsum) {
  __shared__ unsigned long long shu[8];
  __shared__ double shd[8];
  const int y = blockIdx.x;
  unsigned long long c[4] = {0, 0, 0, 0};
  double s = 0.0;
  for (int i = threadIdx.x; i < W * D; i += blockDim.x) {
    const int xx = i % W, d = i / W;
    const float v = x[(static_cast<size_t>(d) * H + y) * W + xx];
    const bool g = seg.p != nullptr && seg.at(y, xx, d) > 0.f;
    const bool m = mask.at(y, xx, d) > 0.f;
    const bool p = v > thr;
    c[0] += p;
    c[1] += g;
    c[2] += (p && g);
    c[3] += m;
    if (m) s += static_cast<double>(v);
  }
  for (int k = 0; k < 4; ++k) {
    const unsigned long long t = block_sum_u(c[k], shu);
    if (threadIdx.x == 0) rows[y * 4 + k] = t;
  }
  const double t = block_sum(s, shd);
  if (threadIdx.x == 0) rowsum[y] = t;
}

// ---------------------------------------------------------------------------------------------- ranking metrics
__global__ void gather_scores_kernel(const float* __restrict__ x, View seg, int H, int W, int D,
                                     float* __restrict__ keys, unsigned int* __restrict__ labels) {
  const long long n = static_cast<long long>(H) * W * D;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const int xx = static_cast<int>(idx % W);
  const int y = static_cast<int>((idx / W) % H);
  const int d = static_cast<int>(idx / (static_cast<long long>(W) * H));
  keys[idx] = x[idx];
  labels[idx] = seg.at(y, xx, d) > 0.f ? 1 : 0;
}

struct MaxOp {
  __host__ __device__ int operator()(int a, int b) const { return a > b ? a : b; }
};

__global__ void end_index_kernel(const float* __restrict__ s, int n, int* __restrict__ endidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool end = (i == n - 1) || (s[i] != s[i + 1]);
  endidx[i] = end ? i : -1;
}

__global__ void ranking_terms_kernel(const float* __restrict__ s, const unsigned int* __restrict__ tps,
                                     const int* __restrict__ lastend, int n, double* __restrict__ auc_terms,
                                     double* __restrict__ ap_terms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool end = (i == n - 1) || (s[i] != s[i + 1]);
  double a = 0.0, p = 0.0;
  if (end) {
    const double P = static_cast<double>(tps[n - 1]);
    const double N = static_cast<double>(n) - P;
    const int j = (i > 0) ? lastend[i - 1] : -1;
    const double tp = tps[i], fp = static_cast<double>(i + 1) - tp;
    const double tp0 = (j >= 0) ? static_cast<double>(tps[j]) : 0.0;
    const double fp0 = (j >= 0) ? (static_cast<double>(j + 1) - tp0) : 0.0;
    a = (fp / N - fp0 / N) * (tp / P + tp0 / P) * 0.5;
    p = (tp / P - tp0 / P) * (tp / (tp + fp));
  }
  auc_terms[i] = a;
  ap_terms[i] = p;
}

// ---------------------------------------------------------------------------------------------- Dice bisection
// find_best_val (utils_eval.py:508-539) for ONE volume, on the device, in one launch of one warp: the ranking pass has
// left the scores sorted in descending order (keys) with the running count of positives (tps), so the counts of a
// threshold q are two look-ups - P = #(x > q) by a 32-ary search over the sorted scores (4 probes rounds for 460 800
// voxels), PG = tps[P - 1], G = tps[n - 1] - and the ten bisection decisions need no pass over the volume and no host
// round trip.  Arithmetic as the reference evaluates it under NumPy >= 2: the range in float32 (every operation rounded
// on its own), Dice = 2 PG / (P + G) in float64 from exact integer counts, NaN (0 / 0) compares false.  (When max(x)
// == 0 the reference restarts from the Python ints (0, 1) and computes in float64: the thresholds are then dyadic
// fractions, identical in float32.)
__device__ __forceinline__ int count_greater_desc(const float* __restrict__ keys, int n, float q, int lane) {
  int lo = 0, len = n;  // the answer lies in [lo, lo + len]
  while (len > 0) {
    const int chunk = (len + 31) / 32;
    const int idx = lo + (lane + 1) * chunk - 1;  // last element of this lane's chunk
    const bool gt = idx < lo + len ? (keys[idx] > q) : false;
    const unsigned ball = __ballot_sync(0xffffffffu, gt);
    const int full = __popc(ball);  // chunks that lie entirely above q (the sequence is descending)
    const int end = lo + len;
    lo += full * chunk;
    // the first chunk whose last element is <= q holds the boundary; that last element itself is out
    len = full == 32 ? 0 : max(0, min(chunk - 1, end - lo));
  }
  return lo;
}

__global__ void dice_bisect_kernel(const float* __restrict__ keys, const unsigned int* __restrict__ tps, int n,
                                   int max_steps, double* __restrict__ result) {
  const int lane = threadIdx.x;
  const long long G = tps[n - 1];
  float bottom = 0.f, top = keys[0];  // val_range = (0, np.max(x))
  double max_val = 0.0, max_point = 0.0;
  auto dice_at = [&](float q) {
    const int P = count_greater_desc(keys, n, q, lane);
    const long long PG = P > 0 ? static_cast<long long>(tps[P - 1]) : 0;
    return static_cast<double>(2 * PG) / static_cast<double>(static_cast<long long>(P) + G);
  };
  for (int step = 0; step < max_steps; ++step) {
    if (bottom == top) top = 1.0f;
    const float span = __fsub_rn(top, bottom);
    const float center = __fadd_rn(bottom, __fmul_rn(span, 0.5f));
    const float q_bottom = __fadd_rn(bottom, __fmul_rn(span, 0.25f));
    const float q_top = __fadd_rn(bottom, __fmul_rn(span, 0.75f));
    const double val_bottom = dice_at(q_bottom);
    const double val_top = dice_at(q_top);
    if (val_bottom >= val_top) {
      if (val_bottom >= max_val) {
        max_val = val_bottom;
        max_point = static_cast<double>(q_bottom);
      }
      top = center;
    } else {
      if (val_top >= max_val) {
        max_val = val_top;
        max_point = static_cast<double>(q_top);
      }
      bottom = center;
    }
  }
  if (lane == 0) {
    result[0] = max_val;
    result[1] = max_point;
    result[2] = static_cast<double>(keys[0]);  // np.max(x): 0 makes the reference continue in Python floats
  }
}

struct RankingLayout {
  size_t keys_in, keys_out, lab_in, lab_out, tps, endidx, lastend, auc, ap, cub, total, cub_bytes;
};
RankingLayout ranking_layout(int64_t n) {
  RankingLayout L{};
  auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
  size_t off = 0;
  L.keys_in = off; off = align(off + n * 4);
  L.keys_out = off; off = align(off + n * 4);
  L.lab_in = off; off = align(off + n * 4);
  L.lab_out = off; off = align(off + n * 4);
  L.tps = off; off = align(off + n * 4);
  L.endidx = off; off = align(off + n * 4);
  L.lastend = off; off = align(off + n * 4);
  L.auc = off; off = align(off + n * 8);
  L.ap = off; off = align(off + n * 8);
  size_t b_sort = 0, b_scan = 0, b_scan2 = 0, b_red = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, b_sort, (const float*)nullptr, (float*)nullptr,
                                            (const unsigned int*)nullptr, (unsigned int*)nullptr, static_cast<int>(n));
  cub::DeviceScan::InclusiveSum(nullptr, b_scan, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                static_cast<int>(n));
  cub::DeviceScan::InclusiveScan(nullptr, b_scan2, (const int*)nullptr, (int*)nullptr, MaxOp(), static_cast<int>(n));
  cub::DeviceReduce::Sum(nullptr, b_red, (const double*)nullptr, (double*)nullptr, static_cast<int>(n));
  L.cub_bytes = std::max(std::max(b_sort, b_scan), std::max(b_scan2, b_red));
  L.cub = off;
  off = align(off + L.cub_bytes);
  L.total = off;
  return L;
}

}  // namespace

int launch_residual_erode(const VolView& orig, const VolView& reco, const VolView& seg, const VolView& mask, int H,
                          int W, int D, int iterations, int erode, float* diff_masked, double* sums,
                          cudaStream_t stream) {
  if (!orig.p || !reco.p || !diff_masked) return fail(kInvalidArgument, "residual: null pointer");
  if (erode && !mask.p) return fail(kInvalidArgument, "residual: erosion needs the brain mask");
  if (sums) CDDPM_CUDA(cudaMemsetAsync(sums, 0, 7 * sizeof(double), stream));
  const long long n = static_cast<long long>(H) * W * D;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  residual_erode_kernel<<<blocks, 256, 0, stream>>>(dev_view(orig), dev_view(reco), dev_view(seg), dev_view(mask), H, W,
                                                    D, iterations, erode, diff_masked, sums);
  return check_launch("residual_erode_kernel");
}

int launch_trilinear_resize(const VolView& src, int H, int W, int D, float* dst, int Ho, int Wo, int Do,
                            cudaStream_t stream) {
  if (!src.p || !dst) return fail(kInvalidArgument, "trilinear_resize: null pointer");
  if (H < 1 || W < 1 || D < 1 || Ho < 1 || Wo < 1 || Do < 1) return fail(kInvalidArgument, "trilinear_resize: empty extent");
  const long long n = static_cast<long long>(Ho) * Wo * Do;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
  trilinear_kernel<<<blocks, 256, 0, stream>>>(dev_view(src), H, W, D, dst, Ho, Wo, Do);
  return check_launch("trilinear_kernel");
}

// ---- 5x5x5 median, second generation: forgetful selection over a shared-memory tile, two x-neighbours per thread.
// Forgetful selection (Perrot, Domas, Couturier 2014): the median of n keys is found by keeping a working set of
// n/2 + 2 keys, discarding its minimum and maximum (neither can be the median) and admitting the next unseen key,
// until one key is left; min / max of a set of S keys cost ~1.5 S compare-exchanges.
// Sharing: the windows of voxels (x, y, d) and (x + 1, y, d) have 100 of their 125 keys in common.  A common key of
// rank r among the 100 has rank r .. r + 25 in either window, so only ranks 37 .. 62 can be either median: the same
// procedure with a working set of 64 drops the 37 smallest and 37 largest common keys ONCE for both voxels; each
// voxel then needs the median of its 26 candidates + 25 own keys (rank 62 - 37 = 25 of 51).  ~2 x 1850 compare-
// exchanges per voxel against ~24 000 ALU instructions of the radix descent, every key in a register with a
// compile-time index.
__device__ __forceinline__ void cex(uint32_t& lo, uint32_t& hi) {
  const uint32_t a = lo, b = hi;
  lo = min(a, b);
  hi = max(a, b);
}
// minimum of a[0..S) -> a[0], maximum -> a[S-1]: pair the two halves, then two log-depth tournaments (the same S/2 +
// 2 (S/2 - 1) compare-exchanges as two linear scans, without their serial dependence chains)
template <int S, int N>
__device__ __forceinline__ void extract_minmax(uint32_t (&a)[N]) {
  constexpr int h = S / 2;
  constexpr int hm = (S + 1) / 2;  // an odd set's middle key takes part in both tournaments
#pragma unroll
  for (int i = 0; i < h; ++i) cex(a[i], a[S - 1 - i]);
#pragma unroll
  for (int stride = 1; stride < hm; stride *= 2)
#pragma unroll
    for (int i = 0; i + stride < hm; i += 2 * stride) {
      cex(a[i], a[i + stride]);
      cex(a[S - 1 - i - stride], a[S - 1 - i]);
    }
}
// From a working set a[0..S): drop min and max, admit next(S) into the freed slot 0, until the set has END keys; then
// drop min and max once more: the survivors are a[1 .. END-1).
template <int S, int END, int N, typename F>
__device__ __forceinline__ void forget(uint32_t (&a)[N], F&& next) {
  extract_minmax<S, N>(a);
  if constexpr (S > END) {
    a[0] = next(S);
    forget<S - 1, END, N>(a, next);
  }
}

__global__ void __launch_bounds__(256, 2) median5_pair_kernel(const float* __restrict__ in, float* __restrict__ out, int H,
                                                           int W, int D) {
  constexpr int TX = 32, TY = 4, TD = 4;  // 512 output voxels per CTA, thread = two x-neighbours
  constexpr int SX = TX + 4, SY = TY + 4, SD = TD + 4;
  constexpr uint32_t zero_key = 0x80000000u;
  __shared__ uint32_t tile[SD][SY][SX];
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, d0 = blockIdx.z * TD;
  int nonzero = 0;
  for (int i = threadIdx.x; i < SD * SY * SX; i += blockDim.x) {
    const int tx = i % SX, ty = (i / SX) % SY, td = i / (SX * SY);
    const int gx = reflect(x0 + tx - 2, W), gy = reflect(y0 + ty - 2, H), gd = reflect(d0 + td - 2, D);
    const uint32_t kv = sort_key(in[(static_cast<size_t>(gd) * H + gy) * W + gx]);
    tile[td][ty][tx] = kv;
    nonzero |= (kv != zero_key);
  }
  const int any = __syncthreads_or(nonzero);
  const int lx = (threadIdx.x % 16) * 2, ly = (threadIdx.x / 16) % TY, ld = threadIdx.x / (16 * TY);
  const int x = x0 + lx, y = y0 + ly, d = d0 + ld;
  if (x >= W || y >= H || d >= D) return;
  const size_t oidx = (static_cast<size_t>(d) * H + y) * W + x;
  const bool has_b = x + 1 < W;
  if (!any) {  // the whole tile is +0 (outside the eroded brain mask)
    out[oidx] = 0.f;
    if (has_b) out[oidx + 1] = 0.f;
    return;
  }
  // the residual volume is >= 0 and mostly exactly 0: the median is 0 as soon as more than half of a window is 0 and
  // nothing in it is negative
  {
    int zc = 0, za = 0, zb = 0, neg = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = 0; b < 5; ++b) {
#pragma unroll
        for (int c = 1; c < 5; ++c) {
          const uint32_t kv = tile[ld + a][ly + b][lx + c];
          zc += (kv == zero_key);
          neg += (kv < zero_key);
        }
        const uint32_t ka = tile[ld + a][ly + b][lx], kb = tile[ld + a][ly + b][lx + 5];
        za += (ka == zero_key);
        zb += (kb == zero_key);
        neg += (ka < zero_key) + (kb < zero_key);
      }
    if (neg == 0 && zc + za > 62 && zc + zb > 62) {
      out[oidx] = 0.f;
      if (has_b) out[oidx + 1] = 0.f;
      return;
    }
  }
  // ---- the 100 common keys (x offsets 1..4): 64 in the working set, 36 admitted one by one -> 26 candidates
  auto common = [&](int j) { return tile[ld + j / 20][ly + (j / 4) % 5][lx + 1 + j % 4]; };
  uint32_t w[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) w[j] = common(j);
  forget<64, 28, 64>(w, [&](int S) { return common(128 - S); });
  // ---- each voxel: median of the 26 candidates w[1..26] + its own 25 keys (x offset 0 resp. 5)
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int xo = lx + 5 * v;
    auto own = [&](int e) { return tile[ld + e / 5][ly + e % 5][xo]; };
    uint32_t m[27];
#pragma unroll
    for (int j = 0; j < 26; ++j) m[j] = w[1 + j];
    m[26] = own(0);
    forget<27, 3, 27>(m, [&](int S) { return own(28 - S); });
    if (v == 0 || has_b) out[oidx + v] = key_value(m[1]);
  }
}

static bool median_v2_enabled() {
  static const bool on = [] {
    const char* e = getenv("CDDPM_MEDIAN_V2");  // A/B switch for measurements: 0 = the radix-descent kernel
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int launch_median3d(const float* in, float* out, int H, int W, int D, int k, cudaStream_t stream) {
  if (!in || !out) return fail(kInvalidArgument, "median: null pointer");
  if (in == out) return fail(kInvalidArgument, "median: in-place filtering is not supported");
  dim3 grid((W + 15) / 16, (H + 3) / 4, (D + 3) / 4);
  if (k == 5 && median_v2_enabled()) {
    median5_pair_kernel<<<dim3((W + 31) / 32, (H + 3) / 4, (D + 3) / 4), 256, 0, stream>>>(in, out, H, W, D);
  } else if (k == 5) {
    median3d_kernel<5><<<grid, 256, 0, stream>>>(in, out, H, W, D);
  } else if (k == 3) {
    median3d_kernel<3><<<grid, 256, 0, stream>>>(in, out, H, W, D);
  } else if (k == 1) {
    return check_cuda(cudaMemcpyAsync(out, in, static_cast<size_t>(H) * W * D * 4, cudaMemcpyDeviceToDevice, stream),
                      "median k=1 copy");
  } else {
    return fail(kUnsupported, "median: kernel size must be 1, 3 or 5");
  }
  return check_launch("median3d_kernel");
}

// ---- output image grid (log_images, utils_eval.py:586-628): one RGB row of four panels for one axial slice
// Eight anchor colours of matplotlib's 'inferno' at equal spacing (its published 8-class palette), interpolated
// linearly; matplotlib itself is absent here, so the colour table is an approximation.
__constant__ float kInferno[8][3] = {{0.f, 0.f, 4.f},      {40.f, 11.f, 84.f},   {101.f, 21.f, 110.f}, {159.f, 42.f, 99.f},
                                     {212.f, 72.f, 66.f},  {245.f, 125.f, 21.f}, {250.f, 193.f, 39.f}, {252.f, 255.f, 164.f}};
// panels [4][H][W] fp32 (original, reconstruction, difference, segmentation), ranges [4][2] = (vmin, vmax) of each
// panel's Normalize; every panel is drawn rotated by torch.rot90(., 3) (out[i][j] = in[H-1-j][i], shape [W][H]) and the
// four are laid side by side: rgb [W][4 H][3] uint8.  Panel 2 uses the inferno table, the others 'gray'.
__global__ void compose_grid_kernel(const float* __restrict__ panels, const float* __restrict__ ranges, int H, int W,
                                    uint8_t* __restrict__ rgb) {
  const int total = W * 4 * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int col = i % (4 * H), row = i / (4 * H);
    const int pnl = col / H, j = col - pnl * H;
    const float v = panels[(static_cast<size_t>(pnl) * H + (H - 1 - j)) * W + row];
    const float lo = ranges[2 * pnl], hi = ranges[2 * pnl + 1];
    float t = hi > lo ? (v - lo) / (hi - lo) : 0.f;  // matplotlib maps a constant image to the bottom of the table
    t = fminf(fmaxf(t, 0.f), 1.f);
    float c[3];
    if (pnl == 2) {
      const float u = t * 7.f;
      const int k = min(static_cast<int>(u), 6);
      const float f = u - static_cast<float>(k);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) c[ch] = kInferno[k][ch] + f * (kInferno[k + 1][ch] - kInferno[k][ch]);
    } else {
      c[0] = c[1] = c[2] = t * 255.f;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) rgb[static_cast<size_t>(i) * 3 + ch] = static_cast<uint8_t>(__float2int_rn(c[ch]));
  }
}

int launch_compose_grid(const float* panels, const float* ranges, int H, int W, uint8_t* rgb, cudaStream_t stream) {
  if (!panels || !ranges || !rgb) return fail(kInvalidArgument, "compose_grid: null pointer");
  if (H < 1 || W < 1) return fail(kInvalidArgument, "compose_grid: empty image");
  const int total = W * 4 * H;
  compose_grid_kernel<<<std::min((total + 255) / 256, 148 * 8), 256, 0, stream>>>(panels, ranges, H, W, rgb);
  return check_launch("compose_grid_kernel");
}

int launch_max(const float* x, int64_t n, float* out_max, cudaStream_t stream) {
  if (!x || !out_max) return fail(kInvalidArgument, "max: null pointer");
  CDDPM_CUDA(cudaMemsetAsync(out_max, 0, 4, stream));
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 4));
  max_key_kernel<<<blocks, 256, 0, stream>>>(x, n, reinterpret_cast<unsigned int*>(out_max));
  max_decode_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned int*>(out_max));
  return check_launch("max kernels");
}

int launch_threshold_counts(const float* x, const VolView& seg, int H, int W, int D, const float* q_host, int nq,
                            unsigned long long* counts, cudaStream_t stream) {
  if (!x || !seg.p || !q_host || !counts) return fail(kInvalidArgument, "threshold_counts: null pointer");
  if (nq < 1 || nq > 4) return fail(kInvalidArgument, "threshold_counts: 1..4 thresholds per call");
  QList ql{};
  ql.nq = nq;
  for (int i = 0; i < nq; ++i) ql.q[i] = q_host[i];
  const long long n = static_cast<long long>(H) * W * D;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 4));
  threshold_counts_kernel<<<blocks, 256, 0, stream>>>(x, dev_view(seg), H, W, D, ql, counts);
  return check_launch("threshold_counts_kernel");
}

int launch_threshold_mask(const float* x, int64_t n, float thr, unsigned char* out, cudaStream_t stream) {
  if (!x || !out) return fail(kInvalidArgument, "threshold_mask: null pointer");
  threshold_mask_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(x, n, thr, out);
  return check_launch("threshold_mask_kernel");
}

int launch_row_stats(const float* x, const VolView& seg, const VolView& mask, int H, int W, int D, float thr,
                     unsigned long long* rows, double* rowsum, cudaStream_t stream) {
  if (!x || !mask.p || !rows || !rowsum) return fail(kInvalidArgument, "row_stats: null pointer");
  row_stats_kernel<<<H, 256, 0, stream>>>(x, dev_view(seg), dev_view(mask), H, W, D, thr, rows, rowsum);
  return check_launch("row_stats_kernel");
}

size_t ranking_workspace_bytes(int64_t n) { return ranking_layout(n).total; }

int launch_ranking_metrics(const float* x, const VolView& seg, int H, int W, int D, void* workspace,
                           size_t workspace_bytes, double* result, cudaStream_t stream) {
  if (!x || !seg.p || !workspace || !result) return fail(kInvalidArgument, "ranking: null pointer");
  const int64_t n64 = static_cast<int64_t>(H) * W * D;
  if (n64 > (1ll << 30)) return fail(kUnsupported, "ranking: too many voxels");
  const int n = static_cast<int>(n64);
  const RankingLayout L = ranking_layout(n);
  if (workspace_bytes < L.total) return fail(kInvalidArgument, "ranking: workspace too small");
  char* ws = reinterpret_cast<char*>(workspace);
  float* keys_in = reinterpret_cast<float*>(ws + L.keys_in);
  float* keys_out = reinterpret_cast<float*>(ws + L.keys_out);
  unsigned int* lab_in = reinterpret_cast<unsigned int*>(ws + L.lab_in);
  unsigned int* lab_out = reinterpret_cast<unsigned int*>(ws + L.lab_out);
  unsigned int* tps = reinterpret_cast<unsigned int*>(ws + L.tps);
  int* endidx = reinterpret_cast<int*>(ws + L.endidx);
  int* lastend = reinterpret_cast<int*>(ws + L.lastend);
  double* auc = reinterpret_cast<double*>(ws + L.auc);
  double* ap = reinterpret_cast<double*>(ws + L.ap);
  void* cubtmp = ws + L.cub;
  size_t cb = L.cub_bytes;
  const int blocks = (n + 255) / 256;
  gather_scores_kernel<<<blocks, 256, 0, stream>>>(x, dev_view(seg), H, W, D, keys_in, lab_in);
  CDDPM_TRY(check_launch("gather_scores_kernel"));
  CDDPM_CUDA(cub::DeviceRadixSort::SortPairsDescending(cubtmp, cb, keys_in, keys_out, lab_in, lab_out, n, 0, 32, stream));
  cb = L.cub_bytes;
  CDDPM_CUDA(cub::DeviceScan::InclusiveSum(cubtmp, cb, lab_out, tps, n, stream));
  end_index_kernel<<<blocks, 256, 0, stream>>>(keys_out, n, endidx);
  CDDPM_TRY(check_launch("end_index_kernel"));
  cb = L.cub_bytes;
  CDDPM_CUDA(cub::DeviceScan::InclusiveScan(cubtmp, cb, endidx, lastend, MaxOp(), n, stream));
  ranking_terms_kernel<<<blocks, 256, 0, stream>>>(keys_out, tps, lastend, n, auc, ap);
  CDDPM_TRY(check_launch("ranking_terms_kernel"));
  cb = L.cub_bytes;
  CDDPM_CUDA(cub::DeviceReduce::Sum(cubtmp, cb, auc, result, n, stream));
  cb = L.cub_bytes;
  CDDPM_CUDA(cub::DeviceReduce::Sum(cubtmp, cb, ap, result + 1, n, stream));
  return kOk;
}

int launch_dice_bisect(const void* ranking_workspace, int64_t n64, int max_steps, double* result, cudaStream_t stream) {
  if (!ranking_workspace || !result) return fail(kInvalidArgument, "dice_bisect: null pointer");
  if (n64 < 1 || n64 > (1ll << 30)) return fail(kInvalidArgument, "dice_bisect: bad voxel count");
  if (max_steps < 0 || max_steps > 64) return fail(kInvalidArgument, "dice_bisect: bad step count");
  const RankingLayout L = ranking_layout(n64);
  const char* ws = reinterpret_cast<const char*>(ranking_workspace);
  dice_bisect_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const float*>(ws + L.keys_out),
                                           reinterpret_cast<const unsigned int*>(ws + L.tps), static_cast<int>(n64),
                                           max_steps, result);
  return check_launch("dice_bisect_kernel");
}

}  // namespace cddpm
