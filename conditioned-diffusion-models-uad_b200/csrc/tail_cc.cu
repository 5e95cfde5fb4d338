// Anomaly-scoring tail, SURVEY.md §8 f-4: what remained on the host after the median / threshold kernels.
//   * filter_3d_connected_components (src/utils/utils_eval.py:489-503): skimage label(connectivity=3) +
//     regionprops.filled_area <= 7.  scikit-image fills holes with a full 3x3x3 structuring element, so a hole would need
//     all 26 neighbours of a voxel set: a component of <= 7 voxels has none and filled_area == area there, while area > 7
//     implies filled_area > 7.  The filter is therefore "drop 26-connected components of at most 7 voxels", which needs
//     no global labelling: every foreground voxel walks its own component and stops at the 8th voxel.
//   * confusion counts of the filtered prediction against seg > 0 (utils_eval.py:104-110).
//   * monai.metrics.compute_hausdorff_distance(pred, seg, euclidean, percentile=None, directed=False)
//     (utils_eval.py:134; monai 0.9: get_mask_edges = mask ^ binary_erosion(mask) with the 6-neighbour cross,
//     get_surface_distance = scipy distance_transform_edt of the complement of the other edge set).  Integer-exact:
//     squared distances stay int32 through a separable min-plus transform; the host takes one float64 sqrt.
// All index / integer work: bit-exact against the oracle (oracle/tail_port.py).
#include <stdint.h>

#include <algorithm>

#include "common.h"
#include "tail.cuh"

namespace cddpm {
namespace {

constexpr int kMaxComp = 15;       // largest supported size threshold
constexpr int kInf = 1 << 28;      // "no edge voxel on this line"; + 3 * 2^16 stays inside int32

struct View {
  const float* p;
  long long sy, sx, sd;
  __device__ __forceinline__ float at(int y, int x, int d) const { return p[y * sy + x * sx + d * sd]; }
};
View dev_view(const VolView& v) { return View{v.p, v.sy, v.sx, v.sd}; }

__global__ void filter_small_components_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                               int H, int W, int D, int max_size) {
  const int n = H * W * D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (in[i] == 0) {
    out[i] = 0;
    return;
  }
  if (max_size < 1) {
    out[i] = 1;
    return;
  }
  int list[kMaxComp + 1];
  int count = 1, head = 0;
  list[0] = i;
  bool big = false;
  while (head < count && !big) {
    const int v = list[head++];
    const int x = v % W, y = (v / W) % H, d = v / (W * H);
    for (int dd = -1; dd <= 1 && !big; ++dd) {
      const int d2 = d + dd;
      if (d2 < 0 || d2 >= D) continue;
      for (int dy = -1; dy <= 1 && !big; ++dy) {
        const int y2 = y + dy;
        if (y2 < 0 || y2 >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int x2 = x + dx;
          if (x2 < 0 || x2 >= W) continue;
          const int u = (d2 * H + y2) * W + x2;
          if (in[u] == 0) continue;
          bool seen = false;
          for (int k = 0; k < count; ++k) seen = seen || (list[k] == u);
          if (seen) continue;
          if (count == max_size) {  // one voxel more than the threshold: the component stays
            big = true;
            break;
          }
          list[count++] = u;
        }
      }
    }
  }
  out[i] = big ? 1 : 0;
}

__global__ void confusion_counts_kernel(const unsigned char* __restrict__ pred, View seg, int H, int W, int D,
                                        unsigned long long* __restrict__ counts) {
  const long long n = static_cast<long long>(H) * W * D;
  unsigned int c11 = 0, c10 = 0, c01 = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W), y = static_cast<int>((i / W) % H), d = static_cast<int>(i / (W * H));
    const bool p = pred[i] != 0, g = seg.at(y, x, d) > 0.f;
    c11 += p && g;
    c10 += p && !g;
    c01 += !p && g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c11 += __shfl_xor_sync(0xffffffffu, c11, o);
    c10 += __shfl_xor_sync(0xffffffffu, c10, o);
    c01 += __shfl_xor_sync(0xffffffffu, c01, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c11) atomicAdd(counts + 0, static_cast<unsigned long long>(c11));
    if (c10) atomicAdd(counts + 1, static_cast<unsigned long long>(c10));
    if (c01) atomicAdd(counts + 2, static_cast<unsigned long long>(c01));
  }
}

// which: 0 = the prediction buffer, 1 = seg > 0 through its view
__device__ __forceinline__ bool fg_at(const unsigned char* pred, const View& seg, int which, int y, int x, int d, int H,
                                      int W, int D) {
  if (y < 0 || y >= H || x < 0 || x >= W || d < 0 || d >= D) return false;
  return which == 0 ? pred[(d * H + y) * W + x] != 0 : seg.at(y, x, d) > 0.f;
}

// edge = mask and not eroded(mask) with the 6-neighbour cross and a zero border (scipy binary_erosion defaults).
// Writes the first min-plus pass input directly: 0 on an edge voxel, kInf elsewhere; counts edge voxels.
__global__ void edge_seed_kernel(const unsigned char* __restrict__ pred, View seg, int H, int W, int D,
                                 int* __restrict__ seed_pred, int* __restrict__ seed_gt,
                                 long long* __restrict__ result) {
  const int n = H * W * D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int ep = 0, eg = 0;
  if (i < n) {
    const int x = i % W, y = (i / W) % H, d = i / (W * H);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bool e = false;
      if (fg_at(pred, seg, which, y, x, d, H, W, D)) {
        e = !(fg_at(pred, seg, which, y - 1, x, d, H, W, D) && fg_at(pred, seg, which, y + 1, x, d, H, W, D) &&
              fg_at(pred, seg, which, y, x - 1, d, H, W, D) && fg_at(pred, seg, which, y, x + 1, d, H, W, D) &&
              fg_at(pred, seg, which, y, x, d - 1, H, W, D) && fg_at(pred, seg, which, y, x, d + 1, H, W, D));
      }
      (which == 0 ? seed_pred : seed_gt)[i] = e ? 0 : kInf;
      (which == 0 ? ep : eg) = e ? 1 : 0;
    }
  }
  const unsigned bp = __ballot_sync(0xffffffffu, ep), bg = __ballot_sync(0xffffffffu, eg);
  if ((threadIdx.x & 31) == 0) {
    if (bp) atomicAdd(reinterpret_cast<unsigned long long*>(result + 2), static_cast<unsigned long long>(__popc(bp)));
    if (bg) atomicAdd(reinterpret_cast<unsigned long long*>(result + 3), static_cast<unsigned long long>(__popc(bg)));
  }
}

// One axis of the exact squared Euclidean distance transform as a brute-force min-plus product over a line of
// length L <= 96: out(p) = min_j in(line(p), j) + (pos(p) - j)^2.  `stride` is the element stride of the axis.
// Two fields (blockIdx.y) per launch.
__global__ void edt_axis_kernel(const int* __restrict__ in0, int* __restrict__ out0, const int* __restrict__ in1,
                                int* __restrict__ out1, int n, int L, int stride) {
  const int* in = blockIdx.y == 0 ? in0 : in1;
  int* out = blockIdx.y == 0 ? out0 : out1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = (i / stride) % L;
  const int base = i - pos * stride;
  int best = kInf;
  for (int j = 0; j < L; ++j) {
    const int v = in[base + j * stride];
    const int dj = pos - j;
    best = min(best, v + dj * dj);
  }
  out[i] = best >= kInf ? kInf : best;
}

// result[0] = max over pred-edge voxels of dist^2 to the gt edge set, result[1] = the reverse direction.
__global__ void edge_max_kernel(const int* __restrict__ seed_pred, const int* __restrict__ seed_gt,
                                const int* __restrict__ dist_pred, const int* __restrict__ dist_gt, int n,
                                long long* __restrict__ result) {
  int m0 = -1, m1 = -1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (seed_pred[i] == 0) m0 = max(m0, dist_gt[i]);
    if (seed_gt[i] == 0) m1 = max(m1, dist_pred[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (m0 >= 0) atomicMax(result + 0, static_cast<long long>(m0));
    if (m1 >= 0) atomicMax(result + 1, static_cast<long long>(m1));
  }
}

__global__ void hausdorff_init_kernel(long long* result) {
  result[0] = -1;
  result[1] = -1;
  result[2] = 0;
  result[3] = 0;
}

}  // namespace

int launch_filter_small_components(const unsigned char* in, unsigned char* out, int H, int W, int D, int max_size,
                                   cudaStream_t stream) {
  if (!in || !out) return fail(kInvalidArgument, "filter_small_components: null pointer");
  if (in == out) return fail(kInvalidArgument, "filter_small_components: in-place filtering is not supported");
  if (max_size < 0 || max_size > kMaxComp)
    return fail(kInvalidArgument, "filter_small_components: size threshold must be in 0..15");
  const long long n = static_cast<long long>(H) * W * D;
  if (n <= 0) return kOk;
  if (n > (1ll << 30)) return fail(kInvalidArgument, "filter_small_components: volume too large");
  filter_small_components_kernel<<<static_cast<int>((n + 127) / 128), 128, 0, stream>>>(in, out, H, W, D, max_size);
  return check_launch("filter_small_components_kernel");
}

int launch_confusion_counts(const unsigned char* pred, const VolView& seg, int H, int W, int D,
                            unsigned long long* counts, cudaStream_t stream) {
  if (!pred || !seg.p || !counts) return fail(kInvalidArgument, "confusion_counts: null pointer");
  const long long n = static_cast<long long>(H) * W * D;
  if (n <= 0) return kOk;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  confusion_counts_kernel<<<blocks, 256, 0, stream>>>(pred, dev_view(seg), H, W, D, counts);
  return check_launch("confusion_counts_kernel");
}

size_t hausdorff_workspace_bytes(int H, int W, int D) {
  return static_cast<size_t>(H) * W * D * sizeof(int) * 6;  // two seeds, two ping-pong pairs
}

int launch_hausdorff(const unsigned char* pred, const VolView& seg, int H, int W, int D, void* workspace,
                     size_t workspace_bytes, long long* result, cudaStream_t stream) {
  if (!pred || !seg.p || !workspace || !result) return fail(kInvalidArgument, "hausdorff: null pointer");
  if (H < 1 || W < 1 || D < 1 || H > 4096 || W > 4096 || D > 4096)
    return fail(kInvalidArgument, "hausdorff: extents must be in 1..4096");
  const long long n64 = static_cast<long long>(H) * W * D;
  if (n64 > (1ll << 28)) return fail(kInvalidArgument, "hausdorff: volume too large");
  if (workspace_bytes < hausdorff_workspace_bytes(H, W, D)) return fail(kInvalidArgument, "hausdorff: workspace too small");
  const int n = static_cast<int>(n64);
  int* seed_p = reinterpret_cast<int*>(workspace);
  int* seed_g = seed_p + n;
  int* a_p = seed_g + n;
  int* a_g = a_p + n;
  int* b_p = a_g + n;
  int* b_g = b_p + n;
  const int blocks = (n + 255) / 256;
  hausdorff_init_kernel<<<1, 1, 0, stream>>>(result);
  CDDPM_TRY(check_launch("hausdorff_init_kernel"));
  edge_seed_kernel<<<blocks, 256, 0, stream>>>(pred, dev_view(seg), H, W, D, seed_p, seed_g, result);
  CDDPM_TRY(check_launch("edge_seed_kernel"));
  const dim3 grid(blocks, 2);
  edt_axis_kernel<<<grid, 256, 0, stream>>>(seed_p, a_p, seed_g, a_g, n, W, 1);
  CDDPM_TRY(check_launch("edt_axis_kernel (x)"));
  edt_axis_kernel<<<grid, 256, 0, stream>>>(a_p, b_p, a_g, b_g, n, H, W);
  CDDPM_TRY(check_launch("edt_axis_kernel (y)"));
  edt_axis_kernel<<<grid, 256, 0, stream>>>(b_p, a_p, b_g, a_g, n, D, W * H);
  CDDPM_TRY(check_launch("edt_axis_kernel (d)"));
  edge_max_kernel<<<std::min(blocks, 148 * 8), 256, 0, stream>>>(seed_p, seed_g, a_p, a_g, n, result);
  return check_launch("edge_max_kernel");
}

}  // namespace cddpm
