// GPU-resident preprocessing of a volume: the once-per-volume transforms the reference's datamodule
// applies on the host through torchio / SimpleITK (src/datamodules/create_dataset.py:196-218):
//   tio.CropOrPad((h, w, d), padding_mode=0)
//   tio.RescaleIntensity((0, 1), percentiles=(1, 99), masking_method='mask')
//   tio.Resample(rescaleFactor, image_interpolation='bspline')   (label maps: nearest neighbour)
// Volumes are contiguous [H][W][D] fp32 (the layout of the reference's [C,H,W,D] tensors for one channel).  These are
// bandwidth- and latency-bound one-off kernels (a 192x192x100 volume is 14.7 MB); the point is that the volume never
// leaves HBM between the loader and the UNet.
#include "preprocess.cuh"

#include <algorithm>
#include <cub/cub.cuh>

namespace cddpm {

namespace {

// ---------------------------------------------------------------------------------------------- CropOrPad
__global__ void __launch_bounds__(256) crop_or_pad_kernel(const float* __restrict__ in, int H, int W, int D,
                                                          float* __restrict__ out, int h, int w, int d, int oy, int ox,
                                                          int od, float pad) {
  const long long n = static_cast<long long>(h) * w * d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(i % d);
    const int x = static_cast<int>((i / d) % w);
    const int y = static_cast<int>(i / (static_cast<long long>(d) * w));
    const int sy = y + oy, sx = x + ox, sz = z + od;
    out[i] = (sy >= 0 && sy < H && sx >= 0 && sx < W && sz >= 0 && sz < D)
                 ? in[(static_cast<long long>(sy) * W + sx) * D + sz]
                 : pad;
  }
}

// ---------------------------------------------------------------------------------------------- RescaleIntensity
struct RescaleScalars {
  unsigned long long count;  // voxels under the mask
  double cutoff[2];          // np.percentile(values, percentiles) in float64
  unsigned int min_key, max_key;  // order-preserving keys of min / max of the clipped array
  int skip;                  // empty mask (or, later, zero range): the volume is left unchanged
};

__device__ __forceinline__ unsigned int fkey(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(256) mask_keys_kernel(const float* __restrict__ vol, const float* __restrict__ mask,
                                                        long long n, float* __restrict__ keys,
                                                        RescaleScalars* __restrict__ sc) {
  unsigned long long c = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool m = mask[i] > 0.f;
    keys[i] = m ? vol[i] : __int_as_float(0x7f800000);  // +inf sorts behind every masked value
    c += m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sc->count, c);
}

// np.percentile(values, (p_lo, p_hi)) with the default 'linear' method as NumPy 2.x evaluates it for a float32 array
// and a float64 array of quantiles: virtual index (n - 1) * (p / 100) in float64, gamma = its fractional part,
// neighbours a <= b taken from the sorted float32 values, their difference ROUNDED to float32, then
// a + diff * gamma (gamma < 0.5) or b - diff * (1 - gamma) in float64 (numpy/lib/_function_base_impl.py: _lerp).
__global__ void percentile_kernel(const float* __restrict__ sorted, double p_lo, double p_hi, RescaleScalars* sc) {
  const long long n = static_cast<long long>(sc->count);
  if (n == 0) {
    sc->skip = 1;
    return;
  }
  const double p[2] = {p_lo, p_hi};
  for (int k = 0; k < 2; ++k) {
    const double q = __ddiv_rn(p[k], 100.0);
    const double virt = __dmul_rn(static_cast<double>(n - 1), q);
    long long prev = static_cast<long long>(floor(virt));
    const double gamma = __dsub_rn(virt, static_cast<double>(prev));
    long long next = prev + 1;
    prev = min(max(prev, 0ll), n - 1);
    next = min(max(next, 0ll), n - 1);
    const float a = sorted[prev], b = sorted[next];
    const double diff = static_cast<double>(__fsub_rn(b, a));
    sc->cutoff[k] = gamma >= 0.5 ? __dsub_rn(static_cast<double>(b), __dmul_rn(diff, __dsub_rn(1.0, gamma)))
                                 : __dadd_rn(static_cast<double>(a), __dmul_rn(diff, gamma));
  }
  sc->min_key = 0xFFFFFFFFu;
  sc->max_key = 0u;
  sc->skip = 0;
}

// np.clip(array, lo, hi, out=array) with float64 bounds on a float32 array: min(max(x, lo), hi) in float64, stored
// back as float32; plus the min / max of the clipped array.
__global__ void __launch_bounds__(256) clip_minmax_kernel(float* __restrict__ vol, long long n, RescaleScalars* sc) {
  if (sc->skip) return;
  const double lo = sc->cutoff[0], hi = sc->cutoff[1];
  unsigned int kmin = 0xFFFFFFFFu, kmax = 0u;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = static_cast<float>(fmin(fmax(static_cast<double>(vol[i]), lo), hi));
    vol[i] = v;
    const unsigned int k = fkey(v);
    kmin = min(kmin, k);
    kmax = max(kmax, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&sc->min_key, kmin);
    atomicMax(&sc->max_key, kmax);
  }
}

// array -= in_min; array /= in_range; array *= out_range; array += out_min - four separately rounded float32 steps
__global__ void __launch_bounds__(256) rescale_kernel(float* __restrict__ vol, long long n, float out_min, float out_max,
                                                      const RescaleScalars* __restrict__ sc) {
  if (sc->skip) return;
  const float in_min = fkey_inv(sc->min_key), in_max = fkey_inv(sc->max_key);
  const float in_range = __fsub_rn(in_max, in_min);
  if (in_range == 0.f) return;  // torchio warns and returns the (clipped-from) input; see RescaleIntensity in preprocess.py
  const float out_range = __fsub_rn(out_max, out_min);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = __fsub_rn(vol[i], in_min);
    v = __fdiv_rn(v, in_range);
    v = __fmul_rn(v, out_range);
    vol[i] = __fadd_rn(v, out_min);
  }
}

struct RescaleLayout {
  size_t keys_in, keys_out, scalars, cub, cub_bytes, total;
};
RescaleLayout rescale_layout(int64_t n) {
  RescaleLayout L{};
  auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
  size_t off = 0;
  L.keys_in = off; off = align(off + n * 4);
  L.keys_out = off; off = align(off + n * 4);
  L.scalars = off; off = align(off + sizeof(RescaleScalars));
  size_t b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, b, (const float*)nullptr, (float*)nullptr, static_cast<int>(n));
  L.cub_bytes = b;
  L.cub = off;
  off = align(off + b);
  L.total = off;
  return L;
}

// ---------------------------------------------------------------------------------------------- B-spline resampling
// Cubic B-spline decomposition of one axis, in place on float64 coefficients: gain 6, one pole z = sqrt(3) - 2, causal
// and anti-causal recursions with the exact mirror (whole-sample symmetric) initial values - the algorithm of
// scipy.ndimage.spline_filter1d(order=3, mode='mirror') (ni_splines.c) and, up to its 1e-10 truncated initialisation,
// of itk::BSplineDecompositionImageFilter behind SimpleITK's sitkBSpline interpolator.
__global__ void __launch_bounds__(128) bspline_prefilter_kernel(double* __restrict__ c, int n_lines, int len,
                                                                long long stride, int inner, long long inner_stride,
                                                                long long outer_stride) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines || len < 2) return;
  // lines are indexed (outer, inner): first element at outer * outer_stride + inner * inner_stride
  double* p = c + static_cast<long long>(line / inner) * outer_stride + static_cast<long long>(line % inner) * inner_stride;
  const double z = sqrt(3.0) - 2.0;
  const double lambda = (1.0 - z) * (1.0 - 1.0 / z);
  for (int i = 0; i < len; ++i) p[i * stride] *= lambda;
  // causal initial value, mirror boundary: exact sum
  {
    const double z_n_1 = pow(z, static_cast<double>(len - 1));
    double z_i = z;
    double c0 = p[0] + z_n_1 * p[(len - 1) * stride];
    for (int i = 1; i < len - 1; ++i) {
      c0 += z_i * (p[i * stride] + z_n_1 * p[(len - 1 - i) * stride]);
      z_i *= z;
    }
    p[0] = c0 / (1.0 - z_n_1 * z_n_1);
  }
  for (int i = 1; i < len; ++i) p[i * stride] += z * p[(i - 1) * stride];
  p[(len - 1) * stride] = (z * p[(len - 2) * stride] + p[(len - 1) * stride]) * z / (z * z - 1.0);
  for (int i = len - 2; i >= 0; --i) p[i * stride] = z * (p[(i + 1) * stride] - p[i * stride]);
}

__global__ void __launch_bounds__(256) to_double_kernel(const float* __restrict__ in, double* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = static_cast<double>(in[i]);
}

__device__ __forceinline__ int mirror_index(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i = i < 0 ? -i : i;
  i %= period;
  return i >= n ? period - i : i;
}

// Output voxel (i, j, k) samples the input at continuous index x = 0.5 (f - 1) + f * i per axis (voxel EDGES of the two
// grids coincide: torchio.Resample's reference image).  A sample outside [-0.5, N - 0.5) is outside the image buffer
// and reads the default value 0 (itk::ResampleImageFilter).  kBspline: cubic B-spline over mirrored coefficients;
// otherwise nearest neighbour with halves rounded up (label maps).
template <bool kBspline>
__global__ void __launch_bounds__(256) resample_kernel(const double* __restrict__ coef, const float* __restrict__ in,
                                                       int H, int W, int D, float* __restrict__ out, int h, int w, int d,
                                                       double fy, double fx, double fz) {
  const long long n = static_cast<long long>(h) * w * d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(i % d);
    const int x = static_cast<int>((i / d) % w);
    const int y = static_cast<int>(i / (static_cast<long long>(d) * w));
    const double cy = 0.5 * (fy - 1.0) + fy * y, cx = 0.5 * (fx - 1.0) + fx * x, cz = 0.5 * (fz - 1.0) + fz * z;
    if (cy < -0.5 || cy >= H - 0.5 || cx < -0.5 || cx >= W - 0.5 || cz < -0.5 || cz >= D - 0.5) {
      out[i] = 0.f;
      continue;
    }
    if (!kBspline) {
      const int iy = min(static_cast<int>(floor(cy + 0.5)), H - 1), ix = min(static_cast<int>(floor(cx + 0.5)), W - 1),
                iz = min(static_cast<int>(floor(cz + 0.5)), D - 1);
      out[i] = in[(static_cast<long long>(iy) * W + ix) * D + iz];
      continue;
    }
    const double c3[3] = {cy, cx, cz};
    const int N3[3] = {H, W, D};
    int idx[3][4];
    double wgt[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int base = static_cast<int>(floor(c3[a]));
      const double t = c3[a] - base;
      // cubic B-spline weights at offsets -1, 0, 1, 2 from floor(x)
      wgt[a][0] = (1.0 - t) * (1.0 - t) * (1.0 - t) / 6.0;
      wgt[a][1] = (4.0 - 6.0 * t * t + 3.0 * t * t * t) / 6.0;
      wgt[a][2] = (1.0 + 3.0 * t + 3.0 * t * t - 3.0 * t * t * t) / 6.0;
      wgt[a][3] = t * t * t / 6.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) idx[a][k] = mirror_index(base - 1 + k, N3[a]);
    }
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double sa = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double* row = coef + (static_cast<long long>(idx[0][a]) * W + idx[1][b]) * D;
        double sb = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) sb += wgt[2][k] * row[idx[2][k]];
        sa += wgt[1][b] * sb;
      }
      acc += wgt[0][a] * sa;
    }
    out[i] = static_cast<float>(acc);
  }
}

int grid_for(long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16)); }

}  // namespace

void crop_or_pad_offsets(int source, int target, int* offset) {
  // torchio.CropOrPad._get_six_bounds_parameters: the odd voxel of a difference goes to the START (ceil at the start,
  // floor at the end), for padding and for cropping alike.  offset = source index of target index 0.
  const int diff = target - source;
  *offset = diff >= 0 ? -((diff + 1) / 2) : ((-diff + 1) / 2);
}

int launch_crop_or_pad(const float* in, int H, int W, int D, float* out, int h, int w, int d, float pad,
                       cudaStream_t stream) {
  if (!in || !out) return fail(kInvalidArgument, "crop_or_pad: null pointer");
  if (H < 1 || W < 1 || D < 1 || h < 1 || w < 1 || d < 1) return fail(kInvalidArgument, "crop_or_pad: empty extent");
  int oy, ox, od;
  crop_or_pad_offsets(H, h, &oy);
  crop_or_pad_offsets(W, w, &ox);
  crop_or_pad_offsets(D, d, &od);
  const long long n = static_cast<long long>(h) * w * d;
  crop_or_pad_kernel<<<grid_for(n), 256, 0, stream>>>(in, H, W, D, out, h, w, d, oy, ox, od, pad);
  return check_launch("crop_or_pad_kernel");
}

size_t rescale_workspace_bytes(int64_t n) { return rescale_layout(n).total; }

int launch_rescale_intensity(float* vol, const float* mask, int64_t n, double p_lo, double p_hi, float out_min,
                             float out_max, void* workspace, size_t workspace_bytes, double* cutoffs,
                             cudaStream_t stream) {
  if (!vol || !mask || !workspace) return fail(kInvalidArgument, "rescale_intensity: null pointer");
  if (n < 1 || n > (1ll << 30)) return fail(kInvalidArgument, "rescale_intensity: bad voxel count");
  if (!(p_lo >= 0.0 && p_lo <= p_hi && p_hi <= 100.0)) return fail(kInvalidArgument, "rescale_intensity: bad percentiles");
  const RescaleLayout L = rescale_layout(n);
  if (workspace_bytes < L.total) return fail(kInvalidArgument, "rescale_intensity: workspace too small");
  char* ws = reinterpret_cast<char*>(workspace);
  float* keys_in = reinterpret_cast<float*>(ws + L.keys_in);
  float* keys_out = reinterpret_cast<float*>(ws + L.keys_out);
  RescaleScalars* sc = reinterpret_cast<RescaleScalars*>(ws + L.scalars);
  CDDPM_CUDA(cudaMemsetAsync(sc, 0, sizeof(RescaleScalars), stream));
  mask_keys_kernel<<<grid_for(n), 256, 0, stream>>>(vol, mask, n, keys_in, sc);
  CDDPM_TRY(check_launch("mask_keys_kernel"));
  size_t cb = L.cub_bytes;
  CDDPM_CUDA(cub::DeviceRadixSort::SortKeys(ws + L.cub, cb, keys_in, keys_out, static_cast<int>(n), 0, 32, stream));
  percentile_kernel<<<1, 1, 0, stream>>>(keys_out, p_lo, p_hi, sc);
  CDDPM_TRY(check_launch("percentile_kernel"));
  clip_minmax_kernel<<<grid_for(n), 256, 0, stream>>>(vol, n, sc);
  CDDPM_TRY(check_launch("clip_minmax_kernel"));
  rescale_kernel<<<grid_for(n), 256, 0, stream>>>(vol, n, out_min, out_max, sc);
  CDDPM_TRY(check_launch("rescale_kernel"));
  if (cutoffs)
    CDDPM_CUDA(cudaMemcpyAsync(cutoffs, sc->cutoff, 2 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  return kOk;
}

void resample_size(int source, double factor, int* target) {
  // torchio.Resample.get_reference_image: new_size = ceil(old_size * old_spacing / new_spacing), at least 1
  const int t = static_cast<int>(ceil(static_cast<double>(source) / factor));
  *target = t < 1 ? 1 : t;
}

size_t resample_workspace_bytes(int H, int W, int D) { return static_cast<size_t>(H) * W * D * sizeof(double) + 256; }

int launch_resample(const float* in, int H, int W, int D, double fy, double fx, double fz, int bspline, float* out,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!in || !out) return fail(kInvalidArgument, "resample: null pointer");
  if (H < 1 || W < 1 || D < 1 || !(fy > 0) || !(fx > 0) || !(fz > 0)) return fail(kInvalidArgument, "resample: bad geometry");
  int h, w, d;
  resample_size(H, fy, &h);
  resample_size(W, fx, &w);
  resample_size(D, fz, &d);
  const long long n_out = static_cast<long long>(h) * w * d;
  if (!bspline) {
    resample_kernel<false><<<grid_for(n_out), 256, 0, stream>>>(nullptr, in, H, W, D, out, h, w, d, fy, fx, fz);
    return check_launch("resample_kernel<nearest>");
  }
  if (!workspace || workspace_bytes < resample_workspace_bytes(H, W, D))
    return fail(kInvalidArgument, "resample: workspace too small");
  double* coef = reinterpret_cast<double*>(workspace);
  const long long n = static_cast<long long>(H) * W * D;
  to_double_kernel<<<grid_for(n), 256, 0, stream>>>(in, coef, n);
  CDDPM_TRY(check_launch("to_double_kernel"));
  // axis D (stride 1): H*W lines, line l starts at l * D
  if (D > 1) {
    bspline_prefilter_kernel<<<(H * W + 127) / 128, 128, 0, stream>>>(coef, H * W, D, 1, H * W, D, 0);
    CDDPM_TRY(check_launch("bspline_prefilter_kernel(D)"));
  }
  // axis W (stride D): H*D lines, (outer = y, inner = z)
  if (W > 1) {
    bspline_prefilter_kernel<<<(H * D + 127) / 128, 128, 0, stream>>>(coef, H * D, W, D, D, 1, static_cast<long long>(W) * D);
    CDDPM_TRY(check_launch("bspline_prefilter_kernel(W)"));
  }
  // axis H (stride W*D): W*D lines, all in the first plane
  if (H > 1) {
    bspline_prefilter_kernel<<<(W * D + 127) / 128, 128, 0, stream>>>(coef, W * D, H, static_cast<long long>(W) * D, W * D, 1, 0);
    CDDPM_TRY(check_launch("bspline_prefilter_kernel(H)"));
  }
  resample_kernel<true><<<grid_for(n_out), 256, 0, stream>>>(coef, in, H, W, D, out, h, w, d, fy, fx, fz);
  return check_launch("resample_kernel<bspline>");
}

}  // namespace cddpm
