// OpenSimplex-2D fractal noise on the GPU (the reference's host-side numba producer, src/utils/generate_noise.py:
// rand_2d_octaves :97-114, _noise2 :252-349, _extrapolate2 :235-239, _noise2a :352-358).
//
// This translation unit is compiled with -fmad=false (see build.py: *_nofma.cu): the reference evaluates every
// product and sum in IEEE float64 without fusion, and we reproduce its operation order so that the float16 field is
// bit-identical.  The 256-entry permutation (generate_noise.py:214-232, a 64-bit LCG) is computed on the host and
// passed by value.  One thread per pixel; 6 octaves x ~60 flops — the point is removing a 5 ms host round trip from
// every reverse-diffusion step, not throughput.
#include <cuda_fp16.h>

#include "simplex.cuh"

namespace cddpm {

namespace {

struct PermTable {
  unsigned char p[256];
};

__device__ __forceinline__ double grad_dot(const PermTable& pt, long long xsb, long long ysb, double dx, double dy) {
  const int g[16] = {5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5};
  const int idx = pt.p[(pt.p[xsb & 0xFF] + ysb) & 0xFF] & 0x0E;
  return static_cast<double>(g[idx]) * dx + static_cast<double>(g[idx + 1]) * dy;
}

__device__ double noise2(const PermTable& pt, double x, double y) {
  const double kStretch = -0.211324865405187;
  const double kSquish = 0.366025403784439;
  const double stretch = (x + y) * kStretch;
  const double xs = x + stretch;
  const double ys = y + stretch;
  long long xsb = static_cast<long long>(floor(xs));
  long long ysb = static_cast<long long>(floor(ys));
  const double squish = static_cast<double>(xsb + ysb) * kSquish;
  const double xb = static_cast<double>(xsb) + squish;
  const double yb = static_cast<double>(ysb) + squish;
  const double xins = xs - static_cast<double>(xsb);
  const double yins = ys - static_cast<double>(ysb);
  const double in_sum = xins + yins;
  double dx0 = x - xb;
  double dy0 = y - yb;
  double value = 0.0;

  const double dx1 = dx0 - 1 - kSquish;
  const double dy1 = dy0 - 0 - kSquish;
  double attn1 = 2 - dx1 * dx1 - dy1 * dy1;
  if (attn1 > 0) {
    attn1 *= attn1;
    value += attn1 * attn1 * grad_dot(pt, xsb + 1, ysb + 0, dx1, dy1);
  }
  const double dx2 = dx0 - 0 - kSquish;
  const double dy2 = dy0 - 1 - kSquish;
  double attn2 = 2 - dx2 * dx2 - dy2 * dy2;
  if (attn2 > 0) {
    attn2 *= attn2;
    value += attn2 * attn2 * grad_dot(pt, xsb + 0, ysb + 1, dx2, dy2);
  }

  long long xsv_ext, ysv_ext;
  double dx_ext, dy_ext;
  if (in_sum <= 1) {
    const double zins = 1 - in_sum;
    if (zins > xins || zins > yins) {
      if (xins > yins) {
        xsv_ext = xsb + 1;
        ysv_ext = ysb - 1;
        dx_ext = dx0 - 1;
        dy_ext = dy0 + 1;
      } else {
        xsv_ext = xsb - 1;
        ysv_ext = ysb + 1;
        dx_ext = dx0 + 1;
        dy_ext = dy0 - 1;
      }
    } else {
      xsv_ext = xsb + 1;
      ysv_ext = ysb + 1;
      dx_ext = dx0 - 1 - 2 * kSquish;
      dy_ext = dy0 - 1 - 2 * kSquish;
    }
  } else {
    const double zins = 2 - in_sum;
    if (zins < xins || zins < yins) {
      if (xins > yins) {
        xsv_ext = xsb + 2;
        ysv_ext = ysb + 0;
        dx_ext = dx0 - 2 - 2 * kSquish;
        dy_ext = dy0 + 0 - 2 * kSquish;
      } else {
        xsv_ext = xsb + 0;
        ysv_ext = ysb + 2;
        dx_ext = dx0 + 0 - 2 * kSquish;
        dy_ext = dy0 - 2 - 2 * kSquish;
      }
    } else {
      dx_ext = dx0;
      dy_ext = dy0;
      xsv_ext = xsb;
      ysv_ext = ysb;
    }
    xsb += 1;
    ysb += 1;
    dx0 = dx0 - 1 - 2 * kSquish;
    dy0 = dy0 - 1 - 2 * kSquish;
  }
  double attn0 = 2 - dx0 * dx0 - dy0 * dy0;
  if (attn0 > 0) {
    attn0 *= attn0;
    value += attn0 * attn0 * grad_dot(pt, xsb, ysb, dx0, dy0);
  }
  double attn_ext = 2 - dx_ext * dx_ext - dy_ext * dy_ext;
  if (attn_ext > 0) {
    attn_ext *= attn_ext;
    value += attn_ext * attn_ext * grad_dot(pt, xsv_ext, ysv_ext, dx_ext, dy_ext);
  }
  return value / 47;
}

__global__ void simplex_fractal_kernel(const PermTable pt, __half* __restrict__ out, float* __restrict__ out_f32,
                                       int H, int W, int B, int octaves, double persistence, double frequency) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int i = idx / W;  // row  -> y[i]
  const int j = idx - i * W;  // col -> x[j]
  double acc = 0.0;
  double amp = 1.0;
  double f = frequency;
  for (int o = 0; o < octaves; ++o) {
    acc += amp * noise2(pt, static_cast<double>(j) / f, static_cast<double>(i) / f);
    f /= 2;
    amp *= persistence;
  }
  // torch: float64 tensor .half() converts through float32
  const float v32 = static_cast<float>(acc);
  if (out != nullptr) {
    const __half h = __float2half_rn(v32);
    for (int b = 0; b < B; ++b) out[static_cast<size_t>(b) * H * W + idx] = h;
  }
  if (out_f32 != nullptr) out_f32[idx] = v32;
}

}  // namespace

int launch_simplex_noise(const unsigned char* perm_host, void* out_f16, float* out_f32, int B, int H, int W,
                         int octaves, double persistence, double frequency, cudaStream_t stream) {
  if (!perm_host) return fail(kInvalidArgument, "simplex: null permutation");
  if (H != W) return fail(kUnsupported, "simplex: the reference's indexing is only defined for square images");
  PermTable pt;
  for (int k = 0; k < 256; ++k) pt.p[k] = perm_host[k];
  const int n = H * W;
  simplex_fractal_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pt, reinterpret_cast<__half*>(out_f16), out_f32, H, W, B,
                                                              octaves, persistence, frequency);
  return check_launch("simplex_fractal_kernel");
}

}  // namespace cddpm
