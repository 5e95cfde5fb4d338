#include "resnet_engine.cuh"

#include <cuda_fp16.h>

#include "elementwise.cuh"

namespace cddpm {

namespace {

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float from16(uint16_t u, int fmt) {
  if (fmt == 1) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}

// Fold eval-mode BatchNorm (y = (conv - mean) * gamma / sqrt(var + eps) + beta) into the packed 16-bit weight matrix
// [cout][tap][cin] and an fp32 bias.
__global__ void fold_bn_pack_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, int cout, int cin, int taps, uint16_t* __restrict__ wp,
                                    float* __restrict__ bias, int fmt) {
  const size_t total = static_cast<size_t>(cout) * taps * cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    const int tap = static_cast<int>((i / cin) % taps);
    const int co = static_cast<int>(i / (static_cast<size_t>(cin) * taps));
    const float s = gamma[co] / sqrtf(var[co] + 1e-5f);
    wp[i] = to16(w[(static_cast<size_t>(co) * cin + ci) * taps + tap] * s, fmt);
    if (ci == 0 && tap == 0) bias[co] = beta[co] - mean[co] * s;
  }
}

// Stem: 7x7 stride 2 pad 3 over one input channel, folded BN, ReLU -> NHWC 16-bit.  ws: [49][64] folded fp32 weights.
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   const float* __restrict__ mean, const float* __restrict__ var,
                                                   uint16_t* __restrict__ out, int B, int H, int W, int fmt) {
  __shared__ float sw[49 * 64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) {
    const int tap = i / 64, co = i % 64;
    sw[i] = w[co * 49 + tap] * (gamma[co] / sqrtf(var[co] + 1e-5f));
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = beta[i] - mean[i] * (gamma[i] / sqrtf(var[i] + 1e-5f));
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const int g = threadIdx.x & 7;  // 8 channel groups of 8
  const size_t total = static_cast<size_t>(B) * Ho * Wo;
  for (size_t pix = static_cast<size_t>(blockIdx.x) * 32 + (threadIdx.x >> 3); pix < total;
       pix += static_cast<size_t>(gridDim.x) * 32) {
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>((pix / Wo) % Ho);
    const size_t n = pix / (static_cast<size_t>(Wo) * Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sb[g * 8 + j];
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = oy * 2 + ky - 3;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 7; ++kx) {
        const int ix = ox * 2 + kx - 3;
        if (ix < 0 || ix >= W) continue;
        const float v = __ldg(x + (n * H + iy) * W + ix);
        const float* wr = &sw[(ky * 7 + kx) * 64 + g * 8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
      }
    }
    uint16_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = to16(fmaxf(acc[j], 0.f), fmt);
    *reinterpret_cast<uint4*>(out + pix * 64 + g * 8) = *reinterpret_cast<uint4*>(o);
  }
}

__global__ void maxpool3s2_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int B, int H, int W, int C,
                                  int fmt) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int ox = static_cast<int>((i / C) % Wo), oy = static_cast<int>((i / (static_cast<size_t>(C) * Wo)) % Ho);
    const size_t n = i / (static_cast<size_t>(C) * Wo * Ho);
    float m = -INFINITY;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = oy * 2 + ky - 1, ix = ox * 2 + kx - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        m = fmaxf(m, from16(in[((n * H + iy) * W + ix) * C + c], fmt));
      }
    out[i] = to16(m, fmt);
  }
}

// col[(n,oy,ox)][tap][c] = in[n][oy*stride+ky-pad][ox*stride+kx-pad][c] (zero outside); 8 channels per thread.
__global__ void im2col_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ col, int B, int H, int W, int C,
                              int k, int stride, int pad, int Ho, int Wo) {
  const int cv = C >> 3;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * k * k * cv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    const int tap = static_cast<int>((i / cv) % (k * k));
    const size_t pix = i / (static_cast<size_t>(cv) * k * k);
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>((pix / Wo) % Ho);
    const size_t n = pix / (static_cast<size_t>(Wo) * Ho);
    const int iy = oy * stride + tap / k - pad, ix = ox * stride + tap % k - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      val = __ldg(reinterpret_cast<const uint4*>(in + ((n * H + iy) * W + ix) * C + v * 8));
    *reinterpret_cast<uint4*>(col + (pix * k * k + tap) * C + v * 8) = val;
  }
}

__global__ void avgpool_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int B, int HW, int C, int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, n = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += from16(in[(static_cast<size_t>(n) * HW + p) * C + c], fmt);
  out[i] = s / static_cast<float>(HW);
}

}  // namespace

ResNetEngine::~ResNetEngine() {
  free_acts();
  for (void* p : owned_) cudaFree(p);
}
void ResNetEngine::free_acts() {
  for (void* p : act_owned_) cudaFree(p);
  act_owned_.clear();
  ops_.clear();
  planned_B_ = 0;
}

template <typename T>
int ResNetEngine::dalloc(T** p, size_t n, std::vector<void*>* pool) {
  void* q = nullptr;
  CDDPM_CUDA(cudaMalloc(&q, n * sizeof(T) + 256));
  pool->push_back(q);
  *p = reinterpret_cast<T*>(q);
  return kOk;
}

int ResNetEngine::add_param(const std::string& name, int64_t numel, float** dst) {
  CDDPM_TRY(dalloc(dst, static_cast<size_t>(numel), &owned_));
  Param p;
  p.name = name;
  p.numel = numel;
  p.dst = *dst;
  index_[name] = static_cast<int>(params_.size());
  params_.push_back(p);
  return kOk;
}

int ResNetEngine::add_convbn(const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride,
                             int pad) {
  ConvBN c;
  c.conv = conv;
  c.bn = bn;
  c.cin = cin;
  c.cout = cout;
  c.k = k;
  c.stride = stride;
  c.pad = pad;
  CDDPM_TRY(add_param(conv + ".weight", static_cast<int64_t>(cout) * cin * k * k, &c.w));
  CDDPM_TRY(add_param(bn + ".weight", cout, &c.gamma));
  CDDPM_TRY(add_param(bn + ".bias", cout, &c.beta));
  CDDPM_TRY(add_param(bn + ".running_mean", cout, &c.mean));
  CDDPM_TRY(add_param(bn + ".running_var", cout, &c.var));
  uint16_t* wp = nullptr;
  CDDPM_TRY(dalloc(&wp, static_cast<size_t>(cout) * cin * k * k, &owned_));
  c.wpacked = wp;
  CDDPM_TRY(dalloc(&c.bias, static_cast<size_t>(cout), &owned_));
  convs_.push_back(c);
  return kOk;
}

int ResNetEngine::init(int image_h, int image_w, int cond_dim, int fmt) {
  H_ = image_h;
  W_ = image_w;
  cond_dim_ = cond_dim;
  fmt_ = fmt;
  if (image_h % 32 != 0 || image_w % 32 != 0) return fail(kUnsupported, "encoder: image size must be a multiple of 32");
  CDDPM_TRY(add_convbn("conv1", "bn1", 1, 64, 7, 2, 3));  // index 0: stem (direct kernel, not packed)
  const int layers[4] = {3, 4, 6, 3};
  const int widths[4] = {64, 128, 256, 512};
  int cin = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < layers[li]; ++bi) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      const int w = widths[li];
      const int stride = (bi == 0 && li > 0) ? 2 : 1;
      Block b;
      CDDPM_TRY(add_convbn(p + ".conv1", p + ".bn1", cin, w, 1, 1, 0));
      b.c1 = static_cast<int>(convs_.size()) - 1;
      CDDPM_TRY(add_convbn(p + ".conv2", p + ".bn2", w, w, 3, stride, 1));
      b.c2 = static_cast<int>(convs_.size()) - 1;
      CDDPM_TRY(add_convbn(p + ".conv3", p + ".bn3", w, 4 * w, 1, 1, 0));
      b.c3 = static_cast<int>(convs_.size()) - 1;
      if (bi == 0) {
        CDDPM_TRY(add_convbn(p + ".downsample.0", p + ".downsample.1", cin, 4 * w, 1, stride, 0));
        b.down = static_cast<int>(convs_.size()) - 1;
      }
      blocks_.push_back(b);
      cin = 4 * w;
    }
  }
  CDDPM_TRY(add_param("fc.weight", static_cast<int64_t>(cond_dim) * 2048, &fc_w_));
  CDDPM_TRY(add_param("fc.bias", cond_dim, &fc_b_));
  return kOk;
}

int ResNetEngine::param_info(int i, const char** name, int64_t* numel) const {
  if (i < 0 || i >= param_count()) return fail(kInvalidArgument, "param index out of range");
  *name = params_[i].name.c_str();
  *numel = params_[i].numel;
  return kOk;
}

int ResNetEngine::set_param(const char* name, const float* dev_ptr, int64_t numel, cudaStream_t stream) {
  auto it = index_.find(name);
  if (it == index_.end()) return fail(kInvalidArgument, std::string("unknown encoder parameter: ") + name);
  Param& p = params_[it->second];
  if (p.numel != numel) return fail(kInvalidArgument, std::string("size mismatch for ") + name);
  if (!dev_ptr) return fail(kInvalidArgument, "null parameter pointer");
  CDDPM_CUDA(cudaMemcpyAsync(p.dst, dev_ptr, numel * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  p.set = true;
  dirty_ = true;
  return kOk;
}

int ResNetEngine::fold(cudaStream_t stream) {
  for (size_t i = 1; i < convs_.size(); ++i) {
    const ConvBN& c = convs_[i];
    const size_t total = static_cast<size_t>(c.cout) * c.cin * c.k * c.k;
    int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
    fold_bn_pack_kernel<<<blocks, 256, 0, stream>>>(c.w, c.gamma, c.beta, c.mean, c.var, c.cout, c.cin, c.k * c.k,
                                                    reinterpret_cast<uint16_t*>(c.wpacked), c.bias, fmt_);
    CDDPM_TRY(check_launch("fold_bn_pack_kernel"));
  }
  dirty_ = false;
  return kOk;
}

int ResNetEngine::plan(int B) {
  free_acts();
  const int fmt = fmt_;
  auto alloc16 = [&](uint16_t** p, size_t n) { return dalloc(p, n, &act_owned_); };
  int H = H_ / 2, W = W_ / 2;
  uint16_t* stem = nullptr;
  CDDPM_TRY(alloc16(&stem, static_cast<size_t>(B) * H * W * 64));
  {
    const ConvBN c = convs_[0];
    const int Hin = H_, Win = W_;
    ops_.push_back([=](cudaStream_t s) {
      const size_t total = static_cast<size_t>(B) * (Hin / 2) * (Win / 2);
      int blocks = static_cast<int>(std::min<size_t>((total + 31) / 32, 148 * 8));
      stem_kernel<<<blocks, 256, 0, s>>>(cur_x_, c.w, c.gamma, c.beta, c.mean, c.var, stem, B, Hin, Win, fmt);
      return check_launch("stem_kernel");
    });
  }
  const int Hp = (H + 2 - 3) / 2 + 1, Wp = (W + 2 - 3) / 2 + 1;
  uint16_t* x = nullptr;
  CDDPM_TRY(alloc16(&x, static_cast<size_t>(B) * Hp * Wp * 64));
  {
    const int h = H, w = W;
    uint16_t* dst = x;
    ops_.push_back([=](cudaStream_t s) {
      const size_t total = static_cast<size_t>(B) * Hp * Wp * 64;
      maxpool3s2_kernel<<<static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16)), 256, 0, s>>>(stem, dst, B, h, w,
                                                                                                    64, fmt);
      return check_launch("maxpool3s2_kernel");
    });
  }
  H = Hp;
  W = Wp;
  int C = 64;

  auto gemm = [&](const uint16_t* a, int rows, int K, const ConvBN& c, const uint16_t* residual, int relu,
                  uint16_t* out) -> int {
    ConvDesc d;
    d.num_src = 1;
    d.src[0] = a;
    d.src_c[0] = K;
    d.src_taps[0] = 1;
    d.flat_rows = rows;
    d.Cout = c.cout;
    d.wpacked = c.wpacked;
    d.bias = c.bias;
    d.residual = residual;
    d.out = out;
    d.ab_format = fmt;
    d.relu = relu;
    auto p = std::make_shared<ConvIgemmParams>();
    CDDPM_TRY(build_conv_params(d, p.get()));
    ops_.push_back([p](cudaStream_t s) { return launch_conv_igemm(*p, s); });
    return kOk;
  };
  auto im2col = [&](const uint16_t* in, int h, int w, int c, int k, int stride, int pad, int ho, int wo,
                    uint16_t* col) {
    ops_.push_back([=](cudaStream_t s) {
      const size_t total = static_cast<size_t>(B) * ho * wo * k * k * (c / 8);
      im2col_kernel<<<static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16)), 256, 0, s>>>(in, col, B, h, w, c, k,
                                                                                                stride, pad, ho, wo);
      return check_launch("im2col_kernel");
    });
  };

  for (const Block& b : blocks_) {
    const ConvBN& c1 = convs_[b.c1];
    const ConvBN& c2 = convs_[b.c2];
    const ConvBN& c3 = convs_[b.c3];
    const int stride = c2.stride;
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    const int rows_in = B * H * W, rows_out = B * Ho * Wo;
    uint16_t *t1 = nullptr, *col = nullptr, *t2 = nullptr, *idt = nullptr, *out = nullptr;
    CDDPM_TRY(alloc16(&t1, static_cast<size_t>(rows_in) * c1.cout));
    CDDPM_TRY(gemm(x, rows_in, C, c1, nullptr, 1, t1));
    CDDPM_TRY(alloc16(&col, static_cast<size_t>(rows_out) * 9 * c2.cin));
    im2col(t1, H, W, c2.cin, 3, stride, 1, Ho, Wo, col);
    CDDPM_TRY(alloc16(&t2, static_cast<size_t>(rows_out) * c2.cout));
    CDDPM_TRY(gemm(col, rows_out, 9 * c2.cin, c2, nullptr, 1, t2));
    const uint16_t* identity = x;
    if (b.down >= 0) {
      const ConvBN& cd = convs_[b.down];
      const uint16_t* src = x;
      if (stride != 1) {
        uint16_t* sub = nullptr;
        CDDPM_TRY(alloc16(&sub, static_cast<size_t>(rows_out) * C));
        im2col(x, H, W, C, 1, stride, 0, Ho, Wo, sub);
        src = sub;
      }
      CDDPM_TRY(alloc16(&idt, static_cast<size_t>(rows_out) * cd.cout));
      CDDPM_TRY(gemm(src, rows_out, C, cd, nullptr, 0, idt));
      identity = idt;
    }
    CDDPM_TRY(alloc16(&out, static_cast<size_t>(rows_out) * c3.cout));
    CDDPM_TRY(gemm(t2, rows_out, c3.cin, c3, identity, 1, out));
    x = out;
    H = Ho;
    W = Wo;
    C = c3.cout;
  }
  float* pooled = nullptr;
  CDDPM_TRY(dalloc(&pooled, static_cast<size_t>(B) * C, &act_owned_));
  {
    const uint16_t* src = x;
    const int hw = H * W, cc = C;
    ops_.push_back([=](cudaStream_t s) {
      avgpool_kernel<<<(B * cc + 255) / 256, 256, 0, s>>>(src, pooled, B, hw, cc, fmt);
      return check_launch("avgpool_kernel");
    });
    const int cd = cond_dim_;
    ops_.push_back([=](cudaStream_t s) { return launch_linear(pooled, cc, fc_w_, fc_b_, cur_out_, cd, B, cc, cd, 0, s); });
  }
  planned_B_ = B;
  return kOk;
}

int ResNetEngine::forward(const float* x, float* c, int B, cudaStream_t stream) {
  if (!x || !c) return fail(kInvalidArgument, "encoder_forward: null pointer");
  if (B < 1) return fail(kInvalidArgument, "encoder_forward: empty batch");
  for (const Param& p : params_)
    if (!p.set) return fail(kNotReady, "encoder parameter not set: " + p.name);
  if (dirty_) CDDPM_TRY(fold(stream));
  if (B != planned_B_) {
    CDDPM_CUDA(cudaDeviceSynchronize());
    int st = plan(B);
    if (st != kOk) {
      free_acts();
      return st;
    }
  }
  cur_x_ = x;
  cur_out_ = c;
  for (auto& op : ops_) CDDPM_TRY(op(stream));
  return kOk;
}

}  // namespace cddpm
