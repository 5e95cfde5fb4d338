// Condition encoder: ResNet-50 v1.5, 1 input channel, fc -> cond_dim, eval mode (BatchNorm folded into the
// convolutions).  Reference: SparK_2D_encoder.forward (src/models/modules/spark/Spark_2D.py:285-290) -> timm
// ResNet.forward(pyramid=0) (spark/resnet.py:13-46), built by build_encoder (spark/models.py:89-109).
// Every convolution after the 7x7 stem runs on the tcgen05 GEMM (flat mode of conv_igemm) over an im2col'ed operand.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "conv_igemm.cuh"

namespace cddpm {

class ResNetEngine {
 public:
  ~ResNetEngine();
  int init(int image_h, int image_w, int cond_dim, int fmt);
  int param_count() const { return static_cast<int>(params_.size()); }
  int param_info(int i, const char** name, int64_t* numel) const;
  int set_param(const char* name, const float* dev_ptr, int64_t numel, cudaStream_t stream);
  // x [B,1,H,W] fp32 -> c [B,cond_dim] fp32
  int forward(const float* x, float* c, int B, cudaStream_t stream);

 private:
  struct Param {
    std::string name;
    int64_t numel = 0;
    bool set = false;
    float* dst = nullptr;
  };
  struct ConvBN {
    std::string conv, bn;
    int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
    float *w = nullptr, *gamma = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr;
    void* wpacked = nullptr;  // [cout][k*k*cin] 16-bit, BN scale folded in
    float* bias = nullptr;    // folded BN shift
  };
  struct Block {
    int c1 = -1, c2 = -1, c3 = -1, down = -1;
  };
  template <typename T>
  int dalloc(T** p, size_t n, std::vector<void*>* pool);
  int add_param(const std::string& name, int64_t numel, float** dst);
  int add_convbn(const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int pad);
  int fold(cudaStream_t stream);
  int plan(int B);
  void free_acts();

  int H_ = 0, W_ = 0, cond_dim_ = 0, fmt_ = 0;
  std::vector<Param> params_;
  std::map<std::string, int> index_;
  std::vector<ConvBN> convs_;
  std::vector<Block> blocks_;
  float *fc_w_ = nullptr, *fc_b_ = nullptr, *stem_w_ = nullptr, *stem_b_ = nullptr;
  std::vector<void*> owned_, act_owned_;
  bool dirty_ = true;
  int planned_B_ = 0;
  const float* cur_x_ = nullptr;
  float* cur_out_ = nullptr;
  std::vector<std::function<int(cudaStream_t)>> ops_;
};

}  // namespace cddpm
