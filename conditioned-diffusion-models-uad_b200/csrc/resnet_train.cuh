// Condition encoder in TRAINING mode: ResNet-50 v1.5 (1 input channel, fc -> cond_dim) with batch-statistics
// BatchNorm, forward AND backward, hand-written.  Reference: DDPM_2D.training_step -> self(input) -> SparK_2D_encoder ->
// timm ResNet.forward in train() mode (src/models/DDPM_2D.py:101-122, spark/resnet.py:13-46, spark/models.py:89-109) and
// torch autograd over it.
//
// Every convolution is a GEMM on the tcgen05 kernels: forward and data gradient on the flat mode of conv_igemm (over an
// im2col'ed operand for the 3x3 / strided layers and for the 7x7 stem, whose 49 taps are padded to 64), weight gradient
// on flat_wgrad_tc_kernel (resnet_train.cu: the contraction runs over pixels, both operands MN-major straight from the
// row-major activations).  Operands are bf16 (BASELINE configs[4]), accumulation fp32; the raw convolution outputs, the
// BatchNorm statistics (fp32 per-block partial rows folded in float64), the normalisation and every gradient that is
// summed over several paths stay fp32.  Both launch lists are replayed as CUDA graphs keyed by the pointers they bind.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "conv_igemm.cuh"

namespace cddpm {

// dw[co][k] += sum_m dy[m][co] * x[m][k] for row-major bf16 matrices dy [M][Cout], x [M][K] (tcgen05, both operands
// MN-major); dw fp32 [Cout][K], accumulated with atomics (zero it first).
int launch_flat_wgrad_bf16(const void* dy, const void* x, int M, int Cout, int K, float* dw, cudaStream_t stream);

class ResNetTrainEngine {
 public:
  ~ResNetTrainEngine();
  int init(int image_h, int image_w, int cond_dim);
  // Entries in the module's state_dict order without num_batches_tracked (the list ResNetEngine uses): parameters and
  // BatchNorm running statistics.
  int entry_count() const { return static_cast<int>(entries_.size()); }
  int entry_info(int i, const char** name, int64_t* numel, int* is_param) const;
  int64_t grad_total() const { return grad_total_; }
  int grad_offset(int i, int64_t* off) const;  // -1 for running statistics
  // values[i]: the caller's fp32 device tensor of entry i.  Running statistics are UPDATED IN PLACE (momentum 0.1,
  // unbiased variance).  drop_scale: optional [blocks][B] per-sample scale of each bottleneck's residual branch
  // (timm DropPath: 0 or 1 / keep), or NULL.  x [B,1,H,W] fp32 -> out [B,cond_dim] fp32.
  int forward(const float* const* values, int count, const float* x, const float* drop_scale, float* out, int B,
              cudaStream_t stream);
  // Backward of the last forward: dout [B,cond_dim] fp32 -> every parameter gradient (fp32, reference layouts) at its
  // offset of `grads` (grad_total() floats, overwritten).
  int backward(const float* dout, float* grads, int B, cudaStream_t stream);
  int num_blocks() const { return static_cast<int>(blocks_.size()); }
  int launches_forward() const { return static_cast<int>(fwd_ops_.size()); }
  int launches_backward() const { return static_cast<int>(bwd_ops_.size()); }

 private:
  struct Entry {
    std::string name;
    int64_t numel = 0;
    int is_param = 0;
    int64_t goff = -1;
  };
  struct Unit {  // convolution + BatchNorm
    int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
    int e_w = -1, e_gamma = -1, e_beta = -1, e_mean = -1, e_var = -1;  // entry indices
    uint16_t *panel = nullptr, *panel_t = nullptr;                      // bf16 [cout][K], [K][cout]
    // planned per batch size
    int Hin = 0, Win = 0, Hout = 0, Wout = 0;
    const uint16_t* in = nullptr;  // input activation [M_in][cin]
    uint16_t* col = nullptr;       // [M_out][K] when k > 1 or stride > 1 (else `in` itself is the GEMM operand)
    float* y = nullptr;            // raw convolution output [M_out][cout]
    float *mean = nullptr, *rstd = nullptr;
    uint16_t* a = nullptr;         // output activation [M_out][cout]
    uint16_t* dy = nullptr;        // gradient of y, bf16 GEMM operand
    float* bsum = nullptr;         // [2][cout] sum dz, sum dz xhat
    float* dwp = nullptr;          // [cout][K] weight gradient in panel order
    float* dx = nullptr;           // data gradient: [M_out][K] (im2col space) or [M_in][cin] directly
    float* dxin = nullptr;         // [M_in][cin] after col2im (3x3 layers)
  };
  struct Block {
    int c1 = -1, c2 = -1, c3 = -1, down = -1;
    float* E = nullptr;  // relu-masked gradient of the block output [M_out][4w]
  };
  struct GradSrc {
    const float* p = nullptr;
    int mode = 0;  // 0 none, 1 same shape, 2 stride-2 subsampled source, 3 per-sample vector broadcast / HW
  };
  template <typename T>
  int dalloc(T** p, size_t n, std::vector<void*>* pool);
  int add_entry(const std::string& name, int64_t numel, int is_param);
  int add_unit(const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int pad);
  int plan(int B);
  void free_acts();
  int plan_unit_forward(Unit& u, const uint16_t* in, int Hin, int Win, int B);
  int plan_unit_backward(Unit& u, int B, bool need_dx);
  void push_stats(const Unit& u, int M);
  int launch_bn_backward(const Unit& u, const float* up, int relu, const float* scale, int rows_per_sample, int M,
                         cudaStream_t s);
  int push_gemm(std::vector<std::function<int(cudaStream_t)>>* ops, const void* a, int rows, int K, const void* panel,
                int N, void* out, bool out_f32);

  int H_ = 0, W_ = 0, cond_dim_ = 0;
  std::vector<Entry> entries_;
  std::vector<Unit> units_;
  std::vector<Block> blocks_;
  int e_fc_w_ = -1, e_fc_b_ = -1;
  int64_t grad_total_ = 0;
  std::vector<void*> owned_, act_owned_;
  int planned_B_ = 0;
  bool forward_done_ = false;
  // per-call bindings
  const float* const* cur_values_ = nullptr;
  const float* cur_x_ = nullptr;
  const float* cur_drop_ = nullptr;
  float* cur_out_ = nullptr;
  const float* cur_dout_ = nullptr;
  float* cur_grads_ = nullptr;
  const float** values_store_ = nullptr;  // host copy of the pointer table of the last forward
  std::vector<const float*> values_;
  std::vector<std::function<int(cudaStream_t)>> fwd_ops_, bwd_ops_;
  // The two launch lists as CUDA graphs, keyed by every pointer the captured launches bind (caller tensors and
  // parameters): a list runs eagerly the first time a key is seen, is captured the second time, replayed afterwards.
  struct GraphSlot {
    std::vector<const void*> key;
    cudaGraphExec_t exec = nullptr;
    int seen = 0;
    uint64_t last_use = 0;
  };
  // side: optional flags, one per op - flagged ops (the weight-gradient GEMMs: off the data-gradient chain) run on a
  // forked side stream / graph branch that is joined at the end of the list
  int run_list(std::vector<std::function<int(cudaStream_t)>>& ops, std::vector<GraphSlot>& slots,
               const std::vector<const void*>& key, cudaStream_t stream, const std::vector<uint8_t>* side = nullptr);
  int run_ops(std::vector<std::function<int(cudaStream_t)>>& ops, const std::vector<uint8_t>* side, cudaStream_t stream);
  std::vector<uint8_t> bwd_side_;
  cudaStream_t side_stream_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
  void drop_graphs();
  std::vector<GraphSlot> fwd_graphs_, bwd_graphs_;
  cudaStream_t cap_stream_ = nullptr;
  uint64_t use_clock_ = 0;
  int graph_captures_ = 0, graph_replays_ = 0;
  // stem / pool / head buffers
  uint16_t* pool_a_ = nullptr;
  float *pooled_ = nullptr, *dpooled_ = nullptr, *stem_g_ = nullptr;
  uint8_t* pool_arg_ = nullptr;  // argmax position of every max-pool window (forward -> backward)
  uint16_t* stem_col_ = nullptr;  // [B*Ho*Wo][64] im2col of the input (49 taps + padding), forward and weight gradient
  float* stem_dwp_ = nullptr;     // [64][64] stem weight gradient in panel order
  float* partial_ = nullptr;      // per-block partial column sums of the BatchNorm reductions (one reduction at a time)
};

}  // namespace cddpm
