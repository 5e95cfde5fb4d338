// The conditioned denoising UNet forward as a pre-planned sequence of sm_100a kernels.
//
// Mirrors the module tree of the reference's UNetModel (src/models/modules/OpenAI_Unet.py:483-1006) and speaks its
// state_dict vocabulary: parameters are pushed by their reference key ("input_blocks.4.0.in_layers.2.weight", ...)
// as fp32 device arrays and re-laid-out here (16-bit K-major conv panels, concatenated FiLM projection, ...).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cddpm_b200.h"
#include "common.h"
#include "conv_igemm.cuh"
#include "elementwise.cuh"

namespace cddpm {

struct ActTensor {
  void* p = nullptr;
  int C = 0, H = 0, W = 0;
  double* stats = nullptr;  // [B][C/4][2] GroupNorm (sum, sumsq) buckets emitted by the producer, or nullptr
  size_t elems(int B) const { return static_cast<size_t>(B) * H * W * C; }
};

class UNetEngine {
 public:
  ~UNetEngine();
  int init(const cddpm_unet_config& cfg);
  int param_count() const { return static_cast<int>(params_.size()); }
  int param_info(int i, const char** name, int64_t* numel) const;
  int set_param(const char* name, const float* dev_ptr, int64_t numel, cudaStream_t stream);
  // All parameters at once (values[i] for parameter i, no NULLs): the ~500 small re-layout launches are captured into
  // a CUDA graph keyed by the pointer list (optimizers update in place, so the pointers are stable across steps) and
  // replayed with one launch per step.
  int set_params_all(const float* const* values, int count, cudaStream_t stream);
  int forward(const float* x, const int64_t* t, const float* cond, float* out, int B, cudaStream_t stream);
  int tap(const char* layer, void** ptr, int* C, int* H, int* W) const;
  int film(const float** ptr, int* stride) const;
  // Arm per-launch CUDA-event timing of the convolution kernels for the NEXT forward; read it back afterwards.
  int profile_arm();
  int profile_read(double* conv_ms, int* conv_launches);
  int64_t conv_flops_per_sample() const { return conv_flops_; }
  int launches_per_forward() const { return static_cast<int>(ops_.size()); }
  int fmt() const { return cfg_.fmt; }
  // Backward of the LAST forward (training step): dout = dL/d out [B,1,H,W] fp32; every parameter gradient is written
  // into the flat fp32 buffer `grads` (grad_total() floats, parameter i at grad_offset(i), reference layouts);
  // dcond (optional) receives dL/d cond [B, num_classes].
  int backward(const float* dout, float* grads, float* dcond, int B, cudaStream_t stream);
  // Training engines keep every intermediate the backward pass reads; inference engines (the default) may plan
  // inference-only fusions (the out_layers GroupNorm finished inside the producing convolution: its raw result is
  // never stored).  Switching drops the current plan.
  int set_training(bool on);
  bool training() const { return training_; }
  int64_t grad_total() const { return grad_total_; }
  int grad_offset(int i, int64_t* off) const;
  int64_t bwd_flops_per_sample() const { return bwd_flops_; }
  int bwd_launches() const { return static_cast<int>(bwd_ops_.size()); }

 private:
  struct Param {
    std::string name;
    int64_t numel = 0;
    bool set = false;
    std::function<int(const float*, cudaStream_t)> load;
    int64_t goff = 0;  // offset of this parameter's gradient in the flat gradient buffer
  };
  struct ResLayer {
    std::string prefix;
    int cin = 0, cout = 0, mode = 0, film_off = 0;
    bool has_skip = false;
    float *gn1_w = nullptr, *gn1_b = nullptr, *gn2_w = nullptr, *gn2_b = nullptr;
    void *w1 = nullptr, *w2 = nullptr;
    void *w1t = nullptr, *w2t = nullptr, *wskipt = nullptr;  // data-gradient panels (transposed, taps flipped)
    float *b1 = nullptr, *b2 = nullptr, *bskip = nullptr, *b2sum = nullptr;
    int in_c0 = 0, in_c1 = 0;  // channel split of the (possibly concatenated) block input
  };
  struct AttnLayer {
    std::string prefix;
    int ch = 0;
    float *gn_w = nullptr, *gn_b = nullptr, *bqkv = nullptr, *bproj = nullptr;
    void *wqkv = nullptr, *wproj = nullptr;
    void *wqkvt = nullptr, *wprojt = nullptr;
  };
  struct Layer {
    int kind = 0;  // 0 stem conv, 1 res, 2 attn
    int idx = 0;
  };

  // what the backward pass needs to know about one planned layer (forward order)
  struct ResPlan {
    int layer = 0;
    bool has_a1 = false;
    ActTensor a0, a1, tA, tH, tB, tS, out;
  };
  struct AttnPlan {
    int layer = 0;
    ActTensor x, tN, tQ, tA, out;
  };
  struct Step {
    int kind = 0;  // 1 res, 2 attn
    int idx = 0;   // index into res_plans_ / attn_plans_
  };
  std::vector<ResPlan> res_plans_;
  std::vector<AttnPlan> attn_plans_;
  std::vector<Step> steps_;
  ActTensor stem_out_, head_in_, head_tN_;
  int plan_backward(int B);
  float* grad_of(const std::string& name) const;
  std::vector<std::function<int(cudaStream_t)>> bwd_ops_;
  bool bwd_planned_ = false;
  // the backward launch list as CUDA graphs keyed by every caller pointer the launches bind (eager the first time a
  // key is seen, captured the second time, replayed afterwards; at most four, least recently used replaced)
  struct BwdGraph {
    std::vector<const void*> key;
    cudaGraphExec_t exec = nullptr;
    int seen = 0;
    uint64_t last_use = 0;
  };
  std::vector<BwdGraph> bwd_graphs_;
  uint64_t bwd_clock_ = 0;
  int bwd_captures_ = 0, bwd_replays_ = 0;
  float* cur_grads_ = nullptr;
  const float* cur_dout_ = nullptr;
  float* cur_dcond_ = nullptr;
  int64_t grad_total_ = 0;
  int64_t bwd_flops_ = 0;

  template <typename T>
  int dalloc(T** p, size_t n);
  int add_param(const std::string& name, int64_t numel, std::function<int(const float*, cudaStream_t)> load);
  int add_copy_param(const std::string& name, int64_t numel, float** dst);
  int build_layers();
  int add_res(const std::string& prefix, int c0, int c1, int cout, int mode);
  int add_attn(const std::string& prefix, int ch);
  int plan(int B);
  int plan_res(const ResLayer& L, const ActTensor& a0, const ActTensor* a1, ActTensor* out, int B);
  int plan_attn(const AttnLayer& L, const ActTensor& x, ActTensor* out, int B);
  int act_alloc(ActTensor* t, int C, int H, int W, int B, bool with_stats = false);
  void free_acts();
  void push_conv(const ConvDesc& d, int* status);
  void push_gn(GnApplyArgs g);
  bool fused_stats_ = true;
  bool training_ = false;
  bool plan_fused_ = false;  // the current plan contains an inference-only fusion (backward refuses it)
  bool plan_fused_head_ = false;  // ... the head's GroupNorm inside conv_out (no normalised tensor kept)

  cddpm_unet_config cfg_{};
  int emb_dim_ = 0, half_dim_ = 0, film_total_ = 0;
  std::vector<Param> params_;
  std::map<std::string, int> param_index_;
  std::vector<void*> owned_;       // parameter-side allocations
  std::vector<void*> act_owned_;   // activation-side allocations (re-planned when B changes)
  std::vector<ResLayer> res_;
  std::vector<AttnLayer> attn_;
  std::vector<std::vector<Layer>> in_blocks_, out_blocks_;
  std::vector<Layer> mid_;
  std::vector<int> in_block_ch_;
  // embedding parameters
  float *te0_w = nullptr, *te0_b = nullptr, *te2_w = nullptr, *te2_b = nullptr;
  float *le0_w = nullptr, *le0_b = nullptr, *le2_w = nullptr, *le2_b = nullptr;
  uint16_t *te0_w16 = nullptr, *te2_w16 = nullptr, *le0_w16 = nullptr, *le2_w16 = nullptr;
  uint16_t* film_w16 = nullptr;  // [film_total][emb_dim] 16-bit K-major
  uint16_t* film_w16t = nullptr; // [emb_dim][film_total]: the panel of the FiLM projection's data gradient
  uint16_t *sin16_ = nullptr, *hid16_ = nullptr, *cond16_ = nullptr, *hidc16_ = nullptr;  // 16-bit MLP hand-offs
  float* film_b = nullptr;
  uint16_t* emb_act16_ = nullptr;
  float *stem_w = nullptr, *stem_b = nullptr, *head_gn_w = nullptr, *head_gn_b = nullptr, *head_w = nullptr,
        *head_b = nullptr;
  // per-forward bindings
  const float* cur_x_ = nullptr;
  const int64_t* cur_t_ = nullptr;
  const float* cur_cond_ = nullptr;
  float* cur_out_ = nullptr;
  // plan
  int planned_B_ = 0;
  std::vector<std::function<int(cudaStream_t)>> ops_;
  std::map<std::string, ActTensor> taps_;
  float *sinus_ = nullptr, *hid_t_ = nullptr, *hid_c_ = nullptr, *emb_act_ = nullptr, *film_out_ = nullptr,
        *gn_partial_ = nullptr;
  double* stats_arena_ = nullptr;
  size_t stats_cap_ = 0, stats_used_ = 0;
  int64_t conv_flops_ = 0;
  bool profile_armed_ = false;
  std::vector<cudaEvent_t> profile_events_;
  // CUDA-graph replay of the planned launch sequence: captured on the second forward of a plan (the first runs
  // directly and finishes every one-time setup), replayed afterwards.  Inputs and the output go through engine-owned
  // staging buffers because the graph bakes pointers in.  CDDPM_GRAPH=0 disables it.
  void drop_graph();
  cudaStream_t cap_stream_ = nullptr;
  // forked capture of the embedding path (capture_ops): op index ranges of the timestep chain, the condition chain,
  // the FiLM projection and the first reader of the FiLM table
  int capture_ops();
  int run_ops_inline(cudaStream_t stream);
  int memset_op_ = -1;  // index of the statistics-arena memset in ops_ (not a kernel: no programmatic launch behind it)
  int emb_t_begin_ = 0, emb_t_end_ = 0, emb_c_begin_ = 0, emb_c_end_ = 0, film_op_ = -1, first_film_use_ = -1;
  cudaStream_t side_stream_[2] = {nullptr, nullptr};
  cudaEvent_t fork_ev_[3] = {nullptr, nullptr, nullptr};
  cudaGraphExec_t graph_exec_ = nullptr;
  ParamPushTable* push_table_ = nullptr;  // recorded re-layout jobs of a whole-model push (param_push.cu)
  std::vector<const float*> push_key_;
  int forwards_on_plan_ = 0;
  float *stage_x_ = nullptr, *stage_cond_ = nullptr, *stage_out_ = nullptr;
  int64_t* stage_t_ = nullptr;
};

}  // namespace cddpm
