// Spatial self-attention core of the UNet's AttentionBlock (QKVAttention, "new attention order"):
// src/models/modules/OpenAI_Unet.py:457-476.  qkv is the NHWC output of the 1x1 qkv convolution.
#pragma once
#include "common.h"

namespace cddpm {

// qkv [B, L, 3*C] 16-bit with q = channels [0,C), k = [C,2C), v = [2C,3C); head h owns channels [h*64,(h+1)*64) of
// each.  out[b, t, h*64+c] = sum_s softmax_s((q_t . k_s) / sqrt(64)) * v[s, c]   (fp32 softmax), 16-bit.
int launch_attention(const void* qkv, void* out, int B, int L, int C, int fmt, cudaStream_t stream);

}  // namespace cddpm

namespace cddpm {
// Backward of launch_attention (bf16 only): dqkv [B, L, 3*C] from dout [B, L, C] and the forward's qkv.
// scratch: attention_bwd_scratch_elems(B, L, C) 16-bit elements (probabilities and score gradients).
int64_t attention_bwd_scratch_elems(int B, int L, int C);
int launch_attention_bwd(const void* qkv, const void* dout, void* dqkv, void* scratch, int B, int L, int C, int fmt,
                         cudaStream_t stream);
// tcgen05 version (attention_bwd_tc.cu; the default - CDDPM_ATTN_BWD_TC=0 selects the mma.sync kernels): same arguments,
// same scratch layout.
bool attention_bwd_tc_enabled();
int launch_attention_bwd_tc(const void* qkv, const void* dout, void* dqkv, void* scratch, int B, int L, int C, int fmt,
                            cudaStream_t stream);
}  // namespace cddpm
