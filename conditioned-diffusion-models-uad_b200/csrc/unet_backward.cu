// Backward pass of the planned UNet forward (training step: DDPM_2D.py:114-138, torch autograd over
// OpenAI_Unet.py:823-1006 in the reference).  The forward plan keeps every layer output in its own buffer, so the
// backward reads the saved activations and GroupNorm statistics directly; it is planned once per batch size as a
// second launch list:
//   data gradients    - the forward tcgen05 convolution kernels over transposed / flipped weight panels,
//   weight gradients  - conv_wgrad.cu (tcgen05, MN-major operands), fp32 split-K accumulation, then re-laid-out
//                       into the reference's OIHW parameter layout inside the caller's flat gradient buffer,
//   GroupNorm/FiLM/SiLU, resampling, biases, stem/head, embedding MLPs - backward.cu,
//   attention         - attention_bwd.cu.
#include <string.h>

#include <vector>

#include "attention.cuh"
#include "backward.cuh"
#include "conv_bwd.cuh"
#include "unet_engine.cuh"

#include <stdlib.h>

namespace cddpm {

int UNetEngine::grad_offset(int i, int64_t* off) const {
  if (i < 0 || i >= param_count()) return fail(kInvalidArgument, "param index out of range");
  *off = params_[i].goff;
  return kOk;
}

int UNetEngine::plan_backward(int B) {
  bwd_ops_.clear();
  bwd_flops_ = 0;
  const int fmt = cfg_.fmt;
  if (fmt != 1) return fail(kUnsupported, "the training step runs in bf16: create the engine with CDDPM_FMT_BF16");
  if (!fused_stats_ || cfg_.model_channels != 128)
    return fail(kUnsupported, "the backward pass needs model_channels == 128 (fused GroupNorm statistics, 128-channel tiles)");
  UNetEngine* self = this;
  const int mc = cfg_.model_channels;
  const int H0 = cfg_.image_h, W0 = cfg_.image_w;

  auto offset_of = [this](const std::string& name, int64_t* off) -> int {
    auto it = param_index_.find(name);
    if (it == param_index_.end()) return fail(kInvalidArgument, "backward: unknown parameter " + name);
    *off = params_[it->second].goff;
    return kOk;
  };
  auto balloc = [this](void** p, size_t bytes) -> int {
    void* q = nullptr;
    CDDPM_CUDA(cudaMalloc(&q, bytes + 256));
    act_owned_.push_back(q);
    *p = q;
    return kOk;
  };

  // ---- gradient buffers of the activations that carry gradients between layers
  std::map<const void*, void*> G, Gskip;
  std::map<const void*, int64_t> dbias_off;  // tensor -> gradient slot of the bias of the convolution that made it
  auto galloc = [&](std::map<const void*, void*>& m, const ActTensor& t) -> int {
    if (m.count(t.p)) return kOk;
    void* q = nullptr;
    CDDPM_TRY(balloc(&q, t.elems(B) * 2));
    m[t.p] = q;
    return kOk;
  };
  CDDPM_TRY(galloc(G, stem_out_));
  CDDPM_TRY(offset_of("input_blocks.0.0.bias", &dbias_off[stem_out_.p]));
  size_t max_elems = head_in_.elems(B);
  for (const ResPlan& rp : res_plans_) {
    const ResLayer& L = res_[rp.layer];
    CDDPM_TRY(galloc(G, rp.out));
    CDDPM_TRY(offset_of(L.prefix + ".out_layers.3.bias", &dbias_off[rp.out.p]));
    if (rp.has_a1) CDDPM_TRY(galloc(Gskip, rp.a1));
    const size_t big = static_cast<size_t>(B) * rp.out.H * rp.out.W * (L.cin > L.cout ? L.cin : L.cout);
    const size_t big_in = static_cast<size_t>(B) * rp.a0.H * rp.a0.W * L.cin;
    if (big > max_elems) max_elems = big;
    if (big_in > max_elems) max_elems = big_in;
  }
  for (const AttnPlan& ap : attn_plans_) {
    const AttnLayer& L = attn_[ap.layer];
    CDDPM_TRY(galloc(G, ap.out));
    CDDPM_TRY(offset_of(L.prefix + ".proj_out.bias", &dbias_off[ap.out.p]));
    if (ap.tQ.elems(B) > max_elems) max_elems = ap.tQ.elems(B);
  }
  void* scr[6];
  for (int i = 0; i < 6; ++i) CDDPM_TRY(balloc(&scr[i], max_elems * 2));

  // ---- fp32 work arenas, zeroed at the start of every backward: packed weight gradients, GroupNorm sums, dfilm
  size_t dwp_total = 0, sums_total = 0;
  for (const ResLayer& L : res_) dwp_total += static_cast<size_t>(L.cout) * 9 * L.cin + static_cast<size_t>(L.cout) * (9 * L.cout + L.cin);
  for (const AttnLayer& L : attn_) dwp_total += static_cast<size_t>(4) * L.ch * L.ch;
  for (const ResLayer& L : res_) sums_total += static_cast<size_t>(B) * 2 * (L.cin + L.cout);
  for (const AttnLayer& L : attn_) sums_total += static_cast<size_t>(B) * 2 * L.ch;
  sums_total += static_cast<size_t>(B) * 2 * mc;
  const size_t dfilm_total = static_cast<size_t>(B) * film_total_;
  float* arena = nullptr;
  const size_t arena_floats = dwp_total + sums_total + dfilm_total;
  {
    void* q = nullptr;
    CDDPM_TRY(balloc(&q, arena_floats * sizeof(float)));
    arena = reinterpret_cast<float*>(q);
  }
  float* dwp_next = arena;
  float* sums_next = arena + dwp_total;
  float* dfilm = arena + dwp_total + sums_total;
  auto take = [](float** next, size_t n) {
    float* p = *next;
    *next += n;
    return p;
  };
  bwd_ops_.push_back([=](cudaStream_t s) {
    CDDPM_CUDA(cudaMemsetAsync(arena, 0, arena_floats * sizeof(float), s));
    return check_cuda(cudaMemsetAsync(self->cur_grads_, 0, static_cast<size_t>(self->grad_total_) * sizeof(float), s),
                      "gradient memset");
  });

  int st = kOk;
  auto push_dgrad = [&](const void* dy, int cdy, int taps, int hh, int ww, const void* panel, int cout_t, void* out) {
    ConvDesc d;
    d.num_src = 1;
    d.src[0] = dy;
    d.src_c[0] = cdy;
    d.src_taps[0] = taps;
    d.B = B;
    d.H = hh;
    d.W = ww;
    d.Cout = cout_t;
    d.wpacked = panel;
    d.out = out;
    d.ab_format = fmt;
    if (st != kOk) return;
    if (!conv2_supported(d)) {
      st = fail(kUnsupported, "backward: data-gradient convolution outside the tcgen05 kernel's geometry");
      return;
    }
    std::shared_ptr<void> holder;
    st = build_conv2(d, &holder);
    if (st != kOk) return;
    bwd_flops_ += 2ll * hh * ww * cout_t * taps * cdy;
    bwd_ops_.push_back([holder](cudaStream_t s) { return launch_conv2(holder, s); });
  };
  auto push_wgrad = [&](WgradDesc d) {
    if (st != kOk) return;
    d.B = B;
    d.ab_format = fmt;
    std::shared_ptr<void> holder;
    st = build_wgrad(d, &holder);
    if (st != kOk) return;
    bwd_flops_ += wgrad_flops(d) / B;
    bwd_ops_.push_back([holder](cudaStream_t s) { return launch_wgrad(holder, s); });
  };
  // dw_packed -> the parameter's OIHW gradient slot
  auto push_unpack = [&](const float* dwp, int cout, int cin_total, int ks, int cin_off, int c_s, const std::string& name,
                         int ktot, int koff) {
    if (st != kOk) return;
    int64_t off = 0;
    st = offset_of(name, &off);
    if (st != kOk) return;
    bwd_ops_.push_back([=](cudaStream_t s) {
      return launch_unpack_conv_grad(dwp, cout, cin_total, ks, cin_off, c_s, self->cur_grads_ + off, ktot, koff, s);
    });
  };
  auto push_gn_bwd = [&](GnBwdArgs g, int C, const std::string& gamma_name, const std::string& beta_name,
                         const void* out0_tensor /* whose producer's bias gradient gets the column sums */) {
    if (st != kOk) return;
    int64_t og = 0, ob = 0, obias = -1;
    st = offset_of(gamma_name, &og);
    if (st == kOk) st = offset_of(beta_name, &ob);
    if (st != kOk) return;
    if (out0_tensor != nullptr) {
      auto it = dbias_off.find(out0_tensor);
      if (it != dbias_off.end()) obias = it->second;
    }
    g.B = B;
    g.fmt = fmt;
    g.sums = take(&sums_next, static_cast<size_t>(B) * 2 * C);
    bwd_ops_.push_back([=](cudaStream_t s) {
      GnBwdArgs a = g;
      a.dgamma = self->cur_grads_ + og;
      a.dbeta = self->cur_grads_ + ob;
      if (obias >= 0) a.bsum0 = self->cur_grads_ + obias;
      return launch_gn_bwd(a, s);
    });
  };

  // ================================================================== head: out = conv3x3(SiLU(GN(h)))
  {
    int64_t ow = 0, obias = 0;
    CDDPM_TRY(offset_of("out.2.weight", &ow));
    CDDPM_TRY(offset_of("out.2.bias", &obias));
    const void* tn = head_tN_.p;
    void* dN = scr[0];
    const int cc = head_in_.C;
    bwd_ops_.push_back([=](cudaStream_t s) {
      CDDPM_TRY(launch_sum_f32(self->cur_dout_, static_cast<int64_t>(B) * H0 * W0, self->cur_grads_ + obias, s));
      CDDPM_TRY(launch_wgrad_1ch(tn, self->cur_dout_, self->cur_grads_ + ow, B, H0, W0, cc, -1, fmt, s));
      return launch_head_bwd_data(self->cur_dout_, self->head_w, dN, B, H0, W0, cc, fmt, s);
    });
    GnBwdArgs g;
    g.x.p0 = head_in_.p;
    g.x.c0 = cc;
    g.H = H0;
    g.W = W0;
    g.stats0 = head_in_.stats;
    g.gamma = head_gn_w;
    g.beta = head_gn_b;
    g.silu = 1;
    g.dy = dN;
    g.out0 = G[head_in_.p];
    push_gn_bwd(g, cc, "out.0.weight", "out.0.bias", head_in_.p);
    CDDPM_TRY(st);
  }

  // ================================================================== blocks, last to first
  for (int si = static_cast<int>(steps_.size()) - 1; si >= 0; --si) {
    if (steps_[si].kind == 1) {
      const ResPlan& rp = res_plans_[steps_[si].idx];
      const ResLayer& L = res_[rp.layer];
      const int Ho = rp.out.H, Wo = rp.out.W, Hi = rp.a0.H, Wi = rp.a0.W;
      const void* dOut = G[rp.out.p];
      void* dtB = scr[0];
      void* dH = scr[1];
      void* dtA = scr[2];
      void* dXs = scr[3];
      void* dAr = scr[4];
      void* dSr = scr[5];
      const int k2 = 9 * L.cout + L.cin;
      // second convolution: out = conv3x3(tB; W2) + skip(x)
      push_dgrad(dOut, L.cout, 9, Ho, Wo, L.w2t, L.cout, dtB);
      {
        float* dwp = take(&dwp_next, static_cast<size_t>(L.cout) * k2);
        WgradDesc d;
        d.num_src = 1;
        d.src[0] = rp.tB.p;
        d.src_c[0] = L.cout;
        d.src_taps[0] = 9;
        if (L.has_skip) {
          d.src[d.num_src] = rp.a0.p;
          d.src_c[d.num_src] = rp.a0.C;
          d.src_taps[d.num_src] = 1;
          d.num_src++;
          if (rp.has_a1) {
            d.src[d.num_src] = rp.a1.p;
            d.src_c[d.num_src] = rp.a1.C;
            d.src_taps[d.num_src] = 1;
            d.num_src++;
          }
        } else {
          d.src[1] = nullptr;  // identity block: no parameters behind it
          d.src_c[1] = L.cout;
          d.src_taps[1] = 1;
          d.src_skip[1] = 1;
          d.num_src = 2;
        }
        d.dy = dOut;
        d.H = Ho;
        d.W = Wo;
        d.Cout = L.cout;
        d.dw = dwp;
        push_wgrad(d);
        push_unpack(dwp, L.cout, L.cout, 3, 0, L.cout, L.prefix + ".out_layers.3.weight", k2, 0);
        if (L.has_skip) {
          push_unpack(dwp, L.cout, L.cin, 1, 0, L.in_c0, L.prefix + ".skip_connection.weight", k2, 9 * L.cout);
          if (L.in_c1 > 0)
            push_unpack(dwp, L.cout, L.cin, 1, L.in_c0, L.in_c1, L.prefix + ".skip_connection.weight", k2,
                        9 * L.cout + L.in_c0);
          push_dgrad(dOut, L.cout, 1, Ho, Wo, L.wskipt, L.cin, dXs);
          // both biases see the same gradient
          int64_t o3 = 0, osk = 0;
          CDDPM_TRY(offset_of(L.prefix + ".out_layers.3.bias", &o3));
          CDDPM_TRY(offset_of(L.prefix + ".skip_connection.bias", &osk));
          const int n = L.cout;
          bwd_ops_.push_back([=](cudaStream_t s) {
            return launch_copy_f32(self->cur_grads_ + o3, self->cur_grads_ + osk, n, s);
          });
        }
        CDDPM_TRY(st);
      }
      // out_layers GroupNorm * (1 + scale) + shift, SiLU
      {
        GnBwdArgs g;
        g.x.p0 = rp.tH.p;
        g.x.c0 = L.cout;
        g.H = Ho;
        g.W = Wo;
        g.stats0 = rp.tH.stats;
        g.gamma = L.gn2_w;
        g.beta = L.gn2_b;
        g.film = film_out_;
        g.film_stride = film_total_;
        g.film_off = L.film_off;
        g.dfilm = dfilm;
        g.silu = 1;
        g.dy = dtB;
        g.out0 = dH;
        int64_t ob1 = 0;
        CDDPM_TRY(offset_of(L.prefix + ".in_layers.2.bias", &ob1));
        dbias_off[rp.tH.p] = ob1;
        push_gn_bwd(g, L.cout, L.prefix + ".out_layers.0.weight", L.prefix + ".out_layers.0.bias", rp.tH.p);
      }
      // first convolution: tH = conv3x3(tA; W1)
      push_dgrad(dH, L.cout, 9, Ho, Wo, L.w1t, L.cin, dtA);
      {
        float* dwp = take(&dwp_next, static_cast<size_t>(L.cout) * 9 * L.cin);
        WgradDesc d;
        d.num_src = 1;
        d.src[0] = rp.tA.p;
        d.src_c[0] = L.cin;
        d.src_taps[0] = 9;
        d.dy = dH;
        d.H = Ho;
        d.W = Wo;
        d.Cout = L.cout;
        d.dw = dwp;
        if (L.cin % 128 != 0) return fail(kUnsupported, "backward: block input channels must be a multiple of 128");
        push_wgrad(d);
        push_unpack(dwp, L.cout, L.cin, 3, 0, L.cin, L.prefix + ".in_layers.2.weight", 9 * L.cin, 0);
      }
      CDDPM_TRY(st);
      // resampling of both branches
      const void* dy1 = dtA;
      const void* skip_grad = L.has_skip ? dXs : dOut;
      if (L.mode != kResampleNone) {
        const int mode = L.mode, cin = L.cin;
        bwd_ops_.push_back([=](cudaStream_t s) {
          CDDPM_TRY(launch_resample_bwd(dtA, dAr, B, Hi, Wi, cin, mode, fmt, s));
          return launch_resample_bwd(dOut, dSr, B, Hi, Wi, cin, mode, fmt, s);
        });
        dy1 = dAr;
        skip_grad = dSr;
      }
      // in_layers GroupNorm + SiLU over the (possibly concatenated) block input
      {
        GnBwdArgs g;
        g.x.p0 = rp.a0.p;
        g.x.c0 = rp.a0.C;
        if (rp.has_a1) {
          g.x.p1 = rp.a1.p;
          g.x.c1 = rp.a1.C;
          g.stats1 = rp.a1.stats;
          g.out1 = Gskip[rp.a1.p];
        }
        g.H = Hi;
        g.W = Wi;
        g.stats0 = rp.a0.stats;
        g.gamma = L.gn1_w;
        g.beta = L.gn1_b;
        g.silu = 1;
        g.dy = dy1;
        g.add0 = skip_grad;
        if (!rp.has_a1 && Gskip.count(rp.a0.p)) g.add1 = Gskip[rp.a0.p];
        if (rp.has_a1 && Gskip.count(rp.a0.p)) return fail(kUnsupported, "backward: unexpected skip topology");
        if (!G.count(rp.a0.p)) return fail(kInvalidArgument, "backward: block input without a gradient buffer");
        g.out0 = G[rp.a0.p];
        push_gn_bwd(g, L.cin, L.prefix + ".in_layers.0.weight", L.prefix + ".in_layers.0.bias", rp.a0.p);
      }
      CDDPM_TRY(st);
    } else {
      const AttnPlan& ap = attn_plans_[steps_[si].idx];
      const AttnLayer& L = attn_[ap.layer];
      const int C = L.ch, hh = ap.x.H, ww = ap.x.W;
      const void* dOut = G[ap.out.p];
      void* dtA = scr[0];
      void* dQKV = scr[1];
      void* dtN = scr[2];
      // proj_out (1x1) + residual
      push_dgrad(dOut, C, 1, hh, ww, L.wprojt, C, dtA);
      {
        float* dwp = take(&dwp_next, static_cast<size_t>(C) * C);
        WgradDesc d;
        d.num_src = 1;
        d.src[0] = ap.tA.p;
        d.src_c[0] = C;
        d.src_taps[0] = 1;
        d.dy = dOut;
        d.H = hh;
        d.W = ww;
        d.Cout = C;
        d.dw = dwp;
        push_wgrad(d);
        push_unpack(dwp, C, C, 1, 0, C, L.prefix + ".proj_out.weight", C, 0);
      }
      // attention core
      {
        void* ascr = nullptr;
        CDDPM_TRY(balloc(&ascr, static_cast<size_t>(attention_bwd_scratch_elems(B, hh * ww, C)) * 2));
        const void* q = ap.tQ.p;
        int64_t obq = 0;
        CDDPM_TRY(offset_of(L.prefix + ".qkv.bias", &obq));
        const int Lq = hh * ww;
        bwd_ops_.push_back([=](cudaStream_t s) {
          CDDPM_TRY(launch_attention_bwd(q, dtA, dQKV, ascr, B, Lq, C, fmt, s));
          return launch_col_sum(dQKV, static_cast<int64_t>(B) * Lq, 3 * C, self->cur_grads_ + obq, fmt, s);
        });
        bwd_flops_ += 2ll * 7 * (C / 64) * Lq * Lq * 64;  // S, dP, dQ, dV, dK (+ the two recomputed in spirit)
      }
      // qkv (1x1)
      push_dgrad(dQKV, 3 * C, 1, hh, ww, L.wqkvt, C, dtN);
      {
        float* dwp = take(&dwp_next, static_cast<size_t>(3) * C * C);
        WgradDesc d;
        d.num_src = 1;
        d.src[0] = ap.tN.p;
        d.src_c[0] = C;
        d.src_taps[0] = 1;
        d.dy = dQKV;
        d.H = hh;
        d.W = ww;
        d.Cout = 3 * C;
        d.dw = dwp;
        push_wgrad(d);
        push_unpack(dwp, 3 * C, C, 1, 0, C, L.prefix + ".qkv.weight", C, 0);
      }
      CDDPM_TRY(st);
      {
        GnBwdArgs g;
        g.x.p0 = ap.x.p;
        g.x.c0 = C;
        g.H = hh;
        g.W = ww;
        g.stats0 = ap.x.stats;
        g.gamma = L.gn_w;
        g.beta = L.gn_b;
        g.silu = 0;
        g.dy = dtN;
        g.add0 = dOut;
        if (Gskip.count(ap.x.p)) g.add1 = Gskip[ap.x.p];
        if (!G.count(ap.x.p)) return fail(kInvalidArgument, "backward: attention input without a gradient buffer");
        g.out0 = G[ap.x.p];
        push_gn_bwd(g, C, L.prefix + ".norm.weight", L.prefix + ".norm.bias", ap.x.p);
      }
      CDDPM_TRY(st);
    }
  }

  // ================================================================== stem: weight gradient only (x is data)
  {
    int64_t ow = 0;
    CDDPM_TRY(offset_of("input_blocks.0.0.weight", &ow));
    const void* dstem = G[stem_out_.p];
    bwd_ops_.push_back([=](cudaStream_t s) {
      return launch_wgrad_1ch(dstem, self->cur_x_, self->cur_grads_ + ow, B, H0, W0, mc, +1, fmt, s);
    });
  }

  // ================================================================== embedding path (fp32, B rows)
  {
    const int E = emb_dim_, Hd = half_dim_, nc = cfg_.num_classes;
    float* buf = nullptr;
    // z1_t [B][Hd], z1_c [B][Hd], z2 [B][E], dz2 [B][E], dz1_t [B][Hd], dz1_c [B][Hd], sinus [B][mc]
    const size_t nfl = static_cast<size_t>(B) * (4 * Hd + 2 * E + mc);
    {
      void* q = nullptr;
      CDDPM_TRY(balloc(&q, nfl * sizeof(float)));
      buf = reinterpret_cast<float*>(q);
    }
    float* z1t = buf;
    float* z1c = z1t + static_cast<size_t>(B) * Hd;
    float* z2 = z1c + static_cast<size_t>(B) * Hd;
    float* dz2 = z2 + static_cast<size_t>(B) * E;
    float* dz1t = dz2 + static_cast<size_t>(B) * E;
    float* dz1c = dz1t + static_cast<size_t>(B) * Hd;
    float* sinus = dz1c + static_cast<size_t>(B) * Hd;
    int64_t o_te0w = 0, o_te0b = 0, o_te2w = 0, o_te2b = 0, o_le0w = 0, o_le0b = 0, o_le2w = 0, o_le2b = 0;
    CDDPM_TRY(offset_of("time_embed.0.weight", &o_te0w));
    CDDPM_TRY(offset_of("time_embed.0.bias", &o_te0b));
    CDDPM_TRY(offset_of("time_embed.2.weight", &o_te2w));
    CDDPM_TRY(offset_of("time_embed.2.bias", &o_te2b));
    if (nc > 0) {
      CDDPM_TRY(offset_of("label_emb.0.weight", &o_le0w));
      CDDPM_TRY(offset_of("label_emb.0.bias", &o_le0b));
      CDDPM_TRY(offset_of("label_emb.2.weight", &o_le2w));
      CDDPM_TRY(offset_of("label_emb.2.bias", &o_le2b));
    }
    // recompute the pre-activations of the two MLPs (the forward keeps only SiLU(.) in 16 bits): the same tensor-core
    // GEMMs as the forward, without the activation, fp32 out
    if (sin16_ == nullptr) return fail(kUnsupported, "backward: needs the tensor-core embedding path");
    auto push_flat = [&](const void* in16, int I, const void* w16, const float* bias, float* out, int O, int stride,
                         int col) {
      if (st != kOk) return;
      ConvDesc d;
      d.num_src = 1;
      d.src[0] = in16;
      d.src_c[0] = I;
      d.src_taps[0] = 1;
      d.flat_rows = B;
      d.Cout = O;
      d.wpacked = w16;
      d.bias = bias;
      d.out = out;
      d.out_is_f32 = 1;
      d.ab_format = fmt;
      d.out_stride = stride;
      d.out_col_off = col;
      auto p = std::make_shared<ConvIgemmParams>();
      st = build_conv_params(d, p.get());
      if (st != kOk) return;
      bwd_ops_.push_back([p](cudaStream_t s) { return launch_conv_igemm(*p, s); });
    };
    bwd_ops_.push_back([=](cudaStream_t s) { return launch_timestep_embedding(self->cur_t_, sinus, B, mc, s); });
    push_flat(sin16_, mc, te0_w16, te0_b, z1t, Hd, Hd, 0);
    push_flat(hid16_, Hd, te2_w16, te2_b, z2, Hd, E, 0);
    if (nc > 0) {
      push_flat(cond16_, nc, le0_w16, le0_b, z1c, Hd, Hd, 0);
      push_flat(hidc16_, Hd, le2_w16, le2_b, z2, Hd, E, Hd);
    }
    CDDPM_TRY(st);
    // FiLM projection: film = SiLU(z2) Wf^T + bf.  Every emb_layers.1 parameter is a row block of Wf.
    // d SiLU(z2) = dfilm Wf: one tensor-core GEMM over the transposed panel (K = all 2 * sum(cout) FiLM outputs)
    const int ftot = film_total_;
    {
      void* q = nullptr;
      CDDPM_TRY(balloc(&q, static_cast<size_t>(B) * ftot * 2));
      void* dfilm16 = q;
      bwd_ops_.push_back([=](cudaStream_t s) { return launch_to16(dfilm, dfilm16, B * ftot, fmt, s); });
      push_flat(dfilm16, ftot, film_w16t, nullptr, dz2, E, E, 0);
      CDDPM_TRY(st);
      bwd_ops_.push_back([=](cudaStream_t s) { return launch_mul_silu_grad(dz2, z2, static_cast<int64_t>(B) * E, s); });
    }
    {
      // all emb_layers.1 weight / bias gradients in ONE launch: per 32-row tile of the concatenated projection, the
      // offset of those rows inside the flat gradient buffer
      std::vector<int64_t> tw(ftot / 32), tb(ftot / 32);
      for (const ResLayer& L : res_) {
        int64_t ow = 0, ob = 0;
        CDDPM_TRY(offset_of(L.prefix + ".emb_layers.1.weight", &ow));
        CDDPM_TRY(offset_of(L.prefix + ".emb_layers.1.bias", &ob));
        if (L.film_off % 32 != 0 || (2 * L.cout) % 32 != 0) return fail(kUnsupported, "backward: FiLM rows not tiled by 32");
        for (int r = 0; r < 2 * L.cout; r += 32) {
          tw[(L.film_off + r) / 32] = ow + static_cast<int64_t>(r) * E;
          tb[(L.film_off + r) / 32] = ob + r;
        }
      }
      void* q = nullptr;
      CDDPM_TRY(balloc(&q, tw.size() * 2 * sizeof(int64_t)));
      int64_t* d_tw = reinterpret_cast<int64_t*>(q);
      int64_t* d_tb = d_tw + tw.size();
      CDDPM_CUDA(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
      CDDPM_CUDA(cudaMemcpy(d_tb, tb.data(), tb.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
      bwd_ops_.push_back([=](cudaStream_t s) {
        return launch_linear_bwd_weight_tiled(dfilm, ftot, z2, E, 1, self->cur_grads_, d_tw, d_tb, B, E, ftot, s);
      });
    }
    // time_embed: z2_t = W2 SiLU(z1_t) + b2, z1_t = W0 sin + b0
    bwd_ops_.push_back([=](cudaStream_t s) {
      float* g = self->cur_grads_;
      CDDPM_TRY(launch_linear_bwd_weight(dz2, E, z1t, Hd, 1, g + o_te2w, g + o_te2b, B, Hd, Hd, s));
      CDDPM_TRY(launch_linear_bwd_input(dz2, E, self->te2_w, nullptr, fmt, dz1t, Hd, z1t, Hd, B, Hd, Hd, s));
      CDDPM_TRY(launch_linear_bwd_weight(dz1t, Hd, sinus, mc, 0, g + o_te0w, g + o_te0b, B, mc, Hd, s));
      if (nc > 0) {
        CDDPM_TRY(launch_linear_bwd_weight(dz2 + Hd, E, z1c, Hd, 1, g + o_le2w, g + o_le2b, B, Hd, Hd, s));
        CDDPM_TRY(launch_linear_bwd_input(dz2 + Hd, E, self->le2_w, nullptr, fmt, dz1c, Hd, z1c, Hd, B, Hd, Hd, s));
        CDDPM_TRY(launch_linear_bwd_weight(dz1c, Hd, self->cur_cond_, nc, 0, g + o_le0w, g + o_le0b, B, nc, Hd, s));
        if (self->cur_dcond_ != nullptr)
          CDDPM_TRY(launch_linear_bwd_input(dz1c, Hd, self->le0_w, nullptr, fmt, self->cur_dcond_, nc, nullptr, 0, B, nc,
                                            Hd, s));
      }
      return static_cast<int>(kOk);
    });
  }
  bwd_planned_ = true;
  return kOk;
}

int UNetEngine::backward(const float* dout, float* grads, float* dcond, int B, cudaStream_t stream) {
  if (!dout || !grads) return fail(kInvalidArgument, "unet_backward: null pointer");
  if (plan_fused_ || plan_fused_head_)
    return fail(kNotReady, "unet_backward: the forward ran with an inference-only fusion (the head's GroupNorm inside "
                           "conv_out, or CDDPM_FUSE_GN=1); call cddpm_unet_set_training(h, 1) before the forward of a "
                           "training step");
  if (B != planned_B_ || forwards_on_plan_ < 1)
    return fail(kNotReady, "unet_backward: run the forward of this batch first");
  if (!bwd_planned_) {
    CDDPM_CUDA(cudaDeviceSynchronize());
    int st = plan_backward(B);
    if (st != kOk) {
      bwd_ops_.clear();
      bwd_planned_ = false;
      return st;
    }
  }
  cur_dout_ = dout;
  cur_grads_ = grads;
  cur_dcond_ = dcond;
  auto eager = [&]() -> int {
    for (auto& op : bwd_ops_) CDDPM_TRY(op(stream));
    return kOk;
  };
  static const bool graphs = [] {
    const char* e = getenv("CDDPM_BWD_GRAPH");
    return !(e != nullptr && e[0] == '0');
  }();
  // (a caller whose buffers never repeat would pay a capture on every second call: stay eager once captures clearly
  // outnumber replays)
  if (!graphs || (bwd_captures_ >= 8 && bwd_replays_ < 2 * bwd_captures_)) return eager();
  const std::vector<const void*> key = {dout, grads, dcond, cur_x_, cur_t_, cur_cond_};
  BwdGraph* slot = nullptr;
  for (BwdGraph& g : bwd_graphs_)
    if (g.key == key) slot = &g;
  if (slot == nullptr) {
    if (bwd_graphs_.size() < 4) {
      bwd_graphs_.emplace_back();
      slot = &bwd_graphs_.back();
    } else {
      slot = &bwd_graphs_[0];
      for (BwdGraph& g : bwd_graphs_)
        if (g.last_use < slot->last_use) slot = &g;
      if (slot->exec != nullptr) cudaGraphExecDestroy(slot->exec);
      *slot = BwdGraph();
    }
    slot->key = key;
  }
  slot->last_use = ++bwd_clock_;
  if (slot->exec == nullptr) {
    if (slot->seen++ == 0) return eager();
    if (cap_stream_ == nullptr) CDDPM_CUDA(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
    CDDPM_CUDA(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal));
    int st = kOk;
    for (auto& op : bwd_ops_) {
      st = op(cap_stream_);
      if (st != kOk) break;
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cap_stream_, &graph);
    if (st != kOk || ce != cudaSuccess) {
      if (graph != nullptr) cudaGraphDestroy(graph);
      return st != kOk ? st : check_cuda(ce, "cudaStreamEndCapture (backward)");
    }
    const cudaError_t ie = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    CDDPM_TRY(check_cuda(ie, "cudaGraphInstantiate (backward)"));
    ++bwd_captures_;
  } else {
    ++bwd_replays_;
  }
  CDDPM_CUDA(cudaGraphLaunch(slot->exec, stream));
  return kOk;
}

}  // namespace cddpm
