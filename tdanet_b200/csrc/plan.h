// Workspace plan shared by the forward engine (engine.cu) and the backward pass (backward.cu).
//
// Inference: one block arena reused by every UConvBlock iteration.  Training: the block arena is
// replicated num_blocks times (every GlobLN-delimited tensor of every iteration is kept for the backward
// pass), followed by a gradient arena that is reused by every iteration of the backward sweep.
// Host-only code: compiles under nvcc and, for the CPU emulation build of the backward pass, under g++.
#pragma once
#include "kernels.h"
#include <cstring>
#include <string>
#include <vector>

namespace td {

constexpr int TDANET_DW_REPLICAS = 8;

struct Named {
  std::string name;
  size_t off;  // bytes
  int64_t dims[3];
  int esize;   // 4: float, 8: double
};

struct Plan {
  int B, T, Tp, rest;
  int depth, L[TDANET_MAX_DEPTH], Lb;
  size_t bytes = 0;
  std::vector<Named> named;
  // ---- common activations
  size_t enc, x0, u[2], masked;
  size_t st_enc;  // [B,2] double
  // TF32 auxiliary weight copies
  size_t aux_proj, aux_res, aux_in, aux_out, aux_fc1, aux_fc2, aux_pool[TDANET_MAX_DEPTH];
  size_t aux_pool_dwT[TDANET_MAX_DEPTH];  // FORK: conv_pool[j].dw_conv.weight transposed to [k][C]
  // ---- block arena [blk_begin, blk_begin + blk_stride)
  size_t blk_begin = 0, blk_stride = 0;
  int n_blk = 1;
  size_t proj, spp[TDANET_MAX_DEPTH], expanded[TDANET_MAX_DEPTH];
  size_t ga_in, attn_in, qkv, attn_ctx, attn_out, ga_mid, fc1, ffn_dw, fc2, ga_out;
  size_t pool_dw[TDANET_MAX_DEPTH], pool_pw[TDANET_MAX_DEPTH];
  size_t fused_a, fused_b;  // x_fused[depth-2] and its "global" partner of the first top-down step, materialised
  // closed-form loc_glo_fus coefficient tables (BEST), [B,6,C] per scale
  size_t inj_coef[TDANET_MAX_DEPTH];
  // statistics arena (zeroed once per block)
  size_t stats_begin, stats_end;
  // deterministic mode (inference, det_mode()): the integer pairs the sums are accumulated in (DetRef, common.cuh),
  // 4 bytes of shadow per byte of the statistics arena / of st_enc
  bool det = false;
  size_t det_shadow = 0, det_enc_shadow = 0;
  // per-item sum / sum of squares in double: [B,2] ([B,2,2] for st_la_g: global_act, global_embedding)
  size_t st_proj, st_fc1, st_fc2, st_pool[TDANET_MAX_DEPTH], st_spp[TDANET_MAX_DEPTH],
      st_la_l[TDANET_MAX_DEPTH], st_la_g[TDANET_MAX_DEPTH];
  // per-channel sums in float [B,2,C] (BEST: inputs of the closed-form loc_glo_fus statistics)
  size_t st_spp_ch[TDANET_MAX_DEPTH], st_g;
  // ---- training only
  bool train = false;
  size_t bin, y;                    // block arena: input of the block (blocks > 0) / res_conv output before concat_block
  size_t fused[TDANET_MAX_DEPTH];   // block arena: x_fused[k] materialised
  size_t st_lgf[TDANET_MAX_DEPTH];  // block arena: [B,3,2] double sums of loc_glo_fus[k] local / act / embedding conv outputs
  // block arena: dropout keep-masks (bytes) of this iteration; valid when drop_elem / drop_item
  bool drop_elem = false, drop_item = false;  // cfg->dropout > 0 / cfg->drop_path > 0
  size_t m_begin = 0, m_att = 0, m_ao = 0, m_f1 = 0, m_f2 = 0, m_dp = 0, m_end = 0;
  size_t n_att = 0;                 // bytes of m_att
  size_t mlogit;                    // mask_net output before ReLU [B, L0, n_src*Nb]
  size_t nenc;                      // GlobLN(enc) [B, L0, Nb] (operand of the bottleneck weight gradient)
  // transposed weights for the data-gradient GEMMs (+ their TF32 copies)
  size_t wt_proj, wt_res, wt_in, wt_out, wt_fc1, wt_fc2, wt_pool[TDANET_MAX_DEPTH];
  size_t auxt_proj, auxt_res, auxt_in, auxt_out, auxt_fc1, auxt_fc2, auxt_pool[TDANET_MAX_DEPTH];
  // FORK conv_pool backward: gradient w.r.t. the pw_conv output (through its GlobLN) per scale, and w.r.t. the dw output
  size_t g_pool_x[TDANET_MAX_DEPTH], g_pool_dw;
  // gradient arena
  size_t g_masked, g_enc, g_x0, g_u[2], g_proj, g_spp[TDANET_MAX_DEPTH], g_fused[TDANET_MAX_DEPTH],
      g_exp[TDANET_MAX_DEPTH];
  // LA backward temporaries, TDANET_LA_TEMP_SETS sets: the top-down steps alternate between sets 0 / 1 so that the
  // local-branch kernels of step i (side stream) overlap step i+1; the loc_glo_fus chain (side stream) owns set 2
  size_t t_dloc[4], t_rawa[4], t_dact[4], t_demb[4], t_rawb[4], t_rawe[4];  // sets 0/1: top-down steps, 2/3: loc_glo_fus chains
  // per top-down step: raw conv outputs of the global branch (pass G of the LA backward), computed ahead of the chain
  size_t t_rawb_step[TDANET_MAX_DEPTH], t_rawe_step[TDANET_MAX_DEPTH];
  size_t g_ga_out2;  // second accumulator of the global feature's gradient (the two loc_glo_fus streams), summed at the join
  size_t g_ga_out, g_fc2, g_ffn, g_fc1, g_ga_mid, g_attn_out, g_ctx, g_qkv, g_attn_in, g_ga_in;
  size_t att_p, att_ds, ln_rows;
  // replicated accumulators of the depthwise weight / bias gradients: [REP_COUNT][rep_floats]
  size_t rep_arena, rep_floats;
  // backward GlobLN sums S1 = sum(gamma*dy), S2 = sum(gamma*dy*xhat): [B,2] double each
  size_t bs_begin, bs_end, bs_enc, bs_proj, bs_fc1, bs_fc2, bs_spp[TDANET_MAX_DEPTH], bs_la[TDANET_MAX_DEPTH][3],
      bs_lgf[TDANET_MAX_DEPTH][3], bs_pool[TDANET_MAX_DEPTH];

  size_t take(size_t nbytes) {
    size_t o = bytes;
    bytes += (nbytes + 255) / 256 * 256;
    return o;
  }
  size_t act(const char* name, int64_t L_, int64_t C_) {
    size_t o = take((size_t)B * L_ * C_ * sizeof(float));
    if (name) named.push_back({name, o, {B, L_, C_}, 4});
    return o;
  }
  // one of the LARGE activations (proj_1x1 output, spp_dw outputs, expanded, materialised x_fused): stored as bf16 in
  // the first half of its fp32-sized buffer when act_dtype is bf16
  size_t big(const char* name, int64_t L_, int64_t C_) {
    size_t o = take((size_t)B * L_ * C_ * sizeof(float));
    if (name) named.push_back({name, o, {B, L_, C_}, act_bf16 ? 2 : 4});
    return o;
  }
  bool act_bf16 = false;
  size_t dstat(const char* name, int n) {
    size_t o = take((size_t)B * n * 2 * sizeof(double));
    if (name) named.push_back({name, o, {B, n, 2}, 8});
    return o;
  }
  // byte offset of a block-arena tensor for iteration `blk`
  size_t blk_off(size_t off, int blk) const {
    return (off >= blk_begin && off < blk_begin + blk_stride) ? off + (size_t)blk * blk_stride : off;
  }
};

// Process-wide switch of the deterministic-statistics mode (tdanet_set_deterministic; TDANET_DETERMINISTIC=1 in the
// environment turns it on from the start).  It changes the inference workspace size, so a caller sets it before
// sizing a workspace, not between tdanet_workspace_bytes and tdanet_forward.
inline int& det_mode() {
  static int mode = [] {
    const char* e = getenv("TDANET_DETERMINISTIC");
    return (e && e[0] != '0' && e[0] != 0) ? 1 : 0;
  }();
  return mode;
}

static inline int check_config(const tdanet_config_t* c) {
  TD_REQUIRE(c != nullptr, "config is NULL");
  TD_REQUIRE(c->variant >= TDANET_BEST && c->variant <= TDANET_ORIGIN, "unknown variant %d", c->variant);
  TD_REQUIRE(c->depth >= 2 && c->depth <= TDANET_MAX_DEPTH, "upsampling_depth %d outside [2, %d]", c->depth, TDANET_MAX_DEPTH);
  TD_REQUIRE(c->num_blocks >= 1, "num_blocks %d", c->num_blocks);
  TD_REQUIRE(c->out_channels > 0 && c->out_channels % 16 == 0, "out_channels %d must be a multiple of 16", c->out_channels);
  TD_REQUIRE(c->in_channels > 0 && c->in_channels % 16 == 0, "in_channels %d must be a multiple of 16", c->in_channels);
  TD_REQUIRE(c->n_head > 0 && c->in_channels % c->n_head == 0, "in_channels %d not divisible by n_head %d", c->in_channels, c->n_head);
  TD_REQUIRE(c->enc_kernel > 0 && c->enc_kernel % 4 == 0 && c->enc_stride == c->enc_kernel / 4, "encoder window %d / hop %d", c->enc_kernel, c->enc_stride);
  TD_REQUIRE(c->num_sources == 2 || c->num_sources == 3, "num_sources %d", c->num_sources);
  TD_REQUIRE(c->gemm_mode >= TDANET_GEMM_FP32 && c->gemm_mode <= TDANET_GEMM_TF32X3, "gemm_mode %d", c->gemm_mode);
  TD_REQUIRE(c->act_dtype == TDANET_ACT_F32 || c->act_dtype == TDANET_ACT_BF16, "act_dtype %d", c->act_dtype);
  if (c->act_dtype == TDANET_ACT_BF16) {
    TD_REQUIRE(c->gemm_mode != TDANET_GEMM_FP32, "bf16 activation storage needs a tensor-core gemm_mode");
    TD_REQUIRE(c->out_channels % 32 == 0 && c->in_channels % 64 == 0,
               "bf16 activation storage needs out_channels %% 32 == 0 and in_channels %% 64 == 0 (got %d / %d)",
               c->out_channels, c->in_channels);
  }
  if (c->variant == TDANET_MULTRES) {
    TD_REQUIRE(c->enc_convs >= 1 && c->enc_convs <= TDANET_MAX_ENC && c->out_channels % c->enc_convs == 0,
               "MULTRES: out_channels %d not divisible by kernels %d", c->out_channels, c->enc_convs);
    TD_REQUIRE(c->n_basis == c->out_channels, "MULTRES: n_basis %d != out_channels %d", c->n_basis, c->out_channels);
  } else {
    TD_REQUIRE(c->enc_convs == 1, "enc_convs %d", c->enc_convs);
    TD_REQUIRE(c->n_basis == c->enc_kernel / 2 + 1, "n_basis %d != K/2+1", c->n_basis);
  }
  return 0;
}

// What the training path supports in this build (everything else raises instead of falling back).
static inline int check_train_config(const tdanet_config_t* c) {
  if (int e = check_config(c)) return e;
  // act_dtype bf16: the large activations a forward keeps for the backward pass (proj, spp_dw outputs, x_fused,
  // expanded) are stored as bf16; arithmetic, statistics, gradients and parameters stay fp32 ("precision: 16" of
  // configs/tdanet.yml:41 in this implementation's terms)
  TD_REQUIRE(c->dropout >= 0.f && c->dropout < 1.f && c->drop_path >= 0.f && c->drop_path < 1.f,
             "dropout %g / drop_path %g outside [0, 1)", (double)c->dropout, (double)c->drop_path);
  return 0;
}

// index of the x_fused tensor the first top-down step takes as its "global" input (python x_fused[i-1], i = depth-2)
static inline int first_step_partner(int depth) { return (depth - 3 + depth) % depth; }

static inline int make_plan(const tdanet_config_t* c, int B, int T, Plan& p, bool train = false) {
  if (int e = train ? check_train_config(c) : check_config(c)) return e;
  TD_REQUIRE(B > 0 && T > 0, "batch %d / n_samples %d", B, T);
  const int K = c->enc_kernel, S = c->enc_stride, C = c->in_channels, cc = c->out_channels;
  p.B = B;
  p.T = T;
  p.train = train;
  p.act_bf16 = c->act_dtype == TDANET_ACT_BF16;
  // pad_input (TDANet_best.py:465-479)
  p.rest = K - (S + T % K) % K;
  p.Tp = T + p.rest + 2 * (K - S);
  p.depth = c->depth;
  p.L[0] = p.Tp / S + 1;  // Conv1d(k, stride S, padding k/2), k even
  for (int k = 1; k < c->depth; ++k) p.L[k] = (p.L[k - 1] - 1) / 2 + 1;
  p.Lb = p.L[c->depth - 1];
  const int depth = c->depth, L0 = p.L[0], Lb = p.Lb, Nb = c->n_basis, NS = c->num_sources;
  char nm[32];

  // ---- common
  p.enc = p.act("enc", L0, Nb);
  p.x0 = p.act("x0", L0, cc);
  p.u[0] = p.act("u0", L0, cc);
  p.u[1] = p.act("u1", L0, cc);
  p.masked = p.act("masked", L0, NS * Nb);
  p.st_enc = p.dstat("st_enc", 1);
  auto wbuf = [&](size_t n) { return p.take(n * sizeof(float)); };
  p.aux_proj = wbuf((size_t)C * cc);
  p.aux_res = wbuf((size_t)cc * C);
  p.aux_in = wbuf((size_t)3 * C * C);
  p.aux_out = wbuf((size_t)C * C);
  p.aux_fc1 = wbuf((size_t)2 * C * C);
  p.aux_fc2 = wbuf((size_t)2 * C * C);
  for (int k = 0; k < depth; ++k) {
    p.aux_pool[k] = c->variant == TDANET_FORK ? wbuf((size_t)C * C) : 0;
    p.aux_pool_dwT[k] = c->variant == TDANET_FORK ? wbuf((size_t)C * (k == 0 ? 5 : 2 * (1 << k) + 1)) : 0;
  }
  if (train) {
    p.mlogit = p.act("mlogit", L0, NS * Nb);
    p.nenc = p.act("nenc", L0, Nb);
    p.wt_proj = wbuf((size_t)C * cc); p.auxt_proj = wbuf((size_t)C * cc);
    p.wt_res = wbuf((size_t)C * cc); p.auxt_res = wbuf((size_t)C * cc);
    p.wt_in = wbuf((size_t)3 * C * C); p.auxt_in = wbuf((size_t)3 * C * C);
    p.wt_out = wbuf((size_t)C * C); p.auxt_out = wbuf((size_t)C * C);
    p.wt_fc1 = wbuf((size_t)2 * C * C); p.auxt_fc1 = wbuf((size_t)2 * C * C);
    p.wt_fc2 = wbuf((size_t)2 * C * C); p.auxt_fc2 = wbuf((size_t)2 * C * C);
    for (int k = 0; k < depth; ++k) {
      p.wt_pool[k] = c->variant == TDANET_FORK ? wbuf((size_t)C * C) : 0;
      p.auxt_pool[k] = c->variant == TDANET_FORK ? wbuf((size_t)C * C) : 0;
    }
  }

  // ---- block arena
  p.blk_begin = p.bytes;
  p.proj = p.big("proj", L0, C);
  for (int k = 0; k < depth; ++k) {
    snprintf(nm, sizeof nm, "spp%d", k);
    p.spp[k] = p.big(nm, p.L[k], C);
  }
  for (int k = 0; k < depth - 1; ++k) {
    snprintf(nm, sizeof nm, "expanded%d", k);
    p.expanded[k] = p.big(nm, p.L[k], C);
  }
  p.ga_in = p.act("ga_in", Lb, C);
  p.attn_in = p.act("attn_in", Lb, C);
  p.qkv = p.act("qkv", Lb, 3 * C);
  p.attn_ctx = p.act("attn_ctx", Lb, C);
  p.attn_out = p.act("attn_out", Lb, C);
  p.ga_mid = p.act("ga_mid", Lb, C);
  p.fc1 = p.act("fc1", Lb, 2 * C);
  p.ffn_dw = p.act("ffn_dw", Lb, 2 * C);
  p.fc2 = p.act("fc2", Lb, C);
  p.ga_out = p.act("ga_out", Lb, C);
  for (int k = 0; k < depth; ++k) {
    // FORK: conv_pool dw / pw outputs.  BEST / MULTRES: pool_pw[k] holds the pooled raw spp_dw[k] output.
    if (c->variant == TDANET_FORK) {
      snprintf(nm, sizeof nm, "pool_dw%d", k);
      p.pool_dw[k] = p.act(nm, Lb, C);
    }
    snprintf(nm, sizeof nm, "pool_pw%d", k);
    p.pool_pw[k] = p.act(nm, Lb, C);
  }
  if (train) {
    p.bin = p.act("bin", L0, cc);
    p.y = p.act("y", L0, cc);
    for (int k = 0; k < depth; ++k) {
      snprintf(nm, sizeof nm, "fused%d", k);
      p.fused[k] = p.big(nm, p.L[k], C);
    }
    p.fused_a = p.fused[depth - 2];
    p.fused_b = p.fused[first_step_partner(depth)];
    p.drop_elem = c->dropout > 0.f;
    p.drop_item = c->drop_path > 0.f;
    if (p.drop_elem || p.drop_item) {
      auto bytes_of = [&](const char* name, int64_t d0, int64_t d1, int64_t d2) {
        size_t o = p.take((size_t)(d0 * d1 * d2));
        p.named.push_back({name, o, {d0, d1, d2}, 1});
        return o;
      };
      // attention problems: (batch group, time index) with `group` tokens each; MULTRES: one per item with Lb tokens
      const bool time_axis = c->variant == TDANET_MULTRES;
      const int group = time_axis ? Lb : (c->attn_group > 0 ? c->attn_group : B);
      const int64_t nprob = time_axis ? B : (int64_t)(B / group) * Lb;
      p.m_begin = p.bytes;
      p.n_att = (size_t)nprob * c->n_head * group * group;
      if (p.drop_elem) {
        p.m_att = bytes_of("m_att", nprob * c->n_head, group, group);
        p.m_ao = bytes_of("m_ao", B, Lb, C);
        p.m_f1 = bytes_of("m_f1", B, Lb, 2 * C);
        p.m_f2 = bytes_of("m_f2", B, Lb, C);
      }
      if (p.drop_item) p.m_dp = bytes_of("m_dp", 2, B, 1);
      p.m_end = p.bytes;
    }
  } else {
    p.fused_a = p.act("fused_a", p.L[depth - 2], C);
    p.fused_b = p.act("fused_b", p.L[first_step_partner(depth)], C);
  }
  auto tab = [&](const char* name, int planes, int ch) {
    size_t o = p.take((size_t)B * planes * ch * sizeof(float));
    if (name) p.named.push_back({name, o, {B, planes, ch}, 4});
    return o;
  };
  for (int k = 0; k < depth; ++k) {
    snprintf(nm, sizeof nm, "inj_coef%d", k);
    p.inj_coef[k] = tab(nm, 6, C);
  }
  p.stats_begin = p.bytes;
  p.st_proj = p.dstat("st_proj", 1);
  p.st_fc1 = p.dstat("st_fc1", 1);
  p.st_fc2 = p.dstat("st_fc2", 1);
  for (int k = 0; k < depth; ++k) {
    snprintf(nm, sizeof nm, "st_pool%d", k);
    p.st_pool[k] = p.dstat(nm, 1);
    snprintf(nm, sizeof nm, "st_spp%d", k);
    p.st_spp[k] = p.dstat(nm, 1);
    snprintf(nm, sizeof nm, "st_la_l%d", k);
    p.st_la_l[k] = p.dstat(nm, 1);
    snprintf(nm, sizeof nm, "st_la_g%d", k);
    p.st_la_g[k] = p.dstat(nm, 2);
    p.st_spp_ch[k] = tab(nullptr, 2, C);
    if (train) {
      snprintf(nm, sizeof nm, "st_lgf%d", k);
      p.st_lgf[k] = p.dstat(nm, 3);
    }
  }
  p.st_g = tab(nullptr, 2, C);
  p.stats_end = p.bytes;
  p.blk_stride = p.bytes - p.blk_begin;
  p.n_blk = train ? c->num_blocks : 1;
  p.bytes = p.blk_begin + (size_t)p.n_blk * p.blk_stride;

  // "block_out" aliases the u buffer the last block writes
  p.named.push_back({"block_out", p.u[(c->num_blocks - 1) & 1], {B, L0, cc}, 4});
  p.named.push_back({"u", p.u[(c->num_blocks & 1)], {B, L0, cc}, 4});
  if (!train && det_mode()) {
    p.det = true;
    p.det_shadow = p.take(4 * (p.stats_end - p.stats_begin));
    p.det_enc_shadow = p.take(4 * (size_t)B * 2 * sizeof(double));
  }
  if (!train) return 0;

  // ---- gradient arena (reused by every iteration of the backward sweep)
  p.g_masked = p.act("g_masked", L0, NS * Nb);
  p.g_enc = p.act("g_enc", L0, Nb);
  p.g_x0 = p.act("g_x0", L0, cc);
  p.g_u[0] = p.act("g_u0", L0, cc);
  p.g_u[1] = p.act("g_u1", L0, cc);
  p.g_proj = p.act("g_proj", L0, C);
  int Lg_max = Lb;  // longest "global" operand of any LA (the first top-down step takes a finer tensor)
  for (int k = 0; k < depth; ++k) {
    snprintf(nm, sizeof nm, "g_spp%d", k);
    p.g_spp[k] = p.act(nm, p.L[k], C);
    // BEST: x_fused[k] = loc_glo_fus(n_k, g) has its own gradient; FORK: x_fused[k] = near(g) + n_k, so the gradient
    // w.r.t. x_fused[k] IS a gradient w.r.t. n_k and lands directly in g_spp[k]
    snprintf(nm, sizeof nm, "g_fused%d", k);
    p.g_fused[k] = c->variant == TDANET_BEST ? p.act(nm, p.L[k], C) : p.g_spp[k];
    if (k < depth - 1) {
      snprintf(nm, sizeof nm, "g_exp%d", k);
      p.g_exp[k] = p.act(nm, p.L[k], C);
    }
  }
  for (int i = 0; i < depth - 1; ++i) {
    const int Lg = i == depth - 2 ? p.L[first_step_partner(depth)] : p.L[i + 1];
    Lg_max = Lg > Lg_max ? Lg : Lg_max;
  }
  for (int s = 0; s < 4; ++s) {
    p.t_dloc[s] = p.act(nullptr, L0, C);
    p.t_rawa[s] = p.act(nullptr, L0, C);
    p.t_dact[s] = p.act(nullptr, Lg_max, C);
    p.t_demb[s] = p.act(nullptr, Lg_max, C);
    p.t_rawb[s] = p.act(nullptr, Lg_max, C);
    p.t_rawe[s] = p.act(nullptr, Lg_max, C);
  }
  for (int i = 0; i < depth - 1; ++i) {
    const int Lg = i == depth - 2 ? p.L[first_step_partner(depth)] : p.L[i + 1];
    p.t_rawb_step[i] = p.act(nullptr, Lg, C);
    p.t_rawe_step[i] = p.act(nullptr, Lg, C);
  }
  p.g_ga_out = p.act("g_ga_out", Lb, C);
  p.g_ga_out2 = p.act(nullptr, Lb, C);
  p.g_fc2 = p.act("g_fc2", Lb, C);
  p.g_ffn = p.act("g_ffn", Lb, 2 * C);
  p.g_fc1 = p.act("g_fc1", Lb, 2 * C);
  p.g_ga_mid = p.act("g_ga_mid", Lb, C);
  p.g_attn_out = p.act("g_attn_out", Lb, C);
  p.g_ctx = p.act("g_ctx", Lb, C);
  p.g_qkv = p.act("g_qkv", Lb, 3 * C);
  p.g_attn_in = p.act("g_attn_in", Lb, C);
  p.g_ga_in = p.act("g_ga_in", Lb, C);
  if (c->variant == TDANET_FORK) {
    for (int k = 0; k < depth; ++k) p.g_pool_x[k] = p.act(nullptr, Lb, C);
    p.g_pool_dw = p.act(nullptr, Lb, C);
  }
  {
    const bool time_axis = c->variant == TDANET_MULTRES;
    const int group = time_axis ? Lb : (c->attn_group > 0 ? c->attn_group : B);
    const size_t n = (size_t)B * Lb * c->n_head * group;  // [problem, head, query, key]
    p.att_p = p.take(n * sizeof(float));
    p.att_ds = p.take(n * sizeof(float));
    p.ln_rows = p.take((size_t)B * Lb * 4 * sizeof(float));
  }
  {
    // every depthwise conv weight (+ bias) of the block: spp_dw, ffn dwconv, last_layer (k5), loc_glo_fus (k1)
    size_t n = (size_t)depth * (C * 5 + C) + (size_t)(2 * C * 5 + 2 * C) + (size_t)(depth - 1) * 3 * C * 5 + (size_t)depth * 3 * C;
    // (the FORK conv_pool depthwise weights accumulate straight into the caller's buffers)
    p.rep_floats = (n + 63) / 64 * 64;
    p.rep_arena = p.take(p.rep_floats * TDANET_DW_REPLICAS * sizeof(float));
  }
  p.bs_enc = p.dstat("bs_enc", 1);
  p.bs_begin = p.bytes;
  p.bs_proj = p.dstat("bs_proj", 1);
  p.bs_fc1 = p.dstat("bs_fc1", 1);
  p.bs_fc2 = p.dstat("bs_fc2", 1);
  for (int k = 0; k < depth; ++k) {
    p.bs_spp[k] = p.dstat(nullptr, 1);
    p.bs_pool[k] = p.dstat(nullptr, 1);
    for (int j = 0; j < 3; ++j) {
      p.bs_la[k][j] = p.dstat(nullptr, 1);
      p.bs_lgf[k][j] = p.dstat(nullptr, 1);
    }
  }
  p.bs_end = p.bytes;
  return 0;
}

// Launch context of one pass over the model.
struct Ctx {
  const tdanet_config_t* c;
  const tdanet_weights_t* w;
  const Plan* p;
  char* ws;
  cudaStream_t st;
  int blk = 0;  // UConvBlock iteration whose arena block-local offsets resolve to (training plan)
  // producers of GEMM-only operands store TF32-rounded values when the tensor-core path is on
  int rnd() const { return c->gemm_mode != TDANET_GEMM_FP32; }
  // large activations (proj, spp, expanded, materialised x_fused) stored as bf16
  int bf() const { return c->act_dtype == TDANET_ACT_BF16; }
  template <class T = float>
  T* at(size_t off) const { return reinterpret_cast<T*>(ws + p->blk_off(off, blk)); }
  template <class T = float>
  T* at_blk(size_t off, int b) const { return reinterpret_cast<T*>(ws + p->blk_off(off, b)); }
  // deterministic mode: the exact accumulators of the statistics arena / of the encoder's statistics
  DetRef det() const { return p->det ? DetRef{ws + p->stats_begin, ws + p->det_shadow} : DetRef{}; }
  DetRef det_enc() const { return p->det ? DetRef{ws + p->st_enc, ws + p->det_enc_shadow} : DetRef{}; }
};

static inline NormRef norm_ref(const Ctx& x, size_t stats_off, int item_stride, double count, const float* gamma,
                               const float* beta) {
  return NormRef{x.at<double>(stats_off), item_stride, count, gamma, beta};
}

}  // namespace td
