// Host-side orchestration of one TDANet forward pass: workspace planning and the launch sequence.
//
// Replaces the Python control flow of TDANet*.forward / Recurrent.forward / UConvBlock.forward
// (TDANet_best.py:342-399, 482-521; TDANet.py:586-636, 769-785, 869-909; TDANet_mult_tes.py:391-434,
// 540-579).  Everything between two GlobLN-delimited tensors is one launch; see DESIGN.md for the
// stage list and the bytes each stage moves.
#include "plan.h"

namespace td {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

size_t pit_scratch_floats(int B);

bool g_pdl = [] {
  const char* e = getenv("TDANET_PDL");
  return !(e && e[0] == '0');
}();

// ----------------------------------------------------------------------------- event profiler
bool g_profile = false;
thread_local const char* g_tag = nullptr;
struct ProfRec {
  const char* name;
  cudaEvent_t ev[2];
};
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;

void profile_mark(const char* name, cudaStream_t st, bool begin) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  cudaEvent_t e;
  if (!g_ev_pool.empty()) {
    e = g_ev_pool.back();
    g_ev_pool.pop_back();
  } else if (cudaEventCreate(&e) != cudaSuccess) {
    return;
  }
  cudaEventRecord(e, st);
  if (begin) {
    g_prof.push_back(ProfRec{g_tag ? g_tag : name, {e, nullptr}});
  } else if (!g_prof.empty()) {
    g_prof.back().ev[1] = e;
  }
}

// ----------------------------------------------------------------------------- forward
// Side stream of the forward pass (lazily created per device, never destroyed): runs the statistics of the LOCAL
// branch of every top-down step, which depend only on spp_dw[*] and the global feature, next to the top-down chain
// (small, dependent launches that leave most of the HBM bandwidth idle).  Fork / join are cudaEventRecord /
// cudaStreamWaitEvent pairs, which stream capture turns into graph edges.
struct FwdSide {
  cudaStream_t s = nullptr;
  std::vector<cudaEvent_t> events;
  size_t next = 0;
  int init() {
    if (s) return 0;
    TD_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    events.resize(256);
    for (auto& e : events) TD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
  }
  // everything enqueued on `from` so far happens before whatever is enqueued on `to` from now on
  int order(cudaStream_t from, cudaStream_t to) {
    cudaEvent_t e = events[next++ % events.size()];
    TD_CUDA(cudaEventRecord(e, from));
    TD_CUDA(cudaStreamWaitEvent(to, e, 0));
    return 0;
  }
};
static FwdSide g_fside[16];

static thread_local FwdSide* t_fside = nullptr;  // set by forward() for the duration of the call

// Deterministic mode (DetRef, common.cuh): the producers of a statistic accumulate it exactly in integer pairs; this
// kernel, enqueued right behind the producer on the producer's stream, writes the value into the slots the consumers
// read (n values of `elem` bytes from `slots`: double sums per item, or float sums per channel).
__global__ void det_finalize_kernel(DetRef d, char* slots, int n, int elem) {
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  char* slot = slots + (size_t)i * elem;
  const double v = det_value(reinterpret_cast<const unsigned long long*>(d.shadow + 4 * (slot - d.base)));
  if (elem == 8) *reinterpret_cast<double*>(slot) = v;
  else *reinterpret_cast<float*>(slot) = (float)v;
}
static int det_finalize(const DetRef& d, void* slots, size_t n, int elem, cudaStream_t st) {
  if (!d.base || !slots || n == 0) return 0;
  TD_LAUNCH(det_finalize_kernel, dim3((unsigned)((n + 255) / 256)), 256, 0, st, d, reinterpret_cast<char*>(slots), (int)n, elem);
  return 0;
}

static int gemm(const Ctx& x, GemmArgs& g, size_t aux_off) {
  if (g.stats) g.det = x.det();
  if (x.c->gemm_mode == TDANET_GEMM_FP32) {
    if (int e = launch_gemm_simt(g, x.st)) return e;
  } else {
    g.W_aux = x.at(aux_off);
    g.narrow = !x.p->train;   // inference: the main stream has the SMs to itself while a GEMM runs
    if (int e = launch_gemm_tc(g, x.c->gemm_mode, x.st)) return e;
  }
  return g.stats ? det_finalize(g.det, g.stats, (size_t)g.B * 2, 8, x.st) : 0;
}

// launch_dw5 with the statistics of its output (per item, optionally per channel) made exact in deterministic mode
static int dw5_stats(const Ctx& x, DwArgs& d, cudaStream_t st) {
  d.det = x.det();
  if (int e = launch_dw5(d, st)) return e;
  if (int e = det_finalize(d.det, d.stats, (size_t)d.B * d.nw * 2, 8, st)) return e;
  return det_finalize(d.det, d.chstats, (size_t)d.B * d.nw * 2 * d.C, 4, st);
}
static int la_local_stats(const Ctx& x, DwArgs* steps, int n, cudaStream_t st) {
  for (int i = 0; i < n; ++i) steps[i].det = x.det();
  if (int e = launch_la_local_stats(steps, n, st)) return e;
  for (int i = 0; i < n; ++i)
    if (int e = det_finalize(steps[i].det, steps[i].stats, (size_t)steps[i].B * 2, 8, st)) return e;
  return 0;
}

static int prepare_weights(const Ctx& x) {
  const tdanet_config_t* c = x.c;
  if (c->variant == TDANET_FORK)  // conv_pool[j] depthwise weights [C][k] -> [k][C] for the input-stationary kernel
    for (int j = 0; j < c->depth; ++j)
      if (int e = launch_weight_transpose(x.w->conv_pool[j].dw_w, x.at(x.p->aux_pool_dwT[j]), c->in_channels,
                                          j == 0 ? 5 : 2 * (1 << j) + 1, x.st)) return e;
  if (c->gemm_mode == TDANET_GEMM_FP32) return 0;
  const size_t C = c->in_channels, cc = c->out_channels;
  const int m = c->gemm_mode;
  if (int e = launch_tf32_prepare(x.w->proj.w, x.at(x.p->aux_proj), C * cc, m, x.st)) return e;
  if (x.bf()) {  // res_conv reads its bf16 operand against a bf16 copy of the weight
    if (int e = launch_bf16_prepare(x.w->res_w, x.at(x.p->aux_res), C * cc, x.st)) return e;
  } else {
    if (int e = launch_tf32_prepare(x.w->res_w, x.at(x.p->aux_res), C * cc, m, x.st)) return e;
  }
  if (int e = launch_tf32_prepare(x.w->in_proj_w, x.at(x.p->aux_in), 3 * C * C, m, x.st)) return e;
  if (int e = launch_tf32_prepare(x.w->out_proj_w, x.at(x.p->aux_out), C * C, m, x.st)) return e;
  if (int e = launch_tf32_prepare(x.w->fc1.w, x.at(x.p->aux_fc1), 2 * C * C, m, x.st)) return e;
  if (int e = launch_tf32_prepare(x.w->fc2.w, x.at(x.p->aux_fc2), 2 * C * C, m, x.st)) return e;
  if (c->variant == TDANET_FORK)
    for (int k = 0; k < c->depth; ++k)
      if (int e = launch_tf32_prepare(x.w->conv_pool[k].pw_w, x.at(x.p->aux_pool[k]), C * C, m, x.st)) return e;
  return 0;
}

// The global-branch statistics of the first top-down step computed by the kernel that materialises its operand.
// Measured on B200: at the training batch of 8 it removes a launch from a launch-bound chain (52.6 -> 53.5 steps/s);
// at B = 64 the fused kernel is 47 us slower per block than the 20 us statistics launch it replaces (17.96 -> 18.40 ms
// per inference step), so inference keeps the separate launch.  TDANET_FIRST_STATS=0 / 1 forces either form.
static bool first_stats_fused(bool train) {
  static const int knob = getenv("TDANET_FIRST_STATS") ? atoi(getenv("TDANET_FIRST_STATS")) : -1;
  return knob < 0 ? train : knob != 0;
}

static SrcDesc plain_src(const float* x, int L) {
  SrcDesc s{};
  s.x = x; s.L = L;
  return s;
}
static SrcDesc affine_src(const float* x, int L, const NormRef& norm, const float* slope = nullptr) {
  SrcDesc s{};
  s.x = x; s.L = L; s.norm = norm; s.slope = slope;
  return s;
}

// The bottom-scale block: GA / GlobalAttention (TDANet_best.py:254-264)
static int global_attention(const Ctx& x, bool ln_pe_done) {
  const tdanet_config_t* c = x.c;
  const tdanet_weights_t* w = x.w;
  const Plan& p = *x.p;
  const int B = p.B, C = c->in_channels, Lb = p.Lb;
  const bool time_axis = c->variant == TDANET_MULTRES;
  const int group = c->attn_group > 0 ? c->attn_group : B;
  TD_REQUIRE(time_axis || B % group == 0, "batch %d is not a multiple of attn_group %d", B, group);
  TD_REQUIRE(w->pe_rows >= Lb, "positional encoding has %d rows, need %d", w->pe_rows, Lb);

  // attn_in_norm + positional encoding
  Tag tag("bottom_misc");
  if (!ln_pe_done)  // else: fused into the kernel that produced ga_in
    if (int e = launch_ln_pe(x.at(p.ga_in), w->ln1_w, w->ln1_b, w->pe, x.at(p.attn_in), B, Lb, C, x.rnd(), x.st)) return e;
  GemmArgs g{};
  // Every item attends alone (attention group 1: the chunk-at-a-time long-form loop, audio_test_css.py:110-111, and
  // any B = 1 call): the softmax over a single key is exactly 1, so the attention output IS the value row - only the
  // V third of in_proj is computed, straight into the context buffer, and the attention launch disappears.
  const bool single_key = group == 1 && !time_axis && !p.train;
  g.A = x.at(p.attn_in); g.W = w->in_proj_w; g.bias = w->in_proj_b; g.D = x.at(p.qkv);
  g.B = B; g.L = Lb; g.N = 3 * C; g.K = C; g.epi = EPI_BIAS;
  if (single_key) {
    g.W = w->in_proj_w + (size_t)2 * C * C; g.bias = w->in_proj_b + 2 * C; g.D = x.at(p.attn_ctx); g.N = C;
    Tag t("gemm_in_proj");
    if (x.c->gemm_mode == TDANET_GEMM_FP32) {
      if (int e = launch_gemm_simt(g, x.st)) return e;
    } else {
      g.W_aux = x.at(p.aux_in) + (size_t)2 * C * C;   // the same rows of the prepared (TF32) copy
      g.narrow = 1;
      if (int e = launch_gemm_tc(g, x.c->gemm_mode, x.st)) return e;
    }
  } else {
    Tag t("gemm_in_proj");
    if (int e = gemm(x, g, p.aux_in)) return e;
  }
  // training-mode multipliers (nn.Dropout / DropPath keep-masks drawn by launch_dropout_masks): null in eval
  const float ik = p.drop_elem ? 1.f / (1.f - c->dropout) : 1.f, ikp = p.drop_item ? 1.f / (1.f - c->drop_path) : 1.f;
  const uint8_t* m_att = p.drop_elem ? x.at<uint8_t>(p.m_att) : nullptr;
  const uint8_t* m_ao = p.drop_elem ? x.at<uint8_t>(p.m_ao) : nullptr;
  const uint8_t* m_f1 = p.drop_elem ? x.at<uint8_t>(p.m_f1) : nullptr;
  const uint8_t* m_f2 = p.drop_elem ? x.at<uint8_t>(p.m_f2) : nullptr;
  const uint8_t* m_dp = p.drop_item ? x.at<uint8_t>(p.m_dp) : nullptr;
  if (!single_key) {
    Tag t("attention");
    if (int e = launch_attention(x.at(p.qkv), x.at(p.attn_ctx), B, Lb, C, c->n_head, group, time_axis, x.rnd(), m_att, ik, x.st)) return e;
  }
  g = GemmArgs{};
  g.A = x.at(p.attn_ctx); g.W = w->out_proj_w; g.bias = w->out_proj_b; g.D = x.at(p.attn_out);
  g.B = B; g.L = Lb; g.N = C; g.K = C; g.epi = EPI_BIAS;
  { Tag t("gemm_out_proj"); if (int e = gemm(x, g, p.aux_out)) return e; }
  // x + DropPath(LayerNorm(out + dropout(out)))  [BEST/FORK]   |   x + LayerNorm(pe_in + out)  [MULTRES]
  // training: attn_out <- out * (1 + mask/keep) in place (the backward pass expects it in this form)
  //           MULTRES: attn_out <- out * mask/keep
  if (m_ao)
    if (int e = launch_mask_scale(x.at(p.attn_out), x.at(p.attn_out), (size_t)B * Lb * C, m_ao, time_axis ? 0.f : 1.f, ik, nullptr, 1.f, 1, 0, x.st)) return e;
  if (int e = launch_ln_residual(x.at(p.attn_out), x.at(p.attn_in), x.at(p.ga_in), w->ln2_w, w->ln2_b,
                                 x.at(p.ga_mid), time_axis ? 0 : (m_ao ? 2 : 1), DropRef{nullptr, 1.f, m_dp, ikp}, B, Lb, C, x.st)) return e;
  // FFN: fc1 (1x1, no bias) -> gLN -> dw k5 + bias -> ReLU -> fc2 (1x1, no bias) -> gLN
  g = GemmArgs{};
  g.A = x.at(p.ga_mid); g.W = w->fc1.w; g.bias = nullptr; g.D = x.at(p.fc1);
  g.B = B; g.L = Lb; g.N = 2 * C; g.K = C; g.epi = EPI_BIAS; g.stats = x.at<double>(p.st_fc1);
  { Tag t("gemm_fc1"); if (int e = gemm(x, g, p.aux_fc1)) return e; }
  DwArgs d{};
  d.src = affine_src(x.at(p.fc1), Lb, norm_ref(x, p.st_fc1, 2, (double)Lb * 2 * C, w->fc1.gamma, w->fc1.beta));
  d.kind = SRC_AFFINE; d.B = B; d.C = 2 * C; d.Lout = Lb; d.stride = 1; d.nw = 1;
  d.w[0] = w->ffn_dw_w; d.bias[0] = w->ffn_dw_b; d.out = x.at(p.ffn_dw); d.relu = 1; d.round_out = x.rnd();
  { Tag t("ffn_dw"); if (int e = launch_dw5(d, x.st)) return e; }
  if (m_f1)  // FFN.drop after the ReLU, in place (x > 0 still marks the live elements for the backward pass)
    if (int e = launch_mask_scale(x.at(p.ffn_dw), x.at(p.ffn_dw), (size_t)B * Lb * 2 * C, m_f1, 0.f, ik, nullptr, 1.f, 1, x.rnd(), x.st)) return e;
  g = GemmArgs{};
  g.A = x.at(p.ffn_dw); g.W = w->fc2.w; g.bias = nullptr; g.D = x.at(p.fc2);
  g.B = B; g.L = Lb; g.N = C; g.K = 2 * C; g.epi = EPI_BIAS; g.stats = x.at<double>(p.st_fc2);
  { Tag t("gemm_fc2"); if (int e = gemm(x, g, p.aux_fc2)) return e; }
  // global_f = x + gLN(fc2); BEST also needs its per-channel sums for the closed-form loc_glo_fus
  if (int e = launch_affine_residual(x.at(p.fc2), norm_ref(x, p.st_fc2, 2, (double)Lb * C, w->fc2.gamma, w->fc2.beta),
                                x.at(p.ga_mid), x.at(p.ga_out),
                                c->variant == TDANET_BEST ? x.at(p.st_g) : nullptr,
                                DropRef{m_f2, ik, m_dp ? m_dp + B : nullptr, ikp}, B, Lb, C, x.st, x.det()))
    return e;
  return c->variant == TDANET_BEST ? det_finalize(x.det(), x.at(p.st_g), (size_t)B * 2 * C, 4, x.st) : 0;
}

// One UConvBlock (TDANet_best.py:342-380) including the concat_block that feeds the next one.
// Item order of the big streaming launches.  A tensor of the two finest scales (132 / 263 MB at B = 64) is larger
// than the 126 MB L2: a consumer that walks the items in the producer's order finds none of it, one that starts
// with the item written last finds the tail of it.  So the launches of the chain alternate their direction:
// proj (first to last) -> spp_dw[0] (last to first) -> spp_dw[1] (first to last) ...; the global statistics of a
// top-down step and res_conv read what a first-to-last la_stream wrote, last to first.  TDANET_L2_ORDER=0: off.
// (bit mask, tuning aid: 1 spp_dw, 2 global statistics, 4 res_conv)
static bool l2_order(int bit) {
  static const int mask = getenv("TDANET_L2_ORDER") ? atoi(getenv("TDANET_L2_ORDER")) : 1;
  return (mask & bit) != 0;
}

static int uconv_block(const Ctx& x, const float* in, float* out, bool last) {
  const tdanet_config_t* c = x.c;
  const tdanet_weights_t* w = x.w;
  const Plan& p = *x.p;
  const int B = p.B, C = c->in_channels, cc = c->out_channels, depth = c->depth, Lb = p.Lb;
  TD_CUDA(cudaMemsetAsync(x.at<char>(p.st_proj), 0, p.stats_end - p.st_proj, x.st));
  if (p.det) TD_CUDA(cudaMemsetAsync(x.ws + p.det_shadow, 0, 4 * (p.stats_end - p.stats_begin), x.st));

  // proj_1x1: 1x1 conv c -> C (+bias); GlobLN + PReLU are applied by the consumer on load
  GemmArgs g{};
  g.A = in; g.W = w->proj.w; g.bias = w->proj.b; g.D = x.at(p.proj);
  g.B = B; g.L = p.L[0]; g.N = C; g.K = cc; g.epi = EPI_BIAS; g.stats = x.at<double>(p.st_proj);
  g.d_bf16 = x.bf();
  { Tag t("gemm_proj"); if (int e = gemm(x, g, p.aux_proj)) return e; }
  Tag tag("coef");
  auto spp_norm = [&](int k) {
    return norm_ref(x, p.st_spp[k], 2, (double)p.L[k] * C, w->spp_dw[k].gamma, w->spp_dw[k].beta);
  };
  // spp_dw[0..depth-1]: depthwise k5 (stride 1, then 2), raw output + per-channel sums
  for (int k = 0; k < depth; ++k) {
    DwArgs d{};
    if (k == 0) {
      d.src = affine_src(x.at(p.proj), p.L[0], norm_ref(x, p.st_proj, 2, (double)p.L[0] * C, w->proj.gamma, w->proj.beta),
                         w->proj_prelu);
      d.kind = SRC_AFFINE_PRELU;
    } else {
      d.src = affine_src(x.at(p.spp[k - 1]), p.L[k - 1], spp_norm(k - 1));
      d.kind = SRC_AFFINE;
    }
    d.B = B; d.C = C; d.Lout = p.L[k]; d.stride = k == 0 ? 1 : 2; d.nw = 1; d.act_bf16 = x.bf();
    d.w[0] = w->spp_dw[k].w; d.bias[0] = w->spp_dw[k].b; d.out = x.at(p.spp[k]);
    d.stats = x.at<double>(p.st_spp[k]);
    d.chstats = c->variant == TDANET_BEST ? x.at(p.st_spp_ch[k]) : nullptr;
    {
      static const int spp_mask = getenv("TDANET_SPP_REV") ? atoi(getenv("TDANET_SPP_REV")) : 1;  // scale 0 only (measured: -0.10 ms; further scales, the global statistics and res_conv: neutral or worse)
      d.rev = l2_order(1) && ((spp_mask >> k) & 1);
    }
    if (c->variant != TDANET_FORK) { d.pool_out = x.at(p.pool_pw[k]); d.Lb = Lb; }  // pooled raw output P_k
    { Tag t(k == 0 ? "spp_dw0" : "spp_dw_s2"); if (int e = dw5_stats(x, d, x.st)) return e; }
  }
  // global feature at the bottom scale
  PoolArgs pa{};
  pa.n = depth; pa.B = B; pa.C = C; pa.Lb = Lb; pa.out = x.at(p.ga_in);
  pa.ln_w = w->ln1_w; pa.ln_b = w->ln1_b; pa.pe = w->pe; pa.ln_out = x.at(p.attn_in); pa.ln_round = x.rnd();
  TD_REQUIRE(w->pe_rows >= Lb, "positional encoding has %d rows, need %d", w->pe_rows, Lb);
  if (c->variant == TDANET_FORK) {
    // conv_pool[depth-1-k](spp[k]): dw (k = 2s+1, stride s = 2^(depth-1-k)) -> 1x1 -> gLN; summed
    for (int k = 0; k < depth; ++k) {
      const int j = depth - 1 - k, s = 1 << j, ks = j == 0 ? 5 : 2 * s + 1;
      const tdanet_sepconvnorm_t& q = w->conv_pool[j];
      Tag tp("conv_pool");
      if (int e = launch_dw_generic(affine_src(x.at(p.spp[k]), p.L[k], spp_norm(k)), SRC_AFFINE, B, C, Lb,
                                    ks, s, q.dw_w, x.at(p.aux_pool_dwT[j]), q.dw_b, x.at(p.pool_dw[k]), x.rnd(), x.bf(), x.st)) return e;
      g = GemmArgs{};
      g.A = x.at(p.pool_dw[k]); g.W = q.pw_w; g.bias = q.pw_b; g.D = x.at(p.pool_pw[k]);
      g.B = B; g.L = Lb; g.N = C; g.K = C; g.epi = EPI_BIAS; g.stats = x.at<double>(p.st_pool[k]);
      if (int e = gemm(x, g, p.aux_pool[j])) return e;
      pa.x[k] = x.at(p.pool_pw[k]); pa.norm[k] = norm_ref(x, p.st_pool[k], 2, (double)Lb * C, q.gamma, q.beta); pa.L[k] = Lb;
    }
    if (int e = launch_affine_sum(pa, x.st)) return e;
  } else {
    // sum_k avgpool(gLN_k(out_k)) = sum_k gLN-affine_k(P_k): the pooling itself rode in the spp_dw kernels
    for (int k = 0; k < depth; ++k) { pa.x[k] = x.at(p.pool_pw[k]); pa.norm[k] = spp_norm(k); pa.L[k] = Lb; }
    Tag tp("pool_sum");
    if (int e = launch_affine_sum(pa, x.st)) return e;
  }
  if (int e = global_attention(x, launch_affine_sum_fuses_ln(pa))) return e;

  // injection of the global feature: never materialised, recomputed on load by the LA kernels
  const int inj_kind = c->variant == TDANET_BEST ? SRC_INJECT_GATE : SRC_INJECT_ADD;
  const float* gf = x.at(p.ga_out);
  auto inj_src = [&](int k) {
    SrcDesc s{};
    s.x = x.at(p.spp[k]); s.L = p.L[k];
    if (c->variant == TDANET_BEST) s.coef = x.at(p.inj_coef[k]);
    else s.norm = spp_norm(k);
    s.g = gf; s.Lg = Lb; s.gscale = nearest_scale(Lb, p.L[k]);
    return s;
  };
  if (c->variant == TDANET_BEST) {
    InjectCoefArgs ia{};
    ia.n = depth; ia.g_stats = x.at(p.st_g); ia.Lg = Lb;
    for (int k = 0; k < depth; ++k) {
      ia.spp_stats[k] = x.at(p.st_spp_ch[k]); ia.L[k] = p.L[k]; ia.spp[k] = w->spp_dw[k];
      ia.la[k] = w->loc_glo_fus[k]; ia.coef[k] = x.at(p.inj_coef[k]);
      ia.conv_stats[k] = p.train ? x.at<double>(p.st_lgf[k]) : nullptr;
    }
    if (int e = launch_coef_inject_gate(ia, B, C, x.st)) return e;
  }
  // training: every live x_fused[k] is kept for the backward pass (the top-down steps below still recompute it on
  // load, like inference: the streaming kernels are the fast path; both evaluate the same expression)
  bool fused_live[TDANET_MAX_DEPTH] = {}, fused_deferred[TDANET_MAX_DEPTH] = {};
  cudaEvent_t fused_ready = nullptr;
  if (p.train) {
    for (int k = 0; k < depth - 1; ++k) fused_live[k] = true;
    fused_live[first_step_partner(depth)] = true;
    Tag t("fused_materialize");
    const int gi = first_step_partner(depth);
    // only the two operands of the first top-down step are read by this forward; the others (the two finest scales,
    // most of the bytes) are kept for the backward pass alone and are written on the side stream, after the local
    // statistics (TDANET_MAT_SIDE=0: everything on the caller's stream, round 2 first session)
    static const bool mat_side = !(getenv("TDANET_MAT_SIDE") && atoi(getenv("TDANET_MAT_SIDE")) == 0);
    const bool side_ok = mat_side && t_fside != nullptr && depth >= 3;
    for (int k = 0; k < depth; ++k)
      if (fused_live[k]) {
        if (side_ok && k != gi && k != depth - 2) { fused_deferred[k] = true; continue; }
        // x_fused[gi] is the "global" operand of the first top-down step: its global-branch GlobLN statistics ride
        // in the kernel that writes it (no statistics launch for that step)
        const bool st = k == gi && first_stats_fused(true);
        const tdanet_la_t& la0 = w->last_layer[depth - 2];
        if (int e = launch_inject_materialize(inj_src(k), inj_kind, B, C, x.at(p.fused[k]), x.bf(), x.st,
                                              st ? la0.global_act.w : nullptr, st ? la0.global_embedding.w : nullptr,
                                              st ? x.at<double>(p.st_la_g[depth - 2]) : nullptr)) return e;
      }
  }
  // statistics of the local branch of every top-down step: independent of the chain below, so they run on the side
  // stream - the coarse scales first (one launch; the first steps of the chain need them), then the finest scale,
  // which is most of the bytes and is only needed by the last step.  Each la_combine waits for its own statistics.
  cudaEvent_t local_ready[TDANET_MAX_DEPTH] = {};
  {
    DwArgs dl[TDANET_MAX_DEPTH];
    for (int i = 0; i < depth - 1; ++i) {
      dl[i] = DwArgs{};
      dl[i].src = inj_src(i); dl[i].kind = inj_kind; dl[i].B = B; dl[i].C = C; dl[i].Lout = p.L[i]; dl[i].stride = 1;
      dl[i].nw = 1; dl[i].w[0] = w->last_layer[i].local_embedding.w; dl[i].stats = x.at<double>(p.st_la_l[i]);
      dl[i].act_bf16 = x.bf();
    }
    Tag t("la_stats_local");
    FwdSide* fs = t_fside;
    if (fs == nullptr || depth < 3) {
      if (int e = la_local_stats(x, dl, depth - 1, x.st)) return e;
    } else {
      if (int e = fs->order(x.st, fs->s)) return e;
      if (int e = la_local_stats(x, dl + 1, depth - 2, fs->s)) return e;  // scales 1 .. depth-2
      cudaEvent_t ec = fs->events[fs->next++ % fs->events.size()];
      TD_CUDA(cudaEventRecord(ec, fs->s));
      for (int i = 1; i < depth - 1; ++i) local_ready[i] = ec;
      if (int e = la_local_stats(x, dl, 1, fs->s)) return e;              // scale 0
      cudaEvent_t e0 = fs->events[fs->next++ % fs->events.size()];
      TD_CUDA(cudaEventRecord(e0, fs->s));
      local_ready[0] = e0;
      bool any = false;
      for (int k = 0; k < depth; ++k)
        if (fused_deferred[k]) {
          Tag tm("fused_materialize");
          if (int e = launch_inject_materialize(inj_src(k), inj_kind, B, C, x.at(p.fused[k]), x.bf(), fs->s)) return e;
          any = true;
        }
      if (any) {
        fused_ready = fs->events[fs->next++ % fs->events.size()];
        TD_CUDA(cudaEventRecord(fused_ready, fs->s));
      }
    }
  }
  // top-down fusion: last_layer[i](x_fused[i], i == depth-2 ? x_fused[i-1] : expanded)
  bool first_global_stats_done = p.train && first_stats_fused(true);   // training: done by fused_materialize above
  for (int i = depth - 2; i >= 0; --i) {
    const tdanet_la_t& la = w->last_layer[i];
    SrcDesc loc = inj_src(i), glo;
    int lkind = inj_kind, gkind;
    if (i == depth - 2 && p.train) {
      // fused_a / fused_b alias x_fused[depth-2] / x_fused[partner], materialised above
      const int gi = first_step_partner(depth);
      loc = plain_src(x.at(p.fused[i]), p.L[i]);
      glo = plain_src(x.at(p.fused[gi]), p.L[gi]);
      lkind = gkind = SRC_PLAIN;
    } else if (i == depth - 2) {
      // python x_fused[i-1]: the finer neighbour (or [-1]).  Both operands of this step are small and its
      // nearest *down*-sampling would re-derive every injected row five times: write the two tensors out once.
      const SrcDesc gsrc = inj_src((i - 1 + depth) % depth);
      Tag t("la_combine_first");
      static const bool one_launch = !(getenv("TDANET_MAT2") && atoi(getenv("TDANET_MAT2")) == 0);
      if (one_launch) {
        const bool st = first_stats_fused(false) && !p.det;  // its statistics are not on the exact path
        if (int e = launch_inject_materialize2(loc, x.at(p.fused_a), gsrc, x.at(p.fused_b), inj_kind, B, C, x.bf(), x.st,
                                               st ? la.global_act.w : nullptr, st ? la.global_embedding.w : nullptr,
                                               st ? x.at<double>(p.st_la_g[i]) : nullptr)) return e;
        first_global_stats_done = st;
      } else {
        if (int e = launch_inject_materialize(loc, inj_kind, B, C, x.at(p.fused_a), x.bf(), x.st)) return e;
        if (int e = launch_inject_materialize(gsrc, inj_kind, B, C, x.at(p.fused_b), x.bf(), x.st)) return e;
      }
      loc = plain_src(x.at(p.fused_a), loc.L);
      glo = plain_src(x.at(p.fused_b), gsrc.L);
      lkind = gkind = SRC_PLAIN;
    } else {
      glo = plain_src(x.at(p.expanded[i + 1]), p.L[i + 1]);
      gkind = SRC_PLAIN;
    }
    DwArgs dg{};
    dg.src = glo; dg.kind = gkind; dg.B = B; dg.C = C; dg.Lout = glo.L; dg.stride = 1; dg.nw = 2; dg.act_bf16 = x.bf();
    dg.w[0] = la.global_act.w; dg.w[1] = la.global_embedding.w; dg.stats = x.at<double>(p.st_la_g[i]);
    dg.rev = l2_order(2);
    if (!(i == depth - 2 && first_global_stats_done)) {
      Tag t("la_stats_global");
      if (int e = dw5_stats(x, dg, x.st)) return e;
    }
    LaArgs l{};
    l.loc = loc; l.glo = glo; l.lkind = lkind; l.gkind = gkind; l.B = B; l.C = C;
    l.wl = la.local_embedding.w; l.wa = la.global_act.w; l.we = la.global_embedding.w;
    l.nL = norm_ref(x, p.st_la_l[i], 2, (double)loc.L * C, la.local_embedding.gamma, la.local_embedding.beta);
    l.nA = norm_ref(x, p.st_la_g[i], 4, (double)glo.L * C, la.global_act.gamma, la.global_act.beta);
    l.nE = norm_ref(x, p.st_la_g[i] + 2 * sizeof(double), 4, (double)glo.L * C, la.global_embedding.gamma, la.global_embedding.beta);
    l.out = x.at(p.expanded[i]); l.scale = nearest_scale(glo.L, loc.L);
    l.round_out = i == 0 && x.rnd() && !x.bf();  // expanded[0] only feeds res_conv
    l.act_bf16 = x.bf();
    if (local_ready[i]) TD_CUDA(cudaStreamWaitEvent(x.st, local_ready[i], 0));  // st_la_l[i] (side stream)
    { Tag t(i == depth - 2 ? "la_combine_first" : "la_combine"); if (int e = launch_la_combine(l, x.st)) return e; }
  }
  if (fused_ready) TD_CUDA(cudaStreamWaitEvent(x.st, fused_ready, 0));  // the side stream's x_fused[k] (long done)
  // res_conv + residual (+ concat_block for the next iteration)
  g = GemmArgs{};
  g.A = x.at(p.expanded[0]); g.W = w->res_w; g.bias = w->res_b; g.D = out;
  g.B = B; g.L = p.L[0]; g.N = cc; g.K = C; g.epi = EPI_RESIDUAL;
  g.resid = in; g.mix = x.at(p.x0); g.cw = w->concat_w; g.cb = w->concat_b; g.cslope = w->concat_prelu; g.last = last;
  g.a_bf16 = x.bf();
  g.rev = l2_order(4);
  Tag tr("gemm_res_conv");
  if (!p.train) return gemm(x, g, p.aux_res);
  // training: keep y = res_conv(expanded) + residual, and apply concat_block in its own launch
  g.D = x.at(p.y); g.last = 1;
  if (int e = gemm(x, g, p.aux_res)) return e;
  if (last) return 0;
  return launch_concat(x.at(p.y), x.at(p.x0), w->concat_w, w->concat_b, w->concat_prelu, out, B * p.L[0], cc, x.st);
}

static int forward(const tdanet_config_t* c, const tdanet_weights_t* w, const float* wav, int B, int T,
                   float* est, void* workspace, size_t ws_bytes, cudaStream_t st, bool train,
                   uint64_t* rng_state = nullptr) {
  Plan p;
  if (int e = make_plan(c, B, T, p, train)) return e;
  TD_REQUIRE(!(p.drop_elem || p.drop_item) || rng_state,
             "dropout %g / drop_path %g > 0 needs tdanet_forward_train_rng", (double)c->dropout, (double)c->drop_path);
  TD_REQUIRE(w && wav && est && workspace, "NULL argument");
  if (ws_bytes < p.bytes) return fail(TDANET_ENOSPACE, "workspace has %zu bytes, need %zu", ws_bytes, p.bytes);
  TD_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  Ctx x{c, w, &p, (char*)workspace, st};
  const int K = c->enc_kernel, S = c->enc_stride, Nb = c->n_basis, cc = c->out_channels, L0 = p.L[0];
  int cur_dev = 0;
  TD_CUDA(cudaGetDevice(&cur_dev));
  std::lock_guard<std::mutex> enqueue_lock(device_enqueue_mutex(cur_dev));
  {
    int dev = cur_dev;
    t_fside = nullptr;
    if (dev >= 0 && dev < 16) {
      if (int e = g_fside[dev].init()) return e;
      t_fside = &g_fside[dev];
    }
  }

  TD_CUDA(cudaMemsetAsync(x.ws + p.st_enc, 0, (size_t)B * 2 * sizeof(double), st));
  if (p.det) TD_CUDA(cudaMemsetAsync(x.ws + p.det_enc_shadow, 0, 4 * (size_t)B * 2 * sizeof(double), st));
  Tag tag("frontend");
  if (int e = prepare_weights(x)) return e;
  // encoder (+ pad_input folded into the indexing) and its GlobLN statistics
  EncArgs ea{};
  ea.wav = wav; ea.B = B; ea.T = T; ea.K = K; ea.S = S; ea.front_pad = K - S; ea.Tp = p.Tp;
  ea.nconv = c->enc_convs; ea.ch_per_conv = Nb / c->enc_convs; ea.Nb = Nb; ea.L0 = L0;
  for (int k = 0; k < c->enc_convs; ++k) { ea.w[k] = w->enc_w[k]; ea.ks[k] = (k + 1) * K; }
  ea.out = x.at(p.enc); ea.stats = x.at<double>(p.st_enc);
  ea.det = x.det_enc();
  if (int e = launch_encoder(ea, st)) return e;
  if (int e = det_finalize(ea.det, ea.stats, (size_t)B * 2, 8, st)) return e;
  const NormRef enc_norm = norm_ref(x, p.st_enc, 2, (double)L0 * Nb, w->ln_gamma, w->ln_beta);
  if (c->variant == TDANET_MULTRES) {
    if (int e = launch_affine(x.at(p.enc), enc_norm, x.at(p.x0), B, L0, Nb, st)) return e;
  } else {
    if (int e = launch_bottleneck(x.at(p.enc), enc_norm, w->bottleneck_w, w->bottleneck_b, x.at(p.x0), B, L0, Nb, cc, st)) return e;
  }
  if (p.drop_elem || p.drop_item) {
    // every keep-mask of every iteration, one launch (dropout.cu)
    MaskRegion reg[5];
    int n = 0;
    const uint32_t te = (uint32_t)((double)c->dropout * 4294967296.0), tp = (uint32_t)((double)c->drop_path * 4294967296.0);
    if (p.drop_elem) {
      reg[n++] = MaskRegion{p.m_att, p.n_att, te, 0};
      reg[n++] = MaskRegion{p.m_ao, (size_t)B * p.Lb * c->in_channels, te, 1};
      reg[n++] = MaskRegion{p.m_f1, (size_t)B * p.Lb * 2 * c->in_channels, te, 2};
      reg[n++] = MaskRegion{p.m_f2, (size_t)B * p.Lb * c->in_channels, te, 3};
    }
    if (p.drop_item) reg[n++] = MaskRegion{p.m_dp, (size_t)2 * B, tp, 4};
    Tag tdm("dropout_masks");
    if (int e = launch_dropout_masks(x.ws, p.blk_stride, c->num_blocks, reg, n, rng_state, st)) return e;
  }
  // Recurrent: num_blocks iterations of one shared UConvBlock
  for (int blk = 0; blk < c->num_blocks; ++blk) {
    if (train) {
      // block `blk` reads its own arena's `bin` (block 0: x0) and writes the next block's `bin`
      Ctx xb = x;
      xb.blk = blk;
      const bool last = blk == c->num_blocks - 1;
      const float* in = blk == 0 ? x.at(p.x0) : xb.at(p.bin);
      if (int e = uconv_block(xb, in, last ? nullptr : x.at_blk(p.bin, blk + 1), last)) return e;
      continue;
    }
    const float* in = blk == 0 ? x.at(p.x0) : x.at(p.u[(blk - 1) & 1]);
    if (int e = uconv_block(x, in, x.at(p.u[blk & 1]), blk == c->num_blocks - 1)) return e;
  }
  // mask_net (PReLU -> 1x1) -> ReLU -> * encoder output, then decoder + crop
  GemmArgs g{};
  g.A = train ? x.at_blk(p.y, c->num_blocks - 1) : x.at(p.u[(c->num_blocks - 1) & 1]);
  g.W = w->mask_w; g.bias = w->mask_b; g.D = x.at(p.masked);
  g.B = B; g.L = L0; g.N = c->num_sources * Nb; g.K = cc; g.epi = EPI_MASK; g.a_slope = w->mask_prelu;
  g.enc = x.at(p.enc); g.Nb = Nb;
  Tag tb("backend");
  if (train) {
    // keep the mask logits: masked = relu(m) * enc is applied by its own launch
    g.D = x.at(p.mlogit); g.epi = EPI_BIAS;
    if (int e = launch_gemm_simt(g, st)) return e;
    if (int e = launch_mask_apply(x.at(p.mlogit), x.at(p.enc), x.at(p.masked), B * L0, c->num_sources, Nb, st)) return e;
  } else if (c->gemm_mode == TDANET_GEMM_TF32 && mask_conv_mma_applies(cc, g.N)) {
    if (int e = launch_mask_conv_mma(g.A, g.W, g.bias, g.a_slope, g.enc, g.D, B * L0, cc, g.N, Nb, st)) return e;
  } else if (int e = launch_gemm_simt(g, st)) return e;
  return launch_decoder(x.at(p.masked), w->dec_w, est, B, L0, Nb, c->num_sources, K, S, T, st);
}

}  // namespace td

// ============================================================================= C ABI
using namespace td;

extern "C" {

int tdanet_abi_version(void) { return TDANET_ABI_VERSION; }

int tdanet_abi_sizes(size_t* config_bytes, size_t* weights_bytes) {
  if (config_bytes) *config_bytes = sizeof(tdanet_config_t);
  if (weights_bytes) *weights_bytes = sizeof(tdanet_weights_t);
  return 0;
}

const char* tdanet_last_error(void) { return g_err; }

uint64_t tdanet_launch_count(void) { return g_launches.load(); }

int tdanet_set_deterministic(int on) {
  det_mode() = on != 0;
  return 0;
}

int tdanet_get_deterministic(void) { return det_mode(); }

int tdanet_device_supported(int dev) {
  cudaDeviceProp prop;
  TD_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(TDANET_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
  return 0;
}

int tdanet_profile_enable(int on) {
  g_profile = on != 0;
  return 0;
}

// JSON: [{"kernel": name, "launches": n, "ms": total}, ...] for everything recorded since the last
// dump; synchronises the device.  Returns the number of bytes needed (excluding the terminator).
int tdanet_profile_dump(char* out, size_t cap) {
  TD_CUDA(cudaDeviceSynchronize());
  struct Agg { std::string name; int n; double ms; };
  std::vector<Agg> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (r.ev[1] && cudaEventElapsedTime(&ms, r.ev[0], r.ev[1]) == cudaSuccess) {
      bool found = false;
      for (auto& a : agg)
        if (a.name == r.name) { a.n++; a.ms += ms; found = true; break; }
      if (!found) agg.push_back({r.name, 1, ms});
    }
    g_ev_pool.push_back(r.ev[0]);
    if (r.ev[1]) g_ev_pool.push_back(r.ev[1]);
  }
  g_prof.clear();
  std::string s = "[";
  char buf[512];
  for (size_t i = 0; i < agg.size(); ++i) {
    std::string nm;
    for (char ch : agg[i].name) if (ch != '"' && ch != '\\') nm += ch;
    snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"launches\": %d, \"ms\": %.6f}", i ? ", " : "", nm.c_str(), agg[i].n, agg[i].ms);
    s += buf;
  }
  s += "]";
  if (out && cap) {
    snprintf(out, cap, "%s", s.c_str());
  }
  return (int)s.size();
}

int tdanet_forward(const tdanet_config_t* cfg, const tdanet_weights_t* w, const float* wav, int batch,
                   int n_samples, float* est, void* workspace, size_t workspace_bytes, tdanet_stream_t stream) {
  return forward(cfg, w, wav, batch, n_samples, est, workspace, workspace_bytes, (cudaStream_t)stream, false);
}

int tdanet_forward_train(const tdanet_config_t* cfg, const tdanet_weights_t* w, const float* wav, int batch,
                         int n_samples, float* est, void* workspace, size_t workspace_bytes, tdanet_stream_t stream) {
  return forward(cfg, w, wav, batch, n_samples, est, workspace, workspace_bytes, (cudaStream_t)stream, true);
}

int tdanet_forward_train_rng(const tdanet_config_t* cfg, const tdanet_weights_t* w, const float* wav, int batch,
                             int n_samples, float* est, void* workspace, size_t workspace_bytes, uint64_t* rng_state,
                             tdanet_stream_t stream) {
  return forward(cfg, w, wav, batch, n_samples, est, workspace, workspace_bytes, (cudaStream_t)stream, true, rng_state);
}

size_t tdanet_gemm_workspace_bytes(int N, int K) { return ((size_t)N * K * sizeof(float) + 255) / 256 * 256; }

int tdanet_gemm(int gemm_mode, const float* A, const float* W, const float* bias, float* D, int batch,
                int rows_per_item, int N, int K, double* stats, void* workspace, size_t workspace_bytes,
                tdanet_stream_t stream) {
  TD_REQUIRE(A && W && D, "NULL argument");
  GemmArgs g{};
  g.A = A; g.W = W; g.bias = bias; g.D = D; g.B = batch; g.L = rows_per_item; g.N = N; g.K = K;
  g.stats = stats; g.epi = EPI_BIAS;
  cudaStream_t st = (cudaStream_t)stream;
  if (gemm_mode == TDANET_GEMM_FP32) return launch_gemm_simt(g, st);
  TD_REQUIRE(gemm_mode == TDANET_GEMM_TF32 || gemm_mode == TDANET_GEMM_TF32X3, "gemm_mode %d", gemm_mode);
  if (workspace_bytes < tdanet_gemm_workspace_bytes(N, K))
    return fail(TDANET_ENOSPACE, "gemm workspace has %zu bytes, need %zu", workspace_bytes, tdanet_gemm_workspace_bytes(N, K));
  if (int e = launch_tf32_prepare(W, (float*)workspace, (size_t)N * K, gemm_mode, st)) return e;
  g.W_aux = (const float*)workspace;
  return launch_gemm_tc(g, gemm_mode, st);
}

int tdanet_css_stitch(const float* est, int n_streams, int n_chunks, int seg_len, int overlap, int out_len,
                      int32_t* swap, float* out, tdanet_stream_t stream) {
  TD_REQUIRE(est && swap && (out || out_len == 0), "NULL argument");
  return launch_css_stitch(est, n_streams, n_chunks, seg_len, overlap, out_len, swap, out, (cudaStream_t)stream);
}

size_t tdanet_pit_loss_scratch_bytes(int batch, int n_src) {
  (void)n_src;
  return pit_scratch_floats(batch) * sizeof(float);
}

int tdanet_pit_loss(const float* est, const float* tgt, int batch, int n_src, int n_samples, int sdr_type,
                    int threshold_byloss, float* loss, float* pw, int32_t* perm, float* grad_est,
                    void* scratch, size_t scratch_bytes, tdanet_stream_t stream) {
  TD_REQUIRE(est && tgt && loss && scratch, "NULL argument");
  if (scratch_bytes < tdanet_pit_loss_scratch_bytes(batch, n_src))
    return fail(TDANET_ENOSPACE, "pit_loss scratch has %zu bytes, need %zu", scratch_bytes, tdanet_pit_loss_scratch_bytes(batch, n_src));
  return launch_pit_loss(est, tgt, batch, n_src, n_samples, sdr_type, threshold_byloss, loss, pw, perm,
                         grad_est, scratch, (cudaStream_t)stream);
}

}  // extern "C"
