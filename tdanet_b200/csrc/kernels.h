// Internal launcher interface between the engine (engine.cu) and the kernel files.
// All tensors are fp32 device pointers, activations channels-last [B, L, C].
#pragma once
#include "common.cuh"
#include <mutex>

namespace td {
// The side streams / event rings of forward (engine.cu) and backward (backward.cu) are per-device library state; a
// host thread holds this lock while it enqueues a forward or backward on that device, so two threads cannot
// interleave their fork / join events (or pull each other's side-stream launches into a stream capture).
inline std::mutex& device_enqueue_mutex(int dev) {
  static std::mutex mu[17];
  return mu[(dev >= 0 && dev < 16) ? dev : 16];
}

// ------------------------------------------------------------------ sources (normalise-on-load)
// How a consumer kernel reads one row of a [B, L, C] activation.
enum SrcKind {
  SRC_PLAIN = 0,         // x
  SRC_AFFINE = 1,        // GlobLN(x) = x*scale + shift        (norm)
  SRC_AFFINE_PRELU = 2,  // prelu(GlobLN(x), *slope)           (norm)
  SRC_INJECT_GATE = 3,   // (al*x+bl) * sigmoid(aa*g[j]+ba) + (ae*g[j]+be), j = nearest(t)  coef [B,6,C]
  SRC_INJECT_ADD = 4     // GlobLN(x) + g[j]                                                (norm)
};

struct SrcDesc {
  const float* x;      // [B, L, C]
  int L;
  NormRef norm;        // AFFINE / AFFINE_PRELU / INJECT_ADD
  const float* coef;   // INJECT_GATE: [B, 6, C] from launch_coef_inject_gate
  const float* slope;  // PReLU slope (1 element) or nullptr
  const float* g;      // injected global feature [B, Lg, C] or nullptr
  int Lg;
  float gscale;        // fl32(Lg / L)
  int bf16;            // backward kernels only: x is stored as bf16 (the forward kernels know it at compile time)
};

// ------------------------------------------------------------------ dwconv.cu
// Depthwise k=5 pad=2 conv over time (stride 1 or 2), NW weight sets sharing one input.
//   w[i]: [C,1,5]  bias[i]: [C] or null
//   out : [B, Lout, C] (only NW==1) or null;  relu applied to the stored value if `relu`
//   stats: [B, NW, 2] per-item sum / sum of squares (double) of the pre-relu conv output, or null
//   chstats: [B, NW, 2, C] the same per channel (float), or null
struct DwArgs {
  SrcDesc src;
  int kind;  // SrcKind
  int B, C, Lout, stride, nw;
  const float* w[2];
  const float* bias[2];
  float* out;
  double* stats;
  float* chstats;
  float* pool_out;  // optional [B, Lb, C]: adaptive-average-pooled raw output (nw == 1, out != null)
  int Lb;
  int act_bf16;     // src.x and out are stored as bf16 (large-activation storage mode), else fp32
  int rev;          // walk the batch items last to first (L2 reuse of what the producer wrote last; engine.cu)
  int relu;
  int round_out;  // store TF32-rounded values (output only feeds a tensor-core GEMM)
  DetRef det;     // deterministic mode: where the sums of stats / chstats are accumulated exactly (common.cuh)
};
int launch_dw5(const DwArgs& a, cudaStream_t st);

// Generic depthwise conv (any odd k, any stride), plain/affine source, writes out, no stats.
// wT (optional): transposed copy [ks][C] of w [C][ks] (launch_weight_transpose) - selects the input-stationary kernel.
int launch_dw_generic(const SrcDesc& src, int kind, int B, int C, int Lout, int ks, int stride,
                      const float* w, const float* wT, const float* bias, float* out, int round_out, int act_bf16,
                      cudaStream_t st);
// wT[k, c] = w[c, k]   (coef.cu)
int launch_weight_transpose(const float* w, float* wT, int C, int ks, cudaStream_t st);

// out[b, t, :] = injected operand (SRC_INJECT_GATE / SRC_INJECT_ADD) written out, [B, src.L, C].
// stats != null: also the GlobLN statistics [B, 2, 2] (double, accumulated) of the two 5-tap depthwise convolutions
// wa / we of the tensor written (zero padding) - what launch_dw5 with nw = 2 and no output would compute from it.
int launch_inject_materialize(const SrcDesc& src, int kind, int B, int C, float* out, int act_bf16, cudaStream_t st,
                              const float* wa = nullptr, const float* we = nullptr, double* stats = nullptr);
// two tensors in one launch (the statistics are those of tensor b)
int launch_inject_materialize2(const SrcDesc& sa, float* out_a, const SrcDesc& sb, float* out_b, int kind, int B, int C,
                               int act_bf16, cudaStream_t st, const float* wa = nullptr, const float* we = nullptr,
                               double* stats = nullptr);

// LA combine (TDANet_best.py:277-292 with the three GlobLN folded into coef tables):
//   out[t] = (cL.s*dw_l(xl)[t] + cL.h) * sigmoid(cA.s*dw_a(xg)[j] + cA.h) + (cE.s*dw_e(xg)[j] + cE.h)
//   j = nearest(t; Lg -> Ll)
struct LaArgs {
  SrcDesc loc, glo;
  int lkind, gkind;
  int B, C;
  const float *wl, *wa, *we;  // [C,1,5]
  NormRef nL, nA, nE;         // GlobLN of local_embedding / global_act / global_embedding outputs
  float* out;                 // [B, Ll, C]
  float scale;                // fl32(Lg / Ll)
  int round_out;
  int act_bf16;               // loc.x, glo.x and out are stored as bf16, else fp32
};
int launch_la_combine(const LaArgs& a, cudaStream_t st);
// statistics of local_embedding(x_fused[i]) for every top-down step in one launch (nw = 1, stats only)
int launch_la_local_stats(const DwArgs* steps, int n, cudaStream_t st);

// ------------------------------------------------------------------ coef.cu
// BEST loc_glo_fus (k=1 LA) closed form -> SRC_INJECT_GATE tables [B,6,C], every scale in one launch
//   spp_stats[k]: per-channel stats [B,2,C] of the raw spp_dw[k] output (L[k] rows), spp[k]: its GlobLN
//   g_stats:      per-channel stats [B,2,C] of global_f (Lg rows)
struct InjectCoefArgs {
  int n;
  const float* spp_stats[TDANET_MAX_DEPTH];
  int L[TDANET_MAX_DEPTH];
  tdanet_convnorm_t spp[TDANET_MAX_DEPTH];
  tdanet_la_t la[TDANET_MAX_DEPTH];
  float* coef[TDANET_MAX_DEPTH];
  double* conv_stats[TDANET_MAX_DEPTH];  // optional [B,3,2]: sum / sum of squares of the local, act, embedding conv outputs
  const float* g_stats;
  int Lg;
};
int launch_coef_inject_gate(const InjectCoefArgs& a, int B, int C, cudaStream_t st);

// ------------------------------------------------------------------ bottom.cu
// sum_k adaptive_avg_pool(affine_k(x_k)) -> out [B, Lb, C]
struct PoolArgs {
  int n;
  const float* x[TDANET_MAX_DEPTH];
  NormRef norm[TDANET_MAX_DEPTH];
  int L[TDANET_MAX_DEPTH];
  int B, C, Lb;
  float* out;
  // optional fused attn_in_norm + positional encoding (GA, TDANet_best.py:254-256): ln_out = LN_C(out)*ln_w + ln_b + pe[t]
  // (honoured when a row fits one CTA, C <= 1024; launch_affine_sum_fuses_ln tells)
  const float* ln_w = nullptr;
  const float* ln_b = nullptr;
  const float* pe = nullptr;
  float* ln_out = nullptr;
  int ln_round = 0;
};
bool launch_affine_sum_fuses_ln(const PoolArgs& a);
int launch_pool_sum(const PoolArgs& a, cudaStream_t st);
// fork: out = sum_k affine_k(x_k), all [B, Lb, C]
int launch_affine_sum(const PoolArgs& a, cudaStream_t st);
// y = LayerNorm_C(x)*w + b + pe[t]      x,y [B, L, C]
int launch_ln_pe(const float* x, const float* w, const float* b, const float* pe, float* y, int B,
                 int L, int C, int round_out, cudaStream_t st);
// multi-head attention core on packed qkv [B*L, 3C]; ctx [B*L, C]
//   time_axis=0: sequence = the `group` batch items sharing a time index; 1: sequence = time
//   amask (training, may be null): keep-mask bytes [problem*head, query, key] of the dropout on the attention
//   weights; kept weights are scaled by inv_keep
int launch_attention(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                     int time_axis, int round_out, const uint8_t* amask, float inv_keep, cudaStream_t st);
// Training-mode multipliers of a residual branch: element keep-mask (nn.Dropout) and per-item keep-mask (DropPath),
// both optional byte tensors; kept entries are scaled by the matching inv_keep.
struct DropRef {
  const uint8_t* mask;       // [B, L, C] or null
  float inv_keep;
  const uint8_t* item_mask;  // [B] or null
  float item_inv_keep;
};
// y = resid + f_b * (LayerNorm_C(post)*w + b) ; post = 2*a (mode 1), xin + a (mode 0) or a (mode 2: the dropout of
// `a + dropout(a)` was already folded into a); f_b: DropPath multiplier of drop.item_mask (1 if null)
int launch_ln_residual(const float* a, const float* xin, const float* resid, const float* w,
                       const float* b, float* y, int mode, const DropRef& drop, int B, int L, int C, cudaStream_t st);
// y = resid + f_b * m * GlobLN(x) (m, f_b: multipliers of `drop`); optional per-channel stats of y -> [B,2,C]
int launch_affine_residual(const float* x, const NormRef& norm, const float* resid, float* y,
                           float* chstats, const DropRef& drop, int B, int L, int C, cudaStream_t st,
                           const DetRef& det = DetRef{});
// ------------------------------------------------------------------ dropout.cu (training only)
struct MaskRegion {
  size_t off;      // byte offset inside one iteration's block arena
  size_t n;        // mask bytes (padded to a multiple of 16 by the generator's index space)
  uint32_t thresh; // keep  <=>  u32 >= thresh
  uint32_t site;   // 0 m_att, 1 m_ao, 2 m_f1, 3 m_f2, 4 m_dp: part of the Philox counter
};
// draws the keep-masks of every region for iterations [0, n_blk): base + blk*blk_stride + region.off; then
// rng_state[1] += 1 (device uint64[2] = {seed, offset})
int launch_dropout_masks(char* base, size_t blk_stride, int n_blk, const MaskRegion* regions, int n_regions,
                         uint64_t* rng_state, cudaStream_t st);
// out[i] = in[i] * (k0 + k1*mask[i]) * (item_mask ? item_mask[i / per_item]*item_scale : 1), optionally rounded to
// TF32; mask / item_mask may be null (mask null: factor k0 + k1); in == out allowed.  (backward.cu)
int launch_mask_scale(const float* in, float* out, size_t n, const uint8_t* mask, float k0, float k1,
                      const uint8_t* item_mask, float item_scale, size_t per_item, int round_out, cudaStream_t st);
// y = GlobLN(x)
int launch_affine(const float* x, const NormRef& norm, float* y, int B, int L, int C, cudaStream_t st);

// ------------------------------------------------------------------ frontend.cu
struct EncArgs {
  const float* wav;  // [B, T]
  int B, T;
  int K, S, front_pad, Tp;  // Tp: length after pad_input
  int nconv;                // 1 or MULTRES kernels
  const float* w[TDANET_MAX_ENC];
  int ks[TDANET_MAX_ENC];
  int ch_per_conv, Nb, L0;
  float* out;     // [B, L0, Nb]
  double* stats;  // [B,2]
  DetRef det;     // deterministic mode (common.cuh)
};
int launch_encoder(const EncArgs& a, cudaStream_t st);
// x0[b,t,:] = Wb . GlobLN(enc) + bb     enc [B,L0,Nb], out [B,L0,c]
int launch_bottleneck(const float* enc, const NormRef& norm, const float* w, const float* bias,
                      float* out, int B, int L0, int Nb, int c, cudaStream_t st);
// mask_net + ReLU mask + encoder product on the tensor cores (TF32 modes, inference): masked [R, N] from y [R, K]
bool mask_conv_mma_applies(int K, int N);
int launch_mask_conv_mma(const float* y, const float* w, const float* bias, const float* slope, const float* enc,
                         float* out, int R, int K, int N, int Nb, cudaStream_t st);
// decoder ConvTranspose1d + crop: masked [B, L0, n_src*Nb] -> est [B, n_src, T]
int launch_decoder(const float* masked, const float* w, float* est, int B, int L0, int Nb,
                   int n_src, int K, int S, int T, cudaStream_t st);

// training forward: out = prelu(cw*(mix + y) + cb, *slope)   (concat_block, TDANet_best.py:388-398), [rows, c]
int launch_concat(const float* y, const float* mix, const float* cw, const float* cb, const float* slope,
                  float* out, int rows, int c, cudaStream_t st);
// training forward: masked[r, s*Nb+n] = relu(m[r, s*Nb+n]) * enc[r, n]   (TDANet_best.py:507-509)
int launch_mask_apply(const float* m, const float* enc, float* masked, int rows, int n_src, int Nb, cudaStream_t st);

// ------------------------------------------------------------------ gemm_simt.cu / gemm_tc.cu
enum GemmEpi {
  EPI_BIAS = 0,       // D = acc + bias
  EPI_RESIDUAL = 1,   // y = acc + bias + resid ; D = last ? y : prelu(cw*(mix+y)+cb, *cslope)
  EPI_MASK = 2        // D = relu(acc + bias) * enc[b, r, n mod Nb]
};
struct GemmArgs {
  const float* A;  // [B, L, K]
  const float* W;  // [N, K]
  const float* bias;
  float* D;        // [B, L, N]
  int B, L, N, K;
  double* stats;   // [B,2] or null (sum, sumsq of D over valid rows)
  int epi;
  // A-operand transform: prelu(A, *a_slope) if non-null (mask_net.0)
  const float* a_slope;
  // EPI_RESIDUAL
  const float *resid, *mix, *cw, *cb, *cslope;
  int last;
  // EPI_MASK
  const float* enc;
  int Nb;
  // tensor-core path: weight split prepared by prepare_tf32_weights (lo part / rounded copy / bf16 copy)
  const float* W_aux;
  // bf16 activation storage: D is written as bf16 (proj_1x1) / A is read as bf16 (res_conv, kind::f16 MMA
  // against a bf16 copy of W in W_aux)
  int d_bf16, a_bf16;
  int rev;  // tensor-core path: tiles of the last batch item first (L2 reuse of what the producer wrote last)
  int narrow;  // tensor-core path: nothing else competes for the SMs (inference forward): narrow tiles for small launches
  DetRef det;  // deterministic mode: where the sums of `stats` are accumulated exactly (common.cuh)
};
int launch_gemm_simt(const GemmArgs& a, cudaStream_t st);
int launch_gemm_tc(const GemmArgs& a, int mode, cudaStream_t st);
// aux[i] = mode==TF32 ? rna_tf32(w[i]) : w[i] - trunc_tf32(w[i])
int launch_tf32_prepare(const float* w, float* aux, size_t n, int mode, cudaStream_t st);
// aux (as bf16[n]) = round-to-nearest-even bf16 copy of w
int launch_bf16_prepare(const float* w, float* aux, size_t n, cudaStream_t st);

// ------------------------------------------------------------------ wgrad_mma.cu
// dW[N, K] += G[R, N]^T A[R, K] (TF32 mma.sync, fp32 accumulate), db[N] += column sums of G (db may be null)
int launch_wgrad_mma(const float* G, const float* A, float* dW, float* db, int R, int N, int K, cudaStream_t st);

// ------------------------------------------------------------------ css.cu
int launch_css_stitch(const float* est, int n_streams, int n_chunks, int seg_len, int overlap, int out_len,
                      int32_t* swap, float* out, cudaStream_t st);

// ------------------------------------------------------------------ loss.cu
int launch_pit_loss(const float* est, const float* tgt, int B, int n_src, int T, int sdr_type,
                    int threshold, float* loss, float* pw, int32_t* perm, float* grad,
                    void* scratch, cudaStream_t st);

}  // namespace td
