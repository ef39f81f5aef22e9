// Backward pass of TDANetBest (TDANet_best.py:342-399, 482-521) over the activations kept by
// tdanet_forward_train: host-side launch sequence + launchers of bwd_kernels.cuh.
//
// Replaces what autograd does for the reference's training_step (system/audio_litmodule.py:83-124): given
// d loss / d est it accumulates d loss / d theta into caller-provided buffers laid out like tdanet_weights_t.
// The 16 UConvBlock iterations share one set of weights, so every iteration adds into the same buffers.
//
// Also compiled by g++ with -DTD_EMU (see emu.h) so that the whole sweep can be checked on a CPU.
#include "plan.h"
#include "bwd_kernels.cuh"

namespace td {

#ifdef TD_EMU
}  // namespace td
namespace emu {
thread_local uint3 tid, bid;
thread_local BlockState* bs = nullptr;
dim3 bdim, gdim;
bool coop_reductions = getenv("TD_EMU_COOP") != nullptr;
}  // namespace emu
namespace td {
// ----------------------------------------------------------------------------- emulation build: host pieces
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
bool g_profile = false;
thread_local const char* g_tag = nullptr;
void profile_mark(const char*, cudaStream_t, bool) {}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
// reference semantics of the GEMM launchers (EPI_BIAS, EPI_RESIDUAL with last = 1), enough for the backward pass
static int emu_gemm(const GemmArgs& a) {
  TD_REQUIRE(a.epi == EPI_BIAS || (a.epi == EPI_RESIDUAL && a.last), "emu_gemm: epilogue %d", a.epi);
  for (int b = 0; b < a.B; ++b)
    for (int r = 0; r < a.L; ++r) {
      const size_t row = (size_t)b * a.L + r;
      for (int n = 0; n < a.N; ++n) {
        double acc = a.bias ? a.bias[n] : 0.0;
        for (int k = 0; k < a.K; ++k) acc += (double)a.A[row * a.K + k] * a.W[(size_t)n * a.K + k];
        if (a.epi == EPI_RESIDUAL) acc += a.resid[row * a.N + n];
        a.D[row * a.N + n] = (float)acc;
      }
    }
  return 0;
}
int launch_tf32_prepare(const float* w, float* aux, size_t n, int, cudaStream_t) {
  memcpy(aux, w, n * sizeof(float));
  return 0;
}
#endif

static int bgemm(const Ctx& x, GemmArgs& g, size_t aux_off) {
#ifdef TD_EMU
  (void)x; (void)aux_off;
  return emu_gemm(g);
#else
  if (x.c->gemm_mode == TDANET_GEMM_FP32) return launch_gemm_simt(g, x.st);
  g.W_aux = x.at(aux_off);
  return launch_gemm_tc(g, x.c->gemm_mode, x.st);
#endif
}

// Side streams of the backward sweep (lazily created per device, never destroyed): `w` runs the weight-gradient
// GEMMs, which nothing in the sweep waits for before the end of a block; `l` runs the local-branch depthwise
// backward of every top-down step while the main stream continues down the global-branch chain; `f[0..1]` run the
// loc_glo_fus chains (four launches per scale) of alternating scales, each with its own temporaries and its own
// accumulator of the global feature's gradient.  (Round 1 ran local branch + loc_glo_fus of every scale on ONE side
// stream: 306 us of serial launches per block against 229 us on the main stream at B = 8 - the join waited.)
// Dependencies are cudaEventRecord / cudaStreamWaitEvent pairs, which stream capture turns into graph edges, so a
// captured training step keeps the concurrency.
struct SideStreams {
  cudaStream_t w = nullptr, l = nullptr, f[2] = {nullptr, nullptr}, g = nullptr;
  std::vector<cudaEvent_t> events;
  size_t next = 0;
  int init() {
    if (w) return 0;
    TD_CUDA(cudaStreamCreateWithFlags(&w, cudaStreamNonBlocking));
    TD_CUDA(cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking));
    TD_CUDA(cudaStreamCreateWithFlags(&f[0], cudaStreamNonBlocking));
    TD_CUDA(cudaStreamCreateWithFlags(&f[1], cudaStreamNonBlocking));
    TD_CUDA(cudaStreamCreateWithFlags(&g, cudaStreamNonBlocking));
    events.resize(512);
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
static SideStreams g_side[16];

struct RepEntry { float* dst; size_t off; int n; };
struct BCtx : Ctx {
  const tdanet_weights_t* g;  // gradient buffers, same layout as the weights
  std::vector<RepEntry>* reps;  // replicated accumulators handed out so far (shared by the per-block copies)
  SideStreams* side;
  cudaEvent_t* w_pending;       // the previous block's weight-gradient GEMMs (stream `w`), not yet joined
  cudaStream_t main_st;         // the caller's stream (x.st is the stream launches currently go to)
  int tset = 0;                 // LA temporary set in use
  BCtx on(cudaStream_t s, int set) const { BCtx y = *this; y.st = s; y.tset = set; return y; }
  float* gp(const float* p) const { return const_cast<float*>(p); }
  // replica 0 of the accumulator that stands in for the depthwise gradient buffer `dst` (n floats) during the sweep
  float* rep_of(const float* dst, int n) const {
    if (!dst) return nullptr;
    for (auto& e : *reps)
      if (e.dst == dst) return at(p->rep_arena) + e.off;
    size_t off = reps->empty() ? 0 : reps->back().off + (size_t)(reps->back().n + 3) / 4 * 4;
    if (off + n > p->rep_floats || reps->size() >= 64) return nullptr;
    reps->push_back({const_cast<float*>(dst), off, n});
    return at(p->rep_arena) + off;
  }
};

// ----------------------------------------------------------------------------- launchers
#ifdef TD_BWD_MINB
constexpr int kRowGridMaxThreads = 128;   // the experiment's launch bounds are for 128-thread CTAs
#else
constexpr int kRowGridMaxThreads = 256;
#endif
static inline void row_grid(int L, int C4, int B, int rows, dim3& grid, int& threads) {
  threads = C4 > kRowGridMaxThreads ? kRowGridMaxThreads : (C4 < 32 ? 32 : C4);
  grid = dim3(cdiv(L, rows), cdiv(C4, threads), B);
}
// rows per thread (a multiple of the 4-row tile): as many as `max_rows`, but keep at least two CTAs per SM
static inline int pick_rows(int L, int C4, int B, int max_rows) {
  const int threads = C4 > kRowGridMaxThreads ? kRowGridMaxThreads : (C4 < 32 ? 32 : C4);
  int rows = max_rows;
  while (rows > 4 && (long)cdiv(L, rows) * cdiv(C4, threads) * B < 2 * 148) rows >>= 1;
  return rows;
}

// CTAs of `kernel` that can be resident on the device at once (occupancy x SM count); cached per (kernel, block size)
static long resident_ctas(const void* kernel, int threads) {
#ifdef TD_EMU
  (void)kernel; (void)threads;
  return 1L << 30;
#else
  static std::vector<std::pair<std::pair<const void*, int>, long>> cache;
  for (auto& e : cache)
    if (e.first.first == kernel && e.first.second == threads) return e.second;
  int per_sm = 0, dev = 0, sms = 148;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cache.push_back({{kernel, threads}, (long)per_sm * sms});
  return cache.back().second;
#endif
}
// Rows per thread for the register-heavy streaming backward kernels (2-3 CTAs of 128 threads per SM): `rows` unless
// the grid would be between one and two waves of resident CTAs - at the training batch of 8 it is 504 CTAs for 296 or
// 444 slots, and the second, mostly empty wave costs as much as the first.  Then the rows are stretched (to a
// multiple of `tile`) so that one wave covers the tensor.
static int fit_wave(int rows, int tile, int L, int ctiles, int B, long slots) {
  const long ctas = (long)cdiv(L, rows) * ctiles * B;
  if (ctas <= slots || ctas >= 2 * slots) return rows;
  const long chunks = slots / ((long)ctiles * B);
  if (chunks < 1) return rows;
  return cdiv(cdiv(L, (int)chunks), tile) * tile;
}

static int launch_gln_bwd_stats_t(int x_bf16, const float* dy, const float* x, const NormRef& norm, float* dgamma, float* dbeta,
                                double* S, int B, int L, int C, cudaStream_t st) {
  dim3 grid;
  int threads;
  const int rows = pick_rows(L, C % 4 == 0 ? C / 4 : C, B, 32);
  if (C % 4 == 0) {
    row_grid(L, C / 4, B, rows, grid, threads);
    if (x_bf16) TD_LAUNCH_RED((gln_bwd_stats_kernel<4, true>), grid, threads, 0, st, dy, x, norm, dgamma, dbeta, S, L, C, rows, x_bf16);
    else TD_LAUNCH_RED((gln_bwd_stats_kernel<4>), grid, threads, 0, st, dy, x, norm, dgamma, dbeta, S, L, C, rows, x_bf16);
  } else {
    row_grid(L, C, B, rows, grid, threads);
    TD_LAUNCH_RED((gln_bwd_stats_kernel<1>), grid, threads, 0, st, dy, x, norm, dgamma, dbeta, S, L, C, rows, x_bf16);
  }
  return 0;
}

static int launch_gln_bwd_stats(const float* dy, const float* x, const NormRef& norm, float* dgamma, float* dbeta,
                                double* S, int B, int L, int C, cudaStream_t st) {
  return launch_gln_bwd_stats_t(0, dy, x, norm, dgamma, dbeta, S, B, L, C, st);
}

static int launch_gln_bwd_apply(const GradSrc& g, float* out, int accumulate, int B, int L, int C, cudaStream_t st) {
  dim3 grid;
  int threads;
  const int rows = pick_rows(L, C % 4 == 0 ? C / 4 : C, B, 16);
  if (C % 4 == 0) {
    row_grid(L, C / 4, B, rows, grid, threads);
    if (g.x_bf16) TD_LAUNCH((gln_bwd_apply_kernel<4, true>), grid, threads, 0, st, g, out, accumulate, L, C, rows);
    else TD_LAUNCH((gln_bwd_apply_kernel<4>), grid, threads, 0, st, g, out, accumulate, L, C, rows);
  } else {
    row_grid(L, C, B, rows, grid, threads);
    TD_LAUNCH((gln_bwd_apply_kernel<1>), grid, threads, 0, st, g, out, accumulate, L, C, rows);
  }
  return 0;
}

static int launch_dw_bwd(const BCtx& x, DwBwdArgs& a, int ks, int nw) {
  cudaStream_t st = x.st;
  TD_REQUIRE(a.C % 4 == 0, "dw_bwd: C=%d", a.C);
  a.rep_stride = x.p->rep_floats;
  a.n_rep = TDANET_DW_REPLICAS;
  for (int g = 0; g < nw; ++g) {
    a.dw[g] = x.rep_of(a.dw[g], a.C * ks);
    TD_REQUIRE(a.dw[g] != nullptr, "dw_bwd: replica arena exhausted");
    if (a.db[g]) {
      a.db[g] = x.rep_of(a.db[g], a.C);
      TD_REQUIRE(a.db[g] != nullptr, "dw_bwd: replica arena exhausted");
    }
  }
  TD_REQUIRE(a.stride == 1 || a.stride == 2, "dw_bwd: stride %d", a.stride);
  TD_REQUIRE(!a.pool_g || (long)(a.Lin + 1) * (a.pool_Lb + 1) < (1L << 31), "dw_bwd: pooling %d -> %d overflows 32-bit bin arithmetic", a.Lin, a.pool_Lb);
  TD_REQUIRE(a.xkind == SRC_PLAIN || a.xkind == SRC_AFFINE || a.xkind == SRC_AFFINE_PRELU, "dw_bwd: source kind %d", a.xkind);
  TD_REQUIRE(a.xin.L == a.Lin, "dw_bwd: input length %d != %d", a.xin.L, a.Lin);
  const bool extra = a.xkind == SRC_AFFINE_PRELU || a.up_S || a.pool_g || a.dslope;
  const int key = ks * 100 + nw * 10 + a.stride;
  dim3 grid;
  int threads;
  a.rows_per_thread = pick_rows(a.Lout, a.C / 4, a.B, 32);
  row_grid(a.Lout, a.C / 4, a.B, a.rows_per_thread, grid, threads);
  // 2 channels per thread: only where 4 channels per thread leave most SMs without a single CTA (very small
  // launches).  Measured on B200 at the training shape (B = 8, 2016 warps per launch): V = 2 is SLOWER (bwd_spp_dw0
  // 0.81 -> 1.30 ms per step, bwd_la_dw 5.20 -> 5.35): the kernels are not occupancy-bound there.
  const long warps4 = (long)grid.x * grid.y * grid.z * (threads / 32);
  if (warps4 < 148L && a.C % 2 == 0 && !(a.xin.bf16 || a.g[0].x_bf16 || (nw == 2 && a.g[1].x_bf16))) {
    a.rows_per_thread = pick_rows(a.Lout, a.C / 2, a.B, 32);
    row_grid(a.Lout, a.C / 2, a.B, a.rows_per_thread, grid, threads);
    if (key == 511 && !extra) TD_LAUNCH_RED((dw_bwd_kernel<5, 1, 1, false, 2>), grid, threads, 0, st, a);
    else if (key == 511) TD_LAUNCH_RED((dw_bwd_kernel<5, 1, 1, true, 2>), grid, threads, 0, st, a);
    else if (key == 512) TD_LAUNCH_RED((dw_bwd_kernel<5, 1, 2, true, 2>), grid, threads, 0, st, a);
    else if (key == 521 && !extra) TD_LAUNCH_RED((dw_bwd_kernel<5, 2, 1, false, 2>), grid, threads, 0, st, a);
    else if (key == 111 && !extra) TD_LAUNCH_RED((dw_bwd_kernel<1, 1, 1, false, 2>), grid, threads, 0, st, a);
    else if (key == 121 && !extra) TD_LAUNCH_RED((dw_bwd_kernel<1, 2, 1, false, 2>), grid, threads, 0, st, a);
    else return fail(TDANET_EINVAL, "dw_bwd: ks=%d nw=%d", ks, nw);
    return 0;
  }
  void (*kfn)(DwBwdArgs) = nullptr;
  int tile = 4;  // output rows per tile of the variant (bwd_kernels.cuh: R)
  // act_dtype bf16: which operands are stored activations in bf16 is a compile-time property of the instantiation.
  // The combinations the sweep produces: conv input only (LA / loc_glo_fus branches: the gradient sources read fp32
  // temporaries) and both (the spp_dw chain); everything at the bottom scale is fp32.
  const bool xbf = a.xin.bf16 != 0;
  const bool gbf = a.g[0].x_bf16 != 0;
  TD_REQUIRE(nw == 1 || (a.g[1].x_bf16 != 0) == gbf, "dw_bwd: mixed storage of the gradient sources");
  if (xbf && !gbf) {
    if (key == 511 && !extra) kfn = dw_bwd_kernel<5, 1, 1, false, 4, false, true>;
    else if (key == 521 && !extra) { kfn = dw_bwd_kernel<5, 2, 1, false, 4, false, true>; tile = 2; }
    else if (key == 111 && !extra) kfn = dw_bwd_kernel<1, 1, 1, false, 4, false, true>;
    else if (key == 121 && !extra) kfn = dw_bwd_kernel<1, 2, 1, false, 4, false, true>;
    else return fail(TDANET_EINVAL, "dw_bwd (bf16 input): ks=%d nw=%d extra=%d", ks, nw, (int)extra);
  } else if (xbf && gbf) {
    if (key == 511) kfn = dw_bwd_kernel<5, 1, 1, true, 4, true, true>;
    else if (key == 512) { kfn = dw_bwd_kernel<5, 1, 2, true, 4, true, true>; tile = 2; }
    else return fail(TDANET_EINVAL, "dw_bwd (bf16 input and sources): ks=%d nw=%d", ks, nw);
  } else if (gbf) {
    return fail(TDANET_EINVAL, "dw_bwd: bf16 gradient sources with an fp32 input");
  }
  else if (key == 511 && !extra) kfn = dw_bwd_kernel<5, 1, 1, false>;
  else if (key == 511) kfn = dw_bwd_kernel<5, 1, 1, true>;
  else if (key == 512) { kfn = dw_bwd_kernel<5, 1, 2, true>; tile = 2; }
  else if (key == 521 && !extra) { kfn = dw_bwd_kernel<5, 2, 1, false>; tile = 2; }
  else if (key == 111 && !extra) kfn = dw_bwd_kernel<1, 1, 1, false>;
  else if (key == 121 && !extra) kfn = dw_bwd_kernel<1, 2, 1, false>;
  else return fail(TDANET_EINVAL, "dw_bwd: ks=%d nw=%d", ks, nw);
  a.rows_per_thread = fit_wave(a.rows_per_thread, tile, a.Lout, grid.y, a.B, resident_ctas((const void*)kfn, threads));
  row_grid(a.Lout, a.C / 4, a.B, a.rows_per_thread, grid, threads);
  TD_LAUNCH_RED(kfn, grid, threads, 0, st, a);
  return 0;
}

// pass G alone (raw_b / raw_e = the global-branch conv outputs: forward quantities, independent of any gradient)
static int launch_la_bwd_g(LaBwdArgs& a, int ks, cudaStream_t st) {
  dim3 grid;
  int threads;
  const int grows = pick_rows(a.Lg, a.C / 4, a.B, 16);
  row_grid(a.Lg, a.C / 4, a.B, grows, grid, threads);
  if (a.glo_bf16) {
    if (ks == 5) TD_LAUNCH((la_bwd_g_kernel<5, true>), grid, threads, 0, st, a, grows);
    else TD_LAUNCH((la_bwd_g_kernel<1, true>), grid, threads, 0, st, a, grows);
  } else if (ks == 5) TD_LAUNCH((la_bwd_g_kernel<5>), grid, threads, 0, st, a, grows);
  else TD_LAUNCH((la_bwd_g_kernel<1>), grid, threads, 0, st, a, grows);
  return 0;
}

static int launch_la_bwd_a(LaBwdArgs& a, int ks, cudaStream_t st, bool g_done = false) {
  TD_REQUIRE(a.C % 4 == 0, "la_bwd: C=%d", a.C);
  TD_REQUIRE(ks == 5 || ks == 1, "la_bwd: ks=%d", ks);
  dim3 grid;
  int threads;
  const int grows = pick_rows(a.Lg, a.C / 4, a.B, 16);  // global rows per thread of the G and F passes
  row_grid(a.Lg, a.C / 4, a.B, grows, grid, threads);
  if (g_done) {
    // raw_b / raw_e were computed ahead of the chain (uconv_block_backward: stream `g`)
  } else if (a.glo_bf16) {
    if (ks == 5) TD_LAUNCH((la_bwd_g_kernel<5, true>), grid, threads, 0, st, a, grows);
    else TD_LAUNCH((la_bwd_g_kernel<1, true>), grid, threads, 0, st, a, grows);
  } else if (ks == 5) TD_LAUNCH((la_bwd_g_kernel<5>), grid, threads, 0, st, a, grows);
  else TD_LAUNCH((la_bwd_g_kernel<1>), grid, threads, 0, st, a, grows);
  // L pass: chunks of centres covering about 32 local rows per thread (one wave of CTAs where that is close, fit_wave)
  int jc = (int)((double)pick_rows(a.loc.L, a.C / 4, a.B, 32) * a.Lg / a.loc.L + 0.5);
  a.jchunk = jc < 1 ? 1 : jc;
  void (*lfn)(LaBwdArgs) = a.loc.bf16 ? (ks == 5 ? la_bwd_l_kernel<5, true> : la_bwd_l_kernel<1, true>)
                                      : (ks == 5 ? la_bwd_l_kernel<5> : la_bwd_l_kernel<1>);
  a.jchunk = fit_wave(a.jchunk, 1, a.Lg, cdiv(a.C / 4, threads), a.B, resident_ctas((const void*)lfn, threads));
  dim3 lgrid;
  row_grid(a.Lg, a.C / 4, a.B, a.jchunk, lgrid, threads);
  TD_LAUNCH_RED(lfn, lgrid, threads, 0, st, a);
  return 0;
}

static int launch_pool_bwd(const float* g, float* dx, int accumulate, int B, int L, int Lb, int C, cudaStream_t st) {
  TD_REQUIRE((long)(L + 1) * (Lb + 1) < (1L << 31), "pool_bwd: pooling %d -> %d overflows 32-bit bin arithmetic", L, Lb);
  dim3 grid;
  int threads;
  const int rows = pick_rows(L, C / 4, B, 16);
  row_grid(L, C / 4, B, rows, grid, threads);
  TD_LAUNCH(pool_bwd_kernel, grid, threads, 0, st, g, dx, accumulate, L, Lb, C, rows);
  return 0;
}

// gradient through  y = resid + LN(k1*x)*w + b :  out = (add ? add : 0) + k1 * dLN(dy)
static int launch_ln_bwd(const float* xin, float k1, const float* w, const float* dy, float* rowstat, const float* add,
                         float* out, float* dw, float* db, int rows, int C, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "ln_bwd: C=%d", C);
  TD_LAUNCH_COOP(ln_bwd_rows_kernel, cdiv(rows, 8), 256, 0, st, xin, k1, w, dy, rowstat, rows, C);
  const int rpt = pick_rows(rows, C / 4, 1, 16);
  const int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  dim3 grid(cdiv(rows, rpt), cdiv(C / 4, threads));
  TD_LAUNCH(ln_bwd_apply_kernel, grid, threads, 0, st, xin, k1, w, dy, rowstat, add, k1, out, dw, db, rows, C, rpt);
  return 0;
}

template <int D>
static int launch_att_bwd_d(const float* qkv, const float* dctx, float* P, float* dS, float* dqkv, int B, int L, int C,
                            int n_head, int group, int time_axis, const uint8_t* amask, float inv_keep, cudaStream_t st) {
  const int n = time_axis ? L : group;
  const int nprob = time_axis ? B : (B / group) * L;
#ifdef TD_EMU
  // the warp kernel runs as OS threads with barriers under emulation (seconds per launch): only when asked for
  static const bool emu_warp = getenv("TD_EMU_WARP_ATT") != nullptr;
  if (n <= 16 && emu_warp) {
#else
  if (n <= 16) {
#endif
    const int warps = nprob * n_head;
    if (n <= 8) TD_LAUNCH_COOP((att_bwd_warp_kernel<D, 8>), cdiv(warps, 4), 128, 0, st, qkv, dctx, dqkv, L, C, n, n_head, group, time_axis, warps, amask, inv_keep);
    else TD_LAUNCH_COOP((att_bwd_warp_kernel<D, 16>), cdiv(warps, 4), 128, 0, st, qkv, dctx, dqkv, L, C, n, n_head, group, time_axis, warps, amask, inv_keep);
    return 0;
  }
  const int total = nprob * n_head * n;
  TD_LAUNCH((att_bwd_dq_kernel<D>), cdiv(total, 128), 128, 0, st, qkv, dctx, P, dS, dqkv, L, C, n, n_head, group, time_axis, total, amask, inv_keep);
  TD_LAUNCH((att_bwd_dkv_kernel<D>), cdiv(total, 128), 128, 0, st, qkv, dctx, P, dS, dqkv, L, C, n, n_head, group, time_axis, total);
  return 0;
}

static int launch_att_bwd(const float* qkv, const float* dctx, float* P, float* dS, float* dqkv, int B, int L, int C,
                          int n_head, int group, int time_axis, const uint8_t* amask, float inv_keep, cudaStream_t st) {
  switch (C / n_head) {
    case 64: return launch_att_bwd_d<64>(qkv, dctx, P, dS, dqkv, B, L, C, n_head, group, time_axis, amask, inv_keep, st);
    case 32: return launch_att_bwd_d<32>(qkv, dctx, P, dS, dqkv, B, L, C, n_head, group, time_axis, amask, inv_keep, st);
    case 16: return launch_att_bwd_d<16>(qkv, dctx, P, dS, dqkv, B, L, C, n_head, group, time_axis, amask, inv_keep, st);
    case 8: return launch_att_bwd_d<8>(qkv, dctx, P, dS, dqkv, B, L, C, n_head, group, time_axis, amask, inv_keep, st);
    case 4: return launch_att_bwd_d<4>(qkv, dctx, P, dS, dqkv, B, L, C, n_head, group, time_axis, amask, inv_keep, st);
  }
  return fail(TDANET_EUNSUPPORTED, "attention backward: head dim %d not in {4,8,16,32,64}", C / n_head);
}

// dW[N, K] += G[R, N]^T f(A[R, K]);  db[N] += column sums of G (db may be null)
int launch_mask_scale(const float* in, float* out, size_t n, const uint8_t* mask, float k0, float k1,
                      const uint8_t* item_mask, float item_scale, size_t per_item, int round_out, cudaStream_t st) {
  TD_REQUIRE(n % 4 == 0 && per_item > 0 && (item_mask == nullptr || per_item % 4 == 0), "mask_scale: n=%zu per_item=%zu", n, per_item);
  const size_t n4 = n / 4;
  TD_LAUNCH(mask_scale_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, in, out, n4, mask, k0, k1, item_mask, item_scale,
            per_item, round_out);
  return 0;
}

static int launch_wgrad(const float* G, const float* A, float* dW, float* db, int R, int N, int K,
                        const float* a_slope, cudaStream_t st, int gemm_mode = TDANET_GEMM_FP32) {
#ifndef TD_EMU
  // tensor-core GEMM modes: TF32 mma.sync kernel (shapes it covers); gemm_mode fp32 keeps the exact fp32 kernel
  if (gemm_mode != TDANET_GEMM_FP32 && !a_slope && N % 4 == 0 && K % 4 == 0)
    return launch_wgrad_mma(G, A, dW, db, R, N, K, st);
#else
  (void)gemm_mode;
#endif
  const int tiles = cdiv(N, WG_T) * cdiv(K, WG_T);
  int splits = cdiv(592, tiles);  // about four CTAs per SM
  int rps = cdiv(R, splits);
  rps = cdiv(rps, WG_R) * WG_R;
  splits = cdiv(R, rps);
  dim3 grid(cdiv(N, WG_T), cdiv(K, WG_T), splits);
  TD_LAUNCH_COOP(wgrad_kernel, grid, 256, 0, st, G, A, dW, R, N, K, a_slope, rps);
  if (db) {
    const int rpt = 64;
    dim3 g2(cdiv(N, 128), cdiv(R, rpt));
    TD_LAUNCH(colsum_kernel, g2, 128, 0, st, G, db, R, N, rpt);
  }
  return 0;
}

static int launch_small_dgrad(const float* G, const float* W, float* D, int R, int Kd, int N, const float* u,
                              const float* slope, float* dslope, cudaStream_t st) {
  const size_t n = (size_t)R * N;
  TD_LAUNCH_RED(small_dgrad_kernel, (unsigned)((n + 255) / 256), 256, 0, st, G, W, D, R, Kd, N, u, slope, dslope);
  return 0;
}

static int launch_transpose(const float* W, float* Wt, int N, int K, cudaStream_t st) {
  const size_t n = (size_t)N * K;
  TD_LAUNCH(transpose_kernel, (unsigned)((n + 255) / 256), 256, 0, st, W, Wt, N, K);
  return 0;
}

static int launch_framed_wgrad(const float* M, const float* sig, float* dW, int B, int L0, int CI, int NO, int K,
                               int S, int T, int shift, int m_stride, cudaStream_t st) {
  TD_REQUIRE(K <= 1024, "framed_wgrad: window %d", K);
  const int R = B * L0;
  int splits = cdiv(148 * 24, CI * NO);   // CTAs of K (64) threads: ~24 per SM, ~150 rows each at the training batch
  const int rps = cdiv(R, splits);
  splits = cdiv(R, rps);
  dim3 grid(CI * NO, splits);
  TD_LAUNCH(framed_wgrad_kernel, grid, K, 0, st, M, sig, dW, B, L0, CI, NO, K, S, T, shift, rps, m_stride);
  return 0;
}

// ----------------------------------------------------------------------------- orchestration
// `bf16`: x is one of the large stored activations and act_dtype is bf16 (everything else the sweep reads is fp32)
static GradSrc gln_grad(const float* dy, const float* x, const NormRef& n, const double* S, int bf16 = 0) {
  GradSrc g{};
  g.dy = dy; g.x = x; g.norm = n; g.S = S; g.kind = G_GLN; g.x_bf16 = bf16;
  return g;
}
static SrcDesc bplain(const float* x, int L, int bf16 = 0) {
  SrcDesc s{};
  s.x = x; s.L = L; s.bf16 = bf16;
  return s;
}
static SrcDesc baffine(const float* x, int L, const NormRef& n, const float* slope = nullptr, int bf16 = 0) {
  SrcDesc s{};
  s.x = x; s.L = L; s.norm = n; s.slope = slope; s.bf16 = bf16;
  return s;
}

// transposed (and, for the tensor-core path, TF32-rounded) copies of the weights the data-gradient GEMMs read
static int prepare_transposed(const BCtx& x) {
  const tdanet_weights_t* w = x.w;
  const Plan& p = *x.p;
  const int C = x.c->in_channels, cc = x.c->out_channels;
  struct { const float* w; size_t wt, aux; int N, K; } list[] = {
      {w->proj.w, p.wt_proj, p.auxt_proj, C, cc},     {w->res_w, p.wt_res, p.auxt_res, cc, C},
      {w->in_proj_w, p.wt_in, p.auxt_in, 3 * C, C},   {w->out_proj_w, p.wt_out, p.auxt_out, C, C},
      {w->fc1.w, p.wt_fc1, p.auxt_fc1, 2 * C, C},     {w->fc2.w, p.wt_fc2, p.auxt_fc2, C, 2 * C},
  };
  for (auto& e : list) {
    if (int r = launch_transpose(e.w, x.at(e.wt), e.N, e.K, x.st)) return r;
    if (x.c->gemm_mode != TDANET_GEMM_FP32)
      if (int r = launch_tf32_prepare(x.at(e.wt), x.at(e.aux), (size_t)e.N * e.K, x.c->gemm_mode, x.st)) return r;
  }
  if (x.c->variant == TDANET_FORK)
    for (int j = 0; j < x.c->depth; ++j) {
      if (int r = launch_transpose(w->conv_pool[j].pw_w, x.at(p.wt_pool[j]), C, C, x.st)) return r;
      if (x.c->gemm_mode != TDANET_GEMM_FP32)
        if (int r = launch_tf32_prepare(x.at(p.wt_pool[j]), x.at(p.auxt_pool[j]), (size_t)C * C, x.c->gemm_mode, x.st)) return r;
    }
  return 0;
}

static int launch_bf16_to_f32(const float* src, float* dst, size_t n, cudaStream_t st) {
  TD_REQUIRE(n % 4 == 0, "bf16_to_f32: n=%zu", n);
  const size_t n4 = n / 4;
  TD_LAUNCH(bf16_to_f32_kernel, (unsigned)((n4 + 255) / 256 > 8192 ? 8192 : (n4 + 255) / 256), 256, 0, st, src, dst, n4);
  return 0;
}

// weight gradient on the side stream `w`: it waits for everything the main stream has enqueued so far (its operands)
// and is joined back at the end of the block (uconv_block_backward) before any operand buffer is reused
static int wgrad_side(const BCtx& x, const float* G, const float* A, float* dW, float* db, int R, int N, int K) {
  if (int e = x.side->order(x.st, x.side->w)) return e;
  return launch_wgrad(G, A, dW, db, R, N, K, nullptr, x.side->w, x.c->gemm_mode);
}

// D[B, L, N] = A[B, L, K] . Wt[N, K]^T (+ resid)
static int dgrad(const BCtx& x, const float* A, size_t wt, size_t aux, float* D, int L, int N, int K, const float* resid) {
  GemmArgs g{};
  g.A = A; g.W = x.at(wt); g.bias = nullptr; g.D = D;
  g.B = x.p->B; g.L = L; g.N = N; g.K = K;
  if (resid) { g.epi = EPI_RESIDUAL; g.resid = resid; g.last = 1; }
  else g.epi = EPI_BIAS;
  return bgemm(x, g, aux);
}

// One LA (last_layer[i] with k = 5, or loc_glo_fus[k] with k = 1) backwards.
//   dout: gradient w.r.t. the LA output [B, Ll, C].  d_loc_in / d_glo_in receive the gradients w.r.t. the two
//   conv inputs (after their on-load transform), written or accumulated.
static int la_backward(const BCtx& x, int ks, const tdanet_la_t& la, const tdanet_la_t& gla, const SrcDesc& loc, int lkind,
                       const float* glo, int Lg, const NormRef& nL, const NormRef& nA, const NormRef& nE,
                       const size_t bs[3], const float* dout, float* d_loc_in, int acc_loc, float* d_glo_in, int acc_glo,
                       cudaStream_t local_st = nullptr, int glo_bf16 = 0, float* pre_rawb = nullptr, float* pre_rawe = nullptr) {
  const Plan& p = *x.p;
  const int B = p.B, C = x.c->in_channels, Ll = loc.L;
  LaBwdArgs a{};
  a.loc = loc; a.lkind = lkind; a.glo = glo; a.glo_bf16 = glo_bf16; a.Lg = Lg; a.B = B; a.C = C;
  a.wl = la.local_embedding.w; a.wa = la.global_act.w; a.we = la.global_embedding.w;
  a.nL = nL; a.nA = nA; a.nE = nE; a.dout = dout; a.scale = nearest_scale(Lg, Ll);
  const int ts = x.tset;
  a.d_loc = x.at(p.t_dloc[ts]); a.raw_a = x.at(p.t_rawa[ts]);
  a.d_act = x.at(p.t_dact[ts]); a.d_emb = x.at(p.t_demb[ts]); a.raw_b = x.at(p.t_rawb[ts]); a.raw_e = x.at(p.t_rawe[ts]);
  if (pre_rawb) { a.raw_b = pre_rawb; a.raw_e = pre_rawe; }   // pass G already done into per-step buffers
  a.dgamma[0] = x.gp(gla.local_embedding.gamma); a.dbeta[0] = x.gp(gla.local_embedding.beta);
  a.dgamma[1] = x.gp(gla.global_act.gamma); a.dbeta[1] = x.gp(gla.global_act.beta);
  a.dgamma[2] = x.gp(gla.global_embedding.gamma); a.dbeta[2] = x.gp(gla.global_embedding.beta);
  for (int i = 0; i < 3; ++i) a.S[i] = x.at<double>(bs[i]);
  { Tag t(ks == 5 ? "bwd_la_a" : "bwd_lgf_a"); if (int e = launch_la_bwd_a(a, ks, x.st, pre_rawb != nullptr)) return e; }
  Tag t(ks == 5 ? "bwd_la_dw" : "bwd_lgf_dw");
  DwBwdArgs d{};
  d.g[0] = gln_grad(a.d_loc, a.raw_a, nL, a.S[0]);
  d.w[0] = la.local_embedding.w; d.dw[0] = x.gp(gla.local_embedding.w);
  d.xin = loc; d.xkind = lkind; d.B = B; d.C = C; d.Lin = Ll; d.Lout = Ll; d.stride = 1;
  d.dx = d_loc_in; d.accumulate = acc_loc;
  if (local_st && local_st != x.st) {
    // the local branch only feeds loc_glo_fus: it runs on the side stream while x.st continues with the global branch
    if (int e = x.side->order(x.st, local_st)) return e;
    if (int e = launch_dw_bwd(x.on(local_st, x.tset), d, ks, 1)) return e;
  } else if (int e = launch_dw_bwd(x, d, ks, 1)) return e;
  d = DwBwdArgs{};
  d.g[0] = gln_grad(a.d_act, a.raw_b, nA, a.S[1]);
  d.g[1] = gln_grad(a.d_emb, a.raw_e, nE, a.S[2]);
  d.w[0] = la.global_act.w; d.dw[0] = x.gp(gla.global_act.w);
  d.w[1] = la.global_embedding.w; d.dw[1] = x.gp(gla.global_embedding.w);
  d.xin = bplain(glo, Lg, glo_bf16); d.xkind = SRC_PLAIN; d.B = B; d.C = C; d.Lin = Lg; d.Lout = Lg; d.stride = 1;
  d.dx = d_glo_in; d.accumulate = acc_glo;
  return launch_dw_bwd(x, d, ks, 2);
}

// GA / GlobalAttention backwards: g_ga_out -> g_ga_in
static int global_attention_backward(const BCtx& x) {
  const tdanet_config_t* c = x.c;
  const tdanet_weights_t* w = x.w;
  const tdanet_weights_t* gw = x.g;
  const Plan& p = *x.p;
  const int B = p.B, C = c->in_channels, Lb = p.Lb, R = B * Lb;
  const bool time_axis = c->variant == TDANET_MULTRES;  // MultiHeadAttentionFixed: sequence = time, LN(x_pe + drop(a))
  const int group = c->attn_group > 0 ? c->attn_group : B;
  TD_REQUIRE(time_axis || B % group == 0, "batch %d is not a multiple of attn_group %d", B, group);
  // ga_out = ga_mid + gLN(fc2)
  const NormRef n_fc2 = norm_ref(x, p.st_fc2, 2, (double)Lb * C, w->fc2.gamma, w->fc2.beta);
  const NormRef n_fc1 = norm_ref(x, p.st_fc1, 2, (double)Lb * 2 * C, w->fc1.gamma, w->fc1.beta);
  // training-mode keep-masks of this iteration (drawn by the forward pass, dropout.cu); null when p = 0
  const float ik = p.drop_elem ? 1.f / (1.f - c->dropout) : 1.f, ikp = p.drop_item ? 1.f / (1.f - c->drop_path) : 1.f;
  const uint8_t* m_att = p.drop_elem ? x.at<uint8_t>(p.m_att) : nullptr;
  const uint8_t* m_ao = p.drop_elem ? x.at<uint8_t>(p.m_ao) : nullptr;
  const uint8_t* m_f1 = p.drop_elem ? x.at<uint8_t>(p.m_f1) : nullptr;
  const uint8_t* m_f2 = p.drop_elem ? x.at<uint8_t>(p.m_f2) : nullptr;
  const uint8_t* m_dp = p.drop_item ? x.at<uint8_t>(p.m_dp) : nullptr;
  const size_t nRC = (size_t)R * C, per_item = (size_t)Lb * C;
  {
    Tag t("bwd_bottom_misc");
    // ga_out = ga_mid + DropPath(drop(gLN(fc2))): gradient w.r.t. the GlobLN output (g_attn_out is free here)
    const float* dy = x.at(p.g_ga_out);
    if (m_f2 || m_dp) {
      if (int e = launch_mask_scale(dy, x.at(p.g_attn_out), nRC, m_f2, 0.f, ik, m_dp ? m_dp + B : nullptr, ikp, per_item, 0, x.st)) return e;
      dy = x.at(p.g_attn_out);
    }
    if (int e = launch_gln_bwd_stats(dy, x.at(p.fc2), n_fc2, x.gp(gw->fc2.gamma), x.gp(gw->fc2.beta),
                                     x.at<double>(p.bs_fc2), B, Lb, C, x.st)) return e;
    if (int e = launch_gln_bwd_apply(gln_grad(dy, x.at(p.fc2), n_fc2, x.at<double>(p.bs_fc2)),
                                     x.at(p.g_fc2), 0, B, Lb, C, x.st)) return e;
  }
  { Tag t("wgrad_fc2"); if (int e = wgrad_side(x, x.at(p.g_fc2), x.at(p.ffn_dw), x.gp(gw->fc2.w), nullptr, R, C, 2 * C)) return e; }
  { Tag t("dgrad_fc2"); if (int e = dgrad(x, x.at(p.g_fc2), p.wt_fc2, p.auxt_fc2, x.at(p.g_ffn), Lb, 2 * C, C, nullptr)) return e; }
  {
    // relu -> dwconv k5 (+bias) on gLN(fc1); FFN.drop sits after the ReLU: ffn_dw holds the dropped tensor, whose
    // sign still marks the live elements, the gradient takes the mask / keep factor here
    Tag t("bwd_ffn_dw");
    if (m_f1)
      if (int e = launch_mask_scale(x.at(p.g_ffn), x.at(p.g_ffn), (size_t)R * 2 * C, m_f1, 0.f, ik, nullptr, 1.f, 1, 0, x.st)) return e;
    DwBwdArgs d{};
    d.g[0].dy = x.at(p.g_ffn); d.g[0].x = x.at(p.ffn_dw); d.g[0].kind = G_RELU;
    d.w[0] = w->ffn_dw_w; d.dw[0] = x.gp(gw->ffn_dw_w); d.db[0] = x.gp(gw->ffn_dw_b);
    d.xin = baffine(x.at(p.fc1), Lb, n_fc1); d.xkind = SRC_AFFINE;
    d.B = B; d.C = 2 * C; d.Lin = Lb; d.Lout = Lb; d.stride = 1; d.dx = x.at(p.g_fc1);
    // sole producer of g_fc1: accumulates the backward sums of the fc1 GlobLN as well
    d.up_dgamma = x.gp(gw->fc1.gamma); d.up_dbeta = x.gp(gw->fc1.beta); d.up_S = x.at<double>(p.bs_fc1);
    if (int e = launch_dw_bwd(x, d, 5, 1)) return e;
  }
  {
    Tag t("bwd_bottom_misc");
    if (int e = launch_gln_bwd_apply(gln_grad(x.at(p.g_fc1), x.at(p.fc1), n_fc1, x.at<double>(p.bs_fc1)),
                                     x.at(p.g_ffn), 0, B, Lb, 2 * C, x.st)) return e;  // g_ffn is free again
  }
  { Tag t("wgrad_fc1"); if (int e = wgrad_side(x, x.at(p.g_ffn), x.at(p.ga_mid), x.gp(gw->fc1.w), nullptr, R, 2 * C, C)) return e; }
  // g_ga_mid = g_ga_out (skip) + fc1 data gradient
  { Tag t("dgrad_fc1"); if (int e = dgrad(x, x.at(p.g_ffn), p.wt_fc1, p.auxt_fc1, x.at(p.g_ga_mid), Lb, C, 2 * C, x.at(p.g_ga_out))) return e; }
  // ga_mid = ga_in + DropPath(LN2(2 * attn_out));  with dropout: LN2(attn_out') where the forward left
  // attn_out' = attn_out * (1 + mask/keep) in the workspace.
  // MULTRES: ga_mid = ga_in + DropPath(LN2(attn_in + attn_out')), attn_out' = attn_out * mask/keep: the gradient w.r.t.
  // the LayerNorm input goes to attn_out (through the mask) AND, as `g_ln_in`, to attn_in (added by dgrad_in_proj).
  const float* g_ln_in = nullptr;
  { Tag t("bwd_bottom_misc");
    const float* dy = x.at(p.g_ga_mid);
    if (m_dp) {  // g_ctx is free here
      if (int e = launch_mask_scale(dy, x.at(p.g_ctx), nRC, nullptr, 1.f, 0.f, m_dp, ikp, per_item, 0, x.st)) return e;
      dy = x.at(p.g_ctx);
    }
    if (time_axis) {
      float* z = x.at(p.g_qkv);  // free until the attention backward writes it
      TD_LAUNCH(add_kernel, (unsigned)((nRC + 255) / 256 > 4096 ? 4096 : (nRC + 255) / 256), 256, 0, x.st, x.at(p.attn_in),
                x.at(p.attn_out), z, nRC);
      float* gz = m_ao ? x.at(p.g_fc1) : x.at(p.g_attn_out);  // g_fc1 is free by now and read by no side-stream launch
      if (int e = launch_ln_bwd(z, 1.f, w->ln2_w, dy, x.at(p.ln_rows), nullptr, gz, x.gp(gw->ln2_w), x.gp(gw->ln2_b), R, C, x.st)) return e;
      if (m_ao)
        if (int e = launch_mask_scale(gz, x.at(p.g_attn_out), nRC, m_ao, 0.f, ik, nullptr, 1.f, 1, 0, x.st)) return e;
      g_ln_in = gz;
    } else {
      if (int e = launch_ln_bwd(x.at(p.attn_out), m_ao ? 1.f : 2.f, w->ln2_w, dy, x.at(p.ln_rows), nullptr,
                                x.at(p.g_attn_out), x.gp(gw->ln2_w), x.gp(gw->ln2_b), R, C, x.st)) return e;
      if (m_ao)
        if (int e = launch_mask_scale(x.at(p.g_attn_out), x.at(p.g_attn_out), nRC, m_ao, 1.f, ik, nullptr, 1.f, 1, 0, x.st)) return e;
    }
  }
  { Tag t("wgrad_out_proj"); if (int e = wgrad_side(x, x.at(p.g_attn_out), x.at(p.attn_ctx), x.gp(gw->out_proj_w), x.gp(gw->out_proj_b), R, C, C)) return e; }
  { Tag t("dgrad_out_proj"); if (int e = dgrad(x, x.at(p.g_attn_out), p.wt_out, p.auxt_out, x.at(p.g_ctx), Lb, C, C, nullptr)) return e; }
  { Tag t("bwd_attention");
    if (int e = launch_att_bwd(x.at(p.qkv), x.at(p.g_ctx), x.at(p.att_p), x.at(p.att_ds), x.at(p.g_qkv), B, Lb, C,
                               c->n_head, group, time_axis, m_att, ik, x.st)) return e; }
  { Tag t("wgrad_in_proj"); if (int e = wgrad_side(x, x.at(p.g_qkv), x.at(p.attn_in), x.gp(gw->in_proj_w), x.gp(gw->in_proj_b), R, 3 * C, C)) return e; }
  { Tag t("dgrad_in_proj"); if (int e = dgrad(x, x.at(p.g_qkv), p.wt_in, p.auxt_in, x.at(p.g_attn_in), Lb, C, 3 * C, g_ln_in)) return e; }
  // attn_in = LN1(ga_in) + pe;  g_ga_in = g_ga_mid (skip) + LN1 backward
  Tag t("bwd_bottom_misc");
  return launch_ln_bwd(x.at(p.ga_in), 1.f, w->ln1_w, x.at(p.g_attn_in), x.at(p.ln_rows), x.at(p.g_ga_mid),
                       x.at(p.g_ga_in), x.gp(gw->ln1_w), x.gp(gw->ln1_b), R, C, x.st);
}

// One UConvBlock iteration backwards.  d_y: gradient w.r.t. the block output y = res_conv(...) + in  [B, L0, c];
// d_in: receives the gradient w.r.t. the block input.
static int uconv_block_backward(const BCtx& x, const float* in, const float* d_y, float* d_in) {
  const tdanet_config_t* c = x.c;
  const tdanet_weights_t* w = x.w;
  const tdanet_weights_t* gw = x.g;
  const Plan& p = *x.p;
  const int B = p.B, C = c->in_channels, cc = c->out_channels, depth = c->depth, Lb = p.Lb, L0 = p.L[0];
  const int R0 = B * L0;
  const int abf = x.bf() ? 1 : 0;   // the large stored activations (proj, spp, x_fused, expanded) are bf16
  cudaEvent_t wres_done = nullptr;
  TD_CUDA(cudaMemsetAsync(x.at<char>(p.bs_begin), 0, p.bs_end - p.bs_begin, x.st));
  auto spp_norm = [&](int k) {
    return norm_ref(x, p.st_spp[k], 2, (double)p.L[k] * C, w->spp_dw[k].gamma, w->spp_dw[k].beta);
  };
  // ---- res_conv
  {
    Tag t("wgrad_res_conv");
    const float* a_op = x.at(p.expanded[0]);
    if (abf) {
      // the weight-gradient kernel reads fp32 rows: widen expanded[0] into a free LA temporary (set 1 is untouched
      // until the second top-down step, which waits for this weight gradient below)
      if (int e = launch_bf16_to_f32(x.at(p.expanded[0]), x.at(p.t_rawa[1]), (size_t)R0 * C, x.st)) return e;
      a_op = x.at(p.t_rawa[1]);
    }
    if (int e = wgrad_side(x, d_y, a_op, x.gp(gw->res_w), x.gp(gw->res_b), R0, cc, C)) return e;
    if (abf) {
      wres_done = x.side->events[x.side->next++ % x.side->events.size()];
      TD_CUDA(cudaEventRecord(wres_done, x.side->w));
    }
  }
  { Tag t("dgrad_res_conv"); if (int e = dgrad(x, d_y, p.wt_res, p.auxt_res, x.at(p.g_exp[0]), L0, C, cc, nullptr)) return e; }
  // ---- top-down fusion, in the reverse of the forward order.  Main stream: passes G/L/F and the global-branch
  // depthwise backward of every step (the chain g_exp[i] -> g_exp[i+1]).  Side stream `l`: the local-branch depthwise
  // backward of step i (-> g_fused[i]) followed by loc_glo_fus[i] backwards (-> g_spp[i], g_ga_out).
  cudaStream_t sl = x.side->l;
  static const int lgf_streams = getenv("TDANET_LGF_STREAMS") ? atoi(getenv("TDANET_LGF_STREAMS")) : 2;  // 0: on `l` (round 1)
  const int gi = first_step_partner(depth);  // x_fused[gi] also receives the first step's global-branch gradient
  bool fused_written[TDANET_MAX_DEPTH] = {};
  bool spp_written[TDANET_MAX_DEPTH] = {};
  bool ga_out_written[2] = {false, false};   // per accumulator (g_ga_out, g_ga_out2)
  int n_lgf = 0;
  cudaEvent_t local_done[TDANET_MAX_DEPTH] = {};
  // loc_glo_fus[k] (1-tap LA): x_fused[k] = LA(gLN(spp_k), ga_out).  Runs on the loc_glo_fus stream n_lgf % 2 with
  // temporary set 2 + n_lgf % 2 and accumulator n_lgf % 2, after `ready` (g_fused[k] complete).
  auto lgf_backward = [&](int k, cudaEvent_t ready) -> int {
    const int q = lgf_streams >= 2 ? (n_lgf & 1) : 0;
    cudaStream_t sf = lgf_streams >= 1 ? x.side->f[q] : sl;
    ++n_lgf;
    if (sf != sl && ready) TD_CUDA(cudaStreamWaitEvent(sf, ready, 0));
    float* g_out = q ? x.at(p.g_ga_out2) : x.at(p.g_ga_out);
    if (c->variant != TDANET_BEST) {
      // x_fused[k] = n_k + near(ga_out): g_fused[k] aliases g_spp[k]; the global feature gets the per-centre sums
      Tag t("bwd_inject_add");
      int jc = (int)((double)pick_rows(p.L[k], C / 4, B, 32) * Lb / p.L[k] + 0.5);
      jc = jc < 1 ? 1 : jc;
      dim3 grid;
      int threads;
      row_grid(Lb, C / 4, B, jc, grid, threads);
      TD_LAUNCH(inject_add_bwd_kernel, grid, threads, 0, sf, x.at(p.g_fused[k]), g_out, (int)ga_out_written[q],
                p.L[k], Lb, C, nearest_scale(Lb, p.L[k]), jc);
      spp_written[k] = true;
      ga_out_written[q] = true;
      return 0;
    }
    const tdanet_la_t& la = w->loc_glo_fus[k];
    const NormRef nL = norm_ref(x, p.st_lgf[k], 6, (double)p.L[k] * C, la.local_embedding.gamma, la.local_embedding.beta);
    const NormRef nA = norm_ref(x, p.st_lgf[k] + 2 * sizeof(double), 6, (double)Lb * C, la.global_act.gamma, la.global_act.beta);
    const NormRef nE = norm_ref(x, p.st_lgf[k] + 4 * sizeof(double), 6, (double)Lb * C, la.global_embedding.gamma, la.global_embedding.beta);
    if (int e = la_backward(x.on(sf, 2 + q), 1, la, gw->loc_glo_fus[k], baffine(x.at(p.spp[k]), p.L[k], spp_norm(k), nullptr, abf), SRC_AFFINE,
                            x.at(p.ga_out), Lb, nL, nA, nE, p.bs_lgf[k], x.at(p.g_fused[k]), x.at(p.g_spp[k]), 0,
                            g_out, ga_out_written[q])) return e;
    spp_written[k] = true;
    ga_out_written[q] = true;
    return 0;
  };
  if (lgf_streams >= 1) {   // the loc_glo_fus streams join the sweep here (their first wait is an event of `l`)
    if (int e = x.side->order(x.st, x.side->f[0])) return e;
    if (int e = x.side->order(x.st, x.side->f[1])) return e;
  }
  // Pass G of every top-down step (raw_b / raw_e = the global-branch conv outputs of the step's "global" operand) reads
  // forward tensors only: all of them run ahead of the chain on stream `g`, into per-step buffers, instead of as the
  // first launch of every step on the critical path (12 + 10 + 8 + 9 us per block at B = 8).
  // Measured neutral (52.6 vs 52.5 steps/s: the join at the end of the top-down phase waits for the loc_glo_fus streams,
  // not for the main chain), so it stays off; TDANET_LA_G_AHEAD=1 enables it.
  static const bool g_ahead = getenv("TDANET_LA_G_AHEAD") && atoi(getenv("TDANET_LA_G_AHEAD")) == 1;
  cudaEvent_t g_done[TDANET_MAX_DEPTH] = {};
  if (g_ahead) {
    if (int e = x.side->order(x.st, x.side->g)) return e;
    for (int i = 0; i <= depth - 2; ++i) {
      const bool first = i == depth - 2;
      LaBwdArgs ga{};
      ga.glo = first ? x.at(p.fused[gi]) : x.at(p.expanded[i + 1]);
      ga.glo_bf16 = abf; ga.Lg = first ? p.L[gi] : p.L[i + 1]; ga.B = B; ga.C = C;
      ga.wa = w->last_layer[i].global_act.w; ga.we = w->last_layer[i].global_embedding.w;
      ga.raw_b = x.at(p.t_rawb_step[i]); ga.raw_e = x.at(p.t_rawe_step[i]);
      Tag t("bwd_la_a");
      if (int e = launch_la_bwd_g(ga, 5, x.side->g)) return e;
      g_done[i] = x.side->events[x.side->next++ % x.side->events.size()];
      TD_CUDA(cudaEventRecord(g_done[i], x.side->g));
    }
  }
  for (int i = 0; i <= depth - 2; ++i) {
    const tdanet_la_t& la = w->last_layer[i];
    const bool first = i == depth - 2;  // the first forward step: its "global" input is x_fused[gi]
    const int Lg = first ? p.L[gi] : p.L[i + 1];
    const float* glo = first ? x.at(p.fused[gi]) : x.at(p.expanded[i + 1]);
    if (g_ahead) TD_CUDA(cudaStreamWaitEvent(x.st, g_done[i], 0));
    const NormRef nL = norm_ref(x, p.st_la_l[i], 2, (double)p.L[i] * C, la.local_embedding.gamma, la.local_embedding.beta);
    const NormRef nA = norm_ref(x, p.st_la_g[i], 4, (double)Lg * C, la.global_act.gamma, la.global_act.beta);
    const NormRef nE = norm_ref(x, p.st_la_g[i] + 2 * sizeof(double), 4, (double)Lg * C, la.global_embedding.gamma, la.global_embedding.beta);
    float* d_glo = first ? x.at(p.g_fused[gi]) : x.at(p.g_exp[i + 1]);
    // this step's passes overwrite temporary set i % 2: the local branch of step i - 2 must have read it
    if (i >= 2) TD_CUDA(cudaStreamWaitEvent(x.st, local_done[i - 2], 0));
    // ... and set 0 holds the G operand of the previous block's proj weight gradient, which is still running on `w`
    // (joining it at the end of that block stalled the main stream for the ~17 us the wgrad outlasts dgrad_proj)
    if (i == 0 && x.w_pending && *x.w_pending) {
      TD_CUDA(cudaStreamWaitEvent(x.st, *x.w_pending, 0));
      *x.w_pending = nullptr;
    }
    // the first step adds into g_fused[gi], which the side stream wrote in step gi
    if (first && fused_written[gi]) TD_CUDA(cudaStreamWaitEvent(x.st, local_done[gi], 0));
    if (i == 1 && wres_done) TD_CUDA(cudaStreamWaitEvent(x.st, wres_done, 0));  // t_rawa[1] held res_conv's fp32 operand
    if (int e = la_backward(x.on(x.st, i & 1), 5, la, gw->last_layer[i], bplain(x.at(p.fused[i]), p.L[i], abf), SRC_PLAIN, glo, Lg,
                            nL, nA, nE, p.bs_la[i], x.at(p.g_exp[i]), x.at(p.g_fused[i]), fused_written[i], d_glo,
                            first ? (int)fused_written[gi] : 0, sl, abf,
                            g_ahead ? x.at(p.t_rawb_step[i]) : nullptr, g_ahead ? x.at(p.t_rawe_step[i]) : nullptr)) return e;
    fused_written[i] = true;
    if (first) fused_written[gi] = true;
    local_done[i] = x.side->events[x.side->next++ % x.side->events.size()];
    TD_CUDA(cudaEventRecord(local_done[i], sl));
    // x_fused[i] is complete unless it still waits for the first step's global-branch gradient
    if (i != gi || first) {
      if (int e = lgf_backward(i, local_done[i])) return e;
    }
  }
  if (gi != depth - 2 && fused_written[gi]) {
    // deferred: x_fused[gi] got its second contribution from the main stream in the last iteration
    cudaEvent_t ev = x.side->events[x.side->next++ % x.side->events.size()];
    TD_CUDA(cudaEventRecord(ev, x.st));
    if (lgf_streams < 1) TD_CUDA(cudaStreamWaitEvent(sl, ev, 0));
    if (int e = lgf_backward(gi, ev)) return e;
  }
  TD_REQUIRE(ga_out_written[0], "no live x_fused tensor");
  if (int e = x.side->order(sl, x.st)) return e;  // g_fused[*] (local branches)
  if (lgf_streams >= 1) {
    if (int e = x.side->order(x.side->f[0], x.st)) return e;  // g_spp[*], g_ga_out
    if (int e = x.side->order(x.side->f[1], x.st)) return e;  // g_spp[*], g_ga_out2
  }
  if (ga_out_written[1]) {
    const size_t n = (size_t)B * Lb * C;
    Tag t("bwd_bottom_misc");
    TD_LAUNCH(add_kernel, (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, x.st, x.at(p.g_ga_out),
              x.at(p.g_ga_out2), x.at(p.g_ga_out), n);
  }
  // ---- bottom-scale block
  if (int e = global_attention_backward(x)) return e;
  if (c->variant == TDANET_FORK) {
    // ---- ga_in = sum_k gLN(pw_conv(dw_conv(gLN(spp_k))))  with conv_pool[depth-1-k] (TDANet.py:605-620)
    const int R = B * Lb;
    for (int k = 0; k < depth; ++k) {
      const int j = depth - 1 - k, s = 1 << j, ks = j == 0 ? 5 : 2 * s + 1;
      const tdanet_sepconvnorm_t& q = w->conv_pool[j];
      const tdanet_sepconvnorm_t& gq = gw->conv_pool[j];
      const NormRef nq = norm_ref(x, p.st_pool[k], 2, (double)Lb * C, q.gamma, q.beta);
      Tag t("bwd_conv_pool");
      if (int e = launch_gln_bwd_stats(x.at(p.g_ga_in), x.at(p.pool_pw[k]), nq, x.gp(gq.gamma), x.gp(gq.beta),
                                       x.at<double>(p.bs_pool[k]), B, Lb, C, x.st)) return e;
      if (int e = launch_gln_bwd_apply(gln_grad(x.at(p.g_ga_in), x.at(p.pool_pw[k]), nq, x.at<double>(p.bs_pool[k])),
                                       x.at(p.g_pool_x[k]), 0, B, Lb, C, x.st)) return e;
      if (int e = wgrad_side(x, x.at(p.g_pool_x[k]), x.at(p.pool_dw[k]), x.gp(gq.pw_w), x.gp(gq.pw_b), R, C, C)) return e;
      if (int e = dgrad(x, x.at(p.g_pool_x[k]), p.wt_pool[j], p.auxt_pool[j], x.at(p.g_pool_dw), Lb, C, C, nullptr)) return e;
      // depthwise conv of conv_pool[j] on gLN(spp_k): data gradient into g_spp[k], weights / bias
      const int rows = pick_rows(p.L[k], C / 4, B, 16);
      dim3 grid;
      int threads;
      row_grid(p.L[k], C / 4, B, rows, grid, threads);
      TD_LAUNCH(dwg_bwd_data_kernel, grid, threads, 0, x.st, x.at(p.g_pool_dw), q.dw_w, x.at(p.g_spp[k]), (int)spp_written[k],
                p.L[k], Lb, C, ks, s, rows);
      spp_written[k] = true;
      // rows per CTA: as long as possible (fewer atomics) while the launch still has ~4 CTAs per SM
      int wrows = 4;
      while (wrows < Lb && (long)cdiv(C / 4, threads) * (ks + 1) * B * cdiv(Lb, wrows * 2) >= 4 * 148) wrows *= 2;
      dim3 wgrid(cdiv(C / 4, threads), ks + 1, B * cdiv(Lb, wrows));
      if (abf) TD_LAUNCH((dwg_bwd_weight_kernel<true>), wgrid, threads, 0, x.st, x.at(p.g_pool_dw), baffine(x.at(p.spp[k]), p.L[k], spp_norm(k), nullptr, abf),
                         (int)SRC_AFFINE, x.gp(gq.dw_w), x.gp(gq.dw_b), B, Lb, C, ks, s, wrows);
      else TD_LAUNCH((dwg_bwd_weight_kernel<false>), wgrid, threads, 0, x.st, x.at(p.g_pool_dw), baffine(x.at(p.spp[k]), p.L[k], spp_norm(k)),
                     (int)SRC_AFFINE, x.gp(gq.dw_w), x.gp(gq.dw_b), B, Lb, C, ks, s, wrows);
    }
  } else {
    // ---- ga_in = sum_k avgpool(gLN(spp_k)): the coarsest scale (identity bins) here, the others inside the spp_dw
    // backward that completes g_spp[k]
    Tag t("bwd_pool");
    const int k = depth - 1;
    if (int e = launch_pool_bwd(x.at(p.g_ga_in), x.at(p.g_spp[k]), spp_written[k], B, p.L[k], Lb, C, x.st)) return e;
  }
  // ---- spp_dw chain.  The depthwise backward of spp_dw[k] is the last launch that adds to g_spp[k-1] (k = 0: the only
  // one that writes g_proj), so it also accumulates the backward sums of that GlobLN; only the coarsest scale, whose
  // gradient is completed by the pooling backward, needs a statistics launch of its own.
  const NormRef n_proj = norm_ref(x, p.st_proj, 2, (double)L0 * C, w->proj.gamma, w->proj.beta);
  for (int k = depth - 1; k >= 0; --k) {
    const NormRef nk = spp_norm(k);
    if (k == depth - 1) {
      Tag t("bwd_gln_stats");
      if (int e = launch_gln_bwd_stats_t(abf, x.at(p.g_spp[k]), x.at(p.spp[k]), nk, x.gp(gw->spp_dw[k].gamma),
                                       x.gp(gw->spp_dw[k].beta), x.at<double>(p.bs_spp[k]), B, p.L[k], C, x.st)) return e;
    }
    DwBwdArgs d{};
    d.g[0] = gln_grad(x.at(p.g_spp[k]), x.at(p.spp[k]), nk, x.at<double>(p.bs_spp[k]), abf);
    d.w[0] = w->spp_dw[k].w; d.dw[0] = x.gp(gw->spp_dw[k].w); d.db[0] = x.gp(gw->spp_dw[k].b);
    d.B = B; d.C = C; d.Lout = p.L[k];
    if (k == 0) {
      d.xin = baffine(x.at(p.proj), L0, n_proj, w->proj_prelu, abf);
      d.xkind = SRC_AFFINE_PRELU; d.Lin = L0; d.stride = 1; d.dx = x.at(p.g_proj); d.accumulate = 0;
      d.dslope = x.gp(gw->proj_prelu);
      d.up_dgamma = x.gp(gw->proj.gamma); d.up_dbeta = x.gp(gw->proj.beta); d.up_S = x.at<double>(p.bs_proj);
    } else {
      d.xin = baffine(x.at(p.spp[k - 1]), p.L[k - 1], spp_norm(k - 1), nullptr, abf);
      d.xkind = SRC_AFFINE; d.Lin = p.L[k - 1]; d.stride = 2; d.dx = x.at(p.g_spp[k - 1]);
      d.accumulate = spp_written[k - 1];  // the loc_glo_fus local branch wrote it (every scale but a dead last one)
      if (c->variant != TDANET_FORK) { d.pool_g = x.at(p.g_ga_in); d.pool_Lb = Lb; }
      d.up_dgamma = x.gp(gw->spp_dw[k - 1].gamma); d.up_dbeta = x.gp(gw->spp_dw[k - 1].beta);
      d.up_S = x.at<double>(p.bs_spp[k - 1]);
    }
    Tag t(k == 0 ? "bwd_spp_dw0" : "bwd_spp_dw_s2");
    if (int e = launch_dw_bwd(x, d, 5, 1)) return e;
  }
  // ---- proj_1x1 (its GlobLN sums were accumulated by the spp_dw[0] backward)
  { Tag t("bwd_gln_apply");
    if (int e = launch_gln_bwd_apply(gln_grad(x.at(p.g_proj), x.at(p.proj), n_proj, x.at<double>(p.bs_proj), abf),
                                     x.at(p.t_dloc[0]), 0, B, L0, C, x.st)) return e; }  // LA temporaries are free by now
  // every weight gradient enqueued so far is joined at the end of this block (they finished long ago); only the proj
  // one, enqueued next, may outlast the block
  cudaEvent_t w_before_proj = x.side->events[x.side->next++ % x.side->events.size()];
  TD_CUDA(cudaEventRecord(w_before_proj, x.side->w));
  { Tag t("wgrad_proj"); if (int e = wgrad_side(x, x.at(p.t_dloc[0]), in, x.gp(gw->proj.w), x.gp(gw->proj.b), R0, C, cc)) return e; }
  Tag t("dgrad_proj");
  if (int e = dgrad(x, x.at(p.t_dloc[0]), p.wt_proj, p.auxt_proj, d_in, L0, cc, C, d_y)) return e;
  // the weight-gradient GEMMs of this block must be done before the next block (or the caller) reuses their operands:
  // the next block waits for this event before its first write to the LA temporaries, backward() after the last block
  if (!x.w_pending) return x.side->order(x.side->w, x.st);
  TD_CUDA(cudaStreamWaitEvent(x.st, w_before_proj, 0));
  cudaEvent_t ev = x.side->events[x.side->next++ % x.side->events.size()];
  TD_CUDA(cudaEventRecord(ev, x.side->w));
  *x.w_pending = ev;
  return 0;
}

static int backward(const tdanet_config_t* c, const tdanet_weights_t* w, const tdanet_weights_t* gw, const float* wav,
                    const float* d_est, int B, int T, void* workspace, size_t ws_bytes, cudaStream_t st) {
  Plan p;
  if (int e = make_plan(c, B, T, p, true)) return e;
  TD_REQUIRE(w && gw && wav && d_est && workspace, "NULL argument");
  if (ws_bytes < p.bytes) return fail(TDANET_ENOSPACE, "workspace has %zu bytes, need %zu", ws_bytes, p.bytes);
  BCtx x{};
  std::vector<RepEntry> reps;
  x.c = c; x.w = w; x.p = &p; x.ws = (char*)workspace; x.st = st; x.blk = 0; x.g = gw; x.reps = &reps;
  int dev = 0;
  TD_CUDA(cudaGetDevice(&dev));
  TD_REQUIRE(dev >= 0 && dev < 16, "device %d", dev);
  std::lock_guard<std::mutex> enqueue_lock(device_enqueue_mutex(dev));
  x.side = &g_side[dev];
  x.main_st = st;
  if (int e = x.side->init()) return e;
  TD_CUDA(cudaMemsetAsync(x.at<char>(p.rep_arena), 0, p.rep_floats * TDANET_DW_REPLICAS * sizeof(float), st));
  const int K = c->enc_kernel, S = c->enc_stride, Nb = c->n_basis, cc = c->out_channels, L0 = p.L[0];
  const int NS = c->num_sources, CI = NS * Nb, R0 = B * L0, nb = c->num_blocks;

  // the transposed weight copies of the data-gradient GEMMs (12 short launches) are first needed by the res_conv
  // data gradient of the last block: they are prepared on the weight-gradient stream while the back end below runs
  cudaEvent_t wt_ready = nullptr;
  {
    if (int e = x.side->order(st, x.side->w)) return e;
    if (int e = prepare_transposed(x.on(x.side->w, 0))) return e;
    wt_ready = x.side->events[x.side->next++ % x.side->events.size()];
    TD_CUDA(cudaEventRecord(wt_ready, x.side->w));
  }
  TD_CUDA(cudaMemsetAsync(x.at<char>(p.bs_enc), 0, (size_t)B * 2 * sizeof(double), st));
  TD_CUDA(cudaMemsetAsync(x.at<char>(p.g_x0), 0, (size_t)R0 * cc * sizeof(float), st));
  {
    Tag t("bwd_backend");
    // decoder ConvTranspose1d + crop
    const int wpad = NS * K + 1;  // odd row stride: conflict-free weight reads
    if ((size_t)NS * ((DEC_FR - 1) * S + K) + (size_t)CI * wpad <= 11264) {
      dim3 dgrid(cdiv(L0, DEC_FR), B);
      TD_LAUNCH_COOP(dec_bwd_data_kernel, dgrid, 256, 0, st, d_est, w->dec_w, x.at(p.g_masked), B, L0, CI, NS, K, S, T, K - S, wpad);
    } else {
      const size_t n = (size_t)R0 * CI;
      TD_LAUNCH(dec_bwd_data_naive_kernel, (unsigned)((n + 127) / 128), 128, 0, st, d_est, w->dec_w, x.at(p.g_masked), B, L0, CI, NS, K, S, T, K - S);
    }
    // the two weight gradients of the back end (decoder basis: ~150 us of latency-bound launch; mask conv) read only
    // inputs of the call / g_masked and nothing else writes their outputs: they ride on the weight-gradient stream,
    // which the first block's event (or join) brings back long before g_masked is reused as d_nenc below
    if (int e = x.side->order(st, x.side->w)) return e;
    if (int e = launch_framed_wgrad(x.at(p.masked), d_est, x.gp(gw->dec_w), B, L0, CI, NS, K, S, T, K - S, CI, x.side->w)) return e;
    // masked = relu(m) * enc
    const size_t ne = (size_t)R0 * Nb;
    TD_LAUNCH(mask_bwd_kernel, (unsigned)((ne + 255) / 256), 256, 0, st, x.at(p.g_masked), x.at(p.mlogit), x.at(p.enc), x.at(p.g_enc), R0, NS, Nb);
    // m = mask_conv(prelu(y_last)) + bias
    const float* y_last = x.at_blk(p.y, nb - 1);
    if (int e = x.side->order(st, x.side->w)) return e;
    if (int e = launch_wgrad(x.at(p.g_masked), y_last, x.gp(gw->mask_w), x.gp(gw->mask_b), R0, CI, cc, w->mask_prelu, x.side->w)) return e;
    if (int e = launch_small_dgrad(x.at(p.g_masked), w->mask_w, x.at(p.g_u[0]), R0, CI, cc, y_last, w->mask_prelu, x.gp(gw->mask_prelu), st)) return e;
  }
  TD_CUDA(cudaStreamWaitEvent(st, wt_ready, 0));
  // Recurrent, backwards: g_u[cur] holds d loss / d y_blk
  int cur = 0;
  cudaEvent_t w_pending = nullptr;
  static const bool defer_w = !(getenv("TDANET_WGRAD_JOIN") && atoi(getenv("TDANET_WGRAD_JOIN")) == 1);
  for (int blk = nb - 1; blk >= 0; --blk) {
    BCtx xb = x;
    xb.blk = blk;
    xb.w_pending = defer_w ? &w_pending : nullptr;
    const float* in = blk == 0 ? x.at(p.x0) : xb.at(p.bin);
    if (int e = uconv_block_backward(xb, in, x.at(p.g_u[cur]), x.at(p.g_u[cur ^ 1]))) return e;
    cur ^= 1;  // g_u[cur] = d loss / d in_blk
    if (blk > 0) {
      // in_blk = concat_block(x0 + y_{blk-1})
      Tag t("bwd_concat");
      const int rpt = 32;
      const int threads = cc / 4 > 256 ? 256 : (cc / 4 < 32 ? 32 : cc / 4);
      dim3 grid(cdiv(R0, rpt), cdiv(cc / 4, threads));
      TD_LAUNCH_RED(concat_bwd_kernel, grid, threads, 0, st, x.at(p.g_u[cur]), x.at_blk(p.y, blk - 1), x.at(p.x0),
                     w->concat_w, w->concat_b, w->concat_prelu, x.at(p.g_u[cur ^ 1]), x.at(p.g_x0), x.gp(gw->concat_w),
                     x.gp(gw->concat_b), x.gp(gw->concat_prelu), R0, cc, rpt);
      cur ^= 1;  // g_u[cur] = d loss / d y_{blk-1}
    }
  }
  if (w_pending) TD_CUDA(cudaStreamWaitEvent(st, w_pending, 0));  // the last block's weight gradients
  Tag t("bwd_frontend");
  // d x0 = d in_0 + the mixture path of every concat_block
  {
    const size_t n = (size_t)R0 * cc;
    TD_LAUNCH(add_kernel, (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st, x.at(p.g_u[cur]), x.at(p.g_x0), x.at(p.g_x0), n);
  }
  // x0 = bottleneck(gLN(enc))   |   MULTRES: x0 = gLN(enc)
  const NormRef n_enc = norm_ref(x, p.st_enc, 2, (double)L0 * Nb, w->ln_gamma, w->ln_beta);
  float* d_nenc = x.at(p.g_masked);  // free by now, large enough
  if (c->variant == TDANET_MULTRES) {
    d_nenc = x.at(p.g_x0);
  } else {
    dim3 grid(cdiv(L0 * Nb, 256), 1, B);
    TD_LAUNCH(gln_fwd_apply_kernel, grid, 256, 0, st, x.at(p.enc), n_enc, x.at(p.nenc), L0, Nb);
    if (int e = launch_wgrad(x.at(p.g_x0), x.at(p.nenc), x.gp(gw->bottleneck_w), x.gp(gw->bottleneck_b), R0, cc, Nb, nullptr, st)) return e;
    if (int e = launch_small_dgrad(x.at(p.g_x0), w->bottleneck_w, d_nenc, R0, cc, Nb, nullptr, nullptr, nullptr, st)) return e;
  }
  if (int e = launch_gln_bwd_stats(d_nenc, x.at(p.enc), n_enc, x.gp(gw->ln_gamma), x.gp(gw->ln_beta), x.at<double>(p.bs_enc), B, L0, Nb, st)) return e;
  if (int e = launch_gln_bwd_apply(gln_grad(d_nenc, x.at(p.enc), n_enc, x.at<double>(p.bs_enc)), x.at(p.g_enc), 1, B, L0, Nb, st)) return e;
  // encoder Conv1d (pad_input folded into the indexing); MULTRES: conv k has window (k+1)*K and its slice of the channels
  {
    const int cpc = Nb / c->enc_convs;
    for (int k = 0; k < c->enc_convs; ++k)
      if (int e = launch_framed_wgrad(x.at(p.g_enc) + k * cpc, wav, x.gp(gw->enc_w[k]), B, L0, cpc, 1, (k + 1) * K, S, T, K - S, Nb, st)) return e;
  }
  // depthwise weight / bias gradients: replicas -> the caller's buffers
  if (!reps.empty()) {
    FoldArgs f{};
    f.count = (int)reps.size(); f.n_rep = TDANET_DW_REPLICAS; f.rep_stride = p.rep_floats;
    int nmax = 0;
    for (int i = 0; i < f.count; ++i) {
      f.e[i].dst = reps[i].dst; f.e[i].rep = x.at(p.rep_arena) + reps[i].off; f.e[i].n = reps[i].n;
      nmax = reps[i].n > nmax ? reps[i].n : nmax;
    }
    dim3 grid(cdiv(nmax, 256), f.count);
    TD_LAUNCH(fold_replicas_kernel, grid, 256, 0, st, f);
  }
  return 0;
}

}  // namespace td

using namespace td;

extern "C" {

int tdanet_backward(const tdanet_config_t* cfg, const tdanet_weights_t* w, const tdanet_weights_t* grads,
                    const float* wav, const float* d_est, int batch, int n_samples, void* workspace,
                    size_t workspace_bytes, tdanet_stream_t stream) {
  return backward(cfg, w, grads, wav, d_est, batch, n_samples, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tdanet_wgrad(int gemm_mode, const float* G, const float* A, float* dW, float* db, int rows, int N, int K,
                 tdanet_stream_t stream) {
  TD_REQUIRE(G && A && dW && rows > 0 && N > 0 && K > 0, "bad argument");
  TD_REQUIRE(gemm_mode >= TDANET_GEMM_FP32 && gemm_mode <= TDANET_GEMM_TF32X3, "gemm_mode %d", gemm_mode);
  return launch_wgrad(G, A, dW, db, rows, N, K, nullptr, (cudaStream_t)stream, gemm_mode);
}

#ifdef TD_EMU
const char* tdanet_last_error(void) { return g_err; }
#endif

}  // extern "C"
