// Backward kernels of the TDANet separation path (TDANetBest), channels-last [B, L, C] like the forward.
//
// What they differentiate (reference lines, all under look2hear/models/TDANet_best.py unless noted):
//   GlobLN                         :47-64     gln_bwd_stats / GradLoad (normalise-on-load, backward form)
//   depthwise Conv1d k5/k1 (+bias) :128-156   dw_bwd_kernel (data + weight + bias gradients)
//   LA gate fusion                 :277-292   la_bwd_a_kernel (sigmoid gate, nearest up-sampling, three GlobLN)
//   adaptive_avg_pool1d sum        :358-364   pool_bwd_kernel
//   LayerNorm / MultiheadAttention :236-252   ln_bwd_*, att_bwd_*
//   1x1 Conv1d / Linear weights      wgrad_kernel, colsum_kernel
//   concat_block, mask, decoder, encoder :388-398, :505-518, :497   concat_bwd, mask_bwd, dec_bwd_data, framed_wgrad
//
// Rule mirrored from the forward: a GlobLN is never applied by a kernel of its own.  For the backward
// pass that means: whoever produces dY (the gradient w.r.t. a GlobLN *output*) also accumulates
//   S1_b = sum gamma_c*dY,  S2_b = sum gamma_c*dY*xhat   (per item, double)   and   dgamma_c, dbeta_c,
// and whoever consumes the gradient w.r.t. the GlobLN *input* forms it on load (GradLoad):
//   dX = r_b * (gamma_c*dY - S1_b/N - xhat*S2_b/N),   xhat = (X - mu_b) * r_b.
//
// The file is written in a subset of CUDA that the CPU emulation shim (emu.h, TD_EMU) can execute, so that
// gradients are checked against autograd of the oracle before they reach a GPU.
#pragma once
#include "kernels.h"

namespace td {

// ----------------------------------------------------------------------------- gradient sources
enum GradKind {
  G_PLAIN = 0,  // dy
  G_GLN = 1,    // gradient through a GlobLN, formed on load from (dy, raw x, forward statistics, S1/S2)
  G_RELU = 2    // dy * [x > 0]   (x: the forward tensor after the ReLU)
};

struct GradSrc {
  const float* dy;  // [B, L, C]
  const float* x;   // G_GLN: raw GlobLN input; G_RELU: forward output
  NormRef norm;     // G_GLN: forward statistics + gamma
  const double* S;  // G_GLN: [B, 2]
  int kind;
  int x_bf16;       // x is a stored activation in bf16 (training with act_dtype bf16); dy is always fp32
};

// BF: `x` is a stored activation in bf16 (act_dtype bf16) - a compile-time property of the instantiation, chosen by
// the launcher from the x_bf16 / bf16 flags of the sources; the fp32 instantiations carry no trace of it.
template <int V, bool BF = false>
struct GradLoad {
  vf<V> gr;
  float r, mur, k1, k2;
  const float* dy;
  const float* x;   // start of the tensor (typed access: act_vload)
  size_t x0;        // element offset of (item b, channel ch)
  int kind, xbf;
  __device__ __forceinline__ void init(const GradSrc& g, int b, int ch, size_t item_elems) {
    kind = g.kind;
    dy = g.dy + (size_t)b * item_elems + ch;
    x = g.x ? g.x : g.dy;  // always loadable: loads are issued unconditionally
    xbf = g.x ? g.x_bf16 : 0;
    x0 = (size_t)b * item_elems + ch;
    r = 1.f; mur = 0.f; k1 = 0.f; k2 = 0.f;
    gr = vzero<V>();
    if (kind == G_GLN) {
      norm_moments(g.norm, b, r, mur);
      const vf<V> gam = vload<V>(g.norm.gamma + ch);
#pragma unroll
      for (int e = 0; e < V; ++e) gr[e] = gam[e] * r;
      const double inv = 1.0 / g.norm.count;
      k1 = (float)((double)r * g.S[2 * b] * inv);
      k2 = (float)((double)r * g.S[2 * b + 1] * inv);
    }
  }
  // off: row * C.  Branch-free (selects on the uniform `kind`) so that callers can issue many loads back to back.
  __device__ __forceinline__ vf<V> load(size_t off) const {
    vf<V> d = vload<V>(dy + off);
    const vf<V> xv = act_vload_t<V, BF>(x, x0 + off);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float xh = fmaf(xv[e], r, -mur);
      const float gln = fmaf(gr[e], d[e], -k1) - xh * k2;
      const float relu = xv[e] > 0.f ? d[e] : 0.f;
      d[e] = kind == G_GLN ? gln : (kind == G_RELU ? relu : d[e]);
    }
    return d;
  }
  // row clamped into [0, L): the load is always issued; rows outside the tensor read as zero
  __device__ __forceinline__ vf<V> load_row(int row, int L, int C) const {
    const int rc = row < 0 ? 0 : (row >= L ? L - 1 : row);
    vf<V> v = load((size_t)rc * C);
    const bool ok = row == rc;
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = ok ? v[e] : 0.f;
    return v;
  }
};

// forward value of a conv input row (normalise-on-load): PLAIN, AFFINE, AFFINE_PRELU
template <int V, bool BF = false>
struct FwdLoad {
  vf<V> sc, sh, gam;
  float slope, r, mur;
  const float* x;   // start of the tensor (typed access: act_vload)
  size_t x0;        // element offset of (item b, channel ch)
  int kind, xbf;
  __device__ __forceinline__ void init(const SrcDesc& s, int kind_, int b, int ch, int C) {
    kind = kind_;
    x = s.x;
    xbf = s.bf16;
    x0 = (size_t)b * s.L * C + ch;
    slope = 1.f; r = 1.f; mur = 0.f;
#pragma unroll
    for (int e = 0; e < V; ++e) { sc[e] = 1.f; sh[e] = 0.f; gam[e] = 0.f; }
    if (kind != SRC_PLAIN) {
      norm_coef<V>(s.norm, b, ch, sc, sh);
      norm_moments(s.norm, b, r, mur);
      gam = vload<V>(s.norm.gamma + ch);
    }
    if (kind == SRC_AFFINE_PRELU) slope = __ldg(s.slope);
  }
  // pre-activation value (GlobLN output; the raw value for PLAIN: sc = 1, sh = 0)
  __device__ __forceinline__ vf<V> pre(size_t off) const {
    vf<V> v = act_vload_t<V, BF>(x, x0 + off);
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
    return v;
  }
  // raw stored value of a row clamped into [0, L)
  __device__ __forceinline__ vf<V> raw_row(int row, int L, int C) const {
    const int rc = row < 0 ? 0 : (row >= L ? L - 1 : row);
    return act_vload_t<V, BF>(x, x0 + (size_t)rc * C);
  }
  // GlobLN output from a raw value; rows outside [0, L) are the conv's zero padding
  __device__ __forceinline__ vf<V> pre_of(const vf<V>& raw, int row, int L) const {
    vf<V> v;
    const bool ok = row >= 0 && row < L;
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = ok ? fmaf(raw[e], sc[e], sh[e]) : 0.f;
    return v;
  }
  // rows outside [0, L) read as zero (the conv pads its input, i.e. the value after the transform)
  __device__ __forceinline__ vf<V> pre_row(int row, int L, int C) const {
    const int rc = row < 0 ? 0 : (row >= L ? L - 1 : row);
    vf<V> v = pre((size_t)rc * C);
    const bool ok = row == rc;
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = ok ? v[e] : 0.f;
    return v;
  }
  __device__ __forceinline__ vf<V> load_row(int row, int L, int C) const {
    vf<V> v = pre_row(row, L, C);
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = preluf_(v[e], slope);  // slope = 1 unless AFFINE_PRELU
    return v;
  }
  __device__ __forceinline__ vf<V> load(size_t off) const {
    vf<V> v = pre(off);
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = preluf_(v[e], slope);
    return v;
  }
};

// ----------------------------------------------------------------------------- GlobLN backward
// statistics pass: dgamma_c += sum dy*xhat, dbeta_c += sum dy, S[b] += (sum gamma*dy, sum gamma*dy*xhat)
template <int V, bool BF = false>
__global__ void gln_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, NormRef norm,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     double* __restrict__ S, int L, int C, int rows_per_thread, int x_bf16) {
  grid_dep_wait();
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const bool active = ch < C;
  double s1 = 0.0, s2 = 0.0;
  if (active) {
    const int t0 = blockIdx.x * rows_per_thread, t1 = min(t0 + rows_per_thread, L);
    float r, mur;
    norm_moments(norm, b, r, mur);
    const vf<V> gam = vload<V>(norm.gamma + ch);
    vf<V> dg = vzero<V>(), db = vzero<V>();
    float a1 = 0.f, a2 = 0.f;
    for (int t = t0; t < t1; t += 4) {
      vf<V> d[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // unconditional, clamped; rows past t1 contribute zero
        const size_t off = ((size_t)b * L + (t + i < L ? t + i : L - 1)) * C + ch;
        d[i] = vload<V>(dy + off);
        xv[i] = act_vload_t<V, BF>(x, off);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = t + i < t1;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float dv = ok ? d[i][e] : 0.f;
          const float xh = fmaf(xv[i][e], r, -mur);
          dg[e] = fmaf(dv, xh, dg[e]);
          db[e] += dv;
          const float gd = gam[e] * dv;
          a1 += gd;
          a2 = fmaf(gd, xh, a2);
        }
      }
    }
    vred_add<V>(dgamma + ch, dg);
    vred_add<V>(dbeta + ch, db);
    s1 = a1;
    s2 = a2;
  }
  block_accum2(S + 2 * b, s1, s2);
}

// out = dX formed on load (out must not alias g.dy: the source is read through the read-only path); accumulate: out += dX
template <int V, bool BF = false>
__global__ void gln_bwd_apply_kernel(GradSrc g, float* __restrict__ out, int accumulate, int L, int C,
                                     int rows_per_thread) {
  grid_dep_wait();
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int t0 = blockIdx.x * rows_per_thread, t1 = min(t0 + rows_per_thread, L);
  GradLoad<V, BF> gl;
  gl.init(g, b, ch, (size_t)L * C);
  float* op = out + (size_t)b * L * C + ch;
  for (int t = t0; t < t1; t += 4) {
    vf<V> v[4], o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t off = (size_t)(t + i < L ? t + i : L - 1) * C;
      v[i] = gl.load(off);
      if (accumulate) o[i] = vload_rw<V>(op + off);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t + i < t1) {
        if (accumulate) {
#pragma unroll
          for (int e = 0; e < V; ++e) v[i][e] += o[i][e];
        }
        vstore<V>(op + (size_t)(t + i) * C, v[i]);
      }
    }
  }
}

// y = GlobLN(x) materialised (operand of a weight-gradient GEMM), any C
__global__ void gln_fwd_apply_kernel(const float* __restrict__ x, NormRef norm, float* __restrict__ y, int L, int C) {
  grid_dep_wait();
  const int b = blockIdx.z;
  float r, mur;
  norm_moments(norm, b, r, mur);
  const size_t n = (size_t)L * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float g = __ldg(norm.gamma + c);
    y[(size_t)b * n + i] = fmaf(x[(size_t)b * n + i], g * r, fmaf(-g, mur, __ldg(norm.beta + c)));
  }
}

// weight of F.adaptive_avg_pool1d bin j = [floor(j*L/Lb), ceil((j+1)*L/Lb)) for input row t: 1/|bin| if t is in it
__device__ __forceinline__ float pool_bin_weight(int j, int t, int L, int Lb) {
  if (j >= Lb) return 0.f;
  // 32-bit arithmetic (the launchers require L * Lb < 2^31): a 64-bit division costs ~100 instructions, and this
  // runs several times per row (it was most of the instructions of the spp_dw backward)
  const unsigned lo = ((unsigned)j * (unsigned)L) / (unsigned)Lb;
  const unsigned hi = (((unsigned)j + 1u) * (unsigned)L + (unsigned)Lb - 1u) / (unsigned)Lb;
  return ((unsigned)t >= lo && (unsigned)t < hi) ? 1.f / (float)(hi - lo) : 0.f;
}

// ----------------------------------------------------------------------------- depthwise conv backward
// Forward: out_g[to] = sum_tap w_g[c, tap] * xin[to*stride + tap - PAD] (+ bias_g), g < NW convs sharing xin.
//   dx[ti]    = sum_g sum_tap w_g[c, tap] * G_g[(ti + PAD - tap) / stride]
//   dw_g[tap] += sum_to G_g[to] * xin[to*stride + tap - PAD];   db_g += sum_to G_g[to]
// xin is read through its forward on-load transform; for SRC_AFFINE(_PRELU) the stored dx is the gradient
// w.r.t. the GlobLN output (PReLU derivative applied, slope gradient accumulated).
struct DwBwdArgs {
  GradSrc g[2];
  const float* w[2];
  float* dw[2];
  float* db[2];
  SrcDesc xin;
  int xkind;
  int B, C, Lin, Lout, stride;
  // dw / db point at replica 0 of an accumulator that exists `n_rep` times (`rep_stride` floats apart): CTAs spread
  // their atomics over the replicas (every depthwise weight lives in ~80 cache lines, and a launch adds ~10^6
  // values to them), fold_replicas_kernel sums them into the real gradient at the end of the backward pass
  size_t rep_stride;
  int n_rep;       // power of two
  float* dx;       // [B, Lin, C]
  int accumulate;  // dx += instead of =
  float* dslope;   // SRC_AFFINE_PRELU
  // optional: this launch completes dx (the gradient w.r.t. the GlobLN output xin is read through), so it also
  // accumulates that GlobLN's backward sums instead of a separate statistics pass
  float *up_dgamma, *up_dbeta;
  double* up_S;
  // optional: xin is also an input of the adaptive average pooling into the bottom scale; the pooling's gradient
  // (gathered from pool_g [B, pool_Lb, C]) is added to dx here instead of by a pass of its own
  const float* pool_g;
  int pool_Lb;
  int rows_per_thread;
};

// Streaming form: a thread owns 4 channels and a run of output rows, walks it in tiles of R = 4 output rows and
// keeps the gradient rows / input rows a tile needs in registers (the rows shared with the previous tile are
// carried over), so every row is fetched once and all loads of a tile are issued back to back.
//   stride 1: tile t..t+R-1 needs G[t-PAD .. t+R-1+PAD] and xin[t-PAD .. t+R-1+PAD]               (carry 2*PAD rows)
//   stride 2: output tile t..t+R-1 = input rows 2t..2t+2R-1 needs G[t-1 .. t+R], xin[2t-2 .. 2t+2R]  (carry 2 / 3 rows)
// EXTRA compiles in the optional epilogues (PReLU derivative + slope gradient, upstream GlobLN sums, pooling gradient);
// the plain variants (LA branches) stay lean in registers.
// V = 4 channels per thread, or 2 for small launches (training batches): twice the warps at about half the registers,
// where the kernel is bound by latency at low occupancy rather than by bytes (rows are still read as whole sectors).
#ifdef TD_BWD_MINB   // experiment: force >= TD_BWD_MINB CTAs of 128 threads per SM on the streaming backward kernels
#define TD_BWD_BOUNDS __launch_bounds__(128, TD_BWD_MINB)
#else
#define TD_BWD_BOUNDS
#endif
// GBF / XBF: the raw GlobLN inputs of the gradient sources (g[*].x) / the conv input (xin) are bf16 activations
template <int KS, int NW, int STRIDE, bool EXTRA, int V = 4, bool GBF = false, bool XBF = false>
__global__ void TD_BWD_BOUNDS dw_bwd_kernel(DwBwdArgs a) {
  grid_dep_wait();
  // output rows per tile: 4, or 2 where the windows are wide (two convs, or stride 2 with its 2x input rows)
  constexpr int PAD = (KS - 1) / 2, R = (NW == 2 && KS == 5) || STRIDE == 2 ? 2 : 4;
  constexpr int GW = STRIDE == 1 ? R + 2 * PAD : R + 2;       // gradient rows held per tile
  constexpr int XW = STRIDE == 1 ? R + 2 * PAD : 2 * R + 3;   // input rows held per tile
  constexpr int GC = GW - R, XC = XW - R * STRIDE;            // rows carried from the previous tile
  constexpr int G0 = STRIDE == 1 ? -PAD : -1, X0 = -PAD;      // row of slot 0 relative to the tile base (t resp. STRIDE*t)
  static_assert(STRIDE == 1 || KS == 5, "stride 2 is implemented for k = 5");
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const bool active = ch < a.C;
  double dsl = 0.0, us[2] = {0.0, 0.0};
  if (active) {
    const int C = a.C, Lin = a.Lin, Lout = a.Lout;
    const int o0 = blockIdx.x * a.rows_per_thread, o1 = min(o0 + a.rows_per_thread, Lout);
    const int i1 = o1 == Lout ? Lin : min(o1 * STRIDE, Lin);
    GradLoad<V, GBF> gl[NW];
    float w[NW][KS][V];
#pragma unroll
    for (int g = 0; g < NW; ++g) {
      gl[g].init(a.g[g], b, ch, (size_t)Lout * C);
#pragma unroll
      for (int e = 0; e < V; ++e)
#pragma unroll
        for (int k = 0; k < KS; ++k) w[g][k][e] = __ldg(a.w[g] + (size_t)(ch + e) * KS + k);
    }
    FwdLoad<V, XBF> fx;
    fx.init(a.xin, a.xkind, b, ch, C);
    const bool prelu = EXTRA && a.xkind == SRC_AFFINE_PRELU;
    float* dxp = a.dx + (size_t)b * Lin * C + ch;
    float dw[NW][KS][V], db[NW][V];
#pragma unroll
    for (int g = 0; g < NW; ++g)
#pragma unroll
      for (int e = 0; e < V; ++e) {
        db[g][e] = 0.f;
#pragma unroll
        for (int k = 0; k < KS; ++k) dw[g][k][e] = 0.f;
      }
    float sl_acc = 0.f;
    vf<V> G[NW][GW], X[XW];  // X holds the stored (raw) rows; the GlobLN / PReLU are applied at the point of use
    auto loadG = [&](int g, int row) { return gl[g].load_row(row, Lout, C); };
    auto loadX = [&](int row) { return fx.raw_row(row, Lin, C); };
    const bool up = EXTRA && a.up_S != nullptr;
    const bool pool = EXTRA && a.pool_g != nullptr;
    vf<V> udg = vzero<V>(), udb = vzero<V>();
    float us1 = 0.f, us2 = 0.f;
    // rows carried into the first tile
#pragma unroll
    for (int j = 0; j < GC; ++j)
#pragma unroll
      for (int g = 0; g < NW; ++g) G[g][j] = loadG(g, o0 + G0 + j);
#pragma unroll
    for (int j = 0; j < XC; ++j) X[j] = loadX(o0 * STRIDE + X0 + j);
    for (int t = o0; t < o1; t += R) {
#pragma unroll
      for (int j = GC; j < GW; ++j)
#pragma unroll
        for (int g = 0; g < NW; ++g) G[g][j] = loadG(g, t + G0 + j);
#pragma unroll
      for (int j = XC; j < XW; ++j) X[j] = loadX(t * STRIDE + X0 + j);
      vf<V> O[R * STRIDE];  // previous content of dx (accumulate mode), fetched with the other loads of the tile
      if (a.accumulate) {
#pragma unroll
        for (int q = 0; q < R * STRIDE; ++q) {
          const int ti = t * STRIDE + q;
          O[q] = vload_rw<V>(dxp + (size_t)(ti < Lin ? ti : Lin - 1) * C);
        }
      }
      if (pool) {
        // d/dx of sum_j mean_{t in bin(j)} x[t]: only bins jc = floor(t*Lb/L) and jc + 1 can contain t
        const int Lb = a.pool_Lb;
        const float* gp = a.pool_g + (size_t)b * Lb * C + ch;
#pragma unroll
        for (int q = 0; q < R * STRIDE; ++q) {
          const int ti = min(t * STRIDE + q, Lin - 1);
          const int jc = (int)(((unsigned)ti * (unsigned)Lb) / (unsigned)Lin);
          const vf<V> g0 = vload<V>(gp + (size_t)jc * C), g1 = vload<V>(gp + (size_t)min(jc + 1, Lb - 1) * C);
          const float w0 = pool_bin_weight(jc, ti, Lin, Lb), w1 = pool_bin_weight(jc + 1, ti, Lin, Lb);
          if (!a.accumulate) O[q] = vzero<V>();
#pragma unroll
          for (int e = 0; e < V; ++e) O[q][e] += fmaf(g0[e], w0, g1[e] * w1);
        }
      }
      const bool add_o = a.accumulate || pool;
      // ---- data gradient of the input rows of this tile
#pragma unroll
      for (int q = 0; q < R * STRIDE; ++q) {
        const int ti = t * STRIDE + q;
        if (ti < i1) {
          vf<V> acc = vzero<V>();
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            // output row (ti + PAD - k) / STRIDE
            if (STRIDE == 2 && ((q + PAD - k) & 1)) continue;
            const int slot = STRIDE == 1 ? q + 2 * PAD - k : (q + PAD - k) / 2 + 1;
#pragma unroll
            for (int g = 0; g < NW; ++g)
#pragma unroll
              for (int e = 0; e < V; ++e) acc[e] = fmaf(w[g][k][e], G[g][slot][e], acc[e]);
          }
          if (prelu) {
            const vf<V> n = fx.pre_of(X[q - X0], ti, Lin);
#pragma unroll
            for (int e = 0; e < V; ++e)
              if (n[e] < 0.f) {
                sl_acc = fmaf(acc[e], n[e], sl_acc);
                acc[e] *= fx.slope;
              }
          }
          if (add_o) {
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] += O[q][e];
          }
          vstore<V>(dxp + (size_t)ti * C, acc);
          if (up) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const float xh = fmaf(X[q - X0][e], fx.r, -fx.mur);
              udg[e] = fmaf(acc[e], xh, udg[e]);
              udb[e] += acc[e];
              const float gd = fx.gam[e] * acc[e];
              us1 += gd;
              us2 = fmaf(gd, xh, us2);
            }
          }
        }
      }
      // ---- weight / bias gradients of the output rows of this tile
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (t + r < o1) {
#pragma unroll
          for (int g = 0; g < NW; ++g)
#pragma unroll
            for (int e = 0; e < V; ++e) db[g][e] += G[g][r - G0][e];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            vf<V> xv = fx.pre_of(X[r * STRIDE + k], (t + r) * STRIDE + k - PAD, Lin);
#pragma unroll
            for (int e = 0; e < V; ++e) xv[e] = preluf_(xv[e], fx.slope);  // slope = 1 unless AFFINE_PRELU
#pragma unroll
            for (int g = 0; g < NW; ++g)
#pragma unroll
              for (int e = 0; e < V; ++e) dw[g][k][e] = fmaf(G[g][r - G0][e], xv[e], dw[g][k][e]);
          }
        }
      }
      // ---- carry
#pragma unroll
      for (int j = 0; j < GC; ++j)
#pragma unroll
        for (int g = 0; g < NW; ++g) G[g][j] = G[g][j + R];
#pragma unroll
      for (int j = 0; j < XC; ++j) X[j] = X[j + R * STRIDE];
    }
    dsl = sl_acc;
    if (up) {
      vred_add<V>(a.up_dgamma + ch, udg);
      vred_add<V>(a.up_dbeta + ch, udb);
      us[0] = us1;
      us[1] = us2;
    }
    // the thread's V*KS weight gradients are contiguous in [C][KS]: KS vector reductions per conv
    const size_t roff = (size_t)((blockIdx.x + blockIdx.z) & (a.n_rep - 1)) * a.rep_stride;
#pragma unroll
    for (int g = 0; g < NW; ++g) {
#pragma unroll
      for (int q = 0; q < KS; ++q) {
        vf<V> v;
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = dw[g][(q * V + i) % KS][(q * V + i) / KS];
        vred_add<V>(a.dw[g] + roff + (size_t)ch * KS + q * V, v);
      }
      if (a.db[g]) {
        vf<V> v;
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = db[g][e];
        vred_add<V>(a.db[g] + roff + ch, v);
      }
    }
  }
  if (EXTRA && a.up_S) block_accum2(a.up_S + 2 * b, us[0], us[1]);  // uniform across the grid
  if (EXTRA && a.dslope) {  // uniform across the grid
    __shared__ double sh[64];
    double z = 0.0;
#ifdef TD_EMU
    if (!emu::bs) { atomicAdd(a.dslope, (float)dsl); return; }
#endif
    block_sum2(dsl, z, sh);
    if (threadIdx.x == 0) atomicAdd(a.dslope, (float)dsl);
  }
}

// grad[i] += sum_r rep[r][i] for every registered accumulator
struct FoldArgs {
  struct Entry { float* dst; const float* rep; int n; } e[64];
  int count, n_rep;
  size_t rep_stride;
};
__global__ void fold_replicas_kernel(FoldArgs a) {
  grid_dep_wait();
  const FoldArgs::Entry& en = a.e[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < en.n; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < a.n_rep; ++r) acc += en.rep[(size_t)r * a.rep_stride + i];
    en.dst[i] += acc;
  }
}

// ----------------------------------------------------------------------------- LA backward, phase A
// Forward (TDANet_best.py:277-292):  A = dw_l(xl), Bt = dw_a(xg), E = dw_e(xg);
//   out[t] = gLN_L(A)[t] * sigmoid(gLN_A(Bt)[j]) + gLN_E(E)[j],   j = nearest(t; Lg -> Ll).
// This phase recomputes the three raw conv outputs, writes them and the gradients w.r.t. the three GlobLN
// outputs, and accumulates the GlobLN backward sums; dw_bwd_kernel finishes the convolutions.
struct LaBwdArgs {
  SrcDesc loc;
  int lkind;         // SRC_PLAIN (x_fused) or SRC_AFFINE (GlobLN(spp_dw[k]) for loc_glo_fus)
  const float* glo;  // [B, Lg, C]
  int glo_bf16;      // glo is a stored activation in bf16
  int Lg, B, C;
  const float *wl, *wa, *we;
  NormRef nL, nA, nE;
  const float* dout;  // [B, Ll, C]
  float scale;        // fl32(Lg / Ll)
  float *d_loc, *raw_a;                  // [B, Ll, C]
  float *d_act, *d_emb, *raw_b, *raw_e;  // [B, Lg, C]
  float *dgamma[3], *dbeta[3];           // local, act, embedding GlobLN
  double* S[3];
  int jchunk;
};

// first local row whose nearest source row is >= j
__device__ __forceinline__ int first_local_row(int j, float scale, int Ll, int Lg) {
  if (j <= 0) return 0;
  if (j >= Lg) return Ll;
  int t = (int)ceilf((float)j / scale);
  t = t < 0 ? 0 : (t > Ll ? Ll : t);
  while (t > 0 && nearest_src(t - 1, scale, Lg) >= j) --t;
  while (t < Ll && nearest_src(t, scale, Lg) < j) ++t;
  return t;
}

// Two regular streaming passes (each thread: 4 channels, tiles of 4 rows, all loads of a tile issued together):
//   G  over the global rows : raw_b = dw_a(xg), raw_e = dw_e(xg)
//   L  over the local rows  : raw_a = dw_l(xl), d_loc = dout * gate[j], GlobLN_L sums; a thread owns whole centres,
//                             so when it leaves centre j it writes d_act[j] = (sum dout*loc) * gate*(1-gate),
//                             d_emb[j] = sum dout and adds them to the GlobLN_A / GlobLN_E sums
template <int KS, bool BF = false>
__global__ void la_bwd_g_kernel(LaBwdArgs a, int rows_per_thread) {
  grid_dep_wait();
  constexpr int V = 4, PAD = (KS - 1) / 2, R = 4, W = R + 2 * PAD;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= a.C) return;
  const int Lg = a.Lg, C = a.C;
  const int j0 = blockIdx.x * rows_per_thread, j1 = min(j0 + rows_per_thread, Lg);
  float wa[KS][V], we[KS][V];
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      wa[k][e] = __ldg(a.wa + (size_t)(ch + e) * KS + k);
      we[k][e] = __ldg(a.we + (size_t)(ch + e) * KS + k);
    }
  const size_t goff = (size_t)b * Lg * C + ch;
  auto loadg = [&](int row) {
    const int rc = row < 0 ? 0 : (row >= Lg ? Lg - 1 : row);
    vf<V> v = act_vload_t<V, BF>(a.glo, goff + (size_t)rc * C);
    const bool ok = row == rc;
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = ok ? v[e] : 0.f;
    return v;
  };
  vf<V> X[W];
#pragma unroll
  for (int j = 0; j < 2 * PAD; ++j) X[j] = loadg(j0 - PAD + j);
  for (int t = j0; t < j1; t += R) {
#pragma unroll
    for (int j = 2 * PAD; j < W; ++j) X[j] = loadg(t - PAD + j);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (t + r < j1) {
        vf<V> Bt = vzero<V>(), Et = vzero<V>();
#pragma unroll
        for (int k = 0; k < KS; ++k)
#pragma unroll
          for (int e = 0; e < V; ++e) {
            Bt[e] = fmaf(wa[k][e], X[r + k][e], Bt[e]);
            Et[e] = fmaf(we[k][e], X[r + k][e], Et[e]);
          }
        vstore<V>(a.raw_b + goff + (size_t)(t + r) * C, Bt);
        vstore<V>(a.raw_e + goff + (size_t)(t + r) * C, Et);
      }
    }
#pragma unroll
    for (int j = 0; j < 2 * PAD; ++j) X[j] = X[j + R];
  }
}

template <int KS, bool BF = false>
__global__ void TD_BWD_BOUNDS la_bwd_l_kernel(LaBwdArgs a) {
  grid_dep_wait();
  constexpr int V = 4, PAD = (KS - 1) / 2, R = 4, W = R + 2 * PAD;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const bool active = ch < a.C;
  double S[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
  if (active) {
    const int Ll = a.loc.L, Lg = a.Lg, C = a.C;
    const int j0 = blockIdx.x * a.jchunk, j1 = min(j0 + a.jchunk, Lg);
    const int t_begin = first_local_row(j0, a.scale, Ll, Lg), t_end = first_local_row(j1, a.scale, Ll, Lg);
    float wl[KS][V];
#pragma unroll
    for (int e = 0; e < V; ++e)
#pragma unroll
      for (int k = 0; k < KS; ++k) wl[k][e] = __ldg(a.wl + (size_t)(ch + e) * KS + k);
    vf<V> scL, shL, scA, shA;
    norm_coef<V>(a.nL, b, ch, scL, shL);
    norm_coef<V>(a.nA, b, ch, scA, shA);
    float rL, murL, rA, murA, rE, murE;
    norm_moments(a.nL, b, rL, murL);
    norm_moments(a.nA, b, rA, murA);
    norm_moments(a.nE, b, rE, murE);
    const vf<V> gL = vload<V>(a.nL.gamma + ch), gA = vload<V>(a.nA.gamma + ch), gE = vload<V>(a.nE.gamma + ch);
    FwdLoad<V, BF> fl;
    fl.init(a.loc, a.lkind, b, ch, C);
    const float* dop = a.dout + (size_t)b * Ll * C + ch;
    float* dlp = a.d_loc + (size_t)b * Ll * C + ch;
    float* rap = a.raw_a + (size_t)b * Ll * C + ch;
    const size_t goff = (size_t)b * Lg * C + ch;
    const float* rbp = a.raw_b + goff;
    const float* rep_ = a.raw_e + goff;
    // zero padding applies to the conv input, i.e. after the on-load transform
    auto loadx = [&](int row) { return fl.load_row(row, Ll, C); };
    vf<V> dg = vzero<V>(), db = vzero<V>(), sum_e = vzero<V>(), sum_a = vzero<V>();
    vf<V> dgA = vzero<V>(), dbA = vzero<V>(), dgE = vzero<V>(), dbE = vzero<V>();
    float s1 = 0.f, s2 = 0.f, s1A = 0.f, s2A = 0.f, s1E = 0.f, s2E = 0.f;
    // centres of this chunk that no local row maps to (down-sampling step) get zero sums
    auto zero_centres = [&](int ja, int jb) {
      for (int j = ja; j < jb; ++j) {
        vstore<V>(a.d_act + goff + (size_t)j * C, vzero<V>());
        vstore<V>(a.d_emb + goff + (size_t)j * C, vzero<V>());
      }
    };
    if (t_begin >= t_end) {
      zero_centres(j0, j1);
    } else {
      zero_centres(j0, nearest_src(t_begin, a.scale, Lg));
      vf<V> X[W];
#pragma unroll
      for (int j = 0; j < 2 * PAD; ++j) X[j] = loadx(t_begin - PAD + j);
      for (int t = t_begin; t < t_end; t += R) {
        vf<V> D[R], Bt[R], Et[R];
        int jr[R + 1];
#pragma unroll
        for (int j = 2 * PAD; j < W; ++j) X[j] = loadx(t - PAD + j);
#pragma unroll
        for (int r = 0; r <= R; ++r) jr[r] = t + r < t_end ? nearest_src(t + r, a.scale, Lg) : j1;
#pragma unroll
        for (int r = 0; r < R; ++r) {  // unconditional, clamped (rows past t_end are not used)
          const int tr = t + r < Ll ? t + r : Ll - 1;
          const int jc = jr[r] < Lg ? jr[r] : Lg - 1;
          D[r] = vload<V>(dop + (size_t)tr * C);
          Bt[r] = vload_rw<V>(rbp + (size_t)jc * C);
          Et[r] = vload_rw<V>(rep_ + (size_t)jc * C);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (t + r < t_end) {
            vf<V> A = vzero<V>(), dl, gate4;
#pragma unroll
            for (int k = 0; k < KS; ++k)
#pragma unroll
              for (int e = 0; e < V; ++e) A[e] = fmaf(wl[k][e], X[r + k][e], A[e]);
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const float gate = sigmoidf_(fmaf(Bt[r][e], scA[e], shA[e]));
              gate4[e] = gate;
              const float loc = fmaf(A[e], scL[e], shL[e]);
              sum_e[e] += D[r][e];
              sum_a[e] = fmaf(D[r][e], loc, sum_a[e]);
              dl[e] = D[r][e] * gate;
              const float xh = fmaf(A[e], rL, -murL);
              dg[e] = fmaf(dl[e], xh, dg[e]);
              db[e] += dl[e];
              const float gd = gL[e] * dl[e];
              s1 += gd;
              s2 = fmaf(gd, xh, s2);
            }
            vstore<V>(dlp + (size_t)(t + r) * C, dl);
            vstore<V>(rap + (size_t)(t + r) * C, A);
            if (jr[r + 1] != jr[r]) {  // last local row of centre jr[r]: its sums are complete
              vf<V> da;
#pragma unroll
              for (int e = 0; e < V; ++e) {
                da[e] = sum_a[e] * gate4[e] * (1.f - gate4[e]);  // through the sigmoid
                const float bh = fmaf(Bt[r][e], rA, -murA), eh = fmaf(Et[r][e], rE, -murE);
                dgA[e] = fmaf(da[e], bh, dgA[e]);
                dbA[e] += da[e];
                const float ga = gA[e] * da[e];
                s1A += ga;
                s2A = fmaf(ga, bh, s2A);
                dgE[e] = fmaf(sum_e[e], eh, dgE[e]);
                dbE[e] += sum_e[e];
                const float ge = gE[e] * sum_e[e];
                s1E += ge;
                s2E = fmaf(ge, eh, s2E);
              }
              vstore<V>(a.d_act + goff + (size_t)jr[r] * C, da);
              vstore<V>(a.d_emb + goff + (size_t)jr[r] * C, sum_e);
              sum_a = vzero<V>();
              sum_e = vzero<V>();
              zero_centres(jr[r] + 1, jr[r + 1]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 2 * PAD; ++j) X[j] = X[j + R];
      }
    }
    vred_add<V>(a.dgamma[0] + ch, dg);
    vred_add<V>(a.dbeta[0] + ch, db);
    vred_add<V>(a.dgamma[1] + ch, dgA);
    vred_add<V>(a.dbeta[1] + ch, dbA);
    vred_add<V>(a.dgamma[2] + ch, dgE);
    vred_add<V>(a.dbeta[2] + ch, dbE);
    S[0][0] = s1; S[0][1] = s2; S[1][0] = s1A; S[1][1] = s2A; S[2][0] = s1E; S[2][1] = s2E;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) block_accum2(a.S[i] + 2 * b, S[i][0], S[i][1]);
}

// ----------------------------------------------------------------------------- FORK: additive injection backward
// x_fused[k][t] = n_k[t] + g[nearest(t)]  (TDANet.py:622-626):  d n_k = d x_fused[k] (same buffer),
// dg[j] (+)= sum_{t : nearest(t) = j} d x_fused[k][t].  A thread owns whole centres j.
__global__ void inject_add_bwd_kernel(const float* __restrict__ dfused, float* __restrict__ dg, int accumulate, int Ll,
                                      int Lg, int C, float scale, int jchunk) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int j0 = blockIdx.x * jchunk, j1 = min(j0 + jchunk, Lg);
  const float* dp = dfused + (size_t)b * Ll * C + ch;
  float* gp = dg + (size_t)b * Lg * C + ch;
  int t = first_local_row(j0, scale, Ll, Lg);
  for (int j = j0; j < j1; ++j) {
    const int tend = first_local_row(j + 1, scale, Ll, Lg);
    vf<V> acc = accumulate ? vload_rw<V>(gp + (size_t)j * C) : vzero<V>();
    for (; t + 4 <= tend; t += 4) {
      vf<V> d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = vload<V>(dp + (size_t)(t + i) * C);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += d[i][e];
    }
    for (; t < tend; ++t) {
      const vf<V> d = vload<V>(dp + (size_t)t * C);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += d[e];
    }
    vstore<V>(gp + (size_t)j * C, acc);
  }
}

// ----------------------------------------------------------------------------- FORK: conv_pool depthwise backward
// Forward (TDANet.py:190-228, 560-569): out[to] = sum_tap w[c, tap] * xin[to*s + tap - pad] + bias, any odd ks, stride s.
//   dx[ti] (+)= sum_{to} w[c, ti + pad - to*s] * G[to]      (taps inside [0, ks))
__global__ void dwg_bwd_data_kernel(const float* __restrict__ G, const float* __restrict__ w, float* __restrict__ dx,
                                    int accumulate, int Lin, int Lout, int C, int ks, int stride, int rows_per_thread) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int pad = (ks - 1) / 2;
  const int t0 = blockIdx.x * rows_per_thread, t1 = min(t0 + rows_per_thread, Lin);
  const float* gp = G + (size_t)b * Lout * C + ch;
  float* dp = dx + (size_t)b * Lin * C + ch;
  for (int ti = t0; ti < t1; ++ti) {
    vf<V> acc = accumulate ? vload_rw<V>(dp + (size_t)ti * C) : vzero<V>();
    // to*s <= ti + pad  and  ti + pad - to*s < ks
    int to_hi = (ti + pad) / stride;
    if (to_hi > Lout - 1) to_hi = Lout - 1;
    int to_lo = ti + pad - ks + 1;
    to_lo = to_lo <= 0 ? 0 : (to_lo + stride - 1) / stride;
    for (int to = to_lo; to <= to_hi; ++to) {
      const int tap = ti + pad - to * stride;
      const vf<V> g = vload<V>(gp + (size_t)to * C);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = fmaf(__ldg(w + (size_t)(ch + e) * ks + tap), g[e], acc[e]);
    }
    vstore<V>(dp + (size_t)ti * C, acc);
  }
}

//   dw[c, tap] += sum_{b, to} G[b, to, c] * xin[b, to*s + tap - pad, c];   db[c] += sum G
// One thread per (4 channels, tap, batch item, chunk of output rows): blockIdx.y = tap (ks taps + one extra row of
// blocks for the bias), blockIdx.z = item * chunks + chunk.  Rows in tiles of 4 with every load issued before use.
// (The first version walked all B * Lout rows in one thread per (tap, 4 channels): 34 CTAs, ~0.5 ms per launch,
// 50.8 of the 68.7 ms of a fork training step.)
template <bool BF = false>
__global__ void dwg_bwd_weight_kernel(const float* __restrict__ G, SrcDesc xin, int xkind, float* __restrict__ dw,
                                      float* __restrict__ db, int B, int Lout, int C, int ks, int stride,
                                      int rows_per_cta) {
  grid_dep_wait();
  constexpr int V = 4;
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  const int tap = blockIdx.y;
  if (ch >= C) return;
  const int chunks = (Lout + rows_per_cta - 1) / rows_per_cta;
  const int b = blockIdx.z / chunks, to0 = (blockIdx.z % chunks) * rows_per_cta;
  const int to1 = min(to0 + rows_per_cta, Lout);
  const int pad = (ks - 1) / 2, Lin = xin.L;
  vf<V> acc = vzero<V>();
  FwdLoad<V, BF> fx;
  fx.init(xin, xkind, b, ch, C);
  const float* gp = G + (size_t)b * Lout * C + ch;
  for (int to = to0; to < to1; to += 4) {
    vf<V> g[4], xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // unconditional, clamped loads
      const int tc = to + i < Lout ? to + i : Lout - 1;
      g[i] = vload<V>(gp + (size_t)tc * C);
      xv[i] = tap == ks ? g[i] : fx.load_row(tc * stride + tap - pad, Lin, C);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (to + i < to1) {
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = tap == ks ? acc[e] + g[i][e] : fmaf(g[i][e], xv[i][e], acc[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    if (tap == ks) atomicAdd(db + ch + e, acc[e]);
    else atomicAdd(dw + (size_t)(ch + e) * ks + tap, acc[e]);
  }
}

// ----------------------------------------------------------------------------- pooling backward
// ga_in[j] = sum_k mean_{t in bin_k(j)} n_k[t]   =>   dn_k[t] = sum_{j : t in bin_k(j)} g[j] / |bin_k(j)|
__global__ void pool_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int accumulate, int L, int Lb,
                                int C, int rows_per_thread) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int t0 = blockIdx.x * rows_per_thread, t1 = min(t0 + rows_per_thread, L);
  const float* gp = g + (size_t)b * Lb * C + ch;
  float* dp = dx + (size_t)b * L * C + ch;
  // only bins jc = floor(t*Lb/L) and jc + 1 can contain t (bins overlap by at most one row)
  auto bin_weight = [&](int j, int t) {
    if (j >= Lb) return 0.f;
    const unsigned lo = ((unsigned)j * (unsigned)L) / (unsigned)Lb;
    const unsigned hi = (((unsigned)j + 1u) * (unsigned)L + (unsigned)Lb - 1u) / (unsigned)Lb;
    return ((unsigned)t >= lo && (unsigned)t < hi) ? 1.f / (float)(hi - lo) : 0.f;
  };
  for (int t = t0; t < t1; t += 4) {
    vf<V> g0[4], g1[4], o[4];
    float w0[4], w1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // unconditional, clamped loads
      const int tt = t + i < L ? t + i : L - 1;
      const int jc = (int)(((unsigned)tt * (unsigned)Lb) / (unsigned)L);
      w0[i] = bin_weight(jc, tt);
      w1[i] = bin_weight(jc + 1, tt);
      g0[i] = vload<V>(gp + (size_t)jc * C);
      g1[i] = vload<V>(gp + (size_t)(jc + 1 < Lb ? jc + 1 : Lb - 1) * C);
      if (accumulate) o[i] = vload_rw<V>(dp + (size_t)tt * C);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t + i < t1) {
        vf<V> acc;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          acc[e] = fmaf(g0[i][e], w0[i], g1[i][e] * w1[i]);
          if (accumulate) acc[e] += o[i][e];
        }
        vstore<V>(dp + (size_t)(t + i) * C, acc);
      }
    }
  }
}

// ----------------------------------------------------------------------------- LayerNorm backward
// y = LN(k1*x)*w + b per row.  Row pass (one warp per row): rowstat[row] = {mu, rstd, mean(w*dy), mean(w*dy*xhat)}
__global__ void ln_bwd_rows_kernel(const float* __restrict__ x, float k1, const float* __restrict__ w,
                                   const float* __restrict__ dy, float* __restrict__ rowstat, int rows, int C) {
  grid_dep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  const float* dr = dy + (size_t)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += k1 * xr[c];
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = k1 * xr[c] - mu;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + kEpsLN);
  float m1 = 0.f, m2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float wd = __ldg(w + c) * dr[c];
    m1 += wd;
    m2 = fmaf(wd, (k1 * xr[c] - mu) * rstd, m2);
  }
  m1 = warp_sum(m1) / (float)C;
  m2 = warp_sum(m2) / (float)C;
  if (lane == 0) {
    float* o = rowstat + (size_t)row * 4;
    o[0] = mu; o[1] = rstd; o[2] = m1; o[3] = m2;
  }
}

// out = (add ? add : 0) + kout * rstd*(w*dy - m1 - xhat*m2);  dw_c += sum dy*xhat;  db_c += sum dy
__global__ void ln_bwd_apply_kernel(const float* __restrict__ x, float k1, const float* __restrict__ w,
                                    const float* __restrict__ dy, const float* __restrict__ rowstat,
                                    const float* __restrict__ add, float kout, float* __restrict__ out,
                                    float* __restrict__ dw, float* __restrict__ db, int rows, int C,
                                    int rows_per_thread) {
  grid_dep_wait();
  constexpr int V = 4;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int r0 = blockIdx.x * rows_per_thread, r1 = min(r0 + rows_per_thread, rows);
  const vf<V> wv = vload<V>(w + ch);
  vf<V> aw = vzero<V>(), ab = vzero<V>();
  for (int r = r0; r < r1; r += 4) {
    vf<V> xv[4], d[4], o[4];
    float4 rs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // unconditional, clamped loads
      const int rr = r + i < rows ? r + i : rows - 1;
      const size_t off = (size_t)rr * C + ch;
      rs[i] = __ldg(reinterpret_cast<const float4*>(rowstat) + rr);
      xv[i] = vload<V>(x + off);
      d[i] = vload<V>(dy + off);
      o[i] = add ? vload<V>(add + off) : vzero<V>();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (r + i < r1) {
        const float mu = rs[i].x, rstd = rs[i].y, m1 = rs[i].z, m2 = rs[i].w;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float xh = (k1 * xv[i][e] - mu) * rstd;
          aw[e] = fmaf(d[i][e], xh, aw[e]);
          ab[e] += d[i][e];
          o[i][e] = fmaf(kout * rstd, wv[e] * d[i][e] - m1 - xh * m2, o[i][e]);
        }
        vstore<V>(out + (size_t)(r + i) * C + ch, o[i]);
      }
    }
  }
  vred_add<V>(dw + ch, aw);
  vred_add<V>(db + ch, ab);
}

// ----------------------------------------------------------------------------- dropout multipliers
// out[i] = in[i] * (k0 + k1*mask[i]) * (item_mask ? item_mask[i / per_item]*item_scale : 1); 4 elements per thread.
// Applies nn.Dropout / DropPath keep-masks to an activation (forward) or to its gradient (backward).
__global__ void mask_scale_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n4,
                                  const uint8_t* __restrict__ mask, float k0, float k1,
                                  const uint8_t* __restrict__ item_mask, float item_scale, size_t per_item,
                                  int round_out) {
  grid_dep_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = *reinterpret_cast<const float4*>(in + i * 4);
  float f[4] = {k0 + k1, k0 + k1, k0 + k1, k0 + k1};
  if (mask) {
    const uint32_t mk = *reinterpret_cast<const uint32_t*>(mask + i * 4);
#pragma unroll
    for (int e = 0; e < 4; ++e) f[e] = ((mk >> (8 * e)) & 0xffu) ? k0 + k1 : k0;
  }
  if (item_mask) {
    const float fi = item_mask[(i * 4) / per_item] ? item_scale : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) f[e] *= fi;
  }
  v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
  if (round_out) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
  *reinterpret_cast<float4*>(out + i * 4) = v;
}

// ----------------------------------------------------------------------------- attention backward
// Problem geometry as in attention_kernel (bottom.cu): token(s) = base + s*stride, n tokens per problem.
// Scratch P, dS: [problem, head, query, key].
__device__ __forceinline__ void att_problem(int prob, int L, int group, int time_axis, long& base, long& stride) {
  if (time_axis) {
    base = (long)prob * L;
    stride = 1;
  } else {
    const int grp = prob / L, t = prob % L;
    base = (long)grp * group * L + t;
    stride = L;
  }
}

// one thread per (problem, head, query i): softmax row, dS row, dq
template <int D>
__global__ void att_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx,
                                  float* __restrict__ P, float* __restrict__ dS, float* __restrict__ dqkv,
                                  int L, int C, int n, int n_head, int group, int time_axis, int total,
                                  const uint8_t* __restrict__ amask, float inv_keep) {
  grid_dep_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i = idx % n, head = (idx / n) % n_head, prob = idx / (n * n_head);
  // training dropout on the weights: ctx = (P o M/keep) V, so dP = (dO V^T) o M/keep and dV uses P o M/keep
  const uint8_t* mrow = amask ? amask + ((size_t)(prob * n_head + head) * n + i) * n : nullptr;
  long base, stride;
  att_problem(prob, L, group, time_axis, base, stride);
  const float scale = rsqrtf((float)D);
  const size_t C3 = (size_t)3 * C;
  const float* qp = qkv + (size_t)(base + (long)i * stride) * C3 + head * D;
  const float* dop = dctx + (size_t)(base + (long)i * stride) * C + head * D;
  float* prow = P + ((size_t)(prob * n_head + head) * n + i) * n;
  float* srow = dS + ((size_t)(prob * n_head + head) * n + i) * n;
  // scores (kept in prow), running maximum
  float m = -3.4e38f;
  for (int j = 0; j < n; ++j) {
    const float* kp = qkv + (size_t)(base + (long)j * stride) * C3 + C + head * D;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(qp[d] * scale, kp[d], acc);
    prow[j] = acc;
    m = fmaxf(m, acc);
  }
  float l = 0.f;
  for (int j = 0; j < n; ++j) {
    const float p = expf(prow[j] - m);
    prow[j] = p;
    l += p;
  }
  const float inv = 1.f / l;
  float dsum = 0.f;
  for (int j = 0; j < n; ++j) {
    const float* vp = qkv + (size_t)(base + (long)j * stride) * C3 + 2 * C + head * D;
    float dp = 0.f;
    for (int d = 0; d < D; ++d) dp = fmaf(dop[d], vp[d], dp);
    const float p = prow[j] * inv;
    if (mrow) dp = mrow[j] ? dp * inv_keep : 0.f;
    prow[j] = p;
    srow[j] = dp;
    dsum = fmaf(p, dp, dsum);
  }
  float dq[D];
#pragma unroll
  for (int d = 0; d < D; ++d) dq[d] = 0.f;
  for (int j = 0; j < n; ++j) {
    const float ds = prow[j] * (srow[j] - dsum);
    srow[j] = ds;
    if (mrow) prow[j] = mrow[j] ? prow[j] * inv_keep : 0.f;  // what att_bwd_dkv_kernel needs for dV
    const float* kp = qkv + (size_t)(base + (long)j * stride) * C3 + C + head * D;
#pragma unroll
    for (int d = 0; d < D; ++d) dq[d] = fmaf(ds, kp[d], dq[d]);
  }
  float* o = dqkv + (size_t)(base + (long)i * stride) * C3 + head * D;
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = dq[d] * scale;
}

// one thread per (problem, head, key j): dk, dv
template <int D>
__global__ void att_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx,
                                   const float* __restrict__ P, const float* __restrict__ dS,
                                   float* __restrict__ dqkv, int L, int C, int n, int n_head, int group,
                                   int time_axis, int total) {
  grid_dep_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = idx % n, head = (idx / n) % n_head, prob = idx / (n * n_head);
  long base, stride;
  att_problem(prob, L, group, time_axis, base, stride);
  const float scale = rsqrtf((float)D);
  const size_t C3 = (size_t)3 * C;
  const float* pm = P + (size_t)(prob * n_head + head) * n * n;
  const float* sm = dS + (size_t)(prob * n_head + head) * n * n;
  float dk[D], dv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
  for (int i = 0; i < n; ++i) {
    const float p = pm[(size_t)i * n + j], ds = sm[(size_t)i * n + j] * scale;
    const float* qp = qkv + (size_t)(base + (long)i * stride) * C3 + head * D;
    const float* dop = dctx + (size_t)(base + (long)i * stride) * C + head * D;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      dk[d] = fmaf(ds, qp[d], dk[d]);
      dv[d] = fmaf(p, dop[d], dv[d]);
    }
  }
  float* o = dqkv + (size_t)(base + (long)j * stride) * C3 + C + head * D;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    o[d] = dk[d];
    o[C + d] = dv[d];
  }
}

// One warp per (problem, head) for short sequences (n <= NMAX; training attends over the batch axis, n = 8):
// lanes split the head dimension, q / k / v / dO of the problem live in registers, scores and dP are warp
// reductions, so nothing goes through a scratch buffer and every global load is issued up front.
template <int D, int NMAX>
__global__ void att_bwd_warp_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx,
                                    float* __restrict__ dqkv, int L, int C, int n, int n_head, int group,
                                    int time_axis, int total_warps, const uint8_t* __restrict__ amask,
                                    float inv_keep) {
  grid_dep_wait();
  constexpr int EPL = D >= 32 ? D / 32 : 1;  // head-dim elements per lane
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= total_warps) return;
  const int head = wid % n_head, prob = wid / n_head;
  long base, stride;
  att_problem(prob, L, group, time_axis, base, stride);
  const bool lane_ok = lane * EPL < D;
  const float scale = rsqrtf((float)D);
  const size_t C3 = (size_t)3 * C;
  float q[NMAX][EPL], k[NMAX][EPL], v[NMAX][EPL], dO[NMAX][EPL];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    const bool ok = i < n && lane_ok;
    const size_t tok = (size_t)(base + (long)(i < n ? i : 0) * stride);
    const float* p = qkv + tok * C3 + head * D + (lane_ok ? lane * EPL : 0);
    const float* dp = dctx + tok * C + head * D + (lane_ok ? lane * EPL : 0);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const float qv = p[e], kv = p[C + e], vv = p[2 * C + e], dv = dp[e];
      q[i][e] = ok ? qv * scale : 0.f;
      k[i][e] = ok ? kv : 0.f;
      v[i][e] = ok ? vv : 0.f;
      dO[i][e] = ok ? dv : 0.f;
    }
  }
  float dq[NMAX][EPL], dk[NMAX][EPL], dvv[NMAX][EPL];
#pragma unroll
  for (int i = 0; i < NMAX; ++i)
#pragma unroll
    for (int e = 0; e < EPL; ++e) { dq[i][e] = 0.f; dk[i][e] = 0.f; dvv[i][e] = 0.f; }
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {  // uniform across the warp
      float s[NMAX], dp[NMAX];
      float m = -3.4e38f;
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          a1 = fmaf(q[i][e], k[j][e], a1);
          a2 = fmaf(dO[i][e], v[j][e], a2);
        }
        s[j] = warp_sum(a1);
        dp[j] = warp_sum(a2);
        if (j < n) m = fmaxf(m, s[j]);
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        s[j] = j < n ? expf(s[j] - m) : 0.f;
        l += s[j];
      }
      const float inv = 1.f / l;
      // training dropout on the weights: mk[j] = keep-mask / keep_prob of (query i, key j), else 1
      float mk[NMAX];
#pragma unroll
      for (int j = 0; j < NMAX; ++j)
        mk[j] = amask && j < n ? (amask[((size_t)wid * n + i) * n + j] ? inv_keep : 0.f) : 1.f;
      float dsum = 0.f;
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        s[j] *= inv;  // P[i][j]
        dp[j] *= mk[j];
        dsum = fmaf(s[j], dp[j], dsum);
      }
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        const float ds = s[j] * (dp[j] - dsum);
        const float pm = s[j] * mk[j];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          dq[i][e] = fmaf(ds, k[j][e], dq[i][e]);    // d s_ij / d q_i (scaled q): k_j
          dk[j][e] = fmaf(ds, q[i][e], dk[j][e]);    // q already carries the 1/sqrt(D)
          dvv[j][e] = fmaf(pm, dO[i][e], dvv[j][e]);
        }
      }
    }
  }
  if (!lane_ok) return;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      float* o = dqkv + (size_t)(base + (long)i * stride) * C3 + head * D + lane * EPL;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        o[e] = dq[i][e] * scale;
        o[C + e] = dk[i][e];
        o[2 * C + e] = dvv[i][e];
      }
    }
  }
}

// ----------------------------------------------------------------------------- weight gradients of 1x1 convs
// dW[n, k] += sum_r G[r, n] * f(A[r, k]),  f = PReLU(*a_slope) or identity.  64x64 tile per CTA, rows split
// over blockIdx.z, fp32 FMA, atomics into dW.
constexpr int WG_T = 64, WG_R = 16;
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ G, const float* __restrict__ A,
                                                    float* __restrict__ dW, int R, int N, int K,
                                                    const float* __restrict__ a_slope, int rows_per_split) {
  grid_dep_wait();
  __shared__ float Gs[WG_R][WG_T + 4];
  __shared__ float As[WG_R][WG_T + 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
  const int r_begin = blockIdx.z * rows_per_split, r_end = min(r_begin + rows_per_split, R);
  const int tx = tid & 15, ty = tid >> 4;  // tx -> k, ty -> n
  const float slope = a_slope ? __ldg(a_slope) : 1.f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += WG_R) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int rr = idx / WG_T, col = idx % WG_T;
      const int r = r0 + rr;
      const bool rok = r < r_end;
      Gs[rr][col] = (rok && n0 + col < N) ? G[(size_t)r * N + n0 + col] : 0.f;
      float av = (rok && k0 + col < K) ? A[(size_t)r * K + k0 + col] : 0.f;
      if (a_slope) av = preluf_(av, slope);
      As[rr][col] = av;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < WG_R; ++rr) {
      float gv[4], av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gv[i] = Gs[rr][ty * 4 + i];
        av[i] = As[rr][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dW + (size_t)n * K + k, acc[i][j]);
    }
  }
}

// db[n] += sum_r G[r, n]
__global__ void colsum_kernel(const float* __restrict__ G, float* __restrict__ db, int R, int N, int rows_per_thread) {
  grid_dep_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int r0 = blockIdx.y * rows_per_thread, r1 = min(r0 + rows_per_thread, R);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += G[(size_t)r * N + n];
  atomicAdd(db + n, acc);
}

// D[r, n] = sum_k G[r, k] * W[k, n]  (W: a forward weight [out = Kd, in = N], so this is its data gradient),
// optionally followed by the PReLU derivative of the forward input u:  D *= (u >= 0 ? 1 : slope),
// dslope += sum D_pre * u * [u < 0].  For the narrow layers (mask conv N_out = 66, bottleneck K = 33).
__global__ void small_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ W, float* __restrict__ D,
                                   int R, int Kd, int N, const float* __restrict__ u,
                                   const float* __restrict__ slope, float* __restrict__ dslope) {
  grid_dep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double dsl = 0.0;
  if (idx < (size_t)R * N) {
    const size_t r = idx / N;
    const int n = (int)(idx % N);
    const float* gr = G + r * Kd;
    float acc = 0.f;
    for (int k = 0; k < Kd; ++k) acc = fmaf(gr[k], __ldg(W + (size_t)k * N + n), acc);
    if (u) {
      const float uv = u[idx];
      if (uv < 0.f) {
        dsl = (double)acc * uv;
        acc *= __ldg(slope);
      }
    }
    D[idx] = acc;
  }
  if (dslope) {
    __shared__ double sh[64];
    double z = 0.0;
#ifdef TD_EMU
    if (!emu::bs) { atomicAdd(dslope, (float)dsl); return; }
#endif
    block_sum2(dsl, z, sh);
    if (threadIdx.x == 0) atomicAdd(dslope, (float)dsl);
  }
}

// Wt[k, n] = W[n, k]
__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int N, int K) {
  grid_dep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * K) return;
  const int k = (int)(idx / N), n = (int)(idx % N);
  Wt[idx] = W[(size_t)n * K + k];
}

// bf16 -> fp32 copy of a stored activation (the fp32 operand of res_conv's weight gradient in bf16 training)
__global__ void bf16_to_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n4) {
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    vstore<4>(dst + 4 * i, act_vload<4>(src, 4 * i, 1));
}

// dst = a + b
__global__ void add_kernel(const float* a, const float* b, float* dst, size_t n) {
  grid_dep_wait();  // dst may alias a or b
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = a[i] + b[i];
}

// ----------------------------------------------------------------------------- concat_block backward
// forward: out = prelu(z), z = cw_c*(mix + y) + cb_c  (TDANet_best.py:388-398)
__global__ void concat_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                  const float* __restrict__ mix, const float* __restrict__ cw,
                                  const float* __restrict__ cb, const float* __restrict__ slope,
                                  float* __restrict__ dy, float* __restrict__ dmix, float* __restrict__ dcw,
                                  float* __restrict__ dcb, float* __restrict__ dslope, int rows, int c,
                                  int rows_per_thread) {
  grid_dep_wait();
  constexpr int V = 4;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  double dsl = 0.0;
  if (ch < c) {
    const int r0 = blockIdx.x * rows_per_thread, r1 = min(r0 + rows_per_thread, rows);
    const vf<V> w = vload<V>(cw + ch), bb = vload<V>(cb + ch);
    const float sl = __ldg(slope);
    vf<V> aw = vzero<V>(), ab = vzero<V>();
    float asl = 0.f;
    for (int r = r0; r < r1; ++r) {
      const size_t off = (size_t)r * c + ch;
      const vf<V> d = vload<V>(dout + off), yv = vload<V>(y + off), mv = vload<V>(mix + off);
      vf<V> o, dm = vload_rw<V>(dmix + off);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float s = mv[e] + yv[e];
        const float z = fmaf(w[e], s, bb[e]);
        float dz = d[e];
        if (z < 0.f) {
          asl = fmaf(d[e], z, asl);
          dz *= sl;
        }
        aw[e] = fmaf(dz, s, aw[e]);
        ab[e] += dz;
        o[e] = dz * w[e];
        dm[e] += o[e];
      }
      vstore<V>(dy + off, o);
      vstore<V>(dmix + off, dm);
    }
    vred_add<V>(dcw + ch, aw);
    vred_add<V>(dcb + ch, ab);
    dsl = asl;
  }
  {
    __shared__ double sh[64];
    double z = 0.0;
#ifdef TD_EMU
    if (!emu::bs) { atomicAdd(dslope, (float)dsl); return; }
#endif
    block_sum2(dsl, z, sh);
    if (threadIdx.x == 0) atomicAdd(dslope, (float)dsl);
  }
}

// ----------------------------------------------------------------------------- mask / decoder / encoder
// masked[r, s*Nb+n] = relu(m) * enc[r, n]:  d_m (in place over d_masked) and d_enc
__global__ void mask_bwd_kernel(float* __restrict__ dmasked, const float* __restrict__ m,
                                const float* __restrict__ enc, float* __restrict__ denc, int rows, int n_src, int Nb) {
  grid_dep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * Nb) return;
  const size_t r = idx / Nb;
  const int n = (int)(idx % Nb);
  const float ev = enc[idx];
  float de = 0.f;
  for (int s = 0; s < n_src; ++s) {
    const size_t off = r * (size_t)(n_src * Nb) + (size_t)s * Nb + n;
    const float dm = dmasked[off], mv = m[off];
    de = fmaf(dm, fmaxf(mv, 0.f), de);
    dmasked[off] = mv > 0.f ? dm * ev : 0.f;
  }
  denc[idx] = de;
}

// full-length signal of the framed (transposed) convolutions: sig_full[o, n] = sig[o, n - shift] inside [0, T)
__device__ __forceinline__ float framed_sig(const float* __restrict__ sig, int n, int shift, int T) {
  const int q = n - shift;
  return (q >= 0 && q < T) ? sig[q] : 0.f;
}

// decoder ConvTranspose1d backward w.r.t. its input:
//   dM[b, l, ci] = sum_o sum_j dfull[b, o, l*S + j - K/2] * W[ci, o, j]
// fallback for configurations whose weights do not fit the shared-memory kernel below: one thread per output
__global__ void dec_bwd_data_naive_kernel(const float* __restrict__ dest, const float* __restrict__ w,
                                          float* __restrict__ dM, int B, int L0, int CI, int NO, int K, int S, int T,
                                          int shift) {
  grid_dep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * L0 * CI) return;
  const int ci = (int)(idx % CI);
  const int l = (int)((idx / CI) % L0);
  const int b = (int)(idx / ((size_t)CI * L0));
  float acc = 0.f;
  for (int o = 0; o < NO; ++o) {
    const float* sp = dest + ((size_t)b * NO + o) * T;
    const float* wp = w + ((size_t)ci * NO + o) * K;
    for (int j = 0; j < K; ++j) acc = fmaf(framed_sig(sp, l * S + j - K / 2, shift, T), __ldg(wp + j), acc);
  }
  dM[idx] = acc;
}

constexpr int DEC_FR = 32;  // frames per CTA
__global__ void __launch_bounds__(256) dec_bwd_data_kernel(const float* __restrict__ dest, const float* __restrict__ w,
                                                           float* __restrict__ dM, int B, int L0, int CI, int NO, int K,
                                                           int S, int T, int shift, int wpad) {
  grid_dep_wait();
  // shared: the signal window of DEC_FR frames for every output channel, and the weights [CI][NO*K (+pad)]
  __shared__ float sh[11264];
  const int span = (DEC_FR - 1) * S + K;
  float* sig = sh;              // [NO][span]
  float* ws = sh + NO * span;   // [CI][wpad]
  const int b = blockIdx.y, l0 = blockIdx.x * DEC_FR;
  const int NK = NO * K;
  for (int i = threadIdx.x; i < NO * span; i += blockDim.x) {
    const int o = i / span, n = l0 * S - K / 2 + i % span;
    sig[i] = framed_sig(dest + ((size_t)b * NO + o) * T, n, shift, T);
  }
  for (int i = threadIdx.x; i < CI * NK; i += blockDim.x) ws[(i / NK) * wpad + i % NK] = __ldg(w + i);
  __syncthreads();
  const int frames = min(DEC_FR, L0 - l0);
  for (int idx = threadIdx.x; idx < frames * CI; idx += blockDim.x) {
    const int f = idx / CI, ci = idx % CI;
    const float* wr = ws + ci * wpad;
    float acc = 0.f;
    for (int o = 0; o < NO; ++o) {
      const float* sp = sig + o * span + f * S;
      for (int j = 0; j < K; ++j) acc = fmaf(sp[j], wr[o * K + j], acc);
    }
    dM[((size_t)b * L0 + l0 + f) * CI + ci] = acc;
  }
}

// dW[ci, o, j] += sum_{b, l} M[b, l, ci] * sig_full[b, o, l*S + j - K/2]
//   decoder weight: M = masked, sig = d_est;   encoder weight: M = d_enc, sig = wav (NO = 1)
__global__ void framed_wgrad_kernel(const float* __restrict__ M, const float* __restrict__ sig,
                                    float* __restrict__ dW, int B, int L0, int CI, int NO, int K, int S, int T,
                                    int shift, int rows_per_split, int m_stride) {
  grid_dep_wait();  // m_stride: row stride of M (>= CI)
  const int j = threadIdx.x;
  const int ci = blockIdx.x / NO, o = blockIdx.x % NO;
  if (j >= K) return;
  const int R = B * L0;
  const int r0 = blockIdx.y * rows_per_split, r1 = min(r0 + rows_per_split, R);
  float acc = 0.f;
  // four rows per trip: their eight loads are independent of the accumulation chain and go out together (the plain
  // loop paid one L2 round trip per row: 148 us for the encoder basis at B = 8); the sum keeps its order
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    float mv[4], sv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int b = (r + q) / L0, l = (r + q) % L0;
      mv[q] = M[(size_t)(r + q) * m_stride + ci];
      sv[q] = framed_sig(sig + ((size_t)b * NO + o) * T, l * S + j - K / 2, shift, T);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc = fmaf(mv[q], sv[q], acc);
  }
  for (; r < r1; ++r) {
    const int b = r / L0, l = r % L0;
    const float mv = M[(size_t)r * m_stride + ci];
    acc = fmaf(mv, framed_sig(sig + ((size_t)b * NO + o) * T, l * S + j - K / 2, shift, T), acc);
  }
  atomicAdd(dW + ((size_t)ci * NO + o) * K + j, acc);
}

}  // namespace td
