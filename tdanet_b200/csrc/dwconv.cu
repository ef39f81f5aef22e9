// Channels-last depthwise convolutions with normalise-on-load sources, and the LA combine.
//
// Replaces, fused: DilatedConvNorm.conv (TDANet_best.py:179-192) + the GlobLN/PReLU of its
// *producer* applied on load, the three ConvNorm.conv of LA (TDANet_best.py:272-289), the
// loc_glo_fus / nearest-interpolate injection (TDANet_best.py:369-371, TDANet.py:624-626)
// recomputed on load, and LA's gate (TDANet_best.py:291).
//
// Memory-bound kernels: one thread owns V consecutive channels and walks over time with a
// register window; all global loads of a chunk of rows are issued before any is consumed.
// CTAs whose rows (halo included) lie inside the tensor run a path without bounds checks; the
// nearest-neighbour row indices of a CTA are tabulated once in shared memory.
#include "kernels.h"

namespace td {

// ----------------------------------------------------------------------------- sources
// EDGE = true: rows outside [0, L) read as exact zeros (conv zero padding).
template <int KIND, int V, bool EDGE>
struct Src {
  const float* x;  // item base + first channel of this thread
  int L, C;
  vf<V> c0_, c1_, c2_, c3_, c4_, c5_;  // coefficient planes
  float slope;
  const float* g;
  const int* jtab;  // nearest global row of local row t at jtab[t - tab0]   (inject kinds)
  int tab0, cur;
  vf<V> sg, eg;

  static constexpr bool kInject = KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD;

  __device__ __forceinline__ void init(const SrcDesc& d, int b, int ch, int C_, const int* jt, int tab0_) {
    C = C_;
    L = d.L;
    x = d.x + (size_t)b * d.L * C_ + ch;
    if constexpr (KIND != SRC_PLAIN) {
      constexpr int planes = (KIND == SRC_INJECT_GATE) ? 6 : 2;
      const float* cf = d.coef + (size_t)b * planes * C_ + ch;
      c0_ = vload<V>(cf);
      c1_ = vload<V>(cf + C_);
      if constexpr (KIND == SRC_INJECT_GATE) {
        c2_ = vload<V>(cf + 2 * C_);
        c3_ = vload<V>(cf + 3 * C_);
        c4_ = vload<V>(cf + 4 * C_);
        c5_ = vload<V>(cf + 5 * C_);
      }
    }
    if constexpr (KIND == SRC_AFFINE_PRELU) slope = __ldg(d.slope);
    if constexpr (kInject) {
      g = d.g + (size_t)b * d.Lg * C_ + ch;
      jtab = jt;
      tab0 = tab0_;
      cur = -1;
    }
  }

  __device__ __forceinline__ vf<V> load_raw(int t) const {
    if (EDGE && (t < 0 || t >= L)) return vzero<V>();
    return vload<V>(x + t * C);
  }

  __device__ __forceinline__ void seek(int j) {
    const vf<V> gr = vload<V>(g + j * C);  // L2-resident: the global feature is small
    cur = j;
    if constexpr (KIND == SRC_INJECT_GATE) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        sg[e] = sigmoidf_(fmaf(c2_[e], gr[e], c3_[e]));
        eg[e] = fmaf(c4_[e], gr[e], c5_[e]);
      }
    } else {
      eg = gr;
    }
  }

  __device__ __forceinline__ vf<V> finalize(vf<V> r, int t) {
    if (EDGE && (t < 0 || t >= L)) return vzero<V>();
    if constexpr (KIND == SRC_PLAIN) {
      return r;
    } else if constexpr (KIND == SRC_AFFINE) {
#pragma unroll
      for (int e = 0; e < V; ++e) r[e] = fmaf(r[e], c0_[e], c1_[e]);
      return r;
    } else if constexpr (KIND == SRC_AFFINE_PRELU) {
#pragma unroll
      for (int e = 0; e < V; ++e) r[e] = preluf_(fmaf(r[e], c0_[e], c1_[e]), slope);
      return r;
    } else {
      const int j = jtab[t - tab0];
      if (j != cur) seek(j);
      if constexpr (KIND == SRC_INJECT_GATE) {
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = fmaf(fmaf(r[e], c0_[e], c1_[e]), sg[e], eg[e]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = fmaf(r[e], c0_[e], c1_[e]) + eg[e];
      }
      return r;
    }
  }
};

// tab[i] = nearest source row (in a tensor of `in_len` rows) of row clamp(t_first + i) of a tensor
// that is `scale` = fl32(in_len / out_len) times shorter/longer; cooperative, caller synchronises.
__device__ __forceinline__ void fill_nearest(int* tab, int n, int t_first, int out_len, float scale, int in_len) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int t = t_first + i;
    t = t < 0 ? 0 : (t >= out_len ? out_len - 1 : t);
    tab[i] = nearest_src(t, scale, in_len);
  }
}

template <int V>
__device__ __forceinline__ void load_taps(const float* __restrict__ w, int ch, vf<V> (&tap)[5]) {
  // w is Conv1d.weight [C,1,5]; channel ch..ch+V-1 are 5*V consecutive floats
  const float* p = w + (size_t)ch * 5;
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int j = 0; j < 5; ++j) tap[j][e] = __ldg(p + e * 5 + j);
}

template <int V>
__device__ __forceinline__ vf<V> conv5(const vf<V> (&tap)[5], const vf<V>& x0, const vf<V>& x1,
                                       const vf<V>& x2, const vf<V>& x3, const vf<V>& x4) {
  vf<V> r;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    float acc = tap[0][e] * x0[e];
    acc = fmaf(tap[1][e], x1[e], acc);
    acc = fmaf(tap[2][e], x2[e], acc);
    acc = fmaf(tap[3][e], x3[e], acc);
    acc = fmaf(tap[4][e], x4[e], acc);
    r[e] = acc;
  }
  return r;
}

// ----------------------------------------------------------------------------- dw k=5
template <int KIND, int V, int NW, int S, int R, bool WRITE, bool STATS, bool EDGE>
__device__ __forceinline__ void dw5_body(const DwArgs& a, int b, int ch, int t0, int t1, const int* jtab, int tab0) {
  Src<KIND, V, EDGE> src;
  src.init(a.src, b, ch, a.C, jtab, tab0);
  vf<V> tap[NW][5], bias[NW], s1[NW], s2[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    load_taps<V>(a.w[i], ch, tap[i]);
    bias[i] = a.bias[i] ? vload<V>(a.bias[i] + ch) : vzero<V>();
    s1[i] = vzero<V>();
    s2[i] = vzero<V>();
  }
  constexpr int NR = (R - 1) * S + 5;  // input rows feeding R outputs
  constexpr int CARRY = 5 - S;         // rows shared with the next chunk
  vf<V> xr[NR];
#pragma unroll
  for (int i = 0; i < CARRY; ++i) {
    const int t = t0 * S - 2 + i;
    xr[R * S + i] = src.finalize(src.load_raw(t), t);
  }
  float* outp = WRITE ? a.out + (size_t)b * a.Lout * a.C + ch : nullptr;

  for (int t = t0; t < t1; t += R) {
#pragma unroll
    for (int i = 0; i < CARRY; ++i) xr[i] = xr[R * S + i];
    const int base = t * S - 2 + CARRY;
#pragma unroll
    for (int i = 0; i < R * S; ++i) xr[CARRY + i] = src.load_raw(base + i);
#pragma unroll
    for (int i = 0; i < R * S; ++i) xr[CARRY + i] = src.finalize(xr[CARRY + i], base + i);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!EDGE || t + r < t1) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          vf<V> y = conv5<V>(tap[i], xr[r * S], xr[r * S + 1], xr[r * S + 2], xr[r * S + 3], xr[r * S + 4]);
#pragma unroll
          for (int e = 0; e < V; ++e) y[e] += bias[i][e];
          if constexpr (STATS) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
              s1[i][e] += y[e];
              s2[i][e] = fmaf(y[e], y[e], s2[i][e]);
            }
          }
          if constexpr (WRITE) {
            if (a.relu) {
#pragma unroll
              for (int e = 0; e < V; ++e) y[e] = fmaxf(y[e], 0.f);
            }
            if (a.round_out) vround_tf32<V>(y);
            vstore<V>(outp + (t + r) * a.C, y);
          }
        }
      }
    }
  }
  if constexpr (STATS) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      float* sp = a.stats + ((size_t)(b * NW + i) * 2) * a.C + ch;
      vred_add<V>(sp, s1[i]);
      vred_add<V>(sp + a.C, s2[i]);
    }
  }
}

template <int KIND, int V, int NW, int S, int R, bool WRITE, bool STATS>
__global__ void __launch_bounds__(256) dw5_kernel(DwArgs a, int rows_per_cta) {
  extern __shared__ int jtab[];  // inject kinds: nearest rows of the input rows this CTA touches
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, a.Lout);
  const int in_first = t0 * S - 2, in_last = (t1 - 1) * S + 2;
  constexpr bool inj = KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD;
  if constexpr (inj) {
    fill_nearest(jtab, rows_per_cta * S + 4, in_first, a.src.L, a.src.gscale, a.src.Lg);
    __syncthreads();
  }
  if (ch >= a.C) return;
  // interior: whole chunks only, every input row inside the tensor
  const bool interior = in_first >= 0 && in_last < a.src.L && (t1 - t0) % R == 0;
  if (interior) dw5_body<KIND, V, NW, S, R, WRITE, STATS, false>(a, b, ch, t0, t1, jtab, in_first);
  else dw5_body<KIND, V, NW, S, R, WRITE, STATS, true>(a, b, ch, t0, t1, jtab, in_first);
}

static void pick_tiling(int B, int L, int ctiles, int R, int* rows_per_cta, int* tiles) {
  // aim at >= ~16 CTAs per SM in total, at most 64 rows per CTA, whole chunks of R rows
  const long target = 148L * 16;
  long per = ((long)B * L * ctiles + target - 1) / target;
  per = (per + R - 1) / R * R;
  if (per < R) per = R;
  if (per > 64) per = 64;
  *rows_per_cta = (int)per;
  *tiles = cdiv(L, (int)per);
}

template <int KIND, int V, int NW, int S, bool WRITE, bool STATS>
static int launch_dw5_t(const DwArgs& a, cudaStream_t st) {
  constexpr int R = (S == 1) ? 8 : 4;
  int threads = a.C / V;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / V, threads);
  int rows, tiles;
  pick_tiling(a.B, a.Lout, ctiles, R, &rows, &tiles);
  dim3 grid(tiles, ctiles, a.B);
  const size_t smem = (KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD) ? (size_t)(rows * S + 4) * sizeof(int) : 0;
  TD_LAUNCH((dw5_kernel<KIND, V, NW, S, R, WRITE, STATS>), grid, threads, smem, st, a, rows);
  return 0;
}

template <int KIND, int V, int NW>
static int launch_dw5_k(const DwArgs& a, cudaStream_t st) {
  const bool wr = a.out != nullptr, stt = a.stats != nullptr;
  if (a.stride == 1) {
    if (wr && stt) return launch_dw5_t<KIND, V, NW, 1, true, true>(a, st);
    if (wr) return launch_dw5_t<KIND, V, NW, 1, true, false>(a, st);
    return launch_dw5_t<KIND, V, NW, 1, false, true>(a, st);
  }
  if constexpr (KIND == SRC_AFFINE && NW == 1) {
    if (wr && stt) return launch_dw5_t<KIND, V, NW, 2, true, true>(a, st);
    if (wr) return launch_dw5_t<KIND, V, NW, 2, true, false>(a, st);
  }
  return fail(TDANET_EINVAL, "dw5: unsupported stride/kind combination");
}

int launch_dw5(const DwArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "dw5: C=%d must be a multiple of 4", a.C);
  TD_REQUIRE(a.nw == 1 || a.nw == 2, "dw5: nw=%d", a.nw);
  TD_REQUIRE(!(a.nw == 2 && a.out), "dw5: writing needs nw == 1");
  TD_REQUIRE(a.out || a.stats, "dw5: nothing to do");
  TD_REQUIRE((long)a.src.L * a.C < (1L << 31) && (long)a.Lout * a.C < (1L << 31), "dw5: item too large for 32-bit offsets");
  if (a.nw == 1) {
    switch (a.kind) {
      case SRC_PLAIN: return launch_dw5_k<SRC_PLAIN, 4, 1>(a, st);
      case SRC_AFFINE: return launch_dw5_k<SRC_AFFINE, 4, 1>(a, st);
      case SRC_AFFINE_PRELU: return launch_dw5_k<SRC_AFFINE_PRELU, 4, 1>(a, st);
      case SRC_INJECT_GATE: return launch_dw5_k<SRC_INJECT_GATE, 4, 1>(a, st);
      case SRC_INJECT_ADD: return launch_dw5_k<SRC_INJECT_ADD, 4, 1>(a, st);
    }
  } else {
    switch (a.kind) {
      case SRC_PLAIN: return launch_dw5_k<SRC_PLAIN, 4, 2>(a, st);
      case SRC_INJECT_GATE: return launch_dw5_k<SRC_INJECT_GATE, 2, 2>(a, st);
      case SRC_INJECT_ADD: return launch_dw5_k<SRC_INJECT_ADD, 4, 2>(a, st);
    }
  }
  return fail(TDANET_EINVAL, "dw5: unsupported source kind %d (nw=%d)", a.kind, a.nw);
}

// ----------------------------------------------------------------------------- generic dw (fork conv_pool)
template <int KIND>
__global__ void dw_generic_kernel(SrcDesc sd, int C, int Lout, int ks, int stride,
                                  const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int round_out) {
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int t = blockIdx.x;
  if (ch >= C || t >= Lout) return;
  Src<KIND, V, true> src;
  src.init(sd, b, ch, C, nullptr, 0);
  const int pad = (ks - 1) / 2;
  vf<V> acc = bias ? vload<V>(bias + ch) : vzero<V>();
  for (int j = 0; j < ks; ++j) {
    const int ti = t * stride - pad + j;
    vf<V> xv = src.finalize(src.load_raw(ti), ti);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = fmaf(__ldg(w + (size_t)(ch + e) * ks + j), xv[e], acc[e]);
  }
  if (round_out) vround_tf32<V>(acc);
  vstore<V>(out + ((size_t)b * Lout + t) * C + ch, acc);
}

int launch_dw_generic(const SrcDesc& src, int kind, int B, int C, int Lout, int ks, int stride,
                      const float* w, const float* bias, float* out, int round_out, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0 && (ks & 1), "dw_generic: C=%d ks=%d", C, ks);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  dim3 grid(Lout, cdiv(C / 4, threads), B);
  if (kind == SRC_AFFINE) {
    TD_LAUNCH((dw_generic_kernel<SRC_AFFINE>), grid, threads, 0, st, src, C, Lout, ks, stride, w, bias, out, round_out);
  } else if (kind == SRC_PLAIN) {
    TD_LAUNCH((dw_generic_kernel<SRC_PLAIN>), grid, threads, 0, st, src, C, Lout, ks, stride, w, bias, out, round_out);
  } else {
    return fail(TDANET_EINVAL, "dw_generic: unsupported source kind %d", kind);
  }
  return 0;
}

// ----------------------------------------------------------------------------- LA combine
// GC > 0 (Lg <= Ll): the global-branch convs are evaluated once per distinct source row of a chunk
//          (at most GC of them) and parked in a thread-private shared-memory column, then gathered.
// GC = 0 (Lg >  Ll): only the first top-down step (reference quirk, TDANet_best.py:375-376);
//          evaluated per output row.
struct LaSmem {
  int* jc;   // [rows]      nearest global row of each output row
  int* jl;   // [rows + 4]  local inject: nearest g row of each local input row
  int* jg;   // [glo span]  global inject: nearest g row of each global input row
  float* scratch;
};

template <int LKIND, int GKIND, int V, int GC, bool EDGE>
__device__ __forceinline__ void la_body(const LaArgs& a, int b, int ch, int t0, int t1, const LaSmem& sm, int g_first) {
  constexpr int R = 8;
  const int Ll = a.loc.L;
  const int colw = blockDim.x * V;
  float* mine = sm.scratch + threadIdx.x * V;

  Src<LKIND, V, EDGE> sl;
  sl.init(a.loc, b, ch, a.C, sm.jl, t0 - 2);
  Src<GKIND, V, EDGE> sg;
  sg.init(a.glo, b, ch, a.C, sm.jg, g_first);
  vf<V> wl[5], wa[5], we[5];
  load_taps<V>(a.wl, ch, wl);
  load_taps<V>(a.wa, ch, wa);
  load_taps<V>(a.we, ch, we);
  const float* cf = a.coef + (size_t)b * 6 * a.C + ch;
  const vf<V> sL = vload<V>(cf), hL = vload<V>(cf + a.C), sA = vload<V>(cf + 2 * a.C),
              hA = vload<V>(cf + 3 * a.C), sE = vload<V>(cf + 4 * a.C), hE = vload<V>(cf + 5 * a.C);

  vf<V> xr[R + 4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 - 2 + i;
    xr[R + i] = sl.finalize(sl.load_raw(t), t);
  }
  float* outp = a.out + (size_t)b * Ll * a.C + ch;

  for (int t = t0; t < t1; t += R) {
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = xr[R + i];
#pragma unroll
    for (int i = 0; i < R; ++i) xr[4 + i] = sl.load_raw(t + 2 + i);
    const int jlo = sm.jc[t - t0];
    if constexpr (GC > 0) {
      // distinct centres of this chunk: jlo .. jc[last row]  (<= GC by construction)
      const int tl = (EDGE ? min(t + R, t1) : t + R) - 1;
      const int nc = sm.jc[tl - t0] - jlo + 1;
      vf<V> gr[GC + 4];
#pragma unroll
      for (int i = 0; i < GC + 4; ++i) gr[i] = sg.load_raw(jlo - 2 + i);
#pragma unroll
      for (int i = 0; i < GC + 4; ++i) gr[i] = sg.finalize(gr[i], jlo - 2 + i);
#pragma unroll
      for (int i = 0; i < GC; ++i) {
        if (i < nc) {
          vf<V> ca = conv5<V>(wa, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]);
          vf<V> ce = conv5<V>(we, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            ca[e] = sigmoidf_(fmaf(sA[e], ca[e], hA[e]));
            ce[e] = fmaf(sE[e], ce[e], hE[e]);
          }
          vstore<V>(mine + (2 * i) * colw, ca);
          vstore<V>(mine + (2 * i + 1) * colw, ce);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) xr[4 + i] = sl.finalize(xr[4 + i], t + 2 + i);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!EDGE || t + r < t1) {
        vf<V> cl = conv5<V>(wl, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        vf<V> ga, ge;
        const int j = sm.jc[t + r - t0];
        if constexpr (GC == 0) {
          vf<V> g5[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) g5[i] = sg.load_raw(j - 2 + i);
#pragma unroll
          for (int i = 0; i < 5; ++i) g5[i] = sg.finalize(g5[i], j - 2 + i);
          ga = conv5<V>(wa, g5[0], g5[1], g5[2], g5[3], g5[4]);
          ge = conv5<V>(we, g5[0], g5[1], g5[2], g5[3], g5[4]);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            ga[e] = sigmoidf_(fmaf(sA[e], ga[e], hA[e]));
            ge[e] = fmaf(sE[e], ge[e], hE[e]);
          }
        } else {
          const float* col = mine + (2 * (j - jlo)) * colw;
          if constexpr (V == 2) {
            const float2 u = *reinterpret_cast<const float2*>(col), w2 = *reinterpret_cast<const float2*>(col + colw);
            ga[0] = u.x; ga[1] = u.y; ge[0] = w2.x; ge[1] = w2.y;
          } else {
            const float4 u = *reinterpret_cast<const float4*>(col), w4 = *reinterpret_cast<const float4*>(col + colw);
            ga[0] = u.x; ga[1] = u.y; ga[2] = u.z; ga[3] = u.w;
            ge[0] = w4.x; ge[1] = w4.y; ge[2] = w4.z; ge[3] = w4.w;
          }
        }
        vf<V> y;
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = fmaf(fmaf(sL[e], cl[e], hL[e]), ga[e], ge[e]);
        if (a.round_out) vround_tf32<V>(y);
        vstore<V>(outp + (t + r) * a.C, y);
      }
    }
  }
}

template <int LKIND, int GKIND, int V, int GC>
__global__ void __launch_bounds__(256, 2) la_combine_kernel(LaArgs a, int rows_per_cta, int gspan) {
  extern __shared__ __align__(16) float la_smem[];
  constexpr int R = 8;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int Ll = a.loc.L, Lg = a.glo.L;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, Ll);
  LaSmem sm;
  sm.scratch = la_smem;
  sm.jc = reinterpret_cast<int*>(la_smem + (GC > 0 ? GC : 0) * 2 * blockDim.x * V);
  sm.jl = sm.jc + rows_per_cta;
  sm.jg = sm.jl + rows_per_cta + 4;
  // first global row any output row of this CTA can touch (halo included)
  const int g_first = nearest_src(t0, a.scale, Lg) - 2;
  fill_nearest(sm.jc, rows_per_cta, t0, Ll, a.scale, Lg);
  if constexpr (LKIND == SRC_INJECT_GATE || LKIND == SRC_INJECT_ADD)
    fill_nearest(sm.jl, rows_per_cta + 4, t0 - 2, Ll, a.loc.gscale, a.loc.Lg);
  if constexpr (GKIND == SRC_INJECT_GATE || GKIND == SRC_INJECT_ADD)
    fill_nearest(sm.jg, gspan, g_first, Lg, a.glo.gscale, a.glo.Lg);
  __syncthreads();
  if (ch >= a.C) return;
  const int g_last = nearest_src(t1 - 1, a.scale, Lg) + 2 + (GC > 0 ? GC : 0);
  const bool interior = t0 - 2 >= 0 && t1 + 2 <= Ll && (t1 - t0) % R == 0 && g_first >= 0 && g_last < Lg;
  if (interior) la_body<LKIND, GKIND, V, GC, false>(a, b, ch, t0, t1, sm, g_first);
  else la_body<LKIND, GKIND, V, GC, true>(a, b, ch, t0, t1, sm, g_first);
}

template <int LKIND, int GKIND>
static int launch_la_t(const LaArgs& a, cudaStream_t st) {
  constexpr int V = 2;
  int threads = a.C / V;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / V, threads);
  int rows, tiles;
  pick_tiling(a.B, a.loc.L, ctiles, 8, &rows, &tiles);
  dim3 grid(tiles, ctiles, a.B);
  // rows of the global tensor one CTA can touch: its rows map to <= rows*scale + 1 centres, + halo
  const int gspan = (int)((double)rows * a.glo.L / a.loc.L) + 16;
  const size_t tabs = (size_t)(2 * rows + 4 + gspan) * sizeof(int);
  if (a.glo.L > a.loc.L) {
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, 0>), grid, threads, tabs, st, a, rows, gspan);
  } else if (7.0 * a.glo.L / a.loc.L <= 3.99) {
    // ratio >= ~2 (every up-sampling step of the U-Net): 8 output rows see at most 5 centres
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, 5>), grid, threads, tabs + (size_t)5 * 2 * threads * V * sizeof(float), st, a, rows, gspan);
  } else {
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, 8>), grid, threads, tabs + (size_t)8 * 2 * threads * V * sizeof(float), st, a, rows, gspan);
  }
  return 0;
}

int launch_la_combine(const LaArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "la: C=%d must be a multiple of 4", a.C);
  TD_REQUIRE((long)a.loc.L * a.C < (1L << 31) && (long)a.glo.L * a.C < (1L << 31), "la: item too large for 32-bit offsets");
  if (a.lkind == SRC_INJECT_GATE && a.gkind == SRC_INJECT_GATE) return launch_la_t<SRC_INJECT_GATE, SRC_INJECT_GATE>(a, st);
  if (a.lkind == SRC_INJECT_GATE && a.gkind == SRC_PLAIN) return launch_la_t<SRC_INJECT_GATE, SRC_PLAIN>(a, st);
  if (a.lkind == SRC_INJECT_ADD && a.gkind == SRC_INJECT_ADD) return launch_la_t<SRC_INJECT_ADD, SRC_INJECT_ADD>(a, st);
  if (a.lkind == SRC_INJECT_ADD && a.gkind == SRC_PLAIN) return launch_la_t<SRC_INJECT_ADD, SRC_PLAIN>(a, st);
  if (a.lkind == SRC_PLAIN && a.gkind == SRC_PLAIN) return launch_la_t<SRC_PLAIN, SRC_PLAIN>(a, st);
  return fail(TDANET_EINVAL, "la: unsupported source kinds %d/%d", a.lkind, a.gkind);
}

}  // namespace td
