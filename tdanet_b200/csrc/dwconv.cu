// Channels-last depthwise convolutions with normalise-on-load sources, and the LA combine.
//
// Replaces, fused: DilatedConvNorm.conv (TDANet_best.py:179-192) + the GlobLN/PReLU of its
// *producer* applied on load, the three ConvNorm.conv of LA (TDANet_best.py:272-289), the
// loc_glo_fus / nearest-interpolate injection (TDANet_best.py:369-371, TDANet.py:624-626)
// recomputed on load, and LA's gate (TDANet_best.py:291).
//
// Memory-bound kernels: one thread owns V consecutive channels and walks over time with a
// register window; all global loads of a chunk of rows are issued before any is consumed.
#include "kernels.h"

namespace td {

// ----------------------------------------------------------------------------- sources
template <int KIND, int V>
struct Src {
  const float* x;
  int L, C;
  vf<V> c0_, c1_, c2_, c3_, c4_, c5_;  // coefficient planes
  float slope;
  const float* g;
  int Lg, cur;
  float gscale;
  vf<V> sg, eg, nxt;

  __device__ __forceinline__ void init(const SrcDesc& d, int b, int ch, int C_) {
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
    if constexpr (KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD) {
      g = d.g + (size_t)b * d.Lg * C_ + ch;
      Lg = d.Lg;
      gscale = d.gscale;
      cur = -2;
    }
  }

  __device__ __forceinline__ vf<V> load_raw(int t) const {
    if (t < 0 || t >= L) return vzero<V>();
    return vload<V>(x + (size_t)t * C);
  }

  __device__ __forceinline__ void seek(int j) {
    vf<V> gr = (j == cur + 1) ? nxt : vload<V>(g + (size_t)j * C);
    cur = j;
    if (j + 1 < Lg) nxt = vload<V>(g + (size_t)(j + 1) * C);  // prefetch the next global row
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

  // value of row t given its raw load; exact zero outside [0, L) (conv zero padding)
  __device__ __forceinline__ vf<V> finalize(vf<V> r, int t) {
    if (t < 0 || t >= L) return vzero<V>();
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
      const int j = nearest_src(t, gscale, Lg);
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
template <int KIND, int V, int NW, int S, int R, bool WRITE, bool STATS>
__global__ void __launch_bounds__(256) dw5_kernel(DwArgs a, int rows_per_cta) {
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= a.C) return;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, a.Lout);

  Src<KIND, V> src;
  src.init(a.src, b, ch, a.C);
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
      if (t + r < t1) {
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
            vstore<V>(outp + (size_t)(t + r) * a.C, y);
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
  TD_LAUNCH((dw5_kernel<KIND, V, NW, S, R, WRITE, STATS>), grid, threads, 0, st, a, rows);
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
  if (a.nw == 1) {
    switch (a.kind) {
      case SRC_PLAIN: return launch_dw5_k<SRC_PLAIN, 4, 1>(a, st);
      case SRC_AFFINE: return launch_dw5_k<SRC_AFFINE, 4, 1>(a, st);
      case SRC_AFFINE_PRELU: return launch_dw5_k<SRC_AFFINE_PRELU, 4, 1>(a, st);
      case SRC_INJECT_GATE: return launch_dw5_k<SRC_INJECT_GATE, 2, 1>(a, st);
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
  Src<KIND, V> src;
  src.init(sd, b, ch, C);
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
// UP   (Lg <= Ll): the global-branch conv is evaluated once per distinct source row of a chunk
//                  and parked in a thread-private shared-memory column, then gathered per row.
// DOWN (Lg >  Ll): only the first top-down step (reference quirk, TDANet_best.py:375-376);
//                  evaluated per output row.
template <int LKIND, int GKIND, int V, bool DOWN>
__global__ void __launch_bounds__(256, 2) la_combine_kernel(LaArgs a, int rows_per_cta) {
  constexpr int R = 8;
  extern __shared__ float scratch[];  // [R][2][blockDim.x * V], column = this thread's channels
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= a.C) return;
  const int Ll = a.loc.L, Lg = a.glo.L;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, Ll);
  const int colw = blockDim.x * V;
  float* mine = scratch + threadIdx.x * V;

  Src<LKIND, V> sl;
  sl.init(a.loc, b, ch, a.C);
  Src<GKIND, V> sg;
  sg.init(a.glo, b, ch, a.C);
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
    // ---- issue every load of the chunk
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = xr[R + i];
#pragma unroll
    for (int i = 0; i < R; ++i) xr[4 + i] = sl.load_raw(t + 2 + i);
    const int jlo = nearest_src(t, a.scale, Lg);
    if constexpr (!DOWN) {
      const int tl = min(t + R, t1) - 1;
      const int nc = nearest_src(tl, a.scale, Lg) - jlo + 1;  // <= R because Lg <= Ll
      vf<V> gr[R + 4];
#pragma unroll
      for (int i = 0; i < R + 4; ++i) gr[i] = (i < nc + 4) ? sg.load_raw(jlo - 2 + i) : vzero<V>();
#pragma unroll
      for (int i = 0; i < R + 4; ++i)
        if (i < nc + 4) gr[i] = sg.finalize(gr[i], jlo - 2 + i);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        if (i < nc) {
          vf<V> ca = conv5<V>(wa, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]);
          vf<V> ce = conv5<V>(we, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            ca[e] = sigmoidf_(fmaf(sA[e], ca[e], hA[e]));
            ce[e] = fmaf(sE[e], ce[e], hE[e]);
          }
          vstore<V>(mine + (size_t)(2 * i) * colw, ca);
          vstore<V>(mine + (size_t)(2 * i + 1) * colw, ce);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) xr[4 + i] = sl.finalize(xr[4 + i], t + 2 + i);
    // ---- combine
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (t + r < t1) {
        vf<V> cl = conv5<V>(wl, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        vf<V> ga, ge;
        const int j = nearest_src(t + r, a.scale, Lg);
        if constexpr (DOWN) {
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
          const float* col = mine + (size_t)(2 * (j - jlo)) * colw;
          if constexpr (V == 2) {
            float2 u = *reinterpret_cast<const float2*>(col), w2 = *reinterpret_cast<const float2*>(col + colw);
            ga[0] = u.x; ga[1] = u.y; ge[0] = w2.x; ge[1] = w2.y;
          } else {
            float4 u = *reinterpret_cast<const float4*>(col), w4 = *reinterpret_cast<const float4*>(col + colw);
            ga[0] = u.x; ga[1] = u.y; ga[2] = u.z; ga[3] = u.w;
            ge[0] = w4.x; ge[1] = w4.y; ge[2] = w4.z; ge[3] = w4.w;
          }
        }
        vf<V> y;
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = fmaf(fmaf(sL[e], cl[e], hL[e]), ga[e], ge[e]);
        if (a.round_out) vround_tf32<V>(y);
        vstore<V>(outp + (size_t)(t + r) * a.C, y);
      }
    }
  }
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
  const size_t smem = (size_t)8 * 2 * threads * V * sizeof(float);
  if (a.glo.L > a.loc.L) {
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, true>), grid, threads, 0, st, a, rows);
  } else {
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, false>), grid, threads, smem, st, a, rows);
  }
  return 0;
}

int launch_la_combine(const LaArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "la: C=%d must be a multiple of 4", a.C);
  if (a.lkind == SRC_INJECT_GATE && a.gkind == SRC_INJECT_GATE) return launch_la_t<SRC_INJECT_GATE, SRC_INJECT_GATE>(a, st);
  if (a.lkind == SRC_INJECT_GATE && a.gkind == SRC_PLAIN) return launch_la_t<SRC_INJECT_GATE, SRC_PLAIN>(a, st);
  if (a.lkind == SRC_INJECT_ADD && a.gkind == SRC_INJECT_ADD) return launch_la_t<SRC_INJECT_ADD, SRC_INJECT_ADD>(a, st);
  if (a.lkind == SRC_INJECT_ADD && a.gkind == SRC_PLAIN) return launch_la_t<SRC_INJECT_ADD, SRC_PLAIN>(a, st);
  if (a.lkind == SRC_PLAIN && a.gkind == SRC_PLAIN) return launch_la_t<SRC_PLAIN, SRC_PLAIN>(a, st);
  return fail(TDANET_EINVAL, "la: unsupported source kinds %d/%d", a.lkind, a.gkind);
}

}  // namespace td
