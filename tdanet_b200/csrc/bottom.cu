// Bottom-scale (most down-sampled sequence) kernels: multi-scale pooling, LayerNorm + positional
// encoding, the fused multi-head attention core, post-attention LayerNorm + residual, FFN glue.
//
// Reference: UConvBlock.forward global-feature gather (TDANet_best.py:358-365), GA / MultiHeadAttention
// (TDANet_best.py:236-264), MultiHeadAttentionFixed (TDANet_mult_tes.py:254-272).
#include "kernels.h"
#include <cfloat>

namespace td {

// ----------------------------------------------------------------------------- pooling
// out[b, j, c] = sum_k mean_{t in bin_k(j)} (x_k[b,t,c]*scale_k + shift_k)
// bin_k(j) = [floor(j*L_k/Lb), ceil((j+1)*L_k/Lb))   (F.adaptive_avg_pool1d)
__global__ void pool_sum_kernel(PoolArgs a) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z, j = blockIdx.x;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= a.C) return;
  vf<V> acc = vzero<V>();
  for (int k = 0; k < a.n; ++k) {
    const int L = a.L[k];
    const int lo = (int)(((long)j * L) / a.Lb);
    const int hi = (int)((((long)j + 1) * L + a.Lb - 1) / a.Lb);
    const float* x = a.x[k] + ((size_t)b * L) * a.C + ch;
    vf<V> s = vzero<V>();
    for (int t = lo; t < hi; ++t) {
      vf<V> v = vload<V>(x + (size_t)t * a.C);
#pragma unroll
      for (int e = 0; e < V; ++e) s[e] += v[e];
    }
    const float inv = 1.f / (float)(hi - lo);
    vf<V> sc, sh;
    norm_coef<V>(a.norm[k], b, ch, sc, sh);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] += fmaf(s[e] * inv, sc[e], sh[e]);
  }
  vstore<V>(a.out + ((size_t)b * a.Lb + j) * a.C + ch, acc);
}

int launch_pool_sum(const PoolArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "pool: C=%d", a.C);
  int threads = a.C / 4 > 256 ? 256 : (a.C / 4 < 32 ? 32 : a.C / 4);
  dim3 grid(a.Lb, cdiv(a.C / 4, threads), a.B);
  TD_LAUNCH(pool_sum_kernel, grid, threads, 0, st, a);
  return 0;
}

__global__ void affine_sum_kernel(PoolArgs a) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z, j = blockIdx.x;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= a.C) return;
  vf<V> acc = vzero<V>();
  for (int k = 0; k < a.n; ++k) {
    vf<V> v = vload<V>(a.x[k] + ((size_t)b * a.Lb + j) * a.C + ch);
    vf<V> sc, sh;
    norm_coef<V>(a.norm[k], b, ch, sc, sh);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] += fmaf(v[e], sc[e], sh[e]);
  }
  vstore<V>(a.out + ((size_t)b * a.Lb + j) * a.C + ch, acc);
}

// The same sum with the LayerNorm + positional encoding that follows it: the CTA holds the whole token row
// (C / 4 threads), so the two-pass row moments are two block reductions and ln_pe's launch disappears.
__global__ void affine_sum_ln_kernel(PoolArgs a) {
  grid_dep_wait();
  constexpr int V = 4;
  __shared__ float red[2][32];
  const int b = blockIdx.z, j = blockIdx.x;
  const int ch = threadIdx.x * V;
  const bool live = ch < a.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  vf<V> acc = vzero<V>();
  if (live) {
    for (int k = 0; k < a.n; ++k) {
      vf<V> v = vload<V>(a.x[k] + ((size_t)b * a.Lb + j) * a.C + ch);
      vf<V> sc, sh;
      norm_coef<V>(a.norm[k], b, ch, sc, sh);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += fmaf(v[e], sc[e], sh[e]);
    }
    vstore<V>(a.out + ((size_t)b * a.Lb + j) * a.C + ch, acc);
  }
  float s = live ? (acc[0] + acc[1]) + (acc[2] + acc[3]) : 0.f;
  s = warp_sum(s);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < nw; ++i) tot += red[0][i];
  const float mu = tot / (float)a.C;
  float q = 0.f;
  if (live) {
    const float d0 = acc[0] - mu, d1 = acc[1] - mu, d2 = acc[2] - mu, d3 = acc[3] - mu;
    q = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  q = warp_sum(q);
  if (lane == 0) red[1][warp] = q;
  __syncthreads();
  float qt = 0.f;
  for (int i = 0; i < nw; ++i) qt += red[1][i];
  const float rstd = rsqrtf(qt / (float)a.C + kEpsLN);
  if (!live) return;
  const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_w + ch));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(a.ln_b + ch));
  const float4 p = __ldg(reinterpret_cast<const float4*>(a.pe + (size_t)j * a.C + ch));
  float4 o;
  o.x = (acc[0] - mu) * rstd * g.x + bb.x + p.x;
  o.y = (acc[1] - mu) * rstd * g.y + bb.y + p.y;
  o.z = (acc[2] - mu) * rstd * g.z + bb.z + p.z;
  o.w = (acc[3] - mu) * rstd * g.w + bb.w + p.w;
  if (a.ln_round) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
  *reinterpret_cast<float4*>(a.ln_out + ((size_t)b * a.Lb + j) * a.C + ch) = o;
}

// One WARP per token row (C a multiple of 128, <= 1024): a lane owns the channel quads lane*4 + 128*q, the row moments
// are warp shuffles - no block barrier, four rows per 128-thread CTA, and the loads of the n pooled tensors of a row
// are all in flight before the first is used.  (The CTA-per-row form above: 8,064 CTAs of two block reductions each,
// 34 us per launch at B = 64.)
template <int Q>
__global__ void __launch_bounds__(128) affine_sum_ln_warp_kernel(PoolArgs a, int rows) {
  grid_dep_wait();
  constexpr int V = 4;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / a.Lb, j = row % a.Lb;
  const size_t roff = (size_t)row * a.C;
  vf<V> acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = vzero<V>();
  for (int k = 0; k < a.n; ++k) {
    vf<V> v[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) v[q] = vload<V>(a.x[k] + roff + lane * 4 + 128 * q);
    float r, mur;
    norm_moments(a.norm[k], b, r, mur);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const vf<V> g = vload<V>(a.norm[k].gamma + lane * 4 + 128 * q), be = vload<V>(a.norm[k].beta + lane * 4 + 128 * q);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[q][e] += fmaf(v[q][e], g[e] * r, fmaf(-g[e], mur, be[e]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    vstore<V>(a.out + roff + lane * 4 + 128 * q, acc[q]);
    s += (acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3]);
  }
  const float mu = warp_sum(s) / (float)a.C;
  float qq = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const float d0 = acc[q][0] - mu, d1 = acc[q][1] - mu, d2 = acc[q][2] - mu, d3 = acc[q][3] - mu;
    qq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = rsqrtf(warp_sum(qq) / (float)a.C + kEpsLN);
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = lane * 4 + 128 * q;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_w + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.ln_b + c));
    const float4 pe = __ldg(reinterpret_cast<const float4*>(a.pe + (size_t)j * a.C + c));
    float4 o;
    o.x = (acc[q][0] - mu) * rstd * g.x + bb.x + pe.x;
    o.y = (acc[q][1] - mu) * rstd * g.y + bb.y + pe.y;
    o.z = (acc[q][2] - mu) * rstd * g.z + bb.z + pe.z;
    o.w = (acc[q][3] - mu) * rstd * g.w + bb.w + pe.w;
    if (a.ln_round) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
    *reinterpret_cast<float4*>(a.ln_out + roff + c) = o;
  }
}

bool launch_affine_sum_fuses_ln(const PoolArgs& a) {
  static const bool off = getenv("TDANET_FUSE_LN_PE") && atoi(getenv("TDANET_FUSE_LN_PE")) == 0;
  return !off && a.ln_out && a.ln_w && a.ln_b && a.pe && a.C % 4 == 0 && a.C / 4 <= 256;
}

int launch_affine_sum(const PoolArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "affine_sum: C=%d", a.C);
  int threads = a.C / 4 > 256 ? 256 : (a.C / 4 < 32 ? 32 : a.C / 4);
  dim3 grid(a.Lb, cdiv(a.C / 4, threads), a.B);
  if (launch_affine_sum_fuses_ln(a)) {
    static const bool warp_rows = !(getenv("TDANET_POOLSUM_WARP") && atoi(getenv("TDANET_POOLSUM_WARP")) == 0);
    const int rows = a.B * a.Lb;
    if (warp_rows && a.C % 128 == 0 && a.C <= 1024) {
      const int q = a.C / 128;
      const unsigned g = (unsigned)cdiv(rows, 4);
      if (q == 4) { TD_LAUNCH((affine_sum_ln_warp_kernel<4>), g, 128, 0, st, a, rows); return 0; }
      if (q == 2) { TD_LAUNCH((affine_sum_ln_warp_kernel<2>), g, 128, 0, st, a, rows); return 0; }
      if (q == 1) { TD_LAUNCH((affine_sum_ln_warp_kernel<1>), g, 128, 0, st, a, rows); return 0; }
      if (q == 8) { TD_LAUNCH((affine_sum_ln_warp_kernel<8>), g, 128, 0, st, a, rows); return 0; }
    }
    threads = (threads + 31) / 32 * 32;
    TD_LAUNCH_COOP(affine_sum_ln_kernel, grid, threads, 0, st, a);
    return 0;
  }
  TD_LAUNCH(affine_sum_kernel, grid, threads, 0, st, a);
  return 0;
}

// ----------------------------------------------------------------------------- LayerNorm rows
// One warp per token row; rows are re-read from L1 instead of being held in registers (C <= 1024,
// tensors at this scale are tiny).
__device__ __forceinline__ void row_moments(const float* __restrict__ x, const float* __restrict__ x2,
                                            float k1, float k2, int C, int lane, float& mu, float& rstd) {
  // statistics of v = k1*x + k2*x2 (x2 may be null)
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(x + c);
    if (x2) {
      float4 u = *reinterpret_cast<const float4*>(x2 + c);
      v.x = k1 * v.x + k2 * u.x; v.y = k1 * v.y + k2 * u.y; v.z = k1 * v.z + k2 * u.z; v.w = k1 * v.w + k2 * u.w;
    } else {
      v.x *= k1; v.y *= k1; v.z *= k1; v.w *= k1;
    }
    s += (v.x + v.y) + (v.z + v.w);
  }
  mu = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(x + c);
    if (x2) {
      float4 u = *reinterpret_cast<const float4*>(x2 + c);
      v.x = k1 * v.x + k2 * u.x; v.y = k1 * v.y + k2 * u.y; v.z = k1 * v.z + k2 * u.z; v.w = k1 * v.w + k2 * u.w;
    } else {
      v.x *= k1; v.y *= k1; v.z *= k1; v.w *= k1;
    }
    const float d0 = v.x - mu, d1 = v.y - mu, d2 = v.z - mu, d3 = v.w - mu;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  rstd = rsqrtf(warp_sum(q) / (float)C + kEpsLN);
}

__global__ void ln_pe_kernel(const float* __restrict__ x, const float* __restrict__ w,
                             const float* __restrict__ bias, const float* __restrict__ pe,
                             float* __restrict__ y, int rows, int L, int C, int round_out) {
  grid_dep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int t = row % L;
  const float* xr = x + (size_t)row * C;
  float mu, rstd;
  row_moments(xr, nullptr, 1.f, 0.f, C, lane, mu, rstd);
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + c);
    float4 g = __ldg(reinterpret_cast<const float4*>(w + c));
    float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 p = __ldg(reinterpret_cast<const float4*>(pe + (size_t)t * C + c));
    float4 o;
    o.x = (v.x - mu) * rstd * g.x + bb.x + p.x;
    o.y = (v.y - mu) * rstd * g.y + bb.y + p.y;
    o.z = (v.z - mu) * rstd * g.z + bb.z + p.z;
    o.w = (v.w - mu) * rstd * g.w + bb.w + p.w;
    if (round_out) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
    *reinterpret_cast<float4*>(y + (size_t)row * C + c) = o;
  }
}

int launch_ln_pe(const float* x, const float* w, const float* b, const float* pe, float* y, int B,
                 int L, int C, int round_out, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "ln_pe: C=%d", C);
  const int rows = B * L;
  TD_LAUNCH(ln_pe_kernel, cdiv(rows, 8), 256, 0, st, x, w, b, pe, y, rows, L, C, round_out);
  return 0;
}

// y = resid + f*(LN(k1*a + k2*xin)*w + b);  f = DropPath multiplier of the row's item (training), else 1
__global__ void ln_residual_kernel(const float* __restrict__ a, const float* __restrict__ xin,
                                   const float* __restrict__ resid, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ y, float k1,
                                   float k2, int rows, int C, const uint8_t* __restrict__ item_mask,
                                   float item_inv_keep, int L) {
  grid_dep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float f = item_mask ? (item_mask[row / L] ? item_inv_keep : 0.f) : 1.f;
  const float* ar = a + (size_t)row * C;
  const float* xr = xin ? xin + (size_t)row * C : nullptr;
  float mu, rstd;
  row_moments(ar, xr, k1, k2, C, lane, mu, rstd);
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(ar + c);
    if (xr) {
      float4 u = *reinterpret_cast<const float4*>(xr + c);
      v.x = k1 * v.x + k2 * u.x; v.y = k1 * v.y + k2 * u.y; v.z = k1 * v.z + k2 * u.z; v.w = k1 * v.w + k2 * u.w;
    } else {
      v.x *= k1; v.y *= k1; v.z *= k1; v.w *= k1;
    }
    float4 g = __ldg(reinterpret_cast<const float4*>(w + c));
    float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 r = *reinterpret_cast<const float4*>(resid + (size_t)row * C + c);
    float4 o;
    if (item_mask) {
      o.x = fmaf(f, (v.x - mu) * rstd * g.x + bb.x, r.x);
      o.y = fmaf(f, (v.y - mu) * rstd * g.y + bb.y, r.y);
      o.z = fmaf(f, (v.z - mu) * rstd * g.z + bb.z, r.z);
      o.w = fmaf(f, (v.w - mu) * rstd * g.w + bb.w, r.w);
    } else {
      o.x = r.x + ((v.x - mu) * rstd * g.x + bb.x);
      o.y = r.y + ((v.y - mu) * rstd * g.y + bb.y);
      o.z = r.z + ((v.z - mu) * rstd * g.z + bb.z);
      o.w = r.w + ((v.w - mu) * rstd * g.w + bb.w);
    }
    *reinterpret_cast<float4*>(y + (size_t)row * C + c) = o;
  }
}

int launch_ln_residual(const float* a, const float* xin, const float* resid, const float* w,
                       const float* b, float* y, int mode, const DropRef& drop, int B, int L, int C, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "ln_residual: C=%d", C);
  TD_REQUIRE(mode >= 0 && mode <= 2, "ln_residual: mode %d", mode);
  const int rows = B * L;
  // mode 1: LayerNorm(out + dropout(out)) == LN(2*out) in eval (TDANet_best.py:251); mode 2: the caller folded the
  // training-mode (1 + mask/keep) into `a`
  const float k1 = mode == 1 ? 2.f : 1.f, k2 = mode == 0 ? 1.f : 0.f;
  TD_LAUNCH(ln_residual_kernel, cdiv(rows, 8), 256, 0, st, a, mode == 0 ? xin : nullptr, resid, w, b, y,
            k1, k2, rows, C, drop.item_mask, drop.item_inv_keep, L);
  return 0;
}

// ----------------------------------------------------------------------------- affine glue
template <bool RES, bool STATS, bool DROP = false>
__global__ void affine_kernel(const float* __restrict__ x, NormRef norm,
                              const float* __restrict__ resid, float* __restrict__ y,
                              float* __restrict__ chstats, int L, int C, int rows_per_cta, DropRef drop,
                              DetRef det) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(t0 + rows_per_cta, L);
  vf<V> sc, sh;
  norm_coef<V>(norm, b, ch, sc, sh);
  float fi = 1.f;
  if constexpr (DROP) fi = drop.item_mask ? (drop.item_mask[b] ? drop.item_inv_keep : 0.f) : 1.f;
  vf<V> s1 = vzero<V>(), s2 = vzero<V>();
  for (int t = t0; t < t1; ++t) {
    const size_t off = ((size_t)b * L + t) * C + ch;
    vf<V> v = vload<V>(x + off);
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
    if constexpr (DROP) {
      // FFN.drop after fc2's GlobLN (TDANet_best.py:212), then GA.drop_path on the branch (:263)
      uchar4 mk = drop.mask ? *reinterpret_cast<const uchar4*>(drop.mask + off) : make_uchar4(1, 1, 1, 1);
      const float fe = drop.mask ? drop.inv_keep * fi : fi;
      v[0] = mk.x ? v[0] * fe : 0.f; v[1] = mk.y ? v[1] * fe : 0.f;
      v[2] = mk.z ? v[2] * fe : 0.f; v[3] = mk.w ? v[3] * fe : 0.f;
    }
    if constexpr (RES) {
      vf<V> r = vload<V>(resid + off);
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] += r[e];
    }
    if constexpr (STATS) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        s1[e] += v[e];
        s2[e] = fmaf(v[e], v[e], s2[e]);
      }
    }
    vstore<V>(y + off, v);
  }
  if constexpr (STATS) {
    float* sp = chstats + (size_t)b * 2 * C + ch;
    vstat_add<V>(det, sp, s1);
    vstat_add<V>(det, sp + C, s2);
  }
}

int launch_affine_residual(const float* x, const NormRef& norm, const float* resid, float* y,
                           float* chstats, const DropRef& drop, int B, int L, int C, cudaStream_t st,
                           const DetRef& det) {
  TD_REQUIRE(C % 4 == 0, "affine_residual: C=%d", C);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  const int rows = 8;
  dim3 grid(cdiv(L, rows), cdiv(C / 4, threads), B);
  if (drop.mask || drop.item_mask) {
    if (chstats) {
      TD_LAUNCH((affine_kernel<true, true, true>), grid, threads, 0, st, x, norm, resid, y, chstats, L, C, rows, drop, det);
    } else {
      TD_LAUNCH((affine_kernel<true, false, true>), grid, threads, 0, st, x, norm, resid, y, chstats, L, C, rows, drop, det);
    }
    return 0;
  }
  if (chstats) {
    TD_LAUNCH((affine_kernel<true, true>), grid, threads, 0, st, x, norm, resid, y, chstats, L, C, rows, DropRef{}, det);
  } else {
    TD_LAUNCH((affine_kernel<true, false>), grid, threads, 0, st, x, norm, resid, y, chstats, L, C, rows, DropRef{}, det);
  }
  return 0;
}

int launch_affine(const float* x, const NormRef& norm, float* y, int B, int L, int C, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "affine: C=%d", C);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  const int rows = 16;
  dim3 grid(cdiv(L, rows), cdiv(C / 4, threads), B);
  TD_LAUNCH((affine_kernel<false, false>), grid, threads, 0, st, x, norm, nullptr, y, nullptr, L, C, rows, DropRef{}, DetRef{});
  return 0;
}

// ----------------------------------------------------------------------------- attention core
// nn.MultiheadAttention eval math (packed in_proj already applied): softmax(q k^T / sqrt(d)) v
// per (problem, head).  A problem is a set of `n` tokens token(s) = base + s*stride:
//   batch-axis (BEST / FORK, batch_first=False fed [B,T',C]): tokens of the `group` batch items
//       that share a time index;  time-axis (MULTRES): the L tokens of one batch item.
// Register tiling against the shared-memory pipe (a broadcast LDS.128 still costs four LSU cycles per warp, which
// bounded the one-query-per-thread form at ~57 us per launch): QL lanes share a group of QPT queries; lane j owns
// the float4 segments {j, j+QL, ...} of the head dimension of q, of the score partial sums (completed by
// xor-shuffles inside the lane group) and of the output accumulators, so every K / V segment read from shared
// memory feeds QPT queries.  K/V chunks of <= 64 keys are staged in shared memory; online softmax per chunk of 8.
template <int D, int QL, int QPT>
__global__ void __launch_bounds__(64 * QL / QPT < 32 ? 32 : 64 * QL / QPT) attention_kernel(const float* __restrict__ qkv,
                                                                 float* __restrict__ ctx, int L, int C, int n,
                                                                 int group, int time_axis, int kchunk, int round_out,
                                                                 const uint8_t* __restrict__ amask, float inv_keep) {
  grid_dep_wait();
  constexpr int NS = D / QL / 4;  // float4 segments per lane
  constexpr int DL = NS * 4;
  extern __shared__ float smem[];  // K [kchunk][D], V [kchunk][D]
  float* Ks = smem;
  float* Vs = smem + (size_t)kchunk * D;
  const int head = blockIdx.y;
  const int prob = blockIdx.x;
  long base, stride;
  if (time_axis) {
    base = (long)prob * L;
    stride = 1;
  } else {
    const int grp = prob / L, t = prob % L;
    base = (long)grp * group * L + t;
    stride = L;
  }
  const int gpc = blockDim.x / QL;  // query groups per CTA
  const int lq = threadIdx.x % QL;
  const int q0 = (blockIdx.z * gpc + threadIdx.x / QL) * QPT;  // first query of this thread's group
  const float scale = rsqrtf((float)D);
  const size_t C3 = (size_t)3 * C;
  float q[QPT][DL], o[QPT][DL], m[QPT], l[QPT];
  const uint8_t* mrow[QPT];
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    const bool active = q0 + a < n;
    m[a] = -FLT_MAX;
    l[a] = 0.f;
    // training: dropout on the normalised weights (nn.MultiheadAttention(dropout)); row of this query's keep-mask
    mrow[a] = amask && active ? amask + (((size_t)prob * gridDim.y + head) * n + (q0 + a)) * n : nullptr;
    const float* qp = qkv + (size_t)(base + (long)(active ? q0 + a : 0) * stride) * C3 + head * D;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      float4 v = *reinterpret_cast<const float4*>(qp + (lq + QL * i) * 4);
      if (!active) v = make_float4(0.f, 0.f, 0.f, 0.f);  // inactive queries run along: the shuffles need whole warps
      q[a][4 * i] = v.x * scale; q[a][4 * i + 1] = v.y * scale; q[a][4 * i + 2] = v.z * scale; q[a][4 * i + 3] = v.w * scale;
      o[a][4 * i] = 0.f; o[a][4 * i + 1] = 0.f; o[a][4 * i + 2] = 0.f; o[a][4 * i + 3] = 0.f;
    }
  }

  for (int k0 = 0; k0 < n; k0 += kchunk) {
    const int kn = min(kchunk, n - k0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kn * (D / 4); idx += blockDim.x) {
      const int s = idx / (D / 4), i = (idx % (D / 4)) * 4;
      const float* kp = qkv + (size_t)(base + (long)(k0 + s) * stride) * C3 + C + head * D + i;
      *reinterpret_cast<float4*>(Ks + s * D + i) = *reinterpret_cast<const float4*>(kp);
      *reinterpret_cast<float4*>(Vs + s * D + i) = *reinterpret_cast<const float4*>(kp + C);
    }
    __syncthreads();
    for (int s0 = 0; s0 < kn; s0 += 8) {
      float sc[QPT][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int a = 0; a < QPT; ++a) sc[a][u] = 0.f;
        if (s0 + u < kn) {
          const float* kr = Ks + (s0 + u) * D + lq * 4;
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            const float4 kv = *reinterpret_cast<const float4*>(kr + QL * 4 * i);
#pragma unroll
            for (int a = 0; a < QPT; ++a) {
              sc[a][u] = fmaf(q[a][4 * i], kv.x, sc[a][u]); sc[a][u] = fmaf(q[a][4 * i + 1], kv.y, sc[a][u]);
              sc[a][u] = fmaf(q[a][4 * i + 2], kv.z, sc[a][u]); sc[a][u] = fmaf(q[a][4 * i + 3], kv.w, sc[a][u]);
            }
          }
        }
      }
#pragma unroll
      for (int off = 1; off < QL; off <<= 1) {
#pragma unroll
        for (int a = 0; a < QPT; ++a)
#pragma unroll
          for (int u = 0; u < 8; ++u) sc[a][u] += __shfl_xor_sync(0xffffffffu, sc[a][u], off);
      }
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        float cm = -FLT_MAX;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (s0 + u >= kn) sc[a][u] = -FLT_MAX;
          cm = fmaxf(cm, sc[a][u]);
        }
        if (cm > m[a]) {
          const float f = expf(m[a] - cm);
          l[a] *= f;
#pragma unroll
          for (int i = 0; i < DL; ++i) o[a][i] *= f;
          m[a] = cm;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float p = s0 + u < kn ? expf(sc[a][u] - m[a]) : 0.f;
          l[a] += p;
          if (mrow[a] && s0 + u < kn) p = mrow[a][k0 + s0 + u] ? p * inv_keep : 0.f;
          sc[a][u] = p;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (s0 + u < kn) {
          const float* vr = Vs + (s0 + u) * D + lq * 4;
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            const float4 vv = *reinterpret_cast<const float4*>(vr + QL * 4 * i);
#pragma unroll
            for (int a = 0; a < QPT; ++a) {
              const float p = sc[a][u];
              o[a][4 * i] = fmaf(p, vv.x, o[a][4 * i]); o[a][4 * i + 1] = fmaf(p, vv.y, o[a][4 * i + 1]);
              o[a][4 * i + 2] = fmaf(p, vv.z, o[a][4 * i + 2]); o[a][4 * i + 3] = fmaf(p, vv.w, o[a][4 * i + 3]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    if (q0 + a < n) {
      const float inv = 1.f / l[a];
      float* op = ctx + (size_t)(base + (long)(q0 + a) * stride) * C + head * D;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        float4 r = make_float4(o[a][4 * i] * inv, o[a][4 * i + 1] * inv, o[a][4 * i + 2] * inv, o[a][4 * i + 3] * inv);
        if (round_out) { r.x = tf32_rna(r.x); r.y = tf32_rna(r.y); r.z = tf32_rna(r.z); r.w = tf32_rna(r.w); }
        *reinterpret_cast<float4*>(op + (lq + QL * i) * 4) = r;
      }
    }
  }
}

template <int D, int QPT>
static int launch_attention_q(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                              int time_axis, int round_out, const uint8_t* amask, float inv_keep, cudaStream_t st) {
  constexpr int QL = D >= 16 ? 4 : D / 4;  // lanes per query group, each owning >= one float4 of the head dimension
  const int n = time_axis ? L : group;
  const int nprob = time_axis ? B : (B / group) * L;
  const int gstep = 32 / QL;  // query groups per warp
  const int groups = cdiv(n, QPT);
  const int cap = 64 / QPT > gstep ? 64 / QPT : gstep;               // <= 64 queries per CTA, at least one warp
  const int gpc = cdiv(groups < cap ? groups : cap, gstep) * gstep;  // query groups per CTA (whole warps)
  int kchunk = n < 64 ? n : 64;  // 2*64*D floats <= 32 KB of static-limit shared memory
  kchunk = (kchunk + 7) / 8 * 8;
  dim3 grid(nprob, n_head, cdiv(groups, gpc));
  const size_t smem = (size_t)2 * kchunk * D * sizeof(float);
  TD_LAUNCH((attention_kernel<D, QL, QPT>), grid, gpc * QL, smem, st, qkv, ctx, L, C, n, group, time_axis, kchunk, round_out,
            amask, inv_keep);
  return 0;
}

// ----------------------------------------------------------------------------- attention core on the tensor cores
// 16 < n <= 64 tokens per problem (the inference batches: attention over the batch axis).  The CUDA-core kernel
// above is bound by instruction issue there (39.6 M warp instructions per launch at B = 64, 94 us, ncu
// profiles/r01_ncu_attention.txt); here S = Q K^T and O = P V are warp-level mma.sync.m16n8k8 TF32 with every
// operand split into hi + lo TF32 parts and three MMAs per tile (hi*hi + lo*hi + hi*lo), which keeps fp32-level
// accuracy (the dropped lo*lo term is ~2^-22), so the same kernel serves gemm_mode fp32.
// One CTA per (problem, head): four warps x 16 query rows; K and V are staged once, already split, in shared
// memory rows padded to D + 4 floats (conflict-free fragment gathers).  Whole softmax rows live in the
// accumulator fragments (n <= 64), so there is no online rescaling.  The P fragment feeds the second product
// without a shuffle: the accumulator columns {2t, 2t+1} of a key tile are taken as the A columns {t, t+4}, and
// the V fragment is gathered with the same relabelling of the keys.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  const float h = tf32_rna(x);
  hi = __float_as_uint(h);
  lo = __float_as_uint(tf32_rna(x - h));
}

// RAW: K and V are staged as stored (half the shared memory: four CTAs per SM instead of three, 1008 CTAs of the
// headline shape in two waves instead of three) and split into TF32 hi / lo when a fragment is gathered - the same
// values as the pre-split staging, so both forms give identical bits.
template <int D, bool RAW>
__global__ void __launch_bounds__(128, RAW ? 4 : 1) attention_mma_kernel(const float* __restrict__ qkv, float* __restrict__ ctx,
                                                            int L, int C, int n, int group, int time_axis,
                                                            int round_out, const uint8_t* __restrict__ amask,
                                                            float inv_keep) {
  grid_dep_wait();
  constexpr int NK = 64, LD = D + 4, KT = D / 8, NT = NK / 8;
  extern __shared__ float smem[];
  float* Khi = smem;
  float* Klo = RAW ? Khi : Khi + NK * LD;
  float* Vhi = Klo + NK * LD;
  float* Vlo = RAW ? Vhi : Vhi + NK * LD;
  const int head = blockIdx.y, prob = blockIdx.x;
  long base, stride;
  if (time_axis) {
    base = (long)prob * L;
    stride = 1;
  } else {
    const int grp = prob / L, t = prob % L;
    base = (long)grp * group * L + t;
    stride = L;
  }
  const size_t C3 = (size_t)3 * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // ---- stage K, V (rows >= n are zero), split into TF32 hi / lo
  for (int idx = tid; idx < NK * (D / 4); idx += 128) {
    const int sidx = idx / (D / 4), i = (idx % (D / 4)) * 4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (sidx < n) {
      const float* kp = qkv + (size_t)(base + (long)sidx * stride) * C3 + C + head * D + i;
      kv = *reinterpret_cast<const float4*>(kp);
      vv = *reinterpret_cast<const float4*>(kp + C);
    }
    if constexpr (RAW) {
      *reinterpret_cast<float4*>(Khi + sidx * LD + i) = kv;
      *reinterpret_cast<float4*>(Vhi + sidx * LD + i) = vv;
    } else {
      uint32_t h[4], l[4];
      split_tf32(kv.x, h[0], l[0]); split_tf32(kv.y, h[1], l[1]); split_tf32(kv.z, h[2], l[2]); split_tf32(kv.w, h[3], l[3]);
      *reinterpret_cast<uint4*>(Khi + sidx * LD + i) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(Klo + sidx * LD + i) = make_uint4(l[0], l[1], l[2], l[3]);
      split_tf32(vv.x, h[0], l[0]); split_tf32(vv.y, h[1], l[1]); split_tf32(vv.z, h[2], l[2]); split_tf32(vv.w, h[3], l[3]);
      *reinterpret_cast<uint4*>(Vhi + sidx * LD + i) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(Vlo + sidx * LD + i) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
  // ---- Q fragments of this warp's 16 rows (rows g and g + 8), scaled, split
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const float scale = rsqrtf((float)D);
  uint32_t qh[KT][4], ql[KT][4];
  {
    const float* q0 = qkv + (size_t)(base + (long)(r0 < n ? r0 : 0) * stride) * C3 + head * D;
    const float* q1 = qkv + (size_t)(base + (long)(r1 < n ? r1 : 0) * stride) * C3 + head * D;
    const float m0 = r0 < n ? scale : 0.f, m1 = r1 < n ? scale : 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      split_tf32(q0[8 * kt + t] * m0, qh[kt][0], ql[kt][0]);
      split_tf32(q1[8 * kt + t] * m1, qh[kt][1], ql[kt][1]);
      split_tf32(q0[8 * kt + t + 4] * m0, qh[kt][2], ql[kt][2]);
      split_tf32(q1[8 * kt + t + 4] * m1, qh[kt][3], ql[kt][3]);
    }
  }
  __syncthreads();
  if (warp * 16 >= n) return;  // no barrier below
  // ---- S = Q K^T : NT key tiles of 8, accumulator (row g: c0 c1 | row g+8: c2 c3), columns 2t, 2t+1
  float S[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
    if (nt * 8 < n) {  // uniform
      const uint32_t* kh = reinterpret_cast<const uint32_t*>(Khi) + (nt * 8 + g) * LD + t;
      const uint32_t* kl = reinterpret_cast<const uint32_t*>(Klo) + (nt * 8 + g) * LD + t;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t bh0 = kh[8 * kt], bh1 = kh[8 * kt + 4], bl0 = kl[8 * kt], bl1 = kl[8 * kt + 4];
        if constexpr (RAW) {
          split_tf32(__uint_as_float(bh0), bh0, bl0);
          split_tf32(__uint_as_float(bh1), bh1, bl1);
        }
        mma_tf32_16x8x8(S[nt], ql[kt], bh0, bh1);
        mma_tf32_16x8x8(S[nt], qh[kt], bl0, bl1);
        mma_tf32_16x8x8(S[nt], qh[kt], bh0, bh1);
      }
    }
  }
  // ---- softmax over the keys of rows r0, r1 (a row is spread over the 4 lanes of a quad)
  float mx0 = -FLT_MAX, mx1 = -FLT_MAX;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (c >= n) { S[nt][0] = -FLT_MAX; S[nt][2] = -FLT_MAX; }
    if (c + 1 >= n) { S[nt][1] = -FLT_MAX; S[nt][3] = -FLT_MAX; }
    mx0 = fmaxf(mx0, fmaxf(S[nt][0], S[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(S[nt][2], S[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
  // training: dropout on the normalised weights (nn.MultiheadAttention(dropout)): the sums use the undropped weights
  const uint8_t* mr0 = amask && r0 < n ? amask + (((size_t)prob * gridDim.y + head) * n + r0) * n : nullptr;
  const uint8_t* mr1 = amask && r1 < n ? amask + (((size_t)prob * gridDim.y + head) * n + r1) * n : nullptr;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = nt * 8 + 2 * t;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = c + (e & 1);
      float p = col < n ? expf(S[nt][e] - (e < 2 ? mx0 : mx1)) : 0.f;
      if (e < 2) l0 += p; else l1 += p;
      const uint8_t* mr = e < 2 ? mr0 : mr1;
      if (mr && col < n) p = mr[col] ? p * inv_keep : 0.f;
      S[nt][e] = p;
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // ---- O = P V : key tile j is the k dimension; A columns {t, t+4} <- keys {2t, 2t+1} of the tile
  float O[KT][4];
#pragma unroll
  for (int dt = 0; dt < KT; ++dt) O[dt][0] = O[dt][1] = O[dt][2] = O[dt][3] = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (j * 8 < n) {  // uniform
      uint32_t ph[4], pl[4];
      split_tf32(S[j][0], ph[0], pl[0]);  // (row g,   key 2t)
      split_tf32(S[j][2], ph[1], pl[1]);  // (row g+8, key 2t)
      split_tf32(S[j][1], ph[2], pl[2]);  // (row g,   key 2t+1)
      split_tf32(S[j][3], ph[3], pl[3]);  // (row g+8, key 2t+1)
      const uint32_t* vh = reinterpret_cast<const uint32_t*>(Vhi) + (j * 8 + 2 * t) * LD + g;
      const uint32_t* vl = reinterpret_cast<const uint32_t*>(Vlo) + (j * 8 + 2 * t) * LD + g;
#pragma unroll
      for (int dt = 0; dt < KT; ++dt) {
        uint32_t bh0 = vh[8 * dt], bh1 = vh[LD + 8 * dt], bl0 = vl[8 * dt], bl1 = vl[LD + 8 * dt];
        if constexpr (RAW) {
          split_tf32(__uint_as_float(bh0), bh0, bl0);
          split_tf32(__uint_as_float(bh1), bh1, bl1);
        }
        mma_tf32_16x8x8(O[dt], pl, bh0, bh1);
        mma_tf32_16x8x8(O[dt], ph, bl0, bl1);
        mma_tf32_16x8x8(O[dt], ph, bh0, bh1);
      }
    }
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
  for (int dt = 0; dt < KT; ++dt) {
    float2 a = make_float2(O[dt][0] * i0, O[dt][1] * i0), b = make_float2(O[dt][2] * i1, O[dt][3] * i1);
    if (round_out) { a.x = tf32_rna(a.x); a.y = tf32_rna(a.y); b.x = tf32_rna(b.x); b.y = tf32_rna(b.y); }
    if (r0 < n) *reinterpret_cast<float2*>(ctx + (size_t)(base + (long)r0 * stride) * C + head * D + 8 * dt + 2 * t) = a;
    if (r1 < n) *reinterpret_cast<float2*>(ctx + (size_t)(base + (long)r1 * stride) * C + head * D + 8 * dt + 2 * t) = b;
  }
}

template <int D>
static int launch_attention_mma(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                                int time_axis, int round_out, const uint8_t* amask, float inv_keep, cudaStream_t st) {
  const int n = time_axis ? L : group;
  const int nprob = time_axis ? B : (B / group) * L;
  static const bool raw = !(getenv("TDANET_ATT_RAW") && atoi(getenv("TDANET_ATT_RAW")) == 0);
  const size_t smem = (size_t)(raw ? 2 : 4) * 64 * (D + 4) * sizeof(float);
  static bool attr_set[16] = {};  // the opt-in to > 48 KB of dynamic shared memory is per device
  int dev = 0;
  TD_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || !attr_set[dev]) {
    TD_CUDA(cudaFuncSetAttribute(attention_mma_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * (D + 4) * (int)sizeof(float)));
    if (dev >= 0 && dev < 16) attr_set[dev] = true;
  }
  dim3 grid(nprob, n_head);
  if (raw) {
    TD_LAUNCH((attention_mma_kernel<D, true>), grid, 128, smem, st, qkv, ctx, L, C, n, group, time_axis, round_out, amask, inv_keep);
  } else {
    TD_LAUNCH((attention_mma_kernel<D, false>), grid, 128, smem, st, qkv, ctx, L, C, n, group, time_axis, round_out, amask, inv_keep);
  }
  return 0;
}

template <int D>
static int launch_attention_d(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                              int time_axis, int round_out, const uint8_t* amask, float inv_keep, cudaStream_t st) {
  const int n = time_axis ? L : group;
  // whole softmax rows in tensor-core fragments for the inference batches (attention over <= 64 batch items)
  if constexpr (D % 8 == 0 && D >= 16)
    if (n > 16 && n <= 64)
      return launch_attention_mma<D>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
  // four queries per thread once there are enough of them to fill warps (time-axis attention)
  if (n >= 32) return launch_attention_q<D, 4>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
  return launch_attention_q<D, 1>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
}

int launch_attention(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                     int time_axis, int round_out, const uint8_t* amask, float inv_keep, cudaStream_t st) {
  TD_REQUIRE(C % n_head == 0, "attention: C=%d n_head=%d", C, n_head);
  TD_REQUIRE(time_axis || (group > 0 && B % group == 0), "attention: batch %d not a multiple of group %d", B, group);
  switch (C / n_head) {
    case 64: return launch_attention_d<64>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
    case 32: return launch_attention_d<32>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
    case 16: return launch_attention_d<16>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
    case 8: return launch_attention_d<8>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
    case 4: return launch_attention_d<4>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, amask, inv_keep, st);
  }
  return fail(TDANET_EUNSUPPORTED, "attention: head dim %d not in {4,8,16,32,64}", C / n_head);
}

}  // namespace td
