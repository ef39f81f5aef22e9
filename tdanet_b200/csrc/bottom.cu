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

int launch_affine_sum(const PoolArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "affine_sum: C=%d", a.C);
  int threads = a.C / 4 > 256 ? 256 : (a.C / 4 < 32 ? 32 : a.C / 4);
  dim3 grid(a.Lb, cdiv(a.C / 4, threads), a.B);
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

// y = resid + LN(k1*a + k2*xin)*w + b
__global__ void ln_residual_kernel(const float* __restrict__ a, const float* __restrict__ xin,
                                   const float* __restrict__ resid, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ y, float k1,
                                   float k2, int rows, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
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
    o.x = r.x + ((v.x - mu) * rstd * g.x + bb.x);
    o.y = r.y + ((v.y - mu) * rstd * g.y + bb.y);
    o.z = r.z + ((v.z - mu) * rstd * g.z + bb.z);
    o.w = r.w + ((v.w - mu) * rstd * g.w + bb.w);
    *reinterpret_cast<float4*>(y + (size_t)row * C + c) = o;
  }
}

int launch_ln_residual(const float* a, const float* xin, const float* resid, const float* w,
                       const float* b, float* y, int doubled, int B, int L, int C, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "ln_residual: C=%d", C);
  const int rows = B * L;
  // doubled: LayerNorm(out + dropout(out)) == LN(2*out) in eval (TDANet_best.py:251)
  const float k1 = doubled ? 2.f : 1.f, k2 = doubled ? 0.f : 1.f;
  TD_LAUNCH(ln_residual_kernel, cdiv(rows, 8), 256, 0, st, a, doubled ? nullptr : xin, resid, w, b, y,
            k1, k2, rows, C);
  return 0;
}

// ----------------------------------------------------------------------------- affine glue
template <bool RES, bool STATS>
__global__ void affine_kernel(const float* __restrict__ x, NormRef norm,
                              const float* __restrict__ resid, float* __restrict__ y,
                              float* __restrict__ chstats, int L, int C, int rows_per_cta) {
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(t0 + rows_per_cta, L);
  vf<V> sc, sh;
  norm_coef<V>(norm, b, ch, sc, sh);
  vf<V> s1 = vzero<V>(), s2 = vzero<V>();
  for (int t = t0; t < t1; ++t) {
    const size_t off = ((size_t)b * L + t) * C + ch;
    vf<V> v = vload<V>(x + off);
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
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
    vred_add<V>(sp, s1);
    vred_add<V>(sp + C, s2);
  }
}

int launch_affine_residual(const float* x, const NormRef& norm, const float* resid, float* y,
                           float* chstats, int B, int L, int C, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "affine_residual: C=%d", C);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  const int rows = 8;
  dim3 grid(cdiv(L, rows), cdiv(C / 4, threads), B);
  if (chstats) {
    TD_LAUNCH((affine_kernel<true, true>), grid, threads, 0, st, x, norm, resid, y, chstats, L, C, rows);
  } else {
    TD_LAUNCH((affine_kernel<true, false>), grid, threads, 0, st, x, norm, resid, y, chstats, L, C, rows);
  }
  return 0;
}

int launch_affine(const float* x, const NormRef& norm, float* y, int B, int L, int C, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0, "affine: C=%d", C);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  const int rows = 16;
  dim3 grid(cdiv(L, rows), cdiv(C / 4, threads), B);
  TD_LAUNCH((affine_kernel<false, false>), grid, threads, 0, st, x, norm, nullptr, y, nullptr, L, C, rows);
  return 0;
}

// ----------------------------------------------------------------------------- attention core
// nn.MultiheadAttention eval math (packed in_proj already applied): softmax(q k^T / sqrt(d)) v
// per (problem, head).  A problem is a set of `n` tokens token(s) = base + s*stride:
//   batch-axis (BEST / FORK, batch_first=False fed [B,T',C]): tokens of the `group` batch items
//       that share a time index;  time-axis (MULTRES): the L tokens of one batch item.
// One thread per query: q and the output accumulator live in registers, K/V chunks are staged in
// shared memory and read as warp-wide broadcasts; online softmax over key chunks of 8.
template <int D>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv,
                                                        float* __restrict__ ctx, int L, int C, int n,
                                                        int group, int time_axis, int kchunk, int round_out) {
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
  const int q0 = blockIdx.z * blockDim.x;
  const int qi = q0 + threadIdx.x;
  const bool active = qi < n;
  const float scale = rsqrtf((float)D);
  const size_t C3 = (size_t)3 * C;
  float q[D], o[D];
  if (active) {
    const float* qp = qkv + (size_t)(base + (long)qi * stride) * C3 + head * D;
#pragma unroll
    for (int i = 0; i < D; i += 4) {
      float4 v = *reinterpret_cast<const float4*>(qp + i);
      q[i] = v.x * scale; q[i + 1] = v.y * scale; q[i + 2] = v.z * scale; q[i + 3] = v.w * scale;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) o[i] = 0.f;
  float m = -FLT_MAX, l = 0.f;

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
    if (active) {
      for (int s0 = 0; s0 < kn; s0 += 8) {
        float sc[8];
        float cm = -FLT_MAX;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (s0 + u < kn) {
            const float* kr = Ks + (s0 + u) * D;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < D; i += 4) {
              float4 kv = *reinterpret_cast<const float4*>(kr + i);
              acc = fmaf(q[i], kv.x, acc); acc = fmaf(q[i + 1], kv.y, acc);
              acc = fmaf(q[i + 2], kv.z, acc); acc = fmaf(q[i + 3], kv.w, acc);
            }
            sc[u] = acc;
            cm = fmaxf(cm, acc);
          } else {
            sc[u] = -FLT_MAX;
          }
        }
        if (cm > m) {
          const float f = expf(m - cm);
          l *= f;
#pragma unroll
          for (int i = 0; i < D; ++i) o[i] *= f;
          m = cm;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (s0 + u < kn) {
            const float p = expf(sc[u] - m);
            l += p;
            const float* vr = Vs + (s0 + u) * D;
#pragma unroll
            for (int i = 0; i < D; i += 4) {
              float4 vv = *reinterpret_cast<const float4*>(vr + i);
              o[i] = fmaf(p, vv.x, o[i]); o[i + 1] = fmaf(p, vv.y, o[i + 1]);
              o[i + 2] = fmaf(p, vv.z, o[i + 2]); o[i + 3] = fmaf(p, vv.w, o[i + 3]);
            }
          }
        }
      }
    }
  }
  if (active) {
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = round_out ? tf32_rna(o[i] * inv) : o[i] * inv;
    float* op = ctx + (size_t)(base + (long)qi * stride) * C + head * D;
#pragma unroll
    for (int i = 0; i < D; i += 4)
      *reinterpret_cast<float4*>(op + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
  }
}

template <int D>
static int launch_attention_d(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                              int time_axis, int round_out, cudaStream_t st) {
  const int n = time_axis ? L : group;
  const int nprob = time_axis ? B : (B / group) * L;
  int threads = n >= 128 ? 128 : (n + 31) / 32 * 32;
  int kchunk = n < 64 ? n : 64;  // 2*64*D floats <= 32 KB of static-limit shared memory
  kchunk = (kchunk + 7) / 8 * 8;
  dim3 grid(nprob, n_head, cdiv(n, threads));
  const size_t smem = (size_t)2 * kchunk * D * sizeof(float);
  TD_LAUNCH((attention_kernel<D>), grid, threads, smem, st, qkv, ctx, L, C, n, group, time_axis, kchunk, round_out);
  return 0;
}

int launch_attention(const float* qkv, float* ctx, int B, int L, int C, int n_head, int group,
                     int time_axis, int round_out, cudaStream_t st) {
  TD_REQUIRE(C % n_head == 0, "attention: C=%d n_head=%d", C, n_head);
  TD_REQUIRE(time_axis || (group > 0 && B % group == 0), "attention: batch %d not a multiple of group %d", B, group);
  switch (C / n_head) {
    case 64: return launch_attention_d<64>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, st);
    case 32: return launch_attention_d<32>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, st);
    case 16: return launch_attention_d<16>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, st);
    case 8: return launch_attention_d<8>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, st);
    case 4: return launch_attention_d<4>(qkv, ctx, B, L, C, n_head, group, time_axis, round_out, st);
  }
  return fail(TDANET_EUNSUPPORTED, "attention: head dim %d not in {4,8,16,32,64}", C / n_head);
}

}  // namespace td
