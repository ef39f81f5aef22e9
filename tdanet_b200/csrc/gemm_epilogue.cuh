// Epilogues shared by the CUDA-core and the tcgen05 GEMM.
//
// Usage: one `Epilogue` per (kernel, batch item); `cols(n)` caches the per-column constants of four
// consecutive output columns, `row(...)` finishes and stores D[r, n..n+3].  Callers arrange that the
// lanes of a warp hold consecutive column groups of the same row so that every global access
// (residual / mixture reads, D store) is a coalesced 128-bit access.
#pragma once
#include "kernels.h"

namespace td {

struct Epilogue {
  const GemmArgs& a;
  size_t row_base;  // first row of this batch item in [B*L]
  float cslope;
  bool vec;
  // per-column constants of the current column group
  int n, cnt;
  float bias[4], cw[4], cb[4];

  __device__ __forceinline__ Epilogue(const GemmArgs& args, int b) : a(args) {
    row_base = (size_t)b * a.L;
    cslope = (a.epi == EPI_RESIDUAL && !a.last) ? __ldg(a.cslope) : 0.f;
    vec = (a.N % 4) == 0;
    n = 0;
    cnt = 0;
  }

  __device__ __forceinline__ void cols(int n_) {
    n = n_;
    cnt = min(4, a.N - n_);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = j < cnt;
      bias[j] = (ok && a.bias) ? __ldg(a.bias + n + j) : 0.f;
      if (a.epi == EPI_RESIDUAL && !a.last) {
        cw[j] = ok ? __ldg(a.cw + n + j) : 0.f;
        cb[j] = ok ? __ldg(a.cb + n + j) : 0.f;
      }
    }
  }

  // v: accumulators of D[r, n..n+3]; adds what is stored to s1 (sum) / s2 (sum of squares) when
  // statistics were requested.
  __device__ __forceinline__ void row(int r, float (&v)[4], float& s1, float& s2) const {
    const size_t off = (row_base + r) * a.N + n;
    if (cnt <= 0) return;
    float res[4] = {0.f, 0.f, 0.f, 0.f}, mx[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.epi == EPI_RESIDUAL) {
      if (vec && cnt == 4) {
        const float4 t = *reinterpret_cast<const float4*>(a.resid + off);
        res[0] = t.x; res[1] = t.y; res[2] = t.z; res[3] = t.w;
        if (!a.last) {
          const float4 m = *reinterpret_cast<const float4*>(a.mix + off);
          mx[0] = m.x; mx[1] = m.y; mx[2] = m.z; mx[3] = m.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cnt) {
            res[j] = a.resid[off + j];
            if (!a.last) mx[j] = a.mix[off + j];
          }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < cnt) {
        float y = v[j] + bias[j];
        if (a.epi == EPI_RESIDUAL) {
          // res_conv(expanded) + residual (TDANet_best.py:380), then for every block but the last
          // concat_block(mixture + x) = PReLU(w_c*(mixture + x) + b_c) (TDANet_best.py:388-398)
          y += res[j];
          if (!a.last) y = preluf_(fmaf(cw[j], mx[j] + y, cb[j]), cslope);
        } else if (a.epi == EPI_MASK) {
          // mask_nl_class(mask) * encoder output (TDANet_best.py:507-509)
          y = fmaxf(y, 0.f) * a.enc[(row_base + r) * a.Nb + (n + j) % a.Nb];
        }
        v[j] = y;
        if (a.stats) {
          s1 += y;
          s2 = fmaf(y, y, s2);
        }
      }
    }
    float* d = a.D + off;
    if (vec && cnt == 4) {
      *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < cnt) d[j] = v[j];
    }
  }
};

}  // namespace td
