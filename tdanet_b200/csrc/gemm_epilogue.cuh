// Epilogues shared by the CUDA-core and the tcgen05 GEMM.
#pragma once
#include "kernels.h"

namespace td {

struct Epilogue {
  const GemmArgs& a;
  size_t row_base;  // first row of this batch item in [B*L]
  float cslope;
  bool vec;

  __device__ __forceinline__ Epilogue(const GemmArgs& args, int b) : a(args) {
    row_base = (size_t)b * a.L;
    cslope = (a.epi == EPI_RESIDUAL && !a.last) ? __ldg(a.cslope) : 0.f;
    vec = (a.N % 4) == 0;
  }

  // v: accumulators of D[r, n..n+3] (columns beyond N are ignored); accumulates the sum and the
  // sum of squares of what is stored into s1/s2 when statistics were requested.
  __device__ __forceinline__ void apply4(int r, int n, float (&v)[4], float& s1, float& s2) const {
    const size_t row = row_base + r;
    const int cnt = min(4, a.N - n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < cnt) {
        float y = v[j] + (a.bias ? __ldg(a.bias + n + j) : 0.f);
        if (a.epi == EPI_RESIDUAL) {
          // res_conv(expanded) + residual (TDANet_best.py:380), then for every block but the last
          // concat_block(mixture + x) = PReLU(w_c*(mixture + x) + b_c) (TDANet_best.py:388-398)
          y += a.resid[row * a.N + n + j];
          if (!a.last)
            y = preluf_(fmaf(__ldg(a.cw + n + j), a.mix[row * a.N + n + j] + y, __ldg(a.cb + n + j)), cslope);
        } else if (a.epi == EPI_MASK) {
          // mask_nl_class(mask) * encoder output (TDANet_best.py:507-509)
          y = fmaxf(y, 0.f) * a.enc[row * a.Nb + (n + j) % a.Nb];
        }
        v[j] = y;
        if (a.stats) {
          s1 += y;
          s2 = fmaf(y, y, s2);
        }
      }
    }
    float* d = a.D + row * a.N + n;
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
