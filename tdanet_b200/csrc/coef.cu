// Closed-form coefficients of BEST's loc_glo_fus (LA with 1-tap depthwise "convs", TDANet_best.py:329-331).
//
// Every GlobLN input there is a per-channel affine map of a tensor whose per-channel sums are known,
// so all three normalisations have closed-form statistics and the whole LA collapses to
//   x_fused[k] = (al*O_k + bl) * sigmoid(aa*g + ba) + (ae*g + be)        per (item, channel),
// which the LA kernels recompute on load (SRC_INJECT_GATE) instead of materialising x_fused.
// One launch covers every scale: grid (B, depth), all arithmetic in double.
//
// (All other GlobLNs need no kernel: producers accumulate per-item sums and consumers derive
// scale/shift themselves, see NormRef in common.cuh.)
#include "kernels.h"

namespace td {

__device__ __forceinline__ void moments(double s, double ss, double n, double& mu, double& r) {
  mu = s / n;
  double var = ss / n - mu * mu;
  if (var < 0.0) var = 0.0;
  r = 1.0 / sqrt(var + (double)kEpsGLN);
}

__global__ void __launch_bounds__(256) coef_inject_gate_kernel(InjectCoefArgs a, int C) {
  grid_dep_wait();
  __shared__ double sh[64];
  const int b = blockIdx.x, k = blockIdx.y;
  const int Lk = a.L[k];
  const tdanet_convnorm_t& spp = a.spp[k];
  const tdanet_la_t& la = a.la[k];
  const float* s1 = a.spp_stats[k] + (size_t)b * 2 * C;
  const float* s2 = s1 + C;
  const float* g1 = a.g_stats + (size_t)b * 2 * C;
  const float* g2 = g1 + C;
  // 1. GlobLN of spp_dw[k]: x_l = a_c*O + b_c
  double t1 = 0.0, t2 = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    t1 += (double)s1[c];
    t2 += (double)s2[c];
  }
  block_sum2(t1, t2, sh);
  double mu, r;
  moments(t1, t2, (double)Lk * C, mu, r);
  // 2. z = w_l * x_l = p*O + q ; statistics of z from the per-channel sums of O
  // 3. global branches: z = w * g
  double zs = 0.0, zq = 0.0, as = 0.0, aq = 0.0, es = 0.0, eq = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double gm = spp.gamma[c];
    const double wl = la.local_embedding.w[c];
    const double p = wl * gm * r, q = wl * ((double)spp.beta[c] - gm * mu * r);
    const double S1 = s1[c], S2 = s2[c];
    zs += p * S1 + q * Lk;
    zq += p * p * S2 + 2.0 * p * q * S1 + q * q * Lk;
    const double wa = la.global_act.w[c], we = la.global_embedding.w[c];
    const double G1 = g1[c], G2 = g2[c];
    as += wa * G1;
    aq += wa * wa * G2;
    es += we * G1;
    eq += we * we * G2;
  }
  block_sum2(zs, zq, sh);
  block_sum2(as, aq, sh);
  block_sum2(es, eq, sh);
  double muz, rz, mua, ra, mue, re;
  moments(zs, zq, (double)Lk * C, muz, rz);
  moments(as, aq, (double)a.Lg * C, mua, ra);
  moments(es, eq, (double)a.Lg * C, mue, re);
  if (a.conv_stats[k] && threadIdx.x == 0) {
    // sums / sums of squares of the three 1-tap conv outputs, in the NormRef layout, for the backward pass
    double* cs = a.conv_stats[k] + (size_t)b * 6;
    cs[0] = zs; cs[1] = zq; cs[2] = as; cs[3] = aq; cs[4] = es; cs[5] = eq;
  }
  float* o = a.coef[k] + (size_t)b * 6 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double gm = spp.gamma[c];
    const double wl = la.local_embedding.w[c];
    const double p = wl * gm * r, q = wl * ((double)spp.beta[c] - gm * mu * r);
    const double gl = la.local_embedding.gamma[c];
    o[c] = (float)(gl * rz * p);
    o[C + c] = (float)(gl * rz * (q - muz) + (double)la.local_embedding.beta[c]);
    const double ga = la.global_act.gamma[c], wa = la.global_act.w[c];
    o[2 * C + c] = (float)(ga * ra * wa);
    o[3 * C + c] = (float)((double)la.global_act.beta[c] - ga * ra * mua);
    const double ge = la.global_embedding.gamma[c], we = la.global_embedding.w[c];
    o[4 * C + c] = (float)(ge * re * we);
    o[5 * C + c] = (float)((double)la.global_embedding.beta[c] - ge * re * mue);
  }
}

// wT[k, c] = w[c, k]: depthwise weights [C][ks] -> [ks][C], so that a thread's 4 channels of one tap are one vector load
__global__ void weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wT, int C, int ks) {
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * ks) return;
  const int k = i / C, c = i % C;
  wT[i] = w[(size_t)c * ks + k];
}

int launch_weight_transpose(const float* w, float* wT, int C, int ks, cudaStream_t st) {
  TD_LAUNCH(weight_transpose_kernel, cdiv(C * ks, 256), 256, 0, st, w, wT, C, ks);
  return 0;
}

int launch_coef_inject_gate(const InjectCoefArgs& a, int B, int C, cudaStream_t st) {
  dim3 grid(B, a.n);
  TD_LAUNCH(coef_inject_gate_kernel, grid, 256, 0, st, a, C);
  return 0;
}

}  // namespace td
