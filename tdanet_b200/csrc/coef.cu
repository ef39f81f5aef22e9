// GlobLN finalisation: turn accumulated sums into per-(item, channel) scale/shift tables.
//
// GlobLN (TDANet_best.py:47-64; nn.GroupNorm(1,C,eps=1e-8) in TDANet.py:59-60) is
//   y = gamma_c * (x - mu_b) / sqrt(var_b + 1e-8) + beta_c,  mu/var over all (C, T) of item b.
// Producers accumulate sum / sum-of-squares; these tiny kernels (one CTA per batch item, all
// arithmetic in double) emit scale = gamma*r and shift = beta - gamma*mu*r so that consumers
// normalise on load with one FMA.
#include "kernels.h"

namespace td {

__device__ __forceinline__ void moments(double s, double ss, double n, double& mu, double& r) {
  mu = s / n;
  double var = ss / n - mu * mu;
  if (var < 0.0) var = 0.0;
  r = rsqrt(var + (double)kEpsGLN);
}

__global__ void coef_item_kernel(const double* __restrict__ stats, double count,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ coef, int C) {
  const int b = blockIdx.x;
  double mu, r;
  moments(stats[2 * b], stats[2 * b + 1], count, mu, r);
  float* o = coef + (size_t)b * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double g = gamma[c];
    o[c] = (float)(g * r);
    o[C + c] = (float)((double)beta[c] - g * mu * r);
  }
}

int launch_coef_item(const double* stats, double count, const float* gamma, const float* beta,
                     float* coef, int B, int C, cudaStream_t st) {
  TD_LAUNCH(coef_item_kernel, B, 128, 0, st, stats, count, gamma, beta, coef, C);
  return 0;
}

// sum over channels of per-channel (sum, sumsq) -> mu, r; valid in all threads
__device__ __forceinline__ void chan_moments(const float* __restrict__ s1, const float* __restrict__ s2,
                                             int C, double n, double* sh, double& mu, double& r) {
  double a = 0.0, q = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    a += (double)s1[c];
    q += (double)s2[c];
  }
  block_sum2(a, q, sh);
  moments(a, q, n, mu, r);
}

__global__ void coef_chan_kernel(const float* __restrict__ chstats, size_t item_stride, int rows,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ coef, int C) {
  __shared__ double sh[64];
  const int b = blockIdx.x;
  const float* s = chstats + (size_t)b * item_stride;
  double mu, r;
  chan_moments(s, s + C, C, (double)rows * C, sh, mu, r);
  float* o = coef + (size_t)b * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double g = gamma[c];
    o[c] = (float)(g * r);
    o[C + c] = (float)((double)beta[c] - g * mu * r);
  }
}

int launch_coef_chan(const float* chstats, size_t item_stride, int rows, const float* gamma,
                     const float* beta, float* coef, int B, int C, cudaStream_t st) {
  TD_LAUNCH(coef_chan_kernel, B, 256, 0, st, chstats, item_stride, rows, gamma, beta, coef, C);
  return 0;
}

__global__ void coef_la_kernel(const float* __restrict__ stats_l, int Ll,
                               const float* __restrict__ stats_g, int Lg, tdanet_la_t la,
                               float* __restrict__ coef, int C) {
  __shared__ double sh[64];
  const int b = blockIdx.x;
  const float* sl = stats_l + (size_t)b * 2 * C;
  const float* sa = stats_g + (size_t)b * 4 * C;  // conv 0 = global_act
  const float* se = sa + 2 * C;                   // conv 1 = global_embedding
  double muL, rL, muA, rA, muE, rE;
  chan_moments(sl, sl + C, C, (double)Ll * C, sh, muL, rL);
  chan_moments(sa, sa + C, C, (double)Lg * C, sh, muA, rA);
  chan_moments(se, se + C, C, (double)Lg * C, sh, muE, rE);
  float* o = coef + (size_t)b * 6 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double g = la.local_embedding.gamma[c];
    o[c] = (float)(g * rL);
    o[C + c] = (float)((double)la.local_embedding.beta[c] - g * muL * rL);
    g = la.global_act.gamma[c];
    o[2 * C + c] = (float)(g * rA);
    o[3 * C + c] = (float)((double)la.global_act.beta[c] - g * muA * rA);
    g = la.global_embedding.gamma[c];
    o[4 * C + c] = (float)(g * rE);
    o[5 * C + c] = (float)((double)la.global_embedding.beta[c] - g * muE * rE);
  }
}

int launch_coef_la(const float* stats_l, int Ll, const float* stats_g, int Lg, const tdanet_la_t* la,
                   float* coef, int B, int C, cudaStream_t st) {
  TD_LAUNCH(coef_la_kernel, B, 256, 0, st, stats_l, Ll, stats_g, Lg, *la, coef, C);
  return 0;
}

// BEST loc_glo_fus[k] = LA with 1-tap depthwise "convs" (TDANet_best.py:329-331): every GlobLN
// input is a per-channel affine map of a tensor whose per-channel sums are known, so all three
// normalisations have closed-form statistics and the whole LA collapses to
//   (al*O + bl) * sigmoid(aa*g + ba) + (ae*g + be)        per (item, channel).
__global__ void coef_inject_gate_kernel(const float* __restrict__ spp_stats, size_t spp_item_stride,
                                        int Lk, tdanet_convnorm_t spp,
                                        const float* __restrict__ g_stats, int Lg, tdanet_la_t la,
                                        float* __restrict__ coef, int C) {
  __shared__ double sh[64];
  const int b = blockIdx.x;
  const float* s1 = spp_stats + (size_t)b * spp_item_stride;
  const float* s2 = s1 + C;
  const float* g1 = g_stats + (size_t)b * 2 * C;
  const float* g2 = g1 + C;
  // 1. GlobLN of spp_dw[k]: x_l = a_c*O + b_c
  double mu, r;
  chan_moments(s1, s2, C, (double)Lk * C, sh, mu, r);
  // 2. z = w_l * x_l = p*O + q ; statistics of z from the per-channel sums of O
  double zs = 0.0, zq = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double gm = spp.gamma[c];
    const double wl = la.local_embedding.w[c];
    const double p = wl * gm * r, q = wl * ((double)spp.beta[c] - gm * mu * r);
    const double S1 = s1[c], S2 = s2[c];
    zs += p * S1 + q * Lk;
    zq += p * p * S2 + 2.0 * p * q * S1 + q * q * Lk;
  }
  block_sum2(zs, zq, sh);
  double muz, rz;
  moments(zs, zq, (double)Lk * C, muz, rz);
  // 3. global branches: z = w * g
  double as = 0.0, aq = 0.0, es = 0.0, eq = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double wa = la.global_act.w[c], we = la.global_embedding.w[c];
    const double G1 = g1[c], G2 = g2[c];
    as += wa * G1;
    aq += wa * wa * G2;
    es += we * G1;
    eq += we * we * G2;
  }
  block_sum2(as, aq, sh);
  block_sum2(es, eq, sh);
  double mua, ra, mue, re;
  moments(as, aq, (double)Lg * C, mua, ra);
  moments(es, eq, (double)Lg * C, mue, re);
  float* o = coef + (size_t)b * 6 * C;
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

int launch_coef_inject_gate(const float* spp_stats, size_t spp_item_stride, int Lk,
                            const tdanet_convnorm_t* spp, const float* g_stats, int Lg,
                            const tdanet_la_t* la, float* coef, int B, int C, cudaStream_t st) {
  TD_LAUNCH(coef_inject_gate_kernel, B, 256, 0, st, spp_stats, spp_item_stride, Lk, *spp, g_stats, Lg,
            *la, coef, C);
  return 0;
}

}  // namespace td
