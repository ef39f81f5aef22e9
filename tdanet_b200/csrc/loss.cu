// Fused PIT (SI-)SDR / SNR loss, forward and backward.
//
// Reference: PairwiseNegSDR.forward (losses/matrix.py:21-56) and PITLossWrapper.forward with
// pit_from="pw_mtx" -> find_best_perm_factorial (losses/pit_wrapper.py:29-67, 106-131).
//
// Kernel 1 (one CTA per batch item): means, centred energies and the noise energies of every
// (estimate, target) pair in three sweeps over the item (each quantity is summed the way the
// reference forms it - centred first, residual formed element-wise - so there is no cancellation),
// the pairwise matrix, the best permutation, and the two coefficients (ca, cb) such that
//     d pw[i, j] / d est_i = ca * (est_i - mean) + cb * (tgt_j - mean).
// Kernel 2: the batch-level threshold (`threshold_byloss`: drop items whose loss is <= -30 dB if any
// item survives), the mean, and grad_est written in one pass.  No host synchronisation anywhere
// (the reference syncs at pit_wrapper.py:60).
#include "kernels.h"
#include <cfloat>

namespace td {

constexpr int kMaxSrc = 3;
__device__ const int kPerms3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
__device__ const int kPerms2[2][2] = {{0, 1}, {1, 0}};
// per item: min_loss, then per estimate: ca, cb, mean_e, mean_t, matched target
constexpr int kItemFloats = 1 + kMaxSrc * 5;

template <int NS>
__device__ __forceinline__ void block_reduce_n(double (&v)[NS * NS + 2 * NS], double* sh) {
  constexpr int N = NS * NS + 2 * NS;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (l == 0)
#pragma unroll
    for (int i = 0; i < N; ++i) sh[i * 32 + w] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int k = 0; k < nw; ++k) s += sh[i * 32 + k];
    v[i] = s;
  }
}

template <int NS>
__global__ void __launch_bounds__(512) pit_item_kernel(const float* __restrict__ est,
                                                       const float* __restrict__ tgt, int T,
                                                       int sdr_type, float* __restrict__ pw_out,
                                                       int32_t* __restrict__ perm_out,
                                                       float* __restrict__ scratch) {
  grid_dep_wait();
  constexpr int N = NS * NS + 2 * NS;
  __shared__ double sh[N * 32];
  const int b = blockIdx.x;
  const float* e = est + (size_t)b * NS * T;
  const float* t = tgt + (size_t)b * NS * T;
  const double eps = 1e-8;
  double v[N];

  // sweep 1: means
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = 0.0;
  {
    float pe[NS], pt[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) pe[i] = pt[i] = 0.f;
    for (int x = threadIdx.x; x < T; x += blockDim.x)
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        pe[i] += e[(size_t)i * T + x];
        pt[i] += t[(size_t)i * T + x];
      }
#pragma unroll
    for (int i = 0; i < NS; ++i) { v[i] = pe[i]; v[NS + i] = pt[i]; }
  }
  block_reduce_n<NS>(v, sh);
  float me[NS], mt[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) { me[i] = (float)(v[i] / T); mt[i] = (float)(v[NS + i] / T); }

  // sweep 2: centred energies, dots (sisdr/sdsdr) or direct noise energies (snr/sdsdr)
  double Ee[NS], Et[NS], X[NS][NS];  // X = dot (sisdr) | noise energy (snr, sdsdr)
  double DOT[NS][NS];
  {
    float pe[NS], pt[NS], px[NS][NS], pd[NS][NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      pe[i] = pt[i] = 0.f;
#pragma unroll
      for (int j = 0; j < NS; ++j) px[i][j] = pd[i][j] = 0.f;
    }
    for (int x = threadIdx.x; x < T; x += blockDim.x) {
      float ev[NS], tv[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) { ev[i] = e[(size_t)i * T + x] - me[i]; tv[i] = t[(size_t)i * T + x] - mt[i]; }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        pe[i] = fmaf(ev[i], ev[i], pe[i]);
        pt[i] = fmaf(tv[i], tv[i], pt[i]);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          pd[i][j] = fmaf(ev[i], tv[j], pd[i][j]);
          const float d = ev[i] - tv[j];
          px[i][j] = fmaf(d, d, px[i][j]);
        }
      }
    }
    // two reductions of N values: (Ee, Et, dots) then (noise energies)
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      v[i] = pe[i]; v[NS + i] = pt[i];
#pragma unroll
      for (int j = 0; j < NS; ++j) v[2 * NS + i * NS + j] = pd[i][j];
    }
    block_reduce_n<NS>(v, sh);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      Ee[i] = v[i]; Et[i] = v[NS + i];
#pragma unroll
      for (int j = 0; j < NS; ++j) DOT[i][j] = v[2 * NS + i * NS + j];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) v[2 * NS + i * NS + j] = px[i][j];
    block_reduce_n<NS>(v, sh);
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) X[i][j] = v[2 * NS + i * NS + j];
  }

  // sweep 3 (sisdr only): residual of the projection, formed element-wise
  if (sdr_type == 1) {
    float al[NS][NS], px[NS][NS];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) { al[i][j] = (float)(DOT[i][j] / (Et[j] + eps)); px[i][j] = 0.f; }
    for (int x = threadIdx.x; x < T; x += blockDim.x) {
      float ev[NS], tv[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) { ev[i] = e[(size_t)i * T + x] - me[i]; tv[i] = t[(size_t)i * T + x] - mt[i]; }
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const float d = fmaf(-al[i][j], tv[j], ev[i]);
          px[i][j] = fmaf(d, d, px[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) v[2 * NS + i * NS + j] = px[i][j];
    block_reduce_n<NS>(v, sh);
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) X[i][j] = v[2 * NS + i * NS + j];
  }

  if (threadIdx.x != 0) return;
  const double k10 = 10.0 / log(10.0);
  double pw[NS][NS], ca[NS][NS], cb[NS][NS];
#pragma unroll
  for (int i = 0; i < NS; ++i)
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const double En = X[i][j];
      double Ep, dEp_t;      // signal energy and d Ep / d e = dEp_t * t'
      double dEn_e, dEn_t;   // d En / d e = dEn_e * e' + dEn_t * t'
      if (sdr_type == 0) {   // snr: proj = t', noise = e' - t'
        Ep = Et[j]; dEp_t = 0.0; dEn_e = 2.0; dEn_t = -2.0;
      } else {
        const double en = Et[j] + eps, dot = DOT[i][j];
        Ep = dot * dot * Et[j] / (en * en);
        dEp_t = 2.0 * dot * Et[j] / (en * en);
        if (sdr_type == 1) { // sisdr: noise = e' - proj
          dEn_e = 2.0; dEn_t = -4.0 * dot / en + 2.0 * dot * Et[j] / (en * en);
        } else {             // sdsdr: noise = e' - t'
          dEn_e = 2.0; dEn_t = -2.0;
        }
      }
      const double R = Ep / (En + eps);
      pw[i][j] = -k10 * log(R + eps);
      const double dR_dEp = 1.0 / (En + eps), dR_dEn = -Ep / ((En + eps) * (En + eps));
      const double g = -k10 / (R + eps);
      ca[i][j] = g * dR_dEn * dEn_e;
      cb[i][j] = g * (dR_dEp * dEp_t + dR_dEn * dEn_t);
      if (pw_out) pw_out[((size_t)b * NS + i) * NS + j] = (float)pw[i][j];
    }
  // best permutation: loss_p = mean_j pw[perm_p[j]][j]; first minimum wins (torch.min)
  constexpr int NP = NS == 2 ? 2 : 6;
  int best = 0;
  float best_loss = FLT_MAX;
  for (int p = 0; p < NP; ++p) {
    float s = 0.f;
    for (int j = 0; j < NS; ++j) {
      const int i = NS == 2 ? kPerms2[p][j] : kPerms3[p][j];
      s += (float)pw[i][j];
    }
    s /= (float)NS;
    if (s < best_loss) { best_loss = s; best = p; }
  }
  float* sc = scratch + (size_t)b * kItemFloats;
  sc[0] = best_loss;
  for (int j = 0; j < NS; ++j) {
    const int i = NS == 2 ? kPerms2[best][j] : kPerms3[best][j];
    if (perm_out) perm_out[(size_t)b * NS + j] = i;
    float* q = sc + 1 + i * 5;
    q[0] = (float)ca[i][j]; q[1] = (float)cb[i][j]; q[2] = me[i]; q[3] = mt[j]; q[4] = (float)j;
  }
}

template <int NS>
__global__ void __launch_bounds__(256) pit_finish_kernel(const float* __restrict__ est,
                                                         const float* __restrict__ tgt, int B, int T,
                                                         int threshold, const float* __restrict__ scratch,
                                                         float* __restrict__ loss, float* __restrict__ grad) {
  grid_dep_wait();
  __shared__ double sh[64];
  // batch-level reduction, recomputed by every CTA (B is small)
  double kept_sum = 0.0, kept_cnt = 0.0, all_sum = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float m = scratch[(size_t)i * kItemFloats];
    all_sum += m;
    if (m > -30.f) { kept_sum += m; kept_cnt += 1.0; }
  }
  block_sum2(kept_sum, kept_cnt, sh);
  double dummy = 0.0;
  block_sum2(all_sum, dummy, sh);
  const bool filter = threshold && kept_cnt > 0.0;
  const double cnt = filter ? kept_cnt : (double)B;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    loss[0] = (float)((filter ? kept_sum : all_sum) / cnt);
  if (!grad) return;
  const int b = blockIdx.y;
  const float* sc = scratch + (size_t)b * kItemFloats;
  const bool kept = !filter || sc[0] > -30.f;
  const float wgt = kept ? (float)(1.0 / (cnt * NS)) : 0.f;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= T) return;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const float* q = sc + 1 + i * 5;
    const int j = (int)q[4];
    const size_t oe = ((size_t)b * NS + i) * T + x, ot = ((size_t)b * NS + j) * T + x;
    grad[oe] = wgt * (q[0] * (est[oe] - q[2]) + q[1] * (tgt[ot] - q[3]));
  }
}

size_t pit_scratch_floats(int B) { return (size_t)B * kItemFloats; }

int launch_pit_loss(const float* est, const float* tgt, int B, int n_src, int T, int sdr_type,
                    int threshold, float* loss, float* pw, int32_t* perm, float* grad, void* scratch,
                    cudaStream_t st) {
  TD_REQUIRE(n_src == 2 || n_src == 3, "pit_loss: n_src=%d (2 or 3; the Hungarian path is out of scope)", n_src);
  TD_REQUIRE(sdr_type >= 0 && sdr_type <= 2, "pit_loss: sdr_type=%d", sdr_type);
  TD_REQUIRE(B > 0 && T > 0, "pit_loss: empty batch");
  float* sc = (float*)scratch;
  dim3 g2(grad ? cdiv(T, 256) : 1, grad ? B : 1);
  if (n_src == 2) {
    TD_LAUNCH((pit_item_kernel<2>), B, 512, 0, st, est, tgt, T, sdr_type, pw, perm, sc);
    TD_LAUNCH((pit_finish_kernel<2>), g2, 256, 0, st, est, tgt, B, T, threshold, sc, loss, grad);
  } else {
    TD_LAUNCH((pit_item_kernel<3>), B, 512, 0, st, est, tgt, T, sdr_type, pw, perm, sc);
    TD_LAUNCH((pit_finish_kernel<3>), g2, 256, 0, st, est, tgt, B, T, threshold, sc, loss, grad);
  }
  return 0;
}

}  // namespace td
