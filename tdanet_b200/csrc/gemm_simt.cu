// fp32 CUDA-core GEMM for the 1x1 convolutions / linear layers in channels-last form:
//   D[b, r, n] = sum_k A[b, r, k] * W[n, k]  (+ epilogue)
// This is the exact-parity path (TDANET_GEMM_FP32) and the numerical reference the tcgen05 path
// is tested against on the GPU; it shares its epilogues with gemm_tc.cu through gemm_epilogue.cuh.
#include "kernels.h"
#include "gemm_epilogue.cuh"

namespace td {

constexpr int BM = 128, BN = 128, BK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a, int tiles_per_item) {
  grid_dep_wait();
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Ws[2][BK][BN];
  __shared__ double red[64];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / tiles_per_item;
  const int r0 = (blockIdx.x % tiles_per_item) * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // loader mapping: 128 rows x 16 k  ->  thread loads 8 consecutive k of one row
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  const bool a_ok = r0 + lrow < a.L;
  const bool w_ok = n0 + lrow < a.N;
  const float* ap = a.A + ((size_t)b * a.L + (a_ok ? r0 + lrow : 0)) * a.K + lk;
  const float* wp = a.W + (size_t)(w_ok ? n0 + lrow : 0) * a.K + lk;
  const float slope = a.a_slope ? __ldg(a.a_slope) : 1.f;

  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ra[i] = a_ok ? __ldg(reinterpret_cast<const float4*>(ap + k0) + i) : make_float4(0, 0, 0, 0);
      rw[i] = w_ok ? __ldg(reinterpret_cast<const float4*>(wp + k0) + i) : make_float4(0, 0, 0, 0);
    }
    if (a.a_slope) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ra[i].x = preluf_(ra[i].x, slope); ra[i].y = preluf_(ra[i].y, slope);
        ra[i].z = preluf_(ra[i].z, slope); ra[i].w = preluf_(ra[i].w, slope);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][lk + 4 * i + 0][lrow] = ra[i].x; As[buf][lk + 4 * i + 1][lrow] = ra[i].y;
      As[buf][lk + 4 * i + 2][lrow] = ra[i].z; As[buf][lk + 4 * i + 3][lrow] = ra[i].w;
      Ws[buf][lk + 4 * i + 0][lrow] = rw[i].x; Ws[buf][lk + 4 * i + 1][lrow] = rw[i].y;
      Ws[buf][lk + 4 * i + 2][lrow] = rw[i].z; Ws[buf][lk + 4 * i + 3][lrow] = rw[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = a.K / BK;
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: rows ty*4+i (+64), cols tx*4+j (+64)
  Epilogue ep(a, b);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int jh = 0; jh < 2; ++jh) {
    const int n = n0 + jh * 64 + tx * 4;
    if (n >= a.N) continue;
    ep.cols(n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
      if (r >= a.L) continue;
      float v[4] = {acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      ep.row(r, v, s1, s2);
    }
  }
  if (a.stats) {
    double d1 = s1, d2 = s2;
    block_sum2(d1, d2, red);
    if (tid == 0) stat_add2(a.det, a.stats + 2 * b, d1, d2);
  }
}

int launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.K % BK == 0, "gemm: K=%d must be a multiple of %d", a.K, BK);
  TD_REQUIRE(a.B > 0 && a.L > 0 && a.N > 0, "gemm: empty problem");
  const int tiles = cdiv(a.L, BM);
  dim3 grid(a.B * tiles, cdiv(a.N, BN));
  TD_LAUNCH(gemm_simt_kernel, grid, 256, 0, st, a, tiles);
  return 0;
}

}  // namespace td
