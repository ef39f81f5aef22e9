// Weight gradients of the 1x1 convolutions / Linear layers on the tensor cores:
//   dW[n, k] += sum_r G[r, n] * A[r, k],   db[n] += sum_r G[r, n]
// G = gradient w.r.t. the layer output [R, N], A = layer input [R, K], both channels-last, so the reduction
// runs over the SLOW axis of both operands.  Warp-level mma.sync.m16n8k8 (TF32 operands rounded to nearest
// when they are staged in shared memory, fp32 accumulate): the fragments are gathered from shared memory by the
// lanes, which is what makes the transposed ("MN-major") operands free; the rows are split over blockIdx.z and
// the partial tiles meet in dW through vector reductions.  The exact fp32 path (gemm_mode fp32) and the odd
// shapes (N = 66 mask conv, K = 33 bottleneck) stay on wgrad_kernel (bwd_kernels.cuh), which is also the
// reference this kernel is tested against on the GPU (tests/test_gpu_train.py::test_wgrad_*).
#include "kernels.h"

namespace td {

constexpr int WM_T = 64, WM_R = 32, WM_S = WM_T + 8;  // tile edge, rows per stage, padded smem row (conflict-free fragments)

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) wgrad_mma_kernel(const float* __restrict__ G, const float* __restrict__ A,
                                                        float* __restrict__ dW, float* __restrict__ db, int R, int N,
                                                        int K, int rows_per_split) {
  grid_dep_wait();
  __shared__ __align__(16) float Gs[2][WM_R][WM_S];
  __shared__ __align__(16) float As[2][WM_R][WM_S];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * WM_T, k0 = blockIdx.y * WM_T;
  const int r_begin = blockIdx.z * rows_per_split, r_end = min(r_begin + rows_per_split, R);
  // loader: 32 rows x 16 float4 columns per operand and stage; a thread always loads the same four columns
  const int lc = (tid & 15) * 4, lr = tid >> 4;
  const bool g_ok = n0 + lc < N, a_ok = k0 + lc < K;  // N, K are multiples of 4
  float4 rg[2], ra[2];
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int r0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + lr + 16 * i;
      const bool ok = r < r_end;
      const int rc = ok ? r : r_end - 1;
      rg[i] = g_ok ? __ldg(reinterpret_cast<const float4*>(G + (size_t)rc * N + n0 + lc)) : make_float4(0.f, 0.f, 0.f, 0.f);
      ra[i] = a_ok ? __ldg(reinterpret_cast<const float4*>(A + (size_t)rc * K + k0 + lc)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (!ok) { rg[i] = make_float4(0.f, 0.f, 0.f, 0.f); ra[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      bsum.x += rg[i].x; bsum.y += rg[i].y; bsum.z += rg[i].z; bsum.w += rg[i].w;
      *reinterpret_cast<float4*>(&Gs[buf][lr + 16 * i][lc]) =
          make_float4(tf32_rna(rg[i].x), tf32_rna(rg[i].y), tf32_rna(rg[i].z), tf32_rna(rg[i].w));
      *reinterpret_cast<float4*>(&As[buf][lr + 16 * i][lc]) =
          make_float4(tf32_rna(ra[i].x), tf32_rna(ra[i].y), tf32_rna(ra[i].z), tf32_rna(ra[i].w));
    }
  };
  // warp tile: 16 (n) x 32 (k); 4 x 2 warps
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int g = lane >> 2, t = lane & 3;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

  gload(r_begin);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += WM_R) {
    const bool more = r0 + WM_R < r_end;
    if (more) gload(r0 + WM_R);
#pragma unroll
    for (int kk = 0; kk < WM_R; kk += 8) {
      uint32_t a[4];
      a[0] = __float_as_uint(Gs[buf][kk + t][wm + g]);
      a[1] = __float_as_uint(Gs[buf][kk + t][wm + g + 8]);
      a[2] = __float_as_uint(Gs[buf][kk + t + 4][wm + g]);
      a[3] = __float_as_uint(Gs[buf][kk + t + 4][wm + g + 8]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t b[2];
        b[0] = __float_as_uint(As[buf][kk + t][wn + j * 8 + g]);
        b[1] = __float_as_uint(As[buf][kk + t + 4][wn + j * 8 + g]);
        mma_tf32(acc[j], a, b);
      }
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  // C fragment: rows g / g+8 of the warp tile, columns 2t, 2t+1 of each n8 block
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = n0 + wm + g + 8 * h;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + wn + j * 8 + 2 * t;
      if (k < K) {  // K is even
        vf<2> v;
        v[0] = acc[j][2 * h];
        v[1] = acc[j][2 * h + 1];
        vred_add<2>(dW + (size_t)n * K + k, v);
      }
    }
  }
  if (db && blockIdx.y == 0 && g_ok) {
    vf<4> v;
    v[0] = bsum.x; v[1] = bsum.y; v[2] = bsum.z; v[3] = bsum.w;
    vred_add<4>(db + n0 + lc, v);
  }
}

int launch_wgrad_mma(const float* G, const float* A, float* dW, float* db, int R, int N, int K, cudaStream_t st) {
  TD_REQUIRE(N % 4 == 0 && K % 4 == 0, "wgrad_mma: N=%d K=%d must be multiples of 4", N, K);
  const int tiles = cdiv(N, WM_T) * cdiv(K, WM_T);
  int splits = cdiv(4 * 148, tiles);  // ~4 CTAs per SM: the row loop of a CTA is a serial chain of global-load latencies
  int rps = cdiv(cdiv(R, splits), WM_R) * WM_R;
  splits = cdiv(R, rps);
  dim3 grid(cdiv(N, WM_T), cdiv(K, WM_T), splits);
  TD_LAUNCH(wgrad_mma_kernel, grid, 256, 0, st, G, A, dW, db, R, N, K, rps);
  return 0;
}

}  // namespace td
