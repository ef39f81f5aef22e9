// Long-form (continuous speech separation) stitching on device.
//
// Reference: audio_test_css.py:108-134.  Every chunk of a recording is separated on its own; the
// permutation of chunk k > 0 is chosen by comparing the cosine similarity of its first `overlap` samples
// with the LAST `overlap` samples of chunk 0 (the reference never updates that tail), then
// est[k][:, overlap:] is appended.  Two kernels: scores (one CTA per stream x chunk) and the copy.
#include "kernels.h"

namespace td {

// swap[s, k] = 1 if chunk k of stream s must be appended with its two sources exchanged
__global__ void __launch_bounds__(256) css_score_kernel(const float* __restrict__ est, int n_chunks, int seg_len,
                                                        int overlap, int32_t* __restrict__ swap) {
  grid_dep_wait();
  __shared__ double sh[64];
  const int s = blockIdx.y, k = blockIdx.x + 1;
  const float* first = est + ((size_t)s * n_chunks) * 2 * seg_len;
  const float* p1 = first + (seg_len - overlap);            // tail of chunk 0, source 1
  const float* p2 = first + seg_len + (seg_len - overlap);  // tail of chunk 0, source 2
  const float* e1 = est + ((size_t)s * n_chunks + k) * 2 * seg_len;
  const float* e2 = e1 + seg_len;
  double d11 = 0, d12 = 0, d21 = 0, d22 = 0, np1 = 0, np2 = 0, ne1 = 0, ne2 = 0;
  for (int i = threadIdx.x; i < overlap; i += blockDim.x) {
    const float a = p1[i], b = p2[i], x = e1[i], y = e2[i];
    d11 += (double)a * x; d12 += (double)a * y; d21 += (double)b * x; d22 += (double)b * y;
    np1 += (double)a * a; np2 += (double)b * b; ne1 += (double)x * x; ne2 += (double)y * y;
  }
  block_sum2(d11, d12, sh);
  block_sum2(d21, d22, sh);
  block_sum2(np1, np2, sh);
  block_sum2(ne1, ne2, sh);
  if (threadIdx.x == 0) {
    // F.cosine_similarity(x, y, dim=0, eps=1e-8) = x.y / sqrt(max(|x|^2 |y|^2, eps^2))
    const double eps2 = 1e-16;
    auto cs = [&](double d, double n1, double n2) { return d / sqrt(fmax(n1 * n2, eps2)); };
    const float comb1 = (float)cs(d11, np1, ne1) + (float)cs(d22, np2, ne2);
    const float comb2 = (float)cs(d12, np1, ne2) + (float)cs(d21, np2, ne1);
    swap[s * n_chunks + k] = comb1 > comb2 ? 0 : 1;
    if (blockIdx.x == 0) swap[s * n_chunks] = 0;
  }
}

// out[s, c, :] = est[s,0,c,:] ++ est[s,k,c^swap,overlap:] for k = 1..n_chunks-1, cut to out_len
__global__ void __launch_bounds__(256) css_copy_kernel(const float* __restrict__ est, const int32_t* __restrict__ swap,
                                                       int n_chunks, int seg_len, int overlap, int out_len,
                                                       float* __restrict__ out) {
  grid_dep_wait();
  const int s = blockIdx.z, c = blockIdx.y;
  const int hop = seg_len - overlap;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < out_len; n += gridDim.x * blockDim.x) {
    int k, i;
    if (n < seg_len) {
      k = 0; i = n;
    } else {
      k = 1 + (n - seg_len) / hop;
      i = overlap + (n - seg_len) % hop;
    }
    const int src = c ^ swap[s * n_chunks + k];
    out[((size_t)s * 2 + c) * out_len + n] = est[(((size_t)s * n_chunks + k) * 2 + src) * seg_len + i];
  }
}

int launch_css_stitch(const float* est, int n_streams, int n_chunks, int seg_len, int overlap, int out_len,
                      int32_t* swap, float* out, cudaStream_t st) {
  TD_REQUIRE(n_streams > 0 && n_chunks > 0 && seg_len > 0, "css_stitch: empty input");
  TD_REQUIRE(overlap >= 0 && overlap < seg_len, "css_stitch: overlap %d outside [0, %d)", overlap, seg_len);
  TD_REQUIRE(out_len >= 0 && out_len <= seg_len + (n_chunks - 1) * (seg_len - overlap), "css_stitch: out_len %d too long", out_len);
  if (n_chunks > 1) {
    dim3 g1(n_chunks - 1, n_streams);
    TD_LAUNCH(css_score_kernel, g1, 256, 0, st, est, n_chunks, seg_len, overlap, swap);
  } else {
    TD_CUDA(cudaMemsetAsync(swap, 0, sizeof(int32_t) * n_streams, st));
  }
  if (out_len > 0) {
    dim3 g2(cdiv(out_len, 256 * 4), 2, n_streams);
    TD_LAUNCH(css_copy_kernel, g2, 256, 0, st, est, swap, n_chunks, seg_len, overlap, out_len, out);
  }
  return 0;
}

}  // namespace td
