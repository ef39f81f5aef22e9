// Keep-masks of the training-mode stochastic layers: nn.Dropout(0.1) after the FFN's ReLU and
// after fc2's GlobLN (TDANet_best.py:210,212), MultiHeadAttention.dropout (:251), the dropout on the attention
// weights inside nn.MultiheadAttention (:241) and DropPath on both GA branches (:7-30,:262-263).
//
// All masks of all UConvBlock iterations of one forward are drawn by ONE launch, before the first block runs, with
// the counter-based Philox4x32-10 generator: key = seed, counter = {group of 4 elements, site | iteration << 8,
// offset}.  They are stored as bytes next to the activations of their iteration, so the backward pass reads what
// the forward used and a test can hand the very same masks to the oracle.  `offset` lives on the device and is
// advanced by a second, single-thread launch: a CUDA-graph replay of the step draws fresh masks.
#include "kernels.h"

namespace td {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr int MAX_REGIONS = 8;
struct MaskJob {
  MaskRegion r[MAX_REGIONS];
  size_t first_group[MAX_REGIONS + 1];  // prefix sums of ceil(n / 16) over the regions
  int n;
};

// one thread per 16 mask bytes (4 Philox calls); blockIdx.y = UConvBlock iteration
__global__ void __launch_bounds__(256) dropout_masks_kernel(char* __restrict__ base, size_t blk_stride, MaskJob job,
                                                            const uint64_t* __restrict__ rng_state) {
  grid_dep_wait();
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= job.first_group[job.n]) return;
  int ri = 0;
  while (g >= job.first_group[ri + 1]) ++ri;
  const MaskRegion reg = job.r[ri];
  const size_t i0 = (g - job.first_group[ri]) * 16;
  const uint64_t seed = rng_state[0], offset = rng_state[1];
  const uint32_t site = reg.site | ((uint32_t)blockIdx.y << 8);
  uint32_t packed[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r4[4];
    const uint64_t ctr = (i0 >> 2) + q;  // group of 4 elements
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32) ^ (uint32_t)(offset >> 32), site, (uint32_t)offset,
                  (uint32_t)seed, (uint32_t)(seed >> 32), r4);
    packed[q] = (r4[0] >= reg.thresh ? 1u : 0u) | (r4[1] >= reg.thresh ? 1u << 8 : 0u) |
                (r4[2] >= reg.thresh ? 1u << 16 : 0u) | (r4[3] >= reg.thresh ? 1u << 24 : 0u);
  }
  char* dst = base + (size_t)blockIdx.y * blk_stride + reg.off + i0;
  if (i0 + 16 <= reg.n) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  } else {
    for (size_t i = i0; i < reg.n; ++i) dst[i - i0] = (char)((packed[(i - i0) >> 2] >> (8 * ((i - i0) & 3))) & 1u);
  }
}

__global__ void rng_advance_kernel(uint64_t* rng_state) {
  grid_dep_wait();
  rng_state[1] += 1;
}

int launch_dropout_masks(char* base, size_t blk_stride, int n_blk, const MaskRegion* regions, int n_regions,
                         uint64_t* rng_state, cudaStream_t st) {
  TD_REQUIRE(n_regions > 0 && n_regions <= MAX_REGIONS, "dropout: %d mask regions", n_regions);
  TD_REQUIRE(rng_state != nullptr, "dropout / drop_path > 0 needs rng_state (tdanet_forward_train_rng)");
  MaskJob job{};
  job.n = n_regions;
  job.first_group[0] = 0;
  for (int i = 0; i < n_regions; ++i) {
    TD_REQUIRE(regions[i].off % 16 == 0, "dropout: mask region %d is not 16-byte aligned", i);
    job.r[i] = regions[i];
    job.first_group[i + 1] = job.first_group[i] + (regions[i].n + 15) / 16;
  }
  const size_t groups = job.first_group[n_regions];
  if (groups > 0) {
    dim3 grid((unsigned)((groups + 255) / 256), n_blk);
    TD_LAUNCH(dropout_masks_kernel, grid, 256, 0, st, base, blk_stride, job, rng_state);
  }
  TD_LAUNCH(rng_advance_kernel, 1, 1, 0, st, rng_state);
  return 0;
}

}  // namespace td
