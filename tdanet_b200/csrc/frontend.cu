// Waveform front end and back end: pad_input + encoder Conv1d (+GlobLN statistics), bottleneck,
// decoder ConvTranspose1d + crop.
//
// Reference: TDANet_best.py:465-479 (pad_input), :430-438/:497 (encoder), :441-444/:502-503 (ln,
// bottleneck), :453-461/:511-518 (decoder + crop); TDANet_mult_tes.py:317-342 (ConvEncoder).
#include "kernels.h"

namespace td {

// ----------------------------------------------------------------------------- encoder
// out[b, t, n] = sum_j w_k[n', j] * xp[t*S + j - ks_k/2], xp = zero-padded input of pad_input.
// The padding is never materialised: xp[p] = wav[p - front_pad] inside [0, T), else 0.
// Wide windows (the reference's default enc_kernel_size = 21 ms: K = 336, 169 basis signals = 228 KB of weights) do
// not fit one CTA's shared memory: blockIdx.y then walks chunks of `chunk` output channels (single-conv encoders only).
__global__ void __launch_bounds__(256) encoder_kernel(EncArgs a, int rows_per_cta, int ksmax, int chunk) {
  grid_dep_wait();
  extern __shared__ float sm[];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, a.L0 - t0);
  const int n_lo = blockIdx.y * chunk;              // first output channel of this CTA
  const int nc = min(chunk, a.Nb - n_lo);           // == Nb when gridDim.y == 1
  const int span = (rows_per_cta - 1) * a.S + ksmax;
  float* xs = sm;         // [span]
  float* ws = sm + span;  // per conv: [ch_per_conv][ks_k + 1]  (chunked: [nc][ks + 1] of conv 0)
  const int wstart = t0 * a.S - ksmax / 2;
  const float* wav = a.wav + (size_t)b * a.T;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int p = wstart + i;
    const int q = p - a.front_pad;
    xs[i] = (p >= 0 && p < a.Tp && q >= 0 && q < a.T) ? __ldg(wav + q) : 0.f;
  }
  int woff[TDANET_MAX_ENC];
  if (gridDim.y > 1) {
    const int ks = a.ks[0];
    woff[0] = 0;
    const float* w0 = a.w[0] + (size_t)n_lo * ks;
    for (int i = threadIdx.x; i < nc * ks; i += blockDim.x) ws[(i / ks) * (ks + 1) + (i % ks)] = __ldg(w0 + i);
  } else {
    int off = 0;
    for (int k = 0; k < a.nconv; ++k) {
      woff[k] = off;
      const int ks = a.ks[k];
      for (int i = threadIdx.x; i < a.ch_per_conv * ks; i += blockDim.x)
        ws[off + (i / ks) * (ks + 1) + (i % ks)] = __ldg(a.w[k] + i);
      off += a.ch_per_conv * (ks + 1);
    }
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  float* out = a.out + ((size_t)b * a.L0 + t0) * a.Nb;
  for (int idx = threadIdx.x; idx < rows * nc; idx += blockDim.x) {
    const int r = idx / nc, n = idx % nc;            // n: channel inside the chunk
    const int k = gridDim.y > 1 ? 0 : n / a.ch_per_conv, nn = gridDim.y > 1 ? n : n % a.ch_per_conv;
    const int ks = a.ks[k];
    const float* wr = ws + woff[k] + nn * (ks + 1);
    const float* xr = xs + r * a.S + (ksmax - ks) / 2;
    float acc = 0.f;
    for (int j = 0; j < ks; ++j) acc = fmaf(wr[j], xr[j], acc);
    out[(size_t)r * a.Nb + n_lo + n] = acc;
    s1 += acc;
    s2 = fmaf(acc, acc, s2);
  }
  __shared__ double red[64];
  double d1 = s1, d2 = s2;
  block_sum2(d1, d2, red);
  if (threadIdx.x == 0) stat_add2(a.det, a.stats + 2 * b, d1, d2);
}

int launch_encoder(const EncArgs& a, cudaStream_t st) {
  int ksmax = 0, wfloats = 0;
  for (int k = 0; k < a.nconv; ++k) {
    TD_REQUIRE(a.ks[k] % 2 == 0, "encoder: window %d must be even", a.ks[k]);
    ksmax = a.ks[k] > ksmax ? a.ks[k] : ksmax;
    wfloats += a.ch_per_conv * (a.ks[k] + 1);
  }
  const int rows = 32;
  const size_t xfloats = (size_t)(rows - 1) * a.S + ksmax;
  size_t smem = (xfloats + wfloats) * sizeof(float);
  int chunk = a.Nb, ychunks = 1;
  if (smem > 96 * 1024 && a.nconv == 1) {   // chunks of output channels, about 64 KB of weights each
    chunk = (int)((64 * 1024 / sizeof(float)) / (size_t)(ksmax + 1));
    if (chunk < 1) chunk = 1;
    ychunks = cdiv(a.Nb, chunk);
    smem = (xfloats + (size_t)chunk * (ksmax + 1)) * sizeof(float);
  }
  TD_REQUIRE(smem <= 200 * 1024, "encoder: %zu bytes of shared memory needed", smem);
  if (smem > 48 * 1024)
    TD_CUDA(cudaFuncSetAttribute(encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(a.L0, rows), ychunks, a.B);
  TD_LAUNCH(encoder_kernel, grid, 256, smem, st, a, rows, ksmax, chunk);
  return 0;
}

// ----------------------------------------------------------------------------- bottleneck
// x0 = Conv1d(Nb -> c, k=1)(GlobLN(enc)); the normalisation is folded into the weights per item.
__global__ void __launch_bounds__(256) bottleneck_kernel(const float* __restrict__ enc, NormRef norm,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         float* __restrict__ out, int L0, int Nb, int c,
                                                         int rows_per_cta) {
  grid_dep_wait();
  extern __shared__ float sm[];
  float* wf = sm;                   // [c][Nb+1]  folded weights (+1: odd stride, no bank conflicts)
  float* bf = wf + c * (Nb + 1);    // [c]
  float* es = bf + c;               // [rows][Nb]
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, L0 - t0);
  float r, mur;
  norm_moments(norm, b, r, mur);
  for (int i = threadIdx.x; i < c * Nb; i += blockDim.x) {
    const int o = i / Nb, n = i % Nb;
    wf[o * (Nb + 1) + n] = __ldg(w + i) * (__ldg(norm.gamma + n) * r);
  }
  for (int o = threadIdx.x; o < c; o += blockDim.x) {
    float acc = __ldg(bias + o);
    for (int n = 0; n < Nb; ++n) {
      const float g = __ldg(norm.gamma + n);
      acc = fmaf(__ldg(w + o * Nb + n), fmaf(-g, mur, __ldg(norm.beta + n)), acc);
    }
    bf[o] = acc;
  }
  const float* e = enc + ((size_t)b * L0 + t0) * Nb;
  for (int i = threadIdx.x; i < rows * Nb; i += blockDim.x) es[i] = __ldg(e + i);
  __syncthreads();
  float* op = out + ((size_t)b * L0 + t0) * c;
  for (int idx = threadIdx.x; idx < rows * c; idx += blockDim.x) {
    const int r = idx / c, o = idx % c;
    const float* wr = wf + o * (Nb + 1);
    const float* er = es + r * Nb;
    float acc = bf[o];
    for (int n = 0; n < Nb; ++n) acc = fmaf(wr[n], er[n], acc);
    op[idx] = acc;
  }
}

// Register-resident form for the reference encoders (Nb = K/2+1 <= NBR): a thread owns ONE output channel, keeps its
// NBR folded weights in registers and walks the rows of the CTA; an encoder row is read from shared memory as
// NBR/4 broadcast float4.  (The kernel above reads two shared-memory words per FMA and re-folds the weights for
// every 32 rows: 226 us per forward at B = 64 against ~15 us of HBM time.)
template <int NBR>
__global__ void __launch_bounds__(256) bottleneck_reg_kernel(const float* __restrict__ enc, NormRef norm,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias,
                                                             float* __restrict__ out, int L0, int Nb, int c,
                                                             int rows_per_cta) {
  grid_dep_wait();
  extern __shared__ float sm[];  // [rows][NBR] encoder rows, zero-padded to NBR
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, L0 - t0);
  float r, mur;
  norm_moments(norm, b, r, mur);
  const float* e = enc + ((size_t)b * L0 + t0) * Nb;
  for (int i = threadIdx.x; i < rows * NBR; i += blockDim.x) {
    const int rr = i / NBR, n = i % NBR;
    sm[i] = n < Nb ? __ldg(e + (size_t)rr * Nb + n) : 0.f;
  }
  // output channel o = threadIdx.x % c, row phase = threadIdx.x / c (blockDim.x is a multiple of c)
  const int o = threadIdx.x % c, ph = threadIdx.x / c, nph = blockDim.x / c;
  float wr[NBR];
  float bo = __ldg(bias + o);
#pragma unroll
  for (int n = 0; n < NBR; ++n) {
    wr[n] = 0.f;
    if (n < Nb) {
      const float g = __ldg(norm.gamma + n), wv = __ldg(w + (size_t)o * Nb + n);
      wr[n] = wv * (g * r);
      bo = fmaf(wv, fmaf(-g, mur, __ldg(norm.beta + n)), bo);
    }
  }
  __syncthreads();
  float* op = out + ((size_t)b * L0 + t0) * c + o;
  for (int rr = ph; rr < rows; rr += nph) {
    const float4* er = reinterpret_cast<const float4*>(sm + rr * NBR);
    float acc = bo;
#pragma unroll
    for (int q = 0; q < NBR / 4; ++q) {
      const float4 v = er[q];
      acc = fmaf(wr[4 * q], v.x, acc); acc = fmaf(wr[4 * q + 1], v.y, acc);
      acc = fmaf(wr[4 * q + 2], v.z, acc); acc = fmaf(wr[4 * q + 3], v.w, acc);
    }
    op[(size_t)rr * c] = acc;
  }
}

int launch_bottleneck(const float* enc, const NormRef& norm, const float* w, const float* bias,
                      float* out, int B, int L0, int Nb, int c, cudaStream_t st) {
  if (Nb <= 68 && c <= 256 && 256 % c == 0) {
    const int rows = 64;
    dim3 grid(cdiv(L0, rows), 1, B);
    if (Nb <= 36) {
      TD_LAUNCH((bottleneck_reg_kernel<36>), grid, 256, (size_t)rows * 36 * sizeof(float), st, enc, norm, w, bias, out, L0, Nb, c, rows);
    } else {
      TD_LAUNCH((bottleneck_reg_kernel<68>), grid, 256, (size_t)rows * 68 * sizeof(float), st, enc, norm, w, bias, out, L0, Nb, c, rows);
    }
    return 0;
  }
  const int rows = 32;
  const size_t smem = ((size_t)c * (Nb + 1) + c + (size_t)rows * Nb) * sizeof(float);
  TD_REQUIRE(smem <= 200 * 1024, "bottleneck: %zu bytes of shared memory needed", smem);
  if (smem > 48 * 1024)
    TD_CUDA(cudaFuncSetAttribute(bottleneck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(L0, rows), 1, B);
  TD_LAUNCH(bottleneck_kernel, grid, 256, smem, st, enc, norm, w, bias, out, L0, Nb, c, rows);
  return 0;
}

// ----------------------------------------------------------------------------- mask_net on the tensor cores
// masked[r, n] = relu(sum_k prelu(y[r, k]) * W[n, k] + bias[n]) * enc[r, n mod Nb]     (TDANet_best.py:505-509)
// y [R, K] (K = out_channels), W [N, K] with N = n_src * Nb (66 for the 4 ms encoder: not a tile multiple, which
// left the CUDA-core GEMM at 9 TFLOP/s, 232 us per forward at B = 64 against ~20 us of HBM time).  TF32 modes only:
// warp-level mma.sync.m16n8k8, one CTA = four warps x 16 rows per tile of 64 rows, W staged once per CTA
// (TF32-rounded, rows padded to K + 4 floats), the A tile staged through shared memory with the PReLU applied,
// NT = ceil(N / 8) accumulator tiles per warp.
__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>
__global__ void __launch_bounds__(128) mask_conv_mma_kernel(const float* __restrict__ y, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ slope_p,
                                                            const float* __restrict__ enc, float* __restrict__ out,
                                                            int R, int K, int N, int Nb, int n_tiles) {
  grid_dep_wait();
  extern __shared__ float sm[];
  const int LD = K + 4;
  float* Ws = sm;                 // [NT*8][LD]
  float* As = sm + NT * 8 * LD;   // [64][LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K4 = K / 4;
  for (int i = tid; i < NT * 8 * K4; i += 128) {
    const int n = i / K4, k = (i % K4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)n * K + k));
    *reinterpret_cast<float4*>(Ws + n * LD + k) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
  }
  const float slope = __ldg(slope_p);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int r0 = tile * 64;
    __syncthreads();  // the previous tile's fragments have been read (first pass: orders the W stores as well)
    for (int i = tid; i < 64 * K4; i += 128) {
      const int rr = i / K4, k = (i % K4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + rr < R) v = *reinterpret_cast<const float4*>(y + (size_t)(r0 + rr) * K + k);
      v.x = tf32_rna(preluf_(v.x, slope)); v.y = tf32_rna(preluf_(v.y, slope));
      v.z = tf32_rna(preluf_(v.z, slope)); v.w = tf32_rna(preluf_(v.w, slope));
      *reinterpret_cast<float4*>(As + rr * LD + k) = v;
    }
    __syncthreads();
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    const uint32_t* ar = reinterpret_cast<const uint32_t*>(As) + (warp * 16 + g) * LD + t;
    const uint32_t* wr = reinterpret_cast<const uint32_t*>(Ws) + g * LD + t;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t a[4] = {ar[k0], ar[8 * LD + k0], ar[k0 + 4], ar[8 * LD + k0 + 4]};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_tf32_m16n8k8(acc[nt], a, wr[nt * 8 * LD + k0], wr[nt * 8 * LD + k0 + 4]);
    }
    // epilogue: rows r0 + 16*warp + g (+8), columns 8*nt + 2t, +1
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + warp * 16 + g + 8 * h;
      if (r < R) {
        const float* er = enc + (size_t)r * Nb;
        float* orow = out + (size_t)r * N;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = nt * 8 + 2 * t;
          if (c < N) {
            const float v0 = fmaxf(acc[nt][2 * h] + __ldg(bias + c), 0.f) * er[c % Nb];
            if (c + 1 < N) {
              const float v1 = fmaxf(acc[nt][2 * h + 1] + __ldg(bias + c + 1), 0.f) * er[(c + 1) % Nb];
              *reinterpret_cast<float2*>(orow + c) = make_float2(v0, v1);
            } else {
              orow[c] = v0;
            }
          }
        }
      }
    }
  }
}

template <int NT>
static int launch_mask_conv_mma_t(const float* y, const float* w, const float* bias, const float* slope, const float* enc,
                                  float* out, int R, int K, int N, int Nb, cudaStream_t st) {
  const size_t smem = (size_t)(NT * 8 + 64) * (K + 4) * sizeof(float);
  static bool attr_set[16] = {};
  int dev = 0;
  TD_CUDA(cudaGetDevice(&dev));
  if (smem > 48 * 1024 && (dev < 0 || dev >= 16 || !attr_set[dev])) {
    TD_CUDA(cudaFuncSetAttribute(mask_conv_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 16) attr_set[dev] = true;
  }
  const int tiles = cdiv(R, 64);
  const int grid = tiles < 148 * 3 ? tiles : 148 * 3;
  TD_LAUNCH((mask_conv_mma_kernel<NT>), grid, 128, smem, st, y, w, bias, slope, enc, out, R, K, N, Nb, tiles);
  return 0;
}

// returns TDANET_EUNSUPPORTED (without setting an error text the caller would report) when the shape is outside
// what the kernel covers; the caller then takes the CUDA-core GEMM
bool mask_conv_mma_applies(int K, int N) { return K % 8 == 0 && K <= 256 && N % 2 == 0 && N <= 136; }

int launch_mask_conv_mma(const float* y, const float* w, const float* bias, const float* slope, const float* enc,
                         float* out, int R, int K, int N, int Nb, cudaStream_t st) {
  TD_REQUIRE(mask_conv_mma_applies(K, N), "mask_conv_mma: K=%d N=%d", K, N);
  const int nt = cdiv(N, 8);
  if (nt <= 5) return launch_mask_conv_mma_t<5>(y, w, bias, slope, enc, out, R, K, N, Nb, st);
  if (nt <= 9) return launch_mask_conv_mma_t<9>(y, w, bias, slope, enc, out, R, K, N, Nb, st);
  if (nt <= 13) return launch_mask_conv_mma_t<13>(y, w, bias, slope, enc, out, R, K, N, Nb, st);
  return launch_mask_conv_mma_t<17>(y, w, bias, slope, enc, out, R, K, N, Nb, st);
}

// ----------------------------------------------------------------------------- decoder
// ConvTranspose1d(n_src*Nb -> n_src, k=K, stride=S=K/4, padding=K/2, no bias) followed by the crop
// [K-S : -(rest+K-S)].  With n the index in the un-cropped output and u = (n + K/2)/S,
// r = (n + K/2) % S:    out[o, n] = sum_{m<4} sum_ci M[u-m, ci] * W[ci, o, r + m*S].
// A thread owns one phase r and four consecutive hops for every source, so each weight load feeds
// four outputs and each frame value feeds up to four taps.
template <int NS>
__global__ void __launch_bounds__(256) decoder_kernel(const float* __restrict__ masked,
                                                      const float* __restrict__ w,
                                                      float* __restrict__ est, int L0, int CI, int K,
                                                      int S, int T, int hops_per_cta, int RS) {
  // RS = phases covered by one pass of the CTA (== S when S <= 64; hops wider than that - the reference's default
  // enc_kernel_size = 21 ms gives S = 84 - are walked in several passes of RS phases)
  grid_dep_wait();
  extern __shared__ float ms[];  // [hops_per_cta + 3][CI]
  const int b = blockIdx.z;
  const int h0 = blockIdx.x * hops_per_cta;  // first hop (n / S) of this CTA
  const int f0 = h0 + 2 - 3;                 // first frame needed: u - 3 with u = h0 + 2
  const int nfr = hops_per_cta + 3;
  for (int i = threadIdx.x; i < nfr * CI; i += blockDim.x) {
    const int f = f0 + i / CI;
    ms[i] = (f >= 0 && f < L0) ? __ldg(masked + ((size_t)b * L0 + f) * CI + i % CI) : 0.f;
  }
  __syncthreads();
  const int qg = threadIdx.x / RS;
  if (qg * 4 >= hops_per_cta) return;
  for (int r = threadIdx.x % RS; r < S; r += RS) {
  float acc[NS][4];
#pragma unroll
  for (int o = 0; o < NS; ++o)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[o][q] = 0.f;
  // frames used by this thread: local index qg*4 + {0..6}  (u_q - m = h0 + qg*4 + q + 2 - m)
  const float* mrow = ms + (size_t)(qg * 4) * CI;
  for (int ci = 0; ci < CI; ++ci) {
    float fr[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) fr[i] = mrow[i * CI + ci];
#pragma unroll
    for (int o = 0; o < NS; ++o) {
      const float* wp = w + ((size_t)ci * NS + o) * K + r;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float wv = __ldg(wp + m * S);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[o][q] = fmaf(fr[q + 3 - m], wv, acc[o][q]);
      }
    }
  }
  const int crop = K - S;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = (h0 + qg * 4 + q) * S + r - crop;
    if (n >= 0 && n < T) {
#pragma unroll
      for (int o = 0; o < NS; ++o) est[((size_t)b * NS + o) * T + n] = acc[o][q];
    }
  }
  }
}

int launch_decoder(const float* masked, const float* w, float* est, int B, int L0, int Nb, int n_src,
                   int K, int S, int T, cudaStream_t st) {
  TD_REQUIRE(K == 4 * S, "decoder: window %d must be 4 hops of %d", K, S);
  const int CI = n_src * Nb;
  const int RS = S <= 64 ? S : 64;
  int hops = (256 / RS) * 4;  // every thread of a 256-thread CTA owns 4 hops of one phase (of several when S > 64)
  if (hops > 64) hops = 64;
  const int threads = RS * (hops / 4);
  const size_t smem = (size_t)(hops + 3) * CI * sizeof(float);
  const int total_hops = cdiv(K - S + T, S);
  dim3 grid(cdiv(total_hops, hops), 1, B);
  if (n_src == 2) {
    if (smem > 48 * 1024)
      TD_CUDA(cudaFuncSetAttribute(decoder_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TD_LAUNCH((decoder_kernel<2>), grid, threads, smem, st, masked, w, est, L0, CI, K, S, T, hops, RS);
  } else if (n_src == 3) {
    if (smem > 48 * 1024)
      TD_CUDA(cudaFuncSetAttribute(decoder_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TD_LAUNCH((decoder_kernel<3>), grid, threads, smem, st, masked, w, est, L0, CI, K, S, T, hops, RS);
  } else {
    return fail(TDANET_EUNSUPPORTED, "decoder: num_sources=%d (2 or 3 supported)", n_src);
  }
  return 0;
}

// ----------------------------------------------------------------------------- training-forward glue
__global__ void concat_kernel(const float* __restrict__ y, const float* __restrict__ mix,
                              const float* __restrict__ cw, const float* __restrict__ cb,
                              const float* __restrict__ slope, float* __restrict__ out, size_t n, int c) {
  grid_dep_wait();
  const float sl = __ldg(slope);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    out[i] = preluf_(fmaf(__ldg(cw + ch), mix[i] + y[i], __ldg(cb + ch)), sl);
  }
}

int launch_concat(const float* y, const float* mix, const float* cw, const float* cb, const float* slope,
                  float* out, int rows, int c, cudaStream_t st) {
  const size_t n = (size_t)rows * c;
  TD_LAUNCH(concat_kernel, (unsigned)((n + 255) / 256), 256, 0, st, y, mix, cw, cb, slope, out, n, c);
  return 0;
}

__global__ void mask_apply_kernel(const float* __restrict__ m, const float* __restrict__ enc,
                                  float* __restrict__ masked, size_t n, int CI, int Nb) {
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / CI;
    const int col = (int)(i % CI);
    masked[i] = fmaxf(m[i], 0.f) * enc[r * Nb + col % Nb];
  }
}

int launch_mask_apply(const float* m, const float* enc, float* masked, int rows, int n_src, int Nb, cudaStream_t st) {
  const int CI = n_src * Nb;
  const size_t n = (size_t)rows * CI;
  TD_LAUNCH(mask_apply_kernel, (unsigned)((n + 255) / 256), 256, 0, st, m, enc, masked, n, CI, Nb);
  return 0;
}

}  // namespace td
