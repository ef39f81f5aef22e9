// Channels-last depthwise convolutions with normalise-on-load sources, and the LA combine.
//
// Replaces, fused: DilatedConvNorm.conv (TDANet_best.py:179-192) + the GlobLN/PReLU of its
// *producer* applied on load, the three ConvNorm.conv of LA (TDANet_best.py:272-289), the
// loc_glo_fus / nearest-interpolate injection (TDANet_best.py:369-371, TDANet.py:624-626)
// recomputed on load, and LA's gate (TDANet_best.py:291).
//
// Memory-bound kernels: one thread owns V consecutive channels and walks over time with a
// register window; all global loads of a chunk of rows are issued before any is consumed.
// CTAs whose rows (halo included) lie inside the tensor run a path without bounds checks; the
// nearest-neighbour row indices of a CTA are tabulated once in shared memory.
// This file is compiled twice (dwconv_f32.cu / dwconv_bf16.cu): ACT_T is the storage type of the large
// activations (proj_1x1 output, spp_dw outputs, expanded, materialised x_fused) - fp32, or bf16 with fp32
// arithmetic and fp32 statistics.  Bottom-scale tensors, pooled outputs and the global feature stay fp32.
#include "kernels.h"

namespace td {
namespace TD_ACT_NS {

// ----------------------------------------------------------------------------- sources
// EDGE = true: rows outside [0, L) read as exact zeros (conv zero padding).
template <int KIND, int V, bool EDGE>
struct Src {
  const ACT_T* x;  // item base + first channel of this thread
  int L, C;
  vf<V> c0_, c1_, c2_, c3_, c4_, c5_;  // coefficient planes
  float slope;
  const float* g;
  const int* jtab;  // nearest global row of local row t at jtab[t - tab0]   (inject kinds)
  int tab0, cur;
  vf<V> sg, eg;

  static constexpr bool kInject = KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD;

  __device__ __forceinline__ void init(const SrcDesc& d, int b, int ch, int C_, const int* jt, int tab0_) {
    C = C_;
    L = d.L;
    x = reinterpret_cast<const ACT_T*>(d.x) + (size_t)b * d.L * C_ + ch;
    if constexpr (KIND == SRC_INJECT_GATE) {
      const float* cf = d.coef + (size_t)b * 6 * C_ + ch;
      c0_ = vload<V>(cf);
      c1_ = vload<V>(cf + C_);
      c2_ = vload<V>(cf + 2 * C_);
      c3_ = vload<V>(cf + 3 * C_);
      c4_ = vload<V>(cf + 4 * C_);
      c5_ = vload<V>(cf + 5 * C_);
    } else if constexpr (KIND != SRC_PLAIN) {
      norm_coef<V>(d.norm, b, ch, c0_, c1_);
    }
    if constexpr (KIND == SRC_AFFINE_PRELU) slope = __ldg(d.slope);
    if constexpr (kInject) {
      g = d.g + (size_t)b * d.Lg * C_ + ch;
      jtab = jt;
      tab0 = tab0_;
      cur = -1;
    }
  }

  __device__ __forceinline__ vf<V> load_raw(int t) const {
    if (EDGE && (t < 0 || t >= L)) return vzero<V>();
    return aload<V>(x + t * C);
  }

  __device__ __forceinline__ void seek(int j) {
    const vf<V> gr = vload<V>(g + j * C);  // L2-resident: the global feature is small
    cur = j;
    if constexpr (KIND == SRC_INJECT_GATE) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        sg[e] = sigmoidf_(fmaf(c2_[e], gr[e], c3_[e]));
        eg[e] = fmaf(c4_[e], gr[e], c5_[e]);
      }
    } else {
      eg = gr;
    }
  }

  __device__ __forceinline__ vf<V> finalize(vf<V> r, int t) {
    if (EDGE && (t < 0 || t >= L)) return vzero<V>();
    if constexpr (KIND == SRC_PLAIN) {
      return r;
    } else if constexpr (KIND == SRC_AFFINE) {
      return vfma<V>(r, c0_, c1_);
    } else if constexpr (KIND == SRC_AFFINE_PRELU) {
#pragma unroll
      for (int e = 0; e < V; ++e) r[e] = preluf_(fmaf(r[e], c0_[e], c1_[e]), slope);
      return r;
    } else {
      const int j = jtab[t - tab0];
      if (j != cur) seek(j);
      if constexpr (KIND == SRC_INJECT_GATE) {
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = fmaf(fmaf(r[e], c0_[e], c1_[e]), sg[e], eg[e]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = fmaf(r[e], c0_[e], c1_[e]) + eg[e];
      }
      return r;
    }
  }
};

// tab[i] = nearest source row (in a tensor of `in_len` rows) of row clamp(t_first + i) of a tensor
// that is `scale` = fl32(in_len / out_len) times shorter/longer; cooperative, caller synchronises.
__device__ __forceinline__ void fill_nearest(int* tab, int n, int t_first, int out_len, float scale, int in_len) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int t = t_first + i;
    t = t < 0 ? 0 : (t >= out_len ? out_len - 1 : t);
    tab[i] = nearest_src(t, scale, in_len);
  }
}

template <int V>
__device__ __forceinline__ void load_taps(const float* __restrict__ w, int ch, vf<V> (&tap)[5]) {
  // w is Conv1d.weight [C,1,5]; channel ch..ch+V-1 are 5*V consecutive floats
  const float* p = w + (size_t)ch * 5;
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int j = 0; j < 5; ++j) tap[j][e] = __ldg(p + e * 5 + j);
}

// 5-tap depthwise convolution of V channels; pairs of channels go through the packed fp32x2 pipe
// (FFMA2 on sm_100: one issue slot for two channels)
template <int V>
__device__ __forceinline__ vf<V> conv5(const vf<V> (&tap)[5], const vf<V>& x0, const vf<V>& x1,
                                       const vf<V>& x2, const vf<V>& x3, const vf<V>& x4) {
  vf<V> r;
  if constexpr (V % 2 == 0) {
#pragma unroll
    for (int e = 0; e < V; e += 2) {
      float2 acc = __fmul2_rn(make_float2(tap[0][e], tap[0][e + 1]), make_float2(x0[e], x0[e + 1]));
      acc = __ffma2_rn(make_float2(tap[1][e], tap[1][e + 1]), make_float2(x1[e], x1[e + 1]), acc);
      acc = __ffma2_rn(make_float2(tap[2][e], tap[2][e + 1]), make_float2(x2[e], x2[e + 1]), acc);
      acc = __ffma2_rn(make_float2(tap[3][e], tap[3][e + 1]), make_float2(x3[e], x3[e + 1]), acc);
      acc = __ffma2_rn(make_float2(tap[4][e], tap[4][e + 1]), make_float2(x4[e], x4[e + 1]), acc);
      r[e] = acc.x;
      r[e + 1] = acc.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float acc = tap[0][e] * x0[e];
      acc = fmaf(tap[1][e], x1[e], acc);
      acc = fmaf(tap[2][e], x2[e], acc);
      acc = fmaf(tap[3][e], x3[e], acc);
      acc = fmaf(tap[4][e], x4[e], acc);
      r[e] = acc;
    }
  }
  return r;
}

// copy the V=4 channels of one row (16 B fp32 / 8 B bf16) into this thread's shared-memory column
__device__ __forceinline__ void cp_async_act(ACT_T* dst, const ACT_T* src, bool valid) {
  if constexpr (sizeof(ACT_T) == 4) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(valid ? 16 : 0)
                 : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(valid ? 8 : 0)
                 : "memory");
  }
}

// ----------------------------------------------------------------------------- bulk (TMA 1-D) staging
// A CTA of 128 threads x 4 channels covers all 512 channels of its rows, so the rows of a chunk are ONE contiguous
// range of global memory and a ring stage [row][512] is its image: one elected thread fills a stage with three
// cp.async.bulk copies that complete on an mbarrier, instead of every thread issuing 20 cp.async of its own
// (~150 of the ~1000 instructions a chunk of la_stream_kernel costs per thread were address arithmetic and copies).
// full[s]: the stage has landed (transaction bytes); empty[s]: one arrival per warp after its last read of the stage.
__device__ __forceinline__ uint32_t sb_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb_u32(bar)) : "memory");
}
__device__ __forceinline__ bool sb_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(sb_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void sb_wait(uint64_t* bar, uint32_t parity) {
  while (!sb_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void sb_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb_u32(dst)),
               "l"(src), "r"(bytes), "r"(sb_u32(bar))
               : "memory");
}

// The stages are read with ordinary ld.shared (generic proxy) and refilled by cp.async.bulk (async proxy).  The
// mbarrier hand-over alone does not order the two proxies: without this fence the refill of a stage could overtake the
// last reads of it - found with the deterministic-statistics mode as 2 deviating forwards in 16,000 (256 bytes of one
// staged row from the NEXT fill of the stage; profiles/r04_bulk_ring_race.txt).  Every thread fences its own reads
// before its warp signals the stage empty (the pattern of CUTLASS's TMA-load epilogues: fence_view_async_shared()
// ahead of consumer_release).
__device__ __forceinline__ void sb_fence_reads() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Two-stage ring protocol of the streaming kernels.  In the LA kernels stage 1 has one extra phase at the start: the
// four rows ahead of the CTA's first chunk are parked in it while chunk 0 lands in stage 0.
struct SbRing {
  uint64_t* full;   // [2]
  uint64_t* empty;  // [2]
  // the copying thread: stage k & 1 may be overwritten with chunk k (PARK: stage 1 began with the parked rows)
  template <bool PARK = true>
  __device__ __forceinline__ void acquire(int k) const {
    const int s = k & 1, u = k >> 1;
    sb_wait(empty + s, (PARK && s) ? (u & 1) : ((u & 1) ^ 1));
  }
  __device__ __forceinline__ void wait_full(int k) const { sb_wait(full + (k & 1), (k >> 1) & 1); }
  // every warp, after its last read of stage s
  __device__ __forceinline__ void release(int s) const {
    sb_fence_reads();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) sb_arrive(empty + s);
  }
};
// thread 0 of a CTA of `nwarps` warps; the caller synchronises the CTA afterwards
__device__ __forceinline__ SbRing sb_ring_init(uint64_t* bars, int nwarps) {
  if (threadIdx.x == 0) {
    sb_init(bars + 0, 1); sb_init(bars + 1, 1);            // full: the copying thread's arrive.expect_tx
    sb_init(bars + 2, nwarps); sb_init(bars + 3, nwarps);  // empty: one arrival per warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  return SbRing{bars, bars + 2};
}
// The same protocol with `ns` stages (2 .. SB_MAX_STAGES) and a prefetch distance of ns - 1 chunks: what limits these
// streaming kernels is the number of bytes in flight per SM (bandwidth x latency, ~90 KB per SM at 6.4 TB/s), and with
// bulk copies a deeper ring costs shared memory only - no registers, no instructions.  PARK: the rows ahead of the
// CTA's first chunk are parked in the last stage, which therefore starts with one extra phase.
constexpr int SB_MAX_STAGES = 4;
struct SbCursor {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void next(int ns) {
    if (++stage == ns) { stage = 0; phase ^= 1; }
  }
};
struct SbRingN {
  uint64_t* full;   // [ns]
  uint64_t* empty;  // [ns]
  int ns;
  template <bool PARK>
  __device__ __forceinline__ void acquire(const SbCursor& c) const {   // the copying thread
    sb_wait(empty + c.stage, (PARK && c.stage == ns - 1) ? c.phase : (c.phase ^ 1));
  }
  __device__ __forceinline__ void wait_full(const SbCursor& c) const { sb_wait(full + c.stage, c.phase); }
  __device__ __forceinline__ void release(int stage) const {          // every warp, after its last read of the stage
    sb_fence_reads();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) sb_arrive(empty + stage);
  }
};
__device__ __forceinline__ SbRingN sb_ringn_init(uint64_t* bars, int ns, int nwarps) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) { sb_init(bars + s, 1); sb_init(bars + ns + s, nwarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  return SbRingN{bars, bars + ns, ns};
}
constexpr size_t SB_BAR_BYTES = 2 * SB_MAX_STAGES * sizeof(uint64_t);
// stages of the bulk rings of the two statistics kernels (TDANET_BULK_STAGES).  Measured on B200 (B = 64): deeper
// rings are SLOWER - 2 / 3 / 4 stages: 17.59 / 18.16 / 18.78 ms per step, every streaming role slower - so more
// bytes in flight is not what these kernels lack; 2 is the default.  spp_dw keeps two stages unconditionally
// (slower with three; its intermittent parity failures at three stages predate sb_fence_reads()).
static inline int bulk_stages() {
  static const int n = getenv("TDANET_BULK_STAGES") ? atoi(getenv("TDANET_BULK_STAGES")) : 2;
  return n < 2 ? 2 : (n > SB_MAX_STAGES ? SB_MAX_STAGES : n);
}
// which streaming kernels stage their rows with cp.async.bulk (bit mask: 1 la_stream, 2 local statistics, 4 global
// statistics, 8 spp_dw); TDANET_BULK=0 restores the per-thread cp.async rings everywhere (A/B measurements)
static inline int bulk_mask() {
  static const int m = getenv("TDANET_BULK") ? atoi(getenv("TDANET_BULK")) : 15;
  return m;
}

// ----------------------------------------------------------------------------- dw k=5
template <int KIND, int V, int NW, int S, int R, bool WRITE, bool STATS, bool EDGE>
__device__ __forceinline__ void dw5_body(const DwArgs& a, int b, int ch, int t0, int t1, const int* jtab, int tab0,
                                         float (&tot1)[NW], float (&tot2)[NW]) {
  Src<KIND, V, EDGE> src;
  src.init(a.src, b, ch, a.C, jtab, tab0);
  vf<V> tap[NW][5], bias[NW], s1[NW], s2[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    load_taps<V>(a.w[i], ch, tap[i]);
    bias[i] = a.bias[i] ? vload<V>(a.bias[i] + ch) : vzero<V>();
    s1[i] = vzero<V>();
    s2[i] = vzero<V>();
  }
  constexpr int NR = (R - 1) * S + 5;  // input rows feeding R outputs
  constexpr int CARRY = 5 - S;         // rows shared with the next chunk
  vf<V> xr[NR];
#pragma unroll
  for (int i = 0; i < CARRY; ++i) {
    const int t = t0 * S - 2 + i;
    xr[R * S + i] = src.finalize(src.load_raw(t), t);
  }
  ACT_T* outp = WRITE ? reinterpret_cast<ACT_T*>(a.out) + (size_t)b * a.Lout * a.C + ch : nullptr;

  for (int t = t0; t < t1; t += R) {
#pragma unroll
    for (int i = 0; i < CARRY; ++i) xr[i] = xr[R * S + i];
    const int base = t * S - 2 + CARRY;
#pragma unroll
    for (int i = 0; i < R * S; ++i) xr[CARRY + i] = src.load_raw(base + i);
#pragma unroll
    for (int i = 0; i < R * S; ++i) xr[CARRY + i] = src.finalize(xr[CARRY + i], base + i);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!EDGE || t + r < t1) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          vf<V> y = vadd<V>(conv5<V>(tap[i], xr[r * S], xr[r * S + 1], xr[r * S + 2], xr[r * S + 3], xr[r * S + 4]), bias[i]);
          if constexpr (STATS) {
            s1[i] = vadd<V>(s1[i], y);
            s2[i] = vfma<V>(y, y, s2[i]);
          }
          if constexpr (WRITE) {
            if (a.relu) {
#pragma unroll
              for (int e = 0; e < V; ++e) y[e] = fmaxf(y[e], 0.f);
            }
            if (a.round_out) vround_tf32<V>(y);
            astore<V>(outp + (t + r) * a.C, y);
          }
        }
      }
    }
  }
  if constexpr (STATS) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      if (a.chstats) {
        float* sp = a.chstats + ((size_t)(b * NW + i) * 2) * a.C + ch;
        vstat_add<V>(a.det, sp, s1[i]);
        vstat_add<V>(a.det, sp + a.C, s2[i]);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        tot1[i] += s1[i][e];
        tot2[i] += s2[i][e];
      }
    }
  }
}

// per-item totals: block reduction, then one pair of double atomics per CTA
template <int NW>
__device__ __forceinline__ void flush_item_stats(double* stats, int b, const float (&tot1)[NW], const float (&tot2)[NW],
                                                 double* red, const DetRef& det = DetRef{}) {
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    double d1 = tot1[i], d2 = tot2[i];
    block_sum2(d1, d2, red);
    if (threadIdx.x == 0) stat_add2(det, stats + ((size_t)b * NW + i) * 2, d1, d2);
  }
}

template <int KIND, int V, int NW, int S, int R, bool WRITE, bool STATS>
__global__ void __launch_bounds__(256) dw5_kernel(DwArgs a, int rows_per_cta) {
  grid_dep_wait();
  extern __shared__ int jtab[];  // inject kinds: nearest rows of the input rows this CTA touches
  __shared__ double red[64];
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, a.Lout);
  const int in_first = t0 * S - 2, in_last = (t1 - 1) * S + 2;
  constexpr bool inj = KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD;
  if constexpr (inj) {
    fill_nearest(jtab, rows_per_cta * S + 4, in_first, a.src.L, a.src.gscale, a.src.Lg);
    __syncthreads();
  }
  float tot1[NW], tot2[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) tot1[i] = tot2[i] = 0.f;
  if (ch < a.C) {
    // interior: whole chunks only, every input row inside the tensor
    const bool interior = in_first >= 0 && in_last < a.src.L && (t1 - t0) % R == 0;
    if (interior) dw5_body<KIND, V, NW, S, R, WRITE, STATS, false>(a, b, ch, t0, t1, jtab, in_first, tot1, tot2);
    else dw5_body<KIND, V, NW, S, R, WRITE, STATS, true>(a, b, ch, t0, t1, jtab, in_first, tot1, tot2);
  }
  if constexpr (STATS) flush_item_stats<NW>(a.stats, b, tot1, tot2, red, a.det);
}

// ----------------------------------------------------------------------------- dw k=5 with fused pooling
// spp_dw[k] for BEST / MULTRES: besides the raw output and its statistics, the kernel emits the
// adaptive-average-pooled raw output P_k [B, Lb, C] (bin j = rows [floor(j*L/Lb), ceil((j+1)*L/Lb)),
// F.adaptive_avg_pool1d).  Pooling commutes with the per-channel affine GlobLN, so
//   sum_k avgpool(gLN_k(out_k)) = sum_k (scale_k * P_k + shift_k)         (TDANet_best.py:358-364)
// and the separate pass that re-read every out_k disappears.  Tiles are whole bins: a tile computes
// rows [lo(ja), hi(jb-1)) and writes / counts rows [lo(ja), lo(jb)), so every bin is owned by one
// thread and stored plainly (no atomics, no zero-fill); at most one row per tile is computed twice.
// RING (stages, 0 = off): the R*S new input rows of a chunk are prefetched RING-1 chunks ahead with cp.async into a ring of
// thread-private shared-memory columns (`ringcol`), so loads stay in flight while the current chunk is computed.
template <int KIND, int S, bool EDGE, int RING, bool BULK = false>
__device__ __forceinline__ void dw5_pool_body(const DwArgs& a, int b, int ch, int t0, int tw, int tc, const int2* bins,
                                              int ja, int nb, ACT_T* ringcol, float& tot1, float& tot2,
                                              SbRingN rb = SbRingN{}) {
  static_assert(!BULK || (!EDGE && RING == 2), "bulk staging: interior CTAs (RING: any non-zero value, rb.ns stages)");
  constexpr int V = 4, R = (S == 1) ? 8 : 4;
  Src<KIND, V, EDGE> src;
  src.init(a.src, b, ch, a.C, nullptr, 0);
  vf<V> tap[5];
  load_taps<V>(a.w[0], ch, tap);
  const vf<V> bias = a.bias[0] ? vload<V>(a.bias[0] + ch) : vzero<V>();
  vf<V> s1 = vzero<V>(), s2 = vzero<V>(), acc = vzero<V>();
  int jrel = 0, hi_cur = bins[0].y;
  float inv_n = 1.f / (float)(bins[0].y - bins[0].x);
  float* pool = a.pool_out + ((size_t)b * a.Lb + ja) * a.C + ch;
  constexpr int NR = (R - 1) * S + 5, CARRY = 5 - S;
  vf<V> xr[NR];
  const int colw = blockDim.x * V;
  SbCursor pc, cc;  // BULK: producer (thread 0) / consumer positions in the ring
  auto issue = [&](int t) {  // the new input rows of the chunk of output rows t .. t+R-1
    const int base = t * S - 2 + CARRY;
    if constexpr (BULK) {  // the CTA covers all 512 channels: the chunk's rows are one contiguous range
      if (threadIdx.x == 0) {
        rb.acquire<false>(pc);
        constexpr uint32_t bytes = R * S * 512 * sizeof(ACT_T);
        sb_expect_tx(rb.full + pc.stage, bytes);
        sb_bulk_g2s(ringcol + (pc.stage * (R * S)) * colw, src.x + (size_t)base * src.C, bytes, rb.full + pc.stage);
        pc.next(rb.ns);
      }
      return;
    }
    ACT_T* st = ringcol + ((((t - t0) / R) % (RING ? RING : 1)) * (R * S)) * colw;
#pragma unroll
    for (int i = 0; i < R * S; ++i) {
      const int row = base + i;
      const bool ok = !EDGE || (row >= 0 && row < src.L);
      cp_async_act(st + i * colw, src.x + (ok ? row : 0) * src.C, ok);
    }
  };
  if constexpr (BULK) {
    for (int k = 0; k < rb.ns - 1; ++k)
      if (t0 + k * R < tc) issue(t0 + k * R);
  } else if constexpr (RING != 0) {
    issue(t0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if constexpr (RING == 3) {
      if (t0 + R < tc) issue(t0 + R);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
#pragma unroll
  for (int i = 0; i < CARRY; ++i) {
    const int t = t0 * S - 2 + i;
    xr[R * S + i] = src.finalize(src.load_raw(t), t);
  }
  ACT_T* outp = reinterpret_cast<ACT_T*>(a.out) + (size_t)b * a.Lout * a.C + ch;
  for (int t = t0; t < tc; t += R) {
#pragma unroll
    for (int i = 0; i < CARRY; ++i) xr[i] = xr[R * S + i];
    const int base = t * S - 2 + CARRY;
    if constexpr (BULK) {
      if (t + (rb.ns - 1) * R < tc) issue(t + (rb.ns - 1) * R);
      rb.wait_full(cc);
      const ACT_T* st = ringcol + (cc.stage * (R * S)) * colw;
#pragma unroll
      for (int i = 0; i < R * S; ++i) xr[CARRY + i] = alds<V>(st + i * colw);
      rb.release(cc.stage);
      cc.next(rb.ns);
    } else if constexpr (RING != 0) {
      if (t + (RING - 1) * R < tc) issue(t + (RING - 1) * R);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if constexpr (RING == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      const ACT_T* st = ringcol + ((((t - t0) / R) % (RING ? RING : 1)) * (R * S)) * colw;
#pragma unroll
      for (int i = 0; i < R * S; ++i) xr[CARRY + i] = alds<V>(st + i * colw);
    } else {
#pragma unroll
      for (int i = 0; i < R * S; ++i) xr[CARRY + i] = src.load_raw(base + i);
    }
#pragma unroll
    for (int i = 0; i < R * S; ++i) xr[CARRY + i] = src.finalize(xr[CARRY + i], base + i);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = t + r;
      if (row < tc) {
        const vf<V> y = vadd<V>(conv5<V>(tap, xr[r * S], xr[r * S + 1], xr[r * S + 2], xr[r * S + 3], xr[r * S + 4]), bias);
        if (row < tw) {
          s1 = vadd<V>(s1, y);
          s2 = vfma<V>(y, y, s2);
          astore<V>(outp + row * a.C, y);
        }
        acc = vadd<V>(acc, y);
        if (row == hi_cur - 1) {
          vf<V> m;
#pragma unroll
          for (int e = 0; e < V; ++e) m[e] = acc[e] * inv_n;
          vstore<V>(pool + jrel * a.C, m);
          ++jrel;
          const int2 nx = bins[jrel];  // the table has one entry past the tile's last bin
          const bool both = jrel < nb && nx.x <= row;  // this row is also the first row of the next bin
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = both ? y[e] : 0.f;
          hi_cur = nx.y;
          inv_n = 1.f / (float)(nx.y - nx.x);
        }
      }
    }
  }
  if (a.chstats) {
    float* sp = a.chstats + ((size_t)b * 2) * a.C + ch;
    vstat_add<V>(a.det, sp, s1);
    vstat_add<V>(a.det, sp + a.C, s2);
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    tot1 += s1[e];
    tot2 += s2[e];
  }
}

template <int KIND, int S, int LB, int RING, bool BULK = false>
__global__ void __launch_bounds__(LB ? 128 : 256, LB ? LB : 1) dw5_pool_kernel(DwArgs a, int bins_per_cta, int ns) {
  // [ring: RING (BULK: ns) stages x R*S rows x blockDim.x*4 ACT_T][bins][BULK: full[ns], empty[ns] mbarriers]
  extern __shared__ __align__(16) unsigned char pool_smem[];
  __shared__ double red[64];
  constexpr int V = 4, R = (S == 1) ? 8 : 4;
  const size_t ring_b = (size_t)(BULK ? ns : RING) * R * S * blockDim.x * V * sizeof(ACT_T);
  int2* bins = reinterpret_cast<int2*>(pool_smem + ring_b);  // (lo, hi) of the tile's bins, plus one
  ACT_T* ringcol = reinterpret_cast<ACT_T*>(pool_smem) + threadIdx.x * V;
  const int b = a.rev ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int L = a.Lout, Lb = a.Lb;
  const int ja = blockIdx.x * bins_per_cta, jb = min(ja + bins_per_cta, Lb), nb = jb - ja;
  for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
    const long j = ja + i;
    bins[i] = make_int2((int)((j * L) / Lb), (int)(((j + 1) * L + Lb - 1) / Lb));
  }
  SbRingN rb{};
  if constexpr (BULK) rb = sb_ringn_init(reinterpret_cast<uint64_t*>(bins + bins_per_cta + 1), ns, blockDim.x / 32);
  __syncthreads();
  grid_dep_wait();   // the bin table does not depend on the previous kernel
  const int t0 = bins[0].x;
  const int tw = jb < Lb ? bins[nb].x : L;  // rows written / counted by this tile
  const int tc = bins[nb - 1].y;            // rows computed (the last bin may reach one row further)
  float tot1[1] = {0.f}, tot2[1] = {0.f};
  if (ch < a.C) {
    const int chunks = (tc - t0 + R - 1) / R;
    const bool interior = t0 * S - 2 >= 0 && (t0 + chunks * R - 1) * S + 2 < a.src.L;
    if (interior) dw5_pool_body<KIND, S, false, RING, BULK>(a, b, ch, t0, tw, tc, bins, ja, nb, ringcol, tot1[0], tot2[0], rb);
    else dw5_pool_body<KIND, S, true, RING>(a, b, ch, t0, tw, tc, bins, ja, nb, ringcol, tot1[0], tot2[0]);
  }
  flush_item_stats<1>(a.stats, b, tot1, tot2, red, a.det);
}

static void pick_tiling(int B, int L, int ctiles, int R, int* rows_per_cta, int* tiles, long target = 148L * 16,
                        int cap = 64) {
  // aim at `target` CTAs in total, at most `cap` rows per CTA, whole chunks of R rows
  // (tuning aid: TDANET_TILE_TARGET / TDANET_TILE_CAP override the defaults of the dw5 / dw5_pool launchers)
  if (target == 148L * 16 && cap == 64) {
    static const long env_target = getenv("TDANET_TILE_TARGET") ? atol(getenv("TDANET_TILE_TARGET")) : 0;
    static const int env_cap = getenv("TDANET_TILE_CAP") ? atoi(getenv("TDANET_TILE_CAP")) : 0;
    if (env_target > 0) target = env_target;
    if (env_cap > 0) cap = env_cap;
  }
  long per = ((long)B * L * ctiles + target - 1) / target;
  per = (per + R - 1) / R * R;
  if (per < R) per = R;
  if (per > cap) per = cap;
  *rows_per_cta = (int)per;
  *tiles = cdiv(L, (int)per);
}

template <int KIND, int V, int NW, int S, bool WRITE, bool STATS>
static int launch_dw5_t(const DwArgs& a, cudaStream_t st) {
  constexpr int R = (S == 1) ? 8 : 4;
  int threads = a.C / V;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / V, threads);
  int rows, tiles;
  pick_tiling(a.B, a.Lout, ctiles, R, &rows, &tiles);
  dim3 grid(tiles, ctiles, a.B);
  const size_t smem = (KIND == SRC_INJECT_GATE || KIND == SRC_INJECT_ADD) ? (size_t)(rows * S + 4) * sizeof(int) : 0;
  TD_LAUNCH((dw5_kernel<KIND, V, NW, S, R, WRITE, STATS>), grid, threads, smem, st, a, rows);
  return 0;
}

template <int KIND, int V, int NW>
static int launch_dw5_k(const DwArgs& a, cudaStream_t st) {
  const bool wr = a.out != nullptr, stt = a.stats != nullptr;
  if (a.stride == 1) {
    if (wr && stt) return launch_dw5_t<KIND, V, NW, 1, true, true>(a, st);
    if (wr) return launch_dw5_t<KIND, V, NW, 1, true, false>(a, st);
    return launch_dw5_t<KIND, V, NW, 1, false, true>(a, st);
  }
  if constexpr (KIND == SRC_AFFINE && NW == 1) {
    if (wr && stt) return launch_dw5_t<KIND, V, NW, 2, true, true>(a, st);
    if (wr) return launch_dw5_t<KIND, V, NW, 2, true, false>(a, st);
  }
  return fail(TDANET_EINVAL, "dw5: unsupported stride/kind combination");
}

template <int KIND, int S>
static int launch_dw5_pool_t(const DwArgs& a, cudaStream_t st) {
  int threads = a.C / 4;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / 4, threads);
  int rows, tiles;
  static const long pool_target = getenv("TDANET_POOL_TARGET") ? atol(getenv("TDANET_POOL_TARGET")) : 148L * 4;
  static const int pool_cap = getenv("TDANET_POOL_CAP") ? atoi(getenv("TDANET_POOL_CAP")) : 128;
  pick_tiling(a.B, a.Lout, ctiles, S == 1 ? 8 : 4, &rows, &tiles, pool_target, pool_cap);
  // the short scales: a CTA pays its prologue (taps, GlobLN coefficients, bin table) for at least this many rows
  static const int min_rows = getenv("TDANET_POOL_MINROWS") ? atoi(getenv("TDANET_POOL_MINROWS")) : 8;
  if (rows < min_rows) rows = min_rows;
  int bpt = (int)(((long)rows * a.Lb + a.Lout / 2) / a.Lout);  // bins per tile ~ rows / (L / Lb)
  if (bpt < 1) bpt = 1;
  dim3 grid(cdiv(a.Lb, bpt), ctiles, a.B);
  static const int lb = getenv("TDANET_POOL_LB") ? atoi(getenv("TDANET_POOL_LB")) : 4;  // 4 CTAs of 128 threads per SM (128 registers)
  static const int ring_env = getenv("TDANET_POOL_RING") ? atoi(getenv("TDANET_POOL_RING")) : 2;  // stages (0: direct loads)
  constexpr int R = S == 1 ? 8 : 4;
  const size_t bins_b = (size_t)(bpt + 1) * sizeof(int2);
  int ring = ring_env == 1 ? 2 : ring_env;
  while (ring > 0 && (size_t)ring * R * S * threads * 4 * sizeof(ACT_T) + bins_b > 100 * 1024) --ring;
  if (ring == 1) ring = 0;
  const size_t ring_b = (size_t)ring * R * S * threads * 4 * sizeof(ACT_T);
  static PerDeviceOnce first_use;
  if (first_use()) {
    TD_CUDA(cudaFuncSetAttribute(dw5_pool_kernel<KIND, S, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    TD_CUDA(cudaFuncSetAttribute(dw5_pool_kernel<KIND, S, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    TD_CUDA(cudaFuncSetAttribute(dw5_pool_kernel<KIND, S, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    TD_CUDA(cudaFuncSetAttribute(dw5_pool_kernel<KIND, S, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    TD_CUDA(cudaFuncSetAttribute(dw5_pool_kernel<KIND, S, 4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  const bool lb4 = lb == 4 && threads <= 128;
  if (ring == 2 && lb4 && a.C == 512 && threads == 128 && (bulk_mask() & 8)) {
    // two stages: three were slower, so the depth is not a knob here.  (With three the B = 64 parity tests also failed
    // intermittently at 1e-3 in the second session - in hindsight the proxy race sb_fence_reads() closes, which a
    // deeper ring makes likelier; not re-measured since.)
    const int ns = 2;
    TD_LAUNCH((dw5_pool_kernel<KIND, S, 4, 2, true>), grid, threads, ring_b / 2 * ns + bins_b + SB_BAR_BYTES, st, a, bpt, ns);
    return 0;
  }
#define TD_POOL_LAUNCH(LB_, RING_) \
  TD_LAUNCH((dw5_pool_kernel<KIND, S, LB_, RING_>), grid, threads, ring_b + bins_b, st, a, bpt, 0)
  if (ring == 3) { if (lb4) { TD_POOL_LAUNCH(4, 3); } else { TD_POOL_LAUNCH(0, 3); } }
  else if (ring == 2) { if (lb4) { TD_POOL_LAUNCH(4, 2); } else { TD_POOL_LAUNCH(0, 2); } }
  else { if (lb4) { TD_POOL_LAUNCH(4, 0); } else { TD_POOL_LAUNCH(0, 0); } }
#undef TD_POOL_LAUNCH
  return 0;
}

static bool gstats_stream_applies(const DwArgs& a);
static int launch_gstats_stream(const DwArgs& a, cudaStream_t st);

int launch_dw5(const DwArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "dw5: C=%d must be a multiple of 4", a.C);
  if (gstats_stream_applies(a)) return launch_gstats_stream(a, st);
  if (a.pool_out) {
    TD_REQUIRE(a.nw == 1 && a.out && a.stats && !a.relu && !a.round_out && a.Lb > 0 && a.Lb <= a.Lout,
               "dw5: pooled output needs nw == 1, out, stats and Lb <= Lout");
    TD_REQUIRE((long)a.src.L * a.C < (1L << 31) && (long)a.Lout * a.C < (1L << 31), "dw5: item too large for 32-bit offsets");
    if (a.kind == SRC_AFFINE_PRELU && a.stride == 1) return launch_dw5_pool_t<SRC_AFFINE_PRELU, 1>(a, st);
    if (a.kind == SRC_AFFINE && a.stride == 2) return launch_dw5_pool_t<SRC_AFFINE, 2>(a, st);
    if (a.kind == SRC_AFFINE && a.stride == 1) return launch_dw5_pool_t<SRC_AFFINE, 1>(a, st);
    return fail(TDANET_EINVAL, "dw5: pooled output unsupported for kind %d stride %d", a.kind, a.stride);
  }
  TD_REQUIRE(a.nw == 1 || a.nw == 2, "dw5: nw=%d", a.nw);
  TD_REQUIRE(!(a.nw == 2 && a.out), "dw5: writing needs nw == 1");
  TD_REQUIRE(a.out || a.stats, "dw5: nothing to do");
  TD_REQUIRE((long)a.src.L * a.C < (1L << 31) && (long)a.Lout * a.C < (1L << 31), "dw5: item too large for 32-bit offsets");
  if (a.nw == 1) {
    switch (a.kind) {
      case SRC_PLAIN: return launch_dw5_k<SRC_PLAIN, 4, 1>(a, st);
      case SRC_AFFINE: return launch_dw5_k<SRC_AFFINE, 4, 1>(a, st);
      case SRC_AFFINE_PRELU: return launch_dw5_k<SRC_AFFINE_PRELU, 4, 1>(a, st);
      case SRC_INJECT_GATE: return launch_dw5_k<SRC_INJECT_GATE, 4, 1>(a, st);
      case SRC_INJECT_ADD: return launch_dw5_k<SRC_INJECT_ADD, 4, 1>(a, st);
    }
  } else {
    switch (a.kind) {
      case SRC_PLAIN: return launch_dw5_k<SRC_PLAIN, 4, 2>(a, st);
      case SRC_INJECT_GATE: return launch_dw5_k<SRC_INJECT_GATE, 2, 2>(a, st);
      case SRC_INJECT_ADD: return launch_dw5_k<SRC_INJECT_ADD, 4, 2>(a, st);
    }
  }
  return fail(TDANET_EINVAL, "dw5: unsupported source kind %d (nw=%d)", a.kind, a.nw);
}

// ----------------------------------------------------------------------------- materialise an injected operand
// y[t] = x_fused[k][t] written out (only for the two small tensors of the first top-down step, whose
// down-sampling access pattern would otherwise recompute every injected row five times)
// Optional: the GlobLN statistics of two 5-tap depthwise convolutions of the tensor being written (global_act /
// global_embedding of the first top-down step, whose "global" operand this is): the CTA finalises two halo rows on
// either side of its range and slides a 5-row window over what it writes, so the separate statistics pass over the
// materialised tensor (one launch per block) disappears.
struct MatStats {
  const float* wa;   // [C,1,5] or null: no statistics
  const float* we;
  double* stats;     // [B, 2, 2]
};

template <int KIND, bool STATS>
__device__ __forceinline__ void inject_materialize_tile(const SrcDesc& sd, int C, float* __restrict__ out, int tile,
                                                        int rows_per_cta, int* jtab, const MatStats& ms) {
  constexpr int V = 4, R = 8;  // 8 rows of loads in flight per thread
  constexpr int H = STATS ? 2 : 0;  // halo rows finalised (not stored) on either side
  __shared__ double red[STATS ? 64 : 1];
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int t0 = tile * rows_per_cta, t1 = min(t0 + rows_per_cta, sd.L);
  fill_nearest(jtab, rows_per_cta + 2 * H, t0 - H, sd.L, sd.gscale, sd.Lg);
  __syncthreads();
  float tot1[2] = {0.f, 0.f}, tot2[2] = {0.f, 0.f};
  if (ch < C) {
    Src<KIND, V, false> src;
    src.init(sd, b, ch, C, jtab, t0 - H);
    ACT_T* op = reinterpret_cast<ACT_T*>(out) + (size_t)b * sd.L * C + ch;
    vf<V> wa[5], we[5], win[5], s1a = vzero<V>(), s2a = vzero<V>(), s1e = vzero<V>(), s2e = vzero<V>();
    if constexpr (STATS) {
      load_taps<V>(ms.wa, ch, wa);
      load_taps<V>(ms.we, ch, we);
#pragma unroll
      for (int i = 0; i < 5; ++i) win[i] = vzero<V>();
    }
    if constexpr (!STATS) {   // the plain copy: every row of the range is inside the tensor
      for (int t = t0; t < t1; t += R) {
        vf<V> r[R];
#pragma unroll
        for (int i = 0; i < R; ++i) r[i] = t + i < t1 ? src.load_raw(t + i) : vzero<V>();
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (t + i < t1) astore<V>(op + (t + i) * C, src.finalize(r[i], t + i));
      }
    }
    for (int t = t0 - H; STATS && t < t1 + H; t += R) {
      vf<V> r[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int row = t + i;
        r[i] = (row < t1 + H && row >= 0 && row < sd.L) ? src.load_raw(row) : vzero<V>();
      }
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int row = t + i;
        if (row < t1 + H) {
          const bool inside = row >= 0 && row < sd.L;   // rows outside the tensor are the convolutions' zero padding
          const vf<V> f = inside ? src.finalize(r[i], row) : vzero<V>();
          if (row >= t0 && row < t1) astore<V>(op + row * C, f);
          if constexpr (STATS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) win[j] = win[j + 1];
            win[4] = f;
            const int c = row - 2;   // the window now holds rows c-2 .. c+2
            if (c >= t0 && c < t1) {
              const vf<V> ya = conv5<V>(wa, win[0], win[1], win[2], win[3], win[4]);
              const vf<V> ye = conv5<V>(we, win[0], win[1], win[2], win[3], win[4]);
              s1a = vadd<V>(s1a, ya); s2a = vfma<V>(ya, ya, s2a);
              s1e = vadd<V>(s1e, ye); s2e = vfma<V>(ye, ye, s2e);
            }
          }
        }
      }
    }
    if constexpr (STATS) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        tot1[0] += s1a[e]; tot2[0] += s2a[e];
        tot1[1] += s1e[e]; tot2[1] += s2e[e];
      }
    }
  }
  if constexpr (STATS) flush_item_stats<2>(ms.stats, b, tot1, tot2, red);
}

template <int KIND, bool STATS>
__global__ void __launch_bounds__(256) inject_materialize_kernel(SrcDesc sd, int C, float* __restrict__ out, int rows_per_cta,
                                                                 MatStats ms) {
  grid_dep_wait();
  extern __shared__ int jtab[];
  inject_materialize_tile<KIND, STATS>(sd, C, out, blockIdx.x, rows_per_cta, jtab, ms);
}

// two tensors in one launch (the two operands of the first top-down step): tiles [0, tiles_a) belong to a; the
// statistics (if any) are those of tensor b, the step's "global" operand
template <int KIND, bool STATS>
__global__ void __launch_bounds__(256) inject_materialize2_kernel(SrcDesc sa, float* __restrict__ out_a, int tiles_a,
                                                                  SrcDesc sb, float* __restrict__ out_b, int C,
                                                                  int rows_per_cta, MatStats ms) {
  grid_dep_wait();
  extern __shared__ int jtab[];
  if ((int)blockIdx.x < tiles_a) inject_materialize_tile<KIND, false>(sa, C, out_a, blockIdx.x, rows_per_cta, jtab, ms);
  else inject_materialize_tile<KIND, STATS>(sb, C, out_b, blockIdx.x - tiles_a, rows_per_cta, jtab, ms);
}

static int materialize_rows(bool stats) {
  static const int rows = getenv("TDANET_MAT_ROWS") ? atoi(getenv("TDANET_MAT_ROWS")) : 16;
  return stats ? 2 * rows : rows;   // longer ranges where every CTA pays four halo rows
}

int launch_inject_materialize(const SrcDesc& src, int kind, int B, int C, float* out, const float* wa, const float* we,
                              double* stats, cudaStream_t st) {  // out: ACT_T
  TD_REQUIRE(C % 4 == 0 && (long)src.L * C < (1L << 31), "inject_materialize: C=%d L=%d", C, src.L);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  const bool with_stats = stats != nullptr;
  const int rows = materialize_rows(with_stats);
  const MatStats ms{wa, we, stats};
  dim3 grid(cdiv(src.L, rows), cdiv(C / 4, threads), B);
  const size_t smem = (rows + 4) * sizeof(int);
  if (kind == SRC_INJECT_GATE) {
    if (with_stats) TD_LAUNCH_COOP((inject_materialize_kernel<SRC_INJECT_GATE, true>), grid, threads, smem, st, src, C, out, rows, ms);
    else TD_LAUNCH_COOP((inject_materialize_kernel<SRC_INJECT_GATE, false>), grid, threads, smem, st, src, C, out, rows, ms);
  } else if (kind == SRC_INJECT_ADD) {
    if (with_stats) TD_LAUNCH_COOP((inject_materialize_kernel<SRC_INJECT_ADD, true>), grid, threads, smem, st, src, C, out, rows, ms);
    else TD_LAUNCH_COOP((inject_materialize_kernel<SRC_INJECT_ADD, false>), grid, threads, smem, st, src, C, out, rows, ms);
  } else {
    return fail(TDANET_EINVAL, "inject_materialize: kind %d", kind);
  }
  return 0;
}

int launch_inject_materialize2(const SrcDesc& sa, float* out_a, const SrcDesc& sb, float* out_b, int kind, int B, int C,
                               const float* wa, const float* we, double* stats, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0 && (long)sa.L * C < (1L << 31) && (long)sb.L * C < (1L << 31), "inject_materialize2: C=%d", C);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  const bool with_stats = stats != nullptr;
  const int rows = materialize_rows(with_stats);
  const MatStats ms{wa, we, stats};
  const int tiles_a = cdiv(sa.L, rows);
  dim3 grid(tiles_a + cdiv(sb.L, rows), cdiv(C / 4, threads), B);
  const size_t smem = (rows + 4) * sizeof(int);
  if (kind == SRC_INJECT_GATE) {
    if (with_stats) TD_LAUNCH_COOP((inject_materialize2_kernel<SRC_INJECT_GATE, true>), grid, threads, smem, st, sa, out_a, tiles_a, sb, out_b, C, rows, ms);
    else TD_LAUNCH_COOP((inject_materialize2_kernel<SRC_INJECT_GATE, false>), grid, threads, smem, st, sa, out_a, tiles_a, sb, out_b, C, rows, ms);
  } else if (kind == SRC_INJECT_ADD) {
    if (with_stats) TD_LAUNCH_COOP((inject_materialize2_kernel<SRC_INJECT_ADD, true>), grid, threads, smem, st, sa, out_a, tiles_a, sb, out_b, C, rows, ms);
    else TD_LAUNCH_COOP((inject_materialize2_kernel<SRC_INJECT_ADD, false>), grid, threads, smem, st, sa, out_a, tiles_a, sb, out_b, C, rows, ms);
  } else {
    return fail(TDANET_EINVAL, "inject_materialize2: kind %d", kind);
  }
  return 0;
}

// ----------------------------------------------------------------------------- generic dw (fork conv_pool)
template <int KIND>
__global__ void dw_generic_kernel(SrcDesc sd, int C, int Lout, int ks, int stride,
                                  const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int round_out) {
  grid_dep_wait();
  constexpr int V = 4;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int t = blockIdx.x;
  if (ch >= C || t >= Lout) return;
  Src<KIND, V, true> src;
  src.init(sd, b, ch, C, nullptr, 0);
  const int pad = (ks - 1) / 2;
  vf<V> acc = bias ? vload<V>(bias + ch) : vzero<V>();
  for (int j = 0; j < ks; ++j) {
    const int ti = t * stride - pad + j;
    vf<V> xv = src.finalize(src.load_raw(ti), ti);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = fmaf(__ldg(w + (size_t)(ch + e) * ks + j), xv[e], acc[e]);
  }
  if (round_out) vround_tf32<V>(acc);
  vstore<V>(out + ((size_t)b * Lout + t) * C + ch, acc);
}

// Input-stationary form of the same convolution for the fork's conv_pool (k = 2s+1, stride s = 2..16, pad s): a thread
// owns 4 channels and RO = 8 consecutive output rows and walks the (RO+1)*s + 1 input rows they cover once.  Input
// row base + q*s + ph (q = 0..RO, ph = 0..s-1) feeds output q through tap ph and output q-1 through tap s+ph, so the
// loop runs over the phase ph with q unrolled: two weight vectors (from the transposed copy wT [k][C]) and RO+1
// independent row loads per iteration, 8 FMAs per row and accumulators with static indices.  The rows q*s also
// close output q-2 through the last tap 2s (a short pass of its own).
// History: the output-stationary kernel re-read every row through ~2 windows with 4 scalar weight loads per tap
// (conv_pool 7.4 of the fork's 25 ms per step at B = 64); a first input-stationary version that tested all RO
// outputs per row moved the right bytes (292 MB for 263 MB) but issued 103 M warp instructions (ncu
// profiles/r01_ncu_dw_strided.txt: 201 us, issue-bound).
template <int KIND>
__global__ void __launch_bounds__(256) dw_strided_kernel(SrcDesc sd, int C, int Lout, int S,
                                                         const float* __restrict__ wT, const float* __restrict__ bias,
                                                         float* __restrict__ out, int round_out) {
  grid_dep_wait();
  constexpr int V = 4, RO = 8;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (ch >= C) return;
  Src<KIND, V, true> src;
  src.init(sd, b, ch, C, nullptr, 0);
  const int to0 = blockIdx.x * RO;
  const int base = to0 * S - S;  // input row of (q = 0, ph = 0); rows outside [0, L) are the conv's zero padding
  vf<V> acc[RO];
  const vf<V> bv = bias ? vload<V>(bias + ch) : vzero<V>();
#pragma unroll
  for (int r = 0; r < RO; ++r) acc[r] = bv;
  {  // last tap: output r takes row base + (r + 2) * S
    const vf<V> w2 = vload<V>(wT + (size_t)(2 * S) * C + ch);
    vf<V> xr[RO];
#pragma unroll
    for (int r = 0; r < RO; ++r) xr[r] = src.load_raw(base + (r + 2) * S);
#pragma unroll
    for (int r = 0; r < RO; ++r) {
      const vf<V> xv = src.finalize(xr[r], base + (r + 2) * S);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[r][e] = fmaf(w2[e], xv[e], acc[r][e]);
    }
  }
  for (int ph = 0; ph < S; ++ph) {
    const vf<V> wlo = vload<V>(wT + (size_t)ph * C + ch), whi = vload<V>(wT + (size_t)(S + ph) * C + ch);
    vf<V> xr[RO + 1];
#pragma unroll
    for (int q = 0; q <= RO; ++q) xr[q] = src.load_raw(base + q * S + ph);
#pragma unroll
    for (int q = 0; q <= RO; ++q) {
      const vf<V> xv = src.finalize(xr[q], base + q * S + ph);
      if (q < RO) {
#pragma unroll
        for (int e = 0; e < V; ++e) acc[q][e] = fmaf(wlo[e], xv[e], acc[q][e]);
      }
      if (q >= 1) {
#pragma unroll
        for (int e = 0; e < V; ++e) acc[q - 1][e] = fmaf(whi[e], xv[e], acc[q - 1][e]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RO; ++r) {
    if (to0 + r < Lout) {
      if (round_out) vround_tf32<V>(acc[r]);
      vstore<V>(out + ((size_t)b * Lout + to0 + r) * C + ch, acc[r]);
    }
  }
}

int launch_dw_generic(const SrcDesc& src, int kind, int B, int C, int Lout, int ks, int stride,
                      const float* w, const float* wT, const float* bias, float* out, int round_out, cudaStream_t st) {
  TD_REQUIRE(C % 4 == 0 && (ks & 1), "dw_generic: C=%d ks=%d", C, ks);
  int threads = C / 4 > 256 ? 256 : (C / 4 < 32 ? 32 : C / 4);
  if (wT && kind == SRC_AFFINE && stride >= 2 && ks == 2 * stride + 1) {
    dim3 grid(cdiv(Lout, 8), cdiv(C / 4, threads), B);
    TD_LAUNCH((dw_strided_kernel<SRC_AFFINE>), grid, threads, 0, st, src, C, Lout, stride, wT, bias, out, round_out);
    return 0;
  }
  dim3 grid(Lout, cdiv(C / 4, threads), B);
  if (kind == SRC_AFFINE) {
    TD_LAUNCH((dw_generic_kernel<SRC_AFFINE>), grid, threads, 0, st, src, C, Lout, ks, stride, w, bias, out, round_out);
  } else if (kind == SRC_PLAIN) {
    TD_LAUNCH((dw_generic_kernel<SRC_PLAIN>), grid, threads, 0, st, src, C, Lout, ks, stride, w, bias, out, round_out);
  } else {
    return fail(TDANET_EINVAL, "dw_generic: unsupported source kind %d", kind);
  }
  return 0;
}

// ----------------------------------------------------------------------------- LA combine
// GC > 0 (Lg <= Ll): the global-branch convs are evaluated once per distinct source row of a chunk
//          (at most GC of them) and parked in a thread-private shared-memory column, then gathered.
// GC = 0 (Lg >  Ll): only the first top-down step (reference quirk, TDANet_best.py:375-376);
//          evaluated per output row.
struct LaSmem {
  int* jc;   // [rows]      nearest global row of each output row
  int* jl;   // [rows + 4]  local inject: nearest g row of each local input row
  int* jg;   // [glo span]  global inject: nearest g row of each global input row
  float* scratch;
};

template <int LKIND, int GKIND, int V, int GC, bool EDGE>
__device__ __forceinline__ void la_body(const LaArgs& a, int b, int ch, int t0, int t1, const LaSmem& sm, int g_first) {
  constexpr int R = 8;
  const int Ll = a.loc.L;
  const int colw = blockDim.x * V;
  float* mine = sm.scratch + threadIdx.x * V;

  Src<LKIND, V, EDGE> sl;
  sl.init(a.loc, b, ch, a.C, sm.jl, t0 - 2);
  Src<GKIND, V, EDGE> sg;
  sg.init(a.glo, b, ch, a.C, sm.jg, g_first);
  vf<V> wl[5], wa[5], we[5];
  load_taps<V>(a.wl, ch, wl);
  load_taps<V>(a.wa, ch, wa);
  load_taps<V>(a.we, ch, we);
  vf<V> sL, hL, sA, hA, sE, hE;
  norm_coef<V>(a.nL, b, ch, sL, hL);
  norm_coef<V>(a.nA, b, ch, sA, hA);
  norm_coef<V>(a.nE, b, ch, sE, hE);

  vf<V> xr[R + 4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 - 2 + i;
    xr[R + i] = sl.finalize(sl.load_raw(t), t);
  }
  ACT_T* outp = reinterpret_cast<ACT_T*>(a.out) + (size_t)b * Ll * a.C + ch;

  for (int t = t0; t < t1; t += R) {
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = xr[R + i];
#pragma unroll
    for (int i = 0; i < R; ++i) xr[4 + i] = sl.load_raw(t + 2 + i);
    const int jlo = sm.jc[t - t0];
    // GC < 0: down-sampling by ~2 (the first top-down step): the 2R+4 global rows a chunk can touch are loaded up
    // front, like the local rows; output row r then takes its window at offset 2r or 2r-1 (see launch_la_t)
    vf<V> gr2[GC < 0 ? 2 * R + 4 : 1];
    if constexpr (GC < 0) {
#pragma unroll
      for (int i = 0; i < 2 * R + 4; ++i) gr2[i] = sg.load_raw(jlo - 2 + i);
#pragma unroll
      for (int i = 0; i < 2 * R + 4; ++i) gr2[i] = sg.finalize(gr2[i], jlo - 2 + i);
    }
    if constexpr (GC > 0) {
      // distinct centres of this chunk: jlo .. jc[last row]  (<= GC by construction)
      const int tl = (EDGE ? min(t + R, t1) : t + R) - 1;
      const int nc = sm.jc[tl - t0] - jlo + 1;
      vf<V> gr[GC + 4];
#pragma unroll
      for (int i = 0; i < GC + 4; ++i) gr[i] = sg.load_raw(jlo - 2 + i);
#pragma unroll
      for (int i = 0; i < GC + 4; ++i) gr[i] = sg.finalize(gr[i], jlo - 2 + i);
#pragma unroll
      for (int i = 0; i < GC; ++i) {
        if (i < nc) {
          vf<V> ca = vfma<V>(sA, conv5<V>(wa, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]), hA);
          const vf<V> ce = vfma<V>(sE, conv5<V>(we, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]), hE);
#pragma unroll
          for (int e = 0; e < V; ++e) ca[e] = sigmoidf_(ca[e]);
          vstore<V>(mine + (2 * i) * colw, ca);
          vstore<V>(mine + (2 * i + 1) * colw, ce);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) xr[4 + i] = sl.finalize(xr[4 + i], t + 2 + i);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!EDGE || t + r < t1) {
        vf<V> cl = conv5<V>(wl, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        vf<V> ga, ge;
        const int j = sm.jc[t + r - t0];
        if constexpr (GC <= 0) {
          vf<V> g5[5];
          if constexpr (GC < 0) {
            const bool even = j - jlo == 2 * r;  // else 2r - 1
#pragma unroll
            for (int i = 0; i < 5; ++i) {
#pragma unroll
              for (int e = 0; e < V; ++e) g5[i][e] = (r == 0 || even) ? gr2[2 * r + i][e] : gr2[(r == 0 ? 1 : 2 * r) - 1 + i][e];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) g5[i] = sg.load_raw(j - 2 + i);
#pragma unroll
            for (int i = 0; i < 5; ++i) g5[i] = sg.finalize(g5[i], j - 2 + i);
          }
          ga = conv5<V>(wa, g5[0], g5[1], g5[2], g5[3], g5[4]);
          ge = conv5<V>(we, g5[0], g5[1], g5[2], g5[3], g5[4]);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            ga[e] = sigmoidf_(fmaf(sA[e], ga[e], hA[e]));
            ge[e] = fmaf(sE[e], ge[e], hE[e]);
          }
        } else {
          const float* col = mine + (2 * (j - jlo)) * colw;
          if constexpr (V == 2) {
            const float2 u = *reinterpret_cast<const float2*>(col), w2 = *reinterpret_cast<const float2*>(col + colw);
            ga[0] = u.x; ga[1] = u.y; ge[0] = w2.x; ge[1] = w2.y;
          } else {
            const float4 u = *reinterpret_cast<const float4*>(col), w4 = *reinterpret_cast<const float4*>(col + colw);
            ga[0] = u.x; ga[1] = u.y; ga[2] = u.z; ga[3] = u.w;
            ge[0] = w4.x; ge[1] = w4.y; ge[2] = w4.z; ge[3] = w4.w;
          }
        }
        vf<V> y;
#pragma unroll
        for (int e = 0; e < V; ++e) y[e] = fmaf(fmaf(sL[e], cl[e], hL[e]), ga[e], ge[e]);
        if (a.round_out) vround_tf32<V>(y);
        astore<V>(outp + (t + r) * a.C, y);
      }
    }
  }
}

template <int LKIND, int GKIND, int V, int GC>
__global__ void __launch_bounds__(256, GC <= 0 ? 1 : 2) la_combine_kernel(LaArgs a, int rows_per_cta, int gspan) {
  grid_dep_wait();
  extern __shared__ __align__(16) float la_smem[];
  constexpr int R = 8;
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int Ll = a.loc.L, Lg = a.glo.L;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, Ll);
  LaSmem sm;
  sm.scratch = la_smem;
  sm.jc = reinterpret_cast<int*>(la_smem + (GC > 0 ? GC : 0) * 2 * blockDim.x * V);
  sm.jl = sm.jc + rows_per_cta;
  sm.jg = sm.jl + rows_per_cta + 4;
  // first global row any output row of this CTA can touch (halo included)
  const int g_first = nearest_src(t0, a.scale, Lg) - 2;
  fill_nearest(sm.jc, rows_per_cta, t0, Ll, a.scale, Lg);
  if constexpr (LKIND == SRC_INJECT_GATE || LKIND == SRC_INJECT_ADD)
    fill_nearest(sm.jl, rows_per_cta + 4, t0 - 2, Ll, a.loc.gscale, a.loc.Lg);
  if constexpr (GKIND == SRC_INJECT_GATE || GKIND == SRC_INJECT_ADD)
    fill_nearest(sm.jg, gspan, g_first, Lg, a.glo.gscale, a.glo.Lg);
  __syncthreads();
  if (ch >= a.C) return;
  const int g_last = nearest_src(t1 - 1, a.scale, Lg) + 2 + (GC > 0 ? GC : (GC < 0 ? 2 : 0));
  const bool interior = t0 - 2 >= 0 && t1 + 2 <= Ll && (t1 - t0) % R == 0 && g_first >= 0 && g_last < Lg;
  if (interior) la_body<LKIND, GKIND, V, GC, false>(a, b, ch, t0, t1, sm, g_first);
  else la_body<LKIND, GKIND, V, GC, true>(a, b, ch, t0, t1, sm, g_first);
}

// ----------------------------------------------------------------------------- injection on staged rows
// x_fused[k][t] recomputed from a raw spp_dw[k] row and the row of the global feature it sees:
//   BEST : (al*x + bl) * sigmoid(aa*g + ba) + (ae*g + be)   (loc_glo_fus, closed-form GlobLNs)
//   FORK : GlobLN(x) + g
// The gate / offset are cached while consecutive rows map to the same global row.
template <int KIND, int V>
struct Injector {
  vf<V> al, bl, aa, ba, ae, be, sg, eg;
  int cur;
  __device__ __forceinline__ void init(const SrcDesc& d, int b, int ch, int C) {
    cur = -1;
    if constexpr (KIND == SRC_INJECT_GATE) {
      const float* cf = d.coef + (size_t)b * 6 * C + ch;
      al = vload<V>(cf); bl = vload<V>(cf + C); aa = vload<V>(cf + 2 * C); ba = vload<V>(cf + 3 * C);
      ae = vload<V>(cf + 4 * C); be = vload<V>(cf + 5 * C);
    } else {
      norm_coef<V>(d.norm, b, ch, al, bl);
    }
  }
  __device__ __forceinline__ vf<V> apply(vf<V> raw, const vf<V>& grow, int j) {
    if (j != cur) {
      cur = j;
      if constexpr (KIND == SRC_INJECT_GATE) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          sg[e] = sigmoidf_(fmaf(aa[e], grow[e], ba[e]));
          eg[e] = fmaf(ae[e], grow[e], be[e]);
        }
      } else {
        eg = grow;
      }
    }
    if constexpr (KIND == SRC_INJECT_GATE) return vfma<V>(vfma<V>(raw, al, bl), sg, eg);
    else return vadd<V>(vfma<V>(raw, al, bl), eg);
  }
};

// ----------------------------------------------------------------------------- LA combine, streaming
// The same arithmetic as la_combine_kernel for the up-sampling steps (Lg <= Ll/2-ish, plain global
// operand), restructured for memory-level parallelism that does not depend on occupancy: a thread
// owns 4 channels and prefetches the rows of the NEXT chunk (8 local rows, 9 global rows, 3 rows of
// the injected global feature) with cp.async into a two-stage ring of thread-private shared-memory
// columns while it computes the current chunk out of the other stage.  No block-level
// synchronisation after the index tables are built.
constexpr int SR = 8;             // output rows per chunk
constexpr int SGC = 5;            // distinct global centres per chunk (ratio >= 2)
constexpr int SGR = SGC + 4;      // global rows per chunk (centres + halo)
constexpr int SGG = 3;            // rows of the injected feature per chunk (ratio to it >= 4)
constexpr int SAROWS = SR + SGR;  // ACT_T rows per ring stage; SGG fp32 rows follow them

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ vf<4> lds4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  vf<4> r;
  r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
  return r;
}

// the same for V channels per thread: fp32 rows of the global feature (16 / 8 bytes), activation rows (fp32: 16 / 8
// bytes, bf16: 8 / 4 bytes)
template <int V>
__device__ __forceinline__ void cp_async_f32(float* dst, const float* src, bool valid) {
  if constexpr (V == 4) {
    cp_async16(dst, src, valid);
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(valid ? 8 : 0)
                 : "memory");
  }
}
template <int V>
__device__ __forceinline__ void cp_async_actv(ACT_T* dst, const ACT_T* src, bool valid) {
  if constexpr (V == 4) {
    cp_async_act(dst, src, valid);
  } else if constexpr (sizeof(ACT_T) == 4) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(valid ? 8 : 0)
                 : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(valid ? 4 : 0)
                 : "memory");
  }
}
template <int V>
__device__ __forceinline__ vf<V> ldsv(const float* p) {
  if constexpr (V == 4) {
    return lds4(p);
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    vf<V> r;
    r[0] = t.x; r[1] = t.y;
    return r;
  }
}

// ring stage s of this thread: ACT_T rows, then fp32 rows of the injected feature
struct StageCol {
  ACT_T* act;
  float* g;
};
template <int AROWS, int GROWS, int V = 4>
__device__ __forceinline__ StageCol stage_col(void* ring, int s, int colw) {
  // sizes in floats (an ACT_T row of colw elements is colw*sizeof(ACT_T)/4 floats; colw is a multiple of 4)
  const int act_f = AROWS * colw * (int)sizeof(ACT_T) / 4;
  const int stage_f = act_f + GROWS * colw;
  float* base = static_cast<float*>(ring) + s * stage_f;
  return StageCol{reinterpret_cast<ACT_T*>(base) + threadIdx.x * V, base + act_f + threadIdx.x * V};
}
template <int AROWS, int GROWS>
static size_t ring_bytes(int threads, int V = 4) {
  return 2 * ((size_t)AROWS * threads * V * sizeof(ACT_T) + (size_t)GROWS * threads * V * sizeof(float));
}

// CT: in_channels when it is the compile-time 512 of every reference configuration (row strides and
// shared-memory offsets then fold into immediates), 0 = run-time
// BULK (interior CTAs of the 512-channel, 4-channels-per-thread form): stages filled by cp.async.bulk, see above
// NT: threads per CTA of the CT != 0 forms (128; 256 for the two-channels-per-thread bulk form, whose CTA of 8 warps
// still covers all 512 channels)
template <int LKIND, bool EDGE, int CT, int V, bool BULK = false, int NT = 128>
__device__ __forceinline__ void la_stream_body(const LaArgs& a, int b, int ch, int t0, int t1, void* ring,
                                               float* scratch, const int* jc, const int* jl, SbRing rb = SbRing{}) {
  static_assert(!BULK || (!EDGE && CT == 512 && NT * V == 512), "bulk staging: interior CTAs covering all 512 channels");
  const int Ll = a.loc.L, Lg = a.glo.L, Lgg = a.loc.Lg;
  const int C = CT ? CT : a.C;
  // CT != 0: NT threads per CTA (the launcher guarantees it), so the column pitch is a compile-time constant too
  const int colw = CT ? NT * V : blockDim.x * V;
  const ACT_T* xl = reinterpret_cast<const ACT_T*>(a.loc.x) + (size_t)b * Ll * C + ch;
  const ACT_T* xg = reinterpret_cast<const ACT_T*>(a.glo.x) + (size_t)b * Lg * C + ch;
  const float* gg = a.loc.g + (size_t)b * Lgg * C + ch;
  ACT_T* outp = reinterpret_cast<ACT_T*>(a.out) + (size_t)b * Ll * C + ch;
  float* mine = scratch + threadIdx.x * V;

  // per-channel constants
  Injector<LKIND, V> inj;
  inj.init(a.loc, b, ch, C);
  auto inject = [&](vf<V> raw, const vf<V>& grow, int j) { return inj.apply(raw, grow, j); };
  vf<V> wl[5], wa[5], we[5];
  load_taps<V>(a.wl, ch, wl);
  load_taps<V>(a.wa, ch, wa);
  load_taps<V>(a.we, ch, we);
  vf<V> sL, hL, sA, hA, sE, hE;
  norm_coef<V>(a.nL, b, ch, sL, hL);
  norm_coef<V>(a.nA, b, ch, sA, hA);
  norm_coef<V>(a.nE, b, ch, sE, hE);

  auto issue = [&](int k) {
    const StageCol st = stage_col<SAROWS, SGG, V>(ring, k & 1, colw);
    const int t = t0 + k * SR;
    if constexpr (BULK) {
      // thread 0 (channel 0: its pointers are the row bases) copies the three row ranges of the chunk
      if (threadIdx.x == 0) {
        uint64_t* full = rb.full;
        const int s = k & 1;
        rb.acquire(k);
        const int jlo = jc[t - t0], j0 = jl[t - t0 + 4];
        const int ngg = min(SGG, Lgg - j0);
        constexpr uint32_t rowb = 512 * sizeof(ACT_T);
        sb_expect_tx(full + s, (SR + SGR) * rowb + ngg * 2048u + (k == 0 ? 4 * rowb : 0u));
        sb_bulk_g2s(st.act, xl + (size_t)(t + 2) * C, SR * rowb, full + s);
        sb_bulk_g2s(st.act + SR * colw, xg + (size_t)(jlo - 2) * C, SGR * rowb, full + s);
        sb_bulk_g2s(st.g, gg + (size_t)j0 * C, ngg * 2048u, full + s);
        if (k == 0)  // rows t0-2 .. t0+1, parked in the other stage
          sb_bulk_g2s(stage_col<SAROWS, SGG, V>(ring, 1, colw).act, xl + (size_t)(t0 - 2) * C, 4 * rowb, full);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      const int row = t + 2 + i;
      const bool ok = !EDGE || row < Ll;
      cp_async_actv<V>(st.act + i * colw, xl + (ok ? row : 0) * C, ok);
    }
    const int jlo = jc[t - t0];
#pragma unroll
    for (int i = 0; i < SGR; ++i) {
      const int row = jlo - 2 + i;
      const bool ok = !EDGE || (row >= 0 && row < Lg);
      cp_async_actv<V>(st.act + (SR + i) * colw, xg + (ok ? row : 0) * C, ok);
    }
    const int j0 = jl[t - t0 + 4];
#pragma unroll
    for (int i = 0; i < SGG; ++i) {
      const int row = j0 + i;
      const bool ok = row < Lgg;
      cp_async_f32<V>(st.g + i * colw, gg + (ok ? row : 0) * C, ok);
    }
  };

  // rows t0-2 .. t0+1 of the local operand ride in chunk 0's copy group (parked in the free stage) and
  // their rows of the injected feature are fetched alongside, so a CTA exposes one memory latency
  vf<V> xr[SR + 4];
  const int nchunks = (t1 - t0 + SR - 1) / SR;
  vf<V> gpre[4];
  {
    const StageCol pre = stage_col<SAROWS, SGG, V>(ring, 1, colw);  // unused until chunk 1 is issued
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 - 2 + i;
      const bool ok = !EDGE || (t >= 0 && t < Ll);
      if constexpr (!BULK) cp_async_actv<V>(pre.act + i * colw, xl + (ok ? t : 0) * C, ok);
      gpre[i] = vload<V>(gg + jl[i] * C);
    }
  }
  issue(0);
  if constexpr (BULK) {
    rb.wait_full(0);
  } else {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  {
    const StageCol pre = stage_col<SAROWS, SGG, V>(ring, 1, colw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 - 2 + i;
      if (EDGE && (t < 0 || t >= Ll)) xr[SR + i] = vzero<V>();
      else xr[SR + i] = inject(alds<V>(pre.act + i * colw), gpre[i], jl[i]);
    }
    if constexpr (BULK) rb.release(1);  // the parked rows are consumed: stage 1 may be filled
  }
  for (int k = 0; k < nchunks; ++k) {
    if (k + 1 < nchunks) issue(k + 1);
    if constexpr (BULK) {
      if (k > 0) rb.wait_full(k);   // chunk 0 was awaited with the prologue rows
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    const StageCol st = stage_col<SAROWS, SGG, V>(ring, k & 1, colw);
    const int t = t0 + k * SR;
    const int jlo = jc[t - t0];
    {
      // global branch: one evaluation per distinct centre of the chunk, parked in the scratch column
      const int tl = (EDGE ? min(t + SR, t1) : t + SR) - 1;
      const int nc = jc[tl - t0] - jlo + 1;
      vf<V> gr[SGR];
#pragma unroll
      for (int i = 0; i < SGR; ++i) gr[i] = alds<V>(st.act + (SR + i) * colw);
#pragma unroll
      for (int i = 0; i < SGC; ++i) {
        if (i < nc) {
          vf<V> ca = vfma<V>(sA, conv5<V>(wa, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]), hA);
          const vf<V> ce = vfma<V>(sE, conv5<V>(we, gr[i], gr[i + 1], gr[i + 2], gr[i + 3], gr[i + 4]), hE);
#pragma unroll
          for (int e = 0; e < V; ++e) ca[e] = sigmoidf_(ca[e]);
          vstore<V>(mine + (2 * i) * colw, ca);
          vstore<V>(mine + (2 * i + 1) * colw, ce);
        }
      }
    }
    // local operand: injection recomputed from the staged raw rows
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = xr[SR + i];
    const int j0 = jl[t - t0 + 4];
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      if (EDGE && t + 2 + i >= Ll) {
        xr[4 + i] = vzero<V>();
      } else {
        const int j = jl[t - t0 + 4 + i];
        xr[4 + i] = inject(alds<V>(st.act + i * colw), ldsv<V>(st.g + (j - j0) * colw), j);
      }
    }
    if constexpr (BULK) rb.release(k & 1);  // last read of this stage by this warp
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      if (!EDGE || t + r < t1) {
        const vf<V> cl = conv5<V>(wl, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        const int j = jc[t + r - t0];
        const vf<V> ga = ldsv<V>(mine + (2 * (j - jlo)) * colw), ge = ldsv<V>(mine + (2 * (j - jlo) + 1) * colw);
        vf<V> y = vfma<V>(vfma<V>(sL, cl, hL), ga, ge);
        if (a.round_out) vround_tf32<V>(y);
        astore<V>(outp + (t + r) * C, y);
      }
    }
  }
}

// V = 4: 245 registers, two CTAs of 128 threads per SM (8 warps: every scheduler has two warps to cover the dependent-issue
// stalls of this arithmetic-heavy loop - issue slots were 38 % busy at 0.65 of the HBM peak).  V = 2: half the
// per-thread state, four CTAs per SM (16 warps), a CTA covers 256 channels.
// BULK, V = 2: 256 threads (8 warps) per CTA, two CTAs per SM at <= 128 registers: the 16 warps per SM of the
// two-channel form without its doubled copy / address overhead (one thread stages the rows for the whole CTA).
// Measured: the same duration as V = 4 (143 vs 139 us for the 661 MB launch: 38 % more instructions at 47 % instead of
// 34 % issue-slot use, profiles/r03_ncu_la_stream_bulk.txt), so V = 4 stays the default.
template <int LKIND, int CT, int V, bool BULK = false>
#ifndef TD_LASTREAM_MINB
#define TD_LASTREAM_MINB 2
#endif
__global__ void __launch_bounds__(BULK && V == 2 ? 256 : 128, V == 4 ? TD_LASTREAM_MINB : (BULK ? 2 : 4))
la_stream_kernel(LaArgs a, int rows_per_cta) {
  constexpr int NT = BULK && V == 2 ? 256 : 128;
  extern __shared__ __align__(16) float la_smem[];
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int Ll = a.loc.L, Lg = a.glo.L;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(t0 + rows_per_cta, Ll);
  const int colw = CT ? NT * V : blockDim.x * V;
  // [ring: 2 stages][scratch 2*SGC rows fp32][tables][BULK: full[2], empty[2] mbarriers]; plain pointer arithmetic on
  // the __shared__ array so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
  void* ring = la_smem;
  float* scratch = la_smem + 2 * ((SAROWS * colw * sizeof(ACT_T)) / sizeof(float) + SGG * colw);
  int* jc = reinterpret_cast<int*>(scratch + 2 * SGC * colw);
  int* jl = jc + rows_per_cta;
  uint64_t* bars = reinterpret_cast<uint64_t*>(jl + rows_per_cta + 4);  // rows_per_cta is a multiple of 8: 16-byte aligned
  fill_nearest(jc, rows_per_cta, t0, Ll, a.scale, Lg);
  fill_nearest(jl, rows_per_cta + 4, t0 - 2, Ll, a.loc.gscale, a.loc.Lg);
  SbRing rb{};
  if constexpr (BULK) rb = sb_ring_init(bars, NT / 32);
  __syncthreads();
  grid_dep_wait();   // everything above is independent of the previous kernel's output
  if (ch >= a.C) return;
  const int g_first = nearest_src(t0, a.scale, Lg) - 2;
  const int g_last = nearest_src(t1 - 1, a.scale, Lg) + SGR;
  const bool interior = t0 - 2 >= 0 && t1 + 2 <= Ll && (t1 - t0) % SR == 0 && g_first >= 0 && g_last < Lg;
  if constexpr (BULK) {
    if (interior) la_stream_body<LKIND, false, CT, V, true, NT>(a, b, ch, t0, t1, ring, scratch, jc, jl, rb);
    else la_stream_body<LKIND, true, CT, V, false, NT>(a, b, ch, t0, t1, ring, scratch, jc, jl);
  } else {
    if (interior) la_stream_body<LKIND, false, CT, V>(a, b, ch, t0, t1, ring, scratch, jc, jl);
    else la_stream_body<LKIND, true, CT, V>(a, b, ch, t0, t1, ring, scratch, jc, jl);
  }
}

// the streaming kernel applies when 8 output rows see <= 5 global centres and <= 3 rows of the
// injected feature, and the global operand needs no transform
static bool la_stream_applies(const LaArgs& a) {
  return a.gkind == SRC_PLAIN && (a.lkind == SRC_INJECT_GATE || a.lkind == SRC_INJECT_ADD) && a.glo.L <= a.loc.L &&
         7.0 * a.glo.L / a.loc.L <= 3.99 && 7.0 * a.loc.Lg / a.loc.L <= 1.99 && a.C % 4 == 0;
}

template <int LKIND, int V>
static int launch_la_stream_v(const LaArgs& a, cudaStream_t st) {
  const bool bulk = (bulk_mask() & 1) != 0 && a.C == 512;
  int threads = a.C / V;
  if (threads > 128 && !(bulk && V == 2)) threads = 128;   // bulk, V = 2: one CTA of 256 threads covers the 512 channels
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / V, threads);
  int rows, tiles;
  static const long ls_target = getenv("TDANET_LASTREAM_TARGET") ? atol(getenv("TDANET_LASTREAM_TARGET")) : 148L * 2 * 4;
  static const int ls_cap = getenv("TDANET_LASTREAM_CAP") ? atoi(getenv("TDANET_LASTREAM_CAP")) : 128;
  pick_tiling(a.B, a.loc.L, ctiles, SR, &rows, &tiles, ls_target, ls_cap);
  // small launches (training batches, a 64-mixture job split over 8 GPUs): four waves' worth of CTAs would leave each
  // with 1-2 chunks after its prologue (index tables, 15 tap loads, the first copy's latency); such a launch gets up to
  // TDANET_LASTREAM_MINROWS rows per CTA as long as one wave of two CTAs per SM stays full.  Measured on B200, training
  // step at B = 8: 19.32 -> 18.93 ms; at B = 64 only the 503-row step changes (32 -> 56 rows: neutral).
  static const int ls_minrows = getenv("TDANET_LASTREAM_MINROWS") ? atoi(getenv("TDANET_LASTREAM_MINROWS")) : 56;
  if (ls_minrows > 0) {
    int rows1, tiles1;
    pick_tiling(a.B, a.loc.L, ctiles, SR, &rows1, &tiles1, 148L * 2, ls_cap);
    if (rows1 > ls_minrows) rows1 = ls_minrows;
    if (rows1 > rows) {
      rows = rows1;
      tiles = cdiv(a.loc.L, rows);
    }
  }
  dim3 grid(tiles, ctiles, a.B);
  const size_t smem = (size_t)2 * SGC * threads * V * sizeof(float) + (size_t)(2 * rows + 4) * sizeof(int) +
                      ring_bytes<SAROWS, SGG>(threads, V) + 4 * sizeof(uint64_t);
  static PerDeviceOnce first_use;
  if (first_use()) {
    TD_CUDA(cudaFuncSetAttribute(la_stream_kernel<LKIND, 0, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    TD_CUDA(cudaFuncSetAttribute(la_stream_kernel<LKIND, 512, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    TD_CUDA(cudaFuncSetAttribute(la_stream_kernel<LKIND, 512, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  }
  if (bulk && threads * V == 512) {
    TD_LAUNCH((la_stream_kernel<LKIND, 512, V, true>), grid, threads, smem, st, a, rows);
  } else if (a.C == 512 && threads == 128) {
    TD_LAUNCH((la_stream_kernel<LKIND, 512, V>), grid, threads, smem, st, a, rows);
  } else {
    TD_LAUNCH((la_stream_kernel<LKIND, 0, V>), grid, threads, smem, st, a, rows);
  }
  return 0;
}

template <int LKIND>
static int launch_la_stream(const LaArgs& a, cudaStream_t st) {
  // channels per thread: 4 (two CTAs of 128 threads per SM).  TDANET_LASTREAM_V=2 selects the two-channel form for A/B
  // measurements (bulk-staged: 256 threads per CTA; measured on B200: 17.60 instead of 17.39 ms per step)
  static const int v_knob = getenv("TDANET_LASTREAM_V") ? atoi(getenv("TDANET_LASTREAM_V")) : 4;
  if (v_knob == 2 && a.C % 2 == 0) return launch_la_stream_v<LKIND, 2>(a, st);
  return launch_la_stream_v<LKIND, 4>(a, st);
}

// ----------------------------------------------------------------------------- LA local statistics, one launch
// GlobLN statistics of local_embedding(x_fused[i]) for EVERY top-down step at once: they depend only on
// spp_dw[i] and the global feature, not on the top-down chain, so the four scales share one launch
// (the small scales ride along with the large one instead of paying their own launch and tail).
// The local operand is streamed (cp.async ring of thread-private columns), the injection recomputed;
// nothing is written but the sums.
constexpr int SSG = 5;            // rows of the injected feature per chunk (ratio to it >= 2)

struct LocalStatsArgs {
  int n;
  DwArgs step[TDANET_MAX_DEPTH];
  int tile_end[TDANET_MAX_DEPTH];  // exclusive prefix sum of tiles per step
  int rows;                        // rows per CTA (same for every step)
  int ns;                          // stages of the bulk ring
};

template <int LKIND, bool EDGE, int CT, bool BULK = false>
__device__ __forceinline__ void stats_stream_body(const DwArgs& a, int b, int ch, int t0, int t1, void* ring,
                                                  const int* jl, float& tot1, float& tot2, SbRingN rb = SbRingN{}) {
  static_assert(!BULK || (!EDGE && CT == 512), "bulk staging: interior CTAs covering all 512 channels");
  constexpr int V = 4;
  const int Ll = a.src.L, Lgg = a.src.Lg;
  const int C = CT ? CT : a.C;
  const int colw = CT ? CT : blockDim.x * V;
  const ACT_T* xl = reinterpret_cast<const ACT_T*>(a.src.x) + (size_t)b * Ll * C + ch;
  const float* gg = a.src.g + (size_t)b * Lgg * C + ch;
  Injector<LKIND, V> inj;
  inj.init(a.src, b, ch, C);
  vf<V> wl[5];
  load_taps<V>(a.w[0], ch, wl);
  vf<V> s1 = vzero<V>(), s2 = vzero<V>();

  SbCursor pc, cc;  // BULK: producer (thread 0) / consumer positions in the ring
  auto issue = [&](int k) {
    const int t = t0 + k * SR;
    if constexpr (BULK) {
      if (threadIdx.x == 0) {
        const StageCol st = stage_col<SR, SSG>(ring, pc.stage, colw);
        uint64_t* full = rb.full + pc.stage;
        rb.acquire<true>(pc);
        const int j0 = jl[t - t0 + 4];
        const int ngg = min(SSG, Lgg - j0);
        constexpr uint32_t rowb = 512 * sizeof(ACT_T);
        sb_expect_tx(full, SR * rowb + ngg * 2048u + (k == 0 ? 4 * rowb : 0u));
        sb_bulk_g2s(st.act, xl + (size_t)(t + 2) * C, SR * rowb, full);
        sb_bulk_g2s(st.g, gg + (size_t)j0 * C, ngg * 2048u, full);
        if (k == 0)  // rows t0-2 .. t0+1, parked in the last stage
          sb_bulk_g2s(stage_col<SR, SSG>(ring, rb.ns - 1, colw).act, xl + (size_t)(t0 - 2) * C, 4 * rowb, full);
        pc.next(rb.ns);
      }
      return;
    }
    const StageCol st = stage_col<SR, SSG>(ring, k & 1, colw);
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      const int row = t + 2 + i;
      const bool ok = !EDGE || row < Ll;
      cp_async_act(st.act + i * colw, xl + (ok ? row : 0) * C, ok);
    }
    const int j0 = jl[t - t0 + 4];
#pragma unroll
    for (int i = 0; i < SSG; ++i) {
      const int row = j0 + i;
      const bool ok = row < Lgg;
      cp_async16(st.g + i * colw, gg + (ok ? row : 0) * C, ok);
    }
  };

  vf<V> xr[SR + 4];
  const int nchunks = (t1 - t0 + SR - 1) / SR;
  vf<V> gpre[4];
  {
    const StageCol pre = stage_col<SR, SSG>(ring, BULK ? rb.ns - 1 : 1, colw);  // unused until that stage's first chunk
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 - 2 + i;
      const bool ok = !EDGE || (t >= 0 && t < Ll);
      if constexpr (!BULK) cp_async_act(pre.act + i * colw, xl + (ok ? t : 0) * C, ok);
      gpre[i] = vload<V>(gg + jl[i] * C);
    }
  }
  issue(0);
  if constexpr (BULK) {
    for (int k = 1; k < rb.ns - 1; ++k)
      if (k < nchunks) issue(k);
    rb.wait_full(cc);   // chunk 0 and the parked rows
  } else {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  {
    const StageCol pre = stage_col<SR, SSG>(ring, BULK ? rb.ns - 1 : 1, colw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 - 2 + i;
      if (EDGE && (t < 0 || t >= Ll)) xr[SR + i] = vzero<V>();
      else xr[SR + i] = inj.apply(alds<V>(pre.act + i * colw), gpre[i], jl[i]);
    }
    if constexpr (BULK) rb.release(rb.ns - 1);
  }
  for (int k = 0; k < nchunks; ++k) {
    if constexpr (BULK) {
      if (k + rb.ns - 1 < nchunks) issue(k + rb.ns - 1);
      rb.wait_full(cc);
    } else {
      if (k + 1 < nchunks) issue(k + 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    const StageCol st = stage_col<SR, SSG>(ring, BULK ? cc.stage : (k & 1), colw);
    const int t = t0 + k * SR;
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = xr[SR + i];
    const int j0 = jl[t - t0 + 4];
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      if (EDGE && t + 2 + i >= Ll) {
        xr[4 + i] = vzero<V>();
      } else {
        const int j = jl[t - t0 + 4 + i];
        xr[4 + i] = inj.apply(alds<V>(st.act + i * colw), lds4(st.g + (j - j0) * colw), j);
      }
    }
    if constexpr (BULK) {
      rb.release(cc.stage);
      cc.next(rb.ns);
    }
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      if (!EDGE || t + r < t1) {
        const vf<V> y = conv5<V>(wl, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        s1 = vadd<V>(s1, y);
        s2 = vfma<V>(y, y, s2);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    tot1 += s1[e];
    tot2 += s2[e];
  }
}

template <int LKIND, int CT, bool BULK = false>
__global__ void __launch_bounds__(128, 3) la_local_stats_kernel(LocalStatsArgs p) {
  extern __shared__ __align__(16) float la_smem[];
  __shared__ double red[64];
  constexpr int V = 4;
  int step = 0;
  while (step + 1 < p.n && (int)blockIdx.x >= p.tile_end[step]) ++step;
  const DwArgs& loc = p.step[step];
  const int tile = blockIdx.x - (step ? p.tile_end[step - 1] : 0);
  const int b = blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int Ll = loc.src.L;
  const int t0 = tile * p.rows, t1 = min(t0 + p.rows, Ll);
  void* ring = la_smem;
  int* jl = reinterpret_cast<int*>(la_smem + (BULK ? p.ns : 2) * ((SR * blockDim.x * V * sizeof(ACT_T)) / sizeof(float) + SSG * blockDim.x * V));
  fill_nearest(jl, p.rows + 4, t0 - 2, Ll, loc.src.gscale, loc.src.Lg);
  SbRingN rb{};
  if constexpr (BULK) rb = sb_ringn_init(reinterpret_cast<uint64_t*>(jl + p.rows + 4), p.ns, 4);  // rows: a multiple of 8
  __syncthreads();
  grid_dep_wait();   // the index table does not depend on the previous kernel
  float tot1[1] = {0.f}, tot2[1] = {0.f};
  if (ch < loc.C) {
    const bool interior = t0 - 2 >= 0 && t1 + 2 <= Ll && (t1 - t0) % SR == 0;
    if (interior) stats_stream_body<LKIND, false, CT, BULK>(loc, b, ch, t0, t1, ring, jl, tot1[0], tot2[0], rb);
    else stats_stream_body<LKIND, true, CT>(loc, b, ch, t0, t1, ring, jl, tot1[0], tot2[0]);
  }
  flush_item_stats<1>(loc.stats, b, tot1, tot2, red, loc.det);
}

template <int LKIND>
static int launch_la_local_stats_t(LocalStatsArgs& p, cudaStream_t st) {
  const DwArgs& a0 = p.step[0];
  int threads = a0.C / 4;
  if (threads > 128) threads = 128;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a0.C / 4, threads);
  long total_rows = 0;
  for (int i = 0; i < p.n; ++i) total_rows += p.step[i].Lout;
  int rows, tiles;
  static const long ls_target = getenv("TDANET_LSTATS_TARGET") ? atol(getenv("TDANET_LSTATS_TARGET")) : 148L * 4;
  static const int ls_cap = getenv("TDANET_LSTATS_CAP") ? atoi(getenv("TDANET_LSTATS_CAP")) : 128;
  pick_tiling(a0.B, (int)total_rows, ctiles, SR, &rows, &tiles, ls_target, ls_cap);
  p.rows = rows;
  int acc = 0;
  for (int i = 0; i < p.n; ++i) {
    acc += cdiv(p.step[i].Lout, rows);
    p.tile_end[i] = acc;
  }
  dim3 grid(acc, ctiles, a0.B);
  const bool bulk = a0.C == 512 && threads == 128 && (bulk_mask() & 2);
  // three CTAs of two stages or two CTAs of up to four stages per SM: the bytes in flight decide (see SbRingN)
  p.ns = bulk ? bulk_stages() : 2;
  const size_t smem = ring_bytes<SR, SSG>(threads) / 2 * p.ns + (size_t)(rows + 4) * sizeof(int) + SB_BAR_BYTES;
  static PerDeviceOnce first_use;
  if (first_use()) {
    TD_CUDA(cudaFuncSetAttribute(la_local_stats_kernel<LKIND, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    TD_CUDA(cudaFuncSetAttribute(la_local_stats_kernel<LKIND, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    TD_CUDA(cudaFuncSetAttribute(la_local_stats_kernel<LKIND, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  }
  if (bulk) {
    TD_LAUNCH((la_local_stats_kernel<LKIND, 512, true>), grid, threads, smem, st, p);
  } else if (a0.C == 512 && threads == 128) {
    TD_LAUNCH((la_local_stats_kernel<LKIND, 512>), grid, threads, smem, st, p);
  } else {
    TD_LAUNCH((la_local_stats_kernel<LKIND, 0>), grid, threads, smem, st, p);
  }
  return 0;
}

int launch_la_local_stats(const DwArgs* steps, int n, cudaStream_t st) {
  TD_REQUIRE(n >= 1 && n <= TDANET_MAX_DEPTH, "la_local_stats: %d steps", n);
  bool stream_ok = true;
  for (int i = 0; i < n; ++i) {
    const DwArgs& a = steps[i];
    TD_REQUIRE(a.nw == 1 && a.stats && !a.out && a.stride == 1, "la_local_stats: bad arguments");
    stream_ok = stream_ok && a.kind == steps[0].kind && (a.kind == SRC_INJECT_GATE || a.kind == SRC_INJECT_ADD) &&
                7.0 * a.src.Lg / a.src.L <= 3.99 && a.C % 4 == 0 && a.C == steps[0].C && a.B == steps[0].B &&
                (long)a.src.L * a.C < (1L << 31);
  }
  if (stream_ok) {
    LocalStatsArgs p{};
    p.n = n;
    for (int i = 0; i < n; ++i) p.step[i] = steps[i];
    if (steps[0].kind == SRC_INJECT_GATE) return launch_la_local_stats_t<SRC_INJECT_GATE>(p, st);
    return launch_la_local_stats_t<SRC_INJECT_ADD>(p, st);
  }
  for (int i = 0; i < n; ++i)
    if (int e = TD_ACT_NS::launch_dw5(steps[i], st)) return e;
  return 0;
}

// ----------------------------------------------------------------------------- LA global statistics, streaming
// GlobLN statistics of global_act(x_g) and global_embedding(x_g) (TDANet_best.py:286-289) for a plain x_g
// (the tensor the previous top-down step just wrote): nothing is written but four sums per item, so the
// kernel is a pure read.  Same structure as the local-statistics kernel: a thread owns 4 channels and
// streams its column through a two-stage cp.async ring, few long CTAs (the tap loads and the 4-row
// prologue are paid once per CTA), four CTAs per SM.
template <bool EDGE, int CT, bool BULK = false>
__device__ __forceinline__ void gstats_stream_body(const DwArgs& a, int b, int ch, int t0, int t1, void* ring,
                                                   float (&tot1)[2], float (&tot2)[2], SbRingN rb = SbRingN{}) {
  static_assert(!BULK || (!EDGE && CT == 512), "bulk staging: interior CTAs covering all 512 channels");
  constexpr int V = 4;
  const int L = a.src.L;
  const int C = CT ? CT : a.C;
  const int colw = CT ? CT : blockDim.x * V;
  const ACT_T* x = reinterpret_cast<const ACT_T*>(a.src.x) + (size_t)b * L * C + ch;
  vf<V> wa[5], we[5];
  load_taps<V>(a.w[0], ch, wa);
  load_taps<V>(a.w[1], ch, we);
  vf<V> s1a = vzero<V>(), s2a = vzero<V>(), s1e = vzero<V>(), s2e = vzero<V>();

  SbCursor pc, cc;  // BULK: producer (thread 0) / consumer positions in the ring
  auto issue = [&](int k) {
    const int t = t0 + k * SR;
    if constexpr (BULK) {
      if (threadIdx.x == 0) {
        const StageCol st = stage_col<SR, 0>(ring, pc.stage, colw);
        uint64_t* full = rb.full + pc.stage;
        rb.acquire<true>(pc);
        constexpr uint32_t rowb = 512 * sizeof(ACT_T);
        sb_expect_tx(full, SR * rowb + (k == 0 ? 4 * rowb : 0u));
        sb_bulk_g2s(st.act, x + (size_t)(t + 2) * C, SR * rowb, full);
        if (k == 0)  // rows t0-2 .. t0+1, parked in the last stage
          sb_bulk_g2s(stage_col<SR, 0>(ring, rb.ns - 1, colw).act, x + (size_t)(t0 - 2) * C, 4 * rowb, full);
        pc.next(rb.ns);
      }
      return;
    }
    const StageCol st = stage_col<SR, 0>(ring, k & 1, colw);
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      const int row = t + 2 + i;
      const bool ok = !EDGE || row < L;
      cp_async_act(st.act + i * colw, x + (ok ? row : 0) * C, ok);
    }
  };

  vf<V> xr[SR + 4];
  const int nchunks = (t1 - t0 + SR - 1) / SR;
  {
    const StageCol pre = stage_col<SR, 0>(ring, BULK ? rb.ns - 1 : 1, colw);  // unused until that stage's first chunk
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 - 2 + i;
      const bool ok = !EDGE || (t >= 0 && t < L);
      if constexpr (!BULK) cp_async_act(pre.act + i * colw, x + (ok ? t : 0) * C, ok);
    }
  }
  issue(0);
  if constexpr (BULK) {
    for (int k = 1; k < rb.ns - 1; ++k)
      if (k < nchunks) issue(k);
    rb.wait_full(cc);   // chunk 0 and the parked rows
  } else {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  {
    const StageCol pre = stage_col<SR, 0>(ring, BULK ? rb.ns - 1 : 1, colw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 - 2 + i;
      if (EDGE && (t < 0 || t >= L)) xr[SR + i] = vzero<V>();
      else xr[SR + i] = alds<V>(pre.act + i * colw);
    }
    if constexpr (BULK) rb.release(rb.ns - 1);
  }
  for (int k = 0; k < nchunks; ++k) {
    if constexpr (BULK) {
      if (k + rb.ns - 1 < nchunks) issue(k + rb.ns - 1);
      rb.wait_full(cc);
    } else {
      if (k + 1 < nchunks) issue(k + 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    const StageCol st = stage_col<SR, 0>(ring, BULK ? cc.stage : (k & 1), colw);
    const int t = t0 + k * SR;
#pragma unroll
    for (int i = 0; i < 4; ++i) xr[i] = xr[SR + i];
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      if (EDGE && t + 2 + i >= L) xr[4 + i] = vzero<V>();
      else xr[4 + i] = alds<V>(st.act + i * colw);
    }
    if constexpr (BULK) {
      rb.release(cc.stage);
      cc.next(rb.ns);
    }
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      if (!EDGE || t + r < t1) {
        const vf<V> ya = conv5<V>(wa, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        const vf<V> ye = conv5<V>(we, xr[r], xr[r + 1], xr[r + 2], xr[r + 3], xr[r + 4]);
        s1a = vadd<V>(s1a, ya);
        s2a = vfma<V>(ya, ya, s2a);
        s1e = vadd<V>(s1e, ye);
        s2e = vfma<V>(ye, ye, s2e);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    tot1[0] += s1a[e];
    tot2[0] += s2a[e];
    tot1[1] += s1e[e];
    tot2[1] += s2e[e];
  }
}

template <int CT, bool BULK = false>
__global__ void __launch_bounds__(128, 4) gstats_stream_kernel(DwArgs a, int rows_per_cta, int ns) {
  extern __shared__ __align__(16) float la_smem[];
  SbRingN rb{};
  if constexpr (BULK) {  // the barriers follow the ring (ns stages of SR rows of 512 channels)
    rb = sb_ringn_init(reinterpret_cast<uint64_t*>(la_smem + ns * (SR * 512 * sizeof(ACT_T)) / sizeof(float)), ns, 4);
    __syncthreads();
  }
  grid_dep_wait();
  __shared__ double red[64];
  constexpr int V = 4;
  const int b = a.rev ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
  const int ch = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const int L = a.src.L;
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(t0 + rows_per_cta, L);
  float tot1[2] = {0.f, 0.f}, tot2[2] = {0.f, 0.f};
  if (ch < a.C) {
    const bool interior = t0 - 2 >= 0 && t1 + 2 <= L && (t1 - t0) % SR == 0;
    if (interior) gstats_stream_body<false, CT, BULK>(a, b, ch, t0, t1, la_smem, tot1, tot2, rb);
    else gstats_stream_body<true, CT>(a, b, ch, t0, t1, la_smem, tot1, tot2);
  }
  flush_item_stats<2>(a.stats, b, tot1, tot2, red, a.det);
}

static bool gstats_stream_applies(const DwArgs& a) {
  static const bool off = getenv("TDANET_GSTATS_STREAM") && atoi(getenv("TDANET_GSTATS_STREAM")) == 0;
  return !off && a.kind == SRC_PLAIN && a.nw == 2 && !a.out && a.stats && !a.chstats && !a.pool_out && a.stride == 1 &&
         !a.bias[0] && !a.bias[1] && a.Lout == a.src.L && a.C % 4 == 0 && (long)a.src.L * a.C < (1L << 31);
}

static int launch_gstats_stream(const DwArgs& a, cudaStream_t st) {
  int threads = a.C / 4;
  if (threads > 128) threads = 128;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / 4, threads);
  // one wave of four CTAs per SM (tuning aid: TDANET_GSTATS_TARGET / TDANET_GSTATS_CAP)
  static const long target = getenv("TDANET_GSTATS_TARGET") ? atol(getenv("TDANET_GSTATS_TARGET")) : 148L * 4;
  static const int cap = getenv("TDANET_GSTATS_CAP") ? atoi(getenv("TDANET_GSTATS_CAP")) : 128;
  int rows, tiles;
  pick_tiling(a.B, a.src.L, ctiles, SR, &rows, &tiles, target, cap);
  dim3 grid(tiles, ctiles, a.B);
  const bool bulk = a.C == 512 && threads == 128 && (bulk_mask() & 4);
  const int ns = bulk ? bulk_stages() : 2;
  const size_t smem = ring_bytes<SR, 0>(threads) / 2 * ns + SB_BAR_BYTES;
  static PerDeviceOnce first_use;
  if (first_use()) {
    TD_CUDA(cudaFuncSetAttribute(gstats_stream_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  if (bulk) {
    TD_LAUNCH((gstats_stream_kernel<512, true>), grid, threads, smem, st, a, rows, ns);
  } else if (a.C == 512 && threads == 128) {
    TD_LAUNCH((gstats_stream_kernel<512>), grid, threads, smem, st, a, rows, 2);
  } else {
    TD_LAUNCH((gstats_stream_kernel<0>), grid, threads, smem, st, a, rows, 2);
  }
  return 0;
}

template <int LKIND, int GKIND>
static int launch_la_t(const LaArgs& a, cudaStream_t st) {
  constexpr int V = 2;
  int threads = a.C / V;
  if (threads > 256) threads = 256;
  // the down-sampling forms hold ~150 registers: CTAs of 128 threads fit three to an SM instead of one of 256
  if (a.glo.L > a.loc.L && threads > 128) threads = 128;
  if (threads < 32) threads = 32;
  const int ctiles = cdiv(a.C / V, threads);
  int rows, tiles;
  static const long lt_target = getenv("TDANET_LAT_TARGET") ? atol(getenv("TDANET_LAT_TARGET")) : 148L * 16;
  static const int lt_cap = getenv("TDANET_LAT_CAP") ? atoi(getenv("TDANET_LAT_CAP")) : 64;
  pick_tiling(a.B, a.loc.L, ctiles, 8, &rows, &tiles, lt_target, lt_cap);
  dim3 grid(tiles, ctiles, a.B);
  // rows of the global tensor one CTA can touch: its rows map to <= rows*scale + 1 centres, + halo
  const int gspan = (int)((double)rows * a.glo.L / a.loc.L) + 16;
  const size_t tabs = (size_t)(2 * rows + 4 + gspan) * sizeof(int);
  if (a.glo.L > a.loc.L && a.scale >= 1.86f && a.scale <= 2.0f) {
    // the first top-down step of every reference configuration (Lg = 2*Ll or 2*Ll - 1): with s = Lg/Ll in
    // (2 - 1/7, 2], floor((t+r)*s) - floor(t*s) is 2r or 2r-1 for r < 8, so a chunk's windows sit at static offsets
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, -2>), grid, threads, tabs, st, a, rows, gspan);
  } else if (a.glo.L > a.loc.L) {
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, 0>), grid, threads, tabs, st, a, rows, gspan);
  } else if (7.0 * a.glo.L / a.loc.L <= 3.99) {
    // ratio >= ~2 (every up-sampling step of the U-Net): 8 output rows see at most 5 centres
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, 5>), grid, threads, tabs + (size_t)5 * 2 * threads * V * sizeof(float), st, a, rows, gspan);
  } else {
    TD_LAUNCH((la_combine_kernel<LKIND, GKIND, V, 8>), grid, threads, tabs + (size_t)8 * 2 * threads * V * sizeof(float), st, a, rows, gspan);
  }
  return 0;
}

int launch_la_combine(const LaArgs& a, cudaStream_t st) {
  TD_REQUIRE(a.C % 4 == 0, "la: C=%d must be a multiple of 4", a.C);
  TD_REQUIRE((long)a.loc.L * a.C < (1L << 31) && (long)a.glo.L * a.C < (1L << 31), "la: item too large for 32-bit offsets");
  if (la_stream_applies(a)) {
    if (a.lkind == SRC_INJECT_GATE) return launch_la_stream<SRC_INJECT_GATE>(a, st);
    return launch_la_stream<SRC_INJECT_ADD>(a, st);
  }
  if (a.lkind == SRC_INJECT_GATE && a.gkind == SRC_INJECT_GATE) return launch_la_t<SRC_INJECT_GATE, SRC_INJECT_GATE>(a, st);
  if (a.lkind == SRC_INJECT_GATE && a.gkind == SRC_PLAIN) return launch_la_t<SRC_INJECT_GATE, SRC_PLAIN>(a, st);
  if (a.lkind == SRC_INJECT_ADD && a.gkind == SRC_INJECT_ADD) return launch_la_t<SRC_INJECT_ADD, SRC_INJECT_ADD>(a, st);
  if (a.lkind == SRC_INJECT_ADD && a.gkind == SRC_PLAIN) return launch_la_t<SRC_INJECT_ADD, SRC_PLAIN>(a, st);
  if (a.lkind == SRC_PLAIN && a.gkind == SRC_PLAIN) return launch_la_t<SRC_PLAIN, SRC_PLAIN>(a, st);
  return fail(TDANET_EINVAL, "la: unsupported source kinds %d/%d", a.lkind, a.gkind);
}

}  // namespace TD_ACT_NS
}  // namespace td
