// tcgen05 / TMEM / TMA GEMM (kind::tf32) for the 1x1 convolutions and attention / FFN projections:
//   D[b, r, n] = sum_k A[b, r, k] * W[n, k]  (+ epilogue shared with gemm_simt.cu)
//
// Both operands are K-major (activations are channels-last, Conv1d / Linear weights are [N, K]), so
// TMA loads 128-byte-swizzled [rows x 32 fp32] boxes that tcgen05.mma consumes in place; the tensor
// core reads fp32 bits as TF32.  Persistent, warp-specialised CTA (one per SM):
//   warp 0      TMA producer        (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time (thread = row), transposed
//               through a per-warp shared-memory patch so that 8 lanes hold the 32 columns of one
//               row, then bias / residual / statistics and coalesced 128-bit st.global.  The
//               accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the loads
//               and MMAs of tile i+1.
// A-tiles are 128 rows of ONE batch item (3-D tensor map, rows past the item are zero-filled by TMA)
// so that the GlobLN statistics of the epilogue never straddle items.
//
// Accuracy modes: TDANET_GEMM_TF32  : W pre-rounded to TF32 (RN), one MMA pass.
//                 TDANET_GEMM_TF32X3: W used as is (tensor core truncates = W_hi) plus a second pass
//                                     over W_lo = W - trunc(W); A is taken as stored in both.
#include "kernels.h"
#include <cuda.h>

namespace td {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // fp32 elements = one 128-byte swizzle row
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (2 + TC_EPI_WARPS) * 32;
constexpr int TC_PATCH = 32 * 36;  // floats per epilogue warp: 32 rows x (32 + 4 pad) columns
constexpr int TC_TPATCH = 2 * 32 * 32;  // TMA-store epilogue: two 32 x 32 boxes (4 KB each, 1024-byte aligned) per warp

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must fault (and be reported by the next CUDA call), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 2 GHz
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMA store of a [32 rows x 32 fp32] box (128-byte swizzled in shared memory) into the 3-D output map; rows / columns
// past the tensor are clipped by the hardware.  Bulk-group completion is tracked per issuing thread.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// full completion: the writes of every bulk group of this thread have been performed (not just its smem reads)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M=128, K=8 (tf32)
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with bf16 operands (K = 16 per instruction)
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of the quarter)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[32]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major: 1) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// sm_100 UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool bf16 = false) {
  const uint32_t fmt = bf16 ? 1u /* BF16 */ : 2u /* TF32 */;
  return (1u << 4) /* D = F32 */ | (fmt << 7) /* A */ | (fmt << 10) /* B */ |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
  int BN;        // columns per tile (multiple of 16, <= 256)
  int stages;    // smem pipeline depth
  int tiles_m;   // row tiles per batch item
  int tiles_n;
  int total;     // tiles in the launch
  int nsplit;    // 1: W only, 2: W then W_lo
  int rev_b;     // > 0: batch items are visited in the order rev_b-1 .. 0 (rev_b = B)
  uint32_t tmem_cols;
  uint32_t idesc;
  int tma_store;  // EPI 0, fp32 output, BN % 32 == 0: the epilogue stores 32 x 32 boxes through TMA (mapD)
};

// EPI: 0 = D = acc + bias                                   (proj_1x1, in/out_proj, fc1, fc2, pw_conv)
//      1 = y = acc + bias + resid; D = prelu(cw*(mix + y) + cb)  (res_conv + concat_block of the next block)
//      2 = D = acc + bias + resid                              (res_conv of the last block)
// STATS: per-item sum / sum of squares of D (GlobLN statistics of the consumer), double atomics.
// AB16: operands are bf16 (K-block = 64 elements = the same 128-byte swizzle row, kind::f16).
// D16:  D is stored as bf16.
template <int EPI, bool STATS, bool AB16, bool D16>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapD, GemmArgs a, TcParams p) {
  grid_dep_wait();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages] x { A 16 KB | W BN*128 B | (W_lo) } then barriers, then the epilogue patches
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 4;
  const uint32_t w_bytes = (uint32_t)p.BN * TC_BK * 4;
  const uint32_t stage_bytes = a_bytes + w_bytes * p.nsplit;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = bars + 2 * p.stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // (TMA-store epilogue: the boxes must sit on 1024-byte boundaries for the 128-byte swizzle)
  float* patches = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + (p.tma_store ? 1024 : 256));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BKE = AB16 ? 2 * TC_BK : TC_BK;  // elements per 128-byte K-block
  const int nkb = a.K / BKE;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x) {
        const int mt = tile / p.tiles_n, nt = tile % p.tiles_n;
        const int b = p.rev_b ? p.rev_b - 1 - mt / p.tiles_m : mt / p.tiles_m, r0 = (mt % p.tiles_m) * TC_BM, n0 = nt * p.BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(full + stage, stage_bytes);
          tma_load_3d(sa, &mapA, full + stage, kb * BKE, r0, b);
          tma_load_2d(sa + a_bytes, &mapW, full + stage, kb * BKE, n0);
          if (p.nsplit == 2) tma_load_2d(sa + a_bytes + w_bytes, &mapW2, full + stage, kb * BKE, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x) {
      mbar_wait(acc_empty + acc, acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = make_smem_desc(sa), dw = make_smem_desc(sa + a_bytes), dw2 = make_smem_desc(sa + a_bytes + w_bytes);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            // advance 8 tf32 / 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            if constexpr (AB16) {
              tc_mma_bf16(tmem_d, da + 2 * k, dw + 2 * k, p.idesc, (kb | k) != 0);
            } else {
              tc_mma_tf32(tmem_d, da + 2 * k, dw + 2 * k, p.idesc, (kb | k) != 0);
              if (p.nsplit == 2) tc_mma_tf32(tmem_d, da + 2 * k, dw2 + 2 * k, p.idesc, 1);
            }
          }
        }
        __syncwarp();
        if (elect_one()) tc_commit(empty + stage);  // frees the smem stage once the MMAs have read it
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(acc_full + acc);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue (8 warps)
    // Two warps per TMEM lane quarter (a warp can only read lanes 32*(warp%4)..+31); the pair splits
    // the 32-column chunks of the tile even/odd.  Two warps per scheduler also hide each other's
    // instruction-fetch and shared-memory latencies.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    float* patch = patches + (warp - 2) * TC_PATCH;
    const int prow = lane >> 3, pcol = (lane & 7) * 4;  // transposed role: row within a group of 4, column group
    const int N = a.N;
    float cslope = 0.f;
    if constexpr (EPI == 1) cslope = __ldg(a.cslope);
    int tbuf = 0;  // TMA-store epilogue: which of the warp's two boxes the next chunk goes to (alternates across tiles too)
    for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x) {
      const int mt = tile / p.tiles_n, nt = tile % p.tiles_n;
      const int b = p.rev_b ? p.rev_b - 1 - mt / p.tiles_m : mt / p.tiles_m, r0 = (mt % p.tiles_m) * TC_BM, n0 = nt * p.BN;
      const int rbase = r0 + quarter * 32;
      const size_t item = (size_t)b * a.L;
      float s1 = 0.f, s2 = 0.f;
      mbar_wait(acc_full + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.BN);
      if constexpr (EPI == 0 && !D16) {
        if (p.tma_store) {
          // thread = row: bias and statistics in registers, the 32 x 32 box goes to shared memory in the 128-byte
          // swizzled layout (16-byte chunk j of row r at chunk j ^ (r & 7): conflict-free for the row-per-thread
          // writes) and one lane hands it to TMA; two boxes per warp so the next chunk never waits for the store
          float* tp = patches + (warp - 2) * TC_TPATCH;
          const bool row_ok = rbase + lane < a.L;
          // (a quarter that lies past the end of the item has nothing to store; it must not touch the boxes either:
          // staging without committing a group would break the "at most one group pending = the other box" invariant
          // of the wait below - observed as a data race at the last row tile of L = 2010 / 4010)
          for (int c0 = half * 32; c0 < p.BN && rbase < a.L; c0 += 64) {
            float v[32];
            tc_ld32(taddr + c0, v);
            if (a.bias) {
              const float4* bp = reinterpret_cast<const float4*>(a.bias + n0 + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(bp + j);
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
            }
            if constexpr (STATS) {
              if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  s1 += v[j];
                  s2 = fmaf(v[j], v[j], s2);
                }
              }
            }
            if (lane == 0) tma_store_wait_read<1>();  // the store that read this box two chunks ago
            __syncwarp();
            float* box = tp + tbuf * (32 * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(box + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_async_shared();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&mapD, box, n0 + c0, rbase, b);
              tma_store_commit();
            }
            tbuf ^= 1;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + acc);
          if constexpr (STATS) {
            const double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
            if (lane == 0) stat_add2(a.det, a.stats + 2 * b, d1, d2);
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
      }
      for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
        float v[32];
        const int ncol = min(32, p.BN - c0);  // BN is a multiple of 16
        if (ncol == 32) tc_ld32(taddr + c0, v); else tc_ld16(taddr + c0, v);
        // thread = row  ->  patch[row][col]; the 36-float row pitch keeps both sides conflict-free
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (j < ncol) *reinterpret_cast<float4*>(patch + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        if (pcol < ncol) {
          const int n = n0 + c0 + pcol;
          const float4 bias = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 cw = bias, cb = bias;
          if constexpr (EPI == 1) {
            cw = __ldg(reinterpret_cast<const float4*>(a.cw + n));
            cb = __ldg(reinterpret_cast<const float4*>(a.cb + n));
          }
          // row of this lane in iteration i: rbase + 4*i + prow; 8 lanes cover the 128 bytes of a row
          const int rfirst = rbase + prow;
          const size_t off0 = (item + rfirst) * N + n;
          float4 res[8], mix[8];
          if constexpr (EPI != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool ok = rfirst + 4 * i < a.L;
              res[i] = ok ? *reinterpret_cast<const float4*>(a.resid + off0 + (size_t)(4 * i) * N) : make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr (EPI == 1)
                mix[i] = ok ? *reinterpret_cast<const float4*>(a.mix + off0 + (size_t)(4 * i) * N) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (rfirst + 4 * i < a.L) {
              float4 t = *reinterpret_cast<const float4*>(patch + (4 * i + prow) * 36 + pcol);
              t.x += bias.x; t.y += bias.y; t.z += bias.z; t.w += bias.w;
              if constexpr (EPI != 0) {
                // res_conv(expanded) + residual (TDANet_best.py:380)
                t.x += res[i].x; t.y += res[i].y; t.z += res[i].z; t.w += res[i].w;
              }
              if constexpr (EPI == 1) {
                // concat_block(mixture + x) = PReLU(w_c*(mixture + x) + b_c) (TDANet_best.py:388-398)
                t.x = preluf_(fmaf(cw.x, mix[i].x + t.x, cb.x), cslope);
                t.y = preluf_(fmaf(cw.y, mix[i].y + t.y, cb.y), cslope);
                t.z = preluf_(fmaf(cw.z, mix[i].z + t.z, cb.z), cslope);
                t.w = preluf_(fmaf(cw.w, mix[i].w + t.w, cb.w), cslope);
              }
              if constexpr (STATS) {
                s1 += (t.x + t.y) + (t.z + t.w);
                s2 = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, s2))));
              }
              if constexpr (D16) {
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.D) + off0 + (size_t)(4 * i) * N) =
                    make_uint2(bf16x2_pack(t.x, t.y), bf16x2_pack(t.z, t.w));
              } else {
                *reinterpret_cast<float4*>(a.D + off0 + (size_t)(4 * i) * N) = t;
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
      if constexpr (STATS) {
        const double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
        if (lane == 0) stat_add2(a.det, a.stats + 2 * b, d1, d2);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  if constexpr (EPI == 0 && !D16) {
    // the stores must have been performed (not only have read their boxes) before this CTA gives up its shared memory
    // and signals completion to a programmatically dependent kernel
    if (p.tma_store && warp >= 2 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box, bool bf16 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(TDANET_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t bdim[3], estr[3] = {1, 1, 1};
  uint64_t stride = bf16 ? 2 : sizeof(float);
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    stride *= dims[i];
    if (i < rank - 1) gstride[i] = stride;
  }
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)ptr, gdim, gstride, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TDANET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

bool gemm_tc_supported(const GemmArgs& a) {
  if (a.K % (a.a_bf16 ? 2 * TC_BK : TC_BK) != 0 || a.N % 16 != 0 || a.a_slope != nullptr || a.epi == EPI_MASK) return false;
  if (a.N > 128 && a.N % 128 != 0) return false;
  const uintptr_t all = (uintptr_t)a.A | (uintptr_t)a.W | (uintptr_t)a.D | (uintptr_t)a.bias | (uintptr_t)a.resid |
                        (uintptr_t)a.mix | (uintptr_t)a.cw | (uintptr_t)a.cb;
  return (all & 15) == 0;
}

int launch_gemm_tc(const GemmArgs& a, int mode, cudaStream_t st) {
  // Shapes the tensor-core tiling cannot express (K not a multiple of 32, N not a multiple of 16,
  // an A-operand transform) run on the CUDA-core kernel of this library.
  if (!gemm_tc_supported(a)) {
    if (a.a_bf16 || a.d_bf16)
      return fail(TDANET_EUNSUPPORTED, "bf16 activation storage needs tensor-core GEMM shapes (K %% 64 == 0, N %% 16 == 0); got N=%d K=%d", a.N, a.K);
    return launch_gemm_simt(a, st);
  }
  TD_REQUIRE(a.W_aux != nullptr, "gemm_tc: prepared weights missing");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    TD_CUDA(cudaGetDevice(&dev));
    TD_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  TcParams p{};
  p.nsplit = (mode == TDANET_GEMM_TF32X3 && !a.a_bf16) ? 2 : 1;
  static const int bn_max = getenv("TDANET_GEMM_BN") ? atoi(getenv("TDANET_GEMM_BN")) : 256;  // tuning aid: 128 | 256
  p.BN = a.N % 256 == 0 && p.nsplit == 1 && bn_max >= 256 ? 256 : (a.N >= 128 ? 128 : a.N);
  p.tiles_m = cdiv(a.L, TC_BM);
  // Tile width by launch size (TDANET_GEMM_BN_MIN=32 ...): the bottom-scale GEMMs of a training batch of 8 have 8 row
  // tiles, i.e. 16-48 CTAs at 256 columns, and a CTA's time is what it pulls through its SM's L2 port (~60 GB/s), so
  // narrower tiles on more SMs shorten the launch itself.  Measured on the whole training step it is a loss, though:
  // TDANetBest +1 % (noise), the fork TDANet 44.0 -> 39.2 steps/s - a persistent CTA takes its SM's whole shared
  // memory, so a 128-CTA GEMM on the main stream evicts the side-stream kernels (weight gradients, loc_glo_fus /
  // conv_pool chains) that a 16-CTA launch lets run beside it.  Off by default (minimum = 256 columns).
  // The inference forward has no such neighbours while its bottom-scale GEMMs run (GemmArgs::narrow): there a small
  // launch (a 64-mixture job split over 8 GPUs: 8 row tiles) takes the narrow tiles.
  static const int bn_env = getenv("TDANET_GEMM_BN_MIN") ? atoi(getenv("TDANET_GEMM_BN_MIN")) : 0;
  const int bn_min = bn_env > 0 ? bn_env : (a.narrow ? 32 : 256);
  while (p.BN > bn_min && p.BN % 32 == 0 && (long)a.B * p.tiles_m * (a.N / p.BN) * 2 < num_sms) p.BN /= 2;
  p.tiles_n = a.N / p.BN;
  p.total = a.B * p.tiles_m * p.tiles_n;
  p.rev_b = a.rev ? a.B : 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.BN)) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = make_idesc(TC_BM, p.BN, a.a_bf16 != 0);
  static const bool tma_store_on = !(getenv("TDANET_GEMM_TMA_STORE") && atoi(getenv("TDANET_GEMM_TMA_STORE")) == 0);
  p.tma_store = tma_store_on && a.epi != EPI_RESIDUAL && !a.d_bf16 && !a.a_bf16 && p.BN % 32 == 0;
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 4 + (size_t)p.BN * TC_BK * 4 * p.nsplit;
  const size_t epi_bytes = p.tma_store ? 1024 + TC_EPI_WARPS * TC_TPATCH * sizeof(float) : 256 + TC_EPI_WARPS * TC_PATCH * sizeof(float);
  const size_t budget = 226 * 1024 - 1024 - epi_bytes;
  int stages = (int)(budget / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(TDANET_EUNSUPPORTED, "gemm_tc: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + epi_bytes;

  CUtensorMap mapA, mapW, mapW2;
  const uint64_t dA[3] = {(uint64_t)a.K, (uint64_t)a.L, (uint64_t)a.B};
  const uint32_t bke = a.a_bf16 ? 2 * TC_BK : TC_BK;
  const uint32_t bA[3] = {bke, TC_BM, 1};
  const uint64_t dW[2] = {(uint64_t)a.K, (uint64_t)a.N};
  const uint32_t bW[2] = {bke, (uint32_t)p.BN};
  if (int e = encode_map(&mapA, a.A, 3, dA, bA, a.a_bf16)) return e;
  // TF32: the rounded copy is the operand.  TF32X3: W itself (hi) and the prepared remainder (lo).
  // bf16 operands: the bf16 copy.
  if (int e = encode_map(&mapW, p.nsplit == 2 ? a.W : a.W_aux, 2, dW, bW, a.a_bf16)) return e;
  if (int e = encode_map(&mapW2, a.W_aux, 2, dW, bW, a.a_bf16)) return e;
  CUtensorMap mapD = mapA;   // (a valid map even when the TMA-store epilogue is off)
  if (p.tma_store) {
    const uint64_t dD[3] = {(uint64_t)a.N, (uint64_t)a.L, (uint64_t)a.B};
    const uint32_t bD[3] = {32, 32, 1};
    if (int e = encode_map(&mapD, a.D, 3, dD, bD, false)) return e;
  }

  const int grid = p.total < num_sms ? p.total : num_sms;
  const int epi = a.epi == EPI_RESIDUAL ? (a.last ? 2 : 1) : 0;
#define TD_TC_LAUNCH(E, S, AB, DB)                                                                             \
  do {                                                                                                         \
    static PerDeviceOnce first_use;                                                                            \
    if (first_use()) {                                                                                         \
      TD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<E, S, AB, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    }                                                                                                          \
    TD_LAUNCH((gemm_tc_kernel<E, S, AB, DB>), grid, TC_THREADS, smem, st, mapA, mapW, mapW2, mapD, a, p);             \
  } while (0)
  if (a.d_bf16) {
    TD_REQUIRE(epi == 0 && a.stats && !a.a_bf16, "gemm_tc: bf16 output is implemented for the statistics epilogue only");
    TD_TC_LAUNCH(0, true, false, true);
  } else if (a.a_bf16) {
    TD_REQUIRE(epi != 0, "gemm_tc: bf16 operands are implemented for the residual epilogues only");
    if (epi == 1) TD_TC_LAUNCH(1, false, true, false);
    else TD_TC_LAUNCH(2, false, true, false);
  } else if (epi == 0 && a.stats) TD_TC_LAUNCH(0, true, false, false);
  else if (epi == 0) TD_TC_LAUNCH(0, false, false, false);
  else if (epi == 1) TD_TC_LAUNCH(1, false, false, false);
  else TD_TC_LAUNCH(2, false, false, false);
#undef TD_TC_LAUNCH
  return 0;
}

// ----------------------------------------------------------------------------- weight preparation
__global__ void tf32_prepare_kernel(const float* __restrict__ w, float* __restrict__ aux, size_t n, int mode) {
  grid_dep_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i];
  if (mode == TDANET_GEMM_TF32) {
    aux[i] = tf32_rna(x);
  } else {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // what the tensor core sees
    aux[i] = x - hi;
  }
}

__global__ void bf16_prepare_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ aux, size_t n) {
  grid_dep_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) aux[i] = __float2bfloat16_rn(w[i]);
}

int launch_bf16_prepare(const float* w, float* aux, size_t n, cudaStream_t st) {
  TD_LAUNCH(bf16_prepare_kernel, (unsigned)((n + 255) / 256), 256, 0, st, w, reinterpret_cast<__nv_bfloat16*>(aux), n);
  return 0;
}

int launch_tf32_prepare(const float* w, float* aux, size_t n, int mode, cudaStream_t st) {
  TD_LAUNCH(tf32_prepare_kernel, (unsigned)((n + 255) / 256), 256, 0, st, w, aux, n, mode);
  return 0;
}

}  // namespace td
