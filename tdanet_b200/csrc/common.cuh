// Shared device/host helpers for the tdanet_b200 kernels (sm_100a only).
#pragma once
#ifdef TD_EMU
#include "emu.h"  // CPU emulation build of the backward pass (tests only)
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#endif
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstdarg>
#include "../../include/tdanet_b200.h"

namespace td {

// ----------------------------------------------------------------------------- host side
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int fail(int code, const char* fmt, ...);

#define TD_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return td::fail(TDANET_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                               \
  } while (0)

#define TD_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) return td::fail(TDANET_EINVAL, __VA_ARGS__);                            \
  } while (0)

// Optional per-launch CUDA-event timing (tdanet_profile_enable): events are recorded on the stream
// the kernel is launched on, immediately before and after it.
extern bool g_profile;
extern thread_local const char* g_tag;  // role of the next launches ("proj", "la_combine", ...), or null
void profile_mark(const char* name, cudaStream_t st, bool begin);
struct Tag {
  const char* prev;
  explicit Tag(const char* t) : prev(g_tag) { g_tag = t; }
  ~Tag() { g_tag = prev; }
};

// Every kernel launch goes through this so that tdanet_launch_count() is honest.
// TD_LAUNCH_COOP marks kernels whose threads communicate (__syncthreads / shuffles); the distinction only
// matters to the CPU emulation build.
#ifdef TD_EMU
#define TD_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                   \
    emu::Timer _t(#kernel);                                                              \
    emu::launch(dim3(grid), dim3(block), false, [&]() { kernel(__VA_ARGS__); });        \
    td::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
  } while (0)
#define TD_LAUNCH_COOP(kernel, grid, block, smem, stream, ...)                           \
  do {                                                                                   \
    emu::Timer _t(#kernel);                                                              \
    emu::launch(dim3(grid), dim3(block), true, [&]() { kernel(__VA_ARGS__); });         \
    td::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
  } while (0)
// kernels whose only intra-block communication is a final block reduction (sequential unless TD_EMU_COOP is set)
#define TD_LAUNCH_RED(kernel, grid, block, smem, stream, ...)                            \
  do {                                                                                   \
    emu::Timer _t(#kernel);                                                              \
    emu::launch(dim3(grid), dim3(block), emu::coop_reductions, [&]() { kernel(__VA_ARGS__); }); \
    td::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
  } while (0)
#else
#define TD_LAUNCH_COOP TD_LAUNCH
#define TD_LAUNCH_RED TD_LAUNCH
#define TD_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                   \
    if (td::g_profile) td::profile_mark(#kernel, (stream), true);                        \
    td::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__); \
    if (td::g_profile) td::profile_mark(#kernel, (stream), false);                       \
    td::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t _e = cudaPeekAtLastError();                                              \
    if (_e != cudaSuccess)                                                               \
      return td::fail(TDANET_ECUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_e)); \
  } while (0)
#endif

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// The opt-in to more than 48 KB of dynamic shared memory (cudaFuncSetAttribute) is per DEVICE: a call site keeps one
// of these as a function-local static and sets its attributes whenever it reports the first use on the current device
// (a process may drive several GPUs, e.g. a model moved from cuda:0 to cuda:1).
struct PerDeviceOnce {
  bool done[16] = {};
  bool operator()() {
#ifdef TD_EMU
    return false;
#else
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 16) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
#endif
  }
};

// Programmatic dependent launch: every kernel of this library begins with grid_dep_wait() (griddepcontrol.wait: block
// until the previous kernel of the stream has completed and its writes are visible) and is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs are scheduled - launch latency, parameter and
// constant setup - while the previous kernel drains instead of after it.  A forward is ~470 dependent launches and
// a training step ~1650, most of them 10-50 us long.  TDANET_PDL=0 in the environment restores plain launches.
#ifdef TD_EMU
__device__ __forceinline__ void grid_dep_wait() {}
#else
#ifndef TD_PDL_EARLY
#define TD_PDL_EARLY 0
#endif
__device__ __forceinline__ void grid_dep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if TD_PDL_EARLY
  // let the next kernel's CTAs become resident as this kernel's CTAs retire (they park in their own wait)
  asm volatile("griddepcontrol.launch_dependents;");
#endif
}
extern bool g_pdl;
template <typename... KArgs, typename... Args>
static inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through cudaPeekAtLastError
}
#endif

// ----------------------------------------------------------------------------- device side
constexpr float kEpsGLN = 1e-8f;
constexpr float kEpsLN = 1e-5f;

// V consecutive channels held by one thread (V = 2 or 4)
template <int V>
struct vf {
  float v[V];
  __device__ __forceinline__ float& operator[](int i) { return v[i]; }
  __device__ __forceinline__ const float& operator[](int i) const { return v[i]; }
};

template <int V>
__device__ __forceinline__ vf<V> vzero() {
  vf<V> r;
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = 0.f;
  return r;
}

template <int V>
__device__ __forceinline__ vf<V> vload(const float* __restrict__ p) {
  vf<V> r;
  if constexpr (V == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

// load of a location the same kernel also writes (read-modify-write): no non-coherent path
template <int V>
__device__ __forceinline__ vf<V> vload_rw(const float* p) {
  vf<V> r;
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (V == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}

template <int V>
__device__ __forceinline__ void vstore(float* __restrict__ p, const vf<V>& r) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

// Typed activation access: the large activations are stored as fp32 or bf16 (arithmetic is always fp32).
template <int V>
__device__ __forceinline__ vf<V> aload(const float* __restrict__ p) { return vload<V>(p); }
template <int V>
__device__ __forceinline__ void astore(float* __restrict__ p, const vf<V>& r) { vstore<V>(p, r); }
template <int V>
__device__ __forceinline__ vf<V> alds(const float* p) {  // shared memory
  vf<V> r;
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}
#ifndef TD_EMU
__device__ __forceinline__ void bf16x2_unpack(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t bf16x2_pack(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);  // round to nearest even
  return *reinterpret_cast<const uint32_t*>(&t);
}
template <int V>
__device__ __forceinline__ vf<V> aload(const __nv_bfloat16* __restrict__ p) {
  vf<V> r;
  if constexpr (V == 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    bf16x2_unpack(t.x, r.v[0], r.v[1]);
    bf16x2_unpack(t.y, r.v[2], r.v[3]);
  } else {
    const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
    bf16x2_unpack(t, r.v[0], r.v[1]);
  }
  return r;
}
template <int V>
__device__ __forceinline__ vf<V> alds(const __nv_bfloat16* p) {
  vf<V> r;
  if constexpr (V == 4) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    bf16x2_unpack(t.x, r.v[0], r.v[1]);
    bf16x2_unpack(t.y, r.v[2], r.v[3]);
  } else {
    const uint32_t t = *reinterpret_cast<const uint32_t*>(p);
    bf16x2_unpack(t, r.v[0], r.v[1]);
  }
  return r;
}
template <int V>
__device__ __forceinline__ void astore(__nv_bfloat16* __restrict__ p, const vf<V>& r) {
  if constexpr (V == 4) {
    *reinterpret_cast<uint2*>(p) = make_uint2(bf16x2_pack(r.v[0], r.v[1]), bf16x2_pack(r.v[2], r.v[3]));
  } else {
    *reinterpret_cast<uint32_t*>(p) = bf16x2_pack(r.v[0], r.v[1]);
  }
}

#endif  // !TD_EMU

// Run-time typed read of a stored activation (the backward kernels are compiled once; which of their inputs were
// stored as bf16 is a launch argument): `base` is the start of the tensor, `off` an ELEMENT offset.  Written so that
// the CPU emulation build (emu.h) can execute it.
template <int V>
__device__ __forceinline__ vf<V> act_vload(const float* __restrict__ base, size_t off, int bf16) {
  if (!bf16) return vload<V>(base + off);
  const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + off;
  vf<V> r;
#ifdef TD_EMU
  for (int e = 0; e < V; ++e) r.v[e] = __uint_as_float((uint32_t)p[e] << 16);
#else
  if constexpr (V == 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    r.v[0] = __uint_as_float(t.x << 16); r.v[1] = __uint_as_float(t.x & 0xffff0000u);
    r.v[2] = __uint_as_float(t.y << 16); r.v[3] = __uint_as_float(t.y & 0xffff0000u);
  } else if constexpr (V == 2) {
    const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
    r.v[0] = __uint_as_float(t << 16); r.v[1] = __uint_as_float(t & 0xffff0000u);
  } else {
    r.v[0] = __uint_as_float((uint32_t)__ldg(p) << 16);
  }
#endif
  return r;
}

// compile-time typed form
template <int V, bool BF>
__device__ __forceinline__ vf<V> act_vload_t(const float* __restrict__ base, size_t off) {
  if constexpr (BF) return act_vload<V>(base, off, 1);
  else return vload<V>(base + off);
}

// fire-and-forget vector reduction into global memory (red.global.add.v{2,4}.f32 on sm_90+)
template <int V>
__device__ __forceinline__ void vred_add(float* p, const vf<V>& r) {
#ifdef TD_EMU
  for (int e = 0; e < V; ++e) atomicAdd(p + e, r.v[e]);
#else
  if constexpr (V == 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]),
                 "f"(r.v[2]), "f"(r.v[3])
                 : "memory");
  } else if constexpr (V == 2) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]) : "memory");
  } else {
    atomicAdd(p, r.v[0]);
  }
#endif
}

// ----------------------------------------------------------------------------- deterministic statistics
// The GlobLN sums are accumulated across CTAs with floating-point atomics, so two runs of one forward differ in the
// last bits (the order of the adds).  The deterministic mode (tdanet_set_deterministic / TDANET_DETERMINISTIC=1,
// inference) makes every such add EXACT instead: a partial v is split into two integers, hi = floor(v * 2^8) and
// lo = trunc((v * 2^8 - hi) * 2^48), which are added to a pair of 64-bit words with integer atomics - integer
// addition is associative, so the pair is independent of the order - and a small kernel (det_finalize_kernel,
// engine.cu) writes hi * 2^-8 + lo * 2^-56 into the slot the consumers read, before they run.  Range |sum| < 2^54,
// resolution 2^-56 per partial, up to 2^15 partials per slot; a non-finite partial saturates (the mode is for
// bisecting, not for diagnosing NaNs).  The pair of a slot at byte offset o of the statistics arena `base` lives at
// shadow + 4 * o: a float slot owns 16 bytes there, a double slot 32 (the first 16 used).
struct DetRef {
  const char* base = nullptr;  // null: off (plain floating-point atomics)
  char* shadow = nullptr;
};
__device__ __forceinline__ void det_add(const DetRef& d, const void* slot, double v) {
  unsigned long long* s =
      reinterpret_cast<unsigned long long*>(d.shadow + 4 * (reinterpret_cast<const char*>(slot) - d.base));
  double sc = v * 256.0;
  if (!(sc < 4.6e18)) sc = 4.6e18;     // also catches NaN
  if (sc < -4.6e18) sc = -4.6e18;
  const double fl = floor(sc);
  const long long hi = (long long)fl;
  const unsigned long long lo = (unsigned long long)((sc - fl) * 281474976710656.0);  // 2^48
  atomicAdd(s, (unsigned long long)hi);
  atomicAdd(s + 1, lo);
}
__device__ __forceinline__ double det_value(const unsigned long long* s) {
  return (double)(long long)s[0] * 0.00390625 + (double)s[1] * 1.3877787807814457e-17;  // 2^-8, 2^-56
}
// dst[0] += a, dst[1] += b (one thread per CTA calls it)
__device__ __forceinline__ void stat_add2(const DetRef& d, double* dst, double a, double b) {
  if (d.base) {
    det_add(d, dst, a);
    det_add(d, dst + 1, b);
    return;
  }
  atomicAdd(dst, a);
  atomicAdd(dst + 1, b);
}
template <int V>
__device__ __forceinline__ void vstat_add(const DetRef& d, float* p, const vf<V>& r) {
  if (d.base) {
#pragma unroll
    for (int e = 0; e < V; ++e) det_add(d, p + e, (double)r.v[e]);
    return;
  }
  vred_add<V>(p, r);
}

// round-to-nearest (ties away) to TF32: what producers of GEMM-only operands store in TF32 mode so
// that the tensor core's truncation of the low 13 mantissa bits is exact
__device__ __forceinline__ float tf32_rna(float x) {
#ifdef TD_EMU
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#else
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
#endif
}
template <int V>
__device__ __forceinline__ void vround_tf32(vf<V>& r) {
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = tf32_rna(r.v[i]);
}

// element-wise a*b+c / a+b on V channels; even V goes through the packed fp32x2 pipe of sm_100
// (FFMA2 / FADD2: one issue slot for two channels)
template <int V>
__device__ __forceinline__ vf<V> vfma(const vf<V>& a, const vf<V>& b, const vf<V>& c) {
  vf<V> r;
  if constexpr (V % 2 == 0) {
#pragma unroll
    for (int e = 0; e < V; e += 2) {
      const float2 t = __ffma2_rn(make_float2(a.v[e], a.v[e + 1]), make_float2(b.v[e], b.v[e + 1]),
                                  make_float2(c.v[e], c.v[e + 1]));
      r.v[e] = t.x;
      r.v[e + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < V; ++e) r.v[e] = fmaf(a.v[e], b.v[e], c.v[e]);
  }
  return r;
}
template <int V>
__device__ __forceinline__ vf<V> vadd(const vf<V>& a, const vf<V>& b) {
  vf<V> r;
  if constexpr (V % 2 == 0) {
#pragma unroll
    for (int e = 0; e < V; e += 2) {
      const float2 t = __fadd2_rn(make_float2(a.v[e], a.v[e + 1]), make_float2(b.v[e], b.v[e + 1]));
      r.v[e] = t.x;
      r.v[e + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < V; ++e) r.v[e] = a.v[e] + b.v[e];
  }
  return r;
}

// sigmoid on the SFU: 1 / (1 + 2^(-x*log2(e))) with ex2.approx / rcp.approx (4 instructions; exact limits
// 0 and 1 for large |x|, relative error ~1e-6)
__device__ __forceinline__ float sigmoidf_(float x) {
#ifdef TD_EMU
  return 1.f / (1.f + expf(-x));
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
#endif
}
__device__ __forceinline__ float preluf_(float x, float a) { return x >= 0.f ? x : a * x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of two doubles; result valid in every thread.  `sh` needs 2*32 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // protect sh against a previous use
  if (l == 0) { sh[w] = a; sh[32 + w] = b; }
  __syncthreads();
  a = 0.0; b = 0.0;
  for (int i = 0; i < nw; ++i) { a += sh[i]; b += sh[32 + i]; }
}

// Adds (a, b) to dst[0], dst[1]: one atomic pair per CTA.  Every thread of the CTA must call it.
__device__ __forceinline__ void block_accum2(double* dst, double a, double b) {
#ifdef TD_EMU
  if (!emu::bs) {  // sequential emulation: no block to reduce over
    atomicAdd(dst, a);
    atomicAdd(dst + 1, b);
    return;
  }
#endif
  __shared__ double sh[64];
  block_sum2(a, b, sh);
  if (threadIdx.x == 0) {
    atomicAdd(dst, a);
    atomicAdd(dst + 1, b);
  }
}

// A GlobLN (TDANet_best.py:47-64) whose per-item statistics a producer kernel has accumulated:
// stats[b*item_stride + {0,1}] = sum, sum of squares (double) over `count` elements of item b.
// Consumers fold the normalisation into one FMA per element: y = x*scale_c + shift_c with
//   scale_c = gamma_c * r_b,  shift_c = beta_c - gamma_c * mu_b * r_b,  r_b = 1/sqrt(var_b + 1e-8).
struct NormRef {
  const double* stats;
  int item_stride;  // doubles between items
  double count;
  const float* gamma;
  const float* beta;
};

__device__ __forceinline__ void norm_moments(const NormRef& n, int b, float& r, float& mur) {
  // the cancellation-prone part (E[x^2] - mu^2) in double, the rest in float: FP64 is scarce on this part
  const double inv = 1.0 / n.count;  // folded by the compiler into the launch constants where possible
  const double mu = n.stats[(size_t)b * n.item_stride] * inv;
  const double var = fma(-mu, mu, n.stats[(size_t)b * n.item_stride + 1] * inv);
  r = rsqrtf(fmaxf((float)var, 0.f) + kEpsGLN);
  mur = (float)mu * r;
}

template <int V>
__device__ __forceinline__ void norm_coef(const NormRef& n, int b, int ch, vf<V>& scale, vf<V>& shift) {
  float r, mur;
  norm_moments(n, b, r, mur);
  const vf<V> g = vload<V>(n.gamma + ch), be = vload<V>(n.beta + ch);
#pragma unroll
  for (int e = 0; e < V; ++e) {
    scale[e] = g[e] * r;
    shift[e] = fmaf(-g[e], mur, be[e]);
  }
}

// index of the source row for F.interpolate(mode="nearest"): min(floor(dst * fl32(in/out)), in-1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_len) {
  int s = (int)floorf((float)dst * scale);
  return s < in_len - 1 ? s : in_len - 1;
}
static inline float nearest_scale(int in_len, int out_len) { return (float)in_len / (float)out_len; }

}  // namespace td
