// CPU emulation shim for the CUDA sources of the backward pass.  TEST INFRASTRUCTURE ONLY.
//
// `g++ -x c++ -DTD_EMU` compiles backward.cu / optim.cu / plan_abi.cu into a host library in which every
// kernel launch runs on the CPU: blocks one after the other, the threads of a block either one after the
// other (kernels without intra-block communication) or as real OS threads with barriers (kernels that use
// __syncthreads / warp shuffles; launched with TD_LAUNCH_COOP).  There is no GPU in the build container,
// so this is how the gradient kernels are checked against autograd of the oracle before they reach a B200
// (tests/test_backward_emu.py).  Nothing in the product path includes this file: the CUDA build never
// defines TD_EMU, and the emulation library is never loaded by tdanet_b200/.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
static inline const char* cudaGetErrorString(cudaError_t) { return "emulation"; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
constexpr int cudaMemcpyDeviceToDevice = 3;
// streams / events: everything runs in program order on the host
typedef void* cudaEvent_t;
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2;
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

namespace emu {
struct BlockState {
  std::barrier<>* block_bar = nullptr;
  std::barrier<>* warp_bar = nullptr;  // barrier of this thread's warp
  double* warp_slots = nullptr;        // 32 exchange slots of this thread's warp
};
extern thread_local uint3 tid, bid;
extern thread_local BlockState* bs;
extern dim3 bdim, gdim;
extern bool coop_reductions;

static inline void sync_block() {
  if (bs && bs->block_bar) bs->block_bar->arrive_and_wait();
}
template <class T>
static inline T shfl_xor(T v, int o) {
  // every launch that reaches here was made with TD_LAUNCH_COOP
  if (!bs) { fprintf(stderr, "emu: warp shuffle in a sequentially emulated kernel\n"); abort(); }
  const int lane = tid.x & 31;
  double* s = bs->warp_slots;
  memcpy(&s[lane], &v, sizeof(T));
  bs->warp_bar->arrive_and_wait();
  T r;
  memcpy(&r, &s[lane ^ o], sizeof(T));
  bs->warp_bar->arrive_and_wait();
  return r;
}

// TD_EMU_TIMING=1: per-launch wall time on stderr (which emulated kernel a slow test spends its time in)
struct Timer {
  const char* name;
  double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
  explicit Timer(const char* n) : name(n), t0(now()) {}
  ~Timer() {
    static const bool on = getenv("TD_EMU_TIMING") != nullptr;
    if (on) { const double dt = now() - t0; if (dt > 0.05) fprintf(stderr, "emu %8.3f s  %s\n", dt, name); }
  }
};

template <class F>
static void launch(dim3 grid, dim3 block, bool coop, F f) {
  gdim = grid;
  bdim = block;
  const unsigned nt = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        if (!coop) {
          bs = nullptr;
          bid = uint3{bx, by, bz};
          for (unsigned tz = 0; tz < block.z; ++tz)
            for (unsigned ty = 0; ty < block.y; ++ty)
              for (unsigned tx = 0; tx < block.x; ++tx) {
                tid = uint3{tx, ty, tz};
                f();
              }
          continue;
        }
        const unsigned nwarp = (nt + 31) / 32;
        std::barrier<> block_bar(nt);
        std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
        for (unsigned w = 0; w < nwarp; ++w)
          warp_bars.emplace_back(new std::barrier<>(std::min(32u, nt - w * 32)));
        std::vector<double> slots((size_t)nwarp * 32);
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t) {
          th.emplace_back([&, t]() {
            BlockState st;
            st.block_bar = &block_bar;
            st.warp_bar = warp_bars[t / 32].get();
            st.warp_slots = slots.data() + (size_t)(t / 32) * 32;
            bs = &st;
            bid = uint3{bx, by, bz};
            tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            f();
            // a thread that returns stops taking part in later barriers, like an exited CUDA thread
            st.warp_bar->arrive_and_drop();
            st.block_bar->arrive_and_drop();
            bs = nullptr;
          });
        }
        for (auto& t : th) t.join();
      }
}
}  // namespace emu

#define threadIdx emu::tid
#define blockIdx emu::bid
#define blockDim emu::bdim
#define gridDim emu::gdim
#define __syncthreads() emu::sync_block()
#define __shfl_xor_sync(mask, v, o) emu::shfl_xor((v), (o))

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return std::atomic_ref<unsigned long long>(*p).fetch_add(v, std::memory_order_relaxed);
}
static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
using std::max;
using std::min;
