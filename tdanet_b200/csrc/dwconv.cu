// Dispatch of the depthwise / LA launchers on the storage type of the large activations.
// The kernels live in dwconv_impl.cuh, compiled once per storage type (dwconv_f32.cu, dwconv_bf16.cu).
#include "kernels.h"

namespace td {
#define TD_DECLARE_ACT(NS)                                                                                    \
  namespace NS {                                                                                              \
  int launch_dw5(const DwArgs& a, cudaStream_t st);                                                           \
  int launch_dw_generic(const SrcDesc& src, int kind, int B, int C, int Lout, int ks, int stride, const float* w, \
                        const float* wT, const float* bias, float* out, int round_out, cudaStream_t st);      \
  int launch_inject_materialize(const SrcDesc& src, int kind, int B, int C, float* out, const float* wa,      \
                                const float* we, double* stats, cudaStream_t st);                             \
  int launch_inject_materialize2(const SrcDesc& sa, float* out_a, const SrcDesc& sb, float* out_b, int kind, int B, \
                                 int C, const float* wa, const float* we, double* stats, cudaStream_t st);    \
  int launch_la_combine(const LaArgs& a, cudaStream_t st);                                                    \
  int launch_la_local_stats(const DwArgs* steps, int n, cudaStream_t st);                                     \
  }
TD_DECLARE_ACT(act_f32)
TD_DECLARE_ACT(act_bf16)
#undef TD_DECLARE_ACT

int launch_dw5(const DwArgs& a, cudaStream_t st) {
  return a.act_bf16 ? act_bf16::launch_dw5(a, st) : act_f32::launch_dw5(a, st);
}
int launch_dw_generic(const SrcDesc& src, int kind, int B, int C, int Lout, int ks, int stride, const float* w,
                      const float* wT, const float* bias, float* out, int round_out, int act_bf16, cudaStream_t st) {
  return act_bf16 ? act_bf16::launch_dw_generic(src, kind, B, C, Lout, ks, stride, w, wT, bias, out, round_out, st)
                  : act_f32::launch_dw_generic(src, kind, B, C, Lout, ks, stride, w, wT, bias, out, round_out, st);
}
int launch_inject_materialize(const SrcDesc& src, int kind, int B, int C, float* out, int act_bf16, cudaStream_t st,
                              const float* wa, const float* we, double* stats) {
  return act_bf16 ? act_bf16::launch_inject_materialize(src, kind, B, C, out, wa, we, stats, st)
                  : act_f32::launch_inject_materialize(src, kind, B, C, out, wa, we, stats, st);
}
int launch_inject_materialize2(const SrcDesc& sa, float* out_a, const SrcDesc& sb, float* out_b, int kind, int B, int C,
                               int act_bf16, cudaStream_t st, const float* wa, const float* we, double* stats) {
  return act_bf16 ? act_bf16::launch_inject_materialize2(sa, out_a, sb, out_b, kind, B, C, wa, we, stats, st)
                  : act_f32::launch_inject_materialize2(sa, out_a, sb, out_b, kind, B, C, wa, we, stats, st);
}
int launch_la_combine(const LaArgs& a, cudaStream_t st) {
  return a.act_bf16 ? act_bf16::launch_la_combine(a, st) : act_f32::launch_la_combine(a, st);
}
int launch_la_local_stats(const DwArgs* steps, int n, cudaStream_t st) {
  return steps[0].act_bf16 ? act_bf16::launch_la_local_stats(steps, n, st) : act_f32::launch_la_local_stats(steps, n, st);
}
}  // namespace td
