// bf16 activation storage instantiation of the depthwise / LA kernels (fp32 arithmetic and statistics)
#include <cuda_bf16.h>
#define ACT_T __nv_bfloat16
#define TD_ACT_NS act_bf16
#include "dwconv_impl.cuh"
