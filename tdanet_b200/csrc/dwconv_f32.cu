// fp32 activation storage instantiation of the depthwise / LA kernels
#define ACT_T float
#define TD_ACT_NS act_f32
#include "dwconv_impl.cuh"
