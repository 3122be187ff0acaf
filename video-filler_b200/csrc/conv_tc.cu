// conv_tc.cu -- placeholder until the tcgen05 path lands: report "shape not handled" so the op-level
// ABI runs the fp32 SIMT kernels.
#include "conv_tc.h"
int tc_conv_fprop_nchw(cenn_state *, const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int) { return 1; }
int tc_conv_dgrad_nchw(cenn_state *, const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int) { return 1; }
int tc_conv_wgrad_nchw(cenn_state *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int, float, int) { return 1; }
