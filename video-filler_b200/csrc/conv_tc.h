// conv_tc.h -- entry points of the BF16 tcgen05 implicit-GEMM convolution path (conv_tc.cu).
// NCHW-fp32 wrappers (op-level ABI, precision CENN_BF16) return 0 = done, >0 = shape not handled
// by the tensor-core path (caller runs the fp32 SIMT kernel instead), <0 = error (message set).
#pragma once
#include "common.cuh"

int tc_conv_fprop_nchw(cenn_state *s, const float *x, const float *w, const float *bias, float *out,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW);
int tc_conv_dgrad_nchw(cenn_state *s, const float *gy, const float *w, const float *bias, float *gx,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW);
int tc_conv_wgrad_nchw(cenn_state *s, const float *x, const float *gy, float *gw,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW,
                       float scale, int overwrite);
