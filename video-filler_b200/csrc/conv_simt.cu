// conv_simt.cu -- fp32 "parity mode" convolutions (CENN_FP32): SIMT implicit GEMM on NCHW fp32,
// any kernel/stride/pad.  Three gather-GEMM kernels serve all six THNN entry points:
//   conv fprop  == full-conv dgrad      
//   conv dgrad  == full-conv fprop
//   conv wgrad  == full-conv wgrad with (input, gradOutput) swapped
// The BF16 tensor-core path (conv_tc.cu) is selected by cenn_set_precision(CENN_BF16).
#include "common.cuh"
#include "conv_tc.h"

namespace {

struct ConvGeom {
    int N, C, H, W;      // "input" side (the side with the larger spatial extent for stride>1)
    int O, oH, oW;       // "output" side of the ordinary convolution
    int kH, kW, dH, dW, pH, pW;
};

// ---- fprop:  out[n,o,oy,ox] = sum_{c,u,v} w[o,c,u,v] x[n,c,oy*d-p+u,ox*d-p+v] (+bias[o])
struct FpropA {
    ConvGeom g; const float *x;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int ox = m % g.oW, t = m / g.oW, oy = t % g.oH, n = t / g.oH;
        int v = k % g.kW, t2 = k / g.kW, u = t2 % g.kH, c = t2 / g.kH;
        int iy = oy * g.dH - g.pH + u, ix = ox * g.dW - g.pW + v;
        if ((unsigned)iy >= (unsigned)g.H || (unsigned)ix >= (unsigned)g.W) return 0.f;
        return __ldg(x + ((int64_t)(n * g.C + c) * g.H + iy) * g.W + ix);
    }
};
struct FpropB {
    const float *w; int K;
    __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(w + (int64_t)n * K + k); }
};
struct FpropEp {
    ConvGeom g; float *out; const float *bias;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        int ox = m % g.oW, t = m / g.oW, oy = t % g.oH, img = t / g.oH;
        if (bias) v += __ldg(bias + n);
        out[((int64_t)(img * g.O + n) * g.oH + oy) * g.oW + ox] = v;
    }
};

// ---- dgrad:  gx[n,c,y,x] = sum_{o,u,v} gy[n,o,(y+p-u)/d,(x+p-v)/d] w[o,c,u,v]   (+bias[c] when used as full-conv fprop)
struct DgradA {
    ConvGeom g; const float *gy;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int x = m % g.W, t = m / g.W, y = t % g.H, n = t / g.H;
        int v = k % g.kW, t2 = k / g.kW, u = t2 % g.kH, o = t2 / g.kH;
        int ty = y + g.pH - u, tx = x + g.pW - v;
        if (ty < 0 || tx < 0) return 0.f;
        int oy = ty / g.dH, ox = tx / g.dW;
        if (oy * g.dH != ty || ox * g.dW != tx || oy >= g.oH || ox >= g.oW) return 0.f;
        return __ldg(gy + ((int64_t)(n * g.O + o) * g.oH + oy) * g.oW + ox);
    }
};
struct DgradB {
    const float *w; int C, kk;
    __device__ __forceinline__ float operator()(int k, int c) const {
        int uv = k % kk, o = k / kk;
        return __ldg(w + ((int64_t)o * C + c) * kk + uv);
    }
};
struct DgradEp {
    ConvGeom g; float *gx; const float *bias;
    __device__ __forceinline__ void operator()(int m, int c, float v) const {
        int x = m % g.W, t = m / g.W, y = t % g.H, n = t / g.H;
        if (bias) v += __ldg(bias + c);
        gx[((int64_t)(n * g.C + c) * g.H + y) * g.W + x] = v;
    }
};

// ---- wgrad:  gw[o,(c,u,v)] += scale * sum_{n,oy,ox} gy[n,o,oy,ox] x[n,c,oy*d-p+u,ox*d-p+v]
struct WgradA {
    ConvGeom g; const float *gy;
    __device__ __forceinline__ float operator()(int o, int k) const {
        int hw = g.oH * g.oW;
        int pix = k % hw, n = k / hw;
        return __ldg(gy + (int64_t)(n * g.O + o) * hw + pix);
    }
};
struct WgradB {
    ConvGeom g; const float *x;
    __device__ __forceinline__ float operator()(int k, int j) const {
        int ox = k % g.oW, t = k / g.oW, oy = t % g.oH, n = t / g.oH;
        int v = j % g.kW, t2 = j / g.kW, u = t2 % g.kH, c = t2 / g.kH;
        int iy = oy * g.dH - g.pH + u, ix = ox * g.dW - g.pW + v;
        if ((unsigned)iy >= (unsigned)g.H || (unsigned)ix >= (unsigned)g.W) return 0.f;
        return __ldg(x + ((int64_t)(n * g.C + c) * g.H + iy) * g.W + ix);
    }
};
struct WgradEp {
    float *gw; int ncols; float scale;
    __device__ __forceinline__ void operator()(int o, int j, float v) const { atomicAdd(gw + (int64_t)o * ncols + j, scale * v); }
};

// C[m,n] = sum_k A(m,k) B(k,n); 64x64x16 tiles, 256 threads, 4x4 micro-tile; blockIdx.z = K split.
template <bool A_KFAST, class AF, class BF, class EP>
__global__ void __launch_bounds__(256) gather_gemm_kernel(int M, int N, int K, int kchunk, AF af, BF bf, EP ep) {
    __shared__ __align__(16) float As[16][68];
    __shared__ __align__(16) float Bs[16][68];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        float ra[4], rb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int ml, kl;
            if (A_KFAST) { kl = tid & 15; ml = (tid >> 4) + 16 * i; } else { ml = tid & 63; kl = (tid >> 6) + 4 * i; }
            int m = m0 + ml, k = k0 + kl;
            ra[i] = (m < M && k < kend) ? af(m, k) : 0.f;
            int kb = tid & 15, nl = (tid >> 4) + 16 * i;
            int n = n0 + nl, k2 = k0 + kb;
            rb[i] = (n < N && k2 < kend) ? bf(k2, n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (A_KFAST) As[tid & 15][(tid >> 4) + 16 * i] = ra[i]; else As[(tid >> 6) + 4 * i][tid & 63] = ra[i];
            Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) ep(m, n, acc[i][j]);
        }
}

// gb[o] += scale * sum_{n,pix} gy[n,o,pix]   (one block per channel, double accumulation)
__global__ void __launch_bounds__(256) bias_grad_kernel(const float *__restrict__ gy, float *__restrict__ gb, int N, int O, int hw, float scale) {
    __shared__ double sh[32];
    int o = blockIdx.x;
    double acc = 0.0;
    for (int n = 0; n < N; ++n) {
        const float *p = gy + (int64_t)(n * O + o) * hw;
        for (int i = threadIdx.x; i < hw; i += blockDim.x) acc += (double)p[i];
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) gb[o] += scale * (float)acc;
}

int check_geom(const ConvGeom &g) {
    REQUIRE(g.N > 0 && g.C > 0 && g.H > 0 && g.W > 0 && g.O > 0, "conv: non-positive size");
    REQUIRE(g.kH > 0 && g.kW > 0 && g.dH > 0 && g.dW > 0 && g.pH >= 0 && g.pW >= 0, "conv: bad kernel/stride/pad");
    REQUIRE(g.oH > 0 && g.oW > 0, "conv: calculated output size is too small (%d x %d)", g.oH, g.oW);
    REQUIRE((int64_t)g.N * g.C * g.H * g.W < (1LL << 31) && (int64_t)g.N * g.O * g.oH * g.oW < (1LL << 31), "conv: tensor too large for 32-bit pixel index");
    return 0;
}

int simt_fprop(cenn_state *s, const ConvGeom &g, const float *x, const float *w, const float *bias, float *out) {
    int M = g.N * g.oH * g.oW, N = g.O, K = g.C * g.kH * g.kW;
    dim3 grid((M + 63) / 64, (N + 63) / 64, 1);
    gather_gemm_kernel<false><<<grid, 256, 0, s->stream>>>(M, N, K, K, FpropA{g, x}, FpropB{w, K}, FpropEp{g, out, bias});
    CK_LAUNCH(s);
    return 0;
}
int simt_dgrad(cenn_state *s, const ConvGeom &g, const float *gy, const float *w, const float *bias, float *gx) {
    int M = g.N * g.H * g.W, N = g.C, K = g.O * g.kH * g.kW;
    dim3 grid((M + 63) / 64, (N + 63) / 64, 1);
    gather_gemm_kernel<false><<<grid, 256, 0, s->stream>>>(M, N, K, K, DgradA{g, gy}, DgradB{w, g.C, g.kH * g.kW}, DgradEp{g, gx, bias});
    CK_LAUNCH(s);
    return 0;
}
int simt_wgrad(cenn_state *s, const ConvGeom &g, const float *x, const float *gy, float *gw, float scale) {
    int M = g.O, N = g.C * g.kH * g.kW, K = g.N * g.oH * g.oW;
    int tiles = ((M + 63) / 64) * ((N + 63) / 64);
    int splits = (2 * s->sm_count + tiles - 1) / tiles;
    int maxsplit = (K + 255) / 256;
    if (splits > maxsplit) splits = maxsplit;
    if (splits < 1) splits = 1;
    int kchunk = (((K + splits - 1) / splits) + 15) / 16 * 16;
    splits = (K + kchunk - 1) / kchunk;
    dim3 grid((M + 63) / 64, (N + 63) / 64, splits);
    gather_gemm_kernel<true><<<grid, 256, 0, s->stream>>>(M, N, K, kchunk, WgradA{g, gy}, WgradB{g, x}, WgradEp{gw, N, scale});
    CK_LAUNCH(s);
    return 0;
}
int simt_bias_grad(cenn_state *s, const float *gy, float *gb, int N, int O, int hw, float scale) {
    bias_grad_kernel<<<O, 256, 0, s->stream>>>(gy, gb, N, O, hw, scale);
    CK_LAUNCH(s);
    return 0;
}

ConvGeom conv_geom(int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH) {
    ConvGeom g;
    g.N = (int)batch; g.C = (int)nIn; g.H = (int)inH; g.W = (int)inW; g.O = (int)nOut;
    g.kH = kH; g.kW = kW; g.dH = dH; g.dW = dW; g.pH = padH; g.pW = padW;
    g.oH = dH > 0 ? (int)((inH + 2 * padH - kH) / dH + 1) : 0;
    g.oW = dW > 0 ? (int)((inW + 2 * padW - kW) / dW + 1) : 0;
    if (inH + 2 * padH < kH) g.oH = 0;
    if (inW + 2 * padW < kW) g.oW = 0;
    return g;
}
// full conv: the *output* of the transposed conv plays the role of the ordinary conv's input
ConvGeom fullconv_geom(int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH,
                       int padW, int padH, int adjW, int adjH) {
    ConvGeom g;
    g.N = (int)batch; g.O = (int)nIn; g.oH = (int)inH; g.oW = (int)inW; g.C = (int)nOut;
    g.kH = kH; g.kW = kW; g.dH = dH; g.dW = dW; g.pH = padH; g.pW = padW;
    g.H = (int)((inH - 1) * dH - 2 * padH + kH + adjH);
    g.W = (int)((inW - 1) * dW - 2 * padW + kW + adjW);
    return g;
}

}  // namespace

extern "C" {

int cenn_SpatialConvolutionMM_updateOutput(cenn_state *s, const float *input, float *output, const float *weight, const float *bias,
        int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH) {
    API_BEGIN(s);
    REQUIRE(input && output && weight, "SpatialConvolutionMM_updateOutput: null tensor");
    ConvGeom g = conv_geom(batch, nIn, inH, inW, nOut, kW, kH, dW, dH, padW, padH);
    if (check_geom(g)) return 1;
    if (s->precision == CENN_BF16) {
        int rc = tc_conv_fprop_nchw(s, input, weight, bias, output, g.N, g.C, g.H, g.W, g.O, kH, kW, dH, dW, padH, padW);
        if (rc <= 0) return -rc;   // 0 = done, <0 = error, >0 = shape not supported by the tensor-core path
    }
    return simt_fprop(s, g, input, weight, bias, output);
}

int cenn_SpatialConvolutionMM_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *weight,
        int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH) {
    API_BEGIN(s);
    REQUIRE(gradOutput && gradInput && weight, "SpatialConvolutionMM_updateGradInput: null tensor");
    ConvGeom g = conv_geom(batch, nIn, inH, inW, nOut, kW, kH, dW, dH, padW, padH);
    if (check_geom(g)) return 1;
    if (s->precision == CENN_BF16) {
        int rc = tc_conv_dgrad_nchw(s, gradOutput, weight, nullptr, gradInput, g.N, g.C, g.H, g.W, g.O, kH, kW, dH, dW, padH, padW);
        if (rc <= 0) return -rc;
    }
    return simt_dgrad(s, g, gradOutput, weight, nullptr, gradInput);
}

int cenn_SpatialConvolutionMM_accGradParameters(cenn_state *s, const float *input, const float *gradOutput, float *gradWeight, float *gradBias,
        int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH, float scale) {
    API_BEGIN(s);
    REQUIRE(input && gradOutput && gradWeight, "SpatialConvolutionMM_accGradParameters: null tensor");
    ConvGeom g = conv_geom(batch, nIn, inH, inW, nOut, kW, kH, dW, dH, padW, padH);
    if (check_geom(g)) return 1;
    if (gradBias && simt_bias_grad(s, gradOutput, gradBias, g.N, g.O, g.oH * g.oW, scale)) return 1;
    if (s->precision == CENN_BF16) {
        int rc = tc_conv_wgrad_nchw(s, input, gradOutput, gradWeight, g.N, g.C, g.H, g.W, g.O, kH, kW, dH, dW, padH, padW, scale, 0);
        if (rc <= 0) return -rc;
    }
    return simt_wgrad(s, g, input, gradOutput, gradWeight, scale);
}

int cenn_SpatialFullConvolution_updateOutput(cenn_state *s, const float *input, float *output, const float *weight, const float *bias,
        int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH) {
    API_BEGIN(s);
    REQUIRE(input && output && weight, "SpatialFullConvolution_updateOutput: null tensor");
    REQUIRE(adjW < dW && adjH < dH || (adjW == 0 && adjH == 0), "SpatialFullConvolution: adj must be smaller than stride");
    ConvGeom g = fullconv_geom(batch, nIn, inH, inW, nOut, kW, kH, dW, dH, padW, padH, adjW, adjH);
    if (check_geom(g)) return 1;
    if (s->precision == CENN_BF16 && adjW == 0 && adjH == 0) {
        int rc = tc_conv_dgrad_nchw(s, input, weight, bias, output, g.N, g.C, g.H, g.W, g.O, kH, kW, dH, dW, padH, padW);
        if (rc <= 0) return -rc;
    }
    return simt_dgrad(s, g, input, weight, bias, output);
}

int cenn_SpatialFullConvolution_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *weight,
        int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH) {
    API_BEGIN(s);
    REQUIRE(gradOutput && gradInput && weight, "SpatialFullConvolution_updateGradInput: null tensor");
    ConvGeom g = fullconv_geom(batch, nIn, inH, inW, nOut, kW, kH, dW, dH, padW, padH, adjW, adjH);
    if (check_geom(g)) return 1;
    if (s->precision == CENN_BF16 && adjW == 0 && adjH == 0) {
        int rc = tc_conv_fprop_nchw(s, gradOutput, weight, nullptr, gradInput, g.N, g.C, g.H, g.W, g.O, kH, kW, dH, dW, padH, padW);
        if (rc <= 0) return -rc;
    }
    return simt_fprop(s, g, gradOutput, weight, nullptr, gradInput);
}

int cenn_SpatialFullConvolution_accGradParameters(cenn_state *s, const float *input, const float *gradOutput, float *gradWeight, float *gradBias,
        int64_t batch, int64_t nIn, int64_t inH, int64_t inW, int64_t nOut, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH, float scale) {
    API_BEGIN(s);
    REQUIRE(input && gradOutput && gradWeight, "SpatialFullConvolution_accGradParameters: null tensor");
    ConvGeom g = fullconv_geom(batch, nIn, inH, inW, nOut, kW, kH, dW, dH, padW, padH, adjW, adjH);
    if (check_geom(g)) return 1;
    if (gradBias && simt_bias_grad(s, gradOutput, gradBias, g.N, g.C, g.H * g.W, scale)) return 1;
    if (s->precision == CENN_BF16 && adjW == 0 && adjH == 0) {
        int rc = tc_conv_wgrad_nchw(s, gradOutput, input, gradWeight, g.N, g.C, g.H, g.W, g.O, kH, kW, dH, dW, padH, padW, scale, 0);
        if (rc <= 0) return -rc;
    }
    // conv wgrad with (x := gradOutput of the full conv, gy := its input) gives gw[o'=nIn][c'=nOut][u][v]
    return simt_wgrad(s, g, gradOutput, input, gradWeight, scale);
}

}  // extern "C"
