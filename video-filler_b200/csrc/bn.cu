// bn.cu -- THNN BatchNormalization_updateOutput / _backward on fp32 [batch, C, spatial] tensors
// (nn.SpatialBatchNormalization, train.lua:79; semantics SURVEY 9.3).  Bandwidth kernels:
//   stats   : per-channel sum / sum-of-squares, double accumulation like THNN's accreal (1 read)
//   finalize: mean, invstd, running stats (momentum, unbiased variance)
//   apply   : y = (x - mean) * invstd * gamma + beta, float4 vectorised (1 read + 1 write)
//   backward: sums (2 reads) then dx (2 reads + 1 write); gradWeight/gradBias accumulate.
#include "common.cuh"

namespace {

// grid (C, S): block (c, sidx) reduces rows n = sidx, sidx+S, ... of channel c
__global__ void __launch_bounds__(256) bn_stats_kernel(const float *__restrict__ x, double *__restrict__ acc, int batch, int C, int64_t sp) {
    __shared__ double sh[32];
    int c = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int n = blockIdx.y; n < batch; n += gridDim.y) {
        const float *p = x + ((int64_t)n * C + c) * sp;
        if ((sp & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const float4 *p4 = reinterpret_cast<const float4 *>(p);
            for (int64_t i = threadIdx.x; i < sp / 4; i += blockDim.x) {
                float4 v = __ldg(p4 + i);
                s1 += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
                s2 += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            }
        } else {
            for (int64_t i = threadIdx.x; i < sp; i += blockDim.x) { double v = p[i]; s1 += v; s2 += v * v; }
        }
    }
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) { atomicAdd(acc + c, s1); atomicAdd(acc + C + c, s2); }
}

__global__ void bn_finalize_kernel(const double *__restrict__ acc, float *__restrict__ running_mean, float *__restrict__ running_var,
        float *__restrict__ save_mean, float *__restrict__ save_std, int C, double n, double momentum, double eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mean = acc[c] / n;
    double S = acc[C + c] - acc[c] * mean;   // sum (x-mean)^2
    if (S < 0) S = 0;
    double invstd = 1.0 / sqrt(S / n + eps);
    save_mean[c] = (float)mean;
    save_std[c] = (float)invstd;
    running_mean[c] = (float)(momentum * mean + (1.0 - momentum) * (double)running_mean[c]);
    double unbiased = S / (n - 1.0);   // n == 1 -> inf, as in the reference
    running_var[c] = (float)(momentum * unbiased + (1.0 - momentum) * (double)running_var[c]);
}

__global__ void bn_eval_stats_kernel(const float *__restrict__ running_mean, const float *__restrict__ running_var,
        float *__restrict__ mean, float *__restrict__ invstd, int C, double eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = running_mean[c];
    invstd[c] = (float)(1.0 / sqrt((double)running_var[c] + eps));
}

// grid-stride over (n*C + c) rows x spatial; one row segment per block iteration keeps the channel uniform
__global__ void __launch_bounds__(256) bn_apply_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ gamma,
        const float *__restrict__ beta, const float *__restrict__ mean, const float *__restrict__ invstd, int64_t rows, int C, int64_t sp) {
    if (sp >= 64) {
        for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
            int c = (int)(r % C);
            float m = mean[c], a = invstd[c] * (gamma ? gamma[c] : 1.f), b = beta ? beta[c] : 0.f;
            const float *p = x + r * sp; float *q = y + r * sp;
            if ((sp & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) & 15) == 0) {
                for (int64_t i = threadIdx.x; i < sp / 4; i += blockDim.x) {
                    float4 v = __ldg(reinterpret_cast<const float4 *>(p) + i);
                    v.x = (v.x - m) * a + b; v.y = (v.y - m) * a + b; v.z = (v.z - m) * a + b; v.w = (v.w - m) * a + b;
                    reinterpret_cast<float4 *>(q)[i] = v;
                }
            } else {
                for (int64_t i = threadIdx.x; i < sp; i += blockDim.x) q[i] = (p[i] - m) * a + b;
            }
        }
    } else {
        int64_t total = rows * sp;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            int c = (int)((i / sp) % C);
            float a = invstd[c] * (gamma ? gamma[c] : 1.f);
            y[i] = (x[i] - mean[c]) * a + (beta ? beta[c] : 0.f);
        }
    }
}

__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float *__restrict__ x, const float *__restrict__ gy, const float *__restrict__ mean,
        double *__restrict__ acc, int batch, int C, int64_t sp) {
    __shared__ double sh[32];
    int c = blockIdx.x;
    double m = mean[c];
    double s = 0.0, d = 0.0;
    for (int n = blockIdx.y; n < batch; n += gridDim.y) {
        const float *p = x + ((int64_t)n * C + c) * sp;
        const float *g = gy + ((int64_t)n * C + c) * sp;
        for (int64_t i = threadIdx.x; i < sp; i += blockDim.x) { double gv = g[i]; s += gv; d += ((double)p[i] - m) * gv; }
    }
    s = block_sum(s, sh);
    d = block_sum(d, sh);
    if (threadIdx.x == 0) { atomicAdd(acc + c, s); atomicAdd(acc + C + c, d); }
}

__global__ void bn_bwd_param_kernel(const double *__restrict__ acc, const float *__restrict__ invstd, float *__restrict__ gw, float *__restrict__ gb,
        int C, double scale) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (gw) gw[c] += (float)(scale * acc[C + c] * (double)invstd[c]);
    if (gb) gb[c] += (float)(scale * acc[c]);
}

__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(const float *__restrict__ x, const float *__restrict__ gy, float *__restrict__ gx,
        const float *__restrict__ gamma, const float *__restrict__ mean, const float *__restrict__ invstd, const double *__restrict__ acc,
        int64_t total, int C, int64_t sp, double n, int train) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)((i / sp) % C);
        float is = invstd[c], g = gamma ? gamma[c] : 1.f;
        if (train) {
            float sN = (float)(acc[c] / n);
            float k = (float)(acc[C + c] / n) * is * is;
            gx[i] = (gy[i] - sN - (x[i] - mean[c]) * k) * is * g;
        } else {
            gx[i] = gy[i] * is * g;
        }
    }
}

}  // namespace

extern "C" {

int cenn_BatchNormalization_updateOutput(cenn_state *s, const float *input, float *output, const float *weight, const float *bias,
        float *runningMean, float *runningVar, float *saveMean, float *saveStd, int64_t batch, int64_t C, int64_t spatial,
        int train, double momentum, double eps) {
    API_BEGIN(s);
    REQUIRE(input && output && runningMean && runningVar && saveMean && saveStd, "BatchNormalization_updateOutput: null tensor");
    REQUIRE(batch > 0 && C > 0 && spatial > 0, "BatchNormalization: empty input");
    if (train) {
        double *acc = (double *)cenn_workspace(s, 2 * C * sizeof(double));
        if (!acc) return 1;
        CK(cudaMemsetAsync(acc, 0, 2 * C * sizeof(double), s->stream));
        int S = (int)((4 * (int64_t)s->sm_count + C - 1) / C);
        if (S > batch) S = (int)batch;
        if (S < 1) S = 1;
        bn_stats_kernel<<<dim3((unsigned)C, (unsigned)S), 256, 0, s->stream>>>(input, acc, (int)batch, (int)C, spatial);
        CK_LAUNCH(s);
        bn_finalize_kernel<<<(unsigned)((C + 127) / 128), 128, 0, s->stream>>>(acc, runningMean, runningVar, saveMean, saveStd, (int)C,
                                                                            (double)(batch * spatial), momentum, eps);
        CK_LAUNCH(s);
    } else {
        bn_eval_stats_kernel<<<(unsigned)((C + 127) / 128), 128, 0, s->stream>>>(runningMean, runningVar, saveMean, saveStd, (int)C, eps);
        CK_LAUNCH(s);
    }
    int64_t rows = batch * C;
    int grid = spatial >= 64 ? (int)(rows < (int64_t)s->sm_count * 8 ? rows : (int64_t)s->sm_count * 8) : bw_grid(s, rows * spatial, 256);
    bn_apply_kernel<<<grid, 256, 0, s->stream>>>(input, output, weight, bias, saveMean, saveStd, rows, (int)C, spatial);
    CK_LAUNCH(s);
    return 0;
}

int cenn_BatchNormalization_backward(cenn_state *s, const float *input, const float *gradOutput, float *gradInput, float *gradWeight,
        float *gradBias, const float *weight, const float *runningMean, const float *runningVar, const float *saveMean,
        const float *saveStd, int64_t batch, int64_t C, int64_t spatial, int train, double scale, double eps) {
    API_BEGIN(s);
    REQUIRE(input && gradOutput, "BatchNormalization_backward: null tensor");
    REQUIRE(batch > 0 && C > 0 && spatial > 0, "BatchNormalization: empty input");
    double *acc = (double *)cenn_workspace(s, 2 * C * sizeof(double) + 2 * C * sizeof(float));
    if (!acc) return 1;
    float *mean_e = (float *)(acc + 2 * C), *invstd_e = mean_e + C;
    const float *mean = saveMean, *invstd = saveStd;
    if (!train) {
        REQUIRE(runningMean && runningVar, "BatchNormalization_backward(eval): null running stats");
        bn_eval_stats_kernel<<<(unsigned)((C + 127) / 128), 128, 0, s->stream>>>(runningMean, runningVar, mean_e, invstd_e, (int)C, eps);
        CK_LAUNCH(s);
        mean = mean_e; invstd = invstd_e;
    } else {
        REQUIRE(saveMean && saveStd, "BatchNormalization_backward(train): null save_mean/save_std");
    }
    CK(cudaMemsetAsync(acc, 0, 2 * C * sizeof(double), s->stream));
    int S = (int)((4 * (int64_t)s->sm_count + C - 1) / C);
    if (S > batch) S = (int)batch;
    if (S < 1) S = 1;
    bn_bwd_stats_kernel<<<dim3((unsigned)C, (unsigned)S), 256, 0, s->stream>>>(input, gradOutput, mean, acc, (int)batch, (int)C, spatial);
    CK_LAUNCH(s);
    if (gradWeight || gradBias) {
        bn_bwd_param_kernel<<<(unsigned)((C + 127) / 128), 128, 0, s->stream>>>(acc, invstd, gradWeight, gradBias, (int)C, scale);
        CK_LAUNCH(s);
    }
    if (gradInput) {
        int64_t total = batch * C * spatial;
        bn_bwd_dx_kernel<<<bw_grid(s, total, 256), 256, 0, s->stream>>>(input, gradOutput, gradInput, weight, mean, invstd, acc, total,
                                                                        (int)C, spatial, (double)(batch * spatial), train);
        CK_LAUNCH(s);
    }
    return 0;
}

}  // extern "C"
