// pointwise.cu -- THNN activations and criteria on fp32 tensors, plus the fused entry points that
// replace the repo-local Lua compositions (MaskedMSECriterion.lua, gdl_criterion.lua, the blend at
// train.lua:377-400 / train_vid_weighted.lua:485-528, inpaint_utils.fillIn, optim.adam).
// All are HBM-bound: one pass, float4 where aligned, warp-shuffle + one atomic per block for the
// loss reductions (double accumulators, like THNN's accreal).
#include "common.cuh"
#include "map.cuh"

namespace {

template <typename F>
__global__ void __launch_bounds__(256) unary_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    int64_t n4 = al ? n / 4 : 0;
    for (int64_t j = i; j < n4; j += st) {
        float4 v = reinterpret_cast<const float4 *>(x)[j];
        v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
        reinterpret_cast<float4 *>(y)[j] = v;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += st) y[j] = f(x[j]);
}
// z = f(a, b)
template <typename F>
__global__ void __launch_bounds__(256) binary_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ z, int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(z)) & 15) == 0;
    int64_t n4 = al ? n / 4 : 0;
    for (int64_t j = i; j < n4; j += st) {
        float4 p = reinterpret_cast<const float4 *>(a)[j], q = reinterpret_cast<const float4 *>(b)[j], r;
        r.x = f(p.x, q.x); r.y = f(p.y, q.y); r.z = f(p.z, q.z); r.w = f(p.w, q.w);
        reinterpret_cast<float4 *>(z)[j] = r;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += st) z[j] = f(a[j], b[j]);
}
// loss reduction: acc[slot] += sum_i f(x_i, t_i) (double)
template <typename F>
__global__ void __launch_bounds__(256) reduce2_kernel(const float *__restrict__ x, const float *__restrict__ t, int64_t n, double *__restrict__ acc, F f) {
    __shared__ double sh[32];
    double a = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) a += f(x[j], t[j]);
    a = block_sum(a, sh);
    if (threadIdx.x == 0) atomicAdd(acc, a);
}

#define UNARY(s, x, y, n, ...) do { if ((n) > 0) { unary_kernel<<<bw_grid(s, (n) / 4 + 1, 256), 256, 0, (s)->stream>>>(x, y, n, __VA_ARGS__); CK_LAUNCH(s); } } while (0)
#define BINARY(s, a, b, z, n, ...) do { if ((n) > 0) { binary_kernel<<<bw_grid(s, (n) / 4 + 1, 256), 256, 0, (s)->stream>>>(a, b, z, n, __VA_ARGS__); CK_LAUNCH(s); } } while (0)

// run a reduction into red[slot], bring it to the host (the criterion:forward sync point)
template <typename F>
int reduce_to_host(cenn_state *s, const float *x, const float *t, int64_t n, double scale, float *loss_host, F f) {
    CK(cudaMemsetAsync(s->red, 0, sizeof(double), s->stream));
    if (n > 0) { reduce2_kernel<<<bw_grid(s, n, 256, 4), 256, 0, s->stream>>>(x, t, n, s->red, f); CK_LAUNCH(s); }
    CK(cudaMemcpyAsync(s->red_host, s->red, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *loss_host = (float)(s->red_host[0] * scale);
    return 0;
}

__global__ void __launch_bounds__(256) masked_mse_kernel(const float *__restrict__ x, const float *__restrict__ t, const float *__restrict__ m,
        float *__restrict__ g, int64_t n, float mW, float two_over_n, double *__restrict__ acc) {
    __shared__ double sh[32];
    double a = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float w = m[j] * (1.f - mW) + mW, d = x[j] - t[j];
        a += (double)(w * d * d);
        if (g) g[j] = w * d * two_over_n;
    }
    a = block_sum(a, sh);
    if (threadIdx.x == 0 && acc) atomicAdd(acc, a);
}

// GDL with flat-index pairing.  One thread per element e of a plane; it owns the two
// "forward" terms whose flat index is k = e restricted to k < H*(W-1), and gathers the four gradient
// contributions that land on pixel (y, x) -- no atomics, no scatter.
__device__ __forceinline__ float sgnp(float z) { return z >= 0.f ? 1.f : -1.f; }
__device__ __forceinline__ void gdl_terms(const float *__restrict__ Y, const float *__restrict__ Yh, int k, int W, float &g12s, float &g34s, float &l) {
    // flat k: a = T[k div (W-1), k mod (W-1) (+1 for j2)], b = T[k div W (+1 row for j1), k mod W]
    int ar = k / (W - 1), ac = k - ar * (W - 1);
    int br = k / W, bc = k - br * W;
    float yi2 = Y[ar * W + ac], yi1 = Y[br * W + bc], hi2 = Yh[ar * W + ac], hi1 = Yh[br * W + bc];
    float yj2 = Y[ar * W + ac + 1], yj1 = Y[(br + 1) * W + bc], hj2 = Yh[ar * W + ac + 1], hj1 = Yh[(br + 1) * W + bc];
    float t12 = fabsf(yi2 - yi1) - fabsf(hi2 - hi1);
    float t34 = fabsf(yj2 - yj1) - fabsf(hj2 - hj1);
    l = fabsf(t12) + fabsf(t34);
    g12s = -sgnp(t12) * sgnp(hi2 - hi1);   // delta2 * n
    g34s = -sgnp(t34) * sgnp(hj2 - hj1);   // delta4 * n
}
__global__ void __launch_bounds__(256) gdl_kernel(const float *__restrict__ inp, const float *__restrict__ tgt, float *__restrict__ g,
        int64_t planes, int H, int W, float inv_n, double *__restrict__ acc) {
    __shared__ double sh[32];
    const int HW = H * W, NK = H * (W - 1);
    double a = 0.0;
    int64_t total = planes * HW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = idx / HW;
        int e = (int)(idx - p * HW);
        const float *Y = tgt + p * HW, *Yh = inp + p * HW;
        float d, q, l;
        if (e < NK) { gdl_terms(Y, Yh, e, W, d, q, l); a += (double)l; }
        if (g) {
            int y = e / W, x = e - y * W;
            float acc_g = 0.f;
            // +d2 at [H, W-1] position (y, x) with x < W-1  -> k = y*(W-1)+x
            if (x < W - 1) { gdl_terms(Y, Yh, y * (W - 1) + x, W, d, q, l); acc_g += d; }
            // -d2 at [H-1, W] position (y, x) with y < H-1  -> k = y*W+x
            if (y < H - 1) { gdl_terms(Y, Yh, y * W + x, W, d, q, l); acc_g -= d; }
            // +d4 at [H, W-1] padded left: pixel (y, x) with x >= 1 -> k = y*(W-1)+(x-1)
            if (x >= 1) { gdl_terms(Y, Yh, y * (W - 1) + x - 1, W, d, q, l); acc_g += q; }
            // -d4 at [H-1, W] padded top: pixel (y, x) with y >= 1 -> k = (y-1)*W+x
            if (y >= 1) { gdl_terms(Y, Yh, (y - 1) * W + x, W, d, q, l); acc_g -= q; }
            g[idx] = acc_g * inv_n;
        }
    }
    a = block_sum(a, sh);
    if (threadIdx.x == 0 && acc) atomicAdd(acc, a);
}

// df_dg = a*df_dg + Wm*2(x-t)/n with the overlapPred ring weight computed from indices; loss = sum (x-t)^2
__global__ void __launch_bounds__(256) blend_overlap_kernel(float *__restrict__ df, const float *__restrict__ x, const float *__restrict__ t,
        int64_t n, int H, int W, int ov, float a, float w_in, float w_ring, float two_over_n, double *__restrict__ acc) {
    __shared__ double sh[32];
    double l = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        int xx = (int)(j % W), yy = (int)((j / W) % H);
        bool inner = (yy >= ov && yy < H - ov && xx >= ov && xx < W - ov);
        float d = x[j] - t[j];
        l += (double)d * d;
        df[j] = df[j] * a + (inner ? w_in : w_ring) * (d * two_over_n);
    }
    l = block_sum(l, sh);
    if (threadIdx.x == 0 && acc) atomicAdd(acc, l);
}
__global__ void __launch_bounds__(256) blend_masked_kernel(float *__restrict__ df, const float *__restrict__ x, const float *__restrict__ t,
        float *__restrict__ mask, int64_t n, float a, float wtl2, float lambda, float wtgdl, float two_over_n, double *__restrict__ acc) {
    __shared__ double sh[32];
    double l = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float d = x[j] - t[j];
        l += (double)d * d;
        float g2 = d * two_over_n, w = 1.f;
        if (lambda != 0.f) { w = mask[j] * (1.f - lambda) + lambda; mask[j] = w; }
        float v = df[j] * a + wtl2 * (g2 * w);
        if (wtgdl != 0.f) v += wtgdl * g2;
        df[j] = v;
    }
    l = block_sum(l, sh);
    if (threadIdx.x == 0 && acc) atomicAdd(acc, l);
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ x, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
        int64_t n, float b1, float b2, float eps, float step) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    int64_t n4 = al ? n / 4 : 0;
    for (int64_t j = i; j < n4; j += st) {
        float4 X = reinterpret_cast<float4 *>(x)[j], G = reinterpret_cast<const float4 *>(g)[j];
        float4 M = reinterpret_cast<float4 *>(m)[j], V = reinterpret_cast<float4 *>(v)[j];
#define ADAM1(c) M.c = M.c * b1 + (1.f - b1) * G.c; V.c = V.c * b2 + (1.f - b2) * G.c * G.c; X.c -= step * M.c / (sqrtf(V.c) + eps);
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        reinterpret_cast<float4 *>(x)[j] = X; reinterpret_cast<float4 *>(m)[j] = M; reinterpret_cast<float4 *>(v)[j] = V;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += st) {
        float gg = g[j];
        float mm = m[j] * b1 + (1.f - b1) * gg, vv = v[j] * b2 + (1.f - b2) * gg * gg;
        m[j] = mm; v[j] = vv;
        x[j] -= step * mm / (sqrtf(vv) + eps);
    }
}

}  // namespace

extern "C" {

// ---- activations -------------------------------------------------------------
int cenn_LeakyReLU_updateOutput(cenn_state *s, const float *in, float *out, int64_t n, double negval, int inplace) {
    API_BEGIN(s); (void)inplace; float nv = (float)negval;
    UNARY(s, in, out, n, [nv] __device__(float v) { return v > 0.f ? v : v * nv; }); return 0;
}
int cenn_LeakyReLU_updateGradInput(cenn_state *s, const float *in, const float *gy, float *gx, int64_t n, double negval, int inplace) {
    API_BEGIN(s); (void)inplace; float nv = (float)negval;
    BINARY(s, in, gy, gx, n, [nv] __device__(float v, float g) { return v > 0.f ? g : g * nv; }); return 0;
}
int cenn_Threshold_updateOutput(cenn_state *s, const float *in, float *out, int64_t n, double threshold, double val, int inplace) {
    API_BEGIN(s); (void)inplace; float th = (float)threshold, vl = (float)val;
    UNARY(s, in, out, n, [th, vl] __device__(float v) { return v > th ? v : vl; }); return 0;
}
int cenn_Threshold_updateGradInput(cenn_state *s, const float *in, const float *gy, float *gx, int64_t n, double threshold, int inplace) {
    API_BEGIN(s); (void)inplace; float th = (float)threshold;
    BINARY(s, in, gy, gx, n, [th] __device__(float v, float g) { return v > th ? g : 0.f; }); return 0;
}
int cenn_Tanh_updateOutput(cenn_state *s, const float *in, float *out, int64_t n) {
    API_BEGIN(s); UNARY(s, in, out, n, [] __device__(float v) { return tanhf(v); }); return 0;
}
int cenn_Tanh_updateGradInput(cenn_state *s, const float *gy, float *gx, const float *out, int64_t n) {
    API_BEGIN(s); BINARY(s, gy, out, gx, n, [] __device__(float g, float y) { return g * (1.f - y * y); }); return 0;
}
int cenn_Sigmoid_updateOutput(cenn_state *s, const float *in, float *out, int64_t n) {
    API_BEGIN(s); UNARY(s, in, out, n, [] __device__(float v) { return 1.f / (1.f + expf(-v)); }); return 0;
}
int cenn_Sigmoid_updateGradInput(cenn_state *s, const float *gy, float *gx, const float *out, int64_t n) {
    API_BEGIN(s); BINARY(s, gy, out, gx, n, [] __device__(float g, float y) { return g * y * (1.f - y); }); return 0;
}
int cenn_Abs_updateOutput(cenn_state *s, const float *in, float *out, int64_t n) {
    API_BEGIN(s); UNARY(s, in, out, n, [] __device__(float v) { return fabsf(v); }); return 0;
}
int cenn_Abs_updateGradInput(cenn_state *s, const float *in, const float *gy, float *gx, int64_t n) {
    API_BEGIN(s); BINARY(s, in, gy, gx, n, [] __device__(float v, float g) { return v >= 0.f ? g : -g; }); return 0;
}
int cenn_Square_updateOutput(cenn_state *s, const float *in, float *out, int64_t n) {
    API_BEGIN(s); UNARY(s, in, out, n, [] __device__(float v) { return v * v; }); return 0;
}
int cenn_Square_updateGradInput(cenn_state *s, const float *in, const float *gy, float *gx, int64_t n) {
    API_BEGIN(s); BINARY(s, in, gy, gx, n, [] __device__(float v, float g) { return 2.f * g * v; }); return 0;
}

// ---- criteria ------------------------------------------------------------------
int cenn_BCECriterion_updateOutput(cenn_state *s, const float *in, const float *tg, int64_t n, int sizeAverage, float *loss) {
    API_BEGIN(s); REQUIRE(in && tg && loss && n > 0, "BCECriterion_updateOutput: bad arguments");
    return reduce_to_host(s, in, tg, n, sizeAverage ? -1.0 / (double)n : -1.0, loss,
        [] __device__(float x, float t) { return (double)(t * logf(x + 1e-12f) + (1.f - t) * logf(1.f - x + 1e-12f)); });
}
int cenn_BCECriterion_updateGradInput(cenn_state *s, const float *in, const float *tg, float *gx, int64_t n, int sizeAverage) {
    API_BEGIN(s); REQUIRE(in && tg && gx && n > 0, "BCECriterion_updateGradInput: bad arguments");
    float norm = sizeAverage ? 1.f / (float)n : 1.f;
    BINARY(s, in, tg, gx, n, [norm] __device__(float x, float t) { return -norm * (t - x) / ((1.f - x + 1e-12f) * (x + 1e-12f)); });
    return 0;
}
int cenn_MSECriterion_updateOutput(cenn_state *s, const float *in, const float *tg, int64_t n, int sizeAverage, float *loss) {
    API_BEGIN(s); REQUIRE(in && tg && loss && n > 0, "MSECriterion_updateOutput: bad arguments");
    return reduce_to_host(s, in, tg, n, sizeAverage ? 1.0 / (double)n : 1.0, loss,
        [] __device__(float x, float t) { double d = (double)x - (double)t; return d * d; });
}
int cenn_MSECriterion_updateGradInput(cenn_state *s, const float *in, const float *tg, float *gx, int64_t n, int sizeAverage) {
    API_BEGIN(s); REQUIRE(in && tg && gx && n > 0, "MSECriterion_updateGradInput: bad arguments");
    float norm = sizeAverage ? 2.f / (float)n : 2.f;
    BINARY(s, in, tg, gx, n, [norm] __device__(float x, float t) { return norm * (x - t); });
    return 0;
}
int cenn_AbsCriterion_updateOutput(cenn_state *s, const float *in, const float *tg, int64_t n, int sizeAverage, float *loss) {
    API_BEGIN(s); REQUIRE(in && tg && loss && n > 0, "AbsCriterion_updateOutput: bad arguments");
    return reduce_to_host(s, in, tg, n, sizeAverage ? 1.0 / (double)n : 1.0, loss,
        [] __device__(float x, float t) { return fabs((double)x - (double)t); });
}
int cenn_AbsCriterion_updateGradInput(cenn_state *s, const float *in, const float *tg, float *gx, int64_t n, int sizeAverage) {
    API_BEGIN(s); REQUIRE(in && tg && gx && n > 0, "AbsCriterion_updateGradInput: bad arguments");
    float norm = sizeAverage ? 1.f / (float)n : 1.f;
    BINARY(s, in, tg, gx, n, [norm] __device__(float x, float t) { return (x - t) >= 0.f ? norm : -norm; });
    return 0;
}

// ---- fused replacements of the Lua compositions --------------------------------------------
static int finish_loss(cenn_state *s, double scale, float *loss_host) {
    if (!loss_host) return 0;
    CK(cudaMemcpyAsync(s->red_host, s->red, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *loss_host = (float)(s->red_host[0] * scale);
    return 0;
}

int cenn_MaskedMSECriterion_forward_backward(cenn_state *s, const float *in, const float *tg, const float *mask, float *gx, int64_t n,
        double mWeight, float *loss_host) {
    API_BEGIN(s); REQUIRE(in && tg && mask && n > 0, "MaskedMSECriterion: bad arguments (setMask not called?)");
    CK(cudaMemsetAsync(s->red, 0, sizeof(double), s->stream));
    masked_mse_kernel<<<bw_grid(s, n, 256, 4), 256, 0, s->stream>>>(in, tg, mask, gx, n, (float)mWeight, (float)(2.0 / (double)n), loss_host ? s->red : nullptr);
    CK_LAUNCH(s);
    return finish_loss(s, 1.0 / (double)n, loss_host);
}

int cenn_GDLCriterion_forward_backward(cenn_state *s, const float *in, const float *tg, float *gx, int64_t batch, int64_t C, int64_t H, int64_t W,
        float *loss_host) {
    API_BEGIN(s); REQUIRE(in && tg && batch > 0 && C > 0, "GDLCriterion: bad arguments");
    REQUIRE(H == W, "GDLCriterion: inconsistent tensor size (needs square maps, got %lld x %lld)", (long long)H, (long long)W);
    REQUIRE(H >= 2, "GDLCriterion: maps must be at least 2x2");
    int64_t planes = batch * C;
    double n = (double)planes * (double)H * (double)(W - 1);
    CK(cudaMemsetAsync(s->red, 0, sizeof(double), s->stream));
    gdl_kernel<<<bw_grid(s, planes * H * W, 256, 4), 256, 0, s->stream>>>(in, tg, gx, planes, (int)H, (int)W, (float)(1.0 / n), loss_host ? s->red : nullptr);
    CK_LAUNCH(s);
    return finish_loss(s, 1.0 / n, loss_host);
}

int cenn_WeightedMSEBlend_overlap(cenn_state *s, float *df, const float *in, const float *tg, int64_t batch, int64_t C, int64_t H, int64_t W,
        double wtl2, int overlapPred, float *loss_host) {
    API_BEGIN(s); REQUIRE(df && in && tg, "WeightedMSEBlend_overlap: null tensor");
    REQUIRE(overlapPred >= 0 && 2 * overlapPred <= H && 2 * overlapPred <= W, "WeightedMSEBlend_overlap: overlapPred too large");
    int64_t n = batch * C * H * W;
    float a = (wtl2 > 0 && wtl2 < 1) ? (float)(1.0 - wtl2) : 1.f;
    float w_in = (float)wtl2, w_ring = overlapPred > 0 ? (float)(10.0 * wtl2) : (float)wtl2;
    CK(cudaMemsetAsync(s->red, 0, sizeof(double), s->stream));
    blend_overlap_kernel<<<bw_grid(s, n, 256, 4), 256, 0, s->stream>>>(df, in, tg, n, (int)H, (int)W, overlapPred, a, w_in, w_ring,
                                                                     (float)(2.0 / (double)n), loss_host ? s->red : nullptr);
    CK_LAUNCH(s);
    return finish_loss(s, 1.0 / (double)n, loss_host);
}

int cenn_WeightedMSEBlend_masked(cenn_state *s, float *df, const float *in, const float *tg, float *mask, int64_t n, double wtl2,
        double weight_nomask, double wtgdl, float *loss_host) {
    API_BEGIN(s); REQUIRE(df && in && tg && n > 0, "WeightedMSEBlend_masked: bad arguments");
    REQUIRE(mask || weight_nomask == 0, "WeightedMSEBlend_masked: mask required when weight_nomask != 0");
    float a = (wtl2 > 0 && wtl2 < 1) ? (float)(1.0 - wtl2) : 1.f;
    CK(cudaMemsetAsync(s->red, 0, sizeof(double), s->stream));
    blend_masked_kernel<<<bw_grid(s, n, 256, 4), 256, 0, s->stream>>>(df, in, tg, mask, n, a, (float)wtl2, (float)weight_nomask, (float)wtgdl,
                                                                    (float)(2.0 / (double)n), loss_host ? s->red : nullptr);
    CK_LAUNCH(s);
    return finish_loss(s, 1.0 / (double)n, loss_host);
}

int cenn_MaskComposite(cenn_state *s, float *dst, const float *mask, const float *src, int64_t n) {
    API_BEGIN(s); REQUIRE(dst && mask && src, "MaskComposite: null tensor");
    LAUNCH_MAP3(s, dst, mask, src, n, [] __device__(float d, float m, float sv) { return m != 0.f ? sv : d; });
    return 0;
}

int cenn_AdamFlat(cenn_state *s, float *x, const float *g, float *m, float *v, int64_t n, double lr, double beta1, double beta2, double eps, int64_t t) {
    API_BEGIN(s); REQUIRE(x && g && m && v && t >= 1, "AdamFlat: bad arguments");
    double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
    float step = (float)(lr * sqrt(bc2) / bc1);
    if (n > 0) { adam_kernel<<<bw_grid(s, n / 4 + 1, 256), 256, 0, s->stream>>>(x, g, m, v, n, (float)beta1, (float)beta2, (float)eps, step); CK_LAUNCH(s); }
    return 0;
}

}  // extern "C"
