// nhwc.cuh -- HBM-bound kernels of the whole-step executor on NHWC bf16 tensors (channels padded to Cp).
// 16-byte vector accesses (8 bf16) along the channel axis, per-channel reductions via registers ->
// shared memory -> one fp32 atomic per channel per CTA, grids sized in multiples of the SM count.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "conv_tc.h"

namespace nhwc {

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };

struct bf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ float act_fwd(float z, int act, float negval) {
    if (act == ACT_LEAKY) return z > 0.f ? z : z * negval;
    if (act == ACT_RELU) return z > 0.f ? z : 0.f;
    if (act == ACT_TANH) return tanhf(z);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-z));
    return z;
}
// derivative expressed through the stored activation output a (what the in-place Torch modules use)
__device__ __forceinline__ float act_bwd(float a, int act, float negval) {
    if (act == ACT_LEAKY) return a > 0.f ? 1.f : negval;
    if (act == ACT_RELU) return a > 0.f ? 1.f : 0.f;
    if (act == ACT_TANH) return 1.f - a * a;
    if (act == ACT_SIGMOID) return a * (1.f - a);
    return 1.f;
}

// ---------------------------------------------------------------- input conversion / im2col
// NCHW (float or uint8) -> NHWC bf16 with Cp channels (pad lanes zero); one thread per pixel
template <typename T>
__global__ void __launch_bounds__(256) to_nhwc_kernel(const T *__restrict__ src, bf16 *__restrict__ dst, int N, int C, int HW, int Cp) {
    pdl_trigger(); pdl_wait();
    int64_t total = (int64_t)N * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(i / HW), pix = (int)(i - (int64_t)n * HW);
        bf16 *o = dst + i * Cp;
        if ((Cp & 7) == 0) {                 // 16-byte stores (12 stacked channels -> Cp = 16)
            for (int c0 = 0; c0 < Cp; c0 += 8) {
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = c0 + k;
                    float v = c < C ? (float)src[((int64_t)n * C + c) * HW + pix] : 0.f;
                    if (sizeof(T) == 1) v = v != 0.f ? 1.f : 0.f;
                    f[k] = v;
                }
                reinterpret_cast<uint4 *>(o)[c0 >> 3] = pack8(f);
            }
            continue;
        }
        for (int c = 0; c < Cp; ++c) {
            float v = c < C ? (float)src[((int64_t)n * C + c) * HW + pix] : 0.f;
            if (sizeof(T) == 1) v = v != 0.f ? 1.f : 0.f;
            o[c] = __float2bfloat16(v);
        }
    }
}
// fast path for 3-channel images (Cp == 4, HW % 4 == 0): one thread converts 4 consecutive pixels -- float4 loads per
// channel plane, two 16-byte stores
__global__ void __launch_bounds__(256) to_nhwc4_kernel(const float *__restrict__ src, bf16 *__restrict__ dst, int N, int C, int HW) {
    pdl_trigger(); pdl_wait();
    const int64_t total = (int64_t)N * HW / 4;
    const int q = HW / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / q), p4 = (int)(i - (int64_t)n * q);
        float4 ch[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ch[c] = c < C ? __ldg(reinterpret_cast<const float4 *>(src + ((int64_t)n * C + c) * HW) + p4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float px[4][4] = {{ch[0].x, ch[1].x, ch[2].x, ch[3].x}, {ch[0].y, ch[1].y, ch[2].y, ch[3].y}, {ch[0].z, ch[1].z, ch[2].z, ch[3].z}, {ch[0].w, ch[1].w, ch[2].w, ch[3].w}};
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(px[k][0], px[k][1]), h1 = __floats2bfloat162_rn(px[k][2], px[k][3]);
            w[2 * k] = *reinterpret_cast<uint32_t *>(&h0); w[2 * k + 1] = *reinterpret_cast<uint32_t *>(&h1);
        }
        uint4 *o = reinterpret_cast<uint4 *>(dst + i * 16);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]); o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}
// NHWC bf16 -> NCHW fp32 (results / debugging)
__global__ void __launch_bounds__(256) to_nchw_kernel(const bf16 *__restrict__ src, float *__restrict__ dst, int N, int C, int HW, int Cp) {
    int64_t total = (int64_t)N * C * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int pix = (int)(i % HW);
        int64_t t = i / HW;
        int c = (int)(t % C), n = (int)(t / C);
        dst[i] = __bfloat162float(src[((int64_t)n * HW + pix) * Cp + c]);
    }
}
// explicit im2col of a thin tensor L [N,2h,2w,Cp] (Cp = 4 or 16): col[pix][(tap, c)] for the 4x4/s2/p1 window.
// One thread moves one window ROW (4 taps = 4 consecutive input pixels): 8-byte loads (the row starts at an odd pixel),
// 16-byte stores (the destination row segment is 32-byte aligned).
template <int CP>
__global__ void __launch_bounds__(256) im2col_kernel(const bf16 *__restrict__ L, bf16 *__restrict__ col, int N, int h, int w) {
    pdl_trigger(); pdl_wait();
    constexpr int V8 = CP / 4;                                   // 8-byte vectors per pixel
    const int64_t total = (int64_t)N * h * w * 4;                // (pixel, window row u)
    const int H2 = 2 * h, W2 = 2 * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int u = (int)(i & 3);
        const int64_t pix = i >> 2;
        const int ox = (int)(pix % w), oy = (int)((pix / w) % h), n = (int)(pix / ((int64_t)w * h));
        const int iy = 2 * oy - 1 + u, ix0 = 2 * ox - 1;
        uint2 v[4 * V8];
        const bool row_ok = (unsigned)iy < (unsigned)H2;
        const uint2 *src = reinterpret_cast<const uint2 *>(L + (((int64_t)n * H2 + (row_ok ? iy : 0)) * W2) * CP);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int ix = ix0 + t;
            const bool ok = row_ok && (unsigned)ix < (unsigned)W2;
#pragma unroll
            for (int k = 0; k < V8; ++k) v[t * V8 + k] = ok ? __ldg(src + (int64_t)ix * V8 + k) : make_uint2(0u, 0u);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(col + (pix * 16 + u * 4) * CP);
#pragma unroll
        for (int k = 0; k < 2 * V8; ++k) dst[k] = make_uint4(v[2 * k].x, v[2 * k].y, v[2 * k + 1].x, v[2 * k + 1].y);
    }
}
// Cp == 16 (12 stacked channels): one thread per 16-byte chunk of the col row -- (tap, channel half) in K order -- so that
// both the load (half a pixel, 16-byte aligned) and the store (consecutive threads, consecutive 16 bytes) are full vectors.
__global__ void __launch_bounds__(256) im2col16_kernel(const bf16 *__restrict__ L, bf16 *__restrict__ col, int N, int h, int w) {
    pdl_trigger(); pdl_wait();
    const int64_t total = (int64_t)N * h * w * 32;               // (pixel, tap, half)
    const int H2 = 2 * h, W2 = 2 * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i & 31), tap = r >> 1, half = r & 1;
        const int64_t pix = i >> 5;
        const int ox = (int)(pix % w), oy = (int)((pix / w) % h), n = (int)(pix / ((int64_t)w * h));
        const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)iy < (unsigned)H2 && (unsigned)ix < (unsigned)W2)
            v = __ldg(reinterpret_cast<const uint4 *>(L + (((int64_t)n * H2 + iy) * W2 + ix) * 16) + half);
        reinterpret_cast<uint4 *>(col)[i] = v;
    }
}

// ---------------------------------------------------------------- batch norm
// sums -> mean / invstd / running stats / (scale, shift); `fold` column groups are summed (G1's GEMM epilogue
// accumulates per (tap, channel) column).  Zeroes the accumulators for the next use.
__device__ __forceinline__ void bn_finalize_channel(float *__restrict__ stats, int stats_stride, int fold, int fold_stride, const float *__restrict__ gamma,
        const float *__restrict__ beta, float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ mean,
        float *__restrict__ invstd, float *__restrict__ scale, float *__restrict__ shift, int c, double n, double momentum, double eps, int update_running) {
    double s1 = 0.0, s2 = 0.0;
    for (int f = 0; f < fold; ++f) {
        s1 += (double)stats[f * fold_stride + c]; s2 += (double)stats[stats_stride + f * fold_stride + c];
        stats[f * fold_stride + c] = 0.f; stats[stats_stride + f * fold_stride + c] = 0.f;
    }
    double m = s1 / n;
    double S = s2 - s1 * m;
    if (S < 0) S = 0;
    double is = 1.0 / sqrt(S / n + eps);
    mean[c] = (float)m; invstd[c] = (float)is;
    float sc = (float)(is * (double)gamma[c]);
    scale[c] = sc; shift[c] = beta[c] - (float)m * sc;
    if (update_running) {
        running_mean[c] = (float)(momentum * m + (1.0 - momentum) * (double)running_mean[c]);
        running_var[c] = (float)(momentum * (S / (n - 1.0)) + (1.0 - momentum) * (double)running_var[c]);
    }
}
__global__ void bn_finalize_kernel(float *__restrict__ stats, int stats_stride, int fold, int fold_stride, const float *__restrict__ gamma,
        const float *__restrict__ beta, float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ mean,
        float *__restrict__ invstd, float *__restrict__ scale, float *__restrict__ shift, int C, double n, double momentum, double eps,
        int update_running) {
    pdl_trigger(); pdl_wait();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    bn_finalize_channel(stats, stats_stride, fold, fold_stride, gamma, beta, running_mean, running_var, mean, invstd, scale, shift, c, n, momentum, eps, update_running);
}

// ---------------------------------------------------------------- one-shot all-reduce over NVLink peer memory
// buf[0..n) <- sum over ranks, in place, by ONE CTA: push the local values into every peer's mailbox (parity slot of
// the exchange counter, row of this rank) over NVLink and add the peers' payloads from the LOCAL mailbox as soon as
// they carry this exchange's tag.  All ranks add in rank order, so replicas get bit-identical sums.  Two parity slots
// suffice: a rank pushes exchange e+2 only after it received every peer's e+1 words, which a peer sends after it has
// finished reading exchange e.
// Replaces a NCCL all-reduce launch (~10-25 us at 8 ranks) by ~3 us inside the consumer kernel.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p) {
    float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v;
}
// LL-style payload: every float travels as an 8-byte {value, exchange tag} word written with ONE store, so a reader that
// sees the tag also sees the value -- no flag, no system fence, one NVLink round trip after the slowest peer has written.
__device__ __forceinline__ void st_ll(uint2 *p, float v, unsigned int tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2 *p) {
    uint2 v; asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v;
}
__device__ void xr_sum_inplace(float *__restrict__ buf, int n, const XrCtx &x) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const unsigned long long e = *x.epoch + 1;
    const unsigned int tag = (unsigned int)e;           // never 0 within 2^32 exchanges; mailboxes start zeroed
    // mailbox of rank r: [parity][source rank][XR_MAXF] words.  PUSH: this rank stores its values into its own row of
    // EVERY rank's mailbox (remote stores are fire-and-forget: one NVLink crossing), then polls its LOCAL mailbox --
    // a poll costs an L2 hit instead of an NVLink round trip, and the wait after the slowest peer has written is one
    // crossing, not a round trip plus the poll period.
    const size_t slot = (size_t)(e & 1) * x.world * XR_MAXF;
    const size_t mine = slot + (size_t)x.rank * XR_MAXF;
    for (int i = tid; i < n; i += nthr) {
        const float v = buf[i];
#pragma unroll
        for (int r = 0; r < XR_MAX_WORLD; ++r)
            if (r < x.world && (x.push ? r != x.rank : r == x.rank)) st_ll(reinterpret_cast<uint2 *>(x.data[r]) + mine + i, v, tag);
    }
    // pull mode (CENN_XR_PULL=1, the round-1 protocol, kept for A/B runs): publish locally, poll the peers' memory
    const uint2 *local = reinterpret_cast<const uint2 *>(x.data[x.rank]) + slot;
#define XR_SRC(r) ((x.push ? local : reinterpret_cast<const uint2 *>(x.data[r]) + slot) + (size_t)(r) * XR_MAXF)
    for (int i = tid; i < n; i += nthr) {
        uint2 v[XR_MAX_WORLD];
        unsigned int pending = 0;
#pragma unroll
        for (int r = 0; r < XR_MAX_WORLD; ++r)
            if (r < x.world && r != x.rank) { v[r] = ld_ll(XR_SRC(r) + i); if (v[r].y != tag) pending |= 1u << r; }
        long long t0 = clock64();
        while (pending) {
#pragma unroll
            for (int r = 0; r < XR_MAX_WORLD; ++r)
                if (pending & (1u << r)) { v[r] = ld_ll(XR_SRC(r) + i); if (v[r].y == tag) pending &= ~(1u << r); }
            if (clock64() - t0 > x.timeout_cycles) { printf("cenn: peer exchange timeout (rank %d, exchange %llu, pending mask %x)\n", x.rank, e, pending); __trap(); }
        }
        float acc = 0.f;                                 // rank order: every replica adds in the same order -> bit-identical sums
#pragma unroll
        for (int r = 0; r < XR_MAX_WORLD; ++r) if (r < x.world) acc += (r == x.rank) ? buf[i] : __uint_as_float(v[r].x);
        buf[i] = acc;
    }
#undef XR_SRC
    __syncthreads();
    if (tid == 0) *x.epoch = e;
}
// BN forward statistics: exchange + finalise in one single-CTA kernel (data parallel)
__global__ void __launch_bounds__(1024) bn_finalize_xr_kernel(const XrCtx x, float *__restrict__ stats, int stats_stride, int fold, int fold_stride, float *__restrict__ cmp,
        const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ mean,
        float *__restrict__ invstd, float *__restrict__ scale, float *__restrict__ shift, int C, int Cp, double n, double momentum, double eps) {
    pdl_wait();      // NO early trigger: this kernel spins on its peers; a successor launched early would hold SM slots while it waits (cross-rank deadlock)
    // fold the column groups (G1: 16 taps per channel) locally first: the exchange carries 2*Cp floats
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
        float s1 = 0.f, s2 = 0.f;
        if (c < C)
            for (int f = 0; f < fold; ++f) {
                s1 += stats[f * fold_stride + c]; s2 += stats[stats_stride + f * fold_stride + c];
                stats[f * fold_stride + c] = 0.f; stats[stats_stride + f * fold_stride + c] = 0.f;
            }
        cmp[c] = s1; cmp[Cp + c] = s2;
    }
    __syncthreads();
    xr_sum_inplace(cmp, 2 * Cp, x);
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        bn_finalize_channel(cmp, Cp, 1, 0, gamma, beta, running_mean, running_var, mean, invstd, scale, shift, c, n, momentum, eps, 1);
}
// BN backward sums [2][Cp] (already folded over this rank's CTAs): exchange + coefficients + affine gradients
__global__ void __launch_bounds__(1024) bn_bwd_coef_xr_kernel(const XrCtx x, float *__restrict__ sums, int Cp, const float *__restrict__ gamma, const float *__restrict__ invstd,
        const float *__restrict__ mean, float *__restrict__ coef, float *__restrict__ ggamma, float *__restrict__ gbeta, int C, double n, float grad_scale, int zero_after) {
    pdl_wait();      // NO early trigger: this kernel spins on its peers; a successor launched early would hold SM slots while it waits (cross-rank deadlock)
    xr_sum_inplace(sums, 2 * Cp, x);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double s = sums[c], d = sums[Cp + c];
        const double is = invstd[c], A = is * (double)gamma[c], k1 = is * is * d / n;
        coef[c] = (float)A; coef[C + c] = (float)(A * k1); coef[2 * C + c] = (float)(((double)mean[c] * k1 - s / n) * A);
        if (ggamma) ggamma[c] += (float)(d * is) * grad_scale;
        if (gbeta) gbeta[c] += (float)s * grad_scale;
    }
    if (zero_after) {                      // `sums` is the accumulator bn_bwd_reduce2_kernel<ACT, true> adds into: ready for its next use
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * Cp; i += blockDim.x) sums[i] = 0.f;
    }
}
__global__ void bn_eval_coef_kernel(const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ running_mean,
        const float *__restrict__ running_var, float *__restrict__ scale, float *__restrict__ shift, int C, double eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float sc = (float)((double)gamma[c] / sqrt((double)running_var[c] + eps));
    scale[c] = sc; shift[c] = beta[c] - running_mean[c] * sc;
}
// a = act(y * scale[c] + shift[c]); vectors of 8 channels
__global__ void __launch_bounds__(256) bn_apply_act_kernel(const bf16 *__restrict__ y, bf16 *__restrict__ a, const float *__restrict__ scale,
        const float *__restrict__ shift, int64_t nvec, int vec_per_pix, int act, float negval) {
    pdl_trigger(); pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        int c0 = (int)(i % vec_per_pix) * 8;
        float f[8];
        unpack8(reinterpret_cast<const uint4 *>(y)[i], f);
        float4 s0 = *reinterpret_cast<const float4 *>(scale + c0), s1 = *reinterpret_cast<const float4 *>(scale + c0 + 4);
        float4 h0 = *reinterpret_cast<const float4 *>(shift + c0), h1 = *reinterpret_cast<const float4 *>(shift + c0 + 4);
        float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = act_fwd(f[k] * sc[k] + sh[k], act, negval);
        reinterpret_cast<uint4 *>(a)[i] = pack8(f);
    }
}

// bn_finalize + bn_apply_act in ONE launch (training mode, C <= 1024): every CTA derives scale / shift of all channels from
// the epilogue sums into shared memory (a few hundred fp64 operations per thread), CTA 0 also publishes mean / invstd /
// scale / shift / running statistics for the backward pass, and the last CTA to have read the sums zeroes them for the
// next use.  Removes one dependent launch (~5 us on the critical path) per BN layer and sweep.
__global__ void __launch_bounds__(256) bn_finalize_apply_act_kernel(float *__restrict__ stats, int stats_stride, int fold, int fold_stride,
        const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ running_mean, float *__restrict__ running_var,
        float *__restrict__ mean, float *__restrict__ invstd, float *__restrict__ scale_g, float *__restrict__ shift_g, int C, int Cp, double n,
        double momentum, double eps, const bf16 *__restrict__ y, bf16 *__restrict__ a, int64_t nvec, int vec_per_pix, int act, float negval,
        unsigned int *__restrict__ done_counter) {
    pdl_trigger(); pdl_wait();
    extern __shared__ float sm_ss[];                   // [2][Cp]
    __shared__ int is_last;
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
        float sc = 0.f, sh = 0.f;
        if (c < C) {
            double s1 = 0.0, s2 = 0.0;
            for (int f = 0; f < fold; ++f) { s1 += (double)stats[f * fold_stride + c]; s2 += (double)stats[stats_stride + f * fold_stride + c]; }
            const double m = s1 / n;
            double S = s2 - s1 * m;
            if (S < 0) S = 0;
            const double is = 1.0 / sqrt(S / n + eps);
            sc = (float)(is * (double)gamma[c]); sh = beta[c] - (float)m * sc;
            if (blockIdx.x == 0) {
                mean[c] = (float)m; invstd[c] = (float)is; scale_g[c] = sc; shift_g[c] = sh;
                running_mean[c] = (float)(momentum * m + (1.0 - momentum) * (double)running_mean[c]);
                running_var[c] = (float)(momentum * (S / (n - 1.0)) + (1.0 - momentum) * (double)running_var[c]);
            }
        }
        sm_ss[c] = sc; sm_ss[Cp + c] = sh;
    }
    __syncthreads();                                   // all reads of `stats` by this CTA are done
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int k = atomicAdd(done_counter, 1u);
        is_last = (k == gridDim.x - 1);
        if (is_last) *done_counter = 0u;
    }
    __syncthreads();
    if (is_last)
        for (int i = threadIdx.x; i < fold * fold_stride; i += blockDim.x) { stats[i] = 0.f; stats[stats_stride + i] = 0.f; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % vec_per_pix) * 8;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(y) + i), f);
        const float4 s0 = *reinterpret_cast<const float4 *>(sm_ss + c0), s1 = *reinterpret_cast<const float4 *>(sm_ss + c0 + 4);
        const float4 h0 = *reinterpret_cast<const float4 *>(sm_ss + Cp + c0), h1 = *reinterpret_cast<const float4 *>(sm_ss + Cp + c0 + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = act_fwd(fmaf(f[k], sc[k], sh[k]), act, negval);
        reinterpret_cast<uint4 *>(a)[i] = pack8(f);
    }
}

// Per-channel reductions over NHWC: blockDim = (TX channel-vectors, TY pixel lanes), grid = (pixel strips, vector groups).
// Each thread owns one 8-channel vector position and walks pixels with stride gridDim.x*TY, U pixels per iteration with
// all loads issued before the arithmetic (bytes in flight hide HBM latency); partial sums are folded with shared-memory
// atomics (TY-way contention at most) and leave the CTA as one global fp32 atomic per channel.
constexpr int RED_U = 4;
constexpr int BN_U = 2;   // BN backward keeps 24-40 per-channel constants in registers: 2 vectors of each tensor in flight per thread

template <int NACC, bool kAtomic = true>
__device__ __forceinline__ void fold_and_flush(float (&acc)[NACC][8], bool active, float *__restrict__ out, int out_stride, int C_valid) {
    extern __shared__ float red_sh[];                 // [NACC][TX*8]
    const int TX = blockDim.x, tx = threadIdx.x, tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const int width = TX * 8;
    for (int i = tid; i < NACC * width; i += nthr) red_sh[i] = 0.f;
    __syncthreads();
    // lanes of a warp that share the same channel vector (TX < 32) are combined with shuffles first, so at most
    // (threads / 32) atomics hit one shared address
    if (TX < 32) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float v = active ? acc[a][k] : 0.f;
                for (int o = 16; o >= TX; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[a][k] = v;
            }
    }
    const bool leader = TX >= 32 || (tid & 31) < TX;
    if (active && leader) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(&red_sh[a * width + tx * 8 + k], acc[a][k]);
    }
    __syncthreads();
    const int c_base = blockIdx.y * width;
    for (int i = tid; i < NACC * width; i += nthr) {
        int a = i / width, c = c_base + (i - a * width);
        if (kAtomic) { if (c < C_valid) atomicAdd(out + a * out_stride + c, red_sh[i]); }
        else if (c < out_stride) out[a * out_stride + c] = c < C_valid ? red_sh[i] : 0.f;   // this CTA's own partial row
    }
}

// The per-channel sums of the backward kernels leave each CTA as one PARTIAL ROW (plain stores): hundreds of CTAs adding
// into the same few cache lines serialise in the L2 atomic unit and used to cost more than the streaming pass itself.
//   part[blockIdx.x][a][c],  row stride = NACC * Cp;  a later kernel (bn_bwd_coef / fold_rows) sums the rows.
// activation derivative from the PRE-activation z = y * scale + shift (recomputed: saves re-reading the activation output)
__device__ __forceinline__ float act_bwd_z(float z, int act, float negval) {
    if (act == ACT_LEAKY) return z > 0.f ? 1.f : negval;
    if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (act == ACT_TANH) { float t = tanhf(z); return 1.f - t * t; }
    if (act == ACT_SIGMOID) { float t = 1.f / (1.f + __expf(-z)); return t * (1.f - t); }
    return 1.f;
}
__device__ __forceinline__ void load8f(const float *__restrict__ p, int c0, int C, float (&o)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (c0 + k < C) ? p[c0 + k] : 0.f;
}

// compile-time activation derivative from the pre-activation z
template <int ACT>
__device__ __forceinline__ float dact_z(float z, float negval) {
    if (ACT == ACT_LEAKY) return z > 0.f ? 1.f : negval;
    if (ACT == ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return act_bwd_z(z, ACT, negval);
}
// bf16 pair -> two floats with two integer ops (a bf16 is the upper half of the fp32 bit pattern)
__device__ __forceinline__ void unpack8b(const uint4 &u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}

// BN backward pass 1: part[cta][0][c] = sum dz, part[cta][1][c] = sum dz * (y - mean[c]),  dz = g * act'(y*scale+shift)
template <int ACT, bool kAtomic = false>
__global__ void __launch_bounds__(256, 4) bn_bwd_reduce2_kernel(const bf16 *__restrict__ g, const bf16 *__restrict__ y, const float *__restrict__ scale,
        const float *__restrict__ shift, const float *__restrict__ mean, float *__restrict__ part, int Cp, int64_t npix, int vec_per_pix, int C,
        float negval) {
    pdl_trigger(); pdl_wait();
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vec_per_pix;
    float acc[2][8] = {};
    if (active) {
        float mu[8], sc[8], sh[8];
        load8f(mean, vec * 8, C, mu); load8f(scale, vec * 8, C, sc); load8f(shift, vec * 8, C, sh);
        const int64_t stride = (int64_t)gridDim.x * blockDim.y;
        for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * BN_U) {
            uint4 rg[BN_U], ry[BN_U];
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) { int64_t vi = p * vec_per_pix + vec; rg[u] = __ldg(reinterpret_cast<const uint4 *>(g) + vi); ry[u] = __ldg(reinterpret_cast<const uint4 *>(y) + vi); }
            }
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                if (p0 + u * stride < npix) {
                    float fg[8], fy[8];
                    unpack8b(rg[u], fg); unpack8b(ry[u], fy);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { float dz = fg[k] * dact_z<ACT>(fmaf(fy[k], sc[k], sh[k]), negval); acc[0][k] += dz; acc[1][k] = fmaf(dz, fy[k] - mu[k], acc[1][k]); }
                }
            }
        }
    }
    // kAtomic: `part` is ONE row [2][Cp] of fp32 sums that every CTA adds into (consumed and re-zeroed by bn_bwd_coef_apply_kernel)
    if (kAtomic) fold_and_flush<2, true>(acc, active, part, Cp, C);
    else fold_and_flush<2, false>(acc, active, part + (size_t)blockIdx.x * 2 * Cp, Cp, C);
}
// BN backward pass 1 WITH the coefficient step (single GPU): every CTA adds its partial sums into one [2][Cp] row (fp32 atomics), and the
// LAST CTA to finish (threadfence + counter) turns the completed sums into the pass-2 coefficients and the BN affine gradients, zeroes the
// row and the counter for the next use.  Removes the dependent bn_bwd_coef2 launch (~10 us on the critical path per BN layer and sweep)
// without touching the apply kernel (putting the coefficients into ITS prologue cost as much as the launch: profiles/r2_notes.md).
template <int ACT>
__global__ void __launch_bounds__(256, 4) bn_bwd_reduce_coef_kernel(const bf16 *__restrict__ g, const bf16 *__restrict__ y, const float *__restrict__ scale,
        const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
        float *__restrict__ sums, float *__restrict__ coef, float *__restrict__ ggamma, float *__restrict__ gbeta, int Cp, int64_t npix, int vec_per_pix,
        int C, float negval, double n, unsigned int *__restrict__ done_counter) {
    pdl_trigger(); pdl_wait();
    __shared__ int is_last;
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vec_per_pix;
    float acc[2][8] = {};
    if (active) {
        float mu[8], sc[8], sh[8];
        load8f(mean, vec * 8, C, mu); load8f(scale, vec * 8, C, sc); load8f(shift, vec * 8, C, sh);
        const int64_t stride = (int64_t)gridDim.x * blockDim.y;
        for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * BN_U) {
            uint4 rg[BN_U], ry[BN_U];
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) { int64_t vi = p * vec_per_pix + vec; rg[u] = __ldg(reinterpret_cast<const uint4 *>(g) + vi); ry[u] = __ldg(reinterpret_cast<const uint4 *>(y) + vi); }
            }
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                if (p0 + u * stride < npix) {
                    float fg[8], fy[8];
                    unpack8b(rg[u], fg); unpack8b(ry[u], fy);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { float dz = fg[k] * dact_z<ACT>(fmaf(fy[k], sc[k], sh[k]), negval); acc[0][k] += dz; acc[1][k] = fmaf(dz, fy[k] - mu[k], acc[1][k]); }
                }
            }
        }
    }
    fold_and_flush<2, true>(acc, active, sums, Cp, C);
    __syncthreads();                                   // this CTA's atomics are issued
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    if (tid == 0) {
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int k = atomicAdd(done_counter, 1u);
        is_last = (k == total - 1);
        if (is_last) { *done_counter = 0u; __threadfence(); }
    }
    __syncthreads();
    if (!is_last) return;
    for (int c = tid; c < Cp; c += nthr) {
        const double s = (double)__ldcg(sums + c), d = (double)__ldcg(sums + Cp + c);
        sums[c] = 0.f; sums[Cp + c] = 0.f;
        if (c < C) {
            const double is = invstd[c], A = is * (double)gamma[c], k1 = is * is * d / n;
            coef[c] = (float)A; coef[C + c] = (float)(A * k1); coef[2 * C + c] = (float)(((double)mean[c] * k1 - s / n) * A);
            if (ggamma) ggamma[c] += (float)(d * is);
            if (gbeta) gbeta[c] += (float)s;
        }
    }
}
// the coefficient step alone, from sums that a dgrad GEMM epilogue accumulated (tc::bwd_sums_slab): pass-2 coefficients, BN affine gradients,
// sums re-zeroed for the next use
__global__ void __launch_bounds__(256) bn_bwd_coef_sums_kernel(float *__restrict__ sums, int Cp, const float *__restrict__ gamma, const float *__restrict__ invstd,
        const float *__restrict__ mean, float *__restrict__ coef, float *__restrict__ ggamma, float *__restrict__ gbeta, int C, double n) {
    pdl_trigger(); pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    const double s = (double)sums[c], d = (double)sums[Cp + c];
    sums[c] = 0.f; sums[Cp + c] = 0.f;
    if (c < C) {
        const double is = invstd[c], A = is * (double)gamma[c], k1 = is * is * d / n;
        coef[c] = (float)A; coef[C + c] = (float)(A * k1); coef[2 * C + c] = (float)(((double)mean[c] * k1 - s / n) * A);
        if (ggamma) ggamma[c] += (float)(d * is);
        if (gbeta) gbeta[c] += (float)s;
    }
}
// sums the partial rows; coefficients for pass 2 + BN parameter gradients.
// sums_io [2][Cp]: rows > 0: written with the folded sums (the buffer a data-parallel run all-reduces); rows == 0: read.
// block = (32 channels, 8 row lanes); grid = ceil(C / 32).  coef for pass 2 is stored pre-combined:
//   g_y = dz * A - y * B + D  with  A = invstd*gamma, B = A * invstd^2 * d/n, D = (mean * invstd^2 * d/n - s/n) * A
__global__ void __launch_bounds__(256) bn_bwd_coef2_kernel(const float *__restrict__ part, int rows, float *__restrict__ sums_io, int Cp, const float *__restrict__ gamma,
        const float *__restrict__ invstd, const float *__restrict__ mean, float *__restrict__ coef, float *__restrict__ ggamma, float *__restrict__ gbeta, int C, double n, int emit_coef, float grad_scale) {
    pdl_trigger(); pdl_wait();
    __shared__ float sh_s[8][33], sh_d[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x, ry = threadIdx.y;
    float s0 = 0.f, d0 = 0.f, s1 = 0.f, d1 = 0.f;
    if (c < C && rows > 0) {
        int r = ry;
        for (; r + 8 < rows; r += 16) {
            const float *q = part + (size_t)r * 2 * Cp + c, *q2 = q + (size_t)16 * Cp;
            s0 += q[0]; d0 += q[Cp]; s1 += q2[0]; d1 += q2[Cp];
        }
        if (r < rows) { const float *q = part + (size_t)r * 2 * Cp + c; s0 += q[0]; d0 += q[Cp]; }
    }
    sh_s[ry][threadIdx.x] = s0 + s1; sh_d[ry][threadIdx.x] = d0 + d1;
    __syncthreads();
    if (ry != 0 || c >= C) return;
    double s = 0.0, d = 0.0;
    if (rows > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += (double)sh_s[i][threadIdx.x]; d += (double)sh_d[i][threadIdx.x]; }
        if (sums_io) { sums_io[c] = (float)s; sums_io[Cp + c] = (float)d; }
    } else { s = sums_io[c]; d = sums_io[Cp + c]; }
    if (!emit_coef) return;
    const double is = invstd[c], A = is * (double)gamma[c], k1 = is * is * d / n;
    coef[c] = (float)A; coef[C + c] = (float)(A * k1); coef[2 * C + c] = (float)(((double)mean[c] * k1 - s / n) * A);
    // data parallel: s and d are already GLOBAL sums and the gradient vector is summed over ranks later -> 1/world here
    if (ggamma) ggamma[c] += (float)(d * is) * grad_scale;
    if (gbeta) gbeta[c] += (float)s * grad_scale;
}
// BN backward pass 2 (in place on g): g_y = dz * A - y * B + D (coefficients above); optional per-CTA partial sums of g_y
// (-> conv gradBias, folded by fold_rows_kernel)
template <int ACT>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply2_kernel(bf16 *__restrict__ g, const bf16 *__restrict__ y, const float *__restrict__ scale,
        const float *__restrict__ shift, const float *__restrict__ coef, float *__restrict__ gb_part, int Cp,
        int64_t npix, int vec_per_pix, int C, float negval) {
    pdl_trigger(); pdl_wait();
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vec_per_pix;
    float acc[1][8] = {};
    if (active) {
        float sc[8], sh[8], cA[8], cB[8], cD[8];
        load8f(scale, vec * 8, C, sc); load8f(shift, vec * 8, C, sh);
        load8f(coef, vec * 8, C, cA); load8f(coef + C, vec * 8, C, cB); load8f(coef + 2 * C, vec * 8, C, cD);
        const int64_t stride = (int64_t)gridDim.x * blockDim.y;
        for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * BN_U) {
            uint4 rg[BN_U], ry[BN_U];
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) { int64_t vi = p * vec_per_pix + vec; rg[u] = reinterpret_cast<const uint4 *>(g)[vi]; ry[u] = __ldg(reinterpret_cast<const uint4 *>(y) + vi); }
            }
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) {
                    float fg[8], fy[8];
                    unpack8b(rg[u], fg); unpack8b(ry[u], fy);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float dz = fg[k] * dact_z<ACT>(fmaf(fy[k], sc[k], sh[k]), negval);
                        float r = fmaf(dz, cA[k], fmaf(-fy[k], cB[k], cD[k]));          // all three are 0 on padded lanes
                        fg[k] = r; acc[0][k] += r;
                    }
                    reinterpret_cast<uint4 *>(g)[p * vec_per_pix + vec] = pack8(fg);
                }
            }
        }
    }
    if (gb_part) fold_and_flush<1, false>(acc, active, gb_part + (size_t)blockIdx.x * Cp, Cp, C);
}
// BN backward pass 2 with the coefficient step folded into its prologue (single GPU): every CTA derives A, B, D of ITS channels
// from the summed [2][Cp] row left by bn_bwd_reduce2_kernel<ACT, true> (a few fp64 operations per thread), CTA row 0 also
// accumulates the BN affine gradients, and the last CTA to have read the sums zeroes them for the next use.  Removes the
// dependent bn_bwd_coef2 launch (~10 us on the critical path) per BN layer and sweep.
template <int ACT>
__global__ void __launch_bounds__(256, 3) bn_bwd_coef_apply_kernel(bf16 *__restrict__ g, const bf16 *__restrict__ y, const float *__restrict__ scale,
        const float *__restrict__ shift, float *__restrict__ sums, const float *__restrict__ gamma, const float *__restrict__ invstd,
        const float *__restrict__ mean, float *__restrict__ ggamma, float *__restrict__ gbeta, float *__restrict__ gb_part, int Cp,
        int64_t npix, int vec_per_pix, int C, float negval, double n, unsigned int *__restrict__ done_counter) {
    pdl_trigger(); pdl_wait();
    __shared__ int is_last;
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vec_per_pix;
    float acc[1][8] = {};
    float sc[8], sh[8], cA[8], cB[8], cD[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = vec * 8 + k;
        sc[k] = sh[k] = cA[k] = cB[k] = cD[k] = 0.f;
        if (active && c < C) {
            const double s = (double)__ldcg(sums + c), d = (double)__ldcg(sums + Cp + c);
            const double is = invstd[c], A = is * (double)gamma[c], k1 = is * is * d / n;
            sc[k] = scale[c]; sh[k] = shift[c];
            cA[k] = (float)A; cB[k] = (float)(A * k1); cD[k] = (float)(((double)mean[c] * k1 - s / n) * A);
            if (blockIdx.x == 0 && threadIdx.y == 0) {
                if (ggamma) ggamma[c] += (float)(d * is);
                if (gbeta) gbeta[c] += (float)s;
            }
        }
    }
    __syncthreads();                                   // all reads of `sums` by this CTA are done
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int k = atomicAdd(done_counter, 1u);
        is_last = (k == total - 1);
        if (is_last) *done_counter = 0u;
    }
    __syncthreads();
    if (is_last) {
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        for (int i = tid; i < 2 * Cp; i += blockDim.x * blockDim.y) sums[i] = 0.f;
    }
    if (active) {
        const int64_t stride = (int64_t)gridDim.x * blockDim.y;
        for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * BN_U) {
            uint4 rg[BN_U], ry[BN_U];
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) { int64_t vi = p * vec_per_pix + vec; rg[u] = reinterpret_cast<const uint4 *>(g)[vi]; ry[u] = __ldg(reinterpret_cast<const uint4 *>(y) + vi); }
            }
#pragma unroll
            for (int u = 0; u < BN_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) {
                    float fg[8], fy[8];
                    unpack8b(rg[u], fg); unpack8b(ry[u], fy);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float dz = fg[k] * dact_z<ACT>(fmaf(fy[k], sc[k], sh[k]), negval);
                        float r = fmaf(dz, cA[k], fmaf(-fy[k], cB[k], cD[k]));          // all three are 0 on padded lanes
                        fg[k] = r; acc[0][k] += r;
                    }
                    reinterpret_cast<uint4 *>(g)[p * vec_per_pix + vec] = pack8(fg);
                }
            }
        }
    }
    if (gb_part) fold_and_flush<1, false>(acc, active, gb_part + (size_t)blockIdx.x * Cp, Cp, C);
}
// The three BN-backward launches of a layer (reduce -> coefficients -> apply: ~12 + 8 + 10 us of mostly launch latency on the
// small layers, and a second HBM read of g and y on the large ones) as ONE kernel with two grid barriers.  Launched
// COOPERATIVELY (cudaLaunchAttributeCooperative: the grid starts only when every CTA can be co-resident), grid capped by the
// host at the kernel's occupancy; the barrier counter only grows, so it needs no reset between launches.  Between phase A
// and phase C the layer's g and y (<= 2 x 33 MB) stay in the 126 MB L2, so phase C's reads do not go back to HBM.
// A CTA that waits longer than ~1 minute traps (protocol bug) instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned int total) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(ctr, 1ULL);
        const unsigned long long target = (old / total + 1ULL) * total;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 120000000000LL) { printf("cenn: grid barrier timeout (cta %d,%d)\n", blockIdx.x, blockIdx.y); __trap(); }
        }
        __threadfence();
    }
    __syncthreads();
}
template <int ACT>
__global__ void __launch_bounds__(256, 3) bn_bwd_fused_kernel(bf16 *__restrict__ g, const bf16 *__restrict__ y, const float *__restrict__ scale,
        const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
        float *__restrict__ part, float *__restrict__ coef, float *__restrict__ ggamma, float *__restrict__ gbeta, int want_gb, int Cp,
        int64_t npix, int vec_per_pix, int C, float negval, double n, unsigned long long *__restrict__ bar) {
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vec_per_pix;
    const unsigned int total = gridDim.x * gridDim.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.y;
    float sc[8], sh[8];
    load8f(scale, vec * 8, active ? C : 0, sc); load8f(shift, vec * 8, active ? C : 0, sh);
    {   // ---- phase A: this CTA's partial row of (sum dz, sum dz * (y - mean))
        float acc[2][8] = {};
        if (active) {
            float mu[8];
            load8f(mean, vec * 8, C, mu);
            for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * BN_U) {
                uint4 rg[BN_U], ry[BN_U];
#pragma unroll
                for (int u = 0; u < BN_U; ++u) {
                    const int64_t p = p0 + u * stride;
                    if (p < npix) { const int64_t vi = p * vec_per_pix + vec; rg[u] = reinterpret_cast<const uint4 *>(g)[vi]; ry[u] = reinterpret_cast<const uint4 *>(y)[vi]; }
                }
#pragma unroll
                for (int u = 0; u < BN_U; ++u) {
                    if (p0 + u * stride < npix) {
                        float fg[8], fy[8];
                        unpack8b(rg[u], fg); unpack8b(ry[u], fy);
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const float dz = fg[k] * dact_z<ACT>(fmaf(fy[k], sc[k], sh[k]), negval); acc[0][k] += dz; acc[1][k] = fmaf(dz, fy[k] - mu[k], acc[1][k]); }
                    }
                }
            }
        }
        fold_and_flush<2, false>(acc, active, part + (size_t)blockIdx.x * 2 * Cp, Cp, C);
    }
    grid_barrier(bar, total);
    {   // ---- phase B: 32-channel groups, round-robin over the CTAs: fold the rows and publish the coefficients
        __shared__ float sh_s[8][33], sh_d[8][33];
        const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
        const int tid = threadIdx.y * blockDim.x + threadIdx.x, lane = tid & 31, ry = tid >> 5;
        for (unsigned int grp = cta; (int)grp * 32 < C; grp += total) {
            const int c = grp * 32 + lane, rows = gridDim.x;
            float s0 = 0.f, d0 = 0.f;
            if (c < C) for (int r = ry; r < rows; r += 8) { const float *q = part + (size_t)r * 2 * Cp + c; s0 += __ldcg(q); d0 += __ldcg(q + Cp); }
            sh_s[ry][lane] = s0; sh_d[ry][lane] = d0;
            __syncthreads();
            if (ry == 0 && c < C) {
                double s = 0.0, d = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { s += (double)sh_s[i][lane]; d += (double)sh_d[i][lane]; }
                const double is = invstd[c], A = is * (double)gamma[c], k1 = is * is * d / n;
                coef[c] = (float)A; coef[C + c] = (float)(A * k1); coef[2 * C + c] = (float)(((double)mean[c] * k1 - s / n) * A);
                if (ggamma) ggamma[c] += (float)(d * is);
                if (gbeta) gbeta[c] += (float)s;
            }
            __syncthreads();                                                  // sh_s / sh_d are reused by the next group
        }
    }
    grid_barrier(bar, total);
    {   // ---- phase C: g_y = dz * A - y * B + D in place; this CTA's partial row of sum g_y (-> conv gradBias)
        float acc[1][8] = {};
        if (active) {
            float cA[8], cB[8], cD[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int c = vec * 8 + k; const bool ok = c < C; cA[k] = ok ? __ldcg(coef + c) : 0.f; cB[k] = ok ? __ldcg(coef + C + c) : 0.f; cD[k] = ok ? __ldcg(coef + 2 * C + c) : 0.f; }
            for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * BN_U) {
                uint4 rg[BN_U], ry[BN_U];
#pragma unroll
                for (int u = 0; u < BN_U; ++u) {
                    const int64_t p = p0 + u * stride;
                    if (p < npix) { const int64_t vi = p * vec_per_pix + vec; rg[u] = reinterpret_cast<const uint4 *>(g)[vi]; ry[u] = reinterpret_cast<const uint4 *>(y)[vi]; }
                }
#pragma unroll
                for (int u = 0; u < BN_U; ++u) {
                    const int64_t p = p0 + u * stride;
                    if (p < npix) {
                        float fg[8], fy[8];
                        unpack8b(rg[u], fg); unpack8b(ry[u], fy);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float dz = fg[k] * dact_z<ACT>(fmaf(fy[k], sc[k], sh[k]), negval);
                            const float r = fmaf(dz, cA[k], fmaf(-fy[k], cB[k], cD[k]));
                            fg[k] = r; acc[0][k] += r;
                        }
                        reinterpret_cast<uint4 *>(g)[p * vec_per_pix + vec] = pack8(fg);
                    }
                }
            }
        }
        if (want_gb) fold_and_flush<1, false>(acc, active, part + (size_t)blockIdx.x * Cp, Cp, C);
    }
}
// activation-only backward (in place on g): g_y = g * act'(a); optional per-CTA partial sums -> gradBias
template <int ACT>
__global__ void __launch_bounds__(256, 4) act_bwd2_kernel(bf16 *__restrict__ g, const bf16 *__restrict__ a, float *__restrict__ gb_part, int Cp, int64_t npix,
        int vec_per_pix, int C, float negval) {
    pdl_trigger(); pdl_wait();
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vec_per_pix;
    float acc[1][8] = {};
    if (active) {
        float keep[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) keep[k] = (vec * 8 + k < C) ? 1.f : 0.f;
        const int64_t stride = (int64_t)gridDim.x * blockDim.y;
        for (int64_t p0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p0 < npix; p0 += stride * RED_U) {
            uint4 rg[RED_U], ra[RED_U];
#pragma unroll
            for (int u = 0; u < RED_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) { int64_t vi = p * vec_per_pix + vec; rg[u] = reinterpret_cast<const uint4 *>(g)[vi]; ra[u] = __ldg(reinterpret_cast<const uint4 *>(a) + vi); }
            }
#pragma unroll
            for (int u = 0; u < RED_U; ++u) {
                int64_t p = p0 + u * stride;
                if (p < npix) {
                    float fg[8], fa[8];
                    unpack8b(rg[u], fg); unpack8b(ra[u], fa);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float d;
                        if (ACT == ACT_LEAKY) d = fa[k] > 0.f ? 1.f : negval; else if (ACT == ACT_RELU) d = fa[k] > 0.f ? 1.f : 0.f; else d = act_bwd(fa[k], ACT, negval);
                        float r = fg[k] * d * keep[k]; fg[k] = r; acc[0][k] += r;
                    }
                    reinterpret_cast<uint4 *>(g)[p * vec_per_pix + vec] = pack8(fg);
                }
            }
        }
    }
    if (gb_part) fold_and_flush<1, false>(acc, active, gb_part + (size_t)blockIdx.x * Cp, Cp, C);
}
// dst[c] += sum_r src[r][c] for a list of jobs (one CTA per job): folds the per-CTA partial rows of a whole backward sweep
struct FoldJob { const float *src; float *dst; int rows, C, stride, fold, fold_stride; };   // fold > 1: column c also sums columns c + k*fold_stride (k < fold)
__global__ void __launch_bounds__(1024) fold_rows_kernel(const FoldJob *__restrict__ jobs) {
    pdl_trigger(); pdl_wait();
    // 32 channels x 32 row lanes per CTA: the row walk is a chain of dependent L2 loads, so its length (rows / 32) is what costs
    __shared__ float sh[32][33];
    const FoldJob j = jobs[blockIdx.x];
    const int tx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    for (int c0 = blockIdx.y * 32; c0 < j.C; c0 += gridDim.y * 32) {      // uniform per CTA
        const int c = c0 + tx;
        float s0 = 0.f, s1 = 0.f;
        if (c < j.C) {
            for (int f = 0; f < j.fold; ++f) {
                const float *q = j.src + c + f * j.fold_stride;
                int r = ry;
                for (; r + 32 < j.rows; r += 64) { s0 += q[(size_t)r * j.stride]; s1 += q[(size_t)(r + 32) * j.stride]; }
                if (r < j.rows) s0 += q[(size_t)r * j.stride];
            }
        }
        sh[ry][tx] = s0 + s1;
        __syncthreads();
        if (ry == 0 && c < j.C) { float t = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) t += sh[i][tx];
            j.dst[c] += t; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 4x4 / stride-1 / pad-0 windows on maps larger than 4x4 (fineSize 256: cfg4)
// The bottleneck convolution, the first decoder layer and the discriminator head are plain GEMMs when their 4x4 window covers the
// whole map (fineSize 128).  At fineSize 256 (train_deepernet at 256 x 256) the map is 8x8 and the window slides over 5x5 positions:
// the GEMMs stay, fed by / feeding an explicit window buffer  col[m = (n, oy, ox)][tap = 4u + v][c] = x[n, oy + u, ox + v, c].
__global__ void __launch_bounds__(256) im2col_v4_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ col, int N, int H, int W, int Cp) {
    pdl_trigger(); pdl_wait();
    const int ho = H - 3, wo = W - 3, vpp = Cp / 8;
    const int64_t total = (int64_t)N * ho * wo * 16 * vpp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % vpp); int64_t r = i / vpp;
        const int tap = (int)(r & 15); r >>= 4;
        const int ox = (int)(r % wo), oy = (int)((r / wo) % ho), n = (int)(r / ((int64_t)wo * ho));
        const int64_t src = (((int64_t)n * H + oy + (tap >> 2)) * W + ox + (tap & 3)) * vpp + cv;
        reinterpret_cast<uint4 *>(col)[i] = __ldg(reinterpret_cast<const uint4 *>(x) + src);
    }
}
// adjoint: out[n, y, x, c] = bias[c] + sum over the windows (oy, ox) = (y - u, x - v) that cover the pixel of col[(n, oy, ox)][4u + v][c];
// optional per-channel sum / sum of squares of the STORED (bf16) values -> BN statistics of the first decoder layer.
// blockDim = (TX channel vectors, TY pixel lanes), grid = (pixel strips, vector groups), as the BN reductions.
__global__ void __launch_bounds__(256) col2im_v4_kernel(const bf16 *__restrict__ col, bf16 *__restrict__ out, const float *__restrict__ bias, float *__restrict__ stats,
        int stats_stride, int N, int H, int W, int Cp, int C) {
    pdl_trigger(); pdl_wait();
    const int ho = H - 3, wo = W - 3, vpp = Cp / 8;
    const int vec = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = vec < vpp;
    float acc[2][8] = {};
    if (active) {
        float bv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) bv[k] = (bias && vec * 8 + k < C) ? bias[vec * 8 + k] : 0.f;
        const int64_t npix = (int64_t)N * H * W, stride = (int64_t)gridDim.x * blockDim.y;
        for (int64_t p = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p < npix; p += stride) {
            const int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = bv[k];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int oy = y - u;
                if (oy < 0 || oy >= ho) continue;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int ox = x - v;
                    if (ox < 0 || ox >= wo) continue;
                    float t[8];
                    unpack8(__ldg(reinterpret_cast<const uint4 *>(col) + ((((int64_t)n * ho + oy) * wo + ox) * 16 + u * 4 + v) * vpp + vec), t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] += t[k];
                }
            }
            const uint4 pk = pack8(f);
            reinterpret_cast<uint4 *>(out)[p * vpp + vec] = pk;
            if (stats) {
                float q[8];
                unpack8(pk, q);
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc[0][k] += q[k]; acc[1][k] = fmaf(q[k], q[k], acc[1][k]); }
            }
        }
    }
    if (stats) fold_and_flush<2, true>(acc, active, stats, stats_stride, C);
}

// ---------------------------------------------------------------- discriminator head: 4x4 valid conv to 1 channel + Sigmoid + BCE
// out[b] = sigmoid(sum_k x[b,k] w[k] + bias): one warp per sample, 16-byte loads
__global__ void __launch_bounds__(256) head_fwd_kernel(const bf16 *__restrict__ x, const bf16 *__restrict__ w, const float *__restrict__ bias,
        float *__restrict__ out, int B, int K) {
    pdl_trigger(); pdl_wait();
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const uint4 *xr = reinterpret_cast<const uint4 *>(x + (int64_t)warp * K), *wr = reinterpret_cast<const uint4 *>(w);
    float acc = 0.f;
    for (int i = lane; i < K / 8; i += 32) {
        float a[8], b[8];
        unpack8(xr[i], a); unpack8(wr[i], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(a[k], b[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[warp] = 1.f / (1.f + expf(-(acc + (bias ? bias[0] : 0.f))));
}
// BCE against a constant label: loss_acc += -sum(...)/n ; gpre[b] = dL/dx * x(1-x)  (Sigmoid backward)
__global__ void __launch_bounds__(256) head_bce_kernel(const float *__restrict__ sig, float label, float *__restrict__ gpre, double *__restrict__ loss_acc,
        int B, double inv_n) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    double l = 0.0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float x = sig[b];
        l += (double)(label * logf(x + 1e-12f) + (1.f - label) * logf(1.f - x + 1e-12f));
        float gx = -(float)inv_n * (label - x) / ((1.f - x + 1e-12f) * (x + 1e-12f));
        if (gpre) gpre[b] = gx * x * (1.f - x);
    }
    l = block_sum(l, sh);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, -l * inv_n);
}
// gx[b,k] = gpre[b] * w[k]
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float *__restrict__ gpre, const bf16 *__restrict__ w, bf16 *__restrict__ gx, int B, int K) {
    pdl_trigger(); pdl_wait();
    int64_t nvec = (int64_t)B * K / 8;
    int kv = K / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(i / kv), k = (int)(i % kv);
        float f[8], g = gpre[b];
        unpack8(reinterpret_cast<const uint4 *>(w)[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= g;
        reinterpret_cast<uint4 *>(gx)[i] = pack8(f);
    }
}
// gw[k] += sum_b gpre[b] x[b,k] ; gb += sum_b gpre[b]   (grid.x over k-vectors, grid.y over batch slices; fp32 atomics)
__global__ void __launch_bounds__(128) head_wgrad_kernel(const float *__restrict__ gpre, const bf16 *__restrict__ x, float *__restrict__ gw,
        float *__restrict__ gb, int B, int K) {
    pdl_trigger(); pdl_wait();
    const int kv = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = (B + gridDim.y - 1) / gridDim.y, b0 = blockIdx.y * per, b1 = min(B, b0 + per);
    if (kv < K / 8) {
        float acc[8] = {};
        int b = b0;
        for (; b + 4 <= b1; b += 4) {             // 4 independent 16-byte loads in flight
            uint4 r[4]; float g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { r[u] = __ldg(reinterpret_cast<const uint4 *>(x + (int64_t)(b + u) * K) + kv); g[u] = gpre[b + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { float f[8]; unpack8(r[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[u], f[j], acc[j]); }
        }
        for (; b < b1; ++b) { float f[8], g = gpre[b]; unpack8(reinterpret_cast<const uint4 *>(x + (int64_t)b * K)[kv], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(g, f[j], acc[j]); }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(gw + kv * 8 + j, acc[j]);
    }
    if (blockIdx.x == 0 && gb && threadIdx.x < 32) {
        float sacc = 0.f;
        for (int b = b0 + threadIdx.x; b < b1; b += 32) sacc += gpre[b];
        sacc = warp_sum(sacc);
        if (threadIdx.x == 0) atomicAdd(gb, sacc);
    }
}

// ---------------------------------------------------------------- generator losses on NHWC tensors (Cp lanes, C valid)
// image variant (train.lua:377-400): g = a*df + Wm(y,x) * 2(x-t)/n ; loss_acc += sum (x-t)^2 / n   (vectors of 4 lanes)
__global__ void __launch_bounds__(256) blend_overlap_kernel(const bf16 *__restrict__ df, const bf16 *__restrict__ x, const bf16 *__restrict__ t,
        bf16 *__restrict__ g, int64_t npix, int H, int W, int Cp, int C, int ov, float a, float w_in, float w_ring, float two_over_n,
        double inv_n, double *__restrict__ loss_acc) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    double l = 0.0;
    int64_t total = npix * Cp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cp);
        int64_t p = i / Cp;
        int xx = (int)(p % W), yy = (int)((p / W) % H);
        float r = 0.f;
        if (c < C) {
            bool inner = (yy >= ov && yy < H - ov && xx >= ov && xx < W - ov);
            float d = __bfloat162float(x[i]) - __bfloat162float(t[i]);
            l += (double)d * d;
            r = (df ? __bfloat162float(df[i]) * a : 0.f) + (inner ? w_in : w_ring) * (d * two_over_n);
        }
        g[i] = __float2bfloat16(r);
    }
    l = block_sum(l, sh);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, l * inv_n);
}
// Cp == 4 fast path of blend_overlap: one thread handles two pixels (16-byte vectors of df, x, t and g)
__global__ void __launch_bounds__(256) blend_overlap4_kernel(const bf16 *__restrict__ df, const bf16 *__restrict__ x, const bf16 *__restrict__ t,
        bf16 *__restrict__ g, int64_t npix, int H, int W, int C, int ov, float a, float w_in, float w_ring, float two_over_n,
        double inv_n, double *__restrict__ loss_acc) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    float l = 0.f;
    const int64_t npair = npix / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += (int64_t)gridDim.x * blockDim.x) {
        float fd[8] = {}, fx[8], ft[8], r[8];
        if (df) unpack8(__ldg(reinterpret_cast<const uint4 *>(df) + i), fd);
        unpack8(__ldg(reinterpret_cast<const uint4 *>(x) + i), fx); unpack8(__ldg(reinterpret_cast<const uint4 *>(t) + i), ft);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t p = 2 * i + h;
            const int xx = (int)(p % W), yy = (int)((p / W) % H);
            const float wgt = (yy >= ov && yy < H - ov && xx >= ov && xx < W - ov) ? w_in : w_ring;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = 4 * h + c;
                float d = fx[k] - ft[k];
                if (c >= C) d = 0.f;
                l = fmaf(d, d, l);
                r[k] = c < C ? fmaf(fd[k], a, wgt * (d * two_over_n)) : 0.f;
            }
        }
        reinterpret_cast<uint4 *>(g)[i] = pack8(r);
    }
    double ld = block_sum((double)l, sh);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, ld * inv_n);
}
// video variant (train_vid_weighted.lua:485-528): g = a*df + (wtl2 * (m(1-lam)+lam) + wtgdl) * 2(x-t)/n
__global__ void __launch_bounds__(256) blend_masked_kernel(const bf16 *__restrict__ df, const bf16 *__restrict__ x, const bf16 *__restrict__ t,
        const bf16 *__restrict__ mask, bf16 *__restrict__ g, int64_t total, int Cp, int C, float a, float wtl2, float lambda, float wtgdl,
        float two_over_n, double inv_n, double *__restrict__ loss_acc) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    double l = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cp);
        float r = 0.f;
        if (c < C) {
            float d = __bfloat162float(x[i]) - __bfloat162float(t[i]);
            l += (double)d * d;
            float w = lambda != 0.f ? __bfloat162float(mask[i]) * (1.f - lambda) + lambda : 1.f;
            r = (df ? __bfloat162float(df[i]) * a : 0.f) + (wtl2 * w + wtgdl) * (d * two_over_n);
        }
        g[i] = __float2bfloat16(r);
    }
    l = block_sum(l, sh);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, l * inv_n);
}
// Cp % 8 == 0 fast path of blend_masked: 16-byte vectors of df, x, t, mask and g
__global__ void __launch_bounds__(256) blend_masked8_kernel(const bf16 *__restrict__ df, const bf16 *__restrict__ x, const bf16 *__restrict__ t,
        const bf16 *__restrict__ mask, bf16 *__restrict__ g, int64_t nvec, int vec_per_pix, int C, float a, float wtl2, float lambda, float wtgdl,
        float two_over_n, double inv_n, double *__restrict__ loss_acc) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    float l = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % vec_per_pix) * 8;
        float fd[8] = {}, fx[8], ft[8], fm[8] = {}, r[8];
        if (df) unpack8(__ldg(reinterpret_cast<const uint4 *>(df) + i), fd);
        unpack8(__ldg(reinterpret_cast<const uint4 *>(x) + i), fx); unpack8(__ldg(reinterpret_cast<const uint4 *>(t) + i), ft);
        if (lambda != 0.f) unpack8(__ldg(reinterpret_cast<const uint4 *>(mask) + i), fm);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float d = fx[k] - ft[k];
            if (c0 + k >= C) d = 0.f;
            l = fmaf(d, d, l);
            const float w = lambda != 0.f ? fm[k] * (1.f - lambda) + lambda : 1.f;
            r[k] = c0 + k < C ? fd[k] * a + (wtl2 * w + wtgdl) * (d * two_over_n) : 0.f;
        }
        reinterpret_cast<uint4 *>(g)[i] = pack8(r);
    }
    double ld = block_sum((double)l, sh);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, ld * inv_n);
}
// dst = mask ? src : dst
__global__ void __launch_bounds__(256) composite_kernel(bf16 *__restrict__ dst, const bf16 *__restrict__ mask, const bf16 *__restrict__ src, int64_t total) {
    pdl_trigger(); pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        if (__bfloat162float(mask[i]) != 0.f) dst[i] = src[i];
}
// GDL loss (forward only: the scripts add criterionMSE:backward as its gradient, train_vid_weighted.lua:525), flat-index pairing
__global__ void __launch_bounds__(256) gdl_loss_kernel(const bf16 *__restrict__ inp, const bf16 *__restrict__ tgt, int64_t N, int H, int W, int Cp, int C,
        double inv_n, double *__restrict__ loss_acc) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    const int NK = H * (W - 1);
    double l = 0.0;
    int64_t total = N * NK * Cp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cp);
        if (c >= C) continue;
        int64_t t = i / Cp;
        int k = (int)(t % NK);
        int64_t n = t / NK;
        int ar = k / (W - 1), ac = k - ar * (W - 1), br = k / W, bc = k - br * W;
        auto at = [&](const bf16 *T, int r, int cc) { return __bfloat162float(T[((n * H + r) * W + cc) * Cp + c]); };
        float t12 = fabsf(at(tgt, ar, ac) - at(tgt, br, bc)) - fabsf(at(inp, ar, ac) - at(inp, br, bc));
        float t34 = fabsf(at(tgt, ar, ac + 1) - at(tgt, br + 1, bc)) - fabsf(at(inp, ar, ac + 1) - at(inp, br + 1, bc));
        l += (double)(fabsf(t12) + fabsf(t34));
    }
    l = block_sum(l, sh);
    if (threadIdx.x == 0) atomicAdd(loss_acc, l * inv_n);
}
// the same with one thread per (sample, pair index k): all CP channels of the four pixels involved come in as 8- or 16-byte vectors
template <int CP>
__global__ void __launch_bounds__(256) gdl_loss_vec_kernel(const bf16 *__restrict__ inp, const bf16 *__restrict__ tgt, int64_t N, int H, int W, int C,
        double inv_n, double *__restrict__ loss_acc) {
    pdl_trigger(); pdl_wait();
    __shared__ double sh[32];
    const int NK = H * (W - 1);
    float l = 0.f;
    const int64_t total = N * NK;
    auto load = [&](const bf16 *T, int64_t n, int r, int cc, float (&f)[CP]) {
        const bf16 *p = T + ((n * H + r) * W + cc) * CP;
        if (CP == 4) {
            const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
            const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162 *>(&u.x), h1 = *reinterpret_cast<const __nv_bfloat162 *>(&u.y);
            f[0] = __low2float(h0); f[1] = __high2float(h0); f[2] = __low2float(h1); f[3] = __high2float(h1);
        } else {
#pragma unroll
            for (int v = 0; v < CP / 8; ++v) {
                float e[8];
                unpack8(__ldg(reinterpret_cast<const uint4 *>(p) + v), e);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[(v * 8 + k) % CP] = e[k];
            }
        }
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % NK);
        const int64_t n = i / NK;
        const int ar = k / (W - 1), ac = k - ar * (W - 1), br = k / W, bc = k - br * W;
        float ta[CP], tb[CP], xa[CP], xb[CP];
        load(tgt, n, ar, ac, ta); load(tgt, n, br, bc, tb); load(inp, n, ar, ac, xa); load(inp, n, br, bc, xb);
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) if (c < C) part += fabsf(fabsf(ta[c] - tb[c]) - fabsf(xa[c] - xb[c]));
        load(tgt, n, ar, ac + 1, ta); load(tgt, n, br + 1, bc, tb); load(inp, n, ar, ac + 1, xa); load(inp, n, br + 1, bc, xb);
#pragma unroll
        for (int c = 0; c < CP; ++c) if (c < C) part += fabsf(fabsf(ta[c] - tb[c]) - fabsf(xa[c] - xb[c]));
        l += part;
    }
    double ld = block_sum((double)l, sh);
    if (threadIdx.x == 0) atomicAdd(loss_acc, ld * inv_n);
}

// ---------------------------------------------------------------- optimiser / parameter maintenance
// optim.adam on the master vector + refresh of the bf16 operand copy in the same pass
__global__ void __launch_bounds__(256) adam_bf16_kernel(float *__restrict__ x, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
        bf16 *__restrict__ xb, int64_t n, float b1, float b2, float eps, const float *__restrict__ step_ptr) {
    pdl_trigger(); pdl_wait();
    const float step = *step_ptr;
    int64_t n4 = n / 4;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (int64_t)gridDim.x * blockDim.x) {
        float4 X = reinterpret_cast<float4 *>(x)[j], G = reinterpret_cast<const float4 *>(g)[j];
        float4 M = reinterpret_cast<float4 *>(m)[j], V = reinterpret_cast<float4 *>(v)[j];
#define ADAM1(c) M.c = M.c * b1 + (1.f - b1) * G.c; V.c = V.c * b2 + (1.f - b2) * G.c * G.c; X.c -= step * M.c / (sqrtf(V.c) + eps);
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        reinterpret_cast<float4 *>(x)[j] = X; reinterpret_cast<float4 *>(m)[j] = M; reinterpret_cast<float4 *>(v)[j] = V;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(X.x, X.y), h1 = __floats2bfloat162_rn(X.z, X.w);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
        reinterpret_cast<uint2 *>(xb)[j] = pk;
    }
}
// Adam with the gradient read from a bf16 bucket (data parallel: the big generator blocks are all-reduced in bf16)
__global__ void __launch_bounds__(256) adam_bf16g_kernel(float *__restrict__ x, const bf16 *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
        bf16 *__restrict__ xb, int64_t n, float b1, float b2, float eps, const float *__restrict__ step_ptr) {
    pdl_trigger(); pdl_wait();
    const float step = *step_ptr;
    int64_t n4 = n / 4;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (int64_t)gridDim.x * blockDim.x) {
        float4 X = reinterpret_cast<float4 *>(x)[j];
        const uint2 gp = reinterpret_cast<const uint2 *>(g)[j];
        const __nv_bfloat162 g0 = *reinterpret_cast<const __nv_bfloat162 *>(&gp.x), g1 = *reinterpret_cast<const __nv_bfloat162 *>(&gp.y);
        const float4 G = make_float4(__low2float(g0), __high2float(g0), __low2float(g1), __high2float(g1));
        float4 M = reinterpret_cast<float4 *>(m)[j], V = reinterpret_cast<float4 *>(v)[j];
#define ADAM1(c) M.c = M.c * b1 + (1.f - b1) * G.c; V.c = V.c * b2 + (1.f - b2) * G.c * G.c; X.c -= step * M.c / (sqrtf(V.c) + eps);
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        reinterpret_cast<float4 *>(x)[j] = X; reinterpret_cast<float4 *>(m)[j] = M; reinterpret_cast<float4 *>(v)[j] = V;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(X.x, X.y), h1 = __floats2bfloat162_rn(X.z, X.w);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
        reinterpret_cast<uint2 *>(xb)[j] = pk;
    }
}
// fp32 -> bf16, 4 elements per thread (n % 4 == 0)
__global__ void __launch_bounds__(256) f32_to_bf16_vec_kernel(const float *__restrict__ x, bf16 *__restrict__ xb, int64_t n4) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (int64_t)gridDim.x * blockDim.x) {
        const float4 X = __ldg(reinterpret_cast<const float4 *>(x) + j);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(X.x, X.y), h1 = __floats2bfloat162_rn(X.z, X.w);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
        reinterpret_cast<uint2 *>(xb)[j] = pk;
    }
}
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ x, bf16 *__restrict__ xb, int64_t n) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) xb[j] = __float2bfloat16(x[j]);
}
// zero a list of (offset, length) segments of a float vector and of its bf16 copy (conv biases, train.lua:279-280)
__global__ void zero_segments_kernel(float *__restrict__ x, bf16 *__restrict__ xb, const int64_t *__restrict__ seg, int nseg) {
    pdl_trigger(); pdl_wait();
    int sidx = blockIdx.x;
    if (sidx >= nseg) return;
    int64_t off = seg[2 * sidx], len = seg[2 * sidx + 1];
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) { x[off + i] = 0.f; if (xb) xb[off + i] = __float2bfloat16(0.f); }
}
// Tiled transpose of the bf16 master copy Wf [rows=Cs][cols=K] into the dgrad / full-conv-fprop operand:
//   mode 0 (plain)  : dst[k][cs]                                   (K x Csp)
//   mode 1 (phases) : dst[((ph*cl_rows + cl)*4 + ab)][cs] with k = tap*Clp + cl, tap = u*4+v, (ph,ab) from (u,v)
__global__ void __launch_bounds__(256) wt_from_wf_kernel(const bf16 *__restrict__ Wf, bf16 *__restrict__ dst, int Cs, int K, int Csp, int Clp, int cl_rows, int mode) {
    pdl_trigger(); pdl_wait();
    __shared__ bf16 tile[32][34];
    int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int cs = c0 + r, k = k0 + tx;
        tile[r][tx] = (cs < Cs && k < K) ? Wf[(int64_t)cs * K + k] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int k = k0 + r, cs = c0 + tx;
        if (k >= K || cs >= Csp) continue;
        int64_t row;
        if (mode == 0) row = k;
        else {
            int tap = k / Clp, cl = k - tap * Clp;
            if (cl >= cl_rows) continue;
            int u = tap >> 2, v = tap & 3;
            // u = UPH[py][a]: {1,3} -> py 0 (a = 0,1); {0,2} -> py 1 (a = 0,1)
            int py = (u & 1) ? 0 : 1, a = (u & 1) ? (u >> 1) : (u >> 1);
            int px = (v & 1) ? 0 : 1, b = (v >> 1);
            row = ((int64_t)((py * 2 + px) * cl_rows + cl)) * 4 + (a * 2 + b);
        }
        dst[row * Csp + cs] = tile[tx][r];
    }
}


// ---------------------------------------------------------------- inference engine (cenn_inpainter_*)
// eval-mode BN folded into the operands: w'[.., c] = w[.., c] * gamma[c] / sqrt(running_var[c] + eps) (bf16 operand copy) and
// bias'[c] = (bias[c] - running_mean[c]) * that scale + beta[c] (fp32, added to the fp32 accumulator in the GEMM epilogue)
__global__ void __launch_bounds__(256) fold_bn_weights_kernel(const float *__restrict__ w, bf16 *__restrict__ wb, int64_t count, int Clp, int chan_is_row,
        const float *__restrict__ gamma, const float *__restrict__ running_var, int Cout, double eps) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = chan_is_row ? (int)(i / (16 * (int64_t)Clp)) : (int)(i % Clp);
        const float sc = c < Cout ? (float)((double)gamma[c] / sqrt((double)running_var[c] + eps)) : 0.f;
        wb[i] = __float2bfloat16(w[i] * sc);
    }
}
// out[rep][Cp]: the folded bias, repeated `reps` times (G1's GEMM columns are (tap, channel)); gamma == nullptr: plain conv bias
__global__ void fold_bn_bias_kernel(const float *__restrict__ bias, const float *__restrict__ gamma, const float *__restrict__ beta,
        const float *__restrict__ running_mean, const float *__restrict__ running_var, float *__restrict__ out, int Cout, int Cp, int reps, double eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= reps * Cp) return;
    const int c = i % Cp;
    float v = 0.f;
    if (c < Cout) {
        if (gamma) { const double sc = (double)gamma[c] / sqrt((double)running_var[c] + eps); v = (float)(((double)bias[c] - (double)running_mean[c]) * sc + (double)beta[c]); }
        else v = bias[c];
    }
    out[i] = v;
}

// Device-side clip preparation of the video loader's per-sample hook (datavid/donkey_folder.lua:161-187), after the crop:
// frames01 [N][C][H][W] in [0,1] and ONE mask plane per sample [N][H][W] -> the three step inputs as NHWC bf16:
//   full = 2f-1, masked = mask ? 2*maskValue-1 : full, mask expanded over the channels; flip[n] != 0 mirrors all three (hflip).
template <int CP>
__global__ void __launch_bounds__(256) clip_prepare_kernel(const float *__restrict__ frames01, const uint8_t *__restrict__ mask1, const uint8_t *__restrict__ flip,
        float maskValue, int N, int C, int H, int W, bf16 *__restrict__ masked, bf16 *__restrict__ full, bf16 *__restrict__ maskx) {
    pdl_trigger(); pdl_wait();
    const int HW = H * W;
    const int64_t total = (int64_t)N * HW;
    const float fillv = 2.f * maskValue - 1.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / HW), r = (int)(i - (int64_t)n * HW), y = r / W, x = r - y * W;
        const int xs = (flip && flip[n]) ? W - 1 - x : x;
        const bool m = mask1[(int64_t)n * HW + y * W + xs] != 0;
        __align__(16) bf16 vf[CP], vm[CP], vk[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            float f = 0.f, k = 0.f, mk = 0.f;
            if (c < C) {
                f = __ldg(frames01 + ((int64_t)n * C + c) * HW + y * W + xs) * 2.f - 1.f;
                k = m ? fillv : f; mk = m ? 1.f : 0.f;
            }
            vf[c] = __float2bfloat16(f); vk[c] = __float2bfloat16(k); vm[c] = __float2bfloat16(mk);
        }
        if (CP == 4) {
            *reinterpret_cast<uint2 *>(full + i * CP) = *reinterpret_cast<const uint2 *>(vf);
            *reinterpret_cast<uint2 *>(masked + i * CP) = *reinterpret_cast<const uint2 *>(vk);
            *reinterpret_cast<uint2 *>(maskx + i * CP) = *reinterpret_cast<const uint2 *>(vm);
        } else {
#pragma unroll
            for (int q = 0; q < CP / 8; ++q) {
                reinterpret_cast<uint4 *>(full + i * CP)[q] = reinterpret_cast<const uint4 *>(vf)[q];
                reinterpret_cast<uint4 *>(masked + i * CP)[q] = reinterpret_cast<const uint4 *>(vk)[q];
                reinterpret_cast<uint4 *>(maskx + i * CP)[q] = reinterpret_cast<const uint4 *>(vm)[q];
            }
        }
    }
}

// Image variant, byte input: what train.lua:286-290 does to a loader batch, from the DECODED BYTES (image.load = byte / 255, then the
// loader's mul(2):add(-1), data/donkey_folder.lua:84-86): real_center = centre F/2 x F/2 crop, real_ctx = image with the centre minus an
// `ov`-wide ring filled with the mean colour (2*117/255-1, 2*104/255-1, 2*123/255-1).  NHWC bf16 with Cp = 4.
__global__ void __launch_bounds__(256) image_u8_prepare_kernel(const uint8_t *__restrict__ img, int N, int F, int ov, bf16 *__restrict__ ctx, bf16 *__restrict__ center) {
    pdl_trigger(); pdl_wait();
    const int64_t total = (int64_t)N * F * F;
    const int q = F / 4, h = F / 2;
    const float fill[3] = {2.f * 117.f / 255.f - 1.f, 2.f * 104.f / 255.f - 1.f, 2.f * 123.f / 255.f - 1.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / ((int64_t)F * F)), r = (int)(i - (int64_t)n * F * F), y = r / F, x = r - y * F;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (float)img[((int64_t)n * 3 + c) * F * F + r] / 255.f * 2.f - 1.f;
        const bool in_center = y >= q && y < q + h && x >= q && x < q + h;
        if (in_center) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], 0.f);
            *reinterpret_cast<uint2 *>(center + (((int64_t)n * h + (y - q)) * h + (x - q)) * 4) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
        }
        if (y >= q + ov && y < q + h - ov && x >= q + ov && x < q + h - ov) { v[0] = fill[0]; v[1] = fill[1]; v[2] = fill[2]; }
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], 0.f);
        *reinterpret_cast<uint2 *>(ctx + i * 4) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
}

// Device-side remainder of the video loader's hook (datavid/donkey_folder.lua:114-129,138-170): random crop of the loaded frames and
// of the logo mask, and the random-block mask that replaces an all-black mask crop.  The RANDOM DRAWS stay on the host (crop origin,
// block count and corners: torch.uniform / torch.random, :149-150,120-123) and arrive as small integer tables; the pixels never do:
//   frames_u8 [N][C][iH][iW] (what image.load decodes: byte / 255), mask_full [iH][iW], crop[n] = (h1, w1) 0-based,
//   blocks[n] = (count, tlx_0, tly_0, ..., tlx_9, tly_9) 0-based inside the crop, block side = floor(F / 6).
// Output = the clip-mode inputs of clip_prepare_kernel: frames01 [N][C][F][F] and one mask plane per sample.
__global__ void __launch_bounds__(256) crop_mask_any_kernel(const uint8_t *__restrict__ mask_full, const int *__restrict__ crop, int iW, int F, int *__restrict__ any) {
    __shared__ int found;
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    const int n = blockIdx.x, h1 = crop[2 * n], w1 = crop[2 * n + 1];
    int f = 0;
    for (int i = threadIdx.x; i < F * F; i += blockDim.x) f |= mask_full[(int64_t)(h1 + i / F) * iW + w1 + i % F] != 0;
    if (f) found = 1;
    __syncthreads();
    if (threadIdx.x == 0) any[n] = found;                   // maskout:max() > 0.5 (:163)
}
__global__ void __launch_bounds__(256) frames_to_clips_kernel(const uint8_t *__restrict__ frames_u8, const uint8_t *__restrict__ mask_full, const int *__restrict__ crop,
        const int *__restrict__ blocks, const int *__restrict__ any, int N, int C, int iH, int iW, int F, float *__restrict__ frames01, uint8_t *__restrict__ mask1) {
    const int64_t total = (int64_t)N * F * F;
    const int bs = F / 6;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / ((int64_t)F * F)), r = (int)(i - (int64_t)n * F * F), y = r / F, x = r - y * F;
        const int h1 = crop[2 * n], w1 = crop[2 * n + 1];
        const int64_t src = (int64_t)(h1 + y) * iW + w1 + x;
        bool m;
        if (any[n]) m = mask_full[src] != 0;
        else {                                               // randomBlockMask (:114-129)
            m = false;
            const int *b = blocks + 21 * n;
            for (int k = 0; k < b[0]; ++k) { const int tlx = b[1 + 2 * k], tly = b[2 + 2 * k]; m = m || (x >= tlx && x < tlx + bs && y >= tly && y < tly + bs); }
        }
        mask1[i] = m ? 1 : 0;
        for (int c = 0; c < C; ++c) frames01[((int64_t)n * C + c) * F * F + r] = (float)frames_u8[((int64_t)n * C + c) * iH * iW + src] / 255.f;
    }
}

// Full-frame sweep of test_vid_wholeim.lua:98-226.  Tile j = ti * groups + g: ti walks the padded frame row-major in FxF tiles,
// g is the frame group (ncin = nc * inputLen stacked channels); the first three tiles of the top row are fed upside down (:167-170).
struct SweepGeom { int P, nc, inh, inw, outh, outw, F, ncin, Cp, groups, tiles_w; };
__device__ __forceinline__ void sweep_tile(const SweepGeom &q, int j, int &h0, int &w0, int &g, bool &flip) {
    const int ti = j / q.groups; g = j - ti * q.groups;
    const int tr = ti / q.tiles_w, tc = ti - tr * q.tiles_w;
    h0 = tr * q.F; w0 = tc * q.F; flip = tr == 0 && tc < 3;
}
// frames01 [P][nc][inh][inw] in [0,1], mask [inh][inw] -> generator input tiles [n][F][F][Cp] bf16 in [-1,1]: masked pixels = maskValue,
// bottom/right padding = 0 (:60-72,109-111).  mid != nullptr (withInit, :179-190): pixels under the tile's slice of the padded mask -- taken
// UN-flipped, as the reference does -- come from the initializer net's output for the same tile.
template <int CP>
__global__ void __launch_bounds__(256) wholeim_gather_kernel(const float *__restrict__ frames01, const uint8_t *__restrict__ mask, float maskValue, SweepGeom q,
        int j0, int n, bf16 *__restrict__ dst, const bf16 *__restrict__ mid) {
    const int FF = q.F * q.F;
    const int64_t total = (int64_t)n * FF;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / FF), r = (int)(i - (int64_t)t * FF), y = r / q.F, x = r - y * q.F;
        int h0, w0, g; bool flip;
        sweep_tile(q, j0 + t, h0, w0, g, flip);
        const int Y = h0 + (flip ? q.F - 1 - y : y), X = w0 + x;
        const bool inside = Y < q.inh && X < q.inw;
        const bool m = inside && mask[(int64_t)Y * q.inw + X] != 0;
        const int Yu = h0 + y;
        const bool fill = mid != nullptr && Yu < q.inh && X < q.inw && mask[(int64_t)Yu * q.inw + X] != 0;
        __align__(16) bf16 v[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            float f = 0.f;
            if (c < q.ncin) {
                const int cc = g * q.ncin + c;
                const float s = inside ? (m ? maskValue : __ldg(frames01 + ((int64_t)cc * q.inh + Y) * q.inw + X)) : 0.f;
                f = s * 2.f - 1.f;
            }
            v[c] = __float2bfloat16(f);
        }
        if (fill) {
            const bf16 *ms = mid + i * CP;
#pragma unroll
            for (int c = 0; c < CP; ++c) if (c < q.ncin) v[c] = ms[c];
        }
        if (CP == 4) *reinterpret_cast<uint2 *>(dst + i * CP) = *reinterpret_cast<const uint2 *>(v);
        else {
#pragma unroll
            for (int k = 0; k < CP / 8; ++k) reinterpret_cast<uint4 *>(dst + i * CP)[k] = reinterpret_cast<const uint4 *>(v)[k];
        }
    }
}
// generator output tiles [n][F][F][Cp] -> outImages / fullImages / inpaintImages [P*nc][outh][outw] fp32 in [0,1] (:194-224):
// un-flip, write back, composite under the padded mask, rescale.
__global__ void __launch_bounds__(256) wholeim_scatter_kernel(const bf16 *__restrict__ gout, const float *__restrict__ frames01, const uint8_t *__restrict__ mask,
        float maskValue, SweepGeom q, int j0, int n, float *__restrict__ out01, float *__restrict__ full01, float *__restrict__ inpaint01) {
    const int FF = q.F * q.F;
    const int64_t total = (int64_t)n * FF;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / FF), r = (int)(i - (int64_t)t * FF), y = r / q.F, x = r - y * q.F;
        int h0, w0, g; bool flip;
        sweep_tile(q, j0 + t, h0, w0, g, flip);
        const int Y = h0 + (flip ? q.F - 1 - y : y), X = w0 + x;
        const bool inside = Y < q.inh && X < q.inw;
        const bool m = inside && mask[(int64_t)Y * q.inw + X] != 0;
        const bf16 *o = gout + i * q.Cp;
        for (int c = 0; c < q.ncin; ++c) {
            const int cc = g * q.ncin + c;
            const float s = inside ? (m ? maskValue : __ldg(frames01 + ((int64_t)cc * q.inh + Y) * q.inw + X)) : 0.f;
            const float full = ((s * 2.f - 1.f) + 1.f) * 0.5f;
            const float ov = (__bfloat162float(o[c]) + 1.f) * 0.5f;
            const int64_t idx = ((int64_t)cc * q.outh + Y) * q.outw + X;
            out01[idx] = ov; full01[idx] = full; inpaint01[idx] = m ? ov : full;
        }
    }
}

// ---------------------------------------------------------------- data parallel: sharded gradient reduction + Adam over peer memory
// The two 32.8 M-element generator blocks (E6, G1: 92 % of the gradient bytes) do not go through an all-reduce.  Every rank owns 1/N of each
// block.  After all ranks have converted their local gradient to bf16 (barrier), the owner READS its shard of every peer's gradient over
// NVLink (peer loads), adds in rank order (fp32), applies Adam to its shard of master / m / v, and WRITES the updated bf16 weights into every
// rank's operand copy (peer stores); a second barrier closes the exchange.  NVLink bytes equal a bf16 ring all-reduce's, the Adam pass
// shrinks to 1/N of the block per rank, and the chain behind the last weight gradient is one short kernel instead of two NCCL all-reduces
// and a full Adam pass.  fp32 master / m / v of a shard live on its owner only (cenn_trainer_get_params_host reads them from there).
struct ShardPeers { const bf16 *grad[XR_MAX_WORLD]; bf16 *wbf[XR_MAX_WORLD]; };
__device__ __forceinline__ uint4 ld_cv16(const void *p) {
    uint4 v; asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
// one-CTA barrier across the ranks on a mailbox sequence: everything this rank's stream did before it is visible to a peer that has passed it
__global__ void __launch_bounds__(32) xr_barrier_kernel(const XrCtx x) {
    __shared__ float token;
    if (threadIdx.x == 0) token = 1.f;
    __threadfence_system();
    __syncthreads();
    xr_sum_inplace(&token, 1, x);
}
__global__ void __launch_bounds__(256) shard_reduce_adam_kernel(const ShardPeers pp, int world, float *__restrict__ x, float *__restrict__ gsum, float *__restrict__ m,
        float *__restrict__ v, int64_t begin, int64_t count, float b1, float b2, float eps, const float *__restrict__ step_ptr) {
    // pointers are relative to the block's first element; this rank owns [begin, begin + count), count % 8 == 0
    const float step = *step_ptr;
    const int64_t n8 = count / 8;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n8; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = begin + 8 * j;
        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < XR_MAX_WORLD; ++r)
            if (r < world) {                                   // rank order: every owner adds its shard the same way, replicas of wbf are bit-identical
                const uint4 u = ld_cv16(pp.grad[r] + i);
                float f[8]; unpack8b(u, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] += f[k];
            }
        float X[8], M[8], V[8];
        *reinterpret_cast<float4 *>(X) = *reinterpret_cast<const float4 *>(x + i); *reinterpret_cast<float4 *>(X + 4) = *reinterpret_cast<const float4 *>(x + i + 4);
        *reinterpret_cast<float4 *>(M) = *reinterpret_cast<const float4 *>(m + i); *reinterpret_cast<float4 *>(M + 4) = *reinterpret_cast<const float4 *>(m + i + 4);
        *reinterpret_cast<float4 *>(V) = *reinterpret_cast<const float4 *>(v + i); *reinterpret_cast<float4 *>(V + 4) = *reinterpret_cast<const float4 *>(v + i + 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) { M[k] = M[k] * b1 + (1.f - b1) * g[k]; V[k] = V[k] * b2 + (1.f - b2) * g[k] * g[k]; X[k] -= step * M[k] / (sqrtf(V[k]) + eps); }
        *reinterpret_cast<float4 *>(x + i) = *reinterpret_cast<float4 *>(X); *reinterpret_cast<float4 *>(x + i + 4) = *reinterpret_cast<float4 *>(X + 4);
        *reinterpret_cast<float4 *>(m + i) = *reinterpret_cast<float4 *>(M); *reinterpret_cast<float4 *>(m + i + 4) = *reinterpret_cast<float4 *>(M + 4);
        *reinterpret_cast<float4 *>(v + i) = *reinterpret_cast<float4 *>(V); *reinterpret_cast<float4 *>(v + i + 4) = *reinterpret_cast<float4 *>(V + 4);
        if (gsum) { *reinterpret_cast<float4 *>(gsum + i) = *reinterpret_cast<float4 *>(g); *reinterpret_cast<float4 *>(gsum + i + 4) = *reinterpret_cast<float4 *>(g + 4); }
        __nv_bfloat162 h0 = __floats2bfloat162_rn(X[0], X[1]), h1 = __floats2bfloat162_rn(X[2], X[3]), h2 = __floats2bfloat162_rn(X[4], X[5]), h3 = __floats2bfloat162_rn(X[6], X[7]);
        const uint4 pk = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1), *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
#pragma unroll
        for (int r = 0; r < XR_MAX_WORLD; ++r)
            if (r < world) *reinterpret_cast<uint4 *>(pp.wbf[r] + i) = pk;
    }
}

// The leftover ranges of a gradient vector (everything outside the buckets: ~1.3 M elements of the generator, ~0.15 M of the discriminator) are
// latency-bound: an 8-rank NCCL all-reduce of them costs ~0.1 ms at the step's tail.  Same scheme as above without the optimizer: the ranges
// are flattened into one index space of 16-byte units, rank r owns units [r * per, (r + 1) * per), reads them from every rank (peer loads,
// rank order -> every replica gets bit-identical sums), and stores the sum into every rank's vector (peer stores).  In place: an element is
// read and written by its owner only.  A mailbox barrier before (all gradients complete) and after (all stores landed) on the same stream.
struct PeerVec { float *p[XR_MAX_WORLD]; };
struct PeerSegs { long long off[16], cnt[16]; int n; };      // every off / cnt is a multiple of 4 floats
__global__ void __launch_bounds__(256) peer_allreduce_f32_kernel(const PeerVec pv, const PeerSegs sg, int world, int rank) {
    long long units = 0;
    for (int k = 0; k < sg.n; ++k) units += sg.cnt[k] / 4;
    const long long per = (units + world - 1) / world, u0 = per * rank, u1 = u0 + per < units ? u0 + per : units;
    for (long long u = u0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; u < u1; u += (long long)gridDim.x * blockDim.x) {
        long long r = u; int k = 0;
        while (k < sg.n - 1 && r >= sg.cnt[k] / 4) { r -= sg.cnt[k] / 4; ++k; }
        const long long i = sg.off[k] + 4 * r;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < XR_MAX_WORLD; ++q)
            if (q < world) {
                float4 v;
                asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(pv.p[q] + i) : "memory");
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
#pragma unroll
        for (int q = 0; q < XR_MAX_WORLD; ++q)
            if (q < world) *reinterpret_cast<float4 *>(pv.p[q] + i) = acc;
    }
}
// the 8 loss accumulators (doubles; the outputs are floats): one mailbox exchange instead of an NCCL launch
__global__ void __launch_bounds__(32) losses_xr_kernel(const XrCtx x, double *__restrict__ acc) {
    __shared__ float buf[8];
    if (threadIdx.x < 8) buf[threadIdx.x] = (float)acc[threadIdx.x];
    __syncthreads();
    xr_sum_inplace(buf, 8, x);
    __syncthreads();
    if (threadIdx.x < 8) acc[threadIdx.x] = (double)buf[threadIdx.x];
}

// ---------------------------------------------------------------- train.lua's optional branches (noiseGen / conditionAdv)
// noiseGen (train.lua:109-124): a 1x1 convolution of the noise vector runs beside the encoder and nn.JoinTable(2) appends its nz outputs
// to the nBottleneck encoder outputs before the bottleneck BN.  The joined tensor is the bottleneck block's conv output y [B][pitch]:
// the encoder GEMM writes columns [0, col0), this kernel columns [col0, col0 + nz) and their BN statistics (of the STORED bf16 values,
// as the GEMM epilogues do).  One CTA per output channel (nz is ~100: 2.5 MFLOP at batch 256).
__global__ void __launch_bounds__(128) noise_fwd_kernel(const float *__restrict__ noise, const bf16 *__restrict__ w, const float *__restrict__ bias,
        bf16 *__restrict__ y, int pitch, int col0, int B, int nz, float *__restrict__ stats, int stats_stride) {
    pdl_trigger(); pdl_wait();
    const int j = blockIdx.x;
    __shared__ float red[2][4];
    float s1 = 0.f, s2 = 0.f;
    for (int n = threadIdx.x; n < B; n += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < nz; ++k) acc = fmaf(__bfloat162float(__float2bfloat16(noise[(int64_t)n * nz + k])), __bfloat162float(w[(int64_t)j * nz + k]), acc);
        const bf16 h = __float2bfloat16(acc + bias[j]);
        y[(int64_t)n * pitch + col0 + j] = h;
        const float v = __bfloat162float(h);
        s1 += v; s2 = fmaf(v, v, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0 && stats) {
        atomicAdd(stats + col0 + j, red[0][0] + red[0][1] + red[0][2] + red[0][3]);
        atomicAdd(stats + stats_stride + col0 + j, red[1][0] + red[1][1] + red[1][2] + red[1][3]);
    }
}
// gradWeight[j][k] += sum_n g_y[n][col0 + j] * noise[n][k]   (the branch's gradBias comes from the block's partial column sums)
__global__ void __launch_bounds__(128) noise_wgrad_kernel(const bf16 *__restrict__ g, int pitch, int col0, const float *__restrict__ noise,
        float *__restrict__ gw, int B, int nz) {
    pdl_trigger(); pdl_wait();
    const int j = blockIdx.x;
    extern __shared__ float gcol[];              // [B]
    for (int n = threadIdx.x; n < B; n += blockDim.x) gcol[n] = __bfloat162float(g[(int64_t)n * pitch + col0 + j]);
    __syncthreads();
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
        float acc = 0.f;
        for (int n = 0; n < B; ++n) acc = fmaf(gcol[n], __bfloat162float(__float2bfloat16(noise[(int64_t)n * nz + k])), acc);
        gw[(int64_t)j * nz + k] += acc;
    }
}
// conditionAdv (train.lua:158-180): the discriminator's first layer is a pair of 5x5 / stride-2 convolutions (context: pad 2 on 128x128,
// prediction: pad 34 on 64x64, both giving 64x64 maps) joined along the channel axis.  Both run as GEMMs over an explicit window buffer
// col[m = (n, oy, ox)][k = 4 * (5u + v) + c] of a 4-channel-padded image, K padded from 100 to 128 with zeros.
static const int K5 = 128;
__global__ void __launch_bounds__(256) im2col5_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ col, int N, int H, int W, int OH, int OW, int pad) {
    pdl_trigger(); pdl_wait();
    const int64_t total = (int64_t)N * OH * OW * 16;         // one 16-byte vector (two taps) per thread
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i & 15); int64_t r = i >> 4;
        const int ox = (int)(r % OW), oy = (int)((r / OW) % OH), n = (int)(r / ((int64_t)OW * OH));
        uint2 v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tap = 2 * j + h, u = tap / 5, vv = tap - 5 * u;
            const int iy = 2 * oy - pad + u, ix = 2 * ox - pad + vv;
            v[h] = make_uint2(0u, 0u);
            if (tap < 25 && iy >= 0 && iy < H && ix >= 0 && ix < W) v[h] = __ldg(reinterpret_cast<const uint2 *>(x) + ((int64_t)n * H + iy) * W + ix);
        }
        reinterpret_cast<uint4 *>(col)[i] = make_uint4(v[0].x, v[0].y, v[1].x, v[1].y);
    }
}
// adjoint of im2col5 (gather form): gx[n, iy, ix, c] = sum over the taps (u, v) with iy + pad - u and ix + pad - v even and in range
__global__ void __launch_bounds__(256) col2im5_kernel(const bf16 *__restrict__ col, bf16 *__restrict__ gx, int N, int H, int W, int OH, int OW, int pad) {
    pdl_trigger(); pdl_wait();
    const int64_t total = (int64_t)N * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ix = (int)(i % W), iy = (int)((i / W) % H), n = (int)(i / ((int64_t)W * H));
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int u = (iy + pad) & 1; u < 5; u += 2) {
            const int oy = (iy + pad - u) >> 1;
            if (oy < 0 || oy >= OH) continue;
            for (int v = (ix + pad) & 1; v < 5; v += 2) {
                const int ox = (ix + pad - v) >> 1;
                if (ox < 0 || ox >= OW) continue;
                const uint2 w = __ldg(reinterpret_cast<const uint2 *>(col + (((int64_t)n * OH + oy) * OW + ox) * K5 + 4 * (5 * u + v)));
                acc[0] += __uint_as_float(w.x << 16); acc[1] += __uint_as_float(w.x & 0xFFFF0000u);
                acc[2] += __uint_as_float(w.y << 16); acc[3] += __uint_as_float(w.y & 0xFFFF0000u);
            }
        }
        __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
        reinterpret_cast<uint2 *>(gx)[i] = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
    }
}

}  // namespace nhwc
