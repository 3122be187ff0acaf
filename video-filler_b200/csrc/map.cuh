// map.cuh -- grid-stride, float4-vectorised elementwise kernel templates shared by the bandwidth ops.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------ elementwise tensor math
template <typename F>
__global__ void __launch_bounds__(256) map1_kernel(float *__restrict__ x, int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
    float4 *x4 = reinterpret_cast<float4 *>(x);
    for (int64_t j = i; j < n4; j += st) {
        float4 v = x4[j];
        v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
        x4[j] = v;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += st) x[j] = f(x[j]);
}
template <typename F>
__global__ void __launch_bounds__(256) map2_kernel(float *__restrict__ y, const float *__restrict__ x, int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    int64_t n4 = al ? n / 4 : 0;
    float4 *y4 = reinterpret_cast<float4 *>(y);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (int64_t j = i; j < n4; j += st) {
        float4 a = y4[j], b = x4[j];
        a.x = f(a.x, b.x); a.y = f(a.y, b.y); a.z = f(a.z, b.z); a.w = f(a.w, b.w);
        y4[j] = a;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += st) y[j] = f(y[j], x[j]);
}
template <typename F>
__global__ void __launch_bounds__(256) map3_kernel(float *__restrict__ y, const float *__restrict__ p, const float *__restrict__ q, int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = i; j < n; j += st) y[j] = f(y[j], p[j], q[j]);
}

#define LAUNCH_MAP1(s, x, n, ...) do { if ((n) > 0) { map1_kernel<<<bw_grid(s, (n) / 4 + 1, 256), 256, 0, (s)->stream>>>(x, n, __VA_ARGS__); CK_LAUNCH(s); } } while (0)
#define LAUNCH_MAP2(s, y, x, n, ...) do { if ((n) > 0) { map2_kernel<<<bw_grid(s, (n) / 4 + 1, 256), 256, 0, (s)->stream>>>(y, x, n, __VA_ARGS__); CK_LAUNCH(s); } } while (0)
#define LAUNCH_MAP3(s, y, p, q, n, ...) do { if ((n) > 0) { map3_kernel<<<bw_grid(s, (n), 256), 256, 0, (s)->stream>>>(y, p, q, n, __VA_ARGS__); CK_LAUNCH(s); } } while (0)

