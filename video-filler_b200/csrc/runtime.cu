// runtime.cu -- state, errors, storage and the cutorch tensor-math calls the scripts make on GPU
// tensors (train.lua:251-256,267-272,297,320-322,383-398; optim.adam).  All kernels are
// grid-stride, 16-byte vectorised where alignment allows, grid sized in multiples of the SM count.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "map.cuh"

static thread_local char g_err[1024] = "";

void cenn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cenn_check_cuda(cudaError_t e, const char *what, const char *file, int line) {
    if (e == cudaSuccess) return 0;
    cenn_set_error("CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    return 1;
}

void *cenn_workspace(cenn_state *s, size_t bytes) {
    if (bytes <= s->ws_bytes) return s->ws;
    if (s->ws) { cudaStreamSynchronize(s->stream); cudaFree(s->ws); s->ws = nullptr; s->ws_bytes = 0; }
    size_t want = bytes + (bytes >> 2) + (1u << 20);
    if (cudaMalloc(&s->ws, want) != cudaSuccess) { s->ws = nullptr; cenn_set_error("workspace alloc of %zu bytes failed", want); return nullptr; }
    s->ws_bytes = want;
    return s->ws;
}
void *cenn_workspace2(cenn_state *s, size_t bytes) {
    if (bytes <= s->ws2_bytes) return s->ws2;
    if (s->ws2) { cudaStreamSynchronize(s->stream); cudaFree(s->ws2); s->ws2 = nullptr; s->ws2_bytes = 0; }
    size_t want = bytes + (bytes >> 2) + (1u << 20);
    if (cudaMalloc(&s->ws2, want) != cudaSuccess) { s->ws2 = nullptr; cenn_set_error("workspace alloc of %zu bytes failed", want); return nullptr; }
    s->ws2_bytes = want;
    return s->ws2;
}

extern "C" {

const char *cenn_last_error(void) { return g_err; }
const char *cenn_version(void) { return "cenn 0.1 (sm_100a)"; }

int cenn_device_count(int *count) {
    if (!count) { cenn_set_error("null count"); return 1; }
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); cenn_set_error("no CUDA device: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

int cenn_init(int device, cenn_state **out) {
    REQUIRE(out, "null out");
    int n = 0;
    if (cenn_device_count(&n)) return 1;
    REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, "libcenn is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    cenn_state *s = new cenn_state();
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    // the compute stream carries the critical path of the step program: highest priority, so that its (short) kernels are
    // dispatched ahead of the side streams' queued GEMM CTAs
    { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&s->own_stream, cudaStreamNonBlocking, hi)); }
    s->stream = s->own_stream;
    CK(cudaMalloc(&s->red, RED_SLOTS * sizeof(double)));
    CK(cudaMemset(s->red, 0, RED_SLOTS * sizeof(double)));
    CK(cudaMallocHost(&s->red_host, RED_SLOTS * sizeof(double)));
    *out = s;
    return 0;
}

int cenn_shutdown(cenn_state *s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (s->ws) cudaFree(s->ws);
    if (s->ws2) cudaFree(s->ws2);
    if (s->red) cudaFree(s->red);
    if (s->red_host) cudaFreeHost(s->red_host);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
    return 0;
}

int cenn_set_precision(cenn_state *s, int mode) {
    REQUIRE(s, "null state");
    REQUIRE(mode == CENN_FP32 || mode == CENN_BF16, "unknown precision mode %d", mode);
    s->precision = mode;
    return 0;
}
int cenn_get_precision(cenn_state *s, int *mode) { REQUIRE(s && mode, "null arg"); *mode = s->precision; return 0; }
int cenn_set_stream(cenn_state *s, void *st) { REQUIRE(s, "null state"); s->stream = st ? (cudaStream_t)st : s->own_stream; return 0; }
int cenn_get_stream(cenn_state *s, void **st) { REQUIRE(s && st, "null arg"); *st = (void *)s->stream; return 0; }
int cenn_synchronize(cenn_state *s) { API_BEGIN(s); CK(cudaStreamSynchronize(s->stream)); return 0; }
int cenn_kernel_launches(cenn_state *s, int64_t *c) { REQUIRE(s && c, "null arg"); *c = s->launches; return 0; }

int cenn_malloc(cenn_state *s, size_t bytes, void **p) { API_BEGIN(s); REQUIRE(p, "null out"); CK(cudaMalloc(p, bytes ? bytes : 16)); return 0; }
int cenn_free(cenn_state *s, void *p) { API_BEGIN(s); if (p) { CK(cudaStreamSynchronize(s->stream)); CK(cudaFree(p)); } return 0; }
int cenn_host_alloc(cenn_state *s, size_t bytes, void **p) { API_BEGIN(s); REQUIRE(p, "null out"); CK(cudaMallocHost(p, bytes ? bytes : 16)); return 0; }
int cenn_host_free(cenn_state *s, void *p) { API_BEGIN(s); if (p) CK(cudaFreeHost(p)); return 0; }
int cenn_copy_h2d(cenn_state *s, void *d, const void *h, size_t n) {
    API_BEGIN(s); CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s->stream)); CK(cudaStreamSynchronize(s->stream)); return 0;
}
int cenn_copy_d2h(cenn_state *s, void *h, const void *d, size_t n) {
    API_BEGIN(s); CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s->stream)); CK(cudaStreamSynchronize(s->stream)); return 0;
}
int cenn_copy_d2d(cenn_state *s, void *d, const void *src, size_t n) {
    API_BEGIN(s); CK(cudaMemcpyAsync(d, src, n, cudaMemcpyDeviceToDevice, s->stream)); return 0;
}

}  // extern "C"

__global__ void __launch_bounds__(256) u8_to_float_kernel(float *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = i; j < n; j += st) dst[j] = src[j] ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(256) fill_box_kernel(float *__restrict__ x, int64_t N, int64_t C, int64_t H, int64_t W,
        int64_t c0, int64_t c1, int64_t y0, int64_t y1, int64_t x0, int64_t x1, float v) {
    int64_t bw = x1 - x0, bh = y1 - y0, bc = c1 - c0, total = N * bc * bh * bw;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = i; j < total; j += st) {
        int64_t xx = j % bw, t = j / bw, yy = t % bh; t /= bh;
        int64_t cc = t % bc, nn = t / bc;
        x[((nn * C + c0 + cc) * H + y0 + yy) * W + x0 + xx] = v;
    }
}
__global__ void __launch_bounds__(256) crop_kernel(float *__restrict__ dst, const float *__restrict__ src, int64_t NC, int64_t H, int64_t W,
        int64_t y0, int64_t x0, int64_t h, int64_t w) {
    int64_t total = NC * h * w;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = i; j < total; j += st) {
        int64_t xx = j % w, t = j / w, yy = t % h, p = t / h;
        dst[j] = src[(p * H + y0 + yy) * W + x0 + xx];
    }
}

// Philox4x32-10 counter RNG (own implementation; stream = seed, counter = element index / 4)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__global__ void __launch_bounds__(256) rng_kernel(float *__restrict__ x, int64_t n, float a, float b, uint64_t seed, int normal) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    int64_t groups = (n + 3) / 4;
    for (int64_t g = i; g < groups; g += st) {
        uint4 r = philox4x32(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        float u[4] = {(r.x + 0.5f) * 2.3283064e-10f, (r.y + 0.5f) * 2.3283064e-10f, (r.z + 0.5f) * 2.3283064e-10f, (r.w + 0.5f) * 2.3283064e-10f};
        float o[4];
        if (normal) {
            float r0 = sqrtf(-2.0f * logf(u[0])), r1 = sqrtf(-2.0f * logf(u[2]));
            o[0] = r0 * cospif(2.0f * u[1]); o[1] = r0 * sinpif(2.0f * u[1]);
            o[2] = r1 * cospif(2.0f * u[3]); o[3] = r1 * sinpif(2.0f * u[3]);
            for (int k = 0; k < 4; ++k) o[k] = a + b * o[k];
        } else {
            for (int k = 0; k < 4; ++k) o[k] = a + (b - a) * u[k];
        }
        for (int k = 0; k < 4; ++k) if (g * 4 + k < n) x[g * 4 + k] = o[k];
    }
}

extern "C" {

int cenn_fill(cenn_state *s, float *x, int64_t n, float v) { API_BEGIN(s); LAUNCH_MAP1(s, x, n, [v] __device__(float) { return v; }); return 0; }
int cenn_mul(cenn_state *s, float *x, int64_t n, float a) { API_BEGIN(s); LAUNCH_MAP1(s, x, n, [a] __device__(float t) { return t * a; }); return 0; }
int cenn_add_scalar(cenn_state *s, float *x, int64_t n, float a) { API_BEGIN(s); LAUNCH_MAP1(s, x, n, [a] __device__(float t) { return t + a; }); return 0; }
int cenn_sqrt(cenn_state *s, float *x, int64_t n) { API_BEGIN(s); LAUNCH_MAP1(s, x, n, [] __device__(float t) { return sqrtf(t); }); return 0; }
int cenn_axpy(cenn_state *s, float *y, const float *x, int64_t n, float a) {
    API_BEGIN(s); LAUNCH_MAP2(s, y, x, n, [a] __device__(float yy, float xx) { return yy + a * xx; }); return 0;
}
int cenn_cmul(cenn_state *s, float *y, const float *x, int64_t n) {
    API_BEGIN(s); LAUNCH_MAP2(s, y, x, n, [] __device__(float yy, float xx) { return yy * xx; }); return 0;
}
int cenn_masked_fill(cenn_state *s, float *x, const float *mask, int64_t n, float v) {
    API_BEGIN(s); LAUNCH_MAP2(s, x, mask, n, [v] __device__(float xx, float m) { return m != 0.0f ? v : xx; }); return 0;
}
int cenn_addcmul(cenn_state *s, float *y, float a, const float *p, const float *q, int64_t n) {
    API_BEGIN(s); LAUNCH_MAP3(s, y, p, q, n, [a] __device__(float yy, float pp, float qq) { return yy + a * pp * qq; }); return 0;
}
int cenn_addcdiv(cenn_state *s, float *y, float a, const float *p, const float *q, int64_t n) {
    API_BEGIN(s); LAUNCH_MAP3(s, y, p, q, n, [a] __device__(float yy, float pp, float qq) { return yy + a * pp / qq; }); return 0;
}
int cenn_u8_to_float(cenn_state *s, float *dst, const uint8_t *src, int64_t n) {
    API_BEGIN(s);
    if (n > 0) { u8_to_float_kernel<<<bw_grid(s, n, 256), 256, 0, s->stream>>>(dst, src, n); CK_LAUNCH(s); }
    return 0;
}
int cenn_fill_box(cenn_state *s, float *x, int64_t N, int64_t C, int64_t H, int64_t W,
                  int64_t c0, int64_t c1, int64_t y0, int64_t y1, int64_t x0, int64_t x1, float v) {
    API_BEGIN(s);
    REQUIRE(0 <= c0 && c0 <= c1 && c1 <= C && 0 <= y0 && y0 <= y1 && y1 <= H && 0 <= x0 && x0 <= x1 && x1 <= W, "fill_box: box out of range");
    int64_t total = N * (c1 - c0) * (y1 - y0) * (x1 - x0);
    if (total > 0) { fill_box_kernel<<<bw_grid(s, total, 256), 256, 0, s->stream>>>(x, N, C, H, W, c0, c1, y0, y1, x0, x1, v); CK_LAUNCH(s); }
    return 0;
}
int cenn_crop(cenn_state *s, float *dst, const float *src, int64_t N, int64_t C, int64_t H, int64_t W,
              int64_t y0, int64_t x0, int64_t h, int64_t w) {
    API_BEGIN(s);
    REQUIRE(y0 >= 0 && x0 >= 0 && y0 + h <= H && x0 + w <= W, "crop: window out of range");
    int64_t total = N * C * h * w;
    if (total > 0) { crop_kernel<<<bw_grid(s, total, 256), 256, 0, s->stream>>>(dst, src, N * C, H, W, y0, x0, h, w); CK_LAUNCH(s); }
    return 0;
}
// nn.JoinTable(2): strided device-to-device copies (one 2-D copy per table member; no kernel)
int cenn_JoinTable_updateOutput(cenn_state *s, float *joined, const float *part, int64_t batch, int64_t joined_per_sample, int64_t offset, int64_t part_per_sample) {
    API_BEGIN(s);
    REQUIRE(joined && part && batch >= 0 && offset >= 0 && part_per_sample >= 0 && offset + part_per_sample <= joined_per_sample, "JoinTable: slice [%lld, %lld) outside %lld",
            (long long)offset, (long long)(offset + part_per_sample), (long long)joined_per_sample);
    if (batch == 0 || part_per_sample == 0) return 0;
    CK(cudaMemcpy2DAsync(joined + offset, (size_t)joined_per_sample * 4, part, (size_t)part_per_sample * 4, (size_t)part_per_sample * 4, (size_t)batch, cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}
int cenn_JoinTable_updateGradInput(cenn_state *s, const float *gradJoined, float *gradPart, int64_t batch, int64_t joined_per_sample, int64_t offset, int64_t part_per_sample) {
    API_BEGIN(s);
    REQUIRE(gradJoined && gradPart && batch >= 0 && offset >= 0 && part_per_sample >= 0 && offset + part_per_sample <= joined_per_sample, "JoinTable: slice [%lld, %lld) outside %lld",
            (long long)offset, (long long)(offset + part_per_sample), (long long)joined_per_sample);
    if (batch == 0 || part_per_sample == 0) return 0;
    CK(cudaMemcpy2DAsync(gradPart, (size_t)part_per_sample * 4, gradJoined + offset, (size_t)joined_per_sample * 4, (size_t)part_per_sample * 4, (size_t)batch, cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}
int cenn_normal(cenn_state *s, float *x, int64_t n, float mean, float std, uint64_t seed) {
    API_BEGIN(s);
    if (n > 0) { rng_kernel<<<bw_grid(s, n / 4 + 1, 256), 256, 0, s->stream>>>(x, n, mean, std, seed, 1); CK_LAUNCH(s); }
    return 0;
}
int cenn_uniform(cenn_state *s, float *x, int64_t n, float a, float b, uint64_t seed) {
    API_BEGIN(s);
    if (n > 0) { rng_kernel<<<bw_grid(s, n / 4 + 1, 256), 256, 0, s->stream>>>(x, n, a, b, seed, 0); CK_LAUNCH(s); }
    return 0;
}

}  // extern "C"
