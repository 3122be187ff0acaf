// probes.cu -- bring-up / measurement probes (tools/*probe*.py).  NOT part of the product path: nothing in the executor or the op-level
// ABI calls them.  Declared in include/cenn_debug.h.
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "conv_tc.h"
#include "tc_gemm.cuh"
#include "../../include/cenn_debug.h"

int tc_make_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box);
int tc_map_2d(CUtensorMap *m, const bf16 *B, uint64_t K, uint64_t rows, int box_rows);
#define make_map tc_make_map
#define map_2d tc_map_2d

// ------------------------------------------------------------------ debug probe (tools/gemm_probe.py)
// Runs one primitive on zero-filled operands `iters` times; returns the mean duration and CTA 0's per-role cycle counters:
//   dbg[0] producer wait(empty) [1] producer total | [2] MMA wait(full) [3] MMA wait(tmem empty) [4] MMA total |
//   [5] epilogue wait(tmem full) [6] epilogue total [7] tiles of CTA 0 [8] epilogue tcgen05.ld cycles
extern "C" int cenn_debug_gemm_probe(cenn_state *s, int kind, int N, int h, int w, int Cs, int Cl, int with_stats, int act,
                                              int iters, float *ms_out, unsigned long long *dbg_out) {
    API_BEGIN(s);
    size_t nL, nS, nW;
    int Csp = round_up(Cs, 8), Clp = round_up(Cl, 64);
    if (kind == 0) { nL = (size_t)N * 4 * h * w * Clp; nS = (size_t)N * h * w * Csp; nW = (size_t)Cs * 16 * Clp; }          // fprop_s2
    else if (kind == 1) { Csp = round_up(Cs, 64); Clp = round_up(Cl, 8); nL = (size_t)N * 4 * h * w * Clp; nS = (size_t)N * h * w * Csp; nW = (size_t)16 * Clp * Csp; }  // dgrad_s2
    else { nL = (size_t)N * Cl; nS = (size_t)N * Csp; nW = (size_t)Cs * Cl; }                                              // gemm: M=N, K=Cl, Nc=Cs
    bf16 *L, *S, *W; float *stats; unsigned long long *dbg;
    CK(cudaMalloc(&L, nL * 2)); CK(cudaMalloc(&S, nS * 2)); CK(cudaMalloc(&W, nW * 2)); CK(cudaMalloc(&stats, 2 * 4096 * 4)); CK(cudaMalloc(&dbg, 16 * 8));
    CK(cudaMemset(L, 0, nL * 2)); CK(cudaMemset(S, 0, nS * 2)); CK(cudaMemset(W, 0, nW * 2)); CK(cudaMemset(stats, 0, 2 * 4096 * 4)); CK(cudaMemset(dbg, 0, 16 * 8));
    TcEpilogue ep; ep.act = act; ep.act_param = 0.2f; ep.dbg = getenv("PROBE_NO_DBG") ? nullptr : dbg;
    if (getenv("PROBE_NO_OUT")) ep.no_bf16 = true;
    if (getenv("PROBE_FLAGS")) ep.dbg_flags = atoi(getenv("PROBE_FLAGS"));
    if (with_stats) { ep.stats = stats; ep.stats_stride = 4096; }
    TcPlan pl; int rc;
    if (kind == 0) rc = tc_plan_fprop_s2(s, &pl, L, W, S, N, h, w, Cs, Csp, Clp, ep);
    else if (kind == 1) rc = tc_plan_dgrad_s2(s, &pl, S, W, L, N, h, w, Csp, Cl, Clp, Clp, ep);
    else rc = tc_plan_gemm(s, &pl, L, W, S, N, Cs, Cl, Csp, ep);
    if (rc) return 1;
    if (getenv("PROBE_STAGES")) { int st = atoi(getenv("PROBE_STAGES")); if (st >= 1 && st < pl.stages) pl.stages = st; }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) if (tc_launch(s, &pl)) return 1;
    CK(cudaEventRecord(e0, s->stream));
    for (int i = 0; i < iters; ++i) if (tc_launch(s, &pl)) return 1;
    CK(cudaEventRecord(e1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    CK(cudaMemcpy(dbg_out, dbg, 16 * 8, cudaMemcpyDeviceToHost));
    dbg_out[15] = ((unsigned long long)pl.grid[0] << 32) | (unsigned)(pl.BN << 8) | (unsigned)pl.stages;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(L); cudaFree(S); cudaFree(W); cudaFree(stats); cudaFree(dbg);
    return 0;
}

// ------------------------------------------------------------------ descriptor probe (tools/desc_probe.py)
// What does a UMMA shared-memory descriptor (K-major, SWIZZLE_128B) read when its start address is a whole number of
// 128-byte rows past the 1024-byte swizzle atom, and when the stride between 8-row groups (SBO) is not a multiple of
// 1024?  A [160 x 64] patch is written by ONE TMA box (so the swizzle phase of every row follows its absolute
// address); B is a 64 x 64 identity, so D[r][n] = A_patch[source_row(r)][n]: the output shows which element was read.
namespace {
__global__ void __launch_bounds__(128, 1)
desc_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *out, int start_row, int base_off, int sbo_bytes) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *a_s = smem, *b_s = smem + 160 * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(b_s + 64 * 128);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(&bar[0], 160 * 128 + 64 * 128);
        tc::tma_load_2d(&tmA, &bar[0], a_s, 0, 0);
        tc::tma_load_2d(&tmB, &bar[0], b_s, 0, 0);
        tc::mbar_wait(&bar[0], 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc(128, 64, 0, 0);
        const uint32_t a_addr = tc::smem_u32(a_s) + (uint32_t)start_row * 128u, b_addr = tc::smem_u32(b_s);
        for (int k = 0; k < 4; ++k) {
            uint64_t ad = tc::make_desc(a_addr + k * 32, 16, (uint32_t)sbo_bytes) | ((uint64_t)(base_off & 7) << 49);
            uint64_t bd = tc::make_desc(b_addr + k * 32, 16, 1024);
            tc::umma_f16(tmem, ad, bd, idesc, k != 0);
        }
        tc::umma_commit(&bar[1]);
    }
    __syncwarp();
    tc::mbar_wait(&bar[1], 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 64); }
}
}  // namespace

// a_host: [160][64] bf16 bit patterns, out_host: [128][64] fp32
extern "C" int cenn_debug_desc_probe(cenn_state *s, const uint16_t *a_host, int start_row, int base_off, int sbo_bytes, float *out_host) {
    API_BEGIN(s);
    bf16 *A, *B; float *out;
    CK(cudaMalloc(&A, 160 * 64 * 2)); CK(cudaMalloc(&B, 64 * 64 * 2)); CK(cudaMalloc(&out, 128 * 64 * 4));
    std::vector<uint16_t> eye(64 * 64, 0);
    for (int i = 0; i < 64; ++i) eye[i * 64 + i] = 0x3F80;   // bf16 1.0
    CK(cudaMemcpy(A, a_host, 160 * 64 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, eye.data(), 64 * 64 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(out, 0, 128 * 64 * 4));
    CUtensorMap ta, tb;
    { uint64_t d[2] = {64, 160}, st[1] = {128}; uint32_t bx[2] = {64, 160}; if (make_map(&ta, A, 2, d, st, bx)) return 1; }
    { uint64_t d[2] = {64, 64}, st[1] = {128}; uint32_t bx[2] = {64, 64}; if (make_map(&tb, B, 2, d, st, bx)) return 1; }
    size_t smem = 1024 + 160 * 128 + 64 * 128 + 64;
    desc_probe_kernel<<<1, 128, smem, s->stream>>>(ta, tb, out, start_row, base_off, sbo_bytes);
    CK_LAUNCH(s);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(out_host, out, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    cudaFree(A); cudaFree(B); cudaFree(out);
    return 0;
}

// ------------------------------------------------------------------ CTA-pair GEMM probe (tools/gemm2sm_probe.py)
// EXPERIMENTAL bring-up vehicle for the next round's kernels; no product path calls it, and at the end
// of round 1 it had only been through ptxas, not through a GPU.  C[M,N] (fp32) = A[M,K] * B[N,K]^T, bf16 operands, K-major,
// M % 256 == 0, N % 256 == 0, K % 64 == 0.  One cluster of two CTAs owns a 256 x 256 tile: each CTA loads ITS 128 rows of A
// and ITS 128 rows of B per k-block, the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256) and every
// CTA reads its own 128 accumulator rows back from its own TMEM.
//   full[s]  (leader's copy only, 2 arrivals + 64 KB of transactions): leader arrive.expect_tx + peer remote arrive; both CTAs'
//            TMA loads complete on it (cta_group::2 load form, barrier address with the peer bit cleared)
//   empty[s] (one per CTA, 1 arrival): tcgen05.commit.cta_group::2 ... multicast::cluster, mask 0b11
//   done     (one per CTA, 1 arrival): same multicast commit after the last k-block
namespace {
constexpr int P2_STAGES = 4;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap *m, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm2sm_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *__restrict__ C, int M, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *a_s = smem, *b_s = smem + P2_STAGES * 16384;
    uint64_t *full = reinterpret_cast<uint64_t *>(b_s + P2_STAGES * 16384), *empty = full + P2_STAGES, *done = empty + P2_STAGES;
    uint32_t *slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int m0 = (blockIdx.x >> 1) * 256 + (int)rank * 128, n0 = blockIdx.y * 256 + (int)rank * 128, num_kb = K / 64;
    if (threadIdx.x == 0) {
        for (int i = 0; i < P2_STAGES; ++i) { tc::mbar_init(&full[i], 2); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(done, 1);
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmA); tc::prefetch_tmap(&tmB);
    }
    if (warp == 1) tmem_alloc_2sm(slot, 256);          // one warp of EACH CTA of the pair
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // the peer's barriers are initialised before anyone signals them
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        if (tc::elect_one()) {                         // ---- TMA producer (both CTAs)
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % P2_STAGES; const uint32_t ph = (kb / P2_STAGES) & 1;
                tc::mbar_wait(&empty[st], ph ^ 1);
                const uint32_t leader_full = tc::smem_u32(&full[st]) & 0xFEFFFFFFu;      // same offset in the even CTA of the pair
                if (leader) tc::mbar_expect_tx(&full[st], 2 * 32768); else mbar_arrive_remote(tc::smem_u32(&full[st]), 0);
                tma_load_2d_2sm(&tmA, leader_full, tc::smem_u32(a_s + st * 16384), kb * 64, m0);
                tma_load_2d_2sm(&tmB, leader_full, tc::smem_u32(b_s + st * 16384), kb * 64, n0);
            }
        }
    } else if (warp == 1) {
        if (leader && tc::elect_one()) {               // ---- MMA issuer (leader CTA only)
            const uint32_t idesc = tc::make_idesc(256, 256, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % P2_STAGES; const uint32_t ph = (kb / P2_STAGES) & 1;
                tc::mbar_wait(&full[st], ph);
                tc::tc_fence_after();
                const uint32_t a_addr = tc::smem_u32(a_s + st * 16384), b_addr = tc::smem_u32(b_s + st * 16384);
                for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(tmem, tc::make_desc(a_addr + k * 32, 16, 1024), tc::make_desc(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0);
                umma_commit_2sm(tc::smem_u32(&empty[st]), 3);
            }
            umma_commit_2sm(tc::smem_u32(done), 3);
        }
    } else {                                           // ---- epilogue: warps 2..5 of both CTAs, 32 accumulator rows each
        const int q = warp & 3;
        tc::mbar_wait(done, 0);
        tc::tc_fence_after();
        float *crow = C + (size_t)(m0 + q * 32 + lane) * N + blockIdx.y * 256;
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t r[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) crow[c0 + j] = __uint_as_float(r[j]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // nobody frees TMEM while the peer may still be read by an MMA
    if (warp == 1) { tc::tc_fence_after(); tmem_dealloc_2sm(tmem, 256); }
}
}  // namespace

// a_host [M][K], b_host [N][K] bf16 bit patterns; c_host [M][N] fp32; ms_out = mean kernel time over `iters` launches
extern "C" int cenn_debug_gemm2sm_probe(cenn_state *s, const uint16_t *a_host, const uint16_t *b_host, int M, int N, int K, int iters,
                                                 float *c_host, float *ms_out) {
    API_BEGIN(s);
    REQUIRE(a_host && b_host && c_host && ms_out && M > 0 && N > 0 && K > 0 && M % 256 == 0 && N % 256 == 0 && K % 64 == 0, "gemm2sm probe: M, N multiples of 256 and K of 64 required");
    bf16 *A, *B; float *Cd;
    CK(cudaMalloc(&A, (size_t)M * K * 2)); CK(cudaMalloc(&B, (size_t)N * K * 2)); CK(cudaMalloc(&Cd, (size_t)M * N * 4));
    CK(cudaMemcpy(A, a_host, (size_t)M * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, b_host, (size_t)N * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(Cd, 0, (size_t)M * N * 4));
    CUtensorMap ta, tb;
    if (map_2d(&ta, A, (uint64_t)K, (uint64_t)M, 128) || map_2d(&tb, B, (uint64_t)K, (uint64_t)N, 128)) return 1;
    const size_t smem = 1024 + 2 * P2_STAGES * 16384 + (2 * P2_STAGES + 1) * 8 + 16;
    CK(cudaFuncSetAttribute(gemm2sm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(2 * (M / 256), N / 256);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gemm2sm_probe_kernel<<<grid, 192, smem, s->stream>>>(ta, tb, Cd, M, N, K);
    CK_LAUNCH(s);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaEventRecord(e0, s->stream));
    for (int i = 0; i < iters; ++i) { gemm2sm_probe_kernel<<<grid, 192, smem, s->stream>>>(ta, tb, Cd, M, N, K); CK_LAUNCH(s); }
    CK(cudaEventRecord(e1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    float ms = 0.f; if (iters > 0) CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = iters > 0 ? ms / iters : 0.f;
    CK(cudaMemcpy(c_host, Cd, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(A); cudaFree(B); cudaFree(Cd);
    return 0;
}

// ------------------------------------------------------------------ TMA box probe (tools/tma_thin_probe.py)
// EXPERIMENT that did NOT make it into the product (round 2): implicit im2col of the 3 / 12-channel first layers through a tiled tensor map
// with OVERLAPPING strides, so that the explicit col buffer (16x the input) disappears.  Measured on B200 / driver 580 / CUDA 12.9
// (profiles/r2_tma_thin_probe.log): the Cp = 4 map (inner box 32 B, stride of dimension 1 larger than that of dimension 2) raises an illegal
// address on its first box; the Cp = 16 map (monotonic strides, windows overlapping along x) returns wrong bytes for boxes of one or two
// rows.  cuTensorMapEncodeTiled accepts both maps.  The executor therefore keeps the explicit im2col for these layers.

// Loads ONE box of the implicit-im2col map of a bordered thin tensor (conv_tc.cu: map_thin_gather) into shared memory and copies the
// raw (still swizzled) bytes out: the host un-swizzles and compares with the im2col rows it expects.
// Implicit im2col of a THIN large-side tensor stored with a physical one-pixel zero border, Lpad [N, H2+2, W2+2, Cp] (Cp = 4 or 16):
// the 4x4 / stride-2 / pad-1 window of output pixel (ox, oy) is rows 2oy .. 2oy+3, pixels 2ox .. 2ox+3 of the padded tensor, and one
// window row (4 pixels x Cp channels) is contiguous in memory.  TMA strides may overlap, so the window becomes a box:
//   Cp == 4 : dims (16 [v,c], 4 [u], w [ox, stride 2 px], h [oy, stride 2 rows], N), box (16, 4, bw, bh, bn): one 128-byte K row (u,v,c)
//             per output pixel -- the whole K = 64 in ONE k-block (coordinates in `a_order`: (k0, u0, x0, y0, n0));
//   Cp == 16: dims (64 [v,c], w, 4 [u, stride 1 row], h, N), box (64, bw, 1, bh, bn): one k-block per window row u (existing order).
static int tc_map_thin_gather(CUtensorMap *m, const bf16 *Lpad, int N, int H2, int W2, int Cp, int bw, int bh, int bn) {
    const uint64_t px = (uint64_t)Cp * 2, row = (uint64_t)(W2 + 2) * px, img = (uint64_t)(H2 + 2) * row;
    if (Cp == 4) {
        uint64_t dims[5] = {16, 4, (uint64_t)W2 / 2, (uint64_t)H2 / 2, (uint64_t)N};
        uint64_t st[4] = {row, 2 * px, 2 * row, img};
        uint32_t box[5] = {16, 4, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
        return tc_make_map(m, Lpad, 5, dims, st, box);
    }
    uint64_t dims[5] = {64, (uint64_t)W2 / 2, 4, (uint64_t)H2 / 2, (uint64_t)N};
    uint64_t st[4] = {2 * px, row, 2 * row, img};
    uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bn};
    return tc_make_map(m, Lpad, 5, dims, st, box);
}
namespace {
__global__ void __launch_bounds__(128, 1) tma_box_probe_kernel(const __grid_constant__ CUtensorMap tm, uint8_t *out, int bytes, int a_order, int c1, int x0, int y0, int n0) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 32768);
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar, (uint32_t)bytes);
        if (a_order) tc::tma_load_5d(&tm, bar, smem, 0, c1, x0, y0, n0); else tc::tma_load_5d(&tm, bar, smem, 0, x0, c1, y0, n0);
    }
    tc::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
}  // namespace
extern "C" int cenn_debug_tma_thin_probe(cenn_state *s, const uint16_t *lpad_host, int N, int H2, int W2, int Cp, int bw, int bh, int bn,
                                         int c1, int x0, int y0, int n0, uint8_t *out_host) {
    API_BEGIN(s);
    const size_t n = (size_t)N * (H2 + 2) * (W2 + 2) * Cp;
    const int bytes = bw * bh * bn * 128;
    REQUIRE(bytes <= 32768, "tma probe: box larger than 32 KB");
    bf16 *L; uint8_t *out;
    CK(cudaMalloc(&L, n * 2)); CK(cudaMalloc(&out, bytes));
    CK(cudaMemcpy(L, lpad_host, n * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(out, 0xEE, bytes));
    CUtensorMap tm;
    if (tc_map_thin_gather(&tm, L, N, H2, W2, Cp, bw, bh, bn)) return 1;
    tma_box_probe_kernel<<<1, 128, 1024 + 32768 + 64, s->stream>>>(tm, out, bytes, Cp == 4 ? 1 : 0, c1, x0, y0, n0);
    CK_LAUNCH(s);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(out_host, out, bytes, cudaMemcpyDeviceToHost));
    cudaFree(L); cudaFree(out);
    return 0;
}
