// tc_gemm.cuh -- hand-written sm_100a implicit-GEMM kernels: TMA (cp.async.bulk.tensor) stages NHWC
// bf16 tiles into 128B-swizzled shared memory, one elected thread issues tcgen05.mma (kind::f16,
// cta_group::1, 128 x N x 16) into a TMEM accumulator, four epilogue warps read it back with
// tcgen05.ld.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue.
//
//  * gather_gemm_kernel  (both operands K-major): conv fprop-type, conv dgrad-type (4 sub-pixel phases
//    in grid.z) and plain GEMMs.  A tile = one TMA box of 128 pixels x 64 channels per k-block, taken from a
//    5-D view of the activation tensor at (tile origin + per-k-block offset); zero padding = TMA OOB fill.
//  * wgrad_gemm_kernel   (both operands MN-major): gW[cs][tap][cl] = sum_pix S[pix,cs] * L[gather_tap(pix),cl];
//    K runs over pixels, split across CTAs (grid.z), fp32 result stored or atomically accumulated.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace tc {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("cenn: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
    }
}
// shared-space (32-bit) address forms: the single-thread pipeline loops keep their barrier / tile addresses as integers
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
static __device__ __noinline__ void mbar_wait_slow_u32(uint32_t bar, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try_wait_u32(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("cenn: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_u32(bar, parity)) return;
    if (mbar_try_wait_u32(bar, parity)) return;
    mbar_wait_slow_u32(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__device__ __forceinline__ void tma_load_2d_u32(const CUtensorMap *m, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_5d_u32(const CUtensorMap *m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_4d_u32(const CUtensorMap *m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *m, const void *src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (SM100 "version 1"), SWIZZLE_128B; addresses/offsets in 16-byte units.
//   K-major : rows of 128 B (64 bf16 of K), 8-row atoms of 1024 B -> SBO = 1024, LBO unused (1)
//   MN-major: rows of 128 B (64 bf16 of M/N) per K index, 8 K-rows = 1024 B -> SBO = 1024,
//             LBO = byte distance between consecutive 64-element M/N chunks
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;   // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D fp32, A/B bf16
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------ kernel parameters
enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };

// K-block kb of a tile = (tap = kb / chunks, c = kb % chunks).  TMA coordinates (all tables live in kernel parameters):
//   A box: (A0[ph][tap] + 64 c,  x0 + A1[ph][tap],  A2[ph][tap],  y0 + A3[ph][tap],  n0)
//   B box: (tap * bk_per_tap + 64 c,  n_tile * BN + B1[ph])
//   O box: (O0[ph] + n_tile * BN + 64 slab,  x0,  O2[ph],  y0,  n0)        (TMA store of the bf16 output tile, BN >= 64)
struct GatherGemmParams {
    int num_kb, chunks, bk_per_tap, num_taps;
    int A0[4][16], A1[4][16], A2[4][16], A3[4][16];
    int B1[4];
    int b_mn;                // 1: B is read MN-major straight from the master layout Wf [K rows][N cols] (no transposed copy):
    int Bc[4][16];           //    box (64 n, 64 k) at column Bc[ph][tap] + n_tile*BN + 64*chunk, row 64*c
    int O0[4], O2[4];
    int o_cols;              // columns of one phase in the output map's dim 0 (slabs starting at or beyond it are not stored)
    int m_tiles, n_tiles, num_phases;   // tiles are enumerated m fastest, then n, then phase
    int box_w, box_h, box_n; // pixels per M tile (product == 128), all powers of two
    int bw_log2, bh_log2;
    int tiles_x, tiles_y;    // m-tile -> (tx, ty, tn)
    int out_w, out_h, out_n; // logical extent of the output pixel grid (for masking)
    int n_valid;             // valid output channels (columns >= n_valid are written as zero / dropped)
    // output addressing of the direct-store path (BN == 32): elem offset = n*sN + y*sY + x*sX + phase_off[phase] + column
    long long sN, sY, sX;
    long long phase_off[4];
    __nv_bfloat16 *out_bf16; // [opt] bf16 output (NULL: results are discarded; probe only)
    const float *bias;       // [opt] per-column
    float *stats;            // [opt] per-column sum / sum of squares (fp32 atomics): stats[c], stats[stats_stride + c]
    int stats_stride;
    int act;                 // applied after bias, before the store (stats are taken before the activation)
    float act_param;
    // [opt] BN-backward sums of the CONSUMER of this output: the output is dLoss/d(activation output) of a BN + activation block whose conv
    // output is bwd_y (same geometry as the output).  With dz = out * act'(y * scale + shift):  stats[c] += sum dz,
    // stats[stats_stride + c] += sum dz * (y - mean)  -- the first pass of the BN backward, taken from the tile while it is in shared memory
    const __nv_bfloat16 *bwd_y;
    const float *bwd_scale, *bwd_shift, *bwd_mean;
    int bwd_act;
    float bwd_negval;
    unsigned long long *dbg; // [opt] per-role cycle counters of CTA 0 (tools/gemm_probe.py)
    int dbg_flags;           // probe only: 1 = issue no MMAs, 2 = no A loads, 4 = no B loads (isolates the TMA feed / the MMA rate)
};

struct WgradParams {
    int num_kb_total;        // pixel blocks (64 pixels each)
    int kb_per_split;
    int box_w, box_h, box_n; // pixels per k-block (product == 64)
    int tiles_x, tiles_y;    // pixel-block grid: kb -> (tx, ty, tn)
    int chunks_per_tap;      // Cl / 64 (row-blocks per tap)
    int num_taps;            // taps in the output (16 for 4x4 windows, 1 for plain GEMM)
    int cl_stride;           // row length of one tap in the output (= Cl padded, master layout [cs][tap][cl])
    int cl_valid;            // valid cl per tap
    int cs_valid;
    int n_chunks;            // cs chunks (64 each) in this N tile = BN/64
    long long out_cs_stride; // taps * cl_stride
    float *out;              // fp32 [cs][tap][cl]
    float scale;
    int accumulate;          // 1: red.add (accumulate / split-K), 0: plain store (requires a single split)
    // gather geometry for L per tap: coords = (chunk*64 + gx0[tap], x0 + gdx[tap], g2[tap], y0 + gdy[tap], n0)
    int gx0[16], gdx[16], g2[16], gdy[16];
};

static constexpr int GEMM_THREADS = 192;

// ------------------------------------------------------------------ K-major gather GEMM (persistent)
// One CTA per SM loops over output tiles (static round-robin).  Three pipelines run concurrently:
//   TMA producer (1 thread) -> smem ring (full/empty mbarriers)      -> MMA issuer (1 thread)
//   MMA issuer              -> 2 TMEM accumulators (tmem_full/empty) -> 4 epilogue warps
//   epilogue warps          -> bf16 tile in swizzled smem            -> TMA store (bulk async group)
// so the epilogue of tile t overlaps the MMAs of tile t+1 and its global stores overlap the epilogue of tile t+1.
// The two single-thread loops are kept to a few dozen instructions per k-block (descriptors advanced by integer adds,
// tap tables in shared memory): at N = 64 the four MMAs of a k-block occupy the tensor pipe for only 128 cycles.
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct TileCoord { int phase, nt, x0, y0, n0; };
__device__ __forceinline__ TileCoord tile_coord(const GatherGemmParams &p, int t, int tiles_per_phase) {
    TileCoord c;
    c.phase = t / tiles_per_phase;
    const int r = t - c.phase * tiles_per_phase;
    c.nt = r / p.m_tiles;
    const int mt = r - c.nt * p.m_tiles;
    const int txy = p.tiles_x * p.tiles_y;
    const int tn = mt / txy, rem = mt - tn * txy;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    c.x0 = tx * p.box_w; c.y0 = ty * p.box_h; c.n0 = tn * p.box_n;
    return c;
}

// MUFU.TANH: 2^-11 relative error, below the bf16 rounding of the stored result (the epilogue warps are latency-bound)
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float apply_act(float f, int act, float param) {
    if (act == ACT_LEAKY) return f > 0.f ? f : f * param;
    if (act == ACT_RELU) return fmaxf(f, 0.f);
    if (act == ACT_TANH) return tanh_fast(f);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-f));
    return f;
}

// derivative of the consumer's activation from its pre-activation z (nhwc::dact_z): LeakyReLU / ReLU only (the BN blocks' activations)
__device__ __forceinline__ float bwd_gate(float z, int act, float negval) { return z > 0.f ? 1.f : (act == ACT_LEAKY ? negval : 0.f); }
// one slab (64 columns) of the BN-backward sums over this warp's 32 tile rows: lane l owns columns 2l, 2l+1; `my_off` is this lane's OWN row
// offset into y (elements, first column of the n-tile; < 0: row outside the tensor); g comes from the staged bf16 tile
__device__ __forceinline__ void bwd_sums_slab(const uint8_t *slab, const __nv_bfloat16 *y, long long my_off, int cc, const float *bwd_c, int BNc,
        int act, float negval, int lane, float *col_acc) {
    const float sc_a = bwd_c[cc], sc_b = bwd_c[cc + 1], sh_a = bwd_c[BNc + cc], sh_b = bwd_c[BNc + cc + 1], mu_a = bwd_c[2 * BNc + cc], mu_b = bwd_c[2 * BNc + cc + 1];
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
        const long long off = __shfl_sync(0xffffffffu, my_off, r);
        if (off < 0) continue;                          // warp-uniform
        const uint32_t wv = *reinterpret_cast<const uint32_t *>(slab + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + ((lane & 3) << 2));
        const uint32_t yv = __ldg(reinterpret_cast<const unsigned int *>(y + off + cc));
        const float ga = __uint_as_float(wv << 16), gb = __uint_as_float(wv & 0xFFFF0000u);
        const float ya = __uint_as_float(yv << 16), yb = __uint_as_float(yv & 0xFFFF0000u);
        const float dza = ga * bwd_gate(fmaf(ya, sc_a, sh_a), act, negval), dzb = gb * bwd_gate(fmaf(yb, sc_b, sh_b), act, negval);
        s1a += dza; s2a = fmaf(dza, ya - mu_a, s2a); s1b += dzb; s2b = fmaf(dzb, yb - mu_b, s2b);
    }
    atomicAdd(&col_acc[cc], s1a); atomicAdd(&col_acc[cc + 1], s1b);
    atomicAdd(&col_acc[BNc + cc], s2a); atomicAdd(&col_acc[BNc + cc + 1], s2b);
}

template <int BN, bool kProbe>
__global__ void __launch_bounds__(GEMM_THREADS, BN <= 128 ? 2 : 1)
gather_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ GatherGemmParams p, int stages) {
    // kProbe = true compiles in the per-role cycle counters and the stream-dropping flags of tools/gemm_probe.py; the
    // production instantiation carries none of it (the two single-thread loops are instruction-issue bound).
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr bool kTma = BN >= 64;               // bf16 tile -> swizzled smem -> TMA store; BN == 32 stores directly (thin outputs)
    constexpr uint32_t A_BYTES = 128 * 128;       // 128 rows x 64 bf16
    constexpr uint32_t B_BYTES = BN * 128;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t OUT_BYTES = kTma ? 128u * BN * 2u : 0u;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, still a __shared__ pointer
    uint8_t *tiles = smem;
    uint8_t *out_stage = smem + (size_t)stages * STAGE_BYTES;                      // [BN/64 slabs][128 rows][128 B], SWIZZLE_128B
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(out_stage + OUT_BYTES);
    uint64_t *empty_bar = full_bar + 8;
    uint64_t *tfull_bar = empty_bar + 8;           // [2] accumulator ready
    uint64_t *tempty_bar = tfull_bar + 2;          // [2] accumulator drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);            // +160 B
    int4 *tapc = reinterpret_cast<int4 *>(reinterpret_cast<uint8_t *>(full_bar) + 176);   // [4][16] A-box coordinates per (phase, tap)
    int4 *phc = tapc + 64;                                                         // [4] (B1, O0, O2, -) per phase
    int *bcs = reinterpret_cast<int *>(phc + 4);                                   // [4][16] B column offsets (MN-major B)
    float *col_acc = reinterpret_cast<float *>(bcs + 64);                           // [2][BN] per-CTA column sums for BN statistics
    float *bias_s = col_acc + 2 * BN;                                              // [BN] bias of the current n-tile
    float *bwd_c = bias_s + BN;                                                    // [3][BN] scale, shift, mean of the consumer BN (bwd_y mode)
    float *stage_f = bwd_c + 3 * BN;                                               // BN == 32 only: [4 warps][32][33] transposition buffer

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_phase = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_per_phase * p.num_phases;
    const bool prof = kProbe && p.dbg != nullptr && blockIdx.x == 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (kTma) prefetch_tmap(&tmO);
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        fence_barrier_init();
    }
    if (threadIdx.x < 64) {
        const int ph = threadIdx.x >> 4, tp = threadIdx.x & 15;
        tapc[threadIdx.x] = make_int4(p.A0[ph][tp], p.A1[ph][tp], p.A2[ph][tp], p.A3[ph][tp]);
        if (tp == 0) phc[ph] = make_int4(p.B1[ph], p.O0[ph], p.O2[ph], 0);
        bcs[threadIdx.x] = p.Bc[ph][tp];
    }
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) col_acc[i] = 0.f;
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();          // the next kernel of the stream may start its prologue
    pdl_wait();             // everything above overlapped the previous kernel's tail; its results are needed from here on

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            long long w_empty = 0, t_begin = prof ? clock64() : 0;
            const int chunks = p.chunks, num_taps = p.num_taps, bk_per_tap = p.bk_per_tap;
            const bool b_mn = p.b_mn != 0;
            const int flags = kProbe ? p.dbg_flags : 0;
            const uint32_t tiles_u32 = smem_u32(tiles), full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(p, t, tiles_per_phase);
                const int4 *tp = tapc + tc.phase * 16;
                const int brow = tc.nt * BN + phc[tc.phase].x;
                for (int tap = 0; tap < num_taps; ++tap) {
                    const int4 c4 = tp[tap];
                    const int ax = tc.x0 + c4.y, ay = tc.y0 + c4.w, bk = tap * bk_per_tap;
                    for (int c = 0; c < chunks; ++c) {
                        long long t0 = prof ? clock64() : 0;
                        mbar_wait_u32(empty_u32 + 8u * s, ph ^ 1u);
                        if (prof) w_empty += clock64() - t0;
                        const uint32_t a_dst = tiles_u32 + (uint32_t)s * STAGE_BYTES, b_dst = a_dst + A_BYTES, fb = full_u32 + 8u * s;
                        if (kProbe && (flags & 6)) {   // probe: drop one or both operand streams
                            const uint32_t bytes = ((flags & 2) ? 0u : A_BYTES) + ((flags & 4) ? 0u : B_BYTES);
                            if (bytes) mbar_expect_tx_u32(fb, bytes); else mbar_arrive(&full_bar[s]);
                            if (!(flags & 2)) tma_load_5d_u32(&tmA, fb, a_dst, c4.x + c * 64, ax, c4.z, ay, tc.n0);
                            if (!(flags & 4)) tma_load_2d_u32(&tmB, fb, b_dst, bk + c * 64, brow);
                        } else if (b_mn) {
                            mbar_expect_tx_u32(fb, STAGE_BYTES);
                            tma_load_5d_u32(&tmA, fb, a_dst, c4.x + c * 64, ax, c4.z, ay, tc.n0);
                            const int bcol = bcs[tc.phase * 16 + tap] + tc.nt * BN;
#pragma unroll
                            for (int ch = 0; ch < (BN >= 64 ? BN / 64 : 1); ++ch) tma_load_2d_u32(&tmB, fb, b_dst + ch * 8192, bcol + ch * 64, c * 64);
                        } else {
                            mbar_expect_tx_u32(fb, STAGE_BYTES);
                            tma_load_5d_u32(&tmA, fb, a_dst, c4.x + c * 64, ax, c4.z, ay, tc.n0);
                            tma_load_2d_u32(&tmB, fb, b_dst, bk + c * 64, brow);
                        }
                        if (++s == stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
            if (prof) { p.dbg[0] = (unsigned long long)w_empty; p.dbg[1] = (unsigned long long)(clock64() - t_begin); }
        }
    } else if (warp == 1) {
        {   // warp-uniform loop: every lane waits on the barriers, one elected lane issues the MMAs and commits
            const bool b_mn = p.b_mn != 0;
            const uint32_t idesc = make_idesc(128, BN < 16 ? 16 : BN, 0, b_mn ? 1 : 0);
            // descriptor words: lo = start address (16-B units) | LBO << 16 ; hi = SBO (1024 B) | version 1 | SWIZZLE_128B
            // MN-major B: 64-column chunks of [64 k][128 B], LBO = chunk stride (8 KB), 16 k-rows (one MMA) = 2048 B
            const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo0 = ((smem_u32(tiles) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lbo_fix = b_mn ? (((8192u >> 4) - 1u) << 16) : 0u;      // replaces LBO = 1 by LBO = 512
            const uint32_t b_kstep = b_mn ? 128u : 2u;
            const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar), tfull_u32 = smem_u32(tfull_bar), tempty_u32 = smem_u32(tempty_bar);
            const int num_kb = p.num_kb;
            const int flags = kProbe ? p.dbg_flags : 0;
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            long long w_full = 0, w_tempty = 0, t_begin = prof ? clock64() : 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                long long t0 = prof ? clock64() : 0;
                mbar_wait_u32(tempty_u32 + 8u * acc, acc_ph ^ 1u);      // epilogue has drained this accumulator
                if (prof) w_tempty += clock64() - t0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (prof) t0 = clock64();
                    mbar_wait_u32(full_u32 + 8u * s, ph);
                    if (prof) w_full += clock64() - t0;
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)s * (STAGE_BYTES >> 4), b_lo = a_lo + (A_BYTES >> 4) + b_lbo_fix;
                    if (elect_one()) {
                        if (!(kProbe && (flags & 1))) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {   // 4 x (K = 16) per 64-wide k-block; +32 B inside the swizzle atom
                                const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_lo + 2u * k), bd = ((uint64_t)desc_hi << 32) | (b_lo + b_kstep * k);
                                umma_f16(d_tmem, ad, bd, idesc, (kb | k) != 0);
                            }
                        }
                        if (kProbe && (flags & 24)) {       // probe: timing experiments only (results are garbage)
                            const int reps = (flags & 16) ? 2 : 1;
                            for (int r = 0; r < reps; ++r)
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_lo + 2u * k), bd = ((uint64_t)desc_hi << 32) | (b_lo + 2u * k);
                                    const uint32_t dd = (flags & 8) ? tmem_base + (uint32_t)(((acc + (k & 1)) & 1) * BN) : d_tmem;
                                    umma_f16(dd, ad, bd, idesc, 1);
                                }
                        }
                        umma_commit_u32(empty_u32 + 8u * s);                        // frees the smem slot when these MMAs retire
                        if (kb == num_kb - 1) umma_commit_u32(tfull_u32 + 8u * acc);   // accumulator complete
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                acc ^= 1; acc_ph ^= (uint32_t)(acc == 0);
            }
            if (prof && lane == 0) { p.dbg[2] = (unsigned long long)w_full; p.dbg[3] = (unsigned long long)w_tempty; p.dbg[4] = (unsigned long long)(clock64() - t_begin); }
        }
    } else {
        // ---------------- epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4) .. +31 (= tile rows)
        const int q = warp & 3;
        const int et = threadIdx.x - 64;                 // 0..127 within the epilogue group
        const int bw_mask = p.box_w - 1, bh_mask = p.box_h - 1, bwh_log2 = p.bw_log2 + p.bh_log2;
        const int row = q * 32 + lane;                   // this lane's TMEM lane = tile row
        const int rx = row & bw_mask, ry = (row >> p.bw_log2) & bh_mask, rn = row >> bwh_log2;
        const int act = p.act; const float act_param = p.act_param;
        const bool has_bias = p.bias != nullptr, has_stats = p.stats != nullptr;
        const bool bwd = kTma && has_stats && p.bwd_y != nullptr;
        int acc = 0; uint32_t acc_ph = 0;
        int stat_key = -1;                               // n-tile the per-CTA column sums belong to
        long long w_tfull = 0, t_begin = prof ? clock64() : 0;
        int ntile = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            ++ntile;
            const TileCoord tc = tile_coord(p, t, tiles_per_phase);
            const int nt = tc.nt, x0 = tc.x0, y0 = tc.y0, n0 = tc.n0;
            const bool row_ok = (x0 + rx) < p.out_w && (y0 + ry) < p.out_h && (n0 + rn) < p.out_n;
            if (has_stats && nt != stat_key) {
                if (stat_key >= 0) {                     // the column sums so far belong to another n-tile: flush them
                    epi_bar_sync();
                    for (int i = et; i < BN; i += 128) {
                        int c = stat_key * BN + i;
                        if (c < p.n_valid) { atomicAdd(p.stats + c, col_acc[i]); atomicAdd(p.stats + p.stats_stride + c, col_acc[BN + i]); }
                        col_acc[i] = 0.f; col_acc[BN + i] = 0.f;
                    }
                    epi_bar_sync();
                }
                stat_key = nt;
            }
            long long t0 = prof ? clock64() : 0;
            mbar_wait(&tfull_bar[acc], acc_ph);
            if (prof) w_tfull += clock64() - t0;
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            int ncols = p.n_valid - nt * BN; ncols = ncols > BN ? BN : ncols;
            if constexpr (kTma) {
                // ---- bf16 tile -> swizzled smem slabs -> one TMA store per 64-column slab
                if (has_bias) for (int i = et; i < BN; i += 128) { const int c = nt * BN + i; bias_s[i] = c < p.n_valid ? __ldg(p.bias + c) : 0.f; }
                if (bwd) for (int i = et; i < BN; i += 128) {
                    const int c = nt * BN + i; const bool ok = c < p.n_valid;
                    bwd_c[i] = ok ? __ldg(p.bwd_scale + c) : 0.f; bwd_c[BN + i] = ok ? __ldg(p.bwd_shift + c) : 0.f; bwd_c[2 * BN + i] = ok ? __ldg(p.bwd_mean + c) : 0.f;
                }
                if (et == 0) bulk_wait_read0();          // the previous tile's stores have finished reading out_stage
                epi_bar_sync();
                const int last_c0 = ((ncols + 31) >> 5) * 32 - 32;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[32];
                    if (c0 < ncols) {
                        uint32_t rr[32];
                        tmem_ld32(t_addr + (uint32_t)c0, rr);
                        tmem_ld_wait();
                        if (c0 == last_c0) {             // last read of this accumulator: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
                        if (has_bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c0 + j);
                                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                            }
                        }
                        switch (act) {                   // uniform: one activation loop is executed
                            case ACT_LEAKY:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * act_param;
                                break;
                            case ACT_RELU:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                                break;
                            case ACT_TANH:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = tanh_fast(v[j]);
                                break;
                            case ACT_SIGMOID:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
                                break;
                            default: break;
                        }
                        if (!row_ok) {                   // rows outside the image: zero (kept out of the statistics; the store clips them)
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0.f;
                        } else if (c0 + 32 > ncols) {    // pad columns stay zero
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (c0 + j >= ncols) v[j] = 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
                    uint8_t *slab_row = out_stage + (size_t)(c0 >> 6) * (128 * 128) + (size_t)row * 128;
                    const int cb = (c0 & 63) >> 3;       // first 16-byte chunk of this 32-column group within the 128-byte row
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                        pk.z = *reinterpret_cast<uint32_t *>(&h2); pk.w = *reinterpret_cast<uint32_t *>(&h3);
                        *reinterpret_cast<uint4 *>(slab_row + (((cb + i) ^ (row & 7)) << 4)) = pk;
                    }
                }
                if (bwd) {
                    __syncwarp();
                    const long long my_off = row_ok ? (long long)(n0 + rn) * p.sN + (long long)(y0 + ry) * p.sY + (long long)(x0 + rx) * p.sX + p.phase_off[tc.phase] + (long long)nt * BN : -1;
                    for (int sl = 0; sl * 64 < ncols; ++sl)
                        bwd_sums_slab(out_stage + (size_t)sl * (128 * 128) + (size_t)(q * 32) * 128, p.bwd_y, my_off, sl * 64 + 2 * lane, bwd_c, BN, p.bwd_act, p.bwd_negval, lane, col_acc);
                } else if (has_stats) {
                    // column sums of the STORED (bf16-rounded) values: lane l owns columns 2l, 2l+1 of each slab, over this warp's 32 rows
                    __syncwarp();
                    for (int sl = 0; sl * 64 < ncols; ++sl) {
                        const uint8_t *slab = out_stage + (size_t)sl * (128 * 128) + (size_t)(q * 32) * 128;
                        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const uint32_t wv = *reinterpret_cast<const uint32_t *>(slab + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + ((lane & 3) << 2));
                            const float a = __uint_as_float(wv << 16), b = __uint_as_float(wv & 0xFFFF0000u);
                            s1a += a; s2a = fmaf(a, a, s2a); s1b += b; s2b = fmaf(b, b, s2b);
                        }
                        atomicAdd(&col_acc[sl * 64 + 2 * lane], s1a); atomicAdd(&col_acc[sl * 64 + 2 * lane + 1], s1b);
                        atomicAdd(&col_acc[BN + sl * 64 + 2 * lane], s2a); atomicAdd(&col_acc[BN + sl * 64 + 2 * lane + 1], s2b);
                    }
                }
                fence_proxy_async();                     // generic-proxy smem writes -> visible to the TMA store
                epi_bar_sync();
                if (et == 0 && p.out_bf16) {
                    const int4 pc = phc[tc.phase];
                    for (int sl = 0; sl < BN / 64; ++sl) {
                        const int col = nt * BN + sl * 64;
                        if (col < p.o_cols) tma_store_5d(&tmO, out_stage + (size_t)sl * (128 * 128), pc.y + col, x0, pc.z, y0, n0);
                    }
                    bulk_commit();
                }
            } else {
                // ---- BN == 32: direct stores (thin outputs: 3 / 4 / 12 / 16 channels)
                const long long tile_base = p.phase_off[tc.phase] + (long long)nt * BN;
                float *my_stage = stage_f + (size_t)(warp - 2) * 32 * 33;
                uint32_t rr[32];
                if (ncols <= 8) { uint32_t r8[8]; tmem_ld8(t_addr, r8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) rr[j] = r8[j];
                } else tmem_ld32(t_addr, rr);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (has_stats) {                         // rare (no thin layer is followed by BN): transpose through smem
#pragma unroll
                    for (int j = 0; j < 32; ++j) my_stage[lane * 33 + j] = row_ok ? __uint_as_float(rr[j]) : 0.f;
                    __syncwarp();
                    const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
                    const float nrows_ok = (float)__popc(okmask);
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                    for (int i = 0; i < 32; ++i) { float tt = my_stage[i * 33 + lane]; s1 += tt; s2 += tt * tt; }
                    const int col_l = nt * BN + lane;
                    if (has_bias && col_l < p.n_valid) { const float bb = __ldg(p.bias + col_l); s2 += 2.f * bb * s1 + nrows_ok * bb * bb; s1 += nrows_ok * bb; }
                    atomicAdd(&col_acc[lane], s1); atomicAdd(&col_acc[BN + lane], s2);
                    __syncwarp();
                }
                if (row_ok && p.out_bf16) {
                    __nv_bfloat16 *o = p.out_bf16 + (long long)(n0 + rn) * p.sN + (long long)(y0 + ry) * p.sY + (long long)(x0 + rx) * p.sX + tile_base;
                    const bool al16 = (reinterpret_cast<uintptr_t>(o) & 15) == 0, al8 = (reinterpret_cast<uintptr_t>(o) & 7) == 0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {        // 8 columns per group
                        if (g * 8 >= ncols) break;
                        float w8[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float f = __uint_as_float(rr[g * 8 + k]);
                            const int c = g * 8 + k;
                            if (c < ncols) { if (has_bias) f += __ldg(p.bias + nt * BN + c); f = apply_act(f, act, act_param); } else f = 0.f;
                            w8[k] = f;
                        }
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(w8[0], w8[1]), h1 = __floats2bfloat162_rn(w8[2], w8[3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(w8[4], w8[5]), h3 = __floats2bfloat162_rn(w8[6], w8[7]);
                        const int cv = ncols - g * 8;
                        if (cv >= 8 && al16) {
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                            pk.z = *reinterpret_cast<uint32_t *>(&h2); pk.w = *reinterpret_cast<uint32_t *>(&h3);
                            *reinterpret_cast<uint4 *>(o + g * 8) = pk;
                        } else if (cv >= 4 && al8) {
                            uint2 pk; pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                            *reinterpret_cast<uint2 *>(o + g * 8) = pk;
                            __nv_bfloat16 hh[4] = {__float2bfloat16(w8[4]), __float2bfloat16(w8[5]), __float2bfloat16(w8[6]), __float2bfloat16(w8[7])};
#pragma unroll
                            for (int k = 4; k < 8; ++k) if (k < cv) o[g * 8 + k] = hh[k - 4];
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) if (k < cv) o[g * 8 + k] = __float2bfloat16(w8[k]);
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        if (has_stats && stat_key >= 0) {
            epi_bar_sync();
            for (int i = et; i < BN; i += 128) {
                int c = stat_key * BN + i;
                if (c < p.n_valid) { atomicAdd(p.stats + c, col_acc[i]); atomicAdd(p.stats + p.stats_stride + c, col_acc[BN + i]); }
            }
        }
        if (kTma && et == 0) bulk_wait_all();            // smem must outlive the last TMA store
        if (prof && et == 0) { p.dbg[5] = (unsigned long long)w_tfull; p.dbg[6] = (unsigned long long)(clock64() - t_begin); p.dbg[7] = (unsigned long long)ntile; }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------ patch kernel, dgrad type (4 sub-pixel phases)
// L[n, 2y+py, 2x+px, cl] = sum_{a,b,cs} S[n, y+dy(py,a), x+dx(px,b), cs] * Wt[ph][cl][ab][cs]        (conv dgrad, full-conv fprop)
// All 16 (phase, tap) MMAs of a 128-pixel tile read SHIFTED WINDOWS of one shared-memory patch of S (halo of one pixel),
// loaded once per 64-channel chunk: the 16 per-tap operand loads of the gather kernel (the L2 -> SM traffic that bounded
// it) collapse into one.  A UMMA descriptor may start any whole number of 128-byte rows into a TMA-written SWIZZLE_128B
// region and step between 8-row groups by any multiple of 16 bytes (profiles/r1_desc_probe.log), so a window is just a
// start address: the tile's rows are ordered (y, n, x) with 8 x-positions per group and the patch is stored [y][n][x]
// with a pitch of 10 pixels, which makes the group stride uniform (1280 B) even when a tile spans several samples.
// Warp roles as in the gather kernel.  Pipelines: patch (2 buffers) and weight stages (one stage = the 4 taps of one
// phase, 16 MMAs per barrier round trip) -> 4 accumulators per tile (one per phase) -> epilogue per phase.
struct PatchDgradParams {
    int chunks;                  // Csp / 64
    int m_tiles, n_tiles;        // tiles enumerated m fastest
    int tiles_x, tiles_y;        // m-tile -> (tx, ty, tn); box = 8 (x) x bn (samples) x bh (y)
    int bh, bn, bn_log2;
    int npatch;                  // patch buffers in flight (2 .. 8): a patch is one HBM round trip, thin layers are bound by how many are in flight
    int b_resident;              // 1: the weight taps of all phases and chunks fit the B stages and do not depend on the tile (one N tile):
                                 //    loaded once per CTA, stage index = chunk * 4 + phase, never released
    int patch_bytes;             // TMA box bytes: 128 * 10 * bn * (bh + 2)
    int patch_stride;            // patch_bytes rounded up to 1024
    int a_off[4][4];             // byte offset of window (phase, tap ab) inside the patch
    int Csp, cl_rows;            // B coordinates (K-major Wt): column ab*Csp + 64*c, row ph*cl_rows + nt*BN
    int b_mn;                    // 1: B taps are read MN-major from Wf [Cs rows][16*Clp cols]: column b_col[ph][ab] + nt*BN + 64*chunk, row 64*c
    int b_col[4][4];
    int out_w, out_h, out_n;     // extent of the S pixel grid
    int n_valid, Clp;            // valid / padded output channels of one phase
    int H2, W2;                  // output image size (direct-store path)
    __nv_bfloat16 *out;          // output tensor (direct-store path, BN == 16)
    const float *bias;
    float *stats; int stats_stride;
    int act; float act_param;
    // [opt] BN-backward sums of the CONSUMER of this output: the output is dLoss/d(activation output) of a BN + activation block whose conv
    // output is bwd_y (same geometry as the output).  With dz = out * act'(y * scale + shift):  stats[c] += sum dz,
    // stats[stats_stride + c] += sum dz * (y - mean)  -- the first pass of the BN backward, taken from the tile while it is in shared memory
    const __nv_bfloat16 *bwd_y;
    const float *bwd_scale, *bwd_shift, *bwd_mean;
    int bwd_act;
    float bwd_negval;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, BN <= 16 ? 2 : 1)      // thin outputs: two co-resident CTAs per SM (the per-tile issue loops, not bytes, bound that variant)
patch_dgrad_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ PatchDgradParams p, int stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr bool kTma = BN >= 64;
    constexpr int NSETS = (8 * BN <= 512) ? 2 : 1;            // accumulator sets (4 phases x BN columns each)
    constexpr uint32_t TMEM_COLS = NSETS * 4 * BN < 32 ? 32 : NSETS * 4 * BN;
    constexpr uint32_t BTAP_BYTES = BN * 128;                 // one tap of one phase: BN rows x 64 cs
    constexpr uint32_t BSTAGE_BYTES = 4 * BTAP_BYTES;
    constexpr uint32_t OUT_BYTES = kTma ? 128u * BN * 2u : 0u; // one phase tile
    constexpr int NOUT = BN >= 128 ? 1 : 2;                    // staging buffers (BN = 128: shared memory allows one)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int npatch = p.npatch;
    uint8_t *patch = smem;                                                    // [npatch][patch_stride]
    uint8_t *bst = patch + (size_t)npatch * (size_t)p.patch_stride;           // [stages][BSTAGE_BYTES]
    uint8_t *out_stage = bst + (size_t)stages * BSTAGE_BYTES;                 // [NOUT][OUT_BYTES]
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_stage + NOUT * OUT_BYTES);
    uint64_t *pfull = bars, *pempty = bars + 8, *bfull = bars + 16, *bempty = bars + 24, *tfull = bars + 32, *tempty = bars + 34;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 36);            // +288 B
    float *col_acc = reinterpret_cast<float *>(bars + 38);                    // [2][BN]
    float *bias_s = col_acc + 2 * BN;                                         // [BN]
    float *bwd_c = bias_s + BN;                                               // [3][BN] scale, shift, mean of the consumer BN (bwd_y mode)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int txy = p.tiles_x * p.tiles_y;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmS); prefetch_tmap(&tmB);
        if (kTma) prefetch_tmap(&tmO);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        for (int i = 0; i < npatch; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], 1); }
        for (int i = 0; i < stages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) col_acc[i] = 0.f;
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();          // the next kernel of the stream may start its prologue
    pdl_wait();             // everything above overlapped the previous kernel's tail; its results are needed from here on

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t patch_u32 = smem_u32(patch), bst_u32 = smem_u32(bst);
            const uint32_t pfull_u32 = smem_u32(pfull), pempty_u32 = smem_u32(pempty), bfull_u32 = smem_u32(bfull), bempty_u32 = smem_u32(bempty);
            const int chunks = p.chunks, Csp = p.Csp, cl_rows = p.cl_rows;
            const bool b_mn = p.b_mn != 0;
            const bool b_res = p.b_resident != 0;
            int pb = 0; uint32_t pph = 0; int s = 0; uint32_t ph = 0;
            bool first = true;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
                const int tn = mt / txy, rem = mt - tn * txy, ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                const int x0 = tx * 8, y0 = ty * p.bh, n0 = tn * p.bn;
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait_u32(pempty_u32 + 8u * pb, pph ^ 1u);
                    mbar_expect_tx_u32(pfull_u32 + 8u * pb, (uint32_t)p.patch_bytes);
                    tma_load_4d_u32(&tmS, pfull_u32 + 8u * pb, patch_u32 + (uint32_t)pb * (uint32_t)p.patch_stride, c * 64, x0 - 1, n0, y0 - 1);
                    if (++pb == npatch) { pb = 0; pph ^= 1u; }
                    if (b_res && !first) continue;                      // resident weights: loaded with the first tile only
#pragma unroll
                    for (int phs = 0; phs < 4; ++phs) {
                        if (b_res) s = c * 4 + phs; else mbar_wait_u32(bempty_u32 + 8u * s, ph ^ 1u);
                        const uint32_t fb = bfull_u32 + 8u * s, dst = bst_u32 + (uint32_t)s * BSTAGE_BYTES;
                        mbar_expect_tx_u32(fb, BSTAGE_BYTES);
                        const int brow = phs * cl_rows + nt * BN;
                        if (b_mn) {
#pragma unroll
                            for (int ab = 0; ab < 4; ++ab)
#pragma unroll
                                for (int ch = 0; ch < (BN >= 64 ? BN / 64 : 1); ++ch)
                                    tma_load_2d_u32(&tmB, fb, dst + ab * BTAP_BYTES + ch * 8192, p.b_col[phs][ab] + nt * BN + ch * 64, c * 64);
                        } else {
#pragma unroll
                            for (int ab = 0; ab < 4; ++ab) tma_load_2d_u32(&tmB, fb, dst + ab * BTAP_BYTES, ab * Csp + c * 64, brow);
                        }
                        if (!b_res && ++s == stages) { s = 0; ph ^= 1u; }
                    }
                }
                first = false;
            }
        }
    } else if (warp == 1) {
        const bool b_mn = p.b_mn != 0;
        const uint32_t idesc = make_idesc(128, BN, 0, b_mn ? 1 : 0);
        const uint32_t b_lbo_fix = b_mn ? (((8192u >> 4) - 1u) << 16) : 0u, b_kstep = b_mn ? 128u : 2u;
        const uint32_t hiA = (1280u >> 4) | (1u << 14) | (2u << 29);          // group stride = one patch row of 10 pixels
        const uint32_t hiB = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t patch_lo = ((smem_u32(patch) & 0x3FFFFu) >> 4) | (1u << 16), bst_lo = ((smem_u32(bst) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t pfull_u32 = smem_u32(pfull), pempty_u32 = smem_u32(pempty), bfull_u32 = smem_u32(bfull), bempty_u32 = smem_u32(bempty);
        const uint32_t tfull_u32 = smem_u32(tfull), tempty_u32 = smem_u32(tempty);
        uint32_t aoff[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) aoff[i][j] = (uint32_t)p.a_off[i][j] >> 4;
        const int chunks = p.chunks;
        const bool b_res = p.b_resident != 0;
        int pb = 0; uint32_t pph = 0; int s = 0; uint32_t ph = 0; int set = 0; uint32_t set_ph = 0;
        bool first = true;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            mbar_wait_u32(tempty_u32 + 8u * set, set_ph ^ 1u);
            tc_fence_after();
            for (int c = 0; c < chunks; ++c) {
                mbar_wait_u32(pfull_u32 + 8u * pb, pph);
                const uint32_t a_base = patch_lo + (((uint32_t)pb * (uint32_t)p.patch_stride) >> 4);
#pragma unroll
                for (int phs = 0; phs < 4; ++phs) {
                    if (b_res) { s = c * 4 + phs; if (first) mbar_wait_u32(bfull_u32 + 8u * s, 0u); }
                    else mbar_wait_u32(bfull_u32 + 8u * s, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((set * 4 + phs) * BN);
                        const uint32_t b_base = bst_lo + (uint32_t)s * (BSTAGE_BYTES >> 4) + b_lbo_fix;
#pragma unroll
                        for (int ab = 0; ab < 4; ++ab)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t ad = ((uint64_t)hiA << 32) | (a_base + aoff[phs][ab] + 2u * k);
                                const uint64_t bd = ((uint64_t)hiB << 32) | (b_base + (uint32_t)ab * (BTAP_BYTES >> 4) + b_kstep * k);
                                umma_f16(d_tmem, ad, bd, idesc, (c | ab | k) != 0);
                            }
                        if (!b_res) umma_commit_u32(bempty_u32 + 8u * s);
                        if (phs == 3) {
                            umma_commit_u32(pempty_u32 + 8u * pb);
                            if (c == chunks - 1) umma_commit_u32(tfull_u32 + 8u * set);
                        }
                    }
                    __syncwarp();
                    if (!b_res && ++s == stages) { s = 0; ph ^= 1u; }
                }
                if (++pb == npatch) { pb = 0; pph ^= 1u; }
            }
            first = false;
            if (NSETS == 2) { set ^= 1; set_ph ^= (uint32_t)(set == 0); } else set_ph ^= 1u;
        }
    } else {
        const int q = warp & 3, et = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const int rx = row & 7, rn = (row >> 3) & (p.bn - 1), ry = row >> (3 + p.bn_log2);
        const int act = p.act; const float act_param = p.act_param;
        const bool has_bias = p.bias != nullptr, has_stats = p.stats != nullptr;
        const bool bwd = kTma && has_stats && p.bwd_y != nullptr;
        int set = 0; uint32_t set_ph = 0;
        int stat_key = -1;
        int ob = 0;                                       // staging buffer of the next phase tile
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
            const int tn = mt / txy, rem = mt - tn * txy, ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
            const int x0 = tx * 8, y0 = ty * p.bh, n0 = tn * p.bn;
            const bool row_ok = (x0 + rx) < p.out_w && (y0 + ry) < p.out_h && (n0 + rn) < p.out_n;
            if (has_stats && nt != stat_key) {
                if (stat_key >= 0) {
                    epi_bar_sync();
                    for (int i = et; i < BN; i += 128) {
                        int c = stat_key * BN + i;
                        if (c < p.n_valid) { atomicAdd(p.stats + c, col_acc[i]); atomicAdd(p.stats + p.stats_stride + c, col_acc[BN + i]); }
                        col_acc[i] = 0.f; col_acc[BN + i] = 0.f;
                    }
                    epi_bar_sync();
                }
                stat_key = nt;
            }
            mbar_wait(&tfull[set], set_ph);
            tc_fence_after();
            int ncols = p.n_valid - nt * BN; ncols = ncols > BN ? BN : ncols;
            if constexpr (kTma) {
                if (has_bias) for (int i = et; i < BN; i += 128) { const int c = nt * BN + i; bias_s[i] = c < p.n_valid ? __ldg(p.bias + c) : 0.f; }
                if (bwd) for (int i = et; i < BN; i += 128) {
                    const int c = nt * BN + i; const bool ok = c < p.n_valid;
                    bwd_c[i] = ok ? __ldg(p.bwd_scale + c) : 0.f; bwd_c[BN + i] = ok ? __ldg(p.bwd_shift + c) : 0.f; bwd_c[2 * BN + i] = ok ? __ldg(p.bwd_mean + c) : 0.f;
                }
                const int last_c0 = ((ncols + 31) >> 5) * 32 - 32;
#pragma unroll 1
                for (int phs = 0; phs < 4; ++phs) {
                    uint8_t *stage = out_stage + (size_t)ob * OUT_BYTES;
                    if (et == 0) { if (NOUT == 2) bulk_wait_read1(); else bulk_wait_read0(); }   // earlier stores have finished reading this buffer
                    epi_bar_sync();
                    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * 4 + phs) * BN);
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        float v[32];
                        if (c0 < ncols) {
                            uint32_t rr[32];
                            tmem_ld32(t_addr + (uint32_t)c0, rr);
                            tmem_ld_wait();
                            if (phs == 3 && c0 == last_c0) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&tempty[set]);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
                            if (has_bias) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c0 + j);
                                    v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                                }
                            }
                            if (act != ACT_NONE) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], act, act_param);
                            }
                            if (!row_ok) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = 0.f;
                            } else if (c0 + 32 > ncols) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (c0 + j >= ncols) v[j] = 0.f;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0.f;
                        }
                        uint8_t *slab_row = stage + (size_t)(c0 >> 6) * (128 * 128) + (size_t)row * 128;
                        const int cb = (c0 & 63) >> 3;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                            pk.z = *reinterpret_cast<uint32_t *>(&h2); pk.w = *reinterpret_cast<uint32_t *>(&h3);
                            *reinterpret_cast<uint4 *>(slab_row + (((cb + i) ^ (row & 7)) << 4)) = pk;
                        }
                    }
                    if (bwd) {
                        __syncwarp();
                        const long long my_off = row_ok ? (((long long)(n0 + rn) * p.H2 + 2 * (y0 + ry) + (phs >> 1)) * p.W2 + 2 * (x0 + rx) + (phs & 1)) * p.Clp + (long long)nt * BN : -1;
                        for (int sl = 0; sl * 64 < ncols; ++sl)
                            bwd_sums_slab(stage + (size_t)sl * (128 * 128) + (size_t)(q * 32) * 128, p.bwd_y, my_off, sl * 64 + 2 * lane, bwd_c, BN, p.bwd_act, p.bwd_negval, lane, col_acc);
                    } else if (has_stats) {
                        __syncwarp();
                        for (int sl = 0; sl * 64 < ncols; ++sl) {
                            const uint8_t *slab = stage + (size_t)sl * (128 * 128) + (size_t)(q * 32) * 128;
                            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                const uint32_t wv = *reinterpret_cast<const uint32_t *>(slab + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + ((lane & 3) << 2));
                                const float a = __uint_as_float(wv << 16), b = __uint_as_float(wv & 0xFFFF0000u);
                                s1a += a; s2a = fmaf(a, a, s2a); s1b += b; s2b = fmaf(b, b, s2b);
                            }
                            atomicAdd(&col_acc[sl * 64 + 2 * lane], s1a); atomicAdd(&col_acc[sl * 64 + 2 * lane + 1], s1b);
                            atomicAdd(&col_acc[BN + sl * 64 + 2 * lane], s2a); atomicAdd(&col_acc[BN + sl * 64 + 2 * lane + 1], s2b);
                        }
                    }
                    fence_proxy_async();
                    epi_bar_sync();
                    if (et == 0) {
                        for (int sl = 0; sl < BN / 64; ++sl) {
                            const int col = nt * BN + sl * 64;
                            if (col < p.Clp) tma_store_5d(&tmO, stage + (size_t)sl * (128 * 128), (phs & 1) * p.Clp + col, x0, n0, y0, phs >> 1);
                        }
                        bulk_commit();
                    }
                    if (NOUT == 2) ob ^= 1;
                }
            } else {
                // ---- thin outputs (Clp = 4 or 16): every lane owns one S pixel = a 2x2 block of output pixels
                uint32_t r4[4][16];
#pragma unroll
                for (int phs = 0; phs < 4; ++phs) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * 4 + phs) * BN), r4[phs]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[set]);
                if (row_ok) {
                    const int Clp = p.Clp;
                    float bv[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) bv[k] = (has_bias && k < ncols) ? __ldg(p.bias + k) : 0.f;
                    // the two sub-pixel columns px = 0, 1 of one output row are adjacent in memory: one 16-byte (Clp = 4) or four 16-byte
                    // (Clp = 16) stores per output row instead of two 8-byte / 32-byte halves
#pragma unroll
                    for (int py = 0; py < 2; ++py) {
                        uint32_t pk[2][8];
#pragma unroll
                        for (int px = 0; px < 2; ++px)
#pragma unroll
                            for (int k = 0; k < 16; k += 2) {
                                const int phs = py * 2 + px;
                                float f0 = k < ncols ? apply_act(__uint_as_float(r4[phs][k]) + bv[k], act, act_param) : 0.f;
                                float f1 = k + 1 < ncols ? apply_act(__uint_as_float(r4[phs][k + 1]) + bv[k + 1], act, act_param) : 0.f;
                                __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
                                pk[px][k >> 1] = *reinterpret_cast<uint32_t *>(&h);
                            }
                        __nv_bfloat16 *o = p.out + (((long long)(n0 + rn) * p.H2 + 2 * (y0 + ry) + py) * p.W2 + 2 * (x0 + rx)) * Clp;
                        if (Clp == 4) *reinterpret_cast<uint4 *>(o) = make_uint4(pk[0][0], pk[0][1], pk[1][0], pk[1][1]);
                        else {
                            reinterpret_cast<uint4 *>(o)[0] = make_uint4(pk[0][0], pk[0][1], pk[0][2], pk[0][3]); reinterpret_cast<uint4 *>(o)[1] = make_uint4(pk[0][4], pk[0][5], pk[0][6], pk[0][7]);
                            reinterpret_cast<uint4 *>(o)[2] = make_uint4(pk[1][0], pk[1][1], pk[1][2], pk[1][3]); reinterpret_cast<uint4 *>(o)[3] = make_uint4(pk[1][4], pk[1][5], pk[1][6], pk[1][7]);
                        }
                    }
                }
            }
            if (NSETS == 2) { set ^= 1; set_ph ^= (uint32_t)(set == 0); } else set_ph ^= 1u;
        }
        if (has_stats && stat_key >= 0) {
            epi_bar_sync();
            for (int i = et; i < BN; i += 128) {
                int c = stat_key * BN + i;
                if (c < p.n_valid) { atomicAdd(p.stats + c, col_acc[i]); atomicAdd(p.stats + p.stats_stride + c, col_acc[BN + i]); }
            }
        }
        if (kTma && et == 0) bulk_wait_all();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------ MN-major wgrad GEMM
// M tile = 128 rows = 2 row-blocks of 64 (tap, cl-chunk); N tile = BN columns of cs; K = pixels.
template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmS, const WgradParams p, int stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t BOX_BYTES = 64 * 128;            // 64 pixels x 64 channels bf16
    constexpr uint32_t A_BYTES = 2 * BOX_BYTES;
    constexpr uint32_t B_BYTES = (BN / 64) * BOX_BYTES;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, still a __shared__ pointer
    uint8_t *tiles = smem;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)stages * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + stages;
    uint64_t *acc_bar = empty_bar + stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, nt = blockIdx.y, split = blockIdx.z;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.num_kb_total, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmL);
        prefetch_tmap(&tmS);
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();          // the next kernel of the stream may start its prologue
    pdl_wait();             // everything above overlapped the previous kernel's tail; its results are needed from here on
    if (nkb <= 0) { __syncthreads(); if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS); return; }

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tiles_u32 = smem_u32(tiles), full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
            // the two (tap, chunk) row-blocks of this M tile: gather coordinates are fixed for the whole K loop
            int ac0[2], adx[2], a2[2], ady[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rb = mt * 2 + h;
                int tap = rb / p.chunks_per_tap; const int chunk = rb - tap * p.chunks_per_tap;
                tap = min(tap, 15);
                ac0[h] = chunk * 64 + p.gx0[tap]; adx[h] = p.gdx[tap]; a2[h] = p.g2[tap]; ady[h] = p.gdy[tap];
            }
            int tx = kb_begin % p.tiles_x, ty = (kb_begin / p.tiles_x) % p.tiles_y, tn = kb_begin / (p.tiles_x * p.tiles_y);
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait_u32(empty_u32 + 8u * s, ph ^ 1u);
                const int x0 = tx * p.box_w, y0 = ty * p.box_h, n0 = tn * p.box_n;
                const uint32_t a_dst = tiles_u32 + (uint32_t)s * STAGE_BYTES, b_dst = a_dst + A_BYTES, fb = full_u32 + 8u * s;
                mbar_expect_tx_u32(fb, STAGE_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h) tma_load_5d_u32(&tmL, fb, a_dst + h * BOX_BYTES, ac0[h], x0 + adx[h], a2[h], y0 + ady[h], n0);
#pragma unroll
                for (int c = 0; c < BN / 64; ++c) tma_load_5d_u32(&tmS, fb, b_dst + c * BOX_BYTES, (nt * (BN / 64) + c) * 64, x0, 0, y0, n0);
                if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++tn; } }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc(128, BN, 1, 1);
        // MN-major descriptors: lo = start address | LBO (distance between the two 64-element M/N chunks) << 16 ; hi = SBO 1024 | v1 | SW128
        const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lo0 = ((smem_u32(tiles) & 0x3FFFFu) >> 4) | ((BOX_BYTES >> 4) << 16);
        const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait_u32(full_u32 + 8u * s, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = a_lo0 + (uint32_t)s * (STAGE_BYTES >> 4), b_lo = a_lo + (A_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 16 pixels (K) per MMA = 2 groups of 8 K-rows = 2048 B
                    const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_lo + 128u * k), bd = ((uint64_t)desc_hi << 32) | (b_lo + 128u * k);
                    umma_f16(tmem_base, ad, bd, idesc, (i | k) != 0);
                }
                umma_commit_u32(empty_u32 + 8u * s);
                if (i == nkb - 1) umma_commit(acc_bar);
            }
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;            // row within the M tile
        const int rb = mt * 2 + (row >> 6);
        const int tap = rb / p.chunks_per_tap, chunk = rb - tap * p.chunks_per_tap;
        const int cl = chunk * 64 + (row & 63);
        const bool row_ok = tap < p.num_taps && cl < p.cl_valid;
        float *obase = p.out + (long long)tap * p.cl_stride + cl;
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        // column chunks are visited in an order rotated by the CTA index: neighbouring CTAs (same cs rows, adjacent 512-byte
        // row segments, 32 KB apart per cs) would otherwise walk the same address pattern in lock step
#pragma unroll 1
        for (int ci = 0; ci < BN; ci += 32) {
            const int c0 = (ci + 32 * (int)(blockIdx.x + blockIdx.z)) % BN;
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    int cs = nt * BN + c0 + j;
                    if (cs < p.cs_valid) {
                        float f = __uint_as_float(r[j]) * p.scale;
                        float *o = obase + (long long)cs * p.out_cs_stride;
                        if (p.accumulate) atomicAdd(o, f); else *o = f;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

}  // namespace tc
