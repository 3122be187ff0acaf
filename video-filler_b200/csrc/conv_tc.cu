// conv_tc.cu -- host side of the BF16 tcgen05 implicit-GEMM path: TMA tensor maps over NHWC bf16
// activations, weight repacking, launch configuration, and NCHW-fp32 wrappers for the op-level ABI.
#include "conv_tc.h"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "tc_gemm.cuh"

namespace {

// ------------------------------------------------------------------ tensor maps
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encoder() {
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cenn_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
        return 1;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return 0;
}

int make_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) {
    if (get_encoder()) return 1;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address not 16-byte aligned");
    for (int i = 0; i < rank - 1; ++i) REQUIRE((gs[i] & 15) == 0, "TMA stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
    for (int i = 0; i < rank; ++i) REQUIRE(bx[i] >= 1 && bx[i] <= 256 && gd[i] >= 1, "TMA box/dim %d out of range (box %u, dim %llu)", i, bx[i], (unsigned long long)gd[i]);
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// 5-D gather view of L [N, H2, W2, C]: (px*C + c, x', py, y', n) with x = 2x'+px, y = 2y'+py
int map_gather(CUtensorMap *m, const bf16 *L, int N, int H2, int W2, int C, int bw, int bh, int bn) {
    uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)W2 / 2, 2, (uint64_t)H2 / 2, (uint64_t)N};
    uint64_t st[4] = {(uint64_t)2 * C * 2, (uint64_t)W2 * C * 2, (uint64_t)2 * W2 * C * 2, (uint64_t)H2 * W2 * C * 2};
    uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bn};
    return make_map(m, L, 5, dims, st, box);
}
// 5-D plain view of S [N, h, w, C]: (c, x, 0, y, n)
int map_plain(CUtensorMap *m, const bf16 *S, int N, int h, int w, int C, int bw, int bh, int bn) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)w, 1, (uint64_t)h, (uint64_t)N};
    uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
    uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bn};
    return make_map(m, S, 5, dims, st, box);
}
int map_2d(CUtensorMap *m, const bf16 *B, uint64_t K, uint64_t rows, int box_rows) {
    uint64_t dims[2] = {K, rows};
    uint64_t st[1] = {K * 2};
    uint32_t box[2] = {64, (uint32_t)box_rows};
    return make_map(m, B, 2, dims, st, box);
}

int map_bmn(CUtensorMap *m, const bf16 *B, uint64_t Ncols, uint64_t Krows) {
    uint64_t dims[2] = {Ncols, Krows};
    uint64_t st[1] = {Ncols * 2};
    uint32_t box[2] = {64, 64};
    return make_map(m, B, 2, dims, st, box);
}
}  // namespace
// helpers shared with probes.cu
int tc_make_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) { return make_map(m, base, rank, dims, strides_bytes, box); }
int tc_map_2d(CUtensorMap *m, const bf16 *B, uint64_t K, uint64_t rows, int box_rows) { return map_2d(m, B, K, rows, box_rows); }
namespace {
const int UPH[2][2] = {{1, 3}, {0, 2}};   // dgrad sub-pixel phase p, tap index a -> window row/column u
int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; }
int pow2_le(int v, int cap) { int p = 1; while (p * 2 <= v && p * 2 <= cap) p *= 2; return p; }
// split `total` pixels per tile over (w, h, n)
void choose_box(int w, int h, int total, int &bw, int &bh, int &bn) {
    bw = pow2_le(w, total);
    bh = pow2_le(h, total / bw);
    bn = total / (bw * bh);
}

const int SMEM_LIMIT = 227 * 1024;

template <int BN>
int set_attr_gather() {
    static bool done = false;
    if (!done) { CK(cudaFuncSetAttribute(tc::gather_gemm_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CK(cudaFuncSetAttribute(tc::gather_gemm_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)); done = true; }
    return 0;
}
template <int BN>
int set_attr_wgrad() {
    static bool done = false;
    if (!done) { CK(cudaFuncSetAttribute(tc::wgrad_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)); done = true; }
    return 0;
}
int config_gather(cenn_state *s, TcPlan *pl, int BN, int num_kb, int total_tiles) {
    const int stage_bytes = 128 * 128 + BN * 128;
    const int out_bytes = BN >= 64 ? 128 * BN * 2 : 0;
    const int fixed = 1024 /*align*/ + out_bytes + 176 /*barriers, tmem slot*/ + 68 * 16 + 256 /*tap tables*/ + 6 * BN * 4 + (BN == 32 ? 4 * 32 * 33 * 4 : 0) + 64;
    // Two co-resident CTAs per SM when the tile is narrow (BN <= 128: 2 x 2*BN TMEM columns fit): the two single-thread
    // loops of a CTA (TMA issue, MMA issue) are instruction-latency bound, a second CTA doubles the issue capacity.
    int per_sm = (BN <= 128 && total_tiles > s->sm_count) ? 2 : 1;
    if (getenv("CENN_CTAS_PER_SM")) per_sm = atoi(getenv("CENN_CTAS_PER_SM")) == 2 && BN <= 128 ? 2 : 1;
    const int budget = per_sm == 2 ? (SMEM_LIMIT + 1024) / 2 - 1024 : SMEM_LIMIT;     // 227 KB usable + 1 KB reserved per CTA
    int stages = (budget - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) { per_sm = 1; stages = (SMEM_LIMIT - fixed) / stage_bytes; if (stages > 8) stages = 8; if (stages < 2) stages = 2; }
    pl->kind = 1; pl->BN = BN; pl->stages = stages;
    pl->smem = (size_t)stages * stage_bytes + fixed;
    const int ctas = s->sm_count * per_sm;
    pl->grid[0] = (unsigned)(total_tiles < ctas ? total_tiles : ctas);   // persistent
    pl->grid[1] = 1; pl->grid[2] = 1;
    (void)num_kb;
    switch (BN) {
        case 32: return set_attr_gather<32>();
        case 64: return set_attr_gather<64>();
        case 128: return set_attr_gather<128>();
        case 256: return set_attr_gather<256>();
    }
    cenn_set_error("unsupported BN %d", BN);
    return 1;
}
int config_wgrad(TcPlan *pl, int BN, int nkb, dim3 grid) {
    const int stage_bytes = 2 * 8192 + (BN / 64) * 8192;
    const int fixed = 1024 + 8 * (2 * 8 + 1) + 16 + 256;
    // two co-resident CTAs per SM (TMEM: BN <= 256 columns each): one's epilogue overlaps the other's main loop and the
    // single-thread issue loops get twice the capacity
    const int budget = getenv("CENN_WGRAD_ONE_CTA") ? SMEM_LIMIT : (SMEM_LIMIT + 1024) / 2 - 1024;
    int stages = (budget - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages > nkb) stages = nkb;
    if (stages < 2) { stages = (SMEM_LIMIT - fixed) / stage_bytes; if (stages > 8) stages = 8; if (stages > nkb) stages = nkb; }
    if (stages < 1) stages = 1;
    pl->kind = 2; pl->BN = BN; pl->stages = stages;
    pl->smem = (size_t)stages * stage_bytes + fixed;
    pl->grid[0] = grid.x; pl->grid[1] = grid.y; pl->grid[2] = grid.z;
    switch (BN) {
        case 64: return set_attr_wgrad<64>();
        case 128: return set_attr_wgrad<128>();
        case 256: return set_attr_wgrad<256>();
    }
    cenn_set_error("unsupported wgrad BN %d", BN);
    return 1;
}

int pick_bn(int n_valid, int m_tiles, int sm_count) {
    // widest N tile (256-wide MMAs run the tensor pipe at full rate with the least smem traffic) that still
    // leaves at least one tile per SM for the persistent kernel
    // a wider N tile halves the operand bytes per FLOP; it is kept as long as its tiles cover >= 85 % of the SMs (measured on B200,
    // image step: 3.248 ms at 0.85 vs 3.295 ms at 1.0; E4 / D3 / G3-dgrad move from 256 tiles of 128 x 128 to 128 tiles of 128 x 256)
    static const double frac = getenv("CENN_BN_FRAC") ? atof(getenv("CENN_BN_FRAC")) : 0.85;
    int bn = 256;
    while (bn > 32 && (bn / 2 >= n_valid)) bn /= 2;
    while (bn > 64 && (double)m_tiles * ((n_valid + bn - 1) / bn) < frac * sm_count) bn /= 2;
    return bn;
}

void fill_epilogue(tc::GatherGemmParams &p, const TcEpilogue &ep, bf16 *out) {
    p.dbg = reinterpret_cast<unsigned long long *>(ep.dbg); p.dbg_flags = ep.dbg_flags;
    p.bias = ep.bias; p.stats = ep.stats; p.stats_stride = ep.stats_stride; p.act = ep.act; p.act_param = ep.act_param;
    p.out_bf16 = ep.no_bf16 ? nullptr : out;
    p.bwd_y = ep.bwd_y; p.bwd_scale = ep.bwd_scale; p.bwd_shift = ep.bwd_shift; p.bwd_mean = ep.bwd_mean; p.bwd_act = ep.bwd_act; p.bwd_negval = ep.bwd_negval;
}

// tap geometry of the 4x4 / stride-2 / pad-1 window: input row 2*oy - 1 + u = 2*(oy + DYS[u]) + PYS[u]
const int DYS[4] = {-1, 0, 0, 1}, PYS[4] = {1, 0, 1, 0};
// dgrad sub-pixel phases: output row 2y'+py receives taps u = UPH[py][a] from S row y' + DYP[py][a]
const int DYP[2][2] = {{0, -1}, {1, 0}};   // (taps are UPH[py][a] = {{1,3},{0,2}}, baked into repack_wt_kernel)

// ------------------------------------------------------------------ layout kernels
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float *__restrict__ src, bf16 *__restrict__ dst, int N, int C, int HW, int Cp) {
    int64_t total = (int64_t)N * HW * Cp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cp);
        int64_t t = i / Cp;
        int pix = (int)(t % HW), n = (int)(t / HW);
        dst[i] = __float2bfloat16(c < C ? src[((int64_t)n * C + c) * HW + pix] : 0.f);
    }
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16 *__restrict__ src, float *__restrict__ dst, int N, int C, int HW, int Cp) {
    int64_t total = (int64_t)N * C * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int pix = (int)(i % HW);
        int64_t t = i / HW;
        int c = (int)(t % C), n = (int)(t / C);
        dst[i] = __bfloat162float(src[((int64_t)n * HW + pix) * Cp + c]);
    }
}
__global__ void __launch_bounds__(256) repack_wf_kernel(const float *__restrict__ w, bf16 *__restrict__ Wf, int Cs, int Cl, int Clp, int kk) {
    int64_t total = (int64_t)Cs * kk * Clp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cl = (int)(i % Clp);
        int64_t t = i / Clp;
        int tap = (int)(t % kk), cs = (int)(t / kk);
        Wf[i] = __float2bfloat16(cl < Cl ? w[((int64_t)cs * Cl + cl) * kk + tap] : 0.f);
    }
}
__global__ void __launch_bounds__(256) repack_wt_kernel(const float *__restrict__ w, bf16 *__restrict__ Wt, int Cs, int Cl, int Csp, int cl_rows) {
    int64_t total = (int64_t)4 * cl_rows * 4 * Csp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cs = (int)(i % Csp);
        int64_t t = i / Csp;
        int ab = (int)(t % 4); t /= 4;
        int cl = (int)(t % cl_rows), ph = (int)(t / cl_rows);
        int py = ph >> 1, px = ph & 1, a = ab >> 1, b = ab & 1;
        const int U[2][2] = {{1, 3}, {0, 2}};
        int u = U[py][a], v = U[px][b];
        Wt[i] = __float2bfloat16((cs < Cs && cl < Cl) ? w[((int64_t)cs * Cl + cl) * 16 + u * 4 + v] : 0.f);
    }
}
__global__ void __launch_bounds__(256) repack_wtp_kernel(const float *__restrict__ w, bf16 *__restrict__ Wtp, int Cs, int Cl, int Clp, int Csp, int kk) {
    int64_t total = (int64_t)kk * Clp * Csp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cs = (int)(i % Csp);
        int64_t t = i / Csp;
        int cl = (int)(t % Clp), tap = (int)(t / Clp);
        Wtp[i] = __float2bfloat16((cs < Cs && cl < Cl) ? w[((int64_t)cs * Cl + cl) * kk + tap] : 0.f);
    }
}
__global__ void __launch_bounds__(256) unpack_grad_add_kernel(const float *__restrict__ g, float *__restrict__ gw, int Cs, int Cl, int Clp, int kk) {
    int64_t total = (int64_t)Cs * Cl * kk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int tap = (int)(i % kk);
        int64_t t = i / kk;
        int cl = (int)(t % Cl), cs = (int)(t / Cl);
        gw[i] += g[((int64_t)cs * kk + tap) * Clp + cl];
    }
}

}  // namespace

#define LAUNCH_1D(s, kern, total, ...) do { int64_t _t = (total); if (_t > 0) { kern<<<bw_grid(s, _t, 256), 256, 0, (s)->stream>>>(__VA_ARGS__); CK_LAUNCH(s); } } while (0)

int tc_nchw_to_nhwc(cenn_state *s, const float *src, bf16 *dst, int N, int C, int H, int W, int Cp) {
    LAUNCH_1D(s, nchw_to_nhwc_kernel, (int64_t)N * H * W * Cp, src, dst, N, C, H * W, Cp); return 0;
}
int tc_nhwc_to_nchw(cenn_state *s, const bf16 *src, float *dst, int N, int C, int H, int W, int Cp, const float *) {
    LAUNCH_1D(s, nhwc_to_nchw_kernel, (int64_t)N * C * H * W, src, dst, N, C, H * W, Cp); return 0;
}
int tc_repack_wf(cenn_state *s, const float *w, bf16 *Wf, int Cs, int Cl, int Clp, int kk) {
    LAUNCH_1D(s, repack_wf_kernel, (int64_t)Cs * kk * Clp, w, Wf, Cs, Cl, Clp, kk); return 0;
}
int tc_repack_wt(cenn_state *s, const float *w, bf16 *Wt, int Cs, int Cl, int Csp, int cl_rows) {
    LAUNCH_1D(s, repack_wt_kernel, (int64_t)16 * cl_rows * Csp, w, Wt, Cs, Cl, Csp, cl_rows); return 0;
}
int tc_repack_wtp(cenn_state *s, const float *w, bf16 *Wtp, int Cs, int Cl, int Clp, int Csp, int kk) {
    LAUNCH_1D(s, repack_wtp_kernel, (int64_t)kk * Clp * Csp, w, Wtp, Cs, Cl, Clp, Csp, kk); return 0;
}
int tc_unpack_grad_add(cenn_state *s, const float *g, float *gw, int Cs, int Cl, int Clp, int kk) {
    LAUNCH_1D(s, unpack_grad_add_kernel, (int64_t)Cs * Cl * kk, g, gw, Cs, Cl, Clp, kk); return 0;
}

// ------------------------------------------------------------------ plan storage helpers
static_assert(sizeof(CUtensorMap) <= 128, "CUtensorMap larger than the plan slot");
static_assert(sizeof(tc::GatherGemmParams) <= 1664 && sizeof(tc::WgradParams) <= 1664 && sizeof(tc::PatchDgradParams) <= 1664, "kernel params larger than the plan slot");
static CUtensorMap *planA(TcPlan *pl) { return reinterpret_cast<CUtensorMap *>(pl->tmA); }
static CUtensorMap *planB(TcPlan *pl) { return reinterpret_cast<CUtensorMap *>(pl->tmB); }
static CUtensorMap *planO(TcPlan *pl) { return reinterpret_cast<CUtensorMap *>(pl->tmO); }

void tc_plan_free(TcPlan *pl) {
    if (pl && pl->kb_dev) { cudaFree(pl->kb_dev); pl->kb_dev = nullptr; }
}

template <int BN>
static int launch_gather_t(cenn_state *s, const TcPlan *pl) {
    const tc::GatherGemmParams &p = *reinterpret_cast<const tc::GatherGemmParams *>(pl->params);
    if (p.dbg || p.dbg_flags)   // tools/gemm_probe.py only
        tc::gather_gemm_kernel<BN, true><<<dim3(pl->grid[0], pl->grid[1], pl->grid[2]), tc::GEMM_THREADS, pl->smem, s->stream>>>(
            *reinterpret_cast<const CUtensorMap *>(pl->tmA), *reinterpret_cast<const CUtensorMap *>(pl->tmB), *reinterpret_cast<const CUtensorMap *>(pl->tmO), p, pl->stages);
    else
        LK(tc::gather_gemm_kernel<BN, false>, dim3(pl->grid[0], pl->grid[1], pl->grid[2]), dim3(tc::GEMM_THREADS), pl->smem, s->stream)(
            *reinterpret_cast<const CUtensorMap *>(pl->tmA), *reinterpret_cast<const CUtensorMap *>(pl->tmB), *reinterpret_cast<const CUtensorMap *>(pl->tmO), p, pl->stages);
    CK_LAUNCH(s);
    return 0;
}
template <int BN>
static int launch_wgrad_t(cenn_state *s, const TcPlan *pl) {
    const tc::WgradParams &p = *reinterpret_cast<const tc::WgradParams *>(pl->params);
    LK(tc::wgrad_gemm_kernel<BN>, dim3(pl->grid[0], pl->grid[1], pl->grid[2]), dim3(tc::GEMM_THREADS), pl->smem, s->stream)(
        *reinterpret_cast<const CUtensorMap *>(pl->tmA), *reinterpret_cast<const CUtensorMap *>(pl->tmB), p, pl->stages);
    CK_LAUNCH(s);
    return 0;
}
template <int BN>
static int launch_patch_dgrad_t(cenn_state *s, const TcPlan *pl) {
    const tc::PatchDgradParams &p = *reinterpret_cast<const tc::PatchDgradParams *>(pl->params);
    LK(tc::patch_dgrad_kernel<BN>, dim3(pl->grid[0], 1, 1), dim3(tc::GEMM_THREADS), pl->smem, s->stream)(
        *reinterpret_cast<const CUtensorMap *>(pl->tmA), *reinterpret_cast<const CUtensorMap *>(pl->tmB), *reinterpret_cast<const CUtensorMap *>(pl->tmO), p, pl->stages);
    CK_LAUNCH(s);
    return 0;
}
int tc_launch(cenn_state *s, const TcPlan *pl) {
    if (pl->kind == 3) {
        switch (pl->BN) {
            case 16: return launch_patch_dgrad_t<16>(s, pl);
            case 64: return launch_patch_dgrad_t<64>(s, pl);
            case 128: return launch_patch_dgrad_t<128>(s, pl);
        }
    }
    if (pl->kind == 1) {
        switch (pl->BN) {
            case 32: return launch_gather_t<32>(s, pl);
            case 64: return launch_gather_t<64>(s, pl);
            case 128: return launch_gather_t<128>(s, pl);
            case 256: return launch_gather_t<256>(s, pl);
        }
    } else if (pl->kind == 2) {
        switch (pl->BN) {
            case 64: return launch_wgrad_t<64>(s, pl);
            case 128: return launch_wgrad_t<128>(s, pl);
            case 256: return launch_wgrad_t<256>(s, pl);
        }
    }
    cenn_set_error("tc_launch: plan not initialised (kind %d, BN %d)", pl->kind, pl->BN);
    return 1;
}

// ------------------------------------------------------------------ P1: fprop-type
int tc_plan_fprop_s2(cenn_state *s, TcPlan *pl, const bf16 *L, const bf16 *Wf, bf16 *S, int N, int h, int w, int Cs, int Csp, int Clp, const TcEpilogue &ep) {
    REQUIRE(Clp % 64 == 0 && Csp % 8 == 0, "tc_fprop_s2: Clp must be a multiple of 64 and Csp of 8 (got %d, %d)", Clp, Csp);
    int bw, bh, bn;
    choose_box(w, h, 128, bw, bh, bn);
    if (map_gather(planA(pl), L, N, 2 * h, 2 * w, Clp, bw, bh, bn)) return 1;
    tc::GatherGemmParams p = {};
    p.box_w = bw; p.box_h = bh; p.box_n = bn; p.bw_log2 = ilog2(bw); p.bh_log2 = ilog2(bh);
    p.tiles_x = (w + bw - 1) / bw; p.tiles_y = (h + bh - 1) / bh;
    int tiles_n = (N + bn - 1) / bn, m_tiles = p.tiles_x * p.tiles_y * tiles_n;
    int BN = pick_bn(Cs, m_tiles, s->sm_count);
    if (Csp < 64) BN = 32;                          // output rows narrower than one 128-byte slab: direct-store path
    if (map_2d(planB(pl), Wf, (uint64_t)16 * Clp, (uint64_t)Cs, BN)) return 1;
    memset(pl->tmO, 0, sizeof(pl->tmO));
    if (BN >= 64 && map_plain(planO(pl), S, N, h, w, Csp, bw, bh, bn)) return 1;
    p.o_cols = Csp; p.num_taps = 16;
    int chunks = Clp / 64;
    for (int t = 0; t < 16; ++t) {
        int u = t / 4, v = t % 4;
        p.A0[0][t] = PYS[v] * Clp; p.A1[0][t] = DYS[v]; p.A2[0][t] = PYS[u]; p.A3[0][t] = DYS[u];
    }
    p.chunks = chunks; p.bk_per_tap = Clp; p.num_kb = 16 * chunks;
    p.m_tiles = m_tiles; p.n_tiles = (Cs + BN - 1) / BN; p.num_phases = 1;
    p.out_w = w; p.out_h = h; p.out_n = N; p.n_valid = Cs;
    p.sX = Csp; p.sY = (long long)w * Csp; p.sN = (long long)h * w * Csp;
    fill_epilogue(p, ep, S);
    REQUIRE(!ep.bwd_y || BN >= 64, "tc_fprop_s2: BN-backward sums need an N tile of 64 or more");
    memcpy(pl->params, &p, sizeof(p));
    pl->flops = 2.0 * N * h * w * (double)Cs * 16.0 * Clp;
    return config_gather(s, pl, BN, p.num_kb, p.m_tiles * p.n_tiles);
}

// ------------------------------------------------------------------ P2: dgrad-type (4 sub-pixel phases)
int tc_plan_dgrad_s2(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *Wt, bf16 *L, int N, int h, int w, int Csp, int Cl, int Clp, int cl_rows, const TcEpilogue &ep) {
    REQUIRE(Csp % 64 == 0, "tc_dgrad_s2: Csp must be a multiple of 64 (got %d)", Csp);
    if (!ep.dbg && !ep.dbg_flags) {      // the patch kernel covers every layer with an 8 x 8 or larger small side
        int rc = tc_plan_dgrad_patch(s, pl, S, Wt, L, N, h, w, Csp, Cl, Clp, cl_rows, ep);
        if (rc != 2) return rc;
    }
    int bw, bh, bn;
    choose_box(w, h, 128, bw, bh, bn);
    if (map_plain(planA(pl), S, N, h, w, Csp, bw, bh, bn)) return 1;
    tc::GatherGemmParams p = {};
    p.box_w = bw; p.box_h = bh; p.box_n = bn; p.bw_log2 = ilog2(bw); p.bh_log2 = ilog2(bh);
    p.tiles_x = (w + bw - 1) / bw; p.tiles_y = (h + bh - 1) / bh;
    int tiles_n = (N + bn - 1) / bn, m_tiles = p.tiles_x * p.tiles_y * tiles_n;
    int BN = pick_bn(Cl, m_tiles * 4, s->sm_count);
    if (Clp % 64 != 0) BN = 32;                     // a 64-column slab would spill into the neighbouring sub-pixel: direct-store path
    REQUIRE(cl_rows >= Cl, "tc_dgrad_s2: cl_rows (%d) too small for Cl %d", cl_rows, Cl);
    const bool b_mn = ep.b_mn && BN >= 64;
    REQUIRE(!ep.b_mn || b_mn, "tc_dgrad_s2: MN-major weights need an N tile of 64 or more (Clp %d)", Clp);
    if (b_mn ? map_bmn(planB(pl), Wt, (uint64_t)16 * Clp, (uint64_t)Csp) : map_2d(planB(pl), Wt, (uint64_t)4 * Csp, (uint64_t)4 * cl_rows, BN)) return 1;
    memset(pl->tmO, 0, sizeof(pl->tmO));
    if (BN >= 64 && map_gather(planO(pl), L, N, 2 * h, 2 * w, Clp, bw, bh, bn)) return 1;
    p.o_cols = Clp; p.num_taps = 4; p.b_mn = b_mn ? 1 : 0;
    for (int ph = 0; ph < 4; ++ph) { p.O0[ph] = (ph & 1) * Clp; p.O2[ph] = ph >> 1; }
    int chunks = Csp / 64;
    for (int ph = 0; ph < 4; ++ph) {
        int py = ph >> 1, px = ph & 1;
        for (int ab = 0; ab < 4; ++ab) {
            int a = ab >> 1, b = ab & 1;
            p.A0[ph][ab] = 0; p.A1[ph][ab] = DYP[px][b]; p.A2[ph][ab] = 0; p.A3[ph][ab] = DYP[py][a];
            p.Bc[ph][ab] = (UPH[py][a] * 4 + UPH[px][b]) * Clp;
        }
        p.B1[ph] = ph * cl_rows;
    }
    p.chunks = chunks; p.bk_per_tap = Csp; p.num_kb = 4 * chunks;
    p.m_tiles = m_tiles; p.n_tiles = (Cl + BN - 1) / BN; p.num_phases = 4;
    p.out_w = w; p.out_h = h; p.out_n = N; p.n_valid = Cl;
    int W2 = 2 * w, H2 = 2 * h;
    p.sX = 2LL * Clp; p.sY = 2LL * W2 * Clp; p.sN = (long long)H2 * W2 * Clp;
    for (int ph = 0; ph < 4; ++ph) p.phase_off[ph] = ((long long)(ph >> 1) * W2 + (ph & 1)) * Clp;
    fill_epilogue(p, ep, L);
    REQUIRE(!ep.bwd_y || BN >= 64, "tc_dgrad_s2: BN-backward sums need an N tile of 64 or more");
    memcpy(pl->params, &p, sizeof(p));
    pl->flops = 2.0 * N * h * w * (double)Cl * 16.0 * Csp;
    return config_gather(s, pl, BN, p.num_kb, p.m_tiles * p.n_tiles * 4);
}

// ------------------------------------------------------------------ P2b: dgrad-type, patch kernel
template <int BN>
static int set_attr_patch_dgrad() {
    static bool done = false;
    if (!done) { CK(cudaFuncSetAttribute(tc::patch_dgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)); done = true; }
    return 0;
}
int tc_plan_dgrad_patch(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *Wt, bf16 *L, int N, int h, int w, int Csp, int Cl, int Clp, int cl_rows, const TcEpilogue &ep) {
    const bool thin = Clp == 4 || Clp == 16;
    if (w < 8 || h < 8 || Csp % 64 != 0 || !(thin || Clp % 64 == 0) || (thin && ep.stats) || getenv("CENN_NO_PATCH")) return 2;
    const int bh = pow2_le(h, 16), bn = 16 / bh;
    const int BN = thin ? 16 : ((Clp % 128 == 0 && !getenv("CENN_PATCH_BN64")) ? 128 : 64);
    REQUIRE(cl_rows >= Cl, "tc_dgrad_patch: cl_rows (%d) too small for Cl %d", cl_rows, Cl);
    {   // S patch: dims (c, x, n, y) so that shared memory holds [y][n][x] rows of 64 channels
        uint64_t d[4] = {(uint64_t)Csp, (uint64_t)w, (uint64_t)N, (uint64_t)h};
        uint64_t st[3] = {(uint64_t)Csp * 2, (uint64_t)h * w * Csp * 2, (uint64_t)w * Csp * 2};
        uint32_t bx[4] = {64, 10, (uint32_t)bn, (uint32_t)bh + 2};
        if (make_map(planA(pl), S, 4, d, st, bx)) return 1;
    }
    const bool b_mn = ep.b_mn && BN >= 64;
    REQUIRE(!ep.b_mn || b_mn, "tc_dgrad_patch: MN-major weights need an N tile of 64 or more (Clp %d)", Clp);
    if (b_mn ? map_bmn(planB(pl), Wt, (uint64_t)16 * Clp, (uint64_t)Csp) : map_2d(planB(pl), Wt, (uint64_t)4 * Csp, (uint64_t)4 * cl_rows, BN)) return 1;
    memset(pl->tmO, 0, sizeof(pl->tmO));
    const int H2 = 2 * h, W2 = 2 * w;
    if (BN >= 64) {   // output: dims (px*Clp + c, x', n, y', py)
        uint64_t d[5] = {(uint64_t)2 * Clp, (uint64_t)w, (uint64_t)N, (uint64_t)h, 2};
        uint64_t st[4] = {(uint64_t)2 * Clp * 2, (uint64_t)H2 * W2 * Clp * 2, (uint64_t)2 * W2 * Clp * 2, (uint64_t)W2 * Clp * 2};
        uint32_t bx[5] = {64, 8, (uint32_t)bn, (uint32_t)bh, 1};
        if (make_map(planO(pl), L, 5, d, st, bx)) return 1;
    }
    tc::PatchDgradParams p = {};
    p.chunks = Csp / 64;
    p.tiles_x = (w + 7) / 8; p.tiles_y = (h + bh - 1) / bh;
    p.m_tiles = p.tiles_x * p.tiles_y * ((N + bn - 1) / bn);
    p.n_tiles = thin ? 1 : (Cl + BN - 1) / BN;
    p.bh = bh; p.bn = bn; p.bn_log2 = ilog2(bn);
    p.patch_bytes = 128 * 10 * bn * (bh + 2);
    p.patch_stride = (p.patch_bytes + 1023) / 1024 * 1024;
    for (int ph = 0; ph < 4; ++ph)
        for (int ab = 0; ab < 4; ++ab) {
            const int dy = DYP[ph >> 1][ab >> 1], dx = DYP[ph & 1][ab & 1];
            p.a_off[ph][ab] = (((1 + dy) * bn) * 10 + (1 + dx)) * 128;
            p.b_col[ph][ab] = (UPH[ph >> 1][ab >> 1] * 4 + UPH[ph & 1][ab & 1]) * Clp;
        }
    p.b_mn = b_mn ? 1 : 0;
    p.Csp = Csp; p.cl_rows = cl_rows;
    p.out_w = w; p.out_h = h; p.out_n = N; p.n_valid = Cl; p.Clp = Clp; p.H2 = H2; p.W2 = W2;
    p.out = ep.no_bf16 ? nullptr : L;
    p.bias = ep.bias; p.stats = ep.stats; p.stats_stride = ep.stats_stride; p.act = ep.act; p.act_param = ep.act_param;
    p.bwd_y = ep.bwd_y; p.bwd_scale = ep.bwd_scale; p.bwd_shift = ep.bwd_shift; p.bwd_mean = ep.bwd_mean; p.bwd_act = ep.bwd_act; p.bwd_negval = ep.bwd_negval;
    REQUIRE(!ep.bwd_y || BN >= 64, "tc_dgrad_patch: BN-backward sums need an N tile of 64 or more");
    memcpy(pl->params, &p, sizeof(p));
    pl->flops = 2.0 * N * h * w * (double)Cl * 16.0 * Csp;
    const int bstage = 4 * BN * 128, out_bytes = BN >= 64 ? 128 * BN * 2 * (BN >= 128 ? 1 : 2) : 0;
    int npatch = 2, stages;
    // thin outputs (3 / 12 channels): the weights of all phases are 8 KB per chunk -- resident for the whole CTA -- and the kernel is bound
    // by the number of S patches (one HBM round trip each) in flight: as many patch buffers as shared memory holds (up to 6)
    const bool resident = thin && p.n_tiles == 1 && 4 * p.chunks <= 8 && getenv("CENN_PATCH_NO_RESIDENT") == nullptr;
    const int base_fixed = 1024 + out_bytes + 38 * 8 + 6 * BN * 4 + 64;
    int per_sm = 1;
    if (resident) {
        stages = 4 * p.chunks;
        // measured (round 2): six patches in flight barely help (E1's dead dgrad 133 -> 126 us): the variant is bound by its two single-thread issue
        // loops and its per-lane epilogue (55 tiles of 2.3 us per CTA), so two CTAs share an SM when half the shared memory holds >= 2 patches
        const int half = (SMEM_LIMIT + 1024) / 2 - 1024;
        const int np2 = (half - base_fixed - stages * bstage) / p.patch_stride;
        if (np2 >= 2 && getenv("CENN_THIN_ONE_CTA") == nullptr) { per_sm = 2; npatch = np2 > 4 ? 4 : np2; }
        else { npatch = (SMEM_LIMIT - base_fixed - stages * bstage) / p.patch_stride; if (npatch > 6) npatch = 6; }
        REQUIRE(npatch >= 2, "tc_dgrad_patch: shared memory too small for two patches (thin)");
    } else {
        stages = (SMEM_LIMIT - base_fixed - 2 * p.patch_stride) / bstage;
        if (stages > 8) stages = 8;
        REQUIRE(stages >= 2, "tc_dgrad_patch: shared memory too small (BN %d)", BN);
    }
    p.npatch = npatch; p.b_resident = resident ? 1 : 0;
    memcpy(pl->params, &p, sizeof(p));
    const int fixed = base_fixed + npatch * p.patch_stride;
    pl->kind = 3; pl->BN = BN; pl->stages = stages;
    pl->smem = (size_t)stages * bstage + fixed;
    const int total = p.m_tiles * p.n_tiles;
    pl->grid[0] = (unsigned)(total < s->sm_count * per_sm ? total : s->sm_count * per_sm); pl->grid[1] = 1; pl->grid[2] = 1;
    switch (BN) {
        case 16: return set_attr_patch_dgrad<16>();
        case 64: return set_attr_patch_dgrad<64>();
        default: return set_attr_patch_dgrad<128>();
    }
}

// ------------------------------------------------------------------ P4: plain GEMM
int tc_plan_gemm(cenn_state *s, TcPlan *pl, const bf16 *A, const bf16 *B, bf16 *out, int M, int Nc, int K, int ldo, const TcEpilogue &ep, int lda) {
    REQUIRE(K % 8 == 0, "tc_gemm: K must be a multiple of 8 (got %d)", K);
    if (lda <= 0) lda = K;                 // row pitch of A in elements (a column slice of a wider matrix: lda > K, columns >= K are never read)
    REQUIRE(lda >= K && lda % 8 == 0, "tc_gemm: lda must be a multiple of 8 and >= K (got %d, K %d)", lda, K);
    uint64_t rowsA = (uint64_t)(M > 128 ? M : 128);
    uint64_t dims[5] = {(uint64_t)K, rowsA, 1, 1, 1};
    uint64_t st[4] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * rowsA, (uint64_t)lda * 2 * rowsA, (uint64_t)lda * 2 * rowsA};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    (void)rowsA;
    dims[1] = (uint64_t)M;                 // the true extent: rows >= M are out of bounds -> zero filled
    if (make_map(planA(pl), A, 5, dims, st, box)) return 1;
    tc::GatherGemmParams p = {};
    p.box_w = 128; p.box_h = 1; p.box_n = 1; p.bw_log2 = 7; p.bh_log2 = 0;
    p.tiles_x = (M + 127) / 128; p.tiles_y = 1;
    int m_tiles = p.tiles_x;
    int BN = pick_bn(Nc, m_tiles, s->sm_count);
    if (ldo < 64 || ldo % 8 != 0) BN = 32;
    const bool b_mn = ep.b_mn && BN >= 64;     // B given as [K][Nc] (row stride Nc)
    REQUIRE(!ep.b_mn || b_mn, "tc_gemm: MN-major B needs an N tile of 64 or more");
    if (b_mn ? map_bmn(planB(pl), B, (uint64_t)Nc, (uint64_t)K) : map_2d(planB(pl), B, (uint64_t)K, (uint64_t)Nc, BN)) return 1;
    memset(pl->tmO, 0, sizeof(pl->tmO));
    if (BN >= 64) {
        uint64_t od[5] = {(uint64_t)ldo, (uint64_t)M, 1, 1, 1};
        uint64_t os[4] = {(uint64_t)ldo * 2, (uint64_t)ldo * 2 * M, (uint64_t)ldo * 2 * M, (uint64_t)ldo * 2 * M};
        uint32_t ob[5] = {64, 128, 1, 1, 1};
        if (make_map(planO(pl), out, 5, od, os, ob)) return 1;
    }
    p.o_cols = ldo; p.num_taps = 1; p.b_mn = b_mn ? 1 : 0;
    int nkb = (K + 63) / 64;
    p.chunks = nkb; p.bk_per_tap = 0; p.num_kb = nkb;          // a single "tap" whose chunks walk K
    p.m_tiles = m_tiles; p.n_tiles = (Nc + BN - 1) / BN; p.num_phases = 1;
    p.out_w = M; p.out_h = 1; p.out_n = 1; p.n_valid = Nc;
    p.sX = ldo; p.sY = 0; p.sN = 0;
    fill_epilogue(p, ep, out);
    REQUIRE(!ep.bwd_y || BN >= 64, "tc_gemm: BN-backward sums need an N tile of 64 or more");
    memcpy(pl->params, &p, sizeof(p));
    pl->flops = 2.0 * M * (double)Nc * K;
    return config_gather(s, pl, BN, p.num_kb, p.m_tiles * p.n_tiles);
}

// ------------------------------------------------------------------ P3 / P5: wgrad
static int wgrad_common(cenn_state *s, TcPlan *pl, tc::WgradParams &p, int num_taps, int Cs, int Clp, int num_kb_total) {
    int chunks = Clp / 64;
    int row_blocks = num_taps * chunks;
    int m_tiles = (row_blocks + 1) / 2;
    int BN = Cs > 128 ? 256 : (Cs > 64 ? 128 : 64);
    int n_tiles = (Cs + BN - 1) / BN;
    int tiles = m_tiles * n_tiles;
    // split K (pixels) over CTAs: minimise waves x (k-blocks per CTA + fixed per-CTA cost), one CTA per SM at a time
    const int max_splits = std::max(1, (num_kb_total + 3) / 4);            // at least 4 k-blocks (256 pixels) per split
    const int fixed_kb = 10;                                               // prologue + epilogue of a CTA, in k-block units
    int splits = 1; long long best = -1;
    for (int sp = 1; sp <= max_splits && sp <= 4 * s->sm_count; ++sp) {
        const long long slots = 2LL * s->sm_count;                          // two CTAs per SM
        const long long waves = ((long long)tiles * sp + slots - 1) / slots;
        const long long cost = waves * ((num_kb_total + sp - 1) / sp + fixed_kb);
        if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
    if (p.accumulate == 2) p.accumulate = splits == 1 ? 0 : 1;
    if (!p.accumulate) splits = 1;
    pl->overwrites = p.accumulate ? 0 : 1;
    p.num_kb_total = num_kb_total;
    p.kb_per_split = (num_kb_total + splits - 1) / splits;
    splits = (num_kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.chunks_per_tap = chunks;
    p.n_chunks = BN / 64;
    memcpy(pl->params, &p, sizeof(p));
    return config_wgrad(pl, BN, p.kb_per_split, dim3(m_tiles, n_tiles, splits));
}

int tc_plan_wgrad_s2(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *L, float *gW, int N, int h, int w, int Cs, int Csp, int Clp, float scale, int accumulate) {
    REQUIRE(Clp % 64 == 0 && Csp % 8 == 0, "tc_wgrad_s2: Clp must be a multiple of 64 and Csp of 8 (got %d, %d)", Clp, Csp);
    int bw, bh, bn;
    choose_box(w, h, 64, bw, bh, bn);
    if (map_gather(planA(pl), L, N, 2 * h, 2 * w, Clp, bw, bh, bn)) return 1;
    if (map_plain(planB(pl), S, N, h, w, Csp, bw, bh, bn)) return 1;
    tc::WgradParams p = {};
    p.box_w = bw; p.box_h = bh; p.box_n = bn;
    p.tiles_x = (w + bw - 1) / bw; p.tiles_y = (h + bh - 1) / bh;
    int tiles_n = (N + bn - 1) / bn;
    for (int t = 0; t < 16; ++t) {
        int u = t / 4, v = t % 4;
        p.gx0[t] = PYS[v] * Clp; p.gdx[t] = DYS[v]; p.g2[t] = PYS[u]; p.gdy[t] = DYS[u];
    }
    p.num_taps = 16;
    p.cl_stride = Clp; p.cl_valid = Clp; p.cs_valid = Cs; p.out_cs_stride = 16LL * Clp;
    p.out = gW; p.scale = scale; p.accumulate = accumulate;
    pl->flops = 2.0 * N * h * w * (double)Cs * 16.0 * Clp;
    return wgrad_common(s, pl, p, 16, Cs, Clp, p.tiles_x * p.tiles_y * tiles_n);
}

int tc_plan_wgrad_plain(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *L, float *gW, int M, int Cs, int Csp, int Clp, float scale, int accumulate) {
    REQUIRE(Clp % 64 == 0 && Csp % 8 == 0, "tc_wgrad_plain: Clp must be a multiple of 64 and Csp of 8 (got %d, %d)", Clp, Csp);
    if (map_plain(planA(pl), L, 1, 1, M, Clp, 64, 1, 1)) return 1;
    if (map_plain(planB(pl), S, 1, 1, M, Csp, 64, 1, 1)) return 1;
    tc::WgradParams p = {};
    p.box_w = 64; p.box_h = 1; p.box_n = 1;
    p.tiles_x = (M + 63) / 64; p.tiles_y = 1;
    p.num_taps = 1;
    p.cl_stride = Clp; p.cl_valid = Clp; p.cs_valid = Cs; p.out_cs_stride = Clp;
    p.out = gW; p.scale = scale; p.accumulate = accumulate;
    pl->flops = 2.0 * M * (double)Cs * Clp;
    return wgrad_common(s, pl, p, 1, Cs, Clp, p.tiles_x);
}

// ---- one-shot forms (plan, launch, free) used by the op-level wrappers --------------------------------------
#define ONE_SHOT(planexpr) do { TcPlan pl; int rc = (planexpr); if (!rc) rc = tc_launch(s, &pl); \
    if (pl.kb_dev) { cudaStreamSynchronize(s->stream); } tc_plan_free(&pl); return rc; } while (0)
int tc_fprop_s2(cenn_state *s, const bf16 *L, const bf16 *Wf, bf16 *S, int N, int h, int w, int Cs, int Csp, int Clp, const TcEpilogue &ep) {
    ONE_SHOT(tc_plan_fprop_s2(s, &pl, L, Wf, S, N, h, w, Cs, Csp, Clp, ep));
}
int tc_dgrad_s2(cenn_state *s, const bf16 *S, const bf16 *Wt, bf16 *L, int N, int h, int w, int Csp, int Cl, int Clp, int cl_rows, const TcEpilogue &ep) {
    ONE_SHOT(tc_plan_dgrad_s2(s, &pl, S, Wt, L, N, h, w, Csp, Cl, Clp, cl_rows, ep));
}
int tc_wgrad_s2(cenn_state *s, const bf16 *S, const bf16 *L, float *gW, int N, int h, int w, int Cs, int Csp, int Clp, float scale, int accumulate) {
    ONE_SHOT(tc_plan_wgrad_s2(s, &pl, S, L, gW, N, h, w, Cs, Csp, Clp, scale, accumulate));
}
int tc_gemm(cenn_state *s, const bf16 *A, const bf16 *B, bf16 *out, int M, int Nc, int K, int ldo, const TcEpilogue &ep) {
    ONE_SHOT(tc_plan_gemm(s, &pl, A, B, out, M, Nc, K, ldo, ep));
}
int tc_wgrad_plain(cenn_state *s, const bf16 *S, const bf16 *L, float *gW, int M, int Cs, int Csp, int Clp, float scale, int accumulate) {
    ONE_SHOT(tc_plan_wgrad_plain(s, &pl, S, L, gW, M, Cs, Csp, Clp, scale, accumulate));
}

// ------------------------------------------------------------------ NCHW fp32 wrappers (op-level ABI, CENN_BF16)
namespace {
struct Bump {          // bump allocator over the state's workspace
    uint8_t *base; size_t off = 0, cap;
    template <typename T> T *take(size_t n) { off = (off + 255) & ~size_t(255); T *p = reinterpret_cast<T *>(base + off); off += n * sizeof(T); return p; }
};
size_t al(size_t b) { return (b + 255) & ~size_t(255); }
bool is_s2(int kH, int kW, int dH, int dW, int pH, int pW, int H, int W) {
    return kH == 4 && kW == 4 && dH == 2 && dW == 2 && pH == 1 && pW == 1 && (H % 2 == 0) && (W % 2 == 0) && H >= 2 && W >= 2;
}
bool is_valid4(int kH, int kW, int dH, int dW, int pH, int pW, int H, int W) {   // 4x4 -> 1x1 bottleneck / head
    return kH == 4 && kW == 4 && dH == 1 && dW == 1 && pH == 0 && pW == 0 && H == 4 && W == 4;
}
}  // namespace

int tc_conv_fprop_nchw(cenn_state *s, const float *x, const float *w, const float *bias, float *out,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW) {
    bool s2 = is_s2(kH, kW, dH, dW, pH, pW, H, W), v4 = is_valid4(kH, kW, dH, dW, pH, pW, H, W);
    if (!s2 && !v4) return 1;
    int Clp = round_up(C, 64), Op = round_up(O, 8);
    int h = s2 ? H / 2 : 1, wd = s2 ? W / 2 : 1;
    size_t need = al((size_t)N * H * W * Clp * 2) + al((size_t)O * 16 * Clp * 2) + al((size_t)N * h * wd * Op * 2) + 4096;
    uint8_t *ws = (uint8_t *)cenn_workspace(s, need);
    if (!ws) return -1;
    Bump b{ws, 0, need};
    bf16 *xl = b.take<bf16>((size_t)N * H * W * Clp), *wf = b.take<bf16>((size_t)O * 16 * Clp), *so = b.take<bf16>((size_t)N * h * wd * Op);
    if (tc_nchw_to_nhwc(s, x, xl, N, C, H, W, Clp) || tc_repack_wf(s, w, wf, O, C, Clp, 16)) return -1;
    TcEpilogue ep; ep.bias = bias;
    int rc = s2 ? tc_fprop_s2(s, xl, wf, so, N, h, wd, O, Op, Clp, ep) : tc_gemm(s, xl, wf, so, N, O, 16 * Clp, Op, ep);
    if (rc) return -1;
    if (tc_nhwc_to_nchw(s, so, out, N, O, h, wd, Op, nullptr)) return -1;
    return 0;
}

int tc_conv_dgrad_nchw(cenn_state *s, const float *gy, const float *w, const float *bias, float *gx,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW) {
    bool s2 = is_s2(kH, kW, dH, dW, pH, pW, H, W), v4 = is_valid4(kH, kW, dH, dW, pH, pW, H, W);
    if (!s2 && !v4) return 1;
    int Csp = round_up(O, 64), Clp = round_up(C, 8);
    int h = s2 ? H / 2 : 1, wd = s2 ? W / 2 : 1;
    size_t wbytes = s2 ? (size_t)16 * Clp * Csp * 2 : (size_t)16 * Clp * Csp * 2;
    size_t need = al((size_t)N * h * wd * Csp * 2) + al(wbytes) + al((size_t)N * H * W * Clp * 2) + 4096;
    uint8_t *ws = (uint8_t *)cenn_workspace(s, need);
    if (!ws) return -1;
    Bump b{ws, 0, need};
    bf16 *gs = b.take<bf16>((size_t)N * h * wd * Csp), *wt = b.take<bf16>(wbytes / 2), *lo = b.take<bf16>((size_t)N * H * W * Clp);
    if (tc_nchw_to_nhwc(s, gy, gs, N, O, h, wd, Csp)) return -1;
    TcEpilogue ep; ep.bias = bias;
    int rc;
    if (s2) {
        if (tc_repack_wt(s, w, wt, O, C, Csp, Clp)) return -1;
        rc = tc_dgrad_s2(s, gs, wt, lo, N, h, wd, Csp, C, Clp, Clp, ep);
    } else {
        // gx[n, (t, c)] = sum_o gy[n, o] * w[o][c][t]; bias (full-conv fprop) is per c -> expand handled by caller path below
        if (tc_repack_wtp(s, w, wt, O, C, Clp, Csp, 16)) return -1;
        if (bias) {
            // per-(t,c) bias vector = bias[c] repeated over the 16 taps
            float *bexp = (float *)cenn_workspace2(s, (size_t)16 * Clp * sizeof(float));
            if (!bexp) return -1;
            std::vector<float> hb(16 * Clp, 0.f), hbias(C);
            if (cudaMemcpyAsync(hbias.data(), bias, C * sizeof(float), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return -1;
            cudaStreamSynchronize(s->stream);
            for (int t = 0; t < 16; ++t) for (int c = 0; c < C; ++c) hb[t * Clp + c] = hbias[c];
            cudaMemcpyAsync(bexp, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream);
            cudaStreamSynchronize(s->stream);
            ep.bias = bexp;
        }
        rc = tc_gemm(s, gs, wt, lo, N, 16 * Clp, Csp, 16 * Clp, ep);
    }
    if (rc) return -1;
    if (tc_nhwc_to_nchw(s, lo, gx, N, C, H, W, Clp, nullptr)) return -1;
    return 0;
}

int tc_conv_wgrad_nchw(cenn_state *s, const float *x, const float *gy, float *gw,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW, float scale, int) {
    bool s2 = is_s2(kH, kW, dH, dW, pH, pW, H, W), v4 = is_valid4(kH, kW, dH, dW, pH, pW, H, W);
    if (!s2 && !v4) return 1;
    int Clp = round_up(C, 64), Csp = round_up(O, 8);
    int h = s2 ? H / 2 : 1, wd = s2 ? W / 2 : 1;
    size_t need = al((size_t)N * H * W * Clp * 2) + al((size_t)N * h * wd * Csp * 2) + al((size_t)O * 16 * Clp * 4) + 4096;
    uint8_t *ws = (uint8_t *)cenn_workspace(s, need);
    if (!ws) return -1;
    Bump b{ws, 0, need};
    bf16 *xl = b.take<bf16>((size_t)N * H * W * Clp), *gs = b.take<bf16>((size_t)N * h * wd * Csp);
    float *g = b.take<float>((size_t)O * 16 * Clp);
    if (tc_nchw_to_nhwc(s, x, xl, N, C, H, W, Clp) || tc_nchw_to_nhwc(s, gy, gs, N, O, h, wd, Csp)) return -1;
    if (cudaMemsetAsync(g, 0, (size_t)O * 16 * Clp * 4, s->stream) != cudaSuccess) return -1;
    int rc = s2 ? tc_wgrad_s2(s, gs, xl, g, N, h, wd, O, Csp, Clp, scale, 1) : tc_wgrad_plain(s, gs, xl, g, N, O, Csp, 16 * Clp, scale, 1);
    if (rc) return -1;
    if (tc_unpack_grad_add(s, g, gw, O, C, Clp, 16)) return -1;
    return 0;
}

