// conv_tc.h -- BF16 tcgen05 implicit-GEMM convolution path (conv_tc.cu): NHWC-bf16 primitives used by the
// whole-step executor, and NCHW-fp32 wrappers used by the op-level ABI in precision mode CENN_BF16.
//
// Naming: every 4x4 / stride-2 / pad-1 layer relates a "small" tensor S [N,h,w,Cs] and a "large" tensor
// L [N,2h,2w,Cl] (conv: S = output, L = input; full conv: S = input, L = output).  THNN weights are
// [Cs][Cl][4][4] for both module types; the internal "master" layout is [Cs][tap][Clp] (tap = u*4+v).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

typedef __nv_bfloat16 bf16;

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct TcEpilogue {
    const float *bias = nullptr;   // per output channel
    float *stats = nullptr;        // [2][stats_stride] fp32 sum / sum-of-squares accumulators (atomics)
    int stats_stride = 0;
    int act = 0;                   // tc::ACT_*
    float act_param = 0.f;
    bool no_bf16 = false;
    void *dbg = nullptr;           // optional per-role cycle counters (debug probe)
    int dbg_flags = 0;             // probe only, see tc::GatherGemmParams::dbg_flags
    // BN-backward sums of the consumer block taken in the epilogue (tc::GatherGemmParams::bwd_y): `stats` / `stats_stride` name the
    // [2][stride] accumulator (fp32 atomics; zeroed by whoever consumes it), N tile >= 64 only
    const bf16 *bwd_y = nullptr;
    const float *bwd_scale = nullptr, *bwd_shift = nullptr, *bwd_mean = nullptr;
    int bwd_act = 0;
    float bwd_negval = 0.2f;
    bool b_mn = false;             // dgrad-type / plain GEMM: the weight operand is the MASTER layout Wf [Cs][16*Clp] (or [K][Nc], row stride
                                   // = Nc) read MN-major, instead of a transposed K-major copy (N tile >= 64 only)
};

// A prepared launch: tensor maps, kernel parameters, k-block table and grid are built once (shapes and
// buffer addresses are static in the executor) and replayed every step / captured into a CUDA graph.
struct TcPlan {
    int kind = 0;                 // 1 = K-major gather GEMM, 2 = MN-major wgrad GEMM, 3 = patch kernel (dgrad type)
    int BN = 0, stages = 0;
    unsigned grid[3] = {1, 1, 1};
    size_t smem = 0;
    alignas(64) unsigned char tmA[128];
    alignas(64) unsigned char tmB[128];
    alignas(64) unsigned char tmO[128];   // output tile store (gather GEMM, BN >= 64)
    alignas(16) unsigned char params[1664];
    void *kb_dev = nullptr;       // owned device k-block table (gather GEMM)
    double flops = 0;             // algorithmic FLOPs of one launch
    int overwrites = 0;           // wgrad: 1 = the launch stores its result (no accumulation; the target need not be zeroed)
};
int tc_plan_fprop_s2(cenn_state *s, TcPlan *pl, const bf16 *L, const bf16 *Wf, bf16 *S, int N, int h, int w, int Cs, int Csp, int Clp, const TcEpilogue &ep);
int tc_plan_dgrad_s2(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *Wt, bf16 *L, int N, int h, int w, int Csp, int Cl, int Clp, int cl_rows, const TcEpilogue &ep);
// patch variant of the dgrad type (w, h >= 8; Csp % 64 == 0; Clp % 64 == 0 or Clp in {4, 16}); returns 2 if the shape is not covered
int tc_plan_dgrad_patch(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *Wt, bf16 *L, int N, int h, int w, int Csp, int Cl, int Clp, int cl_rows, const TcEpilogue &ep);
int tc_plan_wgrad_s2(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *L, float *gW, int N, int h, int w, int Cs, int Csp, int Clp, float scale, int accumulate);
int tc_plan_gemm(cenn_state *s, TcPlan *pl, const bf16 *A, const bf16 *B, bf16 *out, int M, int Nc, int K, int ldo, const TcEpilogue &ep, int lda = 0);
int tc_plan_wgrad_plain(cenn_state *s, TcPlan *pl, const bf16 *S, const bf16 *L, float *gW, int M, int Cs, int Csp, int Clp, float scale, int accumulate);
int tc_launch(cenn_state *s, const TcPlan *pl);
void tc_plan_free(TcPlan *pl);

// ---- NHWC bf16 primitives (plan + launch + free; return 0 on success, 1 on error with message set) -------
// P1: S[pix,cs] = sum_{tap,cl} L[gather_tap(pix),cl] * Wf[cs][tap][cl]      (conv fprop, full-conv dgrad)
int tc_fprop_s2(cenn_state *s, const bf16 *L, const bf16 *Wf, bf16 *S, int N, int h, int w, int Cs, int Csp, int Clp, const TcEpilogue &ep);
// P2: L[pix',cl] = sum_{ab,cs} S[pix+d_ab,cs] * Wt[phase][cl][ab][cs]     (conv dgrad, full-conv fprop)
int tc_dgrad_s2(cenn_state *s, const bf16 *S, const bf16 *Wt, bf16 *L, int N, int h, int w, int Csp, int Cl, int Clp, int cl_rows, const TcEpilogue &ep);
// P3: gW[cs][tap][cl] (+)= scale * sum_pix S[pix,cs] * L[gather_tap(pix),cl]
//     accumulate: 1 = add into gW (red.add, split-K allowed); 0 = store (single split forced); 2 = store when one split
//     covers K anyway, otherwise add (the plan's `overwrites` says which)
int tc_wgrad_s2(cenn_state *s, const bf16 *S, const bf16 *L, float *gW, int N, int h, int w, int Cs, int Csp, int Clp, float scale, int accumulate);
// P4: out[M,Nc] = A[M,K] * B[Nc,K]^T  (both K-major; K multiple of 8; ldo = output row stride in elements)
int tc_gemm(cenn_state *s, const bf16 *A, const bf16 *B, bf16 *out, int M, int Nc, int K, int ldo, const TcEpilogue &ep);
// P5: gW[cs][cl] (+)= scale * sum_m S[m,cs] * L[m,cl]   (S: [M,Csp], L: [M,Clp], out row stride = Clp)
int tc_wgrad_plain(cenn_state *s, const bf16 *S, const bf16 *L, float *gW, int M, int Cs, int Csp, int Clp, float scale, int accumulate);

// ---- layout / repack kernels --------------------------------------------------------------------------------
int tc_nchw_to_nhwc(cenn_state *s, const float *src, bf16 *dst, int N, int C, int H, int W, int Cp);
int tc_nhwc_to_nchw(cenn_state *s, const bf16 *src, float *dst, int N, int C, int H, int W, int Cp, const float *bias_unused);
// THNN weight [Cs][Cl][kk] fp32 -> Wf bf16 [Cs][kk][Clp]
int tc_repack_wf(cenn_state *s, const float *w, bf16 *Wf, int Cs, int Cl, int Clp, int kk);
// THNN weight [Cs][Cl][16] -> Wt bf16 [4 phases][cl_rows][4 ab][Csp]
int tc_repack_wt(cenn_state *s, const float *w, bf16 *Wt, int Cs, int Cl, int Csp, int cl_rows);
// THNN weight [Cs][Cl][kk] -> plain transpose Wtp bf16 [(t,cl) rows = kk*Clp][Csp]
int tc_repack_wtp(cenn_state *s, const float *w, bf16 *Wtp, int Cs, int Cl, int Clp, int Csp, int kk);
// master-layout gradient [Cs][kk][Clp] fp32 -> THNN gradWeight[Cs][Cl][kk] += g
int tc_unpack_grad_add(cenn_state *s, const float *g, float *gradWeight, int Cs, int Cl, int Clp, int kk);

// ---- NCHW fp32 wrappers: 0 = done, >0 = shape not handled (caller uses the fp32 SIMT kernel), <0 = error ----
int tc_conv_fprop_nchw(cenn_state *s, const float *x, const float *w, const float *bias, float *out,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW);
int tc_conv_dgrad_nchw(cenn_state *s, const float *gy, const float *w, const float *bias, float *gx,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW);
int tc_conv_wgrad_nchw(cenn_state *s, const float *x, const float *gy, float *gw,
                       int N, int C, int H, int W, int O, int kH, int kW, int dH, int dW, int pH, int pW,
                       float scale, int overwrite);
