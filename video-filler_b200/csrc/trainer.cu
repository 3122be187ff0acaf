// trainer.cu -- whole-step executor: one optim.adam(fDx) + optim.adam(fGx) (train.lua:421-424,
// train_vid_weighted.lua:373-537) as a static program of kernel launches over preallocated NHWC bf16
// buffers.  Convolutions are tcgen05 implicit GEMMs (conv_tc.cu), everything else is a fused bandwidth
// kernel (nhwc.cuh).  The program is a flat list of ops; "sync" ops mark the buffers a data-parallel
// run must sum across ranks (BN statistics, losses, gradients).  On one GPU the list is captured into a
// CUDA graph.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <string>
#include <type_traits>
#include <vector>
#include "common.cuh"
#include "conv_tc.h"
#include "nhwc.cuh"
#include "tc_gemm.cuh"

namespace {

enum BlockType { CONV_S2, CONV_V4, FULL_S2, FULL_V4, HEAD, JOIN5 };

struct Tensor {
    bf16 *p = nullptr;
    int N = 0, H = 0, W = 0, C = 0, Cp = 0;
    int64_t elems() const { return (int64_t)N * H * W * Cp; }
    int64_t pix() const { return (int64_t)N * H * W; }
};

struct Block {
    BlockType type;
    int Cs = 0, Cl = 0, Csp = 0, Clp = 0;   // small-side / large-side channels (valid, padded)
    bool bn = false;
    int act = 0;
    bool thin = false;                      // large-side tensor is thin (explicit im2col)
    int h = 0, w = 0;                       // spatial size of the small side
    int Cout = 0, Coutp = 0;                // channels of the block output
    // parameter offsets (elements) in the net's master vector
    int64_t w_off = 0, b_off = 0, g_off = -1, be_off = -1, w_count = 0;
    // THNN flat offsets
    int64_t t_w_off = 0, t_b_off = 0, t_g_off = -1, t_be_off = -1;
    // noiseGen (train.lua:109-124), bottleneck block only: nz extra channels appended to the conv output by a 1x1 convolution of the noise
    // vector; Cout / Coutp / the BN cover Cs + nz channels, the block's own conv (weights, bias, GEMMs) the first Cs
    int nz = 0;
    int64_t nw_off = 0, nb_off = 0, t_nw_off = 0, t_nb_off = 0;
    // conditionAdv (train.lua:158-180), JOIN5 block: the context branch's operands (the block's own fields describe the prediction branch;
    // the master weight is one [2*Cs][K5] matrix, context rows first, and one bias vector of 2*Cs entries)
    Tensor in2;                             // the context image (the trainer's real_ctx)
    bf16 *col2 = nullptr;                   // window buffer of the context branch
    int64_t t_w2_off = 0, t_b2_off = 0;
    int pad5 = 0, pad5_2 = 0;
    int nbias() const { return Cout - nz; }
    Tensor in, y, a, g;                     // input, conv output (BN blocks), activation output, gradient buffer
    bf16 *col = nullptr;                    // im2col of the thin large-side tensor (forward input or backward gradient);
                                            // V4 blocks on maps larger than 4x4 (fineSize 256): the 4x4 window buffer [M][16*Clp]
    bf16 *colg = nullptr;                   // V4 blocks, general case: window buffer of the gradient [M][16*Clp] (CONV_V4 / HEAD dgrad output)
    int64_t col_rows = 0, col_k = 0;
    int P = 1;                              // V4 blocks: window positions per sample (1 at fineSize 128, 25 at fineSize 256)
    int Mrows = 0;                          // V4 blocks: GEMM rows = batch * P
    bf16 *Wt = nullptr;                     // transposed operand copy
    int cl_rows = 0;
    unsigned int *done_ctr = nullptr;       // bn_finalize_apply_act: CTAs that have consumed the statistics
    unsigned long long *bar = nullptr;      // experimental fused BN backward (CENN_BN_BWD_FUSED=1): grid-barrier counter
    float *part = nullptr;                  // per-CTA partial rows of the backward reductions [part_rows][2*Coutp]
    int part_rows = 0;
    int red_rows = 0;                       // CTA rows of the atomic-sum variant of the BN backward reduction
    float *stats = nullptr, *bsums = nullptr, *mean = nullptr, *invstd = nullptr, *scale = nullptr, *shift = nullptr, *coef = nullptr;
    float *running = nullptr;               // [2][Cout] running_mean, running_var
    int stats_cols = 0, fold = 1;
    std::vector<cudaEvent_t> ar_ev;         // data parallel: one event per chunk of this block's gradient bucket (recorded behind the chunk's all-reduce)
    bool bwd_epi = false;                   // the first pass of this block's BN backward (sum dz, sum dz (y - mean)) is taken by the epilogue of the
                                            // dgrad GEMM that produces g (the NEXT block's p_dgrad) straight into bsums: no bn_bwd_reduce launch
    float *sig = nullptr, *gpre = nullptr;  // head
    float *bias_exp = nullptr;              // G1: bias expanded over the 16 taps
    float *bias_inf = nullptr;              // inference engine: conv bias with the BN shift folded in (per GEMM column)
    TcPlan p_fwd, p_dgrad, p_wgrad;
    TcPlan p_fwd2, p_wgrad2;                // JOIN5: context branch
    bool has_dgrad = false;
};

struct Net {
    std::vector<Block> blocks;
    int64_t nparam = 0;       // master (padded) element count
    int64_t nparam_thnn = 0;  // Module:getParameters element count
    float *master = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr;
    bf16 *wbf = nullptr;
    bf16 *gradbf = nullptr;   // data parallel: bf16 copies of the big weight-gradient blocks (what crosses NVLink)
    // data parallel, sharded reduce + Adam (nhwc::shard_reduce_adam_kernel): the allocations behind master / grad / wbf / gradbf and the
    // peers' mappings of the same allocations (CUDA IPC); peer pointer of X on rank r = ipc_X[r] + (X - base_X)
    void *base_master = nullptr, *base_grad = nullptr, *base_wbf = nullptr;
    void *ipc_master[XR_MAX_WORLD] = {}, *ipc_grad[XR_MAX_WORLD] = {}, *ipc_wbf[XR_MAX_WORLD] = {}, *ipc_gradbf[XR_MAX_WORLD] = {};
    std::vector<std::pair<int64_t, int64_t>> shard_blocks;   // (offset, count) of the weight blocks whose master / m / v are sharded over the ranks
    int64_t *bias_seg = nullptr;
    int nbias_seg = 0;
    nhwc::FoldJob *fold_jobs = nullptr;   // device: one job per block (conv gradBias <- partial rows)
    std::vector<nhwc::FoldJob> fold_host;
    long long *adam_t = nullptr;   // device step counter (optimState.t)
    float *adam_step = nullptr;    // device: lr * sqrt(1-b2^t)/(1-b1^t)
    float lr = 0.f;
    Tensor input;             // fixed input buffer of the net
};

struct Op {
    std::function<int()> fn;
    float *sync_buf = nullptr;     // non-null: after fn, this buffer must be summed across ranks (DP)
    int64_t sync_count = 0;
    const char *name = "";
    int chain = 0;                 // 1: runs on the trainer's second chain stream (generator forward beside the D-real sweep)
    double flops = 0;              // algorithmic FLOPs (tensor-core ops)
    double bytes = 0;              // algorithmic bytes (bandwidth ops), 0 if not stated
};

}  // namespace

__global__ void fold_bias4_kernel(const float *__restrict__ f8, float *__restrict__ gb, int C) {
    int c = threadIdx.x;
    if (c < C && c < 4) gb[c] += f8[c] + f8[c + 4];
}
__global__ void expand_bias_kernel(const float *__restrict__ bias, float *__restrict__ out, int C, int Cp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 16 * Cp) { int c = i % Cp; out[i] = c < C ? bias[c] : 0.f; }
}
// optim.adam's bias-corrected step size, kept on the device so that the whole step is a static launch sequence
__global__ void adam_step_kernel(long long *__restrict__ t, float *__restrict__ step, float lr, float beta1, float beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long tt = *t + 1;
    *t = tt;
    double bc1 = 1.0 - pow((double)beta1, (double)tt), bc2 = 1.0 - pow((double)beta2, (double)tt);
    *step = (float)((double)lr * sqrt(bc2) / bc1);
}

struct cenn_trainer {
    cenn_state *s = nullptr;
    cenn_trainer_config cfg;
    Net G, D;
    int B = 0, F = 0, nc = 0;
    int64_t Bglobal = 0;
    std::vector<void *> allocs;
    std::vector<Op> prog;
    size_t pc = 0;
    long last_sync = -1;
    // step inputs (device, NHWC bf16)
    Tensor real_ctx, real_aux, mask;      // aux = real_center (image) or real_full (video)
    float *in_a = nullptr, *in_b = nullptr;   // fp32 NCHW staging for host-fed steps
    uint8_t *in_m = nullptr;
    float *pin_a = nullptr, *pin_b = nullptr; uint8_t *pin_m = nullptr; float *pin_loss = nullptr;
    const float *cur_a = nullptr, *cur_b = nullptr; const uint8_t *cur_m = nullptr;
    int64_t n_a = 0, n_b = 0, n_m = 0;
    double *loss_acc = nullptr;           // [8] device accumulators
    float *loss_out = nullptr;            // [8] device floats
    Tensor df_dg;                         // gradient w.r.t. D's input (G step)
    struct GraphEntry { const void *key[3]; cudaGraph_t graph; cudaGraphExec_t exec; int seen; int64_t kernels; unsigned long long last_use; };
    std::vector<GraphEntry> graphs;       // one captured step per set of input buffers (at most GRAPH_CACHE)
    unsigned long long use_clock = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    // pipelined host-fed steps: two staging sets, a copy stream, per-set events, pinned loss slots
    float *in_a2 = nullptr, *in_b2 = nullptr; uint8_t *in_m2 = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_loss[2] = {nullptr, nullptr};
    bool consumed_valid[2] = {false, false};
    float *pin_loss2 = nullptr;           // [2][8]
    long long async_issued = 0, async_read = 0;
    bool graph_failed = false;
    const void *graph_key[3] = {nullptr, nullptr, nullptr}, *seen_key[3] = {nullptr, nullptr, nullptr};
    int64_t graph_kernels = 0;
    int64_t launches_per_step = 0;
    double flops_per_step = 0;
    std::vector<cudaEvent_t> events;      // fork / join events (side stream, overlapped gradient buckets)
    cudaStream_t side3 = nullptr;         // early Adam of the big generator blocks (single GPU): overlaps the rest of the backward sweep
    std::vector<std::pair<int64_t, int64_t>> g_early;   // (offset, count) already updated by an early Adam
    cudaStream_t side2 = nullptr;         // second chain: generator forward beside the discriminator's real sweep
    int emit_chain = 0;                   // chain id given to the ops being emitted
    cudaStream_t side = nullptr;          // weight-gradient GEMMs run here, beside the dgrad / BN-backward chain of the next layer
    bool serial = false;                  // per-op profiling: everything on the main stream
    std::vector<std::pair<int64_t, int64_t>> g_buckets;   // (offset, count) of G's gradient ranges reduced on the bulk communicator
    std::vector<std::pair<int64_t, int64_t>> d_buckets;   // same for D (second sweep of the step only: the first one just accumulates)
    bool d_bucket_sweep = false;
    // clip mode (cenn_trainer_step_clips_*): cur_a == nullptr, cur_b = frames01 [B,C,F,F], cur_m = one mask plane per sample, cur_f = hflip flags
    uint8_t *in_f = nullptr, *in_f2 = nullptr;
    const uint8_t *cur_f = nullptr;
    float clip_mv = -1.f;                 // maskValue baked into the captured clip-mode graphs
    // frame mode (cenn_trainer_step_frames_host): whole decoded frames + loader draws in, crop / mask / random blocks on the device
    uint8_t *fr_u8 = nullptr, *fr_mask = nullptr; int *fr_tab = nullptr; size_t fr_u8_cap = 0, fr_mask_cap = 0;
    bool peer_ar_ok = false;              // data parallel: both nets' gradient vectors are mapped on every rank (peer all-reduce of the leftover ranges)
    bool shard_ok = false;                // data parallel: the peers' generator buffers are mapped (sharded reduce + Adam of the big blocks)
    float *noise = nullptr;               // noiseGen: this step's noise draw [B][nz] fp32 (cenn_trainer_set_noise_*)
    bool join_ctx_done = false;           // program construction: the context branch of conditionAdv's first layer has been emitted for this step
    bool infer = false;                   // inference engine (cenn_inpainter_*): forward plans only, BN folded into weights and bias
    int infer_n = 0;                      // tiles converted in / out by the current forward call
};

namespace {

typedef cenn_trainer T;

// data parallel: are the BN batch statistics those of the GLOBAL batch (exchanged across ranks: the step equals one executor at batchSize *
// world_size) or of this rank's own batch (cfg.bn_local: the usual distributed-data-parallel semantics -- no exchange, every rank normalises
// with its batchSize samples exactly as one reference process would; running statistics are per rank)?
static inline bool bn_sync(const T *t) { return t->cfg.world_size > 1 && !t->cfg.bn_local; }
static inline double bn_batch(const T *t) { return bn_sync(t) ? (double)t->Bglobal : (double)t->B; }

template <typename X>
X *dalloc(T *t, int64_t n, bool zero = true) {
    void *p = nullptr;
    size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(X);
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cenn_set_error("trainer: device allocation of %zu bytes failed", bytes); return nullptr; }
    if (zero) cudaMemset(p, 0, bytes);
    t->allocs.push_back(p);
    return reinterpret_cast<X *>(p);
}
int alloc_tensor(T *t, Tensor &x, int N, int H, int W, int C, int Cp) {
    x.N = N; x.H = H; x.W = W; x.C = C; x.Cp = Cp;
    x.p = dalloc<bf16>(t, x.elems());
    return x.p ? 0 : 1;
}
int pad_thin(int C) { return C <= 4 ? 4 : (C <= 16 ? 16 : round_up(C, 64)); }
int pad_wide(int C) { return C % 64 == 0 ? C : (C < 64 ? 64 : round_up(C, 8)); }

// ---- network description -------------------------------------------------------------------------------------
struct Spec { BlockType type; int Cs, Cl; bool bn; int act; int nz = 0; };

std::vector<Spec> spec_G(const cenn_trainer_config &c) {
    int nc = c.variant == 1 ? c.nc * c.predLen : c.nc, nef = c.nef, ngf = c.ngf, nB = c.nBottleneck;
    std::vector<Spec> v;
    v.push_back({CONV_S2, nef, nc, false, nhwc::ACT_LEAKY});
    v.push_back({CONV_S2, nef, nef, true, nhwc::ACT_LEAKY});
    v.push_back({CONV_S2, nef * 2, nef, true, nhwc::ACT_LEAKY});
    v.push_back({CONV_S2, nef * 4, nef * 2, true, nhwc::ACT_LEAKY});
    v.push_back({CONV_S2, nef * 8, nef * 4, true, nhwc::ACT_LEAKY});
    const int nz = (c.variant == 0 && c.noiseGen) ? c.nz : 0;         // train.lua:109-124: [encoder | 1x1 conv of the noise] -> BN(nBottleneck + nz)
    v.push_back({CONV_V4, nB, nef * 8, true, nhwc::ACT_LEAKY, nz});   // E6 + netG's BN(nz_size) + LeakyReLU
    v.push_back({FULL_V4, nB + nz, ngf * 8, true, nhwc::ACT_RELU});   // G1: Cs = nz_size (input), Cl = ngf*8
    v.push_back({FULL_S2, ngf * 8, ngf * 4, true, nhwc::ACT_RELU});
    v.push_back({FULL_S2, ngf * 4, ngf * 2, true, nhwc::ACT_RELU});
    v.push_back({FULL_S2, ngf * 2, ngf, true, nhwc::ACT_RELU});
    if (c.variant == 1) v.push_back({FULL_S2, ngf, ngf, true, nhwc::ACT_RELU});
    v.push_back({FULL_S2, ngf, nc, false, nhwc::ACT_TANH});
    return v;
}
std::vector<Spec> spec_D(const cenn_trainer_config &c) {
    int nc = c.variant == 1 ? c.nc * c.predLen : c.nc, ndf = c.ndf;
    std::vector<Spec> v;
    if (c.variant == 1) {
        v.push_back({CONV_S2, ndf / 2, nc, false, nhwc::ACT_LEAKY});
        v.push_back({CONV_S2, ndf, ndf / 2, false, nhwc::ACT_LEAKY});
    } else if (c.conditionAdv) {                                       // train.lua:158-180: D also sees the context
        v.push_back({JOIN5, ndf, nc, false, nhwc::ACT_LEAKY});        // two 5x5 / stride-2 convs -> JoinTable(2) -> LeakyReLU: 2*ndf x 64 x 64
        v.push_back({CONV_S2, ndf, ndf * 2, true, nhwc::ACT_LEAKY});
    } else {
        v.push_back({CONV_S2, ndf, nc, false, nhwc::ACT_LEAKY});
    }
    v.push_back({CONV_S2, ndf * 2, ndf, true, nhwc::ACT_LEAKY});
    v.push_back({CONV_S2, ndf * 4, ndf * 2, true, nhwc::ACT_LEAKY});
    v.push_back({CONV_S2, ndf * 8, ndf * 4, true, nhwc::ACT_LEAKY});
    v.push_back({HEAD, 1, ndf * 8, false, nhwc::ACT_SIGMOID});
    return v;
}

int grid1d(const cenn_state *s, int64_t items, int threads = 256, int per_sm = 8) { return bw_grid(s, items, threads, per_sm); }

// ---- building a net: shapes, buffers, parameter layout, plans ---------------------------------------------------
int build_net(T *t, Net &net, const std::vector<Spec> &specs, int in_size, int in_C, bool first_dgrad, bool single_pass) {
    cenn_state *s = t->s;
    const int N = t->B;
    // input tensor (thin: 3 or 12 channels)
    if (alloc_tensor(t, net.input, N, in_size, in_size, in_C, pad_thin(in_C))) return 1;
    Tensor cur = net.input;
    int64_t off = 0, toff = 0;
    net.blocks.resize(specs.size());
    std::vector<int64_t> bias_segs;
    for (size_t i = 0; i < specs.size(); ++i) {
        Block &b = net.blocks[i];
        const Spec &sp = specs[i];
        b.type = sp.type; b.Cs = sp.Cs; b.Cl = sp.Cl; b.bn = sp.bn; b.act = sp.act;
        b.in = cur;
        int oh, ow;
        bool out_small;
        switch (sp.type) {
            case CONV_S2: REQUIRE(cur.H % 2 == 0 && cur.H >= 2, "conv input size %d not even", cur.H); b.h = cur.H / 2; b.w = cur.W / 2; oh = b.h; ow = b.w; out_small = true; break;
            case CONV_V4: case HEAD: REQUIRE(cur.H >= 4 && cur.W >= 4 && cur.H <= 16, "4x4 valid conv expects a 4x4 .. 16x16 input, got %dx%d", cur.H, cur.W);
                b.h = cur.H - 3; b.w = cur.W - 3; oh = b.h; ow = b.w; out_small = true; break;
            case JOIN5: REQUIRE(i == 0 && t->real_ctx.p && t->real_ctx.H == 2 * cur.H && cur.Cp == 4 && t->real_ctx.Cp == 4 && sp.Cs % 64 == 0,
                                "conditionAdv: the joined first layer needs a 4-channel-padded prediction of half the context size and ndf %% 64 == 0");
                b.h = cur.H; b.w = cur.W; oh = b.h; ow = b.w; out_small = false; b.in2 = t->real_ctx; b.pad5 = 2 + cur.H / 2; b.pad5_2 = 2; break;
            case FULL_V4: REQUIRE(cur.H >= 1 && cur.H <= 13, "G1 expects a 1x1 .. 13x13 input, got %dx%d", cur.H, cur.W); b.h = cur.H; b.w = cur.W; oh = cur.H + 3; ow = cur.W + 3; out_small = false; break;
            default: b.h = cur.H; b.w = cur.W; oh = 2 * cur.H; ow = 2 * cur.W; out_small = false; break;
        }
        // channel padding: the large side of an s2 layer is either thin (im2col) or a multiple of 64 (TMA gather)
        if (out_small || sp.type == JOIN5) { b.Clp = cur.Cp; REQUIRE(cur.C == sp.Cl, "channel mismatch at block %zu", i); }
        else { b.Csp = cur.Cp; REQUIRE(cur.C == sp.Cs, "channel mismatch at block %zu", i); }
        bool next_needs_wide = false;   // does the next block gather this output as its large side?
        if (i + 1 < specs.size() && specs[i + 1].type == CONV_S2) next_needs_wide = true;
        if (sp.type == JOIN5) {
            b.Cout = 2 * sp.Cs; b.Csp = b.Coutp = 2 * sp.Cs;
        } else if (out_small) {
            if (sp.nz) REQUIRE(sp.type == CONV_V4 && cur.H == 4 && sp.Cs % 8 == 0, "noiseGen: the noise branch joins a 1x1 bottleneck (fineSize 128) with nBottleneck %% 8 == 0");
            b.nz = sp.nz;
            b.Cout = sp.Cs + sp.nz;
            b.Csp = sp.type == HEAD ? 1 : (next_needs_wide ? pad_wide(b.Cout) : round_up(b.Cout, 8));
            if (sp.type == CONV_S2 && b.Csp % 64 != 0 && b.Csp > 16 && next_needs_wide) b.Csp = round_up(b.Csp, 64);
            b.Coutp = b.Csp;
        } else {
            b.Cout = sp.Cl;
            b.Clp = sp.Cl <= 16 ? pad_thin(sp.Cl) : round_up(sp.Cl, 64);
            b.Coutp = b.Clp;
        }
        if (sp.type == CONV_V4 || sp.type == FULL_V4 || sp.type == HEAD) { b.P = b.h * b.w; b.Mrows = N * b.P; }
        b.thin = (sp.type == CONV_S2 || sp.type == FULL_S2) && b.Clp < 64;
        if (sp.type == FULL_S2 || sp.type == CONV_S2) REQUIRE(b.thin || b.Clp % 64 == 0, "block %zu: large-side channels %d not a multiple of 64", i, b.Clp);
        if (sp.type == FULL_S2) REQUIRE(b.Csp % 64 == 0, "block %zu: small-side channels %d not a multiple of 64", i, b.Csp);
        // parameters: weight master [Cs][16][Clp] then bias[Cout]; BN gamma, beta
        // (G1's weight rows are padded to the input pitch Csp: rows >= Cs stay zero -- the K extent of its GEMMs is the pitch)
        b.w_off = off;
        b.w_count = sp.type == JOIN5 ? (int64_t)2 * sp.Cs * nhwc::K5 : (int64_t)(sp.type == FULL_V4 ? b.Csp : sp.Cs) * 16 * b.Clp;
        off += b.w_count; off = (off + 7) & ~int64_t(7);
        b.b_off = off; off += b.nbias(); off = (off + 7) & ~int64_t(7);
        bias_segs.push_back(b.b_off); bias_segs.push_back(b.nbias());
        if (sp.type == JOIN5) {                 // ParallelTable{context conv, prediction conv}: [w, b] of the context branch first
            b.t_w2_off = toff; toff += (int64_t)sp.Cs * sp.Cl * 25;
            b.t_b2_off = toff; toff += sp.Cs;
            b.t_w_off = toff; toff += (int64_t)sp.Cs * sp.Cl * 25;
            b.t_b_off = toff; toff += sp.Cs;
        } else {
        b.t_w_off = toff; toff += (int64_t)sp.Cs * sp.Cl * 16;
        b.t_b_off = toff; toff += b.nbias();
        }
        if (b.nz) {                             // netG_noise's 1x1 conv follows netE in Module:getParameters order, before the joined BN
            b.nw_off = off; off += (int64_t)b.nz * b.nz; off = (off + 7) & ~int64_t(7);
            b.nb_off = off; off += b.nz; off = (off + 7) & ~int64_t(7);
            bias_segs.push_back(b.nb_off); bias_segs.push_back(b.nz);
            b.t_nw_off = toff; toff += (int64_t)b.nz * b.nz;
            b.t_nb_off = toff; toff += b.nz;
        }
        if (sp.bn) {
            b.g_off = off; off += b.Cout; off = (off + 7) & ~int64_t(7);
            b.be_off = off; off += b.Cout; off = (off + 7) & ~int64_t(7);
            b.t_g_off = toff; toff += b.Cout;
            b.t_be_off = toff; toff += b.Cout;
        }
        // activations
        if (sp.type == HEAD) {
            b.sig = dalloc<float>(t, b.Mrows); b.gpre = dalloc<float>(t, b.Mrows);
            if (!b.sig || !b.gpre) return 1;
        } else {
            if (sp.bn && !t->infer && alloc_tensor(t, b.y, N, oh, ow, b.Cout, b.Coutp)) return 1;
            if (alloc_tensor(t, b.a, N, oh, ow, b.Cout, b.Coutp)) return 1;
            if (!t->infer && alloc_tensor(t, b.g, N, oh, ow, b.Cout, b.Coutp)) return 1;
        }
        if (sp.type != HEAD && !t->infer) {
            // backward reductions: one partial row per CTA of the (pixel-strip x vector-group) grid
            const int vpp = b.Coutp >= 8 ? b.Coutp / 8 : 1;
            int tx = 1; while (tx < vpp && tx < 64) tx *= 2;
            const int ty = 256 / tx, gy = (vpp + tx - 1) / tx;
            const int64_t npix = b.Coutp >= 8 ? (int64_t)N * oh * ow : (int64_t)N * oh * ow / 2;
            b.part_rows = (int)std::max<int64_t>(1, std::min<int64_t>((npix + ty * 4 - 1) / (ty * 4), (int64_t)s->sm_count * 4 / gy));
            // EXPERIMENT (CENN_BN_BWD_FUSED=1, off by default): a BN layer's backward as ONE cooperatively launched kernel with two grid
            // barriers (nhwc::bn_bwd_fused_kernel); the grid (part_rows x gy CTAs) must be co-resident.  Measured on B200 (round 2):
            // 44 us per layer against 46 us for the three launches when run alone, and SLOWER inside the step (3.36 vs 3.18 ms)
            // because a cooperative grid waits until the side streams' GEMM CTAs have left the SMs.  Kept for reference only.
            {
                const char *e = getenv("CENN_BN_BWD_FUSED"), *m = getenv("CENN_BN_FUSED_MAX_ELEMS");
                const int64_t max_elems = m ? atoll(m) : (int64_t)1 << 40;
                if (sp.bn && e && atoi(e) != 0 && t->cfg.world_size <= 1 && b.Coutp >= 8 && (int64_t)N * oh * ow * b.Coutp <= max_elems) {
                    int per_sm_l = 0, per_sm_r = 0;
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_l, nhwc::bn_bwd_fused_kernel<nhwc::ACT_LEAKY>, 256, 2 * tx * 8 * sizeof(float)));
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_r, nhwc::bn_bwd_fused_kernel<nhwc::ACT_RELU>, 256, 2 * tx * 8 * sizeof(float)));
                    const int per_sm = std::max(1, std::min(per_sm_l, per_sm_r));
                    b.part_rows = (int)std::max<int64_t>(1, std::min<int64_t>(b.part_rows, (int64_t)s->sm_count * per_sm / gy));
                    b.bar = dalloc<unsigned long long>(t, 1);
                    if (!b.bar) return 1;
                }
            }
            { const char *e = getenv("CENN_BN_RED_ROWS"); const int cap = e ? atoi(e) : 2 * s->sm_count;     // fewer CTAs add into the same [2][C] row
              b.red_rows = std::max(1, std::min(b.part_rows, std::max(1, cap / gy))); }
            b.part = dalloc<float>(t, (int64_t)b.part_rows * 2 * std::max(b.Coutp, 8));
            if (!b.part) return 1;
        }
        if (b.P > 1) {
            REQUIRE(!t->infer, "the inference engine supports fineSize 128 only");
            b.col_rows = b.Mrows; b.col_k = 16 * b.Clp;
            b.col = dalloc<bf16>(t, b.col_rows * b.col_k);
            if (!b.col) return 1;
            if (sp.type != FULL_V4) { b.colg = dalloc<bf16>(t, b.col_rows * b.col_k); if (!b.colg) return 1; }
        }
        if (sp.type == JOIN5) {
            REQUIRE(!t->infer, "the inference engine builds generators only");
            b.col_rows = (int64_t)N * oh * ow; b.col_k = nhwc::K5;
            b.col = dalloc<bf16>(t, b.col_rows * b.col_k); b.col2 = dalloc<bf16>(t, b.col_rows * b.col_k); b.colg = dalloc<bf16>(t, b.col_rows * b.col_k);
            if (!b.col || !b.col2 || !b.colg) return 1;
        }
        if (b.thin && !(t->infer && sp.type != CONV_S2)) {
            b.col_rows = (int64_t)N * b.h * b.w; b.col_k = 16 * b.Clp;
            b.col = dalloc<bf16>(t, b.col_rows * b.col_k);
            if (!b.col) return 1;
        }
        if (t->infer) {
            b.bias_inf = dalloc<float>(t, sp.type == FULL_V4 ? 16 * (int64_t)b.Clp : b.Coutp);
            if (!b.bias_inf) return 1;
            if (sp.bn) {
                b.running = dalloc<float>(t, 2 * (int64_t)b.Coutp);
                if (!b.running) return 1;
                std::vector<float> ones(b.Coutp, 1.f);
                CK(cudaMemcpy(b.running + b.Coutp, ones.data(), b.Coutp * sizeof(float), cudaMemcpyHostToDevice));
            }
        } else if (sp.bn) {
            b.fold = (sp.type == FULL_V4 && b.P == 1) ? 16 : 1;       // G1 on a 1x1 input: the GEMM epilogue accumulates per (tap, channel) column
            b.stats_cols = b.Coutp * b.fold;
            b.stats = dalloc<float>(t, 2 * (int64_t)b.stats_cols);
            b.bsums = dalloc<float>(t, 2 * (int64_t)b.Coutp);
            b.mean = dalloc<float>(t, b.Coutp); b.invstd = dalloc<float>(t, b.Coutp);
            b.scale = dalloc<float>(t, b.Coutp); b.shift = dalloc<float>(t, b.Coutp);
            b.coef = dalloc<float>(t, 3 * (int64_t)b.Coutp);
            b.running = dalloc<float>(t, 2 * (int64_t)b.Coutp);
            b.done_ctr = dalloc<unsigned int>(t, 1);
            if (!b.stats || !b.bsums || !b.mean || !b.invstd || !b.scale || !b.shift || !b.coef || !b.running) return 1;
            std::vector<float> ones(b.Coutp, 1.f);
            CK(cudaMemcpy(b.running + b.Coutp, ones.data(), b.Coutp * sizeof(float), cudaMemcpyHostToDevice));
        }
        cur = sp.type == HEAD ? Tensor() : b.a;
    }
    net.nparam = off;
    net.nparam_thnn = toff;
    // The Adam kernel streams these five arrays at identical offsets; with identically aligned bases all five streams
    // would walk the same HBM channel sequence in lock step (observed: 3x slower).  Skew the bases by odd multiples of 256 B.
    auto skewed = [&](int k, void **base) -> float * { float *p = dalloc<float>(t, off + 8 * 65536); if (base) *base = p; return p ? p + (size_t)k * (33856 + 64) : nullptr; };
    net.master = skewed(0, &net.base_master);
    if (!t->infer) { net.grad = skewed(1, &net.base_grad); net.m = skewed(2, nullptr); net.v = skewed(3, nullptr); }
    { bf16 *p = dalloc<bf16>(t, off + 8 * 65536); net.base_wbf = p; net.wbf = p ? p + (size_t)4 * (33856 + 64) * 2 : nullptr; }
    net.adam_t = dalloc<long long>(t, 1); net.adam_step = dalloc<float>(t, 1);
    if (!net.master || !net.wbf || !net.adam_t || !net.adam_step) return 1;
    if (t->infer) {
        // forward plans only: bias (with the BN shift folded in) and the activation in the GEMM epilogue, straight into `a`
        for (size_t i = 0; i < net.blocks.size(); ++i) {
            Block &b = net.blocks[i];
            const bf16 *Wf = net.wbf + b.w_off;
            const int M = N * b.h * b.w;
            TcEpilogue ep; ep.bias = b.bias_inf; ep.act = b.act; ep.act_param = 0.2f;
            TcEpilogue ep_m = ep; ep_m.b_mn = true;
            switch (b.type) {
                case CONV_S2:
                    if (b.thin) { if (tc_plan_gemm(s, &b.p_fwd, b.col, Wf, b.a.p, M, b.Cs, (int)b.col_k, b.Csp, ep)) return 1; }
                    else if (tc_plan_fprop_s2(s, &b.p_fwd, b.in.p, Wf, b.a.p, N, b.h, b.w, b.Cs, b.Csp, b.Clp, ep)) return 1;
                    break;
                case CONV_V4:
                    if (tc_plan_gemm(s, &b.p_fwd, b.in.p, Wf, b.a.p, N, b.Cs, 16 * b.Clp, b.Csp, ep)) return 1;
                    break;
                case FULL_V4:
                    if (tc_plan_gemm(s, &b.p_fwd, b.in.p, Wf, b.a.p, N, 16 * b.Clp, b.Csp, 16 * b.Clp, ep_m)) return 1;
                    break;
                case FULL_S2:
                    b.cl_rows = b.Clp;
                    if (b.Clp % 64 == 0) {
                        if (tc_plan_dgrad_s2(s, &b.p_fwd, b.in.p, Wf, b.a.p, N, b.h, b.w, b.Csp, b.Clp, b.Clp, b.cl_rows, ep_m)) return 1;
                    } else {
                        b.Wt = dalloc<bf16>(t, (int64_t)16 * b.cl_rows * b.Csp);
                        if (!b.Wt) return 1;
                        if (tc_plan_dgrad_s2(s, &b.p_fwd, b.in.p, b.Wt, b.a.p, N, b.h, b.w, b.Csp, b.Clp, b.Clp, b.cl_rows, ep)) return 1;
                    }
                    break;
                case HEAD: case JOIN5: REQUIRE(false, "the inference engine builds generators only");
            }
        }
        return 0;
    }
    if (!net.grad || !net.m || !net.v) return 1;
    if (t->cfg.world_size > 1 && s->comm2 && single_pass && getenv("CENN_FP32_BUCKETS") == nullptr) {
        net.gradbf = dalloc<bf16>(t, off);
        if (!net.gradbf) return 1;
    }
    net.nbias_seg = (int)bias_segs.size() / 2;
    net.bias_seg = dalloc<int64_t>(t, bias_segs.size());
    if (!net.bias_seg) return 1;
    CK(cudaMemcpy(net.bias_seg, bias_segs.data(), bias_segs.size() * sizeof(int64_t), cudaMemcpyHostToDevice));

    for (Block &b : net.blocks) {
        if (b.type == HEAD) continue;
        nhwc::FoldJob j;
        j.src = b.part; j.dst = net.grad + b.b_off; j.rows = b.part_rows; j.C = b.nbias();
        if (b.Coutp >= 8) { j.stride = b.Coutp; j.fold = 1; j.fold_stride = 0; } else { j.stride = 8; j.fold = 2; j.fold_stride = 4; }
        net.fold_host.push_back(j);
        if (b.nz) { j.src = b.part + b.nbias(); j.dst = net.grad + b.nb_off; j.C = b.nz; net.fold_host.push_back(j); }   // the noise conv's gradBias
    }
    net.fold_jobs = dalloc<nhwc::FoldJob>(t, net.fold_host.size());
    if (!net.fold_jobs) return 1;
    CK(cudaMemcpy(net.fold_jobs, net.fold_host.data(), net.fold_host.size() * sizeof(nhwc::FoldJob), cudaMemcpyHostToDevice));

    const int acc_mode = single_pass ? 2 : 1;   // gradients of a net with one backward pass per step may be stored instead of accumulated
    // operand copies + plans
    for (size_t i = 0; i < net.blocks.size(); ++i) {
        Block &b = net.blocks[i];
        const bf16 *Wf = net.wbf + b.w_off;
        float *gW = net.grad + b.w_off;
        Block *prev = i > 0 ? &net.blocks[i - 1] : nullptr;
        b.has_dgrad = prev != nullptr || first_dgrad;
        bf16 *dgrad_out = prev ? prev->g.p : nullptr;
        const int M = N * b.h * b.w;
        TcEpilogue ep_f;                       // forward epilogue: bias, then BN statistics or fused activation
        // (conv biases are zero in every training forward -- train.lua:279-280 -- but not when a checkpoint is evaluated)
        ep_f.bias = net.master + b.b_off;
        if (b.type == FULL_V4 && b.P == 1) {   // G1 on a 1x1 input: GEMM columns are (tap, channel) -> per-column copy of the channel bias
            b.bias_exp = dalloc<float>(t, 16 * (int64_t)b.Clp);
            if (!b.bias_exp) return 1;
            ep_f.bias = b.bias_exp;
        }
        if (b.bn) { ep_f.stats = b.stats; ep_f.stats_stride = b.stats_cols; }
        else { ep_f.act = b.act; ep_f.act_param = 0.2f; }
        bf16 *fwd_out = b.bn ? b.y.p : b.a.p;
        TcEpilogue ep_n;
        // EXPERIMENT (CENN_BWD_EPI=1, off by default): BN-backward sums of the previous block from this block's dgrad epilogue -- the dgrad output
        // must be the previous block's gradient tensor itself, with GEMM columns = channels (not the window buffers / (tap, channel) columns of
        // the 4x4 valid layers).  Correct (tests pass with it on) but much SLOWER on B200 (round 2: image step 4.24 vs 2.98 ms, video 3.42 vs
        // 2.77 ms): the epilogue's four warps read y with a few dozen loads in flight where the stand-alone reduction keeps thousands, and a
        // CTA has only one or two tiles to hide that latency behind; staging y by TMA would need the shared memory the operand ring uses.
        if (prev && prev->bn && dgrad_out && b.has_dgrad && getenv("CENN_BWD_EPI") != nullptr && getenv("CENN_BN_BWD_3LAUNCH") == nullptr &&
            getenv("CENN_BN_BWD_2LAUNCH") == nullptr && !prev->bar && prev->Coutp % 64 == 0 &&
            ((b.type == CONV_S2 && !b.thin) || b.type == FULL_S2 || (b.type == FULL_V4 && b.P == 1)) &&
            (!bn_sync(t) || (s->xr_enabled && 2 * prev->Coutp <= XR_MAXF && getenv("CENN_DP_BN_FOLD") == nullptr))) {
            ep_n.stats = prev->bsums; ep_n.stats_stride = prev->Coutp;
            ep_n.bwd_y = prev->y.p; ep_n.bwd_scale = prev->scale; ep_n.bwd_shift = prev->shift; ep_n.bwd_mean = prev->mean;
            ep_n.bwd_act = prev->act; ep_n.bwd_negval = 0.2f;
            prev->bwd_epi = true;
        }
        switch (b.type) {
            case CONV_S2: {
                if (b.thin) { if (tc_plan_gemm(s, &b.p_fwd, b.col, Wf, fwd_out, M, b.Cs, (int)b.col_k, b.Csp, ep_f)) return 1; }
                else if (tc_plan_fprop_s2(s, &b.p_fwd, b.in.p, Wf, fwd_out, N, b.h, b.w, b.Cs, b.Csp, b.Clp, ep_f)) return 1;
                if (b.thin) { if (tc_plan_wgrad_plain(s, &b.p_wgrad, b.g.p, b.col, gW, M, b.Cs, b.Csp, (int)b.col_k, 1.f, 1)) return 1; }
                else if (tc_plan_wgrad_s2(s, &b.p_wgrad, b.g.p, b.in.p, gW, N, b.h, b.w, b.Cs, b.Csp, b.Clp, 1.f, 1)) return 1;
                if (b.has_dgrad) {
                    REQUIRE(b.Csp % 64 == 0, "block %zu: dgrad needs small-side channels padded to 64 (got %d)", i, b.Csp);
                    b.cl_rows = b.Clp;
                    if (!dgrad_out) { dgrad_out = dalloc<bf16>(t, b.in.elems()); if (!dgrad_out) return 1; if (&net == &t->D) { t->df_dg = b.in; t->df_dg.p = dgrad_out; } }
                    if (b.Clp % 64 == 0) {           // weights read MN-major straight from the master copy: no transposed operand
                        TcEpilogue ep_m = ep_n; ep_m.b_mn = true;
                        if (tc_plan_dgrad_s2(s, &b.p_dgrad, b.g.p, Wf, dgrad_out, N, b.h, b.w, b.Csp, b.Clp, b.Clp, b.cl_rows, ep_m)) return 1;
                    } else {
                        b.Wt = dalloc<bf16>(t, (int64_t)16 * b.cl_rows * b.Csp);
                        if (!b.Wt) return 1;
                        if (tc_plan_dgrad_s2(s, &b.p_dgrad, b.g.p, b.Wt, dgrad_out, N, b.h, b.w, b.Csp, b.Clp, b.Clp, b.cl_rows, ep_n)) return 1;
                    }
                }
                break;
            }
            case CONV_V4: {
                // P == 1 (fineSize 128): the 4x4 window is the whole map, the input IS the GEMM operand and the dgrad output IS the previous
                // block's gradient.  P > 1: both go through the window buffers (im2col_v4 / col2im_v4 around the GEMMs).
                int K = 16 * b.Clp;
                const int Mr = b.Mrows;
                const bf16 *a_fwd = b.P > 1 ? b.col : b.in.p;
                if (tc_plan_gemm(s, &b.p_fwd, a_fwd, Wf, fwd_out, Mr, b.Cs, K, b.Csp, ep_f)) return 1;
                if (tc_plan_wgrad_plain(s, &b.p_wgrad, b.g.p, a_fwd, gW, Mr, b.Cs, b.Csp, K, 1.f, acc_mode)) return 1;
                { TcEpilogue ep_m = ep_n; ep_m.b_mn = true;     // dgrad = g [M,Cs] x Wf [Cs][K]
                  if (tc_plan_gemm(s, &b.p_dgrad, b.g.p, Wf, b.P > 1 ? b.colg : dgrad_out, Mr, K, b.nz ? b.Cs : b.Csp, K, ep_m, b.Csp)) return 1; }
                break;
            }
            case FULL_V4: {
                int K = 16 * b.Clp;
                const int Mr = b.Mrows;
                { TcEpilogue ep_m = ep_f; ep_m.b_mn = true;     // G1 = in [M,Cs] x Wf [Cs][K]
                  if (b.P > 1) { ep_m.bias = nullptr; ep_m.stats = nullptr; ep_m.act = 0; }      // bias, statistics (and no activation: BN follows) after the overlap-add
                  if (tc_plan_gemm(s, &b.p_fwd, b.in.p, Wf, b.P > 1 ? b.col : fwd_out, Mr, K, b.Csp, K, ep_m)) return 1; }
                const bf16 *g_cols = b.P > 1 ? b.col : b.g.p;    // backward: the window buffer is reused for im2col_v4(g_y)
                if (tc_plan_wgrad_plain(s, &b.p_wgrad, b.in.p, g_cols, gW, Mr, b.Cs, b.Csp, K, 1.f, acc_mode)) return 1;
                if (tc_plan_gemm(s, &b.p_dgrad, g_cols, Wf, dgrad_out, Mr, b.Cs, K, b.Csp, ep_n)) return 1;
                break;
            }
            case FULL_S2: {
                b.cl_rows = b.Clp;
                if (b.Clp % 64 == 0) {
                    TcEpilogue ep_m = ep_f; ep_m.b_mn = true;
                    if (tc_plan_dgrad_s2(s, &b.p_fwd, b.in.p, Wf, fwd_out, N, b.h, b.w, b.Csp, b.Clp, b.Clp, b.cl_rows, ep_m)) return 1;
                } else {
                    b.Wt = dalloc<bf16>(t, (int64_t)16 * b.cl_rows * b.Csp);
                    if (!b.Wt) return 1;
                    if (tc_plan_dgrad_s2(s, &b.p_fwd, b.in.p, b.Wt, fwd_out, N, b.h, b.w, b.Csp, b.Clp, b.Clp, b.cl_rows, ep_f)) return 1;
                }
                if (b.thin) {
                    if (tc_plan_wgrad_plain(s, &b.p_wgrad, b.in.p, b.col, gW, M, b.Cs, b.Csp, (int)b.col_k, 1.f, 1)) return 1;
                    if (tc_plan_gemm(s, &b.p_dgrad, b.col, Wf, dgrad_out, M, b.Cs, (int)b.col_k, b.Csp, ep_n)) return 1;
                } else {
                    if (tc_plan_wgrad_s2(s, &b.p_wgrad, b.in.p, b.g.p, gW, N, b.h, b.w, b.Cs, b.Csp, b.Clp, 1.f, 1)) return 1;
                    if (tc_plan_fprop_s2(s, &b.p_dgrad, b.g.p, Wf, dgrad_out, N, b.h, b.w, b.Cs, b.Csp, b.Clp, ep_n)) return 1;
                }
                break;
            }
            case JOIN5: {
                // forward: a[:, 0:Cs] = leaky(col2 x Wctx^T + b), a[:, Cs:2Cs] = leaky(col x Wpred^T + b); weights K-major [Cs][K5] per branch
                const int K = nhwc::K5, Cs = b.Cs;
                TcEpilogue ep2 = ep_f, ep1 = ep_f;
                ep1.bias = net.master + b.b_off + Cs;
                if (tc_plan_gemm(s, &b.p_fwd2, b.col2, Wf, fwd_out, M, Cs, K, 2 * Cs, ep2)) return 1;
                if (tc_plan_gemm(s, &b.p_fwd, b.col, Wf + (int64_t)Cs * K, fwd_out + Cs, M, Cs, K, 2 * Cs, ep1)) return 1;
                if (tc_plan_wgrad_plain(s, &b.p_wgrad2, b.g.p, b.col2, gW, M, Cs, 2 * Cs, K, 1.f, 1)) return 1;
                if (tc_plan_wgrad_plain(s, &b.p_wgrad, b.g.p + Cs, b.col, gW + (int64_t)Cs * K, M, Cs, 2 * Cs, K, 1.f, 1)) return 1;
                // gradient w.r.t. the prediction (fGx, train.lua:369-371 takes df_dg[2]): window gradients = g[:, Cs:2Cs] x Wpred, then the overlap-add
                { TcEpilogue ep_m = ep_n; ep_m.b_mn = true;
                  if (tc_plan_gemm(s, &b.p_dgrad, b.g.p + Cs, Wf + (int64_t)Cs * K, b.colg, M, K, Cs, K, ep_m, 2 * Cs)) return 1; }
                if (!dgrad_out) { dgrad_out = dalloc<bf16>(t, b.in.elems()); if (!dgrad_out) return 1; }
                if (&net == &t->D) { t->df_dg = b.in; t->df_dg.p = dgrad_out; }
                break;
            }
            case HEAD: break;
        }
    }
    return 0;
}

// ---- op emission helpers -----------------------------------------------------------------------------------------
#define KLAUNCH(s) do { (s)->launches++; if (cenn_check_cuda(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)) return 1; } while (0)

void emit(T *t, const char *name, std::function<int()> fn, float *sync_buf = nullptr, int64_t sync_count = 0) {
    Op op; op.fn = std::move(fn); op.sync_buf = sync_buf; op.sync_count = sync_count; op.name = name; op.chain = t->emit_chain;
    t->prog.push_back(std::move(op));
}
void emit_plan(T *t, const char *name, TcPlan *pl) {
    cenn_state *s = t->s;
    t->flops_per_step += pl->flops;
    emit(t, name, [s, pl]() { return tc_launch(s, pl); });
    t->prog.back().flops = pl->flops;
}
void emit_im2col(T *t, const Tensor &L, bf16 *col, int h, int w) {
    cenn_state *s = t->s;
    Tensor Lc = L;
    emit(t, "im2col", [s, Lc, col, h, w]() {
        int64_t total = (int64_t)Lc.N * h * w * 4;            // one thread per (pixel, window row)
        if (Lc.Cp == 4) LK(nhwc::im2col_kernel<4>, dim3(grid1d(s, total)), dim3(256), 0, s->stream)(Lc.p, col, Lc.N, h, w);
        else if (Lc.Cp == 16) LK(nhwc::im2col16_kernel, dim3(grid1d(s, total * 8)), dim3(256), 0, s->stream)(Lc.p, col, Lc.N, h, w);
        else { cenn_set_error("im2col: unsupported thin channel count %d", Lc.Cp); return 1; }
        KLAUNCH(s); return 0;
    });
}
// 5x5 / stride-2 windows of a 4-channel-padded image (conditionAdv's first layer): x -> col[(n,oy,ox)][K5]
void emit_im2col5(T *t, const Tensor &x, bf16 *col, int oh, int ow, int pad) {
    cenn_state *s = t->s;
    Tensor xc = x;
    emit(t, "im2col5", [s, xc, col, oh, ow, pad]() {
        LK(nhwc::im2col5_kernel, dim3(grid1d(s, (int64_t)xc.N * oh * ow * 16)), dim3(256), 0, s->stream)(xc.p, col, xc.N, xc.H, xc.W, oh, ow, pad);
        KLAUNCH(s); return 0; });
    t->prog.back().bytes = 2.0 * ((double)x.N * oh * ow * nhwc::K5 + (double)x.pix() * 4);
}
void emit_col2im5(T *t, const bf16 *col, const Tensor &gx, int oh, int ow, int pad) {
    cenn_state *s = t->s;
    Tensor gc = gx;
    emit(t, "col2im5", [s, col, gc, oh, ow, pad]() {
        LK(nhwc::col2im5_kernel, dim3(grid1d(s, gc.pix())), dim3(256), 0, s->stream)(col, gc.p, gc.N, gc.H, gc.W, oh, ow, pad);
        KLAUNCH(s); return 0; });
}
void reduce_dims(int vec_per_pix, dim3 &block, int &gy);
// 4x4 / stride-1 windows of an [N,H,W,Cp] map (fineSize 256, V4 blocks): x -> col[(n,oy,ox)][tap][c]
void emit_im2col_v4(T *t, const Tensor &x, bf16 *col) {
    cenn_state *s = t->s;
    Tensor xc = x;
    emit(t, "im2col_v4", [s, xc, col]() {
        const int64_t total = (int64_t)xc.N * (xc.H - 3) * (xc.W - 3) * 16 * (xc.Cp / 8);
        LK(nhwc::im2col_v4_kernel, dim3(grid1d(s, total)), dim3(256), 0, s->stream)(xc.p, col, xc.N, xc.H, xc.W, xc.Cp);
        KLAUNCH(s); return 0; });
    t->prog.back().bytes = 2.0 * 2.0 * (double)x.N * (x.H - 3) * (x.W - 3) * 16 * x.C;
}
// adjoint (overlap-add of the windows) into out [N,H,W,Cp]; optional bias and BN statistics of the stored values
void emit_col2im_v4(T *t, const bf16 *col, const Tensor &out, const float *bias, float *stats, int stats_stride) {
    cenn_state *s = t->s;
    Tensor oc = out;
    emit(t, "col2im_v4", [s, col, oc, bias, stats, stats_stride]() {
        const int vpp = oc.Cp / 8;
        dim3 blk; int gy; reduce_dims(vpp, blk, gy);
        const int64_t npix = oc.pix();
        const int rows = (int)std::max<int64_t>(1, std::min<int64_t>((npix + blk.y - 1) / blk.y, (int64_t)s->sm_count * 8 / gy));
        LK(nhwc::col2im_v4_kernel, dim3(rows, gy), blk, stats ? 2 * blk.x * 8 * sizeof(float) : 0, s->stream)(col, oc.p, bias, stats, stats_stride, oc.N, oc.H, oc.W, oc.Cp, oc.C);
        KLAUNCH(s); return 0; });
    t->prog.back().bytes = 2.0 * ((double)out.N * (out.H - 3) * (out.W - 3) * 16 * out.C + (double)out.pix() * out.C);
}
void reduce_dims(int vec_per_pix, dim3 &block, int &gy) {
    int tx = 1; while (tx < vec_per_pix && tx < 64) tx *= 2;
    block = dim3(tx, 256 / tx);
    gy = (vec_per_pix + tx - 1) / tx;
}

// forward of one block (train = batch statistics + running update; eval = running statistics)
void emit_forward(T *t, Net &net, size_t i, bool train) {
    cenn_state *s = t->s;
    Block *b = &net.blocks[i];
    float *master = net.master;
    const double n_global = b->type == HEAD ? 1.0 : bn_batch(t) * b->a.H * b->a.W;
    if (b->type == HEAD) {
        const bf16 *w = net.wbf + b->w_off; const float *bias = master + b->b_off;
        int B = b->Mrows, K = 16 * b->Clp;          // one row per 4x4 window (fineSize 128: one per sample)
        if (b->P > 1) emit_im2col_v4(t, b->in, b->col);
        const bf16 *x = b->P > 1 ? b->col : b->in.p;
        emit(t, "head_fwd", [s, x, w, bias, b, B, K]() {
            LK(nhwc::head_fwd_kernel, dim3((B * 32 + 255) / 256), dim3(256), 0, s->stream)(x, w, bias, b->sig, B, K); KLAUNCH(s); return 0; });
        return;
    }
    if (b->type == JOIN5) {
        // the context branch depends only on the context image and D's weights, both unchanged between the real and the fake sweep of fDx:
        // its window buffer and its half of the joined activation are computed by the first sweep of a step only (ctx_done)
        if (!t->join_ctx_done) {
            emit_im2col5(t, b->in2, b->col2, b->h, b->w, b->pad5_2);
            emit_plan(t, "conv_fwd", &b->p_fwd2);
            t->join_ctx_done = true;
        }
        emit_im2col5(t, b->in, b->col, b->h, b->w, b->pad5);
        emit_plan(t, "conv_fwd", &b->p_fwd);
        return;
    }
    if (b->thin && b->type == CONV_S2) emit_im2col(t, b->in, b->col, b->h, b->w);
    if (b->bias_exp) {
        const float *bias = master + b->b_off;
        emit(t, "expand_bias", [s, b, bias]() { expand_bias_kernel<<<(16 * b->Clp + 255) / 256, 256, 0, s->stream>>>(bias, b->bias_exp, b->Cout, b->Clp); KLAUNCH(s); return 0; });
    }
    if (b->type == CONV_V4 && b->P > 1) emit_im2col_v4(t, b->in, b->col);
    emit_plan(t, "conv_fwd", &b->p_fwd);
    if (b->nz) {                              // noiseGen: the 1x1 conv of the noise vector fills columns [Cs, Cs + nz) of the joined tensor (after the
        const bf16 *w = net.wbf + b->nw_off;  // encoder GEMM, whose last N tile zero-fills up to its tile edge) and adds their BN statistics
        const float *bias = master + b->nb_off;
        emit(t, "noise_fwd", [t, s, b, w, bias]() {
            LK(nhwc::noise_fwd_kernel, dim3(b->nz), dim3(128), 0, s->stream)(t->noise, w, bias, b->y.p, b->Coutp, b->Cs, t->B, b->nz, b->stats, b->stats_cols);
            KLAUNCH(s); return 0; });
    }
    if (b->type == FULL_V4 && b->P > 1)       // overlap-add of the 4x4 windows + bias + BN statistics
        emit_col2im_v4(t, b->col, b->bn ? b->y : b->a, master + b->b_off, b->bn ? b->stats : nullptr, b->stats_cols);
    if (b->bn && train && !bn_sync(t) && b->Coutp <= 4096 && getenv("CENN_NO_BN_FUSE") == nullptr) {
        float *gamma = master + b->g_off, *beta = master + b->be_off;
        emit(t, "bn_fin_apply", [s, b, gamma, beta, n_global]() {
            int64_t nvec = b->y.elems() / 8;
            LK(nhwc::bn_finalize_apply_act_kernel, dim3(grid1d(s, nvec)), dim3(256), 2 * b->Coutp * sizeof(float), s->stream)(b->stats, b->stats_cols, b->fold, b->Coutp, gamma, beta,
                b->running, b->running + b->Coutp, b->mean, b->invstd, b->scale, b->shift, b->Cout, b->Coutp, n_global, 0.1, 1e-5,
                b->y.p, b->a.p, nvec, b->Coutp / 8, b->act, 0.2f, b->done_ctr);
            KLAUNCH(s); return 0; });
        t->prog.back().bytes = 2.0 * 2.0 * (double)b->a.pix() * b->Cout;     // read y, write a
        return;
    }
    if (b->bn) {
        float *gamma = master + b->g_off, *beta = master + b->be_off;
        if (train) {
            if (bn_sync(t) && s->xr_enabled && 2 * b->Coutp <= XR_MAXF) {
                // data parallel: the statistics cross NVLink inside the finalize kernel (peer mailboxes), no collective launch
                const int chain = t->emit_chain;
                emit(t, "bn_finalize_xr", [t, s, b, gamma, beta, n_global, chain]() {
                    LK(nhwc::bn_finalize_xr_kernel, dim3(1), dim3(1024), 0, s->stream)((chain == 1 && !t->serial) ? s->xr2 : s->xr, b->stats, b->stats_cols, b->fold, b->Coutp, b->bsums, gamma, beta,
                        b->running, b->running + b->Coutp, b->mean, b->invstd, b->scale, b->shift, b->Cout, b->Coutp, n_global, 0.1, 1e-5);
                    KLAUNCH(s); return 0; });
            } else {
            emit(t, "bn_stats_sync", []() { return 0; }, bn_sync(t) ? b->stats : nullptr, 2 * (int64_t)b->stats_cols);
            emit(t, "bn_finalize", [s, b, gamma, beta, n_global]() {
                LK(nhwc::bn_finalize_kernel, dim3((b->Cout + 127) / 128), dim3(128), 0, s->stream)(b->stats, b->stats_cols, b->fold, b->Coutp, gamma, beta,
                    b->running, b->running + b->Coutp, b->mean, b->invstd, b->scale, b->shift, b->Cout, n_global, 0.1, 1e-5, 1);
                KLAUNCH(s); return 0; });
            }
        } else {
            emit(t, "bn_eval_coef", [s, b, gamma, beta]() {
                nhwc::bn_eval_coef_kernel<<<(b->Cout + 127) / 128, 128, 0, s->stream>>>(gamma, beta, b->running, b->running + b->Coutp, b->scale, b->shift, b->Cout, 1e-5);
                KLAUNCH(s);
                // eval mode must not leave batch sums behind
                return cenn_check_cuda(cudaMemsetAsync(b->stats, 0, 2 * (size_t)b->stats_cols * sizeof(float), s->stream), "memset", __FILE__, __LINE__); });
        }
        emit(t, "bn_apply_act", [s, b]() {
            int64_t nvec = b->y.elems() / 8;
            LK(nhwc::bn_apply_act_kernel, dim3(grid1d(s, nvec)), dim3(256), 0, s->stream)(b->y.p, b->a.p, b->scale, b->shift, nvec, b->Coutp / 8, b->act, 0.2f);
            KLAUNCH(s); return 0; });
        t->prog.back().bytes = 2.0 * 2.0 * (double)b->a.pix() * b->Cout;
    }
}

// smallest weight block (elements) that gets its own gradient bucket / early Adam (generator) or bucket (discriminator, second sweep)
// The generator's first-layer gradInput is dead (the reference computes and discards it).  As the LAST kernel of the step's side stream it
// was exposed at the step's tail (0.13 ms of a 2.98 ms step); nothing observable depends on it, so the program computes it at the START of
// the following step instead -- from the previous step's E1 gradient, beside the low-occupancy opening of the step.  Same FLOPs per step,
// one step late, result discarded either way.  CENN_DEAD_DGRAD_INLINE=1 restores the in-place position.
static bool defer_dead_dgrad() { static const bool v = getenv("CENN_DEAD_DGRAD_INLINE") == nullptr; return v; }
// sharded reduce + Adam (data parallel): blocks of at least 2^23 elements whose count is a multiple of 8; rank r owns [shard_begin(r), shard_begin(r + 1))
static bool shard_block(const T *t, const Net &net, int64_t cnt) { return t->shard_ok && &net == &t->G && cnt >= ((int64_t)1 << 23) && cnt % 8 == 0; }
static int64_t shard_begin(int64_t cnt, int world, int r) { const int64_t per = (((cnt + world - 1) / world) + 7) & ~int64_t(7); return std::min<int64_t>(cnt, per * r); }
template <typename X> static X *peer_ptr(void *const *ipc, int r, const void *base, const X *local) {
    return reinterpret_cast<X *>(reinterpret_cast<char *>(ipc[r]) + (reinterpret_cast<const char *>(local) - reinterpret_cast<const char *>(base)));
}
static int64_t g_bucket_min() { static const int64_t v = (int64_t)1 << (getenv("CENN_G_BUCKET_LOG2") ? atoi(getenv("CENN_G_BUCKET_LOG2")) : 20); return v; }
// chunks of one bucket (elements): multiples of 4, the last one takes the remainder
// (EXPERIMENT, default 1 = off: with CENN_BUCKET_CHUNKS=4 the step at N = 2 went from 3.33 to 3.43 ms, with 8 to 3.73 ms -- every extra NCCL
// launch costs more than the Adam pipelining saves)
static int bucket_chunks(int64_t cnt) { static const int v = getenv("CENN_BUCKET_CHUNKS") ? atoi(getenv("CENN_BUCKET_CHUNKS")) : 1; return (cnt >= ((int64_t)1 << 23) && v > 1) ? v : 1; }
static int64_t chunk_off(int64_t cnt, int nch, int c) { return c >= nch ? cnt : ((cnt / nch) & ~int64_t(3)) * c; }
static int64_t d_bucket_min() { static const int64_t v = (int64_t)1 << (getenv("CENN_D_BUCKET_LOG2") ? atoi(getenv("CENN_D_BUCKET_LOG2")) : 18); return v; }
// backward of one block: b->g holds dLoss/d(activation output); produces parameter gradients (if want_params)
// and the previous block's gradient (if the block has a dgrad plan and want_dgrad)
void emit_backward(T *t, Net &net, size_t i, bool want_params, bool want_dgrad) {
    cenn_state *s = t->s;
    Block *b = &net.blocks[i];
    float *grad = net.grad, *master = net.master;
    if (b->type == HEAD) {
        const bf16 *w = net.wbf + b->w_off;
        int B = b->Mrows, K = 16 * b->Clp;
        const bf16 *x = b->P > 1 ? b->col : b->in.p;
        bf16 *gx = b->P > 1 ? b->colg : net.blocks[i - 1].g.p;
        if (want_params) emit(t, "head_wgrad", [s, b, x, grad, B, K]() {
            LK(nhwc::head_wgrad_kernel, dim3(dim3((K / 8 + 127) / 128, 16)), dim3(128), 0, s->stream)(b->gpre, x, grad + b->w_off, grad + b->b_off, B, K); KLAUNCH(s); return 0; });
        emit(t, "head_dgrad", [s, b, w, gx, B, K]() {
            LK(nhwc::head_dgrad_kernel, dim3(grid1d(s, (int64_t)B * K / 8)), dim3(256), 0, s->stream)(b->gpre, w, gx, B, K); KLAUNCH(s); return 0; });
        if (b->P > 1) emit_col2im_v4(t, b->colg, net.blocks[i - 1].g, nullptr, nullptr, 0);
        return;
    }
    const int vpp = b->Coutp >= 8 ? b->Coutp / 8 : 1;
    float *gb_part = want_params ? b->part : nullptr;     // per-CTA partial sums of g_y -> conv gradBias (fold_gbias at the end of the sweep)
    const int64_t npix = b->a.pix();
    const bool dp = t->cfg.world_size > 1;
    if (b->bn) {
        const double n_global = bn_batch(t) * b->a.H * b->a.W;
        const bool dpb = bn_sync(t);          // the BN sums cross the ranks (global-batch statistics); false with cfg.bn_local
        float *gamma = master + b->g_off;
        float *gg = want_params ? grad + b->g_off : nullptr, *gbeta = want_params ? grad + b->be_off : nullptr;
        if (b->bar && !dpb) {       // single-launch path (nhwc::bn_bwd_fused_kernel), cooperative launch
            const int want_gb = want_params ? 1 : 0;
            emit(t, "bn_bwd_fused", [s, b, gamma, gg, gbeta, want_gb, npix, vpp, n_global]() {
                dim3 blk; int gy; reduce_dims(vpp, blk, gy);
                auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::bn_bwd_fused_kernel<nhwc::ACT_LEAKY> : nhwc::bn_bwd_fused_kernel<nhwc::ACT_RELU>;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(b->part_rows, gy); cfg.blockDim = blk; cfg.dynamicSmemBytes = 2 * blk.x * 8 * sizeof(float); cfg.stream = s->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                bf16 *gp = b->g.p; const bf16 *yp = b->y.p;
                const float *scale = b->scale, *shift = b->shift, *mean = b->mean, *invstd = b->invstd;
                if (cenn_check_cuda(cudaLaunchKernelEx(&cfg, kern, gp, yp, scale, shift, mean, invstd, (const float *)gamma, b->part, b->coef, gg, gbeta, want_gb,
                        b->Coutp, (int64_t)npix, vpp, b->Cout, 0.2f, n_global, b->bar), "cooperative launch", __FILE__, __LINE__)) return 1;
                s->launches++; return 0; });
            t->prog.back().bytes = 5.0 * 2.0 * (double)npix * b->Cout;      // read g, y; (re-read from L2); write g_y: 5 s bytes per element
        } else {
        // CENN_BN_BWD_2LAUNCH=1 (off by default): sums by fp32 atomics + coefficients in the apply prologue.  Measured on B200 (round 2):
        // the prologue's dependent loads and the done-counter cost what the coefficient launch cost (apply 29.6 us vs 17.4 + 11 us per layer)
        // and the step got 4 % SLOWER (3.29 vs 3.15 ms), so the three-launch chain stays the default.
        static const bool two_launch_env = getenv("CENN_BN_BWD_2LAUNCH") != nullptr && atoi(getenv("CENN_BN_BWD_2LAUNCH")) != 0;
        const bool two_launch = !dpb && two_launch_env;
        // data parallel with peer mailboxes: the local sums are accumulated by fp32 atomics straight into the exchange buffer (no fold launch);
        // replicas stay bit-identical because every rank adds the same published values in rank order
        const bool dp_atomic = dpb && s->xr_enabled && 2 * b->Coutp <= XR_MAXF && getenv("CENN_DP_BN_FOLD") == nullptr;
        // single GPU (default): pass 1 and the coefficient step in one launch (the last CTA finishes the sums); CENN_BN_BWD_3LAUNCH=1 restores
        // reduce -> coefficients -> apply
        static const bool fuse_coef_env = getenv("CENN_BN_BWD_3LAUNCH") == nullptr;
        const bool fuse_coef = !dpb && !two_launch && fuse_coef_env;
        if (b->bwd_epi && !dpb) {
            // the sums are already in bsums (epilogue of the dgrad GEMM that wrote g): coefficients + affine gradients, sums re-zeroed
            emit(t, "bn_bwd_coef", [s, b, gamma, gg, gbeta, n_global]() {
                LK(nhwc::bn_bwd_coef_sums_kernel, dim3((b->Coutp + 255) / 256), dim3(256), 0, s->stream)(b->bsums, b->Coutp, gamma, b->invstd, b->mean, b->coef, gg, gbeta, b->Cout, n_global);
                KLAUNCH(s); return 0; });
        } else if (b->bwd_epi) {
            // (data parallel: bn_bwd_coef_xr below exchanges and re-zeroes the same accumulator)
        } else if (fuse_coef) {
            emit(t, "bn_bwd_reduce", [s, b, gamma, gg, gbeta, npix, vpp, n_global]() {
                dim3 blk; int gy; reduce_dims(vpp, blk, gy);
                auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::bn_bwd_reduce_coef_kernel<nhwc::ACT_LEAKY> : nhwc::bn_bwd_reduce_coef_kernel<nhwc::ACT_RELU>;
                LK(kern, dim3(b->red_rows, gy), blk, 2 * blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->y.p, b->scale, b->shift, b->mean, b->invstd, gamma,
                    b->bsums, b->coef, gg, gbeta, b->Coutp, npix, vpp, b->Cout, 0.2f, n_global, b->done_ctr);
                KLAUNCH(s); return 0; });
            t->prog.back().bytes = 2.0 * 2.0 * (double)npix * b->Cout;        // read g, y
        } else
        emit(t, "bn_bwd_reduce", [s, b, npix, vpp, two_launch, dp_atomic]() {
            dim3 blk; int gy; reduce_dims(vpp, blk, gy);
            if (two_launch || dp_atomic) {
                auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::bn_bwd_reduce2_kernel<nhwc::ACT_LEAKY, true> : nhwc::bn_bwd_reduce2_kernel<nhwc::ACT_RELU, true>;
                LK(kern, dim3(dim3(b->red_rows, gy)), dim3(blk), 2 * blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->y.p, b->scale, b->shift, b->mean,
                    b->bsums, b->Coutp, npix, vpp, b->Cout, 0.2f);
                KLAUNCH(s); return 0;
            }
            auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::bn_bwd_reduce2_kernel<nhwc::ACT_LEAKY> : nhwc::bn_bwd_reduce2_kernel<nhwc::ACT_RELU>;
            LK(kern, dim3(dim3(b->part_rows, gy)), dim3(blk), 2 * blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->y.p, b->scale, b->shift, b->mean,
                b->part, b->Coutp, npix, vpp, b->Cout, 0.2f);
            KLAUNCH(s); return 0; });
        if (!b->bwd_epi && !fuse_coef) t->prog.back().bytes = 2.0 * 2.0 * (double)npix * b->Cout;            // read g, y
        if (two_launch) {
            emit(t, "bn_bwd_apply", [s, b, gamma, gg, gbeta, gb_part, npix, vpp, n_global]() {
                dim3 blk; int gy; reduce_dims(vpp, blk, gy);
                auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::bn_bwd_coef_apply_kernel<nhwc::ACT_LEAKY> : nhwc::bn_bwd_coef_apply_kernel<nhwc::ACT_RELU>;
                LK(kern, dim3(dim3(b->part_rows, gy)), dim3(blk), blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->y.p, b->scale, b->shift, b->bsums, gamma, b->invstd, b->mean,
                    gg, gbeta, gb_part, b->Coutp, npix, vpp, b->Cout, 0.2f, n_global, b->done_ctr);
                KLAUNCH(s); return 0; });
            t->prog.back().bytes = 3.0 * 2.0 * (double)npix * b->Cout;        // read g, y; write g_y
        } else {
        if (dp_atomic) {
            const float inv_world = 1.f / (float)t->cfg.world_size;
            emit(t, "bn_bwd_coef_xr", [s, b, gamma, gg, gbeta, n_global, inv_world]() {
                LK(nhwc::bn_bwd_coef_xr_kernel, dim3(1), dim3(1024), 0, s->stream)(s->xr, b->bsums, b->Coutp, gamma, b->invstd, b->mean, b->coef, gg, gbeta, b->Cout, n_global, inv_world, 1);
                KLAUNCH(s); return 0; });
        } else if (dpb && s->xr_enabled && 2 * b->Coutp <= XR_MAXF) {   // fold this rank's partial rows, then exchange + coefficients in one kernel
            const float inv_world = 1.f / (float)t->cfg.world_size;
            emit(t, "bn_bwd_fold", [s, b]() {
                LK(nhwc::bn_bwd_coef2_kernel, dim3((b->Cout + 31) / 32), dim3(dim3(32, 8)), 0, s->stream)(b->part, b->part_rows, b->bsums, b->Coutp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, b->Cout, 1.0, 0, 1.f);
                KLAUNCH(s); return 0; });
            emit(t, "bn_bwd_coef_xr", [s, b, gamma, gg, gbeta, n_global, inv_world]() {
                LK(nhwc::bn_bwd_coef_xr_kernel, dim3(1), dim3(1024), 0, s->stream)(s->xr, b->bsums, b->Coutp, gamma, b->invstd, b->mean, b->coef, gg, gbeta, b->Cout, n_global, inv_world, 0);
                KLAUNCH(s); return 0; });
        } else if (dpb) {   // fold the partial rows into bsums, all-reduce bsums across ranks, then the coefficients
            emit(t, "bn_bwd_fold", [s, b]() {
                LK(nhwc::bn_bwd_coef2_kernel, dim3((b->Cout + 31) / 32), dim3(dim3(32, 8)), 0, s->stream)(b->part, b->part_rows, b->bsums, b->Coutp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, b->Cout, 1.0, 0, 1.f);
                KLAUNCH(s); return 0; }, b->bsums, 2 * (int64_t)b->Coutp);
            const float inv_world = 1.f / (float)t->cfg.world_size;
            emit(t, "bn_bwd_coef", [s, b, gamma, gg, gbeta, n_global, inv_world]() {
                LK(nhwc::bn_bwd_coef2_kernel, dim3((b->Cout + 31) / 32), dim3(dim3(32, 8)), 0, s->stream)(nullptr, 0, b->bsums, b->Coutp, gamma, b->invstd, b->mean, b->coef, gg, gbeta, b->Cout, n_global, 1, inv_world);
                KLAUNCH(s); return 0; });
        } else if (!fuse_coef) {
            emit(t, "bn_bwd_coef", [s, b, gamma, gg, gbeta, n_global]() {
                LK(nhwc::bn_bwd_coef2_kernel, dim3((b->Cout + 31) / 32), dim3(dim3(32, 8)), 0, s->stream)(b->part, b->part_rows, nullptr, b->Coutp, gamma, b->invstd, b->mean, b->coef, gg, gbeta, b->Cout, n_global, 1, 1.f);
                KLAUNCH(s); return 0; });
        }
        emit(t, "bn_bwd_apply", [s, b, gb_part, npix, vpp]() {
            dim3 blk; int gy; reduce_dims(vpp, blk, gy);
            // coef is laid out with stride Cout
            auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::bn_bwd_apply2_kernel<nhwc::ACT_LEAKY> : nhwc::bn_bwd_apply2_kernel<nhwc::ACT_RELU>;
            LK(kern, dim3(dim3(b->part_rows, gy)), dim3(blk), blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->y.p, b->scale, b->shift, b->coef,
                gb_part, b->Coutp, npix, vpp, b->Cout, 0.2f);
            KLAUNCH(s); return 0; });
        t->prog.back().bytes = 3.0 * 2.0 * (double)npix * b->Cout;            // read g, y; write g_y
        }   // coefficient launch + apply
        }   // three-launch path
    } else if (b->Coutp >= 8) {
        emit(t, "act_bwd", [s, b, gb_part, npix, vpp]() {
            dim3 blk; int gy; reduce_dims(vpp, blk, gy);
            auto kern = b->act == nhwc::ACT_LEAKY ? nhwc::act_bwd2_kernel<nhwc::ACT_LEAKY> : (b->act == nhwc::ACT_RELU ? nhwc::act_bwd2_kernel<nhwc::ACT_RELU> : nhwc::act_bwd2_kernel<nhwc::ACT_TANH>);
            LK(kern, dim3(dim3(b->part_rows, gy)), dim3(blk), blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->a.p, gb_part, b->Coutp, npix, vpp, b->Cout, 0.2f);
            KLAUNCH(s); return 0; });
        t->prog.back().bytes = 3.0 * 2.0 * (double)npix * b->Cout;            // read g, a; write g_y
    } else {
        // Cp == 4 (3-channel image output): two pixels form one 8-lane vector, lanes k and k+4 are the same channel
        // (the fold job adds the two halves).  Pad lanes carry zero gradients.
        emit(t, "act_bwd4", [s, b, gb_part, npix]() {
            dim3 blk(1, 256);
            auto kern = b->act == nhwc::ACT_TANH ? nhwc::act_bwd2_kernel<nhwc::ACT_TANH> : (b->act == nhwc::ACT_LEAKY ? nhwc::act_bwd2_kernel<nhwc::ACT_LEAKY> : nhwc::act_bwd2_kernel<nhwc::ACT_RELU>);
            LK(kern, dim3(dim3(b->part_rows, 1)), dim3(blk), blk.x * 8 * sizeof(float), s->stream)(b->g.p, b->a.p, gb_part, 8, npix / 2, 1, 8, 0.2f);
            KLAUNCH(s); return 0; });
        t->prog.back().bytes = 3.0 * 2.0 * (double)npix * b->Cout;
    }
    if (b->nz && want_params) {
        float *gw = grad + b->nw_off;
        emit(t, "noise_wgrad", [t, s, b, gw]() {
            LK(nhwc::noise_wgrad_kernel, dim3(b->nz), dim3(128), (size_t)t->B * sizeof(float), s->stream)(b->g.p, b->Coutp, b->Cs, t->noise, gw, t->B, b->nz);
            KLAUNCH(s); return 0; });
    }
    if (b->type == JOIN5) {
        // both branches' weight gradients on the side stream; the window gradient + overlap-add of the prediction branch only where the
        // caller uses it (fGx: netD:updateGradInput, train.lua:369-371).  The gradInputs netD:backward computes in fDx and discards
        // (both branches) are not computed.
        if (want_params)
            for (TcPlan *pl : {&b->p_wgrad2, &b->p_wgrad}) {
                cudaEvent_t evf; cudaEventCreateWithFlags(&evf, cudaEventDisableTiming); t->events.push_back(evf);
                t->flops_per_step += pl->flops;
                emit(t, "wgrad", [t, s, evf, pl]() {
                    if (t->serial) return tc_launch(s, pl);
                    if (cenn_check_cuda(cudaEventRecord(evf, s->stream), "event record", __FILE__, __LINE__)) return 1;
                    if (cenn_check_cuda(cudaStreamWaitEvent(t->side, evf, 0), "stream wait", __FILE__, __LINE__)) return 1;
                    cudaStream_t keep = s->stream; s->stream = t->side;
                    int rc = tc_launch(s, pl);
                    s->stream = keep;
                    return rc; });
                t->prog.back().flops = pl->flops;
            }
        else if (want_dgrad) {
            emit_plan(t, "dgrad", &b->p_dgrad);
            emit_col2im5(t, b->colg, t->df_dg, b->h, b->w, b->pad5);
        }
        return;
    }
    if (b->type == FULL_V4 && b->P > 1 && (want_params || (want_dgrad && b->has_dgrad))) emit_im2col_v4(t, b->g, b->col);
    // weight gradient
    if (want_params) {
        if (b->thin && b->type == FULL_S2) emit_im2col(t, b->g, b->col, b->h, b->w);
        // The weight gradient of this block depends only on g_y and the block input; the critical path continues with this
        // block's dgrad and the previous block's BN backward (small kernels, and in data-parallel runs a peer exchange).
        // It therefore runs on a side stream (joined at the end of the sweep) and fills the gaps of that chain.
        {
            cudaEvent_t evf; cudaEventCreateWithFlags(&evf, cudaEventDisableTiming); t->events.push_back(evf);
            TcPlan *pl = &b->p_wgrad;
            t->flops_per_step += pl->flops;
            emit(t, "wgrad", [t, s, evf, pl]() {
                if (t->serial) return tc_launch(s, pl);
                if (cenn_check_cuda(cudaEventRecord(evf, s->stream), "event record", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side, evf, 0), "stream wait", __FILE__, __LINE__)) return 1;
                cudaStream_t keep = s->stream; s->stream = t->side;
                int rc = tc_launch(s, pl);
                s->stream = keep;
                return rc; });
            t->prog.back().flops = pl->flops;
        }
        // data parallel, generator: the big weight gradients start their all-reduce now, on the bulk communicator's stream,
        // and overlap the rest of the backward sweep (E6 + G1 are 92 % of the 285 MB)
        const bool bucket_g = &net == &t->G && b->w_count >= g_bucket_min(), bucket_d = &net == &t->D && t->d_bucket_sweep && b->w_count >= d_bucket_min();
        if (dp && s->comm2 && bucket_g && shard_block(t, net, b->w_count)) {
            // sharded reduce + Adam: only the bf16 conversion here (behind the wgrad); the exchange itself follows this block's dgrad, the last
            // reader of the operand copy that the peers will overwrite
            cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); t->events.push_back(ev);
            float *ptr = grad + b->w_off; bf16 *pbf = net.gradbf + b->w_off; const int64_t cnt = b->w_count;
            t->g_buckets.push_back({b->w_off, b->w_count});
            emit(t, "grad_shard_cvt", [t, s, ev, ptr, pbf, cnt]() {
                if (cenn_check_cuda(cudaEventRecord(ev, t->serial ? s->stream : t->side), "event record", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(s->comm_stream, ev, 0), "stream wait", __FILE__, __LINE__)) return 1;
                nhwc::f32_to_bf16_vec_kernel<<<grid1d(s, cnt / 4), 256, 0, s->comm_stream>>>(ptr, pbf, cnt / 4);
                KLAUNCH(s); return 0; });
            t->prog.back().bytes = 6.0 * (double)cnt;
        } else if (dp && s->comm2 && (bucket_g || bucket_d)) {
            cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); t->events.push_back(ev);
            float *ptr = grad + b->w_off; int64_t cnt = b->w_count;
            (bucket_g ? t->g_buckets : t->d_buckets).push_back({b->w_off, b->w_count});
            // generator buckets with an early Adam behind them travel as bf16 (half the NVLink bytes); Adam reads the bf16 sum
            bf16 *pbf = (net.gradbf && cnt % 4 == 0 && getenv("CENN_NO_EARLY_ADAM") == nullptr) ? net.gradbf + b->w_off : nullptr;
            // optional (bucket_chunks): the two 32.8 M-element blocks (E6, G1) travel in chunks, the early Adam of chunk k (third stream) overlaps
            // the all-reduce of chunk k + 1
            const int nch = bucket_chunks(cnt);
            if (bucket_g) { for (int c = 0; c < nch; ++c) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); t->events.push_back(e); b->ar_ev.push_back(e); } }
            emit(t, "grad_bucket_ar", [t, s, b, ev, ptr, pbf, cnt, nch, bucket_g]() {
                if (cenn_check_cuda(cudaEventRecord(ev, t->serial ? s->stream : t->side), "event record", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(s->comm_stream, ev, 0), "stream wait", __FILE__, __LINE__)) return 1;
                for (int c = 0; c < nch; ++c) {
                    const int64_t o = chunk_off(cnt, nch, c), n = chunk_off(cnt, nch, c + 1) - o;
                    if (!pbf) { if (cenn_dist_all_reduce_bulk(s, ptr + o, n)) return 1; }
                    else {
                        nhwc::f32_to_bf16_vec_kernel<<<grid1d(s, n / 4), 256, 0, s->comm_stream>>>(ptr + o, pbf + o, n / 4);
                        KLAUNCH(s);
                        if (cenn_dist_all_reduce_bulk_bf16(s, pbf + o, n)) return 1;
                    }
                    if (bucket_g && cenn_check_cuda(cudaEventRecord(b->ar_ev[c], s->comm_stream), "event record", __FILE__, __LINE__)) return 1;
                }
                return 0; });
        }
    } else if (b->thin && b->type == FULL_S2 && want_dgrad && b->has_dgrad) {
        emit_im2col(t, b->g, b->col, b->h, b->w);
    }
    if (want_dgrad && b->has_dgrad) {
        if (i == 0 && want_params && &net == &t->G && defer_dead_dgrad()) {
            // emitted at the START of the step program (build_program): see there
        } else if (i == 0 && want_params) {
            // first block of a net in a parameter sweep: the reference computes this gradInput and discards it (dead_dgrad); nothing waits for
            // it, so it runs on the side stream after the block's wgrad instead of on the critical path
            cudaEvent_t evf; cudaEventCreateWithFlags(&evf, cudaEventDisableTiming); t->events.push_back(evf);
            TcPlan *pl = &b->p_dgrad;
            t->flops_per_step += pl->flops;
            emit(t, "dgrad", [t, s, evf, pl]() {
                if (t->serial) return tc_launch(s, pl);
                if (cenn_check_cuda(cudaEventRecord(evf, s->stream), "event record", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side, evf, 0), "stream wait", __FILE__, __LINE__)) return 1;
                cudaStream_t keep = s->stream; s->stream = t->side;
                int rc = tc_launch(s, pl);
                s->stream = keep;
                return rc; });
            t->prog.back().flops = pl->flops;
        } else emit_plan(t, "dgrad", &b->p_dgrad);
        if (b->type == CONV_V4 && b->P > 1) emit_col2im_v4(t, b->colg, net.blocks[i - 1].g, nullptr, nullptr, 0);
    }
    // single GPU, generator: once this block's wgrad (side stream) and dgrad (this stream, the last reader of its weights)
    // are queued, its slice of the flat vector can take its Adam update on a third stream while the sweep goes on
    // (E6 + G1 hold 92 % of the parameters: ~0.3 ms of HBM-bound work moved off the critical path)
    // (data parallel: same, chained behind the block's bucket all-reduce on the bulk communicator's stream)
    const bool dp_bulk = dp && s->comm2 != nullptr;
    if (want_params && dp_bulk && shard_block(t, net, b->w_count)) {
        cudaEvent_t e1; cudaEventCreateWithFlags(&e1, cudaEventDisableTiming); t->events.push_back(e1);
        Net *n = &net; const int64_t off = b->w_off, cnt = b->w_count; const float beta1 = t->cfg.beta1;
        const int world = t->cfg.world_size, rank = t->cfg.rank;
        t->g_early.push_back({off, cnt});
        net.shard_blocks.push_back({off, cnt});
        nhwc::ShardPeers pp = {};
        for (int r = 0; r < world; ++r) { pp.grad[r] = peer_ptr<bf16>(net.ipc_gradbf, r, net.gradbf, net.gradbf + off); pp.wbf[r] = peer_ptr<bf16>(net.ipc_wbf, r, net.base_wbf, net.wbf + off); }
        const int64_t begin = shard_begin(cnt, world, rank), mine = shard_begin(cnt, world, rank + 1) - begin;
        emit(t, "shard_adam", [t, s, n, e1, pp, world, off, begin, mine, beta1]() {
            // communication stream, behind this block's dgrad: barrier (every rank has converted its gradient and no longer reads its operand
            // copy) -> owner reduces + updates its shard and stores the new bf16 weights everywhere -> barrier (my operand copy is complete)
            if (cenn_check_cuda(cudaEventRecord(e1, s->stream), "event record", __FILE__, __LINE__)) return 1;
            if (cenn_check_cuda(cudaStreamWaitEvent(s->comm_stream, e1, 0), "stream wait", __FILE__, __LINE__)) return 1;
            nhwc::xr_barrier_kernel<<<1, 32, 0, s->comm_stream>>>(s->xr3); KLAUNCH(s);
            if (mine > 0) {
                nhwc::shard_reduce_adam_kernel<<<grid1d(s, mine / 8), 256, 0, s->comm_stream>>>(pp, world, n->master + off, n->grad + off, n->m + off, n->v + off,
                    begin, mine, beta1, 0.999f, 1e-8f, n->adam_step);
                KLAUNCH(s);
            }
            nhwc::xr_barrier_kernel<<<1, 32, 0, s->comm_stream>>>(s->xr3); KLAUNCH(s);
            return 0; });
        t->prog.back().bytes = (28.0 + 4.0) * (double)mine + 2.0 * 2.0 * (double)mine * (world - 1);
    } else if (want_params && (!dp || dp_bulk) && &net == &t->G && b->w_count >= g_bucket_min() && getenv("CENN_NO_EARLY_ADAM") == nullptr) {
        cudaEvent_t e1, e2; cudaEventCreateWithFlags(&e1, cudaEventDisableTiming); cudaEventCreateWithFlags(&e2, cudaEventDisableTiming);
        t->events.push_back(e1); t->events.push_back(e2);
        Net *n = &net; const int64_t off = b->w_off, cnt = b->w_count; const float beta1 = t->cfg.beta1;
        t->g_early.push_back({off, cnt});
        emit(t, "adam_early", [t, s, b, n, e1, e2, off, cnt, beta1, dp_bulk]() {
            cudaStream_t st = s->stream;
            if (dp_bulk && !b->ar_ev.empty()) {      // third stream: behind this block's dgrad, chunk by chunk behind the chunk's all-reduce (comm_stream)
                if (cenn_check_cuda(cudaEventRecord(e1, s->stream), "event record", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side3, e1, 0), "stream wait", __FILE__, __LINE__)) return 1;
                const int nch = (int)b->ar_ev.size();
                for (int c = 0; c < nch; ++c) {
                    const int64_t o = off + chunk_off(cnt, nch, c), nn = chunk_off(cnt, nch, c + 1) - chunk_off(cnt, nch, c);
                    if (cenn_check_cuda(cudaStreamWaitEvent(t->side3, b->ar_ev[c], 0), "stream wait", __FILE__, __LINE__)) return 1;
                    if (n->gradbf && cnt % 4 == 0)
                        LK(nhwc::adam_bf16g_kernel, dim3(grid1d(s, nn / 4)), dim3(256), 0, t->side3)(n->master + o, n->gradbf + o, n->m + o, n->v + o, n->wbf + o, nn, beta1, 0.999f, 1e-8f, n->adam_step);
                    else
                        LK(nhwc::adam_bf16_kernel, dim3(grid1d(s, nn / 4)), dim3(256), 0, t->side3)(n->master + o, n->grad + o, n->m + o, n->v + o, n->wbf + o, nn, beta1, 0.999f, 1e-8f, n->adam_step);
                    KLAUNCH(s);
                }
                return 0;
            }
            if (dp_bulk) {                           // third stream: waits for this block's bucket all-reduce (comm_stream) and its dgrad
                if (cenn_check_cuda(cudaEventRecord(e1, s->stream), "event record", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaEventRecord(e2, s->comm_stream), "event record", __FILE__, __LINE__)) return 1;   // after the all-reduce
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side3, e1, 0), "stream wait", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side3, e2, 0), "stream wait", __FILE__, __LINE__)) return 1;
                st = t->side3;                       // the next bucket's all-reduce does not queue behind this Adam
            } else if (!t->serial) {
                if (cenn_check_cuda(cudaEventRecord(e1, s->stream), "event record", __FILE__, __LINE__)) return 1;   // after dgrad
                if (cenn_check_cuda(cudaEventRecord(e2, t->side), "event record", __FILE__, __LINE__)) return 1;     // after wgrad
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side3, e1, 0), "stream wait", __FILE__, __LINE__)) return 1;
                if (cenn_check_cuda(cudaStreamWaitEvent(t->side3, e2, 0), "stream wait", __FILE__, __LINE__)) return 1;
                st = t->side3;
            }
            if (dp_bulk && n->gradbf && cnt % 4 == 0)
                LK(nhwc::adam_bf16g_kernel, dim3(grid1d(s, cnt / 4)), dim3(256), 0, st)(n->master + off, n->gradbf + off, n->m + off, n->v + off, n->wbf + off, cnt, beta1, 0.999f, 1e-8f, n->adam_step);
            else
            LK(nhwc::adam_bf16_kernel, dim3(grid1d(s, cnt / 4)), dim3(256), 0, st)(n->master + off, n->grad + off, n->m + off, n->v + off, n->wbf + off, cnt, beta1, 0.999f, 1e-8f, n->adam_step);
            KLAUNCH(s); return 0; });
        t->prog.back().bytes = 30.0 * (double)cnt;
    }
}

}  // namespace

namespace {

// refresh the bf16 operand copies that are transposes of the master copy
void emit_weight_prep(T *t, Net &net) {
    cenn_state *s = t->s;
    for (size_t i = 0; i < net.blocks.size(); ++i) {
        Block *b = &net.blocks[i];
        if (!b->Wt) continue;
        const bf16 *Wf = net.wbf + b->w_off;
        int K = 16 * b->Clp;
        int mode = (b->type == CONV_S2 || b->type == FULL_S2) ? 1 : 0;
        emit(t, "wt_from_wf", [s, b, Wf, K, mode]() {
            dim3 grid((K + 31) / 32, (b->Csp + 31) / 32);
            LK(nhwc::wt_from_wf_kernel, dim3(grid), dim3(256), 0, s->stream)(Wf, b->Wt, b->Cs, K, b->Csp, b->Clp, b->cl_rows, mode);
            KLAUNCH(s); return 0; });
    }
}
void emit_zero_bias(T *t, Net &net) {
    cenn_state *s = t->s;
    Net *n = &net;
    emit(t, "zero_conv_bias", [s, n]() {
        LK(nhwc::zero_segments_kernel, dim3(n->nbias_seg), dim3(128), 0, s->stream)(n->master, n->wbf, n->bias_seg, n->nbias_seg); KLAUNCH(s); return 0; });
}
void emit_fold_gbias(T *t, Net &net) {
    cenn_state *s = t->s;
    Net *n = &net;
    {   // join the side stream (weight gradients of this sweep)
        cudaEvent_t evj; cudaEventCreateWithFlags(&evj, cudaEventDisableTiming); t->events.push_back(evj);
        emit(t, "join_wgrad", [t, s, evj]() {
            if (t->serial) return 0;
            if (cenn_check_cuda(cudaEventRecord(evj, t->side), "event record", __FILE__, __LINE__)) return 1;
            return cenn_check_cuda(cudaStreamWaitEvent(s->stream, evj, 0), "stream wait", __FILE__, __LINE__); });
    }
    emit(t, "fold_gbias", [s, n]() {
        LK(nhwc::fold_rows_kernel, dim3((unsigned)n->fold_host.size(), 8), dim3(1024), 0, s->stream)(n->fold_jobs); KLAUNCH(s); return 0; });
}
void emit_zero_grad(T *t, Net &net) {
    cenn_state *s = t->s;
    Net *n = &net;
    // zero everything except the weight ranges whose wgrad launch stores its result
    std::vector<std::pair<int64_t, int64_t>> segs;
    int64_t cur = 0;
    for (const Block &b : net.blocks)
        if (b.p_wgrad.overwrites && b.w_count > 0) { if (b.w_off > cur) segs.push_back({cur, b.w_off - cur}); cur = b.w_off + b.w_count; }
    if (net.nparam > cur) segs.push_back({cur, net.nparam - cur});
    emit(t, "zero_grad", [s, n, segs]() {
        for (const auto &sg : segs)
            if (cenn_check_cuda(cudaMemsetAsync(n->grad + sg.first, 0, sg.second * sizeof(float), s->stream), "memset", __FILE__, __LINE__)) return 1;
        return 0; });
}
void emit_adam_step(T *t, Net &net) {
    cenn_state *s = t->s;
    Net *n = &net;
    float beta1 = t->cfg.beta1;
    emit(t, "adam_step", [s, n, beta1]() { adam_step_kernel<<<1, 32, 0, s->stream>>>(n->adam_t, n->adam_step, n->lr, beta1, 0.999f); KLAUNCH(s); return 0; });
}
// Adam over the ranges of the flat vector not in `done` (sorted on use)
void emit_adam(T *t, Net &net, std::vector<std::pair<int64_t, int64_t>> done = {}) {
    cenn_state *s = t->s;
    Net *n = &net;
    float beta1 = t->cfg.beta1;
    std::sort(done.begin(), done.end());
    std::vector<std::pair<int64_t, int64_t>> rest;
    int64_t cur = 0;
    for (auto &x : done) { if (x.first > cur) rest.push_back({cur, x.first - cur}); cur = x.first + x.second; }
    if (net.nparam > cur) rest.push_back({cur, net.nparam - cur});
    emit(t, "adam", [s, n, beta1, rest]() {
        for (auto &x : rest) {
            LK(nhwc::adam_bf16_kernel, dim3(grid1d(s, x.second / 4)), dim3(256), 0, s->stream)(n->master + x.first, n->grad + x.first, n->m + x.first, n->v + x.first, n->wbf + x.first,
                x.second, beta1, 0.999f, 1e-8f, n->adam_step);
            KLAUNCH(s);
        }
        return 0; });
    { double by = 0; for (auto &x : rest) by += 30.0 * (double)x.second; t->prog.back().bytes = by; }
}
void emit_bce(T *t, Block *head, float label, int loss_slot, bool want_grad) {
    cenn_state *s = t->s;
    // one BCE term per discriminator output: batchSize at fineSize 128; 25 per sample at fineSize 256, each carrying its sample's
    // label (the cfg4 label convention, DESIGN.md section 6 / oracle/step.py:_label)
    double inv_n = 1.0 / ((double)t->Bglobal * head->P);
    int B = head->Mrows;
    double *acc = t->loss_acc + loss_slot;
    emit(t, "bce", [s, head, label, acc, B, inv_n, want_grad]() {
        LK(nhwc::head_bce_kernel, dim3((B + 255) / 256), dim3(256), 0, s->stream)(head->sig, label, want_grad ? head->gpre : nullptr, acc, B, inv_n); KLAUNCH(s); return 0; });
}
void emit_copy(T *t, const char *name, bf16 *dst, const bf16 *src, int64_t elems) {
    cenn_state *s = t->s;
    emit(t, name, [s, dst, src, elems]() { return cenn_check_cuda(cudaMemcpyAsync(dst, src, elems * sizeof(bf16), cudaMemcpyDeviceToDevice, s->stream), "d2d copy", __FILE__, __LINE__); });
}

// losses: device double accumulators -> float outputs
__global__ void finish_losses_kernel(const double *__restrict__ acc, float *__restrict__ out, float wtl2, float wtgdl) {
    if (threadIdx.x != 0) return;
    double errD_real = acc[CENN_LOSS_ERRD_REAL], errD_fake = acc[CENN_LOSS_ERRD_FAKE], errG = acc[CENN_LOSS_ERRG], l2 = acc[CENN_LOSS_ERRG_L2], gdl = acc[CENN_LOSS_ERRG_GDL];
    out[CENN_LOSS_ERRD] = (float)(errD_real + errD_fake);
    out[CENN_LOSS_ERRG] = (float)errG; out[CENN_LOSS_ERRG_L2] = (float)l2; out[CENN_LOSS_ERRG_GDL] = (float)gdl;
    out[CENN_LOSS_ERRD_REAL] = (float)errD_real; out[CENN_LOSS_ERRD_FAKE] = (float)errD_fake;
    double total = errG;
    if (wtl2 != 0.f) total = (wtl2 > 0.f && wtl2 < 1.f) ? (1.0 - wtl2) * errG + wtl2 * l2 : errG + wtl2 * l2;
    if (wtgdl != 0.f) total += wtgdl * gdl;
    out[CENN_LOSS_ERRG_TOTAL] = (float)total; out[7] = 0.f;
}

// leftover gradient ranges summed over the ranks through peer memory (nhwc::peer_allreduce_f32_kernel) on the compute stream; false if the
// ranges do not fit the kernel's segment table (the caller then uses NCCL)
bool emit_peer_allreduce(T *t, const char *name, Net &net, const std::vector<std::pair<int64_t, int64_t>> &rest, cudaEvent_t ev_buckets) {
    cenn_state *s = t->s;
    if (!t->peer_ar_ok || rest.size() > 16) return false;
    nhwc::PeerSegs sg = {};
    for (const auto &x : rest) { if (x.first % 4 || x.second % 4) return false; sg.off[sg.n] = x.first; sg.cnt[sg.n] = x.second; ++sg.n; }
    nhwc::PeerVec pv = {};
    const int world = t->cfg.world_size, rank = t->cfg.rank;
    for (int r = 0; r < world; ++r) pv.p[r] = peer_ptr<float>(net.ipc_grad, r, net.base_grad, net.grad);
    int64_t total = 0; for (const auto &x : rest) total += x.second;
    emit(t, name, [s, pv, sg, world, rank, total, ev_buckets]() {
        // the bucket all-reduces (bulk communicator, comm_stream) and the sharded blocks finish first
        if (cenn_check_cuda(cudaEventRecord(ev_buckets, s->comm_stream), "event record", __FILE__, __LINE__)) return 1;
        if (cenn_check_cuda(cudaStreamWaitEvent(s->stream, ev_buckets, 0), "stream wait", __FILE__, __LINE__)) return 1;
        nhwc::xr_barrier_kernel<<<1, 32, 0, s->stream>>>(s->xr); KLAUNCH(s);
        nhwc::peer_allreduce_f32_kernel<<<grid1d(s, (total / 4 + world - 1) / world), 256, 0, s->stream>>>(pv, sg, world, rank); KLAUNCH(s);
        nhwc::xr_barrier_kernel<<<1, 32, 0, s->stream>>>(s->xr); KLAUNCH(s);
        return 0; });
    return true;
}

// ---- the step program ---------------------------------------------------------------------------------------------
int build_program(T *t) {
    cenn_state *s = t->s;
    const cenn_trainer_config &c = t->cfg;
    Net &G = t->G, &D = t->D;
    const bool video = c.variant == 1;
    Block *headD = &D.blocks.back();
    Block *lastG = &G.blocks.back();
    t->prog.clear();
    t->join_ctx_done = false;
    t->g_buckets.clear(); t->d_buckets.clear();
    t->flops_per_step = 0;
    // -- inputs: fp32 NCHW (+ uint8 mask) -> NHWC bf16
    emit(t, "convert_inputs", [t, s, video]() {
        const Tensor &a = t->real_ctx, &b = t->real_aux;
        if (!video && !t->cur_a) {      // byte-image mode (cenn_trainer_step_images_u8_*): centre crop + mean fill on the device
            if (a.Cp != 4 || a.C != 3) { cenn_set_error("byte-image mode needs 3-channel images"); return 1; }
            LK(nhwc::image_u8_prepare_kernel, dim3(grid1d(s, a.pix())), dim3(256), 0, s->stream)(reinterpret_cast<const uint8_t *>(t->cur_b), a.N, a.H, t->cfg.overlapPred, a.p, b.p);
            KLAUNCH(s);
            return cenn_check_cuda(cudaMemsetAsync(t->loss_acc, 0, 8 * sizeof(double), s->stream), "memset", __FILE__, __LINE__);
        }
        if (video && !t->cur_a) {       // clip mode: masked / full / expanded mask derived on the device from the frames and one mask plane
            const Tensor &m = t->mask;
            const float mv = t->clip_mv;
            if (a.Cp == 16) LK(nhwc::clip_prepare_kernel<16>, dim3(grid1d(s, a.pix())), dim3(256), 0, s->stream)(t->cur_b, t->cur_m, t->cur_f, mv, a.N, a.C, a.H, a.W, a.p, b.p, m.p);
            else if (a.Cp == 4) LK(nhwc::clip_prepare_kernel<4>, dim3(grid1d(s, a.pix())), dim3(256), 0, s->stream)(t->cur_b, t->cur_m, t->cur_f, mv, a.N, a.C, a.H, a.W, a.p, b.p, m.p);
            else { cenn_set_error("clip mode: unsupported channel padding %d", a.Cp); return 1; }
            KLAUNCH(s);
            return cenn_check_cuda(cudaMemsetAsync(t->loss_acc, 0, 8 * sizeof(double), s->stream), "memset", __FILE__, __LINE__);
        }
        for (const Tensor *x : {&a, &b}) {
            const float *src = x == &a ? t->cur_a : t->cur_b;
            if (x->Cp == 4 && (x->H * x->W) % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0)
                LK(nhwc::to_nhwc4_kernel, dim3(grid1d(s, x->pix() / 4)), dim3(256), 0, s->stream)(src, x->p, x->N, x->C, x->H * x->W);
            else
                LK(nhwc::to_nhwc_kernel<float>, dim3(grid1d(s, x->pix())), dim3(256), 0, s->stream)(src, x->p, x->N, x->C, x->H * x->W, x->Cp);
            KLAUNCH(s);
        }
        if (video) { const Tensor &m = t->mask; LK(nhwc::to_nhwc_kernel<uint8_t>, dim3(grid1d(s, m.pix())), dim3(256), 0, s->stream)(t->cur_m, m.p, m.N, m.C, m.H * m.W, m.Cp); KLAUNCH(s); }
        return cenn_check_cuda(cudaMemsetAsync(t->loss_acc, 0, 8 * sizeof(double), s->stream), "memset", __FILE__, __LINE__); });
    // ================= fDx (train.lua:278-350) =================
    emit_zero_bias(t, D); emit_zero_bias(t, G);
    emit_zero_grad(t, D);
    if (c.dead_dgrad && defer_dead_dgrad() && !G.blocks.empty() && G.blocks[0].has_dgrad) {
        cudaEvent_t evf; cudaEventCreateWithFlags(&evf, cudaEventDisableTiming); t->events.push_back(evf);
        TcPlan *pl = &G.blocks[0].p_dgrad;
        t->flops_per_step += pl->flops;
        emit(t, "dgrad", [t, s, evf, pl]() {          // dead gradInput of the generator's first layer, one step late (defer_dead_dgrad)
            if (t->serial) return tc_launch(s, pl);
            if (cenn_check_cuda(cudaEventRecord(evf, s->stream), "event record", __FILE__, __LINE__)) return 1;
            if (cenn_check_cuda(cudaStreamWaitEvent(t->side, evf, 0), "stream wait", __FILE__, __LINE__)) return 1;
            cudaStream_t keep = s->stream; s->stream = t->side;
            int rc = tc_launch(s, pl);
            s->stream = keep;
            return rc; });
        t->prog.back().flops = pl->flops;
    }
    // The generator forward does not depend on the discriminator's real sweep (and vice versa): it runs as a second chain
    // on its own stream; each chain's small BN kernels / peer exchanges fill the other's gaps.  (Not with NCCL-only
    // data parallelism: one communicator must not be driven from two streams.)
    const bool two_chains = !(bn_sync(t) && !s->xr_enabled) && getenv("CENN_ONE_CHAIN") == nullptr;
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    if (two_chains) {
        cudaEventCreateWithFlags(&ev_fork2, cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_join2, cudaEventDisableTiming);
        t->events.push_back(ev_fork2); t->events.push_back(ev_join2);
        emit(t, "fork_G_fwd", [t, s, ev_fork2]() {
            if (t->serial) return 0;
            if (cenn_check_cuda(cudaEventRecord(ev_fork2, s->stream), "event record", __FILE__, __LINE__)) return 1;
            return cenn_check_cuda(cudaStreamWaitEvent(t->side2, ev_fork2, 0), "stream wait", __FILE__, __LINE__); });
        t->emit_chain = 1;
        emit_copy(t, "g_in<-ctx", G.input.p, t->real_ctx.p, G.input.elems());
        for (size_t i = 0; i < G.blocks.size(); ++i) emit_forward(t, G, i, true);
        t->emit_chain = 0;
    }
    // D on real
    emit_copy(t, "d_in<-real", D.input.p, t->real_aux.p, D.input.elems());
    for (size_t i = 0; i < D.blocks.size(); ++i) emit_forward(t, D, i, true);
    emit_bce(t, headD, 1.f, CENN_LOSS_ERRD_REAL, true);
    for (size_t i = D.blocks.size(); i-- > 0;) emit_backward(t, D, i, true, i > 0 || c.dead_dgrad);
    emit_fold_gbias(t, D);
    // G forward
    if (two_chains) {
        emit(t, "join_G_fwd", [t, s, ev_join2]() {
            if (t->serial) return 0;
            if (cenn_check_cuda(cudaEventRecord(ev_join2, t->side2), "event record", __FILE__, __LINE__)) return 1;
            return cenn_check_cuda(cudaStreamWaitEvent(s->stream, ev_join2, 0), "stream wait", __FILE__, __LINE__); });
    } else {
        emit_copy(t, "g_in<-ctx", G.input.p, t->real_ctx.p, G.input.elems());
        for (size_t i = 0; i < G.blocks.size(); ++i) emit_forward(t, G, i, true);
    }
    // D on fake
    if (video && c.weight_nomask == 0.f) {
        emit_copy(t, "d_in<-real", D.input.p, t->real_aux.p, D.input.elems());
        bf16 *dst = D.input.p; const bf16 *m = t->mask.p, *src = lastG->a.p; int64_t total = D.input.elems();
        emit(t, "composite", [s, dst, m, src, total]() { LK(nhwc::composite_kernel, dim3(grid1d(s, total)), dim3(256), 0, s->stream)(dst, m, src, total); KLAUNCH(s); return 0; });
    } else {
        emit_copy(t, "d_in<-fake", D.input.p, lastG->a.p, D.input.elems());
    }
    for (size_t i = 0; i < D.blocks.size(); ++i) emit_forward(t, D, i, true);
    emit_bce(t, headD, 0.f, CENN_LOSS_ERRD_FAKE, true);
    t->d_bucket_sweep = true;      // D's gradients are complete after this sweep: the big blocks start their all-reduce as they finish
    for (size_t i = D.blocks.size(); i-- > 0;) emit_backward(t, D, i, true, i > 0 || c.dead_dgrad);
    t->d_bucket_sweep = false;
    emit_fold_gbias(t, D);
    if (t->d_buckets.empty()) emit(t, "gradD_sync", []() { return 0; }, D.grad, D.nparam);
    else {
        std::vector<std::pair<int64_t, int64_t>> bk = t->d_buckets, rest;
        std::sort(bk.begin(), bk.end());
        int64_t cur = 0;
        for (auto &x : bk) { if (x.first > cur) rest.push_back({cur, x.first - cur}); cur = x.first + x.second; }
        if (D.nparam > cur) rest.push_back({cur, D.nparam - cur});
        cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); t->events.push_back(ev);
        float *g = D.grad;
        if (emit_peer_allreduce(t, "gradD_sync", D, rest, ev)) {} else
        emit(t, "gradD_sync", [s, rest, g, ev]() {
            // the bucket all-reduces (bulk communicator, comm_stream) finish first: two communicators are never in flight at once
            // (NCCL documents concurrent collectives on two communicators as a hang risk unless both kernels can be co-resident)
            if (cenn_check_cuda(cudaEventRecord(ev, s->comm_stream), "event record", __FILE__, __LINE__)) return 1;
            if (cenn_check_cuda(cudaStreamWaitEvent(s->stream, ev, 0), "stream wait", __FILE__, __LINE__)) return 1;
            if (cenn_dist_group(1)) return 1;
            for (auto &x : rest) if (cenn_dist_all_reduce_on(s, g + x.first, x.second, 0, s->stream)) return 1;
            return cenn_dist_group(0); });
    }
    emit_adam_step(t, D);
    emit_adam(t, D);
    emit_weight_prep(t, D);
    // ================= fGx (train.lua:353-410) =================
    emit_zero_bias(t, D); emit_zero_bias(t, G);
    emit_zero_grad(t, G);
    emit_bce(t, headD, 1.f, CENN_LOSS_ERRG, true);
    for (size_t i = D.blocks.size(); i-- > 0;) emit_backward(t, D, i, false, true);   // netD:updateGradInput
    // blend with the L2 term -> gradient w.r.t. G's output
    {
        const Tensor fake_in = D.input;   // input_center / input_inpainted as seen by D
        const Tensor real = t->real_aux, df = t->df_dg, gout = lastG->g, mk = t->mask;
        const double n = (double)t->Bglobal * fake_in.C * fake_in.H * fake_in.W;
        float a = (c.wtl2 > 0.f && c.wtl2 < 1.f) ? 1.f - c.wtl2 : 1.f;
        double *acc = t->loss_acc + CENN_LOSS_ERRG_L2;
        if (!video) {
            float w_in = c.wtl2, w_ring = c.overlapPred > 0 ? 10.f * c.wtl2 : c.wtl2;
            int ov = c.overlapPred;
            if (c.wtl2 != 0.f) {
                emit(t, "blend_overlap", [s, df, fake_in, real, gout, ov, a, w_in, w_ring, n, acc]() {
                    if (fake_in.Cp == 4 && fake_in.pix() % 2 == 0)
                        LK(nhwc::blend_overlap4_kernel, dim3(grid1d(s, fake_in.pix() / 2, 256, 4)), dim3(256), 0, s->stream)(df.p, fake_in.p, real.p, gout.p, fake_in.pix(), fake_in.H,
                            fake_in.W, fake_in.C, ov, a, w_in, w_ring, (float)(2.0 / n), 1.0 / n, acc);
                    else
                    LK(nhwc::blend_overlap_kernel, dim3(grid1d(s, fake_in.elems(), 256, 4)), dim3(256), 0, s->stream)(df.p, fake_in.p, real.p, gout.p, fake_in.pix(), fake_in.H,
                        fake_in.W, fake_in.Cp, fake_in.C, ov, a, w_in, w_ring, (float)(2.0 / n), 1.0 / n, acc);
                    KLAUNCH(s); return 0; });
                t->prog.back().bytes = 4.0 * 2.0 * (double)fake_in.pix() * fake_in.C;   // read df, x, t; write df (edge weight from indices)
            } else emit_copy(t, "g<-df_dg", gout.p, df.p, gout.elems());
        } else {
            float wtl2 = c.wtl2, lam = c.weight_nomask, wtgdl = c.wtgdl;
            if (wtgdl != 0.f) {
                double *gacc = t->loss_acc + CENN_LOSS_ERRG_GDL;
                double ngdl = (double)t->Bglobal * fake_in.C * fake_in.H * (fake_in.W - 1);
                emit(t, "gdl_loss", [s, fake_in, real, gacc, ngdl]() {
                    const int64_t pairs = (int64_t)fake_in.N * fake_in.H * (fake_in.W - 1);
                    if (fake_in.Cp == 16) LK(nhwc::gdl_loss_vec_kernel<16>, dim3(grid1d(s, pairs, 256, 4)), dim3(256), 0, s->stream)(fake_in.p, real.p, fake_in.N, fake_in.H, fake_in.W, fake_in.C, 1.0 / ngdl, gacc);
                    else if (fake_in.Cp == 4) LK(nhwc::gdl_loss_vec_kernel<4>, dim3(grid1d(s, pairs, 256, 4)), dim3(256), 0, s->stream)(fake_in.p, real.p, fake_in.N, fake_in.H, fake_in.W, fake_in.C, 1.0 / ngdl, gacc);
                    else
                    LK(nhwc::gdl_loss_kernel, dim3(grid1d(s, fake_in.elems(), 256, 4)), dim3(256), 0, s->stream)(fake_in.p, real.p, fake_in.N, fake_in.H, fake_in.W, fake_in.Cp, fake_in.C, 1.0 / ngdl, gacc);
                    KLAUNCH(s); return 0; });
            }
            emit(t, "blend_masked", [s, df, fake_in, real, mk, gout, a, wtl2, lam, wtgdl, n, acc]() {
                if (fake_in.Cp % 8 == 0)
                    LK(nhwc::blend_masked8_kernel, dim3(grid1d(s, fake_in.elems() / 8, 256, 4)), dim3(256), 0, s->stream)(df.p, fake_in.p, real.p, mk.p, gout.p, fake_in.elems() / 8, fake_in.Cp / 8, fake_in.C,
                        a, wtl2, lam, wtgdl, (float)(2.0 / n), 1.0 / n, acc);
                else
                LK(nhwc::blend_masked_kernel, dim3(grid1d(s, fake_in.elems(), 256, 4)), dim3(256), 0, s->stream)(df.p, fake_in.p, real.p, mk.p, gout.p, fake_in.elems(), fake_in.Cp, fake_in.C,
                    a, wtl2, lam, wtgdl, (float)(2.0 / n), 1.0 / n, acc);
                KLAUNCH(s); return 0; });
            t->prog.back().bytes = 5.0 * 2.0 * (double)fake_in.pix() * fake_in.C;   // read df, x, t, mask; write df
        }
    }
    emit_adam_step(t, G);          // optimState.t / step size first: the big blocks are updated as soon as their gradient exists
    t->g_early.clear();
    for (size_t i = G.blocks.size(); i-- > 0;) emit_backward(t, G, i, true, i > 0 || c.dead_dgrad);
    if (!t->g_early.empty()) {
        cudaEvent_t evj; cudaEventCreateWithFlags(&evj, cudaEventDisableTiming); t->events.push_back(evj);
        const bool dp_bulk = c.world_size > 1 && s->comm2 != nullptr;    // data parallel: the early Adams always run on the third stream
        emit(t, "join_adam", [t, s, evj, dp_bulk]() {
            if (t->serial && !dp_bulk) return 0;
            if (cenn_check_cuda(cudaEventRecord(evj, t->side3), "event record", __FILE__, __LINE__)) return 1;
            return cenn_check_cuda(cudaStreamWaitEvent(s->stream, evj, 0), "stream wait", __FILE__, __LINE__); });
    }
    emit_fold_gbias(t, G);
    if (t->g_buckets.empty()) emit(t, "gradG_sync", []() { return 0; }, G.grad, G.nparam);
    else {
        // the ranges not covered by a bucket go through the compute stream's communicator; then wait for the buckets
        std::vector<std::pair<int64_t, int64_t>> bk = t->g_buckets, rest;
        std::sort(bk.begin(), bk.end());
        int64_t cur = 0;
        for (auto &x : bk) { if (x.first > cur) rest.push_back({cur, x.first - cur}); cur = x.first + x.second; }
        if (G.nparam > cur) rest.push_back({cur, G.nparam - cur});
        cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); t->events.push_back(ev);
        float *g = G.grad;
        if (emit_peer_allreduce(t, "gradG_sync", G, rest, ev)) {} else
        emit(t, "gradG_sync", [s, rest, g, ev]() {
            // the bucket all-reduces (bulk communicator, comm_stream) finish first: two communicators are never in flight at once
            // (NCCL documents concurrent collectives on two communicators as a hang risk unless both kernels can be co-resident)
            if (cenn_check_cuda(cudaEventRecord(ev, s->comm_stream), "event record", __FILE__, __LINE__)) return 1;
            if (cenn_check_cuda(cudaStreamWaitEvent(s->stream, ev, 0), "stream wait", __FILE__, __LINE__)) return 1;
            if (cenn_dist_group(1)) return 1;
            for (auto &x : rest) if (cenn_dist_all_reduce_on(s, g + x.first, x.second, 0, s->stream)) return 1;
            return cenn_dist_group(0); });
    }
    emit_adam(t, G, t->g_early);
    emit_weight_prep(t, G);
    float wtl2 = c.wtl2, wtgdl = c.wtgdl;
    if (c.world_size > 1 && s->xr_enabled && getenv("CENN_NO_PEER_AR") == nullptr)      // the 8 accumulators cross the ranks through the mailboxes
        emit(t, "losses_sync", [t, s]() { nhwc::losses_xr_kernel<<<1, 32, 0, s->stream>>>(s->xr, t->loss_acc); KLAUNCH(s); return 0; });
    else
    emit(t, "losses_sync", []() { return 0; }, reinterpret_cast<float *>(t->loss_acc), 0 /* doubles: reduced separately */);
    emit(t, "finish_losses", [t, s, wtl2, wtgdl]() { finish_losses_kernel<<<1, 32, 0, s->stream>>>(t->loss_acc, t->loss_out, wtl2, wtgdl); KLAUNCH(s); return 0; });
    return 0;
}

// sum a sync point's buffer across the data-parallel ranks (library-owned NCCL communicator, same stream as the kernels)
int reduce_sync_point(T *t, const Op &op) {
    cenn_state *s = t->s;
    if (!op.sync_buf || t->cfg.world_size <= 1) return 0;
    if (!s->comm) { cenn_set_error("data-parallel step (world_size %d) without a communicator: call cenn_dist_init first, or drive the step with cenn_trainer_step_phase and reduce the sync points yourself", t->cfg.world_size); return 1; }
    const bool is_loss = op.sync_buf == reinterpret_cast<float *>(t->loss_acc);
    return cenn_dist_all_reduce_on(s, op.sync_buf, is_loss ? 8 : op.sync_count, is_loss ? 1 : 0, s->stream);
}
int run_ops(T *t, size_t from, size_t to) {
    cenn_state *s = t->s;
    static const bool sync_each = getenv("CENN_SYNC_EACH_OP") != nullptr;      // debugging: attribute an asynchronous fault to its op
    for (size_t i = from; i < to; ++i) {
        const Op &op = t->prog[i];
        cudaStream_t keep = s->stream;
        if (op.chain == 1 && !t->serial) s->stream = t->side2;
        int rc = op.fn() || reduce_sync_point(t, op);
        s->stream = keep;
        if (rc) return 1;
        if (sync_each) {
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { cenn_set_error("op %zu (%s) failed: %s", i, op.name, cudaGetErrorString(e)); return 1; }
        }
    }
    return 0;
}

static const int GRAPH_CACHE = 4;
int run_step(T *t) {
    cenn_state *s = t->s;
    static const bool no_graph = getenv("CENN_NO_GRAPH") != nullptr;
    if (no_graph || t->graph_failed) return run_ops(t, 0, t->prog.size());
    const void *key[3] = {t->cur_a, t->cur_b, t->cur_m};
    T::GraphEntry *ent = nullptr;
    for (auto &g : t->graphs) if (memcmp(key, g.key, sizeof(key)) == 0) { ent = &g; break; }
    if (ent && ent->exec) {
        ent->last_use = ++t->use_clock;
        CK(cudaGraphLaunch(ent->exec, s->stream));
        s->launches += ent->kernels;
        return 0;
    }
    // the first step with a given set of input buffers runs eagerly (it also warms lazily loaded kernels);
    // the second one is captured, every later one replays its graph
    if (!ent) {
        if ((int)t->graphs.size() >= GRAPH_CACHE) {        // evict the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < t->graphs.size(); ++i) if (t->graphs[i].last_use < t->graphs[lru].last_use) lru = i;
            if (t->graphs[lru].exec) cudaGraphExecDestroy(t->graphs[lru].exec);
            if (t->graphs[lru].graph) cudaGraphDestroy(t->graphs[lru].graph);
            t->graphs.erase(t->graphs.begin() + lru);
        }
        T::GraphEntry g = {};
        memcpy(g.key, key, sizeof(key)); g.seen = 1; g.last_use = ++t->use_clock;
        t->graphs.push_back(g);
        return run_ops(t, 0, t->prog.size());
    }
    int64_t before = s->launches;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError(); t->graph_failed = true;
        return run_ops(t, 0, t->prog.size());
    }
    int rc = run_ops(t, 0, t->prog.size());
    cudaError_t e = cudaStreamEndCapture(s->stream, &ent->graph);
    ent->kernels = s->launches - before;
    s->launches = before;
    if (rc || e != cudaSuccess || !ent->graph || cudaGraphInstantiate(&ent->exec, ent->graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (ent->graph) { cudaGraphDestroy(ent->graph); ent->graph = nullptr; }
        ent->exec = nullptr; t->graph_failed = true;
        fprintf(stderr, "cenn: CUDA graph capture / instantiation of the step failed (%s); the executor keeps launching eagerly\n", cudaGetErrorString(e));
        return run_ops(t, 0, t->prog.size());
    }
    ent->last_use = ++t->use_clock;
    CK(cudaGraphLaunch(ent->exec, s->stream));
    s->launches += ent->kernels;
    return 0;
}

void refresh_operands(T *t, Net &net) {
    cenn_state *s = t->s;
    nhwc::f32_to_bf16_kernel<<<grid1d(s, net.nparam), 256, 0, s->stream>>>(net.master, net.wbf, net.nparam);
    s->launches++;
    size_t mark = t->prog.size();
    emit_weight_prep(t, net);
    run_ops(t, mark, t->prog.size());
    t->prog.resize(mark);
}

}  // namespace

extern "C" {

int cenn_trainer_create(cenn_state *s, const cenn_trainer_config *cfg, cenn_trainer **out) {
    API_BEGIN(s);
    REQUIRE(cfg && out, "cenn_trainer_create: null argument");
    REQUIRE(cfg->precision == CENN_BF16, "the fused executor runs in BF16 tensor-core mode only; use the op-level modules for CENN_FP32");
    REQUIRE(cfg->variant == 0 || cfg->variant == 1, "unknown variant %d", cfg->variant);
    REQUIRE(cfg->fineSize == 128 || (cfg->fineSize == 256 && cfg->variant == 1),
            "fineSize must be 128, or 256 for the video / deeper nets (train_deepernet at 256 x 256: 5x5 bottleneck and patch head), got %d", cfg->fineSize);
    REQUIRE(cfg->batchSize >= 1 && cfg->nBottleneck % 8 == 0 && cfg->nef % 64 == 0 && cfg->ngf % 64 == 0 && cfg->ndf % 64 == 0,
            "batchSize >= 1, nBottleneck %% 8 == 0 and nef/ngf/ndf %% 64 == 0 required (got %d, %d, %d/%d/%d)", cfg->batchSize, cfg->nBottleneck, cfg->nef, cfg->ngf, cfg->ndf);
    REQUIRE(cfg->variant == 1 || cfg->overlapPred * 2 <= cfg->fineSize / 2, "overlapPred too large");
    REQUIRE(!(cfg->noiseGen || cfg->conditionAdv) || (cfg->variant == 0 && cfg->fineSize == 128),
            "noiseGen / conditionAdv are options of train.lua (image variant, fineSize 128); the video scripts force them off (train_deepernet.lua:55-58)");
    REQUIRE(!cfg->noiseGen || (cfg->nz >= 1 && cfg->nz <= 1024), "noiseGen: nz must be in 1..1024 (got %d)", cfg->nz);
    REQUIRE(cfg->variant == 0 || cfg->overlapPred == 0, "video variant requires overlapPred == 0 (train_vid_weighted.lua:509)");
    int ncv = cfg->variant == 1 ? cfg->nc * cfg->predLen : cfg->nc;
    REQUIRE(ncv >= 1 && ncv <= 16, "1..16 input channels supported (nc*predLen = %d)", ncv);
    cenn_trainer *t = new cenn_trainer();
    t->s = s; t->cfg = *cfg;
    if (t->cfg.world_size < 1) t->cfg.world_size = 1;
    t->B = cfg->batchSize; t->F = cfg->fineSize; t->nc = ncv;
    t->Bglobal = (int64_t)t->B * t->cfg.world_size;
    const bool video = cfg->variant == 1;
    const int dsize = video ? t->F : t->F / 2;
    int rc = 0;
    rc = rc || alloc_tensor(t, t->real_ctx, t->B, t->F, t->F, ncv, pad_thin(ncv));      // (conditionAdv's first D layer reads it directly)
    if (t->cfg.noiseGen) { t->noise = dalloc<float>(t, (int64_t)t->B * t->cfg.nz); rc = rc || !t->noise; }
    rc = rc || build_net(t, t->D, spec_D(*cfg), dsize, ncv, true, false);
    rc = rc || build_net(t, t->G, spec_G(*cfg), t->F, ncv, cfg->dead_dgrad != 0, true);
    if (rc) { cenn_trainer_destroy(t); return 1; }
    t->G.lr = (cfg->wtl2 > 0.f && cfg->wtl2 < 1.f) ? cfg->lr * 10.f : cfg->lr;   // train.lua:219-226
    t->D.lr = cfg->lr;
    rc = rc || alloc_tensor(t, t->real_aux, t->B, dsize, dsize, ncv, pad_thin(ncv));
    if (video) rc = rc || alloc_tensor(t, t->mask, t->B, t->F, t->F, ncv, pad_thin(ncv));
    t->n_a = (int64_t)t->B * ncv * t->F * t->F; t->n_b = (int64_t)t->B * ncv * dsize * dsize; t->n_m = video ? t->n_a : 0;
    t->in_a = dalloc<float>(t, t->n_a); t->in_b = dalloc<float>(t, t->n_b);
    if (video) t->in_m = dalloc<uint8_t>(t, t->n_m);
    t->loss_acc = dalloc<double>(t, 8); t->loss_out = dalloc<float>(t, 8);
    if (rc || !t->in_a || !t->in_b || !t->loss_acc || !t->loss_out) { cenn_trainer_destroy(t); return 1; }
    if (cudaMallocHost(&t->pin_loss, 8 * sizeof(float)) != cudaSuccess) {
        cenn_set_error("trainer: pinned host allocation failed"); cenn_trainer_destroy(t); return 1;
    }
    // G's output must match D's input tensor exactly (same NHWC padding) for the d2d hand-over
    const Tensor &go = t->G.blocks.back().a;
    if (go.H != dsize || go.Cp != t->D.input.Cp) { cenn_set_error("internal: generator output %dx%dx%d does not match discriminator input %dx%dx%d", go.H, go.W, go.Cp, dsize, dsize, t->D.input.Cp); cenn_trainer_destroy(t); return 1; }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);     // side streams: lowest priority (the critical path is on the state's stream)
    if (cudaStreamCreateWithPriority(&t->side3, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { cenn_set_error("trainer: side stream creation failed"); cenn_trainer_destroy(t); return 1; }
    if (cudaStreamCreateWithPriority(&t->side2, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { cenn_set_error("trainer: side stream creation failed"); cenn_trainer_destroy(t); return 1; }
    if (cudaStreamCreateWithPriority(&t->side, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { cenn_set_error("trainer: side stream creation failed"); cenn_trainer_destroy(t); return 1; }
    // data parallel: map the peers' generator buffers for the sharded reduce + Adam of the big blocks (collective: every rank takes the same
    // decisions -- same configuration and environment).  Any failure leaves the NCCL bucket path in place.
    if (t->cfg.world_size > 1 && s->comm && s->comm2 && s->xr_enabled && getenv("CENN_NO_PEER_AR") == nullptr) {
        bool ok = cenn_dist_ipc_map(s, t->G.base_grad, t->G.ipc_grad) == 0;
        ok = ok && cenn_dist_ipc_map(s, t->D.base_grad, t->D.ipc_grad) == 0;
        if (!ok) fprintf(stderr, "cenn: peer mapping of the gradient vectors failed (%s); leftover gradient ranges stay on NCCL\n", cenn_last_error());
        t->peer_ar_ok = ok;
    }
    if (t->peer_ar_ok && t->G.gradbf && getenv("CENN_NO_EARLY_ADAM") == nullptr && getenv("CENN_NO_SHARD_ADAM") == nullptr) {
        bool any = false;
        for (const Block &b : t->G.blocks) any = any || (b.w_count >= ((int64_t)1 << 23) && b.w_count % 8 == 0);
        if (any) {
            Net &n = t->G;
            bool ok = cenn_dist_ipc_map(s, n.gradbf, n.ipc_gradbf) == 0;
            ok = ok && cenn_dist_ipc_map(s, n.base_wbf, n.ipc_wbf) == 0;
            ok = ok && cenn_dist_ipc_map(s, n.base_master, n.ipc_master) == 0;
            if (!ok) { fprintf(stderr, "cenn: peer mapping of the generator buffers failed (%s); gradient buckets stay on NCCL\n", cenn_last_error()); }
            t->shard_ok = ok;
        }
    }
    if (build_program(t)) { cenn_trainer_destroy(t); return 1; }
    int64_t before = s->launches;
    (void)before;
    *out = t;
    return 0;
}

int cenn_trainer_destroy(cenn_trainer *t) {
    if (!t) return 0;
    cudaSetDevice(t->s->device);
    cudaStreamSynchronize(t->s->stream);
    for (auto &g : t->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); if (g.graph) cudaGraphDestroy(g.graph); }
    if (t->shard_ok || t->peer_ar_ok) {       // peers may still read / write this rank's buffers, and this rank theirs: everyone finishes, then everyone unmaps, then frees
        cudaDeviceSynchronize();
        cenn_dist_barrier(t->s);
        cenn_dist_ipc_unmap(t->s, t->G.ipc_gradbf); cenn_dist_ipc_unmap(t->s, t->G.ipc_wbf); cenn_dist_ipc_unmap(t->s, t->G.ipc_master); cenn_dist_ipc_unmap(t->s, t->G.ipc_grad);
        cenn_dist_ipc_unmap(t->s, t->D.ipc_grad);
        cenn_dist_barrier(t->s);
        t->shard_ok = false; t->peer_ar_ok = false;
    }
    if (t->copy_stream) { cudaStreamSynchronize(t->copy_stream); cudaStreamDestroy(t->copy_stream); }
    for (int i = 0; i < 2; ++i) { if (t->ev_copied[i]) cudaEventDestroy(t->ev_copied[i]); if (t->ev_consumed[i]) cudaEventDestroy(t->ev_consumed[i]); if (t->ev_loss[i]) cudaEventDestroy(t->ev_loss[i]); }
    if (t->pin_loss2) cudaFreeHost(t->pin_loss2);
    if (t->fr_u8) cudaFree(t->fr_u8);
    if (t->fr_mask) cudaFree(t->fr_mask);
    if (t->fr_tab) cudaFree(t->fr_tab);
    for (Net *n : {&t->G, &t->D})
        for (Block &b : n->blocks) { tc_plan_free(&b.p_fwd); tc_plan_free(&b.p_dgrad); tc_plan_free(&b.p_wgrad); tc_plan_free(&b.p_fwd2); tc_plan_free(&b.p_wgrad2); }
    if (t->side) { cudaStreamSynchronize(t->side); cudaStreamDestroy(t->side); }
    if (t->side2) { cudaStreamSynchronize(t->side2); cudaStreamDestroy(t->side2); }
    if (t->side3) { cudaStreamSynchronize(t->side3); cudaStreamDestroy(t->side3); }
    for (cudaEvent_t e : t->events) cudaEventDestroy(e);
    for (void *p : t->allocs) cudaFree(p);
    if (t->pin_a) cudaFreeHost(t->pin_a);
    if (t->pin_b) cudaFreeHost(t->pin_b);
    if (t->pin_m) cudaFreeHost(t->pin_m);
    if (t->pin_loss) cudaFreeHost(t->pin_loss);
    delete t;
    return 0;
}

static Net *pick_net(cenn_trainer *t, int net) { return net == CENN_NET_G ? &t->G : (net == CENN_NET_D ? &t->D : nullptr); }

int cenn_trainer_param_count(cenn_trainer *t, int net, int64_t *count) {
    REQUIRE(t && count && pick_net(t, net), "cenn_trainer_param_count: bad argument");
    *count = pick_net(t, net)->nparam_thnn;
    return 0;
}

// THNN flat layout <-> padded master layout (host side; not on the hot path)
static void thnn_to_master(const Net &n, const float *flat, std::vector<float> &m) {
    m.assign(n.nparam, 0.f);
    for (const Block &b : n.blocks) {
        if (b.type == JOIN5) {           // [Cs][nc][5][5] per branch -> [2*Cs][K5], k = 4 * tap + c; context rows first
            for (int br = 0; br < 2; ++br)
                for (int cs = 0; cs < b.Cs; ++cs) {
                    for (int cl = 0; cl < b.Cl; ++cl)
                        for (int tp = 0; tp < 25; ++tp)
                            m[b.w_off + ((int64_t)br * b.Cs + cs) * nhwc::K5 + 4 * tp + cl] = flat[(br ? b.t_w_off : b.t_w2_off) + ((int64_t)cs * b.Cl + cl) * 25 + tp];
                    m[b.b_off + br * b.Cs + cs] = flat[(br ? b.t_b_off : b.t_b2_off) + cs];
                }
            continue;
        }
        if (b.nz) {
            for (int64_t i = 0; i < (int64_t)b.nz * b.nz; ++i) m[b.nw_off + i] = flat[b.t_nw_off + i];
            for (int c = 0; c < b.nz; ++c) m[b.nb_off + c] = flat[b.t_nb_off + c];
        }
        for (int cs = 0; cs < b.Cs; ++cs)
            for (int cl = 0; cl < b.Cl; ++cl)
                for (int tp = 0; tp < 16; ++tp) m[b.w_off + ((int64_t)cs * 16 + tp) * b.Clp + cl] = flat[b.t_w_off + ((int64_t)cs * b.Cl + cl) * 16 + tp];
        for (int c = 0; c < b.nbias(); ++c) m[b.b_off + c] = flat[b.t_b_off + c];
        if (b.bn) for (int c = 0; c < b.Cout; ++c) { m[b.g_off + c] = flat[b.t_g_off + c]; m[b.be_off + c] = flat[b.t_be_off + c]; }
    }
}
static void master_to_thnn(const Net &n, const std::vector<float> &m, float *flat) {
    for (const Block &b : n.blocks) {
        if (b.type == JOIN5) {
            for (int br = 0; br < 2; ++br)
                for (int cs = 0; cs < b.Cs; ++cs) {
                    for (int cl = 0; cl < b.Cl; ++cl)
                        for (int tp = 0; tp < 25; ++tp)
                            flat[(br ? b.t_w_off : b.t_w2_off) + ((int64_t)cs * b.Cl + cl) * 25 + tp] = m[b.w_off + ((int64_t)br * b.Cs + cs) * nhwc::K5 + 4 * tp + cl];
                    flat[(br ? b.t_b_off : b.t_b2_off) + cs] = m[b.b_off + br * b.Cs + cs];
                }
            continue;
        }
        if (b.nz) {
            for (int64_t i = 0; i < (int64_t)b.nz * b.nz; ++i) flat[b.t_nw_off + i] = m[b.nw_off + i];
            for (int c = 0; c < b.nz; ++c) flat[b.t_nb_off + c] = m[b.nb_off + c];
        }
        for (int cs = 0; cs < b.Cs; ++cs)
            for (int cl = 0; cl < b.Cl; ++cl)
                for (int tp = 0; tp < 16; ++tp) flat[b.t_w_off + ((int64_t)cs * b.Cl + cl) * 16 + tp] = m[b.w_off + ((int64_t)cs * 16 + tp) * b.Clp + cl];
        for (int c = 0; c < b.nbias(); ++c) flat[b.t_b_off + c] = m[b.b_off + c];
        if (b.bn) for (int c = 0; c < b.Cout; ++c) { flat[b.t_g_off + c] = m[b.g_off + c]; flat[b.t_be_off + c] = m[b.be_off + c]; }
    }
}

int cenn_trainer_set_params_host(cenn_trainer *t, int net, const float *flat) {
    REQUIRE(t && flat && pick_net(t, net), "cenn_trainer_set_params_host: bad argument");
    API_BEGIN(t->s);
    Net &n = *pick_net(t, net);
    std::vector<float> m;
    thnn_to_master(n, flat, m);
    CK(cudaMemcpyAsync(n.master, m.data(), n.nparam * sizeof(float), cudaMemcpyHostToDevice, t->s->stream));
    CK(cudaStreamSynchronize(t->s->stream));
    refresh_operands(t, n);
    CK(cudaStreamSynchronize(t->s->stream));
    return 0;
}
static int get_vec(cenn_trainer *t, Net &n, const float *dev, float *flat) {
    std::vector<float> m(n.nparam);
    CK(cudaMemcpyAsync(m.data(), dev, n.nparam * sizeof(float), cudaMemcpyDeviceToHost, t->s->stream));
    CK(cudaStreamSynchronize(t->s->stream));
    if (t->shard_ok && !n.shard_blocks.empty() && (dev == n.master || dev == n.grad)) {
        // sharded blocks: the fp32 master copy / the reduced gradient of a shard lives on its owner -- read it from there (one-sided peer copy;
        // valid once this rank's step has completed: its closing barrier means every owner has finished its update)
        void *const *ipc = dev == n.master ? n.ipc_master : n.ipc_grad;
        const void *base = dev == n.master ? n.base_master : n.base_grad;
        const int world = t->cfg.world_size, rank = t->cfg.rank;
        for (const auto &sb : n.shard_blocks)
            for (int r = 0; r < world; ++r) {
                if (r == rank) continue;
                const int64_t b0 = shard_begin(sb.second, world, r), b1 = shard_begin(sb.second, world, r + 1);
                if (b1 > b0) CK(cudaMemcpy(m.data() + sb.first + b0, peer_ptr<float>(ipc, r, base, dev + sb.first + b0), (size_t)(b1 - b0) * sizeof(float), cudaMemcpyDefault));
            }
    }
    master_to_thnn(n, m, flat);
    return 0;
}
int cenn_trainer_get_params_host(cenn_trainer *t, int net, float *flat) {
    REQUIRE(t && flat && pick_net(t, net), "cenn_trainer_get_params_host: bad argument");
    API_BEGIN(t->s);
    return get_vec(t, *pick_net(t, net), pick_net(t, net)->master, flat);
}
int cenn_trainer_get_grads_host(cenn_trainer *t, int net, float *flat) {
    REQUIRE(t && flat && pick_net(t, net), "cenn_trainer_get_grads_host: bad argument");
    API_BEGIN(t->s);
    return get_vec(t, *pick_net(t, net), pick_net(t, net)->grad, flat);
}
int cenn_trainer_bn_stat_count(cenn_trainer *t, int net, int64_t *count) {
    REQUIRE(t && count && pick_net(t, net), "cenn_trainer_bn_stat_count: bad argument");
    int64_t n = 0;
    for (const Block &b : pick_net(t, net)->blocks) if (b.bn) n += 2 * b.Cout;
    *count = n;
    return 0;
}
int cenn_trainer_set_bn_stats_host(cenn_trainer *t, int net, const float *stats) {
    REQUIRE(t && stats && pick_net(t, net), "cenn_trainer_set_bn_stats_host: bad argument");
    API_BEGIN(t->s);
    int64_t off = 0;
    for (Block &b : pick_net(t, net)->blocks) if (b.bn) {
        CK(cudaMemcpy(b.running, stats + off, b.Cout * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b.running + b.Coutp, stats + off + b.Cout, b.Cout * sizeof(float), cudaMemcpyHostToDevice));
        off += 2 * b.Cout;
    }
    return 0;
}
int cenn_trainer_get_bn_stats_host(cenn_trainer *t, int net, float *stats) {
    REQUIRE(t && stats && pick_net(t, net), "cenn_trainer_get_bn_stats_host: bad argument");
    API_BEGIN(t->s);
    CK(cudaStreamSynchronize(t->s->stream));
    int64_t off = 0;
    for (Block &b : pick_net(t, net)->blocks) if (b.bn) {
        CK(cudaMemcpy(stats + off, b.running, b.Cout * sizeof(float), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(stats + off + b.Cout, b.running + b.Coutp, b.Cout * sizeof(float), cudaMemcpyDeviceToHost));
        off += 2 * b.Cout;
    }
    return 0;
}

int cenn_trainer_step_device(cenn_trainer *t, const float *a, const float *b, const uint8_t *mask) {
    REQUIRE(t && a && b, "cenn_trainer_step_device: null input");
    REQUIRE(t->cfg.variant == 0 || mask, "cenn_trainer_step_device: the video variant needs a mask");
    API_BEGIN(t->s);
    t->cur_a = a; t->cur_b = b; t->cur_m = mask;
    int64_t before = t->s->launches;
    int rc = run_step(t);
    t->launches_per_step = t->s->launches - before;
    return rc;
}
static int step_clips_device(cenn_trainer *t, const float *frames, const uint8_t *mask1, const uint8_t *flip, float maskValue) {
    if (maskValue != t->clip_mv) {        // the value is a kernel argument inside the captured graphs: a new value invalidates them
        cudaStreamSynchronize(t->s->stream);
        for (auto &g : t->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); if (g.graph) cudaGraphDestroy(g.graph); }
        t->graphs.clear();
        t->clip_mv = maskValue;
    }
    t->cur_a = nullptr; t->cur_b = frames; t->cur_m = mask1; t->cur_f = flip;
    int64_t before = t->s->launches;
    int rc = run_step(t);
    t->launches_per_step = t->s->launches - before;
    return rc;
}
static int set_noise(cenn_trainer *t, const float *noise, cudaMemcpyKind kind) {
    REQUIRE(t && noise, "cenn_trainer_set_noise: null argument");
    REQUIRE(t->noise, "cenn_trainer_set_noise: the executor was built without noiseGen");
    API_BEGIN(t->s);
    CK(cudaMemcpyAsync(t->noise, noise, (size_t)t->B * t->cfg.nz * sizeof(float), kind, t->s->stream));
    if (kind == cudaMemcpyHostToDevice) CK(cudaStreamSynchronize(t->s->stream));     // the host buffer may be reused on return
    return 0;
}
int cenn_trainer_set_noise_host(cenn_trainer *t, const float *noise) { return set_noise(t, noise, cudaMemcpyHostToDevice); }
int cenn_trainer_set_noise_device(cenn_trainer *t, const float *noise) { return set_noise(t, noise, cudaMemcpyDeviceToDevice); }
int cenn_trainer_read_losses(cenn_trainer *t, float *losses) {
    REQUIRE(t && losses, "cenn_trainer_read_losses: null argument");
    API_BEGIN(t->s);
    CK(cudaMemcpyAsync(t->pin_loss, t->loss_out, 8 * sizeof(float), cudaMemcpyDeviceToHost, t->s->stream));
    CK(cudaStreamSynchronize(t->s->stream));
    memcpy(losses, t->pin_loss, 8 * sizeof(float));
    return 0;
}
int cenn_trainer_step_host(cenn_trainer *t, const float *a, const float *b, const uint8_t *mask, float *losses) {
    REQUIRE(t && a && b && losses, "cenn_trainer_step_host: null argument");
    REQUIRE(t->cfg.variant == 0 || mask, "cenn_trainer_step_host: the video variant needs a mask");
    API_BEGIN(t->s);
    cudaStream_t st = t->s->stream;
    // host -> device copies are part of the end-to-end step (truly asynchronous when the caller's buffers are pinned,
    // e.g. from cenn_host_alloc; staged by the driver otherwise)
    CK(cudaMemcpyAsync(t->in_a, a, t->n_a * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->in_b, b, t->n_b * 4, cudaMemcpyHostToDevice, st));
    if (t->cfg.variant == 1) CK(cudaMemcpyAsync(t->in_m, mask, t->n_m, cudaMemcpyHostToDevice, st));
    if (cenn_trainer_step_device(t, t->in_a, t->in_b, t->in_m)) return 1;
    return cenn_trainer_read_losses(t, losses);
}

// Pipelined host-fed steps: step k's inputs are copied (copy stream, staging set k & 1) while step k-1 computes; losses
// are read one call later.  Host buffers must stay valid until the matching cenn_trainer_wait_losses returns.
int cenn_trainer_step_host_async(cenn_trainer *t, const float *a, const float *b, const uint8_t *mask) {
    REQUIRE(t && a && b, "cenn_trainer_step_host_async: null argument");
    REQUIRE(t->cfg.variant == 0 || mask, "cenn_trainer_step_host_async: the video variant needs a mask");
    API_BEGIN(t->s);
    REQUIRE(t->async_issued - t->async_read < 2, "cenn_trainer_step_host_async: two steps already in flight; call cenn_trainer_wait_losses");
    cudaStream_t st = t->s->stream;
    if (!t->copy_stream) {
        CK(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) { CK(cudaEventCreateWithFlags(&t->ev_copied[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&t->ev_consumed[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&t->ev_loss[i], cudaEventDisableTiming)); }
        t->in_a2 = dalloc<float>(t, t->n_a); t->in_b2 = dalloc<float>(t, t->n_b);
        if (t->cfg.variant == 1) t->in_m2 = dalloc<uint8_t>(t, t->n_m);
        REQUIRE(t->in_a2 && t->in_b2, "trainer: staging allocation failed");
        CK(cudaMallocHost(&t->pin_loss2, 2 * 8 * sizeof(float)));
    }
    const int i = (int)(t->async_issued & 1);
    float *da = i ? t->in_a2 : t->in_a, *db = i ? t->in_b2 : t->in_b; uint8_t *dm = i ? t->in_m2 : t->in_m;
    if (t->consumed_valid[i]) CK(cudaStreamWaitEvent(t->copy_stream, t->ev_consumed[i], 0));
    CK(cudaMemcpyAsync(da, a, t->n_a * 4, cudaMemcpyHostToDevice, t->copy_stream));
    CK(cudaMemcpyAsync(db, b, t->n_b * 4, cudaMemcpyHostToDevice, t->copy_stream));
    if (t->cfg.variant == 1) CK(cudaMemcpyAsync(dm, mask, t->n_m, cudaMemcpyHostToDevice, t->copy_stream));
    CK(cudaEventRecord(t->ev_copied[i], t->copy_stream));
    CK(cudaStreamWaitEvent(st, t->ev_copied[i], 0));
    if (cenn_trainer_step_device(t, da, db, dm)) return 1;
    CK(cudaEventRecord(t->ev_consumed[i], st)); t->consumed_valid[i] = true;
    CK(cudaMemcpyAsync(t->pin_loss2 + 8 * i, t->loss_out, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(t->ev_loss[i], st));
    t->async_issued++;
    return 0;
}
int cenn_trainer_wait_losses(cenn_trainer *t, float *losses) {
    REQUIRE(t && losses, "cenn_trainer_wait_losses: null argument");
    API_BEGIN(t->s);
    REQUIRE(t->async_read < t->async_issued, "cenn_trainer_wait_losses: no step in flight");
    const int i = (int)(t->async_read & 1);
    CK(cudaEventSynchronize(t->ev_loss[i]));
    memcpy(losses, t->pin_loss2 + 8 * i, 8 * sizeof(float));
    t->async_read++;
    return 0;
}

// Byte-image steps (image variant): the host hands over the loader's crop as DECODED BYTES [B,3,F,F] (what image.load yields before its
// division by 255); the device applies the loader's rescale to [-1,1] (data/donkey_folder.lua:84-86) and train.lua:286-290 (centre clone,
// mean fill).  A quarter of the H2D bytes of the FloatTensor form, and one tensor instead of two.
static int step_images_u8_device(cenn_trainer *t, const uint8_t *img_dev) {
    t->cur_a = nullptr; t->cur_b = reinterpret_cast<const float *>(img_dev); t->cur_m = nullptr;
    int64_t before = t->s->launches;
    int rc = run_step(t);
    t->launches_per_step = t->s->launches - before;
    return rc;
}
int cenn_trainer_step_images_u8_host(cenn_trainer *t, const uint8_t *images_u8, float *losses) {
    REQUIRE(t && images_u8 && losses, "cenn_trainer_step_images_u8_host: null argument");
    REQUIRE(t->cfg.variant == 0, "cenn_trainer_step_images_u8_host: byte images belong to the image variant");
    API_BEGIN(t->s);
    CK(cudaMemcpyAsync(t->in_a, images_u8, (size_t)t->n_a, cudaMemcpyHostToDevice, t->s->stream));      // n_a bytes: one byte per element of real_ctx
    if (step_images_u8_device(t, reinterpret_cast<const uint8_t *>(t->in_a))) return 1;
    return cenn_trainer_read_losses(t, losses);
}
int cenn_trainer_step_images_u8_host_async(cenn_trainer *t, const uint8_t *images_u8) {
    REQUIRE(t && images_u8, "cenn_trainer_step_images_u8_host_async: null argument");
    REQUIRE(t->cfg.variant == 0, "cenn_trainer_step_images_u8_host_async: byte images belong to the image variant");
    API_BEGIN(t->s);
    REQUIRE(t->async_issued - t->async_read < 2, "cenn_trainer_step_images_u8_host_async: two steps already in flight; call cenn_trainer_wait_losses");
    cudaStream_t st = t->s->stream;
    if (!t->copy_stream) {
        CK(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) { CK(cudaEventCreateWithFlags(&t->ev_copied[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&t->ev_consumed[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&t->ev_loss[i], cudaEventDisableTiming)); }
        t->in_a2 = dalloc<float>(t, t->n_a); t->in_b2 = dalloc<float>(t, t->n_b);
        REQUIRE(t->in_a2 && t->in_b2, "trainer: staging allocation failed");
        CK(cudaMallocHost(&t->pin_loss2, 2 * 8 * sizeof(float)));
    }
    const int i = (int)(t->async_issued & 1);
    float *da = i ? t->in_a2 : t->in_a;
    if (t->consumed_valid[i]) CK(cudaStreamWaitEvent(t->copy_stream, t->ev_consumed[i], 0));
    CK(cudaMemcpyAsync(da, images_u8, (size_t)t->n_a, cudaMemcpyHostToDevice, t->copy_stream));
    CK(cudaEventRecord(t->ev_copied[i], t->copy_stream));
    CK(cudaStreamWaitEvent(st, t->ev_copied[i], 0));
    if (step_images_u8_device(t, reinterpret_cast<const uint8_t *>(da))) return 1;
    CK(cudaEventRecord(t->ev_consumed[i], st)); t->consumed_valid[i] = true;
    CK(cudaMemcpyAsync(t->pin_loss2 + 8 * i, t->loss_out, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(t->ev_loss[i], st));
    t->async_issued++;
    return 0;
}

// Clip-mode steps (video variant): the host hands over what the loader has after its crop -- frames in [0,1], ONE mask plane per
// sample, hflip flags -- and the device derives real_full, real_ctx (masked fill) and the expanded mask
// (datavid/donkey_folder.lua:161-187).  Less than half the H2D bytes of the three-tensor form.
int cenn_trainer_step_clips_host(cenn_trainer *t, const float *frames01, const uint8_t *mask1, const uint8_t *flip, float maskValue, float *losses) {
    REQUIRE(t && frames01 && mask1 && losses, "cenn_trainer_step_clips_host: null argument");
    REQUIRE(t->cfg.variant == 1, "cenn_trainer_step_clips_host: clips belong to the video variant");
    API_BEGIN(t->s);
    cudaStream_t st = t->s->stream;
    if (!t->in_f) { t->in_f = dalloc<uint8_t>(t, t->B); REQUIRE(t->in_f, "trainer: staging allocation failed"); }
    CK(cudaMemcpyAsync(t->in_b, frames01, t->n_b * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->in_m, mask1, (size_t)t->B * t->F * t->F, cudaMemcpyHostToDevice, st));
    if (flip) CK(cudaMemcpyAsync(t->in_f, flip, t->B, cudaMemcpyHostToDevice, st)); else CK(cudaMemsetAsync(t->in_f, 0, t->B, st));
    if (step_clips_device(t, t->in_b, t->in_m, t->in_f, maskValue)) return 1;   // always a flag buffer: the pointer is baked into the captured graph
    return cenn_trainer_read_losses(t, losses);
}
int cenn_trainer_step_clips_host_async(cenn_trainer *t, const float *frames01, const uint8_t *mask1, const uint8_t *flip, float maskValue) {
    REQUIRE(t && frames01 && mask1, "cenn_trainer_step_clips_host_async: null argument");
    REQUIRE(t->cfg.variant == 1, "cenn_trainer_step_clips_host_async: clips belong to the video variant");
    API_BEGIN(t->s);
    REQUIRE(t->async_issued - t->async_read < 2, "cenn_trainer_step_clips_host_async: two steps already in flight; call cenn_trainer_wait_losses");
    cudaStream_t st = t->s->stream;
    if (!t->copy_stream) {
        CK(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) { CK(cudaEventCreateWithFlags(&t->ev_copied[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&t->ev_consumed[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&t->ev_loss[i], cudaEventDisableTiming)); }
        t->in_a2 = dalloc<float>(t, t->n_a); t->in_b2 = dalloc<float>(t, t->n_b);
        t->in_m2 = dalloc<uint8_t>(t, t->n_m);
        REQUIRE(t->in_a2 && t->in_b2 && t->in_m2, "trainer: staging allocation failed");
        CK(cudaMallocHost(&t->pin_loss2, 2 * 8 * sizeof(float)));
    }
    if (!t->in_f) { t->in_f = dalloc<uint8_t>(t, t->B); REQUIRE(t->in_f, "trainer: staging allocation failed"); }
    if (!t->in_f2) { t->in_f2 = dalloc<uint8_t>(t, t->B); REQUIRE(t->in_f2, "trainer: staging allocation failed"); }
    const int i = (int)(t->async_issued & 1);
    float *db = i ? t->in_b2 : t->in_b; uint8_t *dm = i ? t->in_m2 : t->in_m, *df = i ? t->in_f2 : t->in_f;
    if (t->consumed_valid[i]) CK(cudaStreamWaitEvent(t->copy_stream, t->ev_consumed[i], 0));
    CK(cudaMemcpyAsync(db, frames01, t->n_b * 4, cudaMemcpyHostToDevice, t->copy_stream));
    CK(cudaMemcpyAsync(dm, mask1, (size_t)t->B * t->F * t->F, cudaMemcpyHostToDevice, t->copy_stream));
    if (flip) CK(cudaMemcpyAsync(df, flip, t->B, cudaMemcpyHostToDevice, t->copy_stream)); else CK(cudaMemsetAsync(df, 0, t->B, t->copy_stream));
    CK(cudaEventRecord(t->ev_copied[i], t->copy_stream));
    CK(cudaStreamWaitEvent(st, t->ev_copied[i], 0));
    if (step_clips_device(t, db, dm, df, maskValue)) return 1;
    CK(cudaEventRecord(t->ev_consumed[i], st)); t->consumed_valid[i] = true;
    CK(cudaMemcpyAsync(t->pin_loss2 + 8 * i, t->loss_out, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(t->ev_loss[i], st));
    t->async_issued++;
    return 0;
}

// Frame-mode step (video variant): the host hands over what the loader holds BEFORE its per-sample hook -- decoded frames (bytes), the
// full-size logo mask and the hook's random draws (crop origin, hflip, random-block corners) -- and the device runs the hook
// (datavid/donkey_folder.lua:138-187): crop, mask crop + expand, maskedFill or randomBlockMask, hflip, [0,1] -> [-1,1].
int cenn_trainer_step_frames_host(cenn_trainer *t, const uint8_t *frames_u8, int iH, int iW, const uint8_t *mask_full, const int *crop,
                                  const uint8_t *flip, const int *blocks, float maskValue, float *losses) {
    REQUIRE(t && frames_u8 && mask_full && crop && blocks && losses, "cenn_trainer_step_frames_host: null argument");
    REQUIRE(t->cfg.variant == 1, "cenn_trainer_step_frames_host: frames belong to the video variant");
    const int F = t->F, B = t->B, Cc = t->nc;
    REQUIRE(iH >= F && iW >= F, "cenn_trainer_step_frames_host: frames (%dx%d) smaller than fineSize %d", iH, iW, F);
    for (int n = 0; n < B; ++n) {
        REQUIRE(crop[2 * n] >= 0 && crop[2 * n] + F <= iH && crop[2 * n + 1] >= 0 && crop[2 * n + 1] + F <= iW, "crop %d outside the frame", n);
        REQUIRE(blocks[21 * n] >= 0 && blocks[21 * n] <= 10, "sample %d: at most 10 random blocks (donkey_folder.lua:120)", n);
    }
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    cudaStream_t st = s->stream;
    const size_t n_u8 = (size_t)B * Cc * iH * iW, n_mask = (size_t)iH * iW;
    if (n_u8 > t->fr_u8_cap) { if (t->fr_u8) cudaFree(t->fr_u8); t->fr_u8 = nullptr; CK(cudaMalloc(&t->fr_u8, n_u8)); t->fr_u8_cap = n_u8; }
    if (n_mask > t->fr_mask_cap) { if (t->fr_mask) cudaFree(t->fr_mask); t->fr_mask = nullptr; CK(cudaMalloc(&t->fr_mask, n_mask)); t->fr_mask_cap = n_mask; }
    if (!t->fr_tab) CK(cudaMalloc(&t->fr_tab, (size_t)B * (2 + 21 + 1) * sizeof(int)));
    if (!t->in_f) { t->in_f = dalloc<uint8_t>(t, B); REQUIRE(t->in_f, "trainer: staging allocation failed"); }
    int *d_crop = t->fr_tab, *d_blocks = t->fr_tab + 2 * B, *d_any = t->fr_tab + 23 * B;
    CK(cudaMemcpyAsync(t->fr_u8, frames_u8, n_u8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->fr_mask, mask_full, n_mask, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_crop, crop, (size_t)B * 2 * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_blocks, blocks, (size_t)B * 21 * sizeof(int), cudaMemcpyHostToDevice, st));
    if (flip) CK(cudaMemcpyAsync(t->in_f, flip, B, cudaMemcpyHostToDevice, st)); else CK(cudaMemsetAsync(t->in_f, 0, B, st));
    nhwc::crop_mask_any_kernel<<<B, 256, 0, st>>>(t->fr_mask, d_crop, iW, F, d_any);
    KLAUNCH(s);
    nhwc::frames_to_clips_kernel<<<grid1d(s, (int64_t)B * F * F), 256, 0, st>>>(t->fr_u8, t->fr_mask, d_crop, d_blocks, d_any, B, Cc, iH, iW, F, t->in_b, t->in_m);
    KLAUNCH(s);
    if (step_clips_device(t, t->in_b, t->in_m, t->in_f, maskValue)) return 1;
    return cenn_trainer_read_losses(t, losses);
}

// One step with a CUDA-event pair around every op of the program (on the launching stream); returns the op names
// ('\n'-separated), their durations and algorithmic FLOPs.  Used by bench.py for the roofline of the dominant kernels.
int cenn_trainer_profile_step(cenn_trainer *t, const float *a, const float *b, const uint8_t *mask, char *names, int64_t names_cap,
        float *ms, double *flops, int64_t cap, int64_t *nops) {
    REQUIRE(t && a && b && names && ms && flops && nops, "cenn_trainer_profile_step: null argument");
    API_BEGIN(t->s);
    cudaStream_t st = t->s->stream;
    size_t n = t->prog.size();
    REQUIRE((int64_t)n <= cap, "cenn_trainer_profile_step: capacity %lld < %zu ops", (long long)cap, n);
    std::vector<cudaEvent_t> ev(n + 1, nullptr);
    struct Guard {      // every exit path: events destroyed, the executor back on its streams
        cenn_trainer *t; std::vector<cudaEvent_t> &ev;
        ~Guard() { t->serial = false; for (auto &e : ev) if (e) cudaEventDestroy(e); }
    } guard{t, ev};
    for (auto &e : ev) CK(cudaEventCreate(&e));
    t->cur_a = a; t->cur_b = b; t->cur_m = mask;
    static const bool timeline = getenv("CENN_TIMELINE") != nullptr;   // keep the side streams: events then show the MAIN stream's time line
    t->serial = !timeline;            // per-op timing: no side stream
    CK(cudaEventRecord(ev[0], st));
    int rc = 0;
    for (size_t i = 0; i < n && !rc; ++i) {
        cudaStream_t keep = t->s->stream;
        if (t->prog[i].chain == 1 && !t->serial) t->s->stream = t->side2;
        rc = t->prog[i].fn() || reduce_sync_point(t, t->prog[i]);
        t->s->stream = keep;
        CK(cudaEventRecord(ev[i + 1], st));
    }
    CK(cudaStreamSynchronize(st));
    if (rc) return 1;
    std::string all;
    for (size_t i = 0; i < n; ++i) {
        CK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
        flops[i] = t->prog[i].flops;
        all += t->prog[i].name; all += '\n';
    }
    REQUIRE((int64_t)all.size() + 1 <= names_cap, "cenn_trainer_profile_step: names buffer too small");
    memcpy(names, all.c_str(), all.size() + 1);
    *nops = (int64_t)n;
    return 0;
}

// One EAGER step on the executor's real streams with an event pair around every op on the stream that op launches on:
// start / end of every op in milliseconds since the step began -> which chain is the critical path, where the streams idle.
int cenn_trainer_timeline_step(cenn_trainer *t, const float *a, const float *b, const uint8_t *mask, float *t_start, float *t_end, int *stream_id, int64_t cap, int64_t *nops) {
    REQUIRE(t && a && b && t_start && t_end && stream_id && nops, "cenn_trainer_timeline_step: null argument");
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    cudaStream_t st = s->stream;
    const size_t n = t->prog.size();
    REQUIRE((int64_t)n <= cap, "cenn_trainer_timeline_step: capacity %lld < %zu ops", (long long)cap, n);
    std::vector<cudaEvent_t> ev(2 * n + 1, nullptr);
    struct Guard { std::vector<cudaEvent_t> &ev; ~Guard() { for (auto &e : ev) if (e) cudaEventDestroy(e); } } guard{ev};
    for (auto &e : ev) CK(cudaEventCreate(&e));
    t->cur_a = a; t->cur_b = b; t->cur_m = mask;
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(ev[2 * n], st));
    // the side streams start behind the base event so that every elapsed time is non-negative
    for (cudaStream_t x : {t->side, t->side2, t->side3}) CK(cudaStreamWaitEvent(x, ev[2 * n], 0));
    if (s->comm_stream) CK(cudaStreamWaitEvent(s->comm_stream, ev[2 * n], 0));
    for (size_t i = 0; i < n; ++i) {
        const Op &op = t->prog[i];
        // ops that hop to a side stream themselves (wgrad, dead dgrad, early Adam; data parallel: the bucket all-reduces on the bulk
        // communicator's stream, id 4) are bracketed on that stream
        const bool on_side = !strcmp(op.name, "wgrad") || (!strcmp(op.name, "dgrad") && false);
        const bool on_side3 = !strcmp(op.name, "adam_early");
        const bool on_comm = (!strcmp(op.name, "grad_bucket_ar") || !strcmp(op.name, "grad_shard_cvt") || !strcmp(op.name, "shard_adam")) && s->comm_stream;
        cudaStream_t run = op.chain == 1 ? t->side2 : st;
        cudaStream_t where = on_side ? t->side : (on_side3 ? t->side3 : (on_comm ? s->comm_stream : run));
        stream_id[i] = where == st ? 0 : (where == t->side ? 1 : (where == t->side2 ? 2 : (where == t->side3 ? 3 : 4)));
        cudaStream_t keep = s->stream;
        s->stream = run;
        if (where == run) CK(cudaEventRecord(ev[2 * i], where));
        int rc = op.fn() || reduce_sync_point(t, op);
        s->stream = keep;
        if (rc) return 1;
        if (where != run) CK(cudaEventRecord(ev[2 * i], where));     // start unknown on a hop: recorded after the launch, start := previous end on that stream
        CK(cudaEventRecord(ev[2 * i + 1], where));
    }
    CK(cudaDeviceSynchronize());
    float last_end[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t i = 0; i < n; ++i) {
        float a0 = 0.f, a1 = 0.f;
        CK(cudaEventElapsedTime(&a1, ev[2 * n], ev[2 * i + 1]));
        const Op &op = t->prog[i];
        const bool hop = !strcmp(op.name, "wgrad") || !strcmp(op.name, "adam_early") || stream_id[i] == 4;
        if (hop) a0 = last_end[stream_id[i]]; else CK(cudaEventElapsedTime(&a0, ev[2 * n], ev[2 * i]));
        t_start[i] = a0; t_end[i] = a1;
        last_end[stream_id[i]] = a1;
    }
    *nops = (int64_t)n;
    return 0;
}

// Test / debugging hook: run the step program from its first op up to and including the `occurrence`-th (0-based) op named
// `op_name`, serially on the compute stream, then synchronise -- the stored tensors and gradient vectors can then be fetched
// mid-step (e.g. after the discriminator's real sweep, before the fake sweep overwrites its activations).
int cenn_trainer_step_until(cenn_trainer *t, const float *a, const float *b, const uint8_t *mask, const char *op_name, int occurrence, int64_t *ops_run) {
    REQUIRE(t && a && b && op_name, "cenn_trainer_step_until: null argument");
    API_BEGIN(t->s);
    long idx = -1; int seen = 0;
    for (size_t i = 0; i < t->prog.size(); ++i)
        if (!strcmp(t->prog[i].name, op_name) && seen++ == occurrence) { idx = (long)i; break; }
    REQUIRE(idx >= 0, "cenn_trainer_step_until: the step program has no op '%s' #%d", op_name, occurrence);
    t->cur_a = a; t->cur_b = b; t->cur_m = mask;
    t->serial = true;
    int rc = 0;
    for (long i = 0; i <= idx && !rc; ++i) rc = t->prog[i].fn() || reduce_sync_point(t, t->prog[i]);
    t->serial = false;
    if (rc) return 1;
    CK(cudaStreamSynchronize(t->s->stream));
    if (ops_run) *ops_run = idx + 1;
    return 0;
}

int cenn_trainer_op_bytes(cenn_trainer *t, double *bytes, int64_t cap, int64_t *nops) {
    REQUIRE(t && bytes && nops, "cenn_trainer_op_bytes: null argument");
    REQUIRE((int64_t)t->prog.size() <= cap, "cenn_trainer_op_bytes: capacity %lld < %zu ops", (long long)cap, t->prog.size());
    for (size_t i = 0; i < t->prog.size(); ++i) bytes[i] = t->prog[i].bytes;
    *nops = (int64_t)t->prog.size();
    return 0;
}

int cenn_trainer_grad_buffer(cenn_trainer *t, int net, float **grads, int64_t *count) {
    REQUIRE(t && grads && count && pick_net(t, net), "cenn_trainer_grad_buffer: bad argument");
    *grads = pick_net(t, net)->grad; *count = pick_net(t, net)->nparam;
    return 0;
}
// phase < 0: (re)start a step with the given inputs and run up to the first sync point; phase >= 0: continue.
// After the call, cenn_trainer_grad_buffer-style queries are replaced by the sync record returned through
// cenn_last_error-free out-params of cenn_trainer_sync_info (see DESIGN.md "multi-GPU").
int cenn_trainer_step_phase(cenn_trainer *t, int phase, const float *a, const float *b, const uint8_t *mask) {
    REQUIRE(t, "cenn_trainer_step_phase: null trainer");
    API_BEGIN(t->s);
    t->serial = true;
    if (phase < 0) { REQUIRE(a && b, "cenn_trainer_step_phase: null input"); t->cur_a = a; t->cur_b = b; t->cur_m = mask; t->pc = 0; t->last_sync = -1; }
    while (t->pc < t->prog.size()) {
        Op &op = t->prog[t->pc++];
        if (op.fn()) { t->serial = false; return 1; }
        if (op.sync_buf) { t->last_sync = (long)t->pc - 1; t->serial = false; return 0; }
    }
    t->last_sync = -1;
    t->serial = false;
    return 0;
}
int cenn_trainer_sync_info(cenn_trainer *t, void **buf, int64_t *count, int *is_double, int *done) {
    REQUIRE(t && buf && count && is_double && done, "cenn_trainer_sync_info: null argument");
    *done = t->pc >= t->prog.size() && t->last_sync < 0;
    *buf = nullptr; *count = 0; *is_double = 0;
    if (t->last_sync >= 0) {
        const Op &op = t->prog[t->last_sync];
        *buf = op.sync_buf;
        if (op.sync_buf == reinterpret_cast<float *>(t->loss_acc)) { *count = 8; *is_double = 1; } else *count = op.sync_count;
    }
    return 0;
}

int cenn_trainer_generator_forward_host(cenn_trainer *t, const float *in, float *out, int batch) {
    REQUIRE(t && in && out, "cenn_trainer_generator_forward_host: null argument");
    REQUIRE(batch >= 1 && batch <= t->B, "generator_forward: batch %d outside 1..%d (the executor's buffers are sized for batchSize)", batch, t->B);
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    Net &G = t->G;
    const Tensor &gi = G.input;
    int64_t n_in = (int64_t)batch * gi.C * gi.H * gi.W;
    CK(cudaMemsetAsync(t->in_a, 0, t->n_a * 4, s->stream));
    CK(cudaMemcpyAsync(t->in_a, in, n_in * 4, cudaMemcpyHostToDevice, s->stream));
    LK(nhwc::to_nhwc_kernel<float>, dim3(grid1d(s, gi.pix())), dim3(256), 0, s->stream)(t->in_a, gi.p, gi.N, gi.C, gi.H * gi.W, gi.Cp);
    KLAUNCH(s);
    size_t mark = t->prog.size();
    for (size_t i = 0; i < G.blocks.size(); ++i) emit_forward(t, G, i, false);
    double f = t->flops_per_step;
    int rc = run_ops(t, mark, t->prog.size());
    t->prog.resize(mark);
    t->flops_per_step = f;
    if (rc) return 1;
    const Tensor &go = G.blocks.back().a;
    float *tmp = (float *)cenn_workspace(s, (size_t)go.N * go.C * go.H * go.W * 4);
    if (!tmp) return 1;
    nhwc::to_nchw_kernel<<<grid1d(s, (int64_t)go.N * go.C * go.H * go.W), 256, 0, s->stream>>>(go.p, tmp, go.N, go.C, go.H * go.W, go.Cp);
    KLAUNCH(s);
    CK(cudaMemcpyAsync(out, tmp, (size_t)batch * go.C * go.H * go.W * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

// name = "<G|D>.<block index>.<y|a|g|in>" or "df_dg" / "fake" / "ctx"
int cenn_trainer_fetch_host(cenn_trainer *t, const char *name, float *dst, int64_t capacity, int64_t *count) {
    REQUIRE(t && name && dst && count, "cenn_trainer_fetch_host: null argument");
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    Tensor x;
    std::string nm(name);
    if (nm == "df_dg") x = t->df_dg;
    else if (nm == "fake") x = t->G.blocks.back().a;
    else if (nm == "ctx") x = t->real_ctx;
    else {
        REQUIRE(nm.size() >= 5 && (nm[0] == 'G' || nm[0] == 'D') && nm[1] == '.', "fetch: bad name '%s'", name);
        Net &n = nm[0] == 'G' ? t->G : t->D;
        size_t dot = nm.find('.', 2);
        REQUIRE(dot != std::string::npos, "fetch: bad name '%s'", name);
        int idx = atoi(nm.substr(2, dot - 2).c_str());
        REQUIRE(idx >= 0 && idx < (int)n.blocks.size(), "fetch: block index out of range in '%s'", name);
        std::string f = nm.substr(dot + 1);
        Block &b = n.blocks[idx];
        if (f == "y") x = b.y; else if (f == "a") x = b.a; else if (f == "g") x = b.g; else if (f == "in") x = b.in;
        else if (f == "sig") { *count = b.Mrows; REQUIRE(capacity >= b.Mrows && b.sig, "fetch: no sig"); CK(cudaStreamSynchronize(s->stream)); CK(cudaMemcpy(dst, b.sig, (size_t)b.Mrows * 4, cudaMemcpyDeviceToHost)); return 0; }
        else { cenn_set_error("fetch: unknown field in '%s'", name); return 1; }
    }
    REQUIRE(x.p, "fetch: '%s' has no buffer", name);
    int64_t n = (int64_t)x.N * x.C * x.H * x.W;
    REQUIRE(capacity >= n, "fetch: capacity %lld < %lld", (long long)capacity, (long long)n);
    float *tmp = (float *)cenn_workspace(s, n * 4);
    if (!tmp) return 1;
    nhwc::to_nchw_kernel<<<grid1d(s, n), 256, 0, s->stream>>>(x.p, tmp, x.N, x.C, x.H * x.W, x.Cp);
    KLAUNCH(s);
    CK(cudaMemcpyAsync(dst, tmp, n * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *count = n;
    return 0;
}

int cenn_trainer_kernel_launches_per_step(cenn_trainer *t, int64_t *count) {
    REQUIRE(t && count, "null argument");
    *count = t->launches_per_step;
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// inference engine: eval-mode generator with BN folded into the operands, full-frame sweep on the device
// ------------------------------------------------------------------------------------------------------------------
struct cenn_inpainter {
    cenn_trainer *t = nullptr;          // generator only, t->infer
    cenn_inpainter_config cfg;
    int ncin = 0;
    bool loaded = false;
    float *stage_in = nullptr, *stage_out = nullptr;   // fp32 NCHW staging of forward_host
    // sweep buffers (grown on demand)
    float *frames = nullptr, *img[3] = {nullptr, nullptr, nullptr};
    uint8_t *mask = nullptr;
    size_t frames_cap = 0, img_cap = 0, mask_cap = 0;
};

namespace {

int inpainter_build_program(cenn_inpainter *p) {
    T *t = p->t;
    cenn_state *s = t->s;
    Net &G = t->G;
    t->prog.clear();
    // cur_a / cur_b: fp32 NCHW device input / output (nullptr: the tiles are already in / stay in the NHWC buffers)
    emit(t, "tiles<-nchw", [t, s]() {
        if (!t->cur_a) return 0;
        const Tensor &gi = t->G.input;
        const int n = t->infer_n;
        if (gi.Cp == 4 && (gi.H * gi.W) % 4 == 0) LK(nhwc::to_nhwc4_kernel, dim3(grid1d(s, (int64_t)n * gi.H * gi.W / 4)), dim3(256), 0, s->stream)(t->cur_a, gi.p, n, gi.C, gi.H * gi.W);
        else LK(nhwc::to_nhwc_kernel<float>, dim3(grid1d(s, (int64_t)n * gi.H * gi.W)), dim3(256), 0, s->stream)(t->cur_a, gi.p, n, gi.C, gi.H * gi.W, gi.Cp);
        KLAUNCH(s); return 0; });
    for (size_t i = 0; i < G.blocks.size(); ++i) {
        Block *b = &G.blocks[i];
        if (b->thin && b->type == CONV_S2) emit_im2col(t, b->in, b->col, b->h, b->w);
        emit_plan(t, "conv_fwd", &b->p_fwd);
    }
    emit(t, "tiles->nchw", [t, s]() {
        if (!t->cur_b) return 0;
        const Tensor &go = t->G.blocks.back().a;
        const int n = t->infer_n;
        nhwc::to_nchw_kernel<<<grid1d(s, (int64_t)n * go.C * go.H * go.W), 256, 0, s->stream>>>(go.p, const_cast<float *>(t->cur_b), n, go.C, go.H * go.W, go.Cp);
        KLAUNCH(s); return 0; });
    return 0;
}

int inpainter_run(cenn_inpainter *p, const float *in_dev, float *out_dev, int n) {
    T *t = p->t;
    t->cur_a = in_dev; t->cur_b = out_dev; t->infer_n = n;
    t->cur_m = reinterpret_cast<const uint8_t *>((uintptr_t)n);     // part of the graph-cache key
    return run_step(t);
}

template <typename X>
int grow(X *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    if (cudaMalloc(&ptr, need * sizeof(X)) != cudaSuccess) { cenn_set_error("inpainter: device allocation of %zu bytes failed", need * sizeof(X)); return 1; }
    cap = need;
    return 0;
}

}  // namespace

extern "C" {

int cenn_inpainter_create(cenn_state *s, const cenn_inpainter_config *cfg, cenn_inpainter **out) {
    API_BEGIN(s);
    REQUIRE(cfg && out, "cenn_inpainter_create: null argument");
    REQUIRE(cfg->variant == 0 || cfg->variant == 1, "unknown variant %d", cfg->variant);
    REQUIRE(cfg->fineSize == 128, "fineSize must be 128 (the 4x4 bottleneck of the reference generators), got %d", cfg->fineSize);
    REQUIRE(cfg->batch >= 1 && cfg->nBottleneck % 8 == 0 && cfg->nBottleneck >= 8 && cfg->nef % 64 == 0 && cfg->ngf % 64 == 0 && cfg->nef >= 64 && cfg->ngf >= 64,
            "batch >= 1, nBottleneck %% 8 == 0 and nef/ngf %% 64 == 0 required (got %d, %d, %d/%d)", cfg->batch, cfg->nBottleneck, cfg->nef, cfg->ngf);
    const int ncin = cfg->variant == 1 ? cfg->nc * cfg->inputLen : cfg->nc;
    REQUIRE(cfg->nc >= 1 && cfg->inputLen >= 1 && ncin <= 16, "1..16 input channels supported (nc*inputLen = %d)", ncin);
    cenn_inpainter *p = new cenn_inpainter();
    p->cfg = *cfg; p->ncin = ncin;
    cenn_trainer *t = new cenn_trainer();
    p->t = t;
    t->s = s; t->infer = true;
    cenn_trainer_config tc = {};
    tc.variant = cfg->variant; tc.batchSize = cfg->batch; tc.fineSize = cfg->fineSize; tc.nBottleneck = cfg->nBottleneck;
    tc.nef = cfg->nef; tc.ngf = cfg->ngf; tc.ndf = 64; tc.nc = cfg->nc; tc.predLen = cfg->inputLen; tc.precision = CENN_BF16; tc.world_size = 1;
    t->cfg = tc;
    t->B = cfg->batch; t->F = cfg->fineSize; t->nc = ncin; t->Bglobal = t->B;
    if (build_net(t, t->G, spec_G(tc), t->F, ncin, false, true) || inpainter_build_program(p)) { cenn_inpainter_destroy(p); return 1; }
    *out = p;
    return 0;
}

int cenn_inpainter_destroy(cenn_inpainter *p) {
    if (!p) return 0;
    if (p->t) {
        cudaSetDevice(p->t->s->device);
        cudaStreamSynchronize(p->t->s->stream);
        for (void *q : {(void *)p->stage_in, (void *)p->stage_out, (void *)p->frames, (void *)p->img[0], (void *)p->img[1], (void *)p->img[2], (void *)p->mask}) if (q) cudaFree(q);
        cenn_trainer_destroy(p->t);
    }
    delete p;
    return 0;
}

int cenn_inpainter_param_count(cenn_inpainter *p, int64_t *params, int64_t *bn_stats) {
    REQUIRE(p && p->t, "cenn_inpainter_param_count: null argument");
    if (params) *params = p->t->G.nparam_thnn;
    if (bn_stats) return cenn_trainer_bn_stat_count(p->t, CENN_NET_G, bn_stats);
    return 0;
}

int cenn_inpainter_load_host(cenn_inpainter *p, const float *flat, const float *bn_stats) {
    REQUIRE(p && p->t && flat && bn_stats, "cenn_inpainter_load_host: null argument");
    T *t = p->t;
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    Net &n = t->G;
    std::vector<float> m;
    thnn_to_master(n, flat, m);
    CK(cudaMemcpyAsync(n.master, m.data(), n.nparam * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (cenn_trainer_set_bn_stats_host(t, CENN_NET_G, bn_stats)) return 1;
    nhwc::f32_to_bf16_kernel<<<grid1d(s, n.nparam), 256, 0, s->stream>>>(n.master, n.wbf, n.nparam);
    KLAUNCH(s);
    for (Block &b : n.blocks) {
        const float *gamma = b.bn ? n.master + b.g_off : nullptr, *beta = b.bn ? n.master + b.be_off : nullptr;
        const float *rm = b.bn ? b.running : nullptr, *rv = b.bn ? b.running + b.Coutp : nullptr;
        const bool chan_is_row = b.type == CONV_S2 || b.type == CONV_V4;     // output channel = small side = operand row
        if (b.bn) {
            nhwc::fold_bn_weights_kernel<<<grid1d(s, b.w_count), 256, 0, s->stream>>>(n.master + b.w_off, n.wbf + b.w_off, b.w_count, b.Clp, chan_is_row ? 1 : 0, gamma, rv, b.Cout, 1e-5);
            KLAUNCH(s);
        }
        const int Cp = b.type == FULL_V4 ? b.Clp : b.Coutp, reps = b.type == FULL_V4 ? 16 : 1;
        nhwc::fold_bn_bias_kernel<<<(reps * Cp + 255) / 256, 256, 0, s->stream>>>(n.master + b.b_off, gamma, beta, rm, rv, b.bias_inf, b.Cout, Cp, reps, 1e-5);
        KLAUNCH(s);
    }
    size_t mark = t->prog.size();
    emit_weight_prep(t, n);                                  // transposed operand copies of the thin layers, from the folded bf16 copy
    int rc = run_ops(t, mark, t->prog.size());
    t->prog.resize(mark);
    if (rc) return 1;
    CK(cudaStreamSynchronize(s->stream));
    p->loaded = true;
    return 0;
}

int cenn_inpainter_forward_device(cenn_inpainter *p, const float *in, float *out, int n) {
    REQUIRE(p && p->t && in && out, "cenn_inpainter_forward_device: null argument");
    REQUIRE(p->loaded, "cenn_inpainter_forward: no parameters loaded (cenn_inpainter_load_host)");
    REQUIRE(n >= 1 && n <= p->t->B, "cenn_inpainter_forward: %d tiles outside 1..%d (the engine's batch)", n, p->t->B);
    API_BEGIN(p->t->s);
    return inpainter_run(p, in, out, n);
}

int cenn_inpainter_forward_host(cenn_inpainter *p, const float *in, float *out, int n) {
    REQUIRE(p && p->t && in && out, "cenn_inpainter_forward_host: null argument");
    REQUIRE(p->loaded, "cenn_inpainter_forward: no parameters loaded (cenn_inpainter_load_host)");
    T *t = p->t;
    REQUIRE(n >= 1 && n <= t->B, "cenn_inpainter_forward: %d tiles outside 1..%d (the engine's batch)", n, t->B);
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    const Tensor &gi = t->G.input, &go = t->G.blocks.back().a;
    const size_t per_in = (size_t)gi.C * gi.H * gi.W, per_out = (size_t)go.C * go.H * go.W;
    if (!p->stage_in) {
        if (cudaMalloc(&p->stage_in, per_in * t->B * 4) != cudaSuccess || cudaMalloc(&p->stage_out, per_out * t->B * 4) != cudaSuccess) { cenn_set_error("inpainter: staging allocation failed"); return 1; }
    }
    CK(cudaMemcpyAsync(p->stage_in, in, per_in * n * 4, cudaMemcpyHostToDevice, s->stream));
    if (inpainter_run(p, p->stage_in, p->stage_out, n)) return 1;
    CK(cudaMemcpyAsync(out, p->stage_out, per_out * n * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int cenn_inpainter_sweep_host(cenn_inpainter *p, cenn_inpainter *init, const float *frames01, const uint8_t *mask, int P, int inh, int inw,
                              float maskValue, float *out01, float *full01, float *inpaint01) {
    REQUIRE(p && p->t && frames01 && mask, "cenn_inpainter_sweep_host: null argument");
    REQUIRE(p->loaded && (!init || init->loaded), "cenn_inpainter_sweep: no parameters loaded (cenn_inpainter_load_host)");
    REQUIRE(p->cfg.variant == 1, "the full-frame sweep needs a generator whose output has the input's size (variant 1)");
    REQUIRE(P >= 1 && P % p->cfg.inputLen == 0, "I don't do padding in time dim: %d frames, inputLen %d (test_vid_wholeim.lua:41)", P, p->cfg.inputLen);
    REQUIRE(inh >= 1 && inw >= 1, "empty frames");
    T *t = p->t;
    if (init) REQUIRE(init->t->s == t->s && init->ncin == p->ncin && init->t->B == t->B && init->cfg.variant == 1, "initializer net must share the state, channel count and batch of the main engine");
    API_BEGIN(t->s);
    cenn_state *s = t->s;
    const int F = t->F, nc = p->cfg.nc;
    nhwc::SweepGeom q;
    q.P = P; q.nc = nc; q.inh = inh; q.inw = inw; q.F = F; q.ncin = p->ncin; q.Cp = t->G.input.Cp;
    q.outh = (inh + F - 1) / F * F; q.outw = (inw + F - 1) / F * F;
    q.groups = P * nc / p->ncin; q.tiles_w = q.outw / F;
    const int total_tiles = (q.outh / F) * q.tiles_w * q.groups;
    const size_t n_frames = (size_t)P * nc * inh * inw, n_img = (size_t)P * nc * q.outh * q.outw, n_mask = (size_t)inh * inw;
    if (grow(p->frames, p->frames_cap, n_frames) || grow(p->mask, p->mask_cap, n_mask)) return 1;
    if (n_img > p->img_cap) { size_t c0 = p->img_cap, c1 = p->img_cap, c2 = p->img_cap; if (grow(p->img[0], c0, n_img) || grow(p->img[1], c1, n_img) || grow(p->img[2], c2, n_img)) return 1; p->img_cap = n_img; }
    CK(cudaMemcpyAsync(p->frames, frames01, n_frames * 4, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(p->mask, mask, n_mask, cudaMemcpyHostToDevice, s->stream));
    const Tensor &gi = t->G.input, &go = t->G.blocks.back().a;
    REQUIRE(go.H == F && go.Cp == gi.Cp, "internal: generator output %dx%dx%d does not match its input", go.H, go.W, go.Cp);
    for (int j0 = 0; j0 < total_tiles; j0 += t->B) {
        const int n = std::min(t->B, total_tiles - j0);
        const int64_t px = (int64_t)n * F * F;
        const bf16 *mid = nullptr;
        if (init) {
            const Tensor &ii = init->t->G.input;
            if (q.Cp == 4) nhwc::wholeim_gather_kernel<4><<<grid1d(s, px), 256, 0, s->stream>>>(p->frames, p->mask, maskValue, q, j0, n, ii.p, nullptr);
            else nhwc::wholeim_gather_kernel<16><<<grid1d(s, px), 256, 0, s->stream>>>(p->frames, p->mask, maskValue, q, j0, n, ii.p, nullptr);
            KLAUNCH(s);
            if (inpainter_run(init, nullptr, nullptr, n)) return 1;
            mid = init->t->G.blocks.back().a.p;
        }
        if (q.Cp == 4) nhwc::wholeim_gather_kernel<4><<<grid1d(s, px), 256, 0, s->stream>>>(p->frames, p->mask, maskValue, q, j0, n, gi.p, mid);
        else nhwc::wholeim_gather_kernel<16><<<grid1d(s, px), 256, 0, s->stream>>>(p->frames, p->mask, maskValue, q, j0, n, gi.p, mid);
        KLAUNCH(s);
        if (inpainter_run(p, nullptr, nullptr, n)) return 1;
        nhwc::wholeim_scatter_kernel<<<grid1d(s, px), 256, 0, s->stream>>>(go.p, p->frames, p->mask, maskValue, q, j0, n, p->img[0], p->img[1], p->img[2]);
        KLAUNCH(s);
    }
    float *dst[3] = {out01, full01, inpaint01};
    for (int k = 0; k < 3; ++k) if (dst[k]) CK(cudaMemcpyAsync(dst[k], p->img[k], n_img * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

}  // extern "C"
