/*
 * cenn.h -- C ABI of libcenn.so, the B200-native (sm_100a) replacement for what the
 * reference's Torch7 nn modules call *under* their Lua methods: the per-tensor-type
 * THNN / THCUNN function table (THNN_Cuda<Op>_updateOutput / _updateGradInput /
 * _accGradParameters) plus the cutorch tensor calls the scripts make on GPU tensors.
 *
 * Boundary: every nn module method of the reference forwards to
 *   input.THNN.<Op>_<phase>(input:cdata(), output:cdata(), ...)
 * e.g. nn.SpatialConvolution (train.lua:80,89-104), nn.SpatialFullConvolution
 * (train.lua:81,134-146), nn.SpatialBatchNormalization (train.lua:79), nn.LeakyReLU /
 * nn.ReLU / nn.Tanh / nn.Sigmoid (train.lua:90,135,147,197), nn.BCECriterion /
 * nn.MSECriterion (train.lua:207-210), and the repo-local nn.MaskedMSECriterion
 * (MaskedMSECriterion.lua:4-42) and nn.GDLCriterion (gdl_criterion.lua:4-53).
 * The THNN signatures are the ones restated in SURVEY.md 9.11; THCudaTensor* arguments
 * become (device pointer, explicit sizes) because this ABI carries no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; cenn_last_error()
 *     returns a thread-local message.  Nothing aborts the process (THError would longjmp).
 *   - all tensor pointers are DEVICE pointers to contiguous fp32 NCHW data unless the
 *     name ends in _host.  "[opt]" pointers may be NULL.
 *   - ops run on the state's current stream; scalar-returning criterion forwards are the
 *     only calls that synchronise (they mirror criterion:forward returning a Lua number).
 */
#ifndef CENN_H
#define CENN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CENN_API __attribute__((visibility("default")))
#else
#define CENN_API
#endif

typedef struct cenn_state cenn_state;     /* replaces THCState* (cutorch.getState()) */
typedef struct cenn_trainer cenn_trainer; /* whole-step executor, SURVEY.md 8f rank 1   */

/* precision modes (north_star: fp32 parity mode <= 1e-5, BF16 tensor-core mode <= 2e-2) */
enum { CENN_FP32 = 0, CENN_BF16 = 1 };

/* ---------------------------------------------------------------- runtime ------------ */
/* replaces `require 'cunn'; cutorch.setDevice(opt.gpu)` (train.lua:249-250) */
CENN_API int cenn_init(int device, cenn_state **out);
CENN_API int cenn_shutdown(cenn_state *s);
CENN_API const char *cenn_last_error(void);
CENN_API const char *cenn_version(void);
CENN_API int cenn_device_count(int *count);
CENN_API int cenn_set_precision(cenn_state *s, int mode);
CENN_API int cenn_get_precision(cenn_state *s, int *mode);
CENN_API int cenn_set_stream(cenn_state *s, void *cuda_stream); /* NULL = library's own stream */
CENN_API int cenn_get_stream(cenn_state *s, void **cuda_stream);
CENN_API int cenn_synchronize(cenn_state *s);                   /* cutorch.synchronize() */
CENN_API int cenn_kernel_launches(cenn_state *s, int64_t *count); /* kernels launched so far */

/* ------------------------------------------------- data parallelism (one process per GPU) -- */
/* The reference is single-device (cutorch.setDevice(opt.gpu), train.lua:250); these calls are what a multi-GPU launcher
 * adds: rank 0 creates an id, every rank receives it out of band (file, env, torch.distributed broadcast) and joins.
 * Once a communicator exists, a cenn_trainer created with world_size > 1 all-reduces its BN statistics, gradients and
 * loss accumulators itself (NCCL over NVLink, enqueued on the state's stream and captured in the step's CUDA graph). */
CENN_API int cenn_dist_unique_id(void *id128_host);                       /* rank 0: 128-byte ncclUniqueId */
CENN_API int cenn_dist_init(cenn_state *s, const void *id128_host, int world_size, int rank);
CENN_API int cenn_dist_all_reduce(cenn_state *s, void *buf_dev, int64_t count, int is_double);   /* in-place sum, fp32 / fp64 */
CENN_API int cenn_dist_broadcast(cenn_state *s, void *buf_dev, int64_t bytes, int root);
CENN_API int cenn_dist_shutdown(cenn_state *s);

/* ------------------------------------------------- storage (tensor:cuda(), :float()) -- */
CENN_API int cenn_malloc(cenn_state *s, size_t bytes, void **dptr);
CENN_API int cenn_free(cenn_state *s, void *dptr);
CENN_API int cenn_host_alloc(cenn_state *s, size_t bytes, void **hptr); /* pinned */
CENN_API int cenn_host_free(cenn_state *s, void *hptr);
CENN_API int cenn_copy_h2d(cenn_state *s, void *dst, const void *src_host, size_t bytes);
CENN_API int cenn_copy_d2h(cenn_state *s, void *dst_host, const void *src, size_t bytes);
CENN_API int cenn_copy_d2d(cenn_state *s, void *dst, const void *src, size_t bytes);

/* -------------------------- tensor math the scripts call on GPU tensors -- */
CENN_API int cenn_fill(cenn_state *s, float *x, int64_t n, float v);                      /* x:fill(v)/zero() */
CENN_API int cenn_mul(cenn_state *s, float *x, int64_t n, float a);                       /* x:mul(a) */
CENN_API int cenn_add_scalar(cenn_state *s, float *x, int64_t n, float a);                /* x:add(a) */
CENN_API int cenn_axpy(cenn_state *s, float *y, const float *x, int64_t n, float a);      /* y:add(a, x) */
CENN_API int cenn_cmul(cenn_state *s, float *y, const float *x, int64_t n);               /* y:cmul(x) */
CENN_API int cenn_addcmul(cenn_state *s, float *y, float a, const float *p, const float *q, int64_t n);
CENN_API int cenn_addcdiv(cenn_state *s, float *y, float a, const float *p, const float *q, int64_t n);
CENN_API int cenn_sqrt(cenn_state *s, float *x, int64_t n);
CENN_API int cenn_u8_to_float(cenn_state *s, float *dst, const uint8_t *src, int64_t n);  /* input_mask:copy(byte) */
CENN_API int cenn_masked_fill(cenn_state *s, float *x, const float *mask, int64_t n, float v);
/* fill a [N,C,H,W] sub-box x[:, c0:c1, y0:y1, x0:x1] = v  (train.lua:288-290, :392) */
CENN_API int cenn_fill_box(cenn_state *s, float *x, int64_t N, int64_t C, int64_t H, int64_t W,
                           int64_t c0, int64_t c1, int64_t y0, int64_t y1, int64_t x0, int64_t x1, float v);
/* dst = src[:, :, y0:y0+h, x0:x0+w] cloned contiguous (train.lua:287 centre crop) */
CENN_API int cenn_crop(cenn_state *s, float *dst, const float *src, int64_t N, int64_t C, int64_t H, int64_t W,
                       int64_t y0, int64_t x0, int64_t h, int64_t w);
/* nn.JoinTable(2) on batch-mode tensors (train.lua:120,177; noiseGen / conditionAdv): each sample of `joined` is
 * `joined_per_sample` contiguous floats, of which [offset, offset + part_per_sample) belong to `part`.
 * updateOutput copies part -> joined, updateGradInput copies the matching narrow of gradOutput -> gradPart. */
CENN_API int cenn_JoinTable_updateOutput(cenn_state *s, float *joined, const float *part, int64_t batch,
        int64_t joined_per_sample, int64_t offset, int64_t part_per_sample);
CENN_API int cenn_JoinTable_updateGradInput(cenn_state *s, const float *gradJoined, float *gradPart, int64_t batch,
        int64_t joined_per_sample, int64_t offset, int64_t part_per_sample);
/* Philox-based :normal(mean,std) / :uniform(a,b) (train.lua:61,64,269-272) */
CENN_API int cenn_normal(cenn_state *s, float *x, int64_t n, float mean, float std, uint64_t seed);
CENN_API int cenn_uniform(cenn_state *s, float *x, int64_t n, float a, float b, uint64_t seed);

/* ------------------------------------------------ THNN function table: convolutions --- */
/* THNN_CudaSpatialConvolutionMM_updateOutput(state,input,output,weight,bias,columns,ones,kW,kH,dW,dH,padW,padH)
 * weight [nOutputPlane, nInputPlane, kH, kW]; called by nn.SpatialConvolution:updateOutput (train.lua:89-104,183-196).
 * The columns/ones scratch tensors of THNN are owned by the library. */
CENN_API int cenn_SpatialConvolutionMM_updateOutput(cenn_state *s, const float *input, float *output,
        const float *weight, const float *bias /*opt*/, int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW,
        int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH);
CENN_API int cenn_SpatialConvolutionMM_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput,
        const float *weight, int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW,
        int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH);
/* gradWeight += scale * ..., gradBias += scale * ... (accumulating) */
CENN_API int cenn_SpatialConvolutionMM_accGradParameters(cenn_state *s, const float *input, const float *gradOutput,
        float *gradWeight, float *gradBias /*opt*/, int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW,
        int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, float scale);

/* THNN_CudaSpatialFullConvolution_* ; weight [nInputPlane, nOutputPlane, kH, kW] (train.lua:134-146) */
CENN_API int cenn_SpatialFullConvolution_updateOutput(cenn_state *s, const float *input, float *output,
        const float *weight, const float *bias /*opt*/, int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW,
        int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH);
CENN_API int cenn_SpatialFullConvolution_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput,
        const float *weight, int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW,
        int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH);
CENN_API int cenn_SpatialFullConvolution_accGradParameters(cenn_state *s, const float *input, const float *gradOutput,
        float *gradWeight, float *gradBias /*opt*/, int64_t batch, int64_t nInputPlane, int64_t inH, int64_t inW,
        int64_t nOutputPlane, int kW, int kH, int dW, int dH, int padW, int padH, int adjW, int adjH, float scale);

/* ------------------------------------------------ THNN: BatchNormalization ------------ */
/* THNN_CudaBatchNormalization_updateOutput(state,input,output,weight,bias,runningMean,runningVar,saveMean,saveStd,
 *   train,momentum,eps); input viewed as [batch, C, spatial]; saveStd holds invstd */
CENN_API int cenn_BatchNormalization_updateOutput(cenn_state *s, const float *input, float *output,
        const float *weight /*opt*/, const float *bias /*opt*/, float *runningMean, float *runningVar,
        float *saveMean, float *saveStd, int64_t batch, int64_t C, int64_t spatial, int train, double momentum, double eps);
CENN_API int cenn_BatchNormalization_backward(cenn_state *s, const float *input, const float *gradOutput,
        float *gradInput /*opt*/, float *gradWeight /*opt*/, float *gradBias /*opt*/, const float *weight /*opt*/,
        const float *runningMean, const float *runningVar, const float *saveMean, const float *saveStd,
        int64_t batch, int64_t C, int64_t spatial, int train, double scale, double eps);

/* ------------------------------------------------ THNN: activations ------------------- */
CENN_API int cenn_LeakyReLU_updateOutput(cenn_state *s, const float *input, float *output, int64_t n, double negval, int inplace);
CENN_API int cenn_LeakyReLU_updateGradInput(cenn_state *s, const float *input, const float *gradOutput, float *gradInput,
        int64_t n, double negval, int inplace);
CENN_API int cenn_Threshold_updateOutput(cenn_state *s, const float *input, float *output, int64_t n, double threshold,
        double val, int inplace);
CENN_API int cenn_Threshold_updateGradInput(cenn_state *s, const float *input, const float *gradOutput, float *gradInput,
        int64_t n, double threshold, int inplace);
CENN_API int cenn_Tanh_updateOutput(cenn_state *s, const float *input, float *output, int64_t n);
CENN_API int cenn_Tanh_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *output, int64_t n);
CENN_API int cenn_Sigmoid_updateOutput(cenn_state *s, const float *input, float *output, int64_t n);
CENN_API int cenn_Sigmoid_updateGradInput(cenn_state *s, const float *gradOutput, float *gradInput, const float *output, int64_t n);
CENN_API int cenn_Abs_updateOutput(cenn_state *s, const float *input, float *output, int64_t n);
CENN_API int cenn_Abs_updateGradInput(cenn_state *s, const float *input, const float *gradOutput, float *gradInput, int64_t n);
CENN_API int cenn_Square_updateOutput(cenn_state *s, const float *input, float *output, int64_t n);
CENN_API int cenn_Square_updateGradInput(cenn_state *s, const float *input, const float *gradOutput, float *gradInput, int64_t n);

/* ------------------------------------------------ THNN: criteria ---------------------- */
/* forward writes the loss to *loss_host (host float) and synchronises, like criterion:forward() */
CENN_API int cenn_BCECriterion_updateOutput(cenn_state *s, const float *input, const float *target, int64_t n,
        int sizeAverage, float *loss_host);
CENN_API int cenn_BCECriterion_updateGradInput(cenn_state *s, const float *input, const float *target, float *gradInput,
        int64_t n, int sizeAverage);
CENN_API int cenn_MSECriterion_updateOutput(cenn_state *s, const float *input, const float *target, int64_t n,
        int sizeAverage, float *loss_host);
CENN_API int cenn_MSECriterion_updateGradInput(cenn_state *s, const float *input, const float *target, float *gradInput,
        int64_t n, int sizeAverage);
CENN_API int cenn_AbsCriterion_updateOutput(cenn_state *s, const float *input, const float *target, int64_t n,
        int sizeAverage, float *loss_host);
CENN_API int cenn_AbsCriterion_updateGradInput(cenn_state *s, const float *input, const float *target, float *gradInput,
        int64_t n, int sizeAverage);

/* ------------------------------ fused entry points for the repo-local Lua compositions -- */
/* nn.MaskedMSECriterion (MaskedMSECriterion.lua:29-42): loss = sum(((1-mW)M+mW)(x-t)^2)/n, grad = 2 wM (x-t)/n.
 * mask is float {0,1} (the module keeps m:double(); values are identical).  gradInput [opt] */
CENN_API int cenn_MaskedMSECriterion_forward_backward(cenn_state *s, const float *input, const float *target,
        const float *mask, float *gradInput /*opt*/, int64_t n, double mWeight, float *loss_host /*opt*/);
/* nn.GDLCriterion(1) (gdl_criterion.lua:38-52) with the flat-index pairing of SURVEY 9.8; requires H == W */
CENN_API int cenn_GDLCriterion_forward_backward(cenn_state *s, const float *input, const float *target,
        float *gradInput /*opt*/, int64_t batch, int64_t C, int64_t H, int64_t W, float *loss_host /*opt*/);
/* train.lua:377-400: errG_l2 = MSE(x,t); df_dg = a*df_dg + Wm .* 2(x-t)/n with Wm = 10*wtl2 on the
 * overlapPred border ring and wtl2 inside; a = (0<wtl2<1) ? 1-wtl2 : 1.  In place on df_dg. */
CENN_API int cenn_WeightedMSEBlend_overlap(cenn_state *s, float *df_dg, const float *input, const float *target,
        int64_t batch, int64_t C, int64_t H, int64_t W, double wtl2, int overlapPred, float *errG_l2_host /*opt*/);
/* train_vid_weighted.lua:485-507 (+ :523-528 when wtgdl != 0): weights = mask*(1-lambda)+lambda written IN PLACE
 * over mask (skipped when lambda == 0); df_dg = a*df_dg + (wtl2*weights + wtgdl) .* 2(x-t)/n. */
CENN_API int cenn_WeightedMSEBlend_masked(cenn_state *s, float *df_dg, const float *input, const float *target,
        float *mask_inout, int64_t n, double wtl2, double weight_nomask, double wtgdl, float *errG_l2_host /*opt*/);
/* inpaint_utils.fillIn / train_vid_weighted.lua:429-432: dst = where(mask, src, dst) */
CENN_API int cenn_MaskComposite(cenn_state *s, float *dst, const float *mask, const float *src, int64_t n);
/* optim.adam over a flat vector; t is the 1-based step count AFTER increment */
CENN_API int cenn_AdamFlat(cenn_state *s, float *x, const float *g, float *m, float *v, int64_t n,
        double lr, double beta1, double beta2, double eps, int64_t t);

/* ------------------------------ whole-step executor ------------------ */
typedef struct cenn_trainer_config {
    int variant;       /* 0: train.lua (inpaintCenter image)   1: train_vid_weighted.lua / train_deepernet.lua */
    int batchSize;     /* per-process (per-GPU) batch */
    int fineSize;      /* 128 */
    int nBottleneck, nef, ngf, ndf;
    int nc;            /* channels per frame (3) */
    int predLen;       /* frames per clip; video nets use nc*predLen channels */
    int overlapPred;
    float wtl2, weight_nomask, wtgdl;
    float lr, beta1;
    int precision;     /* CENN_FP32 / CENN_BF16 */
    int world_size;    /* data-parallel replicas; batch statistics / criteria use batchSize*world_size */
    int rank;
    int dead_dgrad;    /* 1: also compute the first-layer dgrads the reference computes and discards */
    /* train.lua's optional branches (image variant, fineSize 128; the video scripts force both off, train_deepernet.lua:55-58) */
    int noiseGen;      /* train.lua:109-124: a 1x1 conv of a noise vector [B,nz,1,1] is joined to the bottleneck (cenn_trainer_set_noise_*) */
    int nz;            /* length of the noise vector (train.lua:11, default 100) */
    int conditionAdv;  /* train.lua:158-180: netD takes {context, prediction}: two 5x5/stride-2 first-layer convs joined along channels */
    /* data parallel (world_size > 1) only.  0 (default): BN batch statistics are those of the GLOBAL batch, exchanged across the ranks inside the
     * step -- N ranks at batchSize equal one executor at N*batchSize (tests/test_dp_multi_gpu.py).  1: every rank normalises with its OWN
     * batchSize samples, as one reference process would (the usual distributed-data-parallel semantics): no statistics cross the ranks, only
     * gradients and losses; running statistics are per rank (save rank 0's). */
    int bn_local;
} cenn_trainer_config;

enum { CENN_NET_G = 0, CENN_NET_D = 1 };
enum { CENN_LOSS_ERRD = 0, CENN_LOSS_ERRG = 1, CENN_LOSS_ERRG_L2 = 2, CENN_LOSS_ERRG_GDL = 3,
       CENN_LOSS_ERRD_REAL = 4, CENN_LOSS_ERRD_FAKE = 5, CENN_LOSS_ERRG_TOTAL = 6, CENN_LOSS_COUNT = 8 };

CENN_API int cenn_trainer_create(cenn_state *s, const cenn_trainer_config *cfg, cenn_trainer **out);
CENN_API int cenn_trainer_destroy(cenn_trainer *t);
/* flat parameter vectors in Module:getParameters order, THNN layouts (train.lua:262-263) */
CENN_API int cenn_trainer_param_count(cenn_trainer *t, int net, int64_t *count);
CENN_API int cenn_trainer_set_params_host(cenn_trainer *t, int net, const float *flat_host);
CENN_API int cenn_trainer_get_params_host(cenn_trainer *t, int net, float *flat_host);
CENN_API int cenn_trainer_get_grads_host(cenn_trainer *t, int net, float *flat_host);
/* BN running statistics, concatenated per BN layer in module order: [running_mean(C), running_var(C)]... */
CENN_API int cenn_trainer_bn_stat_count(cenn_trainer *t, int net, int64_t *count);
CENN_API int cenn_trainer_set_bn_stats_host(cenn_trainer *t, int net, const float *stats_host);
CENN_API int cenn_trainer_get_bn_stats_host(cenn_trainer *t, int net, float *stats_host);
/* One optim.adam(fDx)+optim.adam(fGx) step (train.lua:421-424).  Host inputs are fp32 NCHW:
 *   image variant: a = real_ctx [B,3,F,F] (centre already mean-filled), b = real_center [B,3,F/2,F/2], mask = NULL
 *   video variant: a = real_ctx (masked) [B,nc*predLen,F,F], b = real_full, mask = uint8 same shape
 * Copies inputs H2D, runs the step, copies the CENN_LOSS_COUNT losses back. */
CENN_API int cenn_trainer_step_host(cenn_trainer *t, const float *a_host, const float *b_host,
        const uint8_t *mask_host, float *losses_host /*[CENN_LOSS_COUNT]*/);
/* pipelined form: enqueue step k (H2D of its inputs on a copy stream overlaps step k-1), read the losses one call later.
 * At most two steps may be in flight; host buffers (pinned for true overlap, cenn_host_alloc) must stay valid until the
 * matching cenn_trainer_wait_losses returns.  Results are identical to cenn_trainer_step_host. */
CENN_API int cenn_trainer_step_host_async(cenn_trainer *t, const float *a_host, const float *b_host, const uint8_t *mask_host);
CENN_API int cenn_trainer_wait_losses(cenn_trainer *t, float *losses_host /*[CENN_LOSS_COUNT]*/);
/* same with inputs already resident in HBM (fp32 NCHW device pointers, mask as uint8 device pointer);
 * losses stay on the device until cenn_trainer_read_losses */
CENN_API int cenn_trainer_step_device(cenn_trainer *t, const float *a_dev, const float *b_dev, const uint8_t *mask_dev);
CENN_API int cenn_trainer_read_losses(cenn_trainer *t, float *losses_host);
/* noiseGen: the noise draw of the NEXT step(s), fp32 [batchSize, nz] (train.lua:319-323 redraws `noise` inside fDx; here the caller
 * draws it -- noise:uniform(-1,1) / noise:normal(0,1) -- and hands it over before each step call).  Stream-ordered copy into the
 * executor's own buffer: the source may be reused as soon as the call returns (host form) / the copy has run (device form). */
CENN_API int cenn_trainer_set_noise_host(cenn_trainer *t, const float *noise_host);
CENN_API int cenn_trainer_set_noise_device(cenn_trainer *t, const float *noise_dev);
/* phase-split form for data parallelism: the host inserts all-reduces between phases
 * (phase list and the buffers to reduce are described in DESIGN.md section "multi-GPU") */
CENN_API int cenn_trainer_grad_buffer(cenn_trainer *t, int net, float **grads_dev, int64_t *count);
/* phase < 0 starts a step with the given device inputs; phase >= 0 continues it.  Each call runs up to and
 * including the next synchronisation point and returns; cenn_trainer_sync_info then names the device buffer
 * (fp32, or 8 doubles for the loss accumulators) that must be summed across ranks before the next call,
 * or reports done = 1 when the step has finished. */
CENN_API int cenn_trainer_step_phase(cenn_trainer *t, int phase, const float *a_dev, const float *b_dev,
        const uint8_t *mask_dev);
CENN_API int cenn_trainer_sync_info(cenn_trainer *t, void **buf_dev, int64_t *count, int *is_double, int *done);
/* Clip-mode steps of the video variant: the per-sample hook of the video loader after its crop
 * (datavid/donkey_folder.lua:161-187) runs on the device.  frames01 [B, nc*predLen, F, F] in [0,1]; mask1 [B, F, F], ONE plane
 * per sample (non-zero = hole; the loader expands it over the channels); flip [B] (opt, non-zero = image.hflip of frames and
 * mask).  The device derives real_full = 2f-1, real_ctx = maskedFill(maskValue) and the expanded mask -- less than half the
 * host-to-device bytes of cenn_trainer_step_host.  The async form pairs with cenn_trainer_wait_losses. */
CENN_API int cenn_trainer_step_clips_host(cenn_trainer *t, const float *frames01_host, const uint8_t *mask1_host,
        const uint8_t *flip_host, float maskValue, float *losses_host);
CENN_API int cenn_trainer_step_clips_host_async(cenn_trainer *t, const float *frames01_host, const uint8_t *mask1_host,
        const uint8_t *flip_host, float maskValue);
/* Frame mode (remainder): the per-sample hook of the video loader (datavid/donkey_folder.lua:114-129,138-187) on the
 * device.  frames_u8 [B][nc*predLen][iH][iW] are the decoded frames (image.load = byte / 255), mask_full [iH][iW] the logo mask;
 * the hook's RANDOM DRAWS stay on the host and arrive as tables: crop[b] = (h1, w1) 0-based top-left of the fineSize crop (:149-151),
 * flip[b] (:172), blocks[b] = (count <= 10, tlx_0, tly_0, ...) 0-based inside the crop, block side floor(fineSize / 6), used when the
 * mask crop is all black (:114-129,163-168).  The sample-rejection rule (:152-157) is loader policy and stays with the caller. */
/* Byte-image steps (image variant): images_u8 [B][3][fineSize][fineSize] = the loader's crop as decoded bytes (image.load yields byte / 255,
 * data/donkey_folder.lua:70-88 then rescales to [-1,1]); the device rescales, clones the centre (real_center) and mean-fills it
 * (real_ctx), train.lua:286-290.  12.6 MB of H2D per 256-sample step instead of 62.9 MB.  The async form pipelines like step_host_async. */
CENN_API int cenn_trainer_step_images_u8_host(cenn_trainer *t, const uint8_t *images_u8_host, float *losses_host);
CENN_API int cenn_trainer_step_images_u8_host_async(cenn_trainer *t, const uint8_t *images_u8_host);
CENN_API int cenn_trainer_step_frames_host(cenn_trainer *t, const uint8_t *frames_u8_host, int iH, int iW, const uint8_t *mask_full_host,
        const int *crop_host, const uint8_t *flip_host, const int *blocks_host, float maskValue, float *losses_host);
/* eval-mode generator forward (test_vid_wholeim.lua:180, demo.lua:68): in [B,Cin,F,F] -> out, host fp32 NCHW */
CENN_API int cenn_trainer_generator_forward_host(cenn_trainer *t, const float *in_host, float *out_host, int batch);
/* debugging / parity: copy an internal activation or gradient as fp32 NCHW to the host by name */
CENN_API int cenn_trainer_fetch_host(cenn_trainer *t, const char *name, float *dst_host, int64_t capacity, int64_t *count);
CENN_API int cenn_trainer_kernel_launches_per_step(cenn_trainer *t, int64_t *count);
/* one step with a CUDA-event pair around every op: '\n'-separated op names, per-op milliseconds and algorithmic FLOPs */
CENN_API int cenn_trainer_profile_step(cenn_trainer *t, const float *a_dev, const float *b_dev, const uint8_t *mask_dev,
        char *names, int64_t names_cap, float *ms, double *flops, int64_t cap, int64_t *nops);
/* test / debugging hook: run the step program up to and including the `occurrence`-th (0-based) op named `op_name`
 * (names as returned by cenn_trainer_profile_step), serially, and synchronise: stored tensors can then be fetched mid-step */
CENN_API int cenn_trainer_step_until(cenn_trainer *t, const float *a_dev, const float *b_dev, const uint8_t *mask_dev,
        const char *op_name, int occurrence, int64_t *ops_run);
/* algorithmic bytes of every op of the step program (0 where not stated), same order as cenn_trainer_profile_step:
 * the numerators of the HBM rooflines of the BN / activation / loss / Adam kernels */
CENN_API int cenn_trainer_op_bytes(cenn_trainer *t, double *bytes, int64_t cap, int64_t *nops);
/* one eager step on the executor's own streams: start / end (ms since the step began) of every op and the stream it ran on
 * (0 = compute stream, 1 = weight-gradient side stream, 2 = generator-forward chain, 3 = early-Adam stream) */
CENN_API int cenn_trainer_timeline_step(cenn_trainer *t, const float *a_dev, const float *b_dev, const uint8_t *mask_dev,
        float *t_start, float *t_end, int *stream_id, int64_t cap, int64_t *nops);

/* ------------------------------ inference engine ---------------
 * Eval-mode generator for test.lua:92 / demo.lua:68 (variant 0: [n,3,128,128] -> [n,3,64,64]) and for the full-frame
 * sweep of test_vid_wholeim.lua:98-226 (variant 1: [n,nc*inputLen,128,128] -> same shape).  BatchNorm running statistics
 * are folded into the bf16 operands and the epilogue bias when the parameters are loaded, so a forward is one tensor-core
 * GEMM launch per layer, replayed as a CUDA graph.  Tiles are independent in eval mode: `batch` is the number of tiles
 * one forward processes. */
typedef struct cenn_inpainter cenn_inpainter;
typedef struct cenn_inpainter_config {
    int variant;       /* 0: train.lua generator (centre prediction)   1: train_vid_weighted.lua generator */
    int batch;         /* tiles per forward */
    int fineSize;      /* 128 */
    int nBottleneck, nef, ngf;
    int nc;            /* channels per frame (3) */
    int inputLen;      /* frames stacked along the channel axis (the training scripts' predLen) */
} cenn_inpainter_config;
CENN_API int cenn_inpainter_create(cenn_state *s, const cenn_inpainter_config *cfg, cenn_inpainter **out);
CENN_API int cenn_inpainter_destroy(cenn_inpainter *p);
/* element counts of netG:getParameters() and of the concatenated [running_mean, running_var] per BN layer */
CENN_API int cenn_inpainter_param_count(cenn_inpainter *p, int64_t *params, int64_t *bn_stats);
/* flat parameters (Module:getParameters order, THNN layouts = what a *_net_G.t7 holds) + running statistics; folds BN */
CENN_API int cenn_inpainter_load_host(cenn_inpainter *p, const float *flat_host, const float *bn_stats_host);
/* n <= batch tiles, fp32 NCHW in and out; the device variant neither copies nor synchronises */
CENN_API int cenn_inpainter_forward_device(cenn_inpainter *p, const float *in_dev, float *out_dev, int n);
CENN_API int cenn_inpainter_forward_host(cenn_inpainter *p, const float *in_host, float *out_host, int n);
/* whole sweep on the device: pad to a multiple of fineSize, gather (+ vertical flip of the first three top tiles), forward,
 * write back, composite under the mask.  frames01 [P,nc,inh,inw] in [0,1] (P %% inputLen == 0), mask [inh,inw] (non-zero =
 * hole); out01 / full01 / inpaint01 [P,nc,outh,outw] in [0,1] (test_vid_wholeim.lua:222-224), any of them may be NULL.
 * init (may be NULL) = the optional initializer net of withInit (:179-190), same geometry as p. */
CENN_API int cenn_inpainter_sweep_host(cenn_inpainter *p, cenn_inpainter *init, const float *frames01_host, const uint8_t *mask_host,
        int P, int inh, int inw, float maskValue, float *out01_host, float *full01_host, float *inpaint01_host);

#ifdef __cplusplus
}
#endif
#endif /* CENN_H */
