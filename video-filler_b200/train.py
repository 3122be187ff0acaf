"""The training step of the reference, two ways.

``ClosureTrainer`` is the drop-in path: the ``fDx`` / ``fGx`` closures of ``train.lua:278-410`` and
``train_vid_weighted.lua:373-537`` written against the nn.* mirror, op by op, with the reference's
call pattern (four criterion:forward syncs per step, bias zeroing, gradient accumulation over the
two D passes, fGx reusing D's fake-pass state).  It runs in either precision mode.

``FusedTrainer`` wraps the C++ whole-step executor (``cenn_trainer_*``): NHWC bf16 activations,
tcgen05 implicit-GEMM convolutions, fused BN / loss / Adam kernels, losses kept on the device.
"""
import ctypes as C

import numpy as np

from . import _lib, models, nn, optim
from .tensor import CudaTensor, api, state

MEAN_FILL = (2 * 117.0 / 255.0 - 1.0, 2 * 104.0 / 255.0 - 1.0, 2 * 123.0 / 255.0 - 1.0)


class ClosureTrainer:
    def __init__(self, opt, seed=1234):
        self.opt = opt
        rng = np.random.default_rng(seed)
        self.netG = models.build_netG(opt)
        self.netD = models.build_netD(opt)
        models.weights_init(self.netG, rng)
        models.weights_init(self.netD, rng)
        self.criterion = nn.BCECriterion()
        self.criterionMSE = nn.MSECriterion() if opt["wtl2"] != 0 else None
        self.criterionGDL = nn.GDLCriterion(1) if opt["wtgdl"] != 0 else None
        wtl2 = opt["wtl2"]
        self.optimStateG = dict(learningRate=opt["lr"] * 10 if 0 < wtl2 < 1 else opt["lr"], beta1=opt["beta1"])
        self.optimStateD = dict(learningRate=opt["lr"], beta1=opt["beta1"])
        self.parametersD, self.gradParametersD = self.netD.getParameters()
        self.parametersG, self.gradParametersG = self.netG.getParameters()
        B, F, nc = opt["batchSize"], opt["fineSize"], models.net_channels(opt)
        self.input_ctx = CudaTensor(B, nc, F, F)
        self.label = CudaTensor(B)
        if opt["variant"] == "image":
            self.input_center = CudaTensor(B, nc, F // 2, F // 2)
            self.input_real_center = CudaTensor(B, nc, F // 2, F // 2)
        else:
            self.input_inpainted = CudaTensor(B, nc, F, F)
            self.input_mask = CudaTensor(B, nc, F, F)
            self.input_real = CudaTensor(B, nc, F, F)
        self.errD = self.errG = self.errG_l2 = self.errG_gdl = None

    # ---------------------------------------------------------------- image variant
    def _d_in(self, center):
        return [self.input_ctx, center] if self.opt.get("conditionAdv") else center       # train.lua:300-311

    def _g_in(self):
        return [self.input_ctx, self.noise] if self.opt.get("noiseGen") else self.input_ctx    # train.lua:325-329

    def fDx_image(self, real_ctx, real_center, noise=None):
        if self.opt.get("noiseGen"):
            # train.lua:319-323 redraws the noise here; the caller passes the draw so that both sides of a parity test see the same one
            if noise is None:
                raise ValueError("noiseGen: pass the noise tensor [B, nz, 1, 1] (train.lua:319-323)")
            if getattr(self, "noise", None) is None:
                self.noise = CudaTensor(self.opt["batchSize"], self.opt.get("nz", 100), 1, 1)
            self.noise.copy_(noise)
        models.zero_conv_bias(self.netD)
        models.zero_conv_bias(self.netG)
        self.gradParametersD.zero()
        self.input_ctx.copy_(real_ctx)
        self.input_center.copy_(real_center)
        self.input_real_center.copy_(real_center)
        self.label.fill(1)
        out = self.netD.forward(self._d_in(self.input_center))
        self.errD_real = self.criterion.forward(out, self.label)
        df_do = self.criterion.backward(out, self.label)
        self.netD.backward(self._d_in(self.input_center), df_do)
        fake = self.netG.forward(self._g_in())
        self.input_center.copy_(fake)
        self.label.fill(0)
        out = self.netD.forward(self._d_in(self.input_center))
        self.errD_fake = self.criterion.forward(out, self.label)
        df_do = self.criterion.backward(out, self.label)
        self.netD.backward(self._d_in(self.input_center), df_do)
        self.errD = self.errD_real + self.errD_fake
        return self.errD, self.gradParametersD

    def fGx_image(self):
        o = self.opt
        models.zero_conv_bias(self.netD)
        models.zero_conv_bias(self.netG)
        self.gradParametersG.zero()
        self.label.fill(1)
        out = self.netD.output
        self.errG = self.criterion.forward(out, self.label)
        df_do = self.criterion.backward(out, self.label)
        df_dg = self.netD.updateGradInput(self._d_in(self.input_center), df_do)
        if o.get("conditionAdv"):
            df_dg = df_dg[1]              # train.lua:371: df_dg[2], the prediction branch
        total = self.errG
        wtl2 = o["wtl2"]
        if wtl2 != 0:
            # train.lua:377-400 -- MSE forward + backward + overlap-weighted blend, one fused kernel
            loss = C.c_float()
            N, Cn, H, W = self.input_center.shape
            api().cenn_WeightedMSEBlend_overlap(state(), C.c_void_p(df_dg.ptr), C.c_void_p(self.input_center.ptr),
                                                C.c_void_p(self.input_real_center.ptr), N, Cn, H, W, wtl2,
                                                o["overlapPred"], C.byref(loss))
            self.errG_l2 = loss.value
            total = (1 - wtl2) * self.errG + wtl2 * self.errG_l2 if 0 < wtl2 < 1 else self.errG + wtl2 * self.errG_l2
        self.df_dg = df_dg
        self.netG.backward(self._g_in(), df_dg)
        return total, self.gradParametersG

    # ---------------------------------------------------------------- video variant
    def fDx_video(self, real_ctx, real_full, real_mask):
        o = self.opt
        models.zero_conv_bias(self.netD)
        models.zero_conv_bias(self.netG)
        self.gradParametersD.zero()
        self.input_ctx.copy_(real_ctx)
        self.input_real.copy_(real_full)
        self.label.fill(1)
        self.input_mask.copy_(real_mask)
        out = self.netD.forward(self.input_real)
        self.errD_real = self.criterion.forward(out, self.label)
        df_do = self.criterion.backward(out, self.label)
        self.netD.backward(self.input_real, df_do)
        fake = self.netG.forward(self.input_ctx)
        if o["weight_nomask"] == 0:
            self.input_inpainted.copy_(self.input_real)
            api().cenn_MaskComposite(state(), C.c_void_p(self.input_inpainted.ptr), C.c_void_p(self.input_mask.ptr),
                                     C.c_void_p(fake.ptr), fake.nelement())
        else:
            self.input_inpainted.copy_(fake)
        self.label.fill(0)
        out = self.netD.forward(self.input_inpainted)
        self.errD_fake = self.criterion.forward(out, self.label)
        df_do = self.criterion.backward(out, self.label)
        self.netD.backward(self.input_inpainted, df_do)
        self.errD = self.errD_real + self.errD_fake
        return self.errD, self.gradParametersD

    def fGx_video(self):
        o = self.opt
        models.zero_conv_bias(self.netD)
        models.zero_conv_bias(self.netG)
        self.gradParametersG.zero()
        self.label.fill(1)
        out = self.netD.output
        self.errG = self.criterion.forward(out, self.label)
        df_do = self.criterion.backward(out, self.label)
        df_dg = self.netD.updateGradInput(self.input_real, df_do)
        total = self.errG
        wtl2 = o["wtl2"]
        if o["overlapPred"] != 0:
            raise ValueError("video scripts require overlapPred == 0 (train_vid_weighted.lua:509)")
        if o["wtgdl"] != 0:
            # loss from GDL (:524); its gradient in the script is criterionMSE:backward (:525), folded into the blend
            self.errG_gdl = self.criterionGDL.forward(self.input_inpainted, self.input_real)
        if wtl2 != 0 or o["wtgdl"] != 0:
            # the GDL branch adds wtgdl * criterionMSE:backward to df_dg UNCONDITIONALLY (train_vid_weighted.lua:523-528), also when wtl2 == 0:
            # the fused kernel handles wtl2 = 0 (no L2 term, adversarial weight 1)
            loss = C.c_float()
            api().cenn_WeightedMSEBlend_masked(state(), C.c_void_p(df_dg.ptr), C.c_void_p(self.input_inpainted.ptr),
                                               C.c_void_p(self.input_real.ptr), C.c_void_p(self.input_mask.ptr),
                                               df_dg.nelement(), wtl2, o["weight_nomask"], o["wtgdl"], C.byref(loss))
            if wtl2 != 0:
                self.errG_l2 = loss.value
                total = (1 - wtl2) * self.errG + wtl2 * self.errG_l2 if 0 < wtl2 < 1 else self.errG + wtl2 * self.errG_l2
        if o["wtgdl"] != 0:
            total = total + o["wtgdl"] * self.errG_gdl
        self.df_dg = df_dg
        self.netG.backward(self.input_ctx, df_dg)
        return total, self.gradParametersG

    # ---------------------------------------------------------------- one step (train.lua:421-424)
    def step(self, *batch):
        if self.opt["variant"] == "image":
            optim.adam(lambda x: self.fDx_image(*batch), self.parametersD, self.optimStateD)
            _, fx = optim.adam(lambda x: self.fGx_image(), self.parametersG, self.optimStateG)
        else:
            optim.adam(lambda x: self.fDx_video(*batch), self.parametersD, self.optimStateD)
            _, fx = optim.adam(lambda x: self.fGx_video(), self.parametersG, self.optimStateG)
        return dict(errD=self.errD, errG=self.errG, errG_l2=self.errG_l2, errG_gdl=self.errG_gdl,
                    errD_real=self.errD_real, errD_fake=self.errD_fake, errG_total=fx[0])


LOSS_NAMES = ("errD", "errG", "errG_l2", "errG_gdl", "errD_real", "errD_fake", "errG_total")


def dp_step(trainer, a_ptr, b_ptr, mask_ptr, all_reduce):
    """One data-parallel step: run the executor's program phase by phase; after each phase sum the buffer it names
    (BN batch statistics, BN backward sums, the D / G gradient vectors, the loss accumulators) across ranks with
    ``all_reduce(device_ptr, count, is_double)``.  Every rank holds full parameters and B/world samples; criteria
    and BN use the global batch size, so the summed quantities equal the single-process ones.  Returns the number
    of collectives issued."""
    trainer.step_phase(-1, a_ptr, b_ptr, mask_ptr)
    n = 0
    while True:
        buf, count, is_double, done = trainer.sync_info()
        if done:
            return n
        if buf and count:
            all_reduce(buf, count, is_double)
            n += 1
        trainer.step_phase(0, a_ptr, b_ptr, mask_ptr)


class FusedTrainer:
    """Whole-step executor (cenn_trainer_*): one call per G+D step, host or device inputs."""

    def __init__(self, opt, precision="bf16", world_size=1, rank=0, dead_dgrad=1, bn_local=0):
        self.opt = opt
        cfg = _lib.TrainerConfig(
            variant=0 if opt["variant"] == "image" else 1, batchSize=opt["batchSize"], fineSize=opt["fineSize"],
            nBottleneck=opt["nBottleneck"], nef=opt["nef"], ngf=opt["ngf"], ndf=opt["ndf"], nc=opt["nc"],
            predLen=opt.get("predLen", 1), overlapPred=opt["overlapPred"], wtl2=opt["wtl2"],
            weight_nomask=opt.get("weight_nomask", 0.0), wtgdl=opt.get("wtgdl", 0.0), lr=opt["lr"], beta1=opt["beta1"],
            precision={"fp32": 0, "bf16": 1}[precision], world_size=world_size, rank=rank, dead_dgrad=dead_dgrad,
            noiseGen=1 if opt.get("noiseGen") else 0, nz=opt.get("nz", 100), conditionAdv=1 if opt.get("conditionAdv") else 0, bn_local=int(bn_local))
        self.cfg = cfg
        h = C.c_void_p()
        api().cenn_trainer_create(state(), C.byref(cfg), C.byref(h))
        self.h = h

    def close(self):
        """Destroy the executor (buffers, CUDA graph).  Must precede cenn_dist_shutdown: the graph references the communicator."""
        if getattr(self, "h", None):
            api().cenn_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def param_count(self, net):
        n = C.c_int64()
        api().cenn_trainer_param_count(self.h, net, C.byref(n))
        return n.value

    def set_params(self, net, flat):
        flat = np.ascontiguousarray(flat, np.float32)
        assert flat.size == self.param_count(net)
        api().cenn_trainer_set_params_host(self.h, net, flat.ctypes.data_as(C.c_void_p))

    def get_params(self, net):
        out = np.empty(self.param_count(net), np.float32)
        api().cenn_trainer_get_params_host(self.h, net, out.ctypes.data_as(C.c_void_p))
        return out

    def get_grads(self, net):
        out = np.empty(self.param_count(net), np.float32)
        api().cenn_trainer_get_grads_host(self.h, net, out.ctypes.data_as(C.c_void_p))
        return out

    def bn_stat_count(self, net):
        n = C.c_int64()
        api().cenn_trainer_bn_stat_count(self.h, net, C.byref(n))
        return n.value

    def get_bn_stats(self, net):
        out = np.empty(self.bn_stat_count(net), np.float32)
        api().cenn_trainer_get_bn_stats_host(self.h, net, out.ctypes.data_as(C.c_void_p))
        return out

    def set_bn_stats(self, net, stats):
        stats = np.ascontiguousarray(stats, np.float32)
        assert stats.size == self.bn_stat_count(net)
        api().cenn_trainer_set_bn_stats_host(self.h, net, stats.ctypes.data_as(C.c_void_p))

    def set_noise(self, noise):
        """noiseGen: the noise draw [B, nz(,1,1)] of the next step (train.lua:319-323 redraws it inside fDx; here the caller draws it)."""
        noise = np.ascontiguousarray(noise, np.float32)
        assert noise.size == self.opt["batchSize"] * self.opt.get("nz", 100)
        api().cenn_trainer_set_noise_host(self.h, noise.ctypes.data_as(C.c_void_p))

    def step_host(self, a, b, mask=None, noise=None):
        """One step from host fp32 NCHW arrays (H2D + step + D2H of the losses)."""
        if noise is not None:
            self.set_noise(noise)
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        m = np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p) if mask is not None else None
        losses = np.zeros(8, np.float32)
        api().cenn_trainer_step_host(self.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), m,
                                     losses.ctypes.data_as(C.c_void_p))
        return dict(zip(LOSS_NAMES, losses.tolist()))

    def step_host_async(self, a, b, mask=None):
        """Enqueue one step from host arrays (must stay alive until the matching wait_losses); at most two in flight."""
        assert a.dtype == np.float32 and b.dtype == np.float32 and a.flags.c_contiguous and b.flags.c_contiguous
        m = mask.ctypes.data_as(C.c_void_p) if mask is not None else None
        api().cenn_trainer_step_host_async(self.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), m)

    def wait_losses(self):
        losses = np.zeros(8, np.float32)
        api().cenn_trainer_wait_losses(self.h, losses.ctypes.data_as(C.c_void_p))
        return dict(zip(LOSS_NAMES, losses.tolist()))

    def step_clips_host(self, frames01, mask1, flip=None, maskValue=None):
        """Clip-mode step (video variant): frames01 [B, nc*predLen, F, F] in [0,1], mask1 [B, F, F] uint8 (one plane per
        sample), flip [B] uint8 or None; the device derives real_full, real_ctx and the expanded mask."""
        frames01 = np.ascontiguousarray(frames01, np.float32)
        mask1 = np.ascontiguousarray(mask1, np.uint8)
        f = np.ascontiguousarray(flip, np.uint8).ctypes.data_as(C.c_void_p) if flip is not None else None
        mv = float(self.opt["maskValue"] if maskValue is None else maskValue)
        losses = np.zeros(8, np.float32)
        api().cenn_trainer_step_clips_host(self.h, frames01.ctypes.data_as(C.c_void_p), mask1.ctypes.data_as(C.c_void_p), f, mv,
                                           losses.ctypes.data_as(C.c_void_p))
        return dict(zip(LOSS_NAMES, losses.tolist()))

    def step_images_u8_host(self, images_u8):
        """Byte-image step (image variant): uint8 [B, 3, F, F]; the device rescales to [-1,1], clones the centre and mean-fills it (train.lua:286-290)."""
        images_u8 = np.ascontiguousarray(images_u8, np.uint8)
        losses = np.zeros(8, np.float32)
        api().cenn_trainer_step_images_u8_host(self.h, images_u8.ctypes.data_as(C.c_void_p), losses.ctypes.data_as(C.c_void_p))
        return dict(zip(LOSS_NAMES, losses.tolist()))

    def step_images_u8_host_async(self, images_u8):
        assert images_u8.dtype == np.uint8 and images_u8.flags.c_contiguous
        api().cenn_trainer_step_images_u8_host_async(self.h, images_u8.ctypes.data_as(C.c_void_p))

    def step_frames_host(self, frames_u8, mask_full, crop, flip, blocks, maskValue=None):
        """Frame-mode step (video variant): decoded frames uint8 [B, nc*predLen, iH, iW], the full-size mask uint8 [iH, iW] and the loader
        hook's draws (crop int32 [B,2] 0-based, flip uint8 [B] or None, blocks int32 [B,21]); the device runs the hook
        (datavid/donkey_folder.lua:114-129,138-187)."""
        frames_u8 = np.ascontiguousarray(frames_u8, np.uint8)
        mask_full = np.ascontiguousarray(mask_full, np.uint8)
        crop = np.ascontiguousarray(crop, np.int32)
        blocks = np.ascontiguousarray(blocks, np.int32)
        assert frames_u8.ndim == 4 and mask_full.shape == frames_u8.shape[2:] and crop.shape == (frames_u8.shape[0], 2) and blocks.shape == (frames_u8.shape[0], 21)
        f = np.ascontiguousarray(flip, np.uint8).ctypes.data_as(C.c_void_p) if flip is not None else None
        mv = float(self.opt["maskValue"] if maskValue is None else maskValue)
        losses = np.zeros(8, np.float32)
        api().cenn_trainer_step_frames_host(self.h, frames_u8.ctypes.data_as(C.c_void_p), int(frames_u8.shape[2]), int(frames_u8.shape[3]),
                                            mask_full.ctypes.data_as(C.c_void_p), crop.ctypes.data_as(C.c_void_p), f, blocks.ctypes.data_as(C.c_void_p), mv,
                                            losses.ctypes.data_as(C.c_void_p))
        return dict(zip(LOSS_NAMES, losses.tolist()))

    def step_clips_host_async(self, frames01, mask1, flip=None, maskValue=None):
        assert frames01.dtype == np.float32 and frames01.flags.c_contiguous and mask1.dtype == np.uint8 and mask1.flags.c_contiguous
        f = flip.ctypes.data_as(C.c_void_p) if flip is not None else None
        mv = float(self.opt["maskValue"] if maskValue is None else maskValue)
        api().cenn_trainer_step_clips_host_async(self.h, frames01.ctypes.data_as(C.c_void_p), mask1.ctypes.data_as(C.c_void_p), f, mv)

    def step_device(self, a_ptr, b_ptr, mask_ptr=None):
        api().cenn_trainer_step_device(self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr),
                                       C.c_void_p(mask_ptr) if mask_ptr else None)

    def step_phase(self, phase, a_ptr, b_ptr, mask_ptr=None):
        api().cenn_trainer_step_phase(self.h, int(phase), C.c_void_p(a_ptr), C.c_void_p(b_ptr),
                                      C.c_void_p(mask_ptr) if mask_ptr else None)

    def sync_info(self):
        """(device pointer, element count, is_double, done) of the synchronisation point the last step_phase stopped at."""
        buf, n, dbl, done = C.c_void_p(), C.c_int64(), C.c_int(), C.c_int()
        api().cenn_trainer_sync_info(self.h, C.byref(buf), C.byref(n), C.byref(dbl), C.byref(done))
        return buf.value, n.value, bool(dbl.value), bool(done.value)

    def grad_buffer(self, net):
        p, n = C.c_void_p(), C.c_int64()
        api().cenn_trainer_grad_buffer(self.h, net, C.byref(p), C.byref(n))
        return p.value, n.value

    def read_losses(self):
        losses = np.zeros(8, np.float32)
        api().cenn_trainer_read_losses(self.h, losses.ctypes.data_as(C.c_void_p))
        return dict(zip(LOSS_NAMES, losses.tolist()))

    def generator_forward(self, x):
        x = np.ascontiguousarray(x, np.float32)
        B = x.shape[0]
        F = self.opt["fineSize"]
        nc = models.net_channels(self.opt)
        oF = F // 2 if self.opt["variant"] == "image" else F
        out = np.empty((B, nc, oF, oF), np.float32)
        api().cenn_trainer_generator_forward_host(self.h, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), B)
        return out

    def fetch(self, name, capacity=1 << 28):
        n = C.c_int64()
        buf = np.empty(capacity, np.float32)
        api().cenn_trainer_fetch_host(self.h, name.encode(), buf.ctypes.data_as(C.c_void_p), capacity, C.byref(n))
        return buf[:n.value].copy()

    def profile_step(self, a_ptr, b_ptr, mask_ptr=None, repeats=3):
        """Per-op CUDA-event timing of one step (averaged over `repeats` runs after one untimed run)."""
        cap = 4096
        names = C.create_string_buffer(1 << 16)
        ms = np.zeros(cap, np.float32)
        fl = np.zeros(cap, np.float64)
        n = C.c_int64()
        acc = None
        for r in range(repeats + 1):
            api().cenn_trainer_profile_step(self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(mask_ptr) if mask_ptr else None,
                                            names, len(names), ms.ctypes.data_as(C.c_void_p), fl.ctypes.data_as(C.c_void_p), cap, C.byref(n))
            if r == 0:
                acc = np.zeros(n.value, np.float64)
            else:
                acc += ms[:n.value]
        acc /= repeats
        nm = names.value.decode().split("\n")[:n.value]
        flops = fl[:n.value]
        by_op = {}
        for k, t in zip(nm, acc):
            by_op[k] = by_op.get(k, 0.0) + float(t)
        tc = flops > 0
        by = np.zeros(cap, np.float64)
        nb = C.c_int64()
        api().cenn_trainer_op_bytes(self.h, by.ctypes.data_as(C.c_void_p), cap, C.byref(nb))
        best = None
        for i in np.nonzero(tc)[0]:
            tf = flops[i] / (acc[i] * 1e-3) / 1e12 if acc[i] > 0 else 0.0
            if flops[i] >= 1e10 and (best is None or tf > best["tflops"]):
                best = {"op": nm[i], "index": int(i), "ms": round(float(acc[i]), 4), "gflop": round(float(flops[i]) / 1e9, 2), "tflops": round(float(tf), 1)}
        return {"names": nm, "ms": acc, "flops": flops, "bytes": by[:n.value].copy(), "total_ms": float(acc.sum()), "tc_ms": float(acc[tc].sum()),
                "tc_flops": float(flops[tc].sum()), "tc_launches": int(tc.sum()), "best_tc": best,
                "by_op": {k: round(v, 4) for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])}}

    def timeline(self, a_ptr, b_ptr, mask_ptr=None):
        """One eager step on the executor's own streams: [(name, stream id, start ms, end ms)] per op (cenn_trainer_timeline_step)."""
        cap = 4096
        names = C.create_string_buffer(1 << 16)
        ms = np.zeros(cap, np.float32)
        fl = np.zeros(cap, np.float64)
        n = C.c_int64()
        api().cenn_trainer_profile_step(self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(mask_ptr) if mask_ptr else None,
                                        names, len(names), ms.ctypes.data_as(C.c_void_p), fl.ctypes.data_as(C.c_void_p), cap, C.byref(n))
        nm = names.value.decode().split("\n")[:n.value]
        t0, t1, sid = np.zeros(cap, np.float32), np.zeros(cap, np.float32), np.zeros(cap, np.int32)
        for _ in range(2):      # second run: warm
            api().cenn_trainer_timeline_step(self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(mask_ptr) if mask_ptr else None,
                                             t0.ctypes.data_as(C.c_void_p), t1.ctypes.data_as(C.c_void_p), sid.ctypes.data_as(C.c_void_p), cap, C.byref(n))
        return [(nm[i], int(sid[i]), float(t0[i]), float(t1[i])) for i in range(n.value)]

    def step_until(self, a_ptr, b_ptr, mask_ptr, op_name, occurrence=0):
        """Run the step program up to and including the `occurrence`-th op named `op_name` (test hook, cenn_trainer_step_until)."""
        n = C.c_int64()
        api().cenn_trainer_step_until(self.h, C.c_void_p(a_ptr), C.c_void_p(b_ptr), C.c_void_p(mask_ptr) if mask_ptr else None,
                                      op_name.encode(), int(occurrence), C.byref(n))
        return n.value

    def launches_per_step(self):
        n = C.c_int64()
        api().cenn_trainer_kernel_launches_per_step(self.h, C.byref(n))
        return n.value
