#!/usr/bin/env python
"""bench.py -- train samples/sec of one G+D step (optim.adam(fDx) + optim.adam(fGx), train.lua:421-424).

  python bench.py --gpus N --steps K --warmup W            our arm (fused executor, BF16 tcgen05 path)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU (gpu=0) path = the oracle port

Workload (BASELINE.json configs[1]): inpaintCenter context encoder, fineSize 128, overlapPred 4, nBottleneck 4000,
nef=ngf=ndf=64, batch 256 per GPU (weak scaling), synthetic U(-1,1) RGB, random-init weights N(0,0.02).
One JSON line is printed by rank 0.  `value` is timed on the device with inputs resident in HBM; `e2e` goes through
cenn_trainer_step_host with pinned HOST buffers (H2D of the inputs and D2H of the losses inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "inpaintCenter 128x128 G+D training step, batch 256/GPU, nBottleneck 4000, overlapPred 4, wtl2 0.999"
STEP_GFLOP_PER_SAMPLE = 3.552          # 3 F_G + 7 F_D per sample
# non-default workloads (--workload): BASELINE.json configs[2] per GPU and configs[4]
WORKLOAD_VIDEO = "train_vid_weighted 128x128 clips, predLen 4 (12 stacked channels), mask-weighted L2 + GDL (wtgdl %s), batch %d/GPU"
VIDEO_GFLOP_PER_SAMPLE = 5.242         # cfg3, 3 F_G + 7 F_D per sample
INFER_GFLOP_PER_TILE = 0.853           # cfg5 generator forward per 128x128 tile
# BASELINE.json configs[3]: train_deepernet at 256 x 256, global batch 256 split over the ranks (strong scaling)
WORKLOAD_DEEPER = ("train_deepernet 256x256 clips, predLen 4 (12 stacked channels), mask-weighted L2 (wtgdl %s), nBottleneck 4000 (5x5 bottleneck map), "
                   "patch discriminator head (25 outputs per sample, each labelled with its sample's label), global batch 256 = %d/GPU")
DEEPER_GFLOP_PER_SAMPLE = 29.228       # cfg4, 7482.3 GFLOP per 256-sample step


def opt_for(batch, variant="image"):
    from video_filler_b200 import models
    return models.default_opt(variant, batchSize=batch)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md), polled in-process through NVML every
    few milliseconds so that even a 60 ms region holds several samples; falls back to `nvidia-smi -lms` when NVML is missing."""

    def __init__(self, index, period_s=0.004):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.period = index, [], False, period_s
        self.proc = None
        self.t_begin = self.t_end = None
        self.source = "nvml"

    def mark_begin(self):
        """Samples taken from now on are inside the timed region."""
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def _handle(self, nv):
        try:                                  # honour CUDA_VISIBLE_DEVICES: look the device up by UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            return nv.nvmlDeviceGetHandleByIndex(self.index)

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = self._handle(nv)
            smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
            while not self.stop_flag:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.time(), sm, smax, pw, [n for n, b in names if mask & b]))
                time.sleep(self.period)
            return
        except Exception:
            self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                v = [x.strip() for x in line.split(",")]
                try:
                    rs = [n for n, x in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), v[3:7]) if x.lower().startswith("active")]
                    self.samples.append((time.time(), float(v[0]), float(v[1]), float(v[2]), rs))
                except Exception:
                    continue
        except Exception:
            pass

    def finish(self):
        if self.t_end is None:
            self.mark_end()
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        if self.is_alive():
            self.join(timeout=1.0)
        inside = [x for x in self.samples if (self.t_begin is None or x[0] >= self.t_begin) and x[0] <= self.t_end + self.period]
        sm = [x[1] for x in inside]
        reasons = sorted({r for x in inside for r in x[4]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": inside[0][2] if inside else (self.samples[-1][2] if self.samples else None),
                "reasons": reasons, "samples": len(sm), "power_w_max": max([x[3] for x in inside]) if inside else None, "source": self.source}


def host_threads():
    """Host threads this process may use (affinity-aware; torchrun's OMP_NUM_THREADS=1 is overridden explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_step_rate(batch, steps, warmup, threads, variant="image", wtgdl=0.0, fine=128):
    """The reference's gpu=0 path restated on the host cores: the oracle's fDx / fGx / optim.adam step sequence (oracle/step.py,
    train.lua:278-410) with its heavy ops on the PyTorch-CPU engine (oracle/torch_engine.py: oneDNN / MKL -- BASELINE.md section 4).
    Returns (samples/s over `steps` steps, seconds per step, threads in use)."""
    from oracle import nets as onets
    from oracle import step as ostep
    from oracle import torch_engine
    n_thr = torch_engine.enable(threads)
    kw = dict(batchSize=batch)
    if variant == "video":
        kw["wtgdl"] = wtgdl
        kw["fineSize"] = fine
    orc = ostep.StepOracle(onets.default_opt(variant, **kw), seed=1234, dtype=np.float32)
    rng = np.random.default_rng(1234)
    batches = [orc.synth_batch(rng) for _ in range(min(2, steps + warmup))]
    for i in range(warmup):
        orc.step(*batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps):
        orc.step(*batches[i % len(batches)])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, n_thr


REF_BUDGET_S = 240.0     # the whole reference run (warm-up + timed steps) is sized to end within a few minutes


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the step on this box's host cores, all threads, on the SAME
    workload string as our arm.  Each step processes the full per-GPU batch when `steps + warmup` of them fit REF_BUDGET_S
    (probed with one small step first); otherwise a bounded sample of the batch, and the line says which."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = getattr(args, "workload", "image")
    deeper = wl == "deeper"
    video = wl == "video" or deeper
    world = int(os.environ.get("WORLD_SIZE", "1"))
    full = getattr(args, "batch", None) or ((256 // world) if deeper else (64 if video else 256))
    threads = host_threads()
    variant = "video" if video else "image"
    wtgdl = getattr(args, "wtgdl", 0.5) if video else 0.0
    if deeper and wtgdl == 0.5:
        wtgdl = 0.0
    fine = 256 if deeper else 128
    nsteps = max(1, args.steps + args.warmup)
    probe_b = min(8 if deeper else 16, full)
    probe_rate, _, _ = cpu_port_step_rate(probe_b, 1, 1, threads, variant, wtgdl, fine)    # per-step fixed costs (Adam over 74 M params) make this pessimistic
    batch = full
    if full * nsteps / probe_rate > REF_BUDGET_S:
        batch = int(max(2, min(full, REF_BUDGET_S * probe_rate // nsteps)))
    rate, sec, n_thr = cpu_port_step_rate(batch, args.steps, args.warmup, threads, variant, wtgdl, fine)
    workload = (WORKLOAD_DEEPER % (wtgdl, full)) if deeper else ((WORKLOAD_VIDEO % (wtgdl, full)) if video else WORKLOAD.replace("batch 256", "batch %d" % full))
    sample = ("each step = the full %d-sample batch" % full) if batch == full else ("each step = %d samples of the %d-sample batch (bounded sample)" % (batch, full))
    line = {
        "impl": "reference", "metric": "train samples/sec (G+D step)", "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": sample, "batch_per_step": batch},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": n_thr, "kind": "port",
                         "sample": "oracle step sequence on the PyTorch-CPU engine (oneDNN/MKL; faster than Torch7's im2col+sgemm gpu=0 path would be), fp32, %d threads, %s" % (n_thr, sample)},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_infer(args, torch, dist, api, st, stream, rank, local_rank, world):
    """BASELINE.json configs[4]: eval-mode generator over 128x128 tiles (test_vid_wholeim.lua:159-205), batch 1..1024 tiles per
    forward, tiles resident in HBM (fp32 NCHW in / out, converted inside the timed region); replicas only at N > 1 (no collective).
    e2e = the whole device-side sweep of 32 frames of 360x480 (12 tiles per frame) from host frames to host composites."""
    from video_filler_b200 import infer, models, util
    import video_filler_b200.tensor as T
    opt = models.default_opt("video", predLen=1)
    rng = np.random.default_rng(1234)
    flat = util.params_flat(util.weights_init(util.describe_netG(opt), rng))
    batches = [args.batch] if args.batch else [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
    hbm, tf_burst, tf_sus, peak_src = peaks()
    sweep, stats = [], None
    sampler = ClockSampler(local_rank); sampler.start(); sampler.mark_begin()
    launches0 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches0))
    with torch.cuda.stream(stream):
        for Bt in batches:
            eng = infer.Inpainter(opt, Bt)
            if stats is None:
                n_stats = eng.counts()[1]
                srng = np.random.default_rng(7)
                stats = np.concatenate([srng.normal(0, 0.1, n_stats // 2), srng.uniform(0.5, 1.5, n_stats - n_stats // 2)]).astype(np.float32)
                # [mean(C), var(C)] per layer: keep every variance slot positive whatever the interleaving
                stats = np.abs(stats) + 0.25
            eng.load(flat, stats)
            xs = [T.CudaTensor.from_numpy(rng.uniform(-1, 1, (Bt, 3, 128, 128)).astype(np.float32)) for _ in range(2)]
            y = T.CudaTensor.from_numpy(np.zeros((Bt, 3, 128, 128), np.float32))
            for i in range(max(4, args.warmup)):
                eng.forward_device(xs[i % 2].ptr, y.ptr, Bt)
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(args.steps):
                eng.forward_device(xs[i % 2].ptr, y.ptr, Bt)
            e1.record(stream)
            torch.cuda.synchronize()
            t_ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if dist:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            ms = float(t_ms.item()) / args.steps
            sweep.append({"batch": Bt, "latency_ms": ms, "tiles_per_s": Bt * world / (ms * 1e-3),
                          "tflops": INFER_GFLOP_PER_TILE * 1e-3 * Bt / (ms * 1e-3)})
            eng.close()
            del xs, y
        launches1 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches1))
        clocks = sampler.finish()
        # ---- end to end: host frames -> device sweep -> host composites
        P, inh, inw = 32, 360, 480
        def pinned(shape):
            n = int(np.prod(shape)); ptr = C.c_void_p()
            api.cenn_host_alloc(st, n * 4, C.byref(ptr))
            return np.ctypeslib.as_array((C.c_float * n).from_address(ptr.value)).reshape(shape)
        frames = pinned((P, 3, inh, inw)); frames[:] = rng.uniform(0, 1, frames.shape)
        result = {"inpaint": pinned((P, 3, 384, 512))}
        mask = np.zeros((inh, inw), bool); mask[120:220, 180:330] = True
        eng = infer.Inpainter(opt, 12 * P)
        eng.load(flat, stats)
        sweep_once = lambda: eng.sweep(frames, mask, opt["maskValue"], want=("inpaint",), buffers=result)
        for _ in range(2):
            sweep_once()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        reps = max(3, min(20, args.steps))
        t0 = time.perf_counter()
        for _ in range(reps):
            sweep_once()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e2e_s = (time.perf_counter() - t0) / reps
        eng.close()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    best = max(sweep, key=lambda r: r["tiles_per_s"])
    line = {
        "metric": "inference tiles/sec (eval-mode generator forward, 128x128 tiles)", "value": best["tiles_per_s"], "unit": "tiles/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(4, args.warmup), "ms_per_step": best["latency_ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "test_vid_wholeim tile forward, inputLen 1, nBottleneck 4000; batch sweep %s tiles/GPU, value = best batch (%d)" % (batches, best["batch"]),
                   "parallelism": "replicas x%d (no collective)" % world,
                   "l2": "small batches are L2-resident by nature (latency case); from 64 tiles up the activations exceed the 126 MB L2"},
        "gpu_launches": int(launches1.value - launches0.value), "clocks": clocks, "sweep": sweep,
        "e2e": {"value": 12 * P * world / e2e_s, "unit": "tiles/s", "frames_per_s": P * world / e2e_s, "ms_per_sweep": e2e_s * 1e3,
                "workload": "cenn_inpainter_sweep_host: %d frames of %dx%d (12 tiles each) pinned host frames -> device sweep -> pinned host inpaintImages" % (P, inh, inw),
                "h2d_bytes_per_step": int(frames.nbytes + mask.size), "d2h_bytes_per_step": int(P * 3 * 384 * 512 * 4)},
        "roofline": {"bound": "tensor", "achieved": best["tflops"], "peak": tf_sus, "unit": "TFLOP/s", "frac": best["tflops"] / tf_sus, "traffic": None,
                     "kernel": "whole forward (12 tcgen05 GEMM launches + layout conversion) at the best batch", "peak_source": peak_src + " bf16_tflops_sustained"},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import nets as onets
        from oracle import step as ostep
        torch.set_num_threads(os.cpu_count() or 1)
        orc = ostep.StepOracle(onets.default_opt("video", predLen=1, batchSize=8), seed=1234, dtype=np.float32)
        orc.netG.evaluate()
        x = rng.uniform(-1, 1, (8, 3, 128, 128)).astype(np.float32)
        orc.netG.forward(x)
        t0 = time.perf_counter(); orc.netG.forward(x); dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 8 / dt, "unit": "tiles/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "one 8-tile eval forward of the oracle generator, %.2f s" % dt}
    print(json.dumps(line))
    sys.stdout.flush()
    if dist:
        dist.destroy_process_group()


def wrap_device(ptr, count, dtype, torch):
    """A torch view of a raw device buffer (for torch.distributed collectives on the executor's buffers)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8" if dtype == "f8" else "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")




def measure_train(args, torch, dist, api, st, stream, rank, local_rank, world, video, B, steps, warmup, want_profile=True, fine=128, bn_local=0, want_e2e=True):
    """Time `steps` G+D steps of one workload on this rank.  Returns a dict of raw measurements (device-timed region, end-to-end
    region through the pipelined host API, per-op profile, launch count, clocks); the executor is closed before returning."""
    from video_filler_b200 import synth, train, util
    import video_filler_b200.tensor as T
    opt = opt_for(B, "video" if video else "image")
    if video:
        opt["wtgdl"] = args.wtgdl
    opt["fineSize"] = fine
    trn = train.FusedTrainer(opt, precision="bf16", world_size=world, rank=rank, bn_local=bn_local)
    # identical random-init weights on every rank (parameter broadcast = same seed), train.lua:58-67
    rng = np.random.default_rng(1234)
    nG, nD = trn.param_count(0), trn.param_count(1)
    trn.set_params(0, util.params_flat(util.weights_init(util.describe_netG(opt), rng)))
    trn.set_params(1, util.params_flat(util.weights_init(util.describe_netD(opt), rng)))
    assert video or (nG == 71118691 + 2 * (64 + 128 + 256 + 512 + 4000 + 512 + 256 + 128 + 64) and nD > 2764737)

    drng = np.random.default_rng(1000 + rank)
    n_batches = 2
    host, clips, pinned = [], [], []

    def pin(nbytes):
        p = C.c_void_p(); api.cenn_host_alloc(st, nbytes, C.byref(p)); pinned.append(p); return p

    for _ in range(n_batches):
        if video:
            ctx, center, mask = synth.video_batch(B, 12, fine, opt["maskValue"], drng)
        else:
            ctx, center = synth.image_batch(B, 128, 4, drng)
            mask = None
        pa, pb = pin(ctx.nbytes), pin(center.nbytes)
        ha = np.ctypeslib.as_array((C.c_float * ctx.size).from_address(pa.value)); ha[:] = ctx.ravel()
        hb = np.ctypeslib.as_array((C.c_float * center.size).from_address(pb.value)); hb[:] = center.ravel()
        da, db = T.CudaTensor.from_numpy(ctx), T.CudaTensor.from_numpy(center)
        hm, dm, nbytes = None, None, ctx.nbytes + center.nbytes
        if mask is not None:
            pm = pin(mask.nbytes)
            hm = np.ctypeslib.as_array((C.c_uint8 * mask.size).from_address(pm.value)); hm[:] = mask.ravel()
            dmp = C.c_void_p(); api.cenn_malloc(st, mask.nbytes, C.byref(dmp)); api.cenn_copy_h2d(st, dmp, pm, mask.nbytes)
            dm = dmp.value
            # clip mode (the e2e call of the video workload): frames in [0,1] + ONE mask plane per sample + hflip flags;
            # the device derives real_full, real_ctx and the expanded mask (datavid/donkey_folder.lua:161-187)
            pf = pin(center.nbytes)
            hf = np.ctypeslib.as_array((C.c_float * center.size).from_address(pf.value)); hf[:] = ((center + 1) * 0.5).ravel()
            m1 = np.ascontiguousarray(mask[:, 0]); p1 = pin(m1.nbytes + B)
            h1 = np.ctypeslib.as_array((C.c_uint8 * (m1.size + B)).from_address(p1.value)); h1[:m1.size] = m1.ravel()
            h1[m1.size:] = drng.integers(0, 2, B).astype(np.uint8)
            clips.append((hf, h1[:m1.size], h1[m1.size:], center.nbytes + m1.nbytes + B))
            nbytes = clips[-1][3]
        host.append((ha, hb, da, db, nbytes, hm, dm))

    def step_device(i):
        # world > 1: the executor all-reduces BN statistics, gradients and losses itself (NCCL / peer mailboxes, inside its CUDA graph)
        h = host[i % n_batches]
        trn.step_device(h[2].ptr, h[3].ptr, h[6])

    sampler = ClockSampler(local_rank); sampler.start()
    with torch.cuda.stream(stream):
        for i in range(max(4, warmup)):      # two input sets x (first use runs eagerly, second captures its CUDA graph): never fewer than 4
            step_device(i)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        launches0 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches0))
        sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step_device(i)
        e1.record(stream)
        torch.cuda.synchronize()
        sampler.mark_end()
        if dist:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.finish()
        launches1 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches1))
        losses = trn.read_losses()
        # ---- end-to-end: host buffers in, losses out, every step (single-process API call)
        for i in range(2):
            trn.step_host(host[i % n_batches][0], host[i % n_batches][1], host[i % n_batches][5])

        def step_async(i):
            if video:
                c = clips[i % n_batches]
                trn.step_clips_host_async(c[0], c[1], c[2], opt["maskValue"])
            else:
                trn.step_host_async(host[i % n_batches][0], host[i % n_batches][1])
        for i in range(4):      # both staging sets of the pipelined path: first use runs eagerly, second captures its CUDA graph
            step_async(i)
            if i > 0:
                trn.wait_losses()
        trn.wait_losses()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        # public pipelined API: every step copies ITS inputs from pinned host memory and its losses are read back;
        # the copy of step k+1 overlaps the compute of step k, the losses of step k are read while step k+1 runs
        for i in range(steps):
            step_async(i)
            if i > 0:
                losses = trn.wait_losses()
        losses = trn.wait_losses()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_u8_ms = None
        if not video:
            # the same end-to-end loop from DECODED BYTES (cenn_trainer_step_images_u8_host_async): what the image loader holds before its
            # division by 255; rescale, centre clone and mean fill (train.lua:286-290) run on the device
            u8 = []
            for _ in range(n_batches):
                img = drng.integers(0, 256, (B, 3, 128, 128)).astype(np.uint8)
                pu = pin(img.nbytes)
                hu = np.ctypeslib.as_array((C.c_uint8 * img.size).from_address(pu.value)); hu[:] = img.ravel()
                u8.append(hu)
            for i in range(4):
                trn.step_images_u8_host_async(u8[i % n_batches])
                if i > 0:
                    trn.wait_losses()
            trn.wait_losses()
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                trn.step_images_u8_host_async(u8[i % n_batches])
                if i > 0:
                    trn.wait_losses()
            trn.wait_losses()
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            e2e_u8_ms = (time.perf_counter() - t0) * 1e3
        # ---- per-op CUDA-event profile for the rooflines of the dominant kernels (every rank runs it: it contains the exchanges)
        prof = trn.profile_step(host[0][2].ptr, host[0][3].ptr, host[0][6], repeats=3) if want_profile else None
        torch.cuda.synchronize()
    t_ms = torch.tensor([ms, e2e_ms, e2e_u8_ms if e2e_u8_ms is not None else 0.0], device="cuda")
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    out = {"ms": float(t_ms[0].item()), "e2e_ms": float(t_ms[1].item()), "e2e_u8_ms": float(t_ms[2].item()) if e2e_u8_ms is not None else None, "losses": losses, "launches": int(launches1.value - launches0.value),
           "clocks": clocks, "prof": prof, "h2d_bytes": int(host[0][4]), "nG": nG, "nD": nD, "B": B, "steps": steps}
    # the step's CUDA graph holds the NCCL communicator: release the executor (and this workload's buffers) before anything else
    trn.close()
    del host, clips
    for p in pinned:
        try:
            api.cenn_host_free(st, p)
        except Exception:
            pass
    return out


def run_oplevel(args, torch, api, st, video, local_rank):
    """The op-level drop-in path (what train*.lua hit through lua/cenn.lua without edits): nn.* mirror modules calling the THNN-table
    entries one by one in precision mode CENN_BF16 (tensor-core kernels behind NCHW fp32 tensors), four criterion:forward host
    syncs per step, host batches copied in by the closures (train.lua:292-296).  Single GPU; wall-clock timing because the path
    is host-synchronous by construction."""
    import video_filler_b200.tensor as T
    from video_filler_b200 import synth, train
    T.set_precision("bf16")
    B = args.batch
    opt = opt_for(B, "video" if video else "image")
    if video:
        opt["wtgdl"] = args.wtgdl
    trn = train.ClosureTrainer(opt, seed=1234)
    drng = np.random.default_rng(1000)
    batches = [synth.video_batch(B, 12, 128, opt["maskValue"], drng) if video else synth.image_batch(B, 128, 4, drng) for _ in range(2)]
    launches0 = C.c_int64()
    for i in range(max(3, args.warmup)):
        trn.step(*batches[i % 2])
    torch.cuda.synchronize()
    api.cenn_kernel_launches(st, C.byref(launches0))
    sampler = ClockSampler(local_rank); sampler.start(); sampler.mark_begin()
    t0 = time.perf_counter()
    for i in range(args.steps):
        losses = trn.step(*batches[i % 2])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.finish()
    launches1 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches1))
    value = B * args.steps / dt
    gflop = VIDEO_GFLOP_PER_SAMPLE if video else STEP_GFLOP_PER_SAMPLE
    hbm, tf_burst, tf_sus, peak_src = peaks()
    nbytes = int(sum(np.asarray(x).nbytes for x in batches[0]))
    line = {"metric": "train samples/sec (G+D step)", "value": value, "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (WORKLOAD_VIDEO % (args.wtgdl, B)) if video else WORKLOAD.replace("batch 256", "batch %d" % B), "global_batch": B, "parallelism": "dp1",
                       "path": "op-level drop-in path: one THNN-table call per module phase (cenn_<Op>_*), NCHW fp32 tensors converted to NHWC bf16 inside every call, "
                               "4-5 criterion:forward host syncs per step (the reference's call pattern, train.lua:278-410)",
                       "l2": "working set exceeds the 126 MB L2", "losses_last_step": {k: (round(v, 5) if v is not None else None) for k, v in losses.items()}},
            "gpu_launches": int(launches1.value - launches0.value), "clocks": clocks, "step_tflops": gflop * 1e-3 * value,
            "step_frac_of_bf16_burst": gflop * 1e-3 * value / tf_burst,
            # the same call IS the end-to-end path: host batches in (H2D inside the closures), loss numbers out every step
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 16}}
    print(json.dumps(line))
    sys.stdout.flush()


def class_rooflines(prof, hbm, peak_src):
    """Achieved GB/s per class of bandwidth kernel = algorithmic bytes (stated per op by the executor) / CUDA-event time."""
    if not prof or prof.get("bytes") is None:
        return None
    classes = {"bn_fwd": ("bn_fin_apply", "bn_apply_act"), "bn_bwd": ("bn_bwd_reduce", "bn_bwd_apply", "bn_bwd_fused", "bn_bwd_coef"),
               "act_bwd": ("act_bwd", "act_bwd4"), "loss_blend": ("blend_overlap", "blend_masked", "gdl_loss"), "adam": ("adam", "adam_early")}
    out = {}
    for cls, names in classes.items():
        sel = [i for i, n in enumerate(prof["names"]) if n in names]
        ms = float(sum(prof["ms"][i] for i in sel)); by = float(sum(prof["bytes"][i] for i in sel))
        if ms > 0 and by > 0:
            gbs = by / (ms * 1e-3) / 1e9
            out[cls] = {"achieved": round(gbs, 1), "frac": round(gbs / hbm, 4), "ms": round(ms, 4), "launches": len(sel), "algorithmic_mb": round(by / 1e6, 2)}
    # the largest single launch of the BN / activation chain (the small layers are L2-resident and launch-latency bound)
    best = None
    for i, n in enumerate(prof["names"]):
        if n.startswith("bn_") or n.startswith("act_bwd"):
            by, ms = float(prof["bytes"][i]), float(prof["ms"][i])
            if by >= 32e6 and ms > 0:
                gbs = by / (ms * 1e-3) / 1e9
                if best is None or gbs > best["achieved"]:
                    best = {"kernel": n, "achieved": round(gbs, 1), "frac": round(gbs / hbm, 4), "algorithmic_mb": round(by / 1e6, 2), "ms": round(ms, 4)}
    return {"bound": "hbm", "peak": hbm, "unit": "GB/s", "peak_source": peak_src + " hbm_gbs", "classes": out, "best_large_layer_launch": best}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default 256 image / 64 video)")
    ap.add_argument("--workload", default="image", choices=["image", "video", "deeper", "infer"],
                    help="image = BASELINE.json configs[1] (the headline); video = configs[2] per GPU; deeper = configs[3] (256x256, global batch 256); infer = configs[4] sweep")
    ap.add_argument("--wtgdl", type=float, default=0.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-video-block", action="store_true", help="skip the `video` sub-block (cfg3 at the same N) of the image line")
    ap.add_argument("--no-local-bn-block", action="store_true", help="N > 1: skip the `local_bn` sub-block (the same step with per-rank BN statistics)")
    ap.add_argument("--path", default="fused", choices=["fused", "oplevel"],
                    help="fused = whole-step executor (cenn_trainer_*); oplevel = the drop-in THNN-table path an unchanged script hits (ClosureTrainer, precision bf16)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    deeper = args.workload == "deeper"
    video = args.workload == "video" or deeper
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.batch is None and args.workload != "infer":
        args.batch = (256 // world_env) if deeper else (64 if video else 256)
    if deeper and args.wtgdl == 0.5:
        args.wtgdl = 0.0                 # train_deepernet.lua:27 default

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import video_filler_b200.tensor as T
    T.state(local_rank)
    api, st = T.api(), T.state()
    stream = torch.cuda.Stream(priority=-1)      # the step's critical path; the executor's side streams run at lowest priority
    api.cenn_set_stream(st, C.c_void_p(stream.cuda_stream))

    if args.workload == "infer":
        run_infer(args, torch, dist, api, st, stream, rank, local_rank, world)
        return
    if args.path == "oplevel":
        run_oplevel(args, torch, api, st, video, local_rank)
        return
    B = args.batch
    if world > 1:
        # library-owned NCCL communicator: rank 0 creates the id, torch.distributed carries it to the other ranks
        idbuf = np.zeros(128, np.uint8)
        if rank == 0:
            api.cenn_dist_unique_id(idbuf.ctypes.data_as(C.c_void_p))
        idt = torch.from_numpy(idbuf).cuda()
        dist.broadcast(idt, src=0)
        idbuf = idt.cpu().numpy()
        api.cenn_dist_init(st, idbuf.ctypes.data_as(C.c_void_p), world, rank)

    m = measure_train(args, torch, dist, api, st, stream, rank, local_rank, world, video, B, args.steps, args.warmup, fine=256 if deeper else 128)
    vid = None
    if not video and not args.no_video_block:
        # the config the >= 7x scaling target is quoted on (BASELINE.json configs[2]): 64 clips of 12 stacked channels per GPU, at the same N
        vid = measure_train(args, torch, dist, api, st, stream, rank, local_rank, world, True, 64, args.steps, args.warmup, want_profile=False)
    loc = locv = None
    if world > 1 and not deeper and not args.no_local_bn_block:
        # the same workload(s) with per-rank BN statistics (cfg.bn_local: the usual distributed-data-parallel semantics -- every rank normalises
        # with its own 256 / 64 samples, as one reference process at that batchSize does; no statistics cross the ranks).  The headline `value`
        # stays the global-batch-statistics step, which equals one executor at N x batchSize (tests/test_dp_multi_gpu.py).
        loc = measure_train(args, torch, dist, api, st, stream, rank, local_rank, world, video, B, args.steps, args.warmup, want_profile=False, bn_local=1)
        if vid is not None:
            locv = measure_train(args, torch, dist, api, st, stream, rank, local_rank, world, True, 64, args.steps, args.warmup, want_profile=False, bn_local=1)

    def teardown():
        # ordered teardown, then a NORMAL interpreter exit (atexit hooks run): executors are already closed by measure_train
        if dist:
            api.cenn_dist_shutdown(st)
            dist.barrier()
            dist.destroy_process_group()
        sys.stdout.flush()

    if rank != 0:
        teardown()
        return
    hbm, tf_burst, tf_sus, peak_src = peaks()
    gflop_per_sample = DEEPER_GFLOP_PER_SAMPLE if deeper else (VIDEO_GFLOP_PER_SAMPLE if video else STEP_GFLOP_PER_SAMPLE)
    ms, e2e_ms, prof, losses = m["ms"], m["e2e_ms"], m["prof"], m["losses"]
    value = B * world * args.steps / (ms / 1e3)
    line = {
        "metric": "train samples/sec (G+D step)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(4, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if deeper else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (WORKLOAD_DEEPER % (args.wtgdl, B)) if deeper else ((WORKLOAD_VIDEO % (args.wtgdl, B)) if video else WORKLOAD.replace("batch 256", "batch %d" % B)), "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (>3 GB of activations, weights and optimizer state) exceeds the 126 MB L2; no explicit flush",
                   "timed_region_s": round(ms / 1e3, 4),
                   "losses_last_step": {k: round(v, 5) for k, v in losses.items()} if losses else None},
        "gpu_launches": m["launches"],
        "clocks": m["clocks"],
        "step_tflops": gflop_per_sample * 1e-3 * value,
        "step_frac_of_bf16_sustained": gflop_per_sample * 1e-3 * value / (tf_sus * world),
        "step_frac_of_bf16_burst": gflop_per_sample * 1e-3 * value / (tf_burst * world),
    }
    line["e2e"] = {"value": B * world * args.steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": 32,
                   "call": "cenn_trainer_step_clips_host_async (frames + one mask plane per clip)" if video else "cenn_trainer_step_host_async (the loader's FloatTensor batches: real_ctx + real_center)"}
    if m.get("e2e_u8_ms"):
        line["e2e_bytes"] = {"value": B * world * args.steps / (m["e2e_u8_ms"] / 1e3), "unit": "samples/s", "h2d_bytes_per_step": B * 3 * 128 * 128, "d2h_bytes_per_step": 32,
                             "call": "cenn_trainer_step_images_u8_host_async (decoded bytes in; rescale, centre clone and mean fill on the device)"}
    if vid is not None:
        vvalue = 64 * world * args.steps / (vid["ms"] / 1e3)
        line["video"] = {"workload": WORKLOAD_VIDEO % (args.wtgdl, 64), "value": vvalue, "unit": "samples/s", "ms_per_step": vid["ms"] / args.steps,
                         "global_batch": 64 * world, "step_tflops": VIDEO_GFLOP_PER_SAMPLE * 1e-3 * vvalue, "gpu_launches": vid["launches"],
                         "e2e": {"value": 64 * world * args.steps / (vid["e2e_ms"] / 1e3), "unit": "samples/s", "h2d_bytes_per_step": vid["h2d_bytes"], "d2h_bytes_per_step": 32},
                         "clocks": vid["clocks"]}
    if loc is not None:
        line["local_bn"] = {"semantics": "per-rank BN batch statistics (cfg.bn_local = 1): no statistics exchange, gradients and losses all-reduced",
                            "value": B * world * args.steps / (loc["ms"] / 1e3), "unit": "samples/s", "ms_per_step": loc["ms"] / args.steps,
                            "e2e": {"value": B * world * args.steps / (loc["e2e_ms"] / 1e3), "unit": "samples/s", "h2d_bytes_per_step": loc["h2d_bytes"], "d2h_bytes_per_step": 32},
                            "gpu_launches": loc["launches"]}
        if locv is not None:
            line["local_bn"]["video"] = {"value": 64 * world * args.steps / (locv["ms"] / 1e3), "unit": "samples/s", "ms_per_step": locv["ms"] / args.steps,
                                         "e2e": {"value": 64 * world * args.steps / (locv["e2e_ms"] / 1e3), "unit": "samples/s"}}
    if prof:
        tc_ms, tc_flops, tc_n = prof["tc_ms"], prof["tc_flops"], prof["tc_launches"]
        ach = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")     # dram__bytes_read + dram__bytes_write of the conv launches of one step (ncu --set full)
        if os.path.exists(tp) and not video and B == 256:
            try:
                traffic = json.load(open(tp)).get("conv_launch_dram_bytes_mean")
            except Exception:
                traffic = None
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst, "frac_of_sustained_peak": ach / tf_sus, "traffic": traffic,
                            "kernel": "tc::gather_gemm_kernel / tc::patch_dgrad_kernel / tc::wgrad_gemm_kernel (all conv fprop/dgrad/wgrad launches of one step; algorithmic FLOPs / summed CUDA-event durations)",
                            "launches_per_step": tc_n, "share_of_step": tc_ms / prof["total_ms"], "peak_source": peak_src + " bf16_tflops (burst: each launch is timed alone)",
                            "by_op_ms": prof["by_op"]}
        best = prof.get("best_tc")
        if best:
            line["roofline"]["best_launch"] = best
        cr = class_rooflines(prof, hbm, peak_src)
        if cr:
            line["roofline_hbm"] = cr
        adam_ms = prof["by_op"].get("adam", 0.0) + prof["by_op"].get("adam_early", 0.0)
        if adam_ms > 0:
            gbs = 30.0 * (m["nG"] + m["nD"]) / (adam_ms * 1e-3) / 1e9
            line.setdefault("roofline_hbm", {"bound": "hbm", "peak": hbm, "unit": "GB/s", "peak_source": peak_src + " hbm_gbs"}).update(
                {"kernel": "nhwc::adam_bf16_kernel", "achieved": gbs, "frac": gbs / hbm, "traffic": None, "bytes_per_param": 30, "params": int(m["nG"] + m["nD"])})
    if not args.no_cpu_baseline and world == 1 and not video:
        threads = host_threads()
        rate, sec, n_thr = cpu_port_step_rate(64, 3, 1, threads)
        rate1, sec1, _ = cpu_port_step_rate(16, 1, 1, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": n_thr, "kind": "port",
                                "sample": "BASELINE.json configs[0] (batch 64): 1 warm-up + 3 timed steps of the oracle step sequence on the PyTorch-CPU engine "
                                          "(oneDNN/MKL, fp32; faster than Torch7's im2col+sgemm gpu=0 path), %.2f s/step on %d threads" % (sec, n_thr),
                                "one_thread": {"value": rate1, "unit": "samples/s", "cores": 1,
                                               "sample": "torch.setnumthreads(1) as train.lua:47: 1 warm-up + 1 timed 16-sample step, %.2f s" % sec1}}
    print(json.dumps(line))
    teardown()


if __name__ == "__main__":
    main()
