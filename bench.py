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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None
        self.t_begin = None

    def mark_begin(self):
        """Samples taken from now on are inside the timed region."""
        self.t_begin = time.time()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, reasons, smax = [], set(), None
        inside = [v for (ts, v) in self.samples if self.t_begin is None or ts >= self.t_begin]
        if not inside:                      # timed region shorter than one sampling period: use the last sample before it ended
            inside = [v for (_, v) in self.samples[-1:]]
        for s in inside:
            try:
                sm.append(float(s[0])); smax = float(s[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_step_rate(batch, steps, warmup, threads):
    """The reference's gpu=0 path restated (oracle port, fp32) on the host cores: samples/s over `steps` steps."""
    import torch
    torch.set_num_threads(threads)
    from oracle import nets as onets
    from oracle import step as ostep
    orc = ostep.StepOracle(onets.default_opt("image", batchSize=batch), seed=1234, dtype=np.float32)
    rng = np.random.default_rng(1234)
    batches = [orc.synth_batch(rng) for _ in range(min(2, steps + warmup))]
    for i in range(warmup):
        orc.step(*batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps):
        orc.step(*batches[i % len(batches)])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample of the 256-sample step: the port runs ~8 samples/s on 16 cores, so size each step for ~90 s in total
    batch = int(max(2, min(32, 90 * 8 // max(1, args.steps + args.warmup))))
    rate, sec = cpu_port_step_rate(batch, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "train samples/sec (G+D step)", "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "each step = %d samples of the 256-sample batch" % batch},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": "oracle port (numpy im2col+sgemm restatement of the Torch7 gpu=0 path), %d-sample steps" % batch},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def wrap_device(ptr, count, dtype, torch):
    """A torch view of a raw device buffer (for torch.distributed collectives on the executor's buffers)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8" if dtype == "f8" else "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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
    from video_filler_b200 import synth, train, util
    T.state(local_rank)
    api, st = T.api(), T.state()
    stream = torch.cuda.Stream(priority=-1)      # the step's critical path; the executor's side streams run at lowest priority
    api.cenn_set_stream(st, C.c_void_p(stream.cuda_stream))

    B = args.batch
    opt = opt_for(B)
    if world > 1:
        # library-owned NCCL communicator: rank 0 creates the id, torch.distributed carries it to the other ranks
        idbuf = np.zeros(128, np.uint8)
        if rank == 0:
            api.cenn_dist_unique_id(idbuf.ctypes.data_as(C.c_void_p))
        idt = torch.from_numpy(idbuf).cuda()
        dist.broadcast(idt, src=0)
        idbuf = idt.cpu().numpy()
        api.cenn_dist_init(st, idbuf.ctypes.data_as(C.c_void_p), world, rank)

    trn = train.FusedTrainer(opt, precision="bf16", world_size=world, rank=rank)
    # identical random-init weights on every rank (parameter broadcast = same seed), train.lua:58-67
    rng = np.random.default_rng(1234)
    nG, nD = trn.param_count(0), trn.param_count(1)

    trn.set_params(0, util.params_flat(util.weights_init(util.describe_netG(opt), rng)))
    trn.set_params(1, util.params_flat(util.weights_init(util.describe_netD(opt), rng)))
    assert nG == 71118691 + 2 * (64 + 128 + 256 + 512 + 4000 + 512 + 256 + 128 + 64) and nD > 2764737

    drng = np.random.default_rng(1000 + rank)
    n_batches = 2
    host = []
    for _ in range(n_batches):
        ctx, center = synth.image_batch(B, 128, 4, drng)
        pa, pb = C.c_void_p(), C.c_void_p()
        api.cenn_host_alloc(st, ctx.nbytes, C.byref(pa)); api.cenn_host_alloc(st, center.nbytes, C.byref(pb))
        ha = np.ctypeslib.as_array((C.c_float * ctx.size).from_address(pa.value)); ha[:] = ctx.ravel()
        hb = np.ctypeslib.as_array((C.c_float * center.size).from_address(pb.value)); hb[:] = center.ravel()
        da, db = T.CudaTensor.from_numpy(ctx), T.CudaTensor.from_numpy(center)
        host.append((ha, hb, da, db, ctx.nbytes + center.nbytes))

    def step_device(i):
        # world > 1: the executor all-reduces BN statistics, gradients and losses itself (NCCL, inside its CUDA graph)
        _, _, da, db, _ = host[i % n_batches]
        trn.step_device(da.ptr, db.ptr)

    losses = None
    sampler = ClockSampler(local_rank); sampler.start()
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step_device(i)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        launches0 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches0))
        sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            step_device(i)
        e1.record(stream)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.finish()
        launches1 = C.c_int64(); api.cenn_kernel_launches(st, C.byref(launches1))
        losses = trn.read_losses()
        # ---- end-to-end: host buffers in, losses out, every step (single-process API call)
        e2e_ms = None
        for i in range(2):
            trn.step_host(host[i % n_batches][0], host[i % n_batches][1])
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        # public pipelined API: every step copies ITS inputs from pinned host memory and its losses are read back;
        # the copy of step k+1 overlaps the compute of step k, the losses of step k are read while step k+1 runs
        for i in range(args.steps):
            trn.step_host_async(host[i % n_batches][0], host[i % n_batches][1])
            if i > 0:
                losses = trn.wait_losses()
        losses = trn.wait_losses()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        # ---- per-op CUDA-event profile for the roofline of the dominant kernels
        prof = trn.profile_step(host[0][2].ptr, host[0][3].ptr, repeats=3)    # every rank runs it (it contains the all-reduces)

    t_ms = torch.tensor([ms, e2e_ms], device="cuda")
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t_ms[0].item()), float(t_ms[1].item())
    def finish():
        # the step's CUDA graph holds the NCCL communicator: release the executor first, then the communicator
        nonlocal trn
        trn.close()
        trn = None
        if dist:
            api.cenn_dist_shutdown(st)
            dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)         # nothing left to do; do not depend on interpreter-exit ordering of CUDA / NCCL teardown

    if rank != 0:
        finish()
    hbm, tf_burst, tf_sus, peak_src = peaks()
    total_samples = B * world * args.steps
    value = total_samples / (ms / 1e3)
    line = {
        "metric": "train samples/sec (G+D step)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (>3 GB of activations, weights and optimizer state) exceeds the 126 MB L2; no explicit flush",
                   "losses_last_step": {k: round(v, 5) for k, v in losses.items()} if losses else None},
        "gpu_launches": int(launches1.value - launches0.value),
        "clocks": clocks,
        "step_tflops": STEP_GFLOP_PER_SAMPLE * 1e-3 * value,
        "step_frac_of_bf16_sustained": STEP_GFLOP_PER_SAMPLE * 1e-3 * value / (tf_sus * world),
    }
    if e2e_ms is not None:
        line["e2e"] = {"value": B * world * args.steps / (e2e_ms / 1e3), "unit": "samples/s",
                       "h2d_bytes_per_step": int(host[0][4]), "d2h_bytes_per_step": 32}
    else:
        line["e2e"] = None
    if prof and rank == 0:
        tc_ms, tc_flops, tc_n = prof["tc_ms"], prof["tc_flops"], prof["tc_launches"]
        ach = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus, "traffic": None,
                            "kernel": "tc::gather_gemm_kernel / tc::wgrad_gemm_kernel (all conv fprop/dgrad/wgrad launches)",
                            "launches_per_step": tc_n, "share_of_step": tc_ms / prof["total_ms"], "peak_source": peak_src + " bf16_tflops_sustained",
                            "by_op_ms": prof["by_op"]}
    if prof and rank == 0:
        # the dominant bandwidth kernel (fused Adam + bf16 operand refresh over the flat parameter vectors): 30 B per parameter
        adam_ms = prof["by_op"].get("adam", 0.0) + prof["by_op"].get("adam_early", 0.0)
        if adam_ms > 0:
            gbs = 30.0 * (nG + nD) / (adam_ms * 1e-3) / 1e9
            line["roofline_hbm"] = {"bound": "hbm", "kernel": "nhwc::adam_bf16_kernel", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                    "traffic": None, "bytes_per_param": 30, "params": int(nG + nD), "peak_source": peak_src + " hbm_gbs"}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, sec = cpu_port_step_rate(64, 1, 0, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                                "sample": "one 64-sample step of the oracle port (numpy restatement of the Torch7 gpu=0 path), %.1f s" % sec}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
