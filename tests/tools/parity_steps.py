"""100-step loss parity at the CPU config (BASELINE.json configs[0]: inpaintCenter, batch 64, nBottleneck 4000, fp32).

north_star: "losses after 100 steps within 1 %".  The oracle's 100 steps take ~20 minutes of CPU, so they are run ONCE
where CPU time is free and committed as a fixture; the GPU side replays the same seeded batches through the fused
executor (BF16 tensor-core mode) and compares.

    python tests/tools/parity_steps.py --make-golden      # oracle (fp32, the reference's gpu=0 arithmetic) -> tests/golden/losses_100.npz
    python tests/tools/parity_steps.py                    # executor on cuda:0 vs the fixture; writes gpurun_out/parity_steps.json

Batches: video_filler_b200.synth.image_batch with numpy PCG64 seed 1234, one fresh batch per step; weights
util.weights_init with seed 1234 (train.lua:58-67).  Both sides start from the identical flat parameter vectors.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "losses_100.npz")
NAMES = ("errD", "errG", "errG_l2", "errD_real", "errD_fake", "errG_total")


def config(batch, nB):
    return dict(batchSize=batch, nBottleneck=nB, nef=64, ngf=64, ndf=64)


def batches(batch, steps, seed=1234):
    from video_filler_b200 import synth
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        yield synth.image_batch(batch, 128, 4, rng)


def init_params(opt, seed=1234):
    from video_filler_b200 import util
    rng = np.random.default_rng(seed)
    pG = util.params_flat(util.weights_init(util.describe_netG(opt), rng))
    pD = util.params_flat(util.weights_init(util.describe_netD(opt), rng))
    return pG, pD


def make_golden(args):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import nets as onets
    from oracle import step as ostep
    from video_filler_b200 import models
    opt = models.default_opt("image", **config(args.batch, args.nBottleneck))
    orc = ostep.StepOracle(onets.default_opt("image", **config(args.batch, args.nBottleneck)), seed=1, dtype=np.float32)
    pG, pD = init_params(opt)
    orc.pG[:] = pG
    orc.pD[:] = pD
    hist = []
    for it, (ctx, center) in enumerate(batches(args.batch, args.steps)):
        lo = orc.step(ctx.astype(np.float32), center.astype(np.float32))
        hist.append([lo[k] for k in NAMES])
        if it % 10 == 0:
            print(it, hist[-1], flush=True)
    np.savez(GOLDEN, losses=np.array(hist, np.float64), names=np.array(NAMES), batch=args.batch, nBottleneck=args.nBottleneck, steps=args.steps)
    print("wrote", GOLDEN)


def run_control(args):
    """CONTROL for the adversarial terms: how far do errD / errG of two CPU runs of the SAME oracle step
    sequence drift apart over 100 steps when they differ only at rounding level?  Arms, all from the golden fixture's seeded
    batches and initial weights:
      torch_fp32   the oracle on the PyTorch-CPU engine, fp32 (other summation order than the numpy fp32 golden run)
      perturbed    the same, initial weights multiplied by (1 +- 2^-23) (one fp32 ulp)
      fp64         the same in fp64
    Written to profiles/r2_parity_control.json with the same summary the executor gets."""
    import time
    from oracle import nets as onets
    from oracle import step as ostep
    from oracle import torch_engine
    from video_filler_b200 import models
    torch_engine.enable(os.cpu_count() or 1)
    g = np.load(GOLDEN)
    batch, nB, n = int(g["batch"]), int(g["nBottleneck"]), min(int(g["steps"]), args.steps)
    gold = g["losses"][:n]
    opt = models.default_opt("image", **config(batch, nB))
    pG, pD = init_params(opt)
    out = {"steps": n, "batch": batch, "nBottleneck": nB, "arms": {}}
    for arm in args.arms.split(","):
        dt = np.float64 if arm == "fp64" else np.float32
        orc = ostep.StepOracle(onets.default_opt("image", **config(batch, nB)), seed=1, dtype=dt)
        orc.pG[:] = pG; orc.pD[:] = pD
        if arm == "perturbed":
            prng = np.random.default_rng(99)
            orc.pG *= (1 + np.float32(2.0 ** -23) * prng.choice([-1.0, 1.0], orc.pG.size)).astype(np.float32)
            orc.pD *= (1 + np.float32(2.0 ** -23) * prng.choice([-1.0, 1.0], orc.pD.size)).astype(np.float32)
        hist, t0 = [], time.time()
        for it, (ctx, center) in enumerate(batches(batch, n)):
            lo = orc.step(ctx.astype(dt), center.astype(dt))
            hist.append([lo[k] for k in NAMES])
            if it % 10 == 0:
                print(arm, it, ["%.4f" % v for v in hist[-1]], "%.0fs" % (time.time() - t0), flush=True)
        hist = np.array(hist, np.float64)
        out["arms"][arm] = {"vs_numpy_fp32_golden": summarize(hist, gold), "losses": hist.tolist()}
        with open(os.path.join(ROOT, "profiles", "r2_parity_control.json"), "w") as f:
            json.dump(out, f)
    print(json.dumps({a: {k: v["vs_numpy_fp32_golden"][k]["rel_at_last_step"] for k in NAMES} for a, v in out["arms"].items()}, indent=1))


def run_executor(steps=None):
    """Executor losses on the fixture's batches: returns (ours [steps, 6], golden [steps, 6])."""
    g = np.load(GOLDEN)
    batch, nB, n = int(g["batch"]), int(g["nBottleneck"]), int(g["steps"])
    steps = min(steps or n, n)
    import video_filler_b200.tensor as T
    from video_filler_b200 import models, train
    T.state(0)
    opt = models.default_opt("image", **config(batch, nB))
    trn = train.FusedTrainer(opt, precision="bf16")
    pG, pD = init_params(opt)
    trn.set_params(0, pG)
    trn.set_params(1, pD)
    hist = []
    for ctx, center in batches(batch, steps):
        lg = trn.step_host(ctx, center)
        hist.append([lg[k] for k in NAMES])
    trn.close()
    return np.array(hist, np.float64), g["losses"][:steps]


def summarize(ours, gold):
    rel = np.abs(ours - gold) / np.maximum(np.abs(gold), 1e-12)
    last = slice(-10, None)
    out = {"steps": int(len(ours))}
    for j, k in enumerate(NAMES):
        out[k] = {"max_rel_all_steps": float(rel[:, j].max()), "rel_at_last_step": float(rel[-1, j]),
                  "rel_of_mean_last10": float(abs(ours[last, j].mean() - gold[last, j].mean()) / abs(gold[last, j].mean())),
                  "ours_last": float(ours[-1, j]), "oracle_last": float(gold[-1, j])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--make-golden", action="store_true")
    ap.add_argument("--control", action="store_true", help="CPU control arms for the adversarial terms -> profiles/r2_parity_control.json")
    ap.add_argument("--arms", default="torch_fp32,perturbed,fp64")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--nBottleneck", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    if args.make_golden:
        return make_golden(args)
    if args.control:
        return run_control(args)
    ours, gold = run_executor()
    s = summarize(ours, gold)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_steps.json"), "w") as f:
        json.dump({"summary": s, "ours": ours.tolist(), "oracle": gold.tolist(), "names": NAMES}, f)
    print(json.dumps(s, indent=1))


if __name__ == "__main__":
    main()
