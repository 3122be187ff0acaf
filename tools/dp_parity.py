"""Correctness of the REAL data-parallel data plane (run under torchrun, one rank per GPU; tests/test_dp_multi_gpu.py drives it).

Every rank runs the fused executor with world_size = N on its slice of one global batch: library-owned NCCL communicators,
the peer-mailbox BN exchanges (bn_finalize_xr / bn_bwd_coef_xr), bf16 gradient buckets and the Adam kernels reading the
bucket sums -- everything captured in the step's CUDA graph from the third step on.  Rank 0 then runs ONE executor at the global
batch (world_size 1) on the same inputs and compares: losses, BN running statistics, parameters; all ranks compare their
parameters with each other (replicas must stay bit-identical).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py --variant image
    env knobs exercised by the test: CENN_NO_XR=1 (NCCL-only statistics), CENN_FP32_BUCKETS=1 (fp32 gradient buckets)
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.dot(a, b) / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="image")
    ap.add_argument("--per-rank", type=int, default=8)
    ap.add_argument("--nB", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--bn-local", action="store_true", help="cfg.bn_local = 1: per-rank BN statistics (DDP semantics); checked against one single-GPU executor per shard")
    ap.add_argument("--branches", action="store_true", help="image variant with noiseGen + conditionAdv (train.lua:109-124,158-180)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import video_filler_b200.tensor as T
    from video_filler_b200 import models, synth, train, util
    T.state(local_rank)
    api, st = T.api(), T.state()
    idbuf = np.zeros(128, np.uint8)
    if rank == 0:
        api.cenn_dist_unique_id(idbuf.ctypes.data_as(C.c_void_p))
    idt = torch.from_numpy(idbuf).cuda()
    dist.broadcast(idt, src=0)
    idbuf = idt.cpu().numpy()
    api.cenn_dist_init(st, idbuf.ctypes.data_as(C.c_void_p), world, rank)

    Bl, Bg = args.per_rank, args.per_rank * world
    kw = dict(nBottleneck=args.nB, nef=64, ngf=64, ndf=64)
    if args.variant == "video":
        kw.update(predLen=2, wtgdl=0.5)
    if args.branches:
        kw.update(noiseGen=1, nz=100, conditionAdv=1)
    opt_l = models.default_opt(args.variant, batchSize=Bl, **kw)
    opt_g = models.default_opt(args.variant, batchSize=Bg, **kw)
    rng = np.random.default_rng(7)
    pG = util.params_flat(util.weights_init(util.describe_netG(opt_g), rng))
    pD = util.params_flat(util.weights_init(util.describe_netD(opt_g), rng))
    drng = np.random.default_rng(11)
    batches = []
    noises = []
    for _ in range(args.steps):
        batches.append(synth.image_batch(Bg, 128, 4, drng) if args.variant == "image" else synth.video_batch(Bg, 6, 128, opt_g["maskValue"], drng))
        noises.append(drng.uniform(-1, 1, (Bg, 100)).astype(np.float32) if args.branches else None)       # each rank takes its rows of the global draw

    if args.bn_local:
        return bn_local_check(args, dist, api, st, train, opt_l, pG, pD, batches, rank, world, Bl)
    trn = train.FusedTrainer(opt_l, precision="bf16", world_size=world, rank=rank)
    trn.set_params(0, pG); trn.set_params(1, pD)
    sl = slice(Bl * rank, Bl * rank + Bl)
    losses, bn_after_1, params_after_1 = [], None, None
    for i, b in enumerate(batches):
        losses.append(trn.step_host(*[np.ascontiguousarray(x[sl]) for x in b], noise=None if noises[i] is None else noises[i][sl]))
        if i == 0:
            bn_after_1 = (trn.get_bn_stats(0), trn.get_bn_stats(1))
            params_after_1 = (trn.get_params(0), trn.get_params(1))
    pGr, pDr = trn.get_params(0), trn.get_params(1)
    # eval-mode generator forward of one fixed input with THIS rank's operand copies (the bf16 weights the kernels read: with the sharded
    # reduce + Adam of the big blocks they are written by the peers, and get_params reads the fp32 shards from their owners)
    if args.branches:
        trn.set_noise(noises[0][:Bl])               # the probe must see the same noise on every rank (each rank trained on its own rows)
    probe = trn.generator_forward(np.ascontiguousarray(batches[0][0][:Bl]))
    digest = hashlib.sha256(pGr.tobytes() + pDr.tobytes() + trn.get_bn_stats(0).tobytes() + trn.get_bn_stats(1).tobytes() + probe.tobytes()).hexdigest()
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    all_losses = [None] * world
    dist.all_gather_object(all_losses, losses)
    trn.close()
    torch.cuda.synchronize()
    dist.barrier()

    out = {"variant": args.variant, "branches": bool(args.branches), "world": world, "per_rank": Bl, "steps": args.steps, "env": {k: os.environ.get(k) for k in ("CENN_NO_XR", "CENN_FP32_BUCKETS")},
           "replicas_bit_identical": len(set(digests)) == 1, "ok": True, "checks": {}}
    fail = []
    if rank == 0:
        if not out["replicas_bit_identical"]:
            fail.append("parameters / BN statistics differ between ranks after %d steps" % args.steps)
        for r in range(1, world):
            for i in range(args.steps):
                for k, v in all_losses[0][i].items():
                    if v != all_losses[r][i][k]:
                        fail.append("loss %s of step %d differs between rank 0 and rank %d: %r vs %r" % (k, i, r, v, all_losses[r][i][k]))
        ref = train.FusedTrainer(opt_g, precision="bf16")
        ref.set_params(0, pG); ref.set_params(1, pD)
        ref_losses = []
        for i, b in enumerate(batches):
            ref_losses.append(ref.step_host(*b, noise=noises[i]))
            if i == 0:
                chk = out["checks"]
                chk["bn_G_rel"] = rel_err(bn_after_1[0], ref.get_bn_stats(0)); chk["bn_D_rel"] = rel_err(bn_after_1[1], ref.get_bn_stats(1))
                dG, dGr = params_after_1[0] - pG, ref.get_params(0) - pG
                dD, dDr = params_after_1[1] - pD, ref.get_params(1) - pD
                chk["adam_update_cos_G"] = cos(dG, dGr); chk["adam_update_cos_D"] = cos(dD, dDr)
                if chk["bn_G_rel"] > 5e-3 or chk["bn_D_rel"] > 5e-3:
                    fail.append("BN running statistics after step 1: rel %.3g (G) %.3g (D) > 5e-3" % (chk["bn_G_rel"], chk["bn_D_rel"]))
                # Adam's first update is +-lr per weight (sign of the gradient): direction of the whole update vector
                if chk["adam_update_cos_G"] < 0.9 or chk["adam_update_cos_D"] < 0.9:
                    fail.append("first Adam update: cos %.3f (G) %.3f (D) < 0.9" % (chk["adam_update_cos_G"], chk["adam_update_cos_D"]))
        ref.close()
        out["losses_dp"], out["losses_ref"] = all_losses[0], ref_losses
        # step 1 is a pure function of inputs and initial weights: 1e-2 (bf16 summation order); later steps inherit the
        # run-to-run spread of the GAN game (tests/test_dp_gpu.py quantifies it): finite, and the L2 term within 5e-2
        tol1 = {"errD_real": 1e-2, "errG_l2": 1e-2, "errG_total": 1e-2, "errD": 1e-2, "errD_fake": 2e-2, "errG": 4e-2, "errG_gdl": 1e-2}
        for k, tol in tol1.items():
            a, b = all_losses[0][0][k], ref_losses[0][k]
            out["checks"]["step1_" + k] = [a, b]
            if abs(a - b) > tol * max(abs(b), 1e-3):
                fail.append("step 1 %s: %r (dp) vs %r (global batch), tol %g" % (k, a, b, tol))
        for i in range(1, args.steps):
            a, b = all_losses[0][i]["errG_l2"], ref_losses[i]["errG_l2"]
            if not np.isfinite(list(all_losses[0][i].values())).all() or abs(a - b) > 5e-2 * abs(b):
                fail.append("step %d errG_l2: %r vs %r" % (i + 1, a, b))
        out["ok"] = not fail
        out["failures"] = fail
        print("DP_PARITY " + json.dumps(out))
    api.cenn_dist_shutdown(st)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and fail:
        sys.exit(1)


def bn_local_check(args, dist, api, st, train, opt_l, pG, pD, batches, rank, world, Bl):
    """cfg.bn_local: rank r normalises with its own shard, so after ONE step from common weights
      * rank r's BN running statistics equal those of a single-GPU executor (world 1, batch Bl) run on shard r,
      * the all-reduced losses are the mean over the shards of the single-GPU losses (every loss is a batch mean),
      * the all-reduced gradients are the mean of the single-GPU gradients;
    and the replicas' parameters stay bit-identical over the following steps (the statistics do not)."""
    trn = train.FusedTrainer(opt_l, precision="bf16", world_size=world, rank=rank, bn_local=1)
    trn.set_params(0, pG); trn.set_params(1, pD)
    sl = slice(Bl * rank, Bl * rank + Bl)
    losses, bn1, g1 = [], None, None
    for i, b in enumerate(batches):
        losses.append(trn.step_host(*[np.ascontiguousarray(x[sl]) for x in b]))
        if i == 0:
            bn1 = (trn.get_bn_stats(0), trn.get_bn_stats(1)); g1 = (trn.get_grads(0), trn.get_grads(1))
    digest = hashlib.sha256(trn.get_params(0).tobytes() + trn.get_params(1).tobytes()).hexdigest()
    digests, all_losses = [None] * world, [None] * world
    dist.all_gather_object(digests, digest)
    dist.all_gather_object(all_losses, losses)
    trn.close()
    # this rank's shard on a plain single-GPU executor
    ref = train.FusedTrainer(opt_l, precision="bf16")
    ref.set_params(0, pG); ref.set_params(1, pD)
    rl = ref.step_host(*[np.ascontiguousarray(x[sl]) for x in batches[0]])
    rbn, rg = (ref.get_bn_stats(0), ref.get_bn_stats(1)), (ref.get_grads(0), ref.get_grads(1))
    ref.close()
    mine = {"bn_G_rel": rel_err(bn1[0], rbn[0]), "bn_D_rel": rel_err(bn1[1], rbn[1]), "loss": rl, "gG": rg[0], "gD": rg[1]}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    out = {"variant": args.variant, "bn_local": True, "world": world, "per_rank": Bl, "replicas_bit_identical": len(set(digests)) == 1, "checks": {}}
    fail = []
    if rank == 0:
        if not out["replicas_bit_identical"]:
            fail.append("parameters differ between ranks after %d steps" % len(batches))
        for r in range(world):
            out["checks"]["bn_rel_rank%d" % r] = [gathered[r]["bn_G_rel"], gathered[r]["bn_D_rel"]]
            if gathered[r]["bn_G_rel"] > 5e-3 or gathered[r]["bn_D_rel"] > 5e-3:
                fail.append("rank %d BN running statistics differ from the single-GPU run on its shard: %r" % (r, out["checks"]["bn_rel_rank%d" % r]))
            if all_losses[r] != all_losses[0]:
                fail.append("losses differ between rank 0 and rank %d" % r)
        # Exactly comparable are the quantities computed BEFORE the first Adam update (fDx: the discriminator losses and gradients, the L2 term of
        # the unchanged generator); fGx runs behind D's update, which uses the AVERAGED gradient here and the shard's own in the single-GPU runs,
        # so errG and G's gradients only have to stay close.  Bounds: run-to-run spread of this 8-sample-per-rank GAN step (fp32 atomics order
        # flips bf16 roundings; BN over 8 samples amplifies them) -- the same spread tests/test_dp_gpu.py quantifies.
        for k, tol in (("errD_real", 5e-3), ("errD_fake", 1e-2), ("errD", 1e-2), ("errG_l2", 5e-3), ("errG", 4e-2), ("errG_total", 1e-2)):
            m = float(np.mean([gathered[r]["loss"][k] for r in range(world)]))
            out["checks"]["loss_" + k] = [all_losses[0][0][k], m]
            if abs(all_losses[0][0][k] - m) > tol * max(abs(m), 1e-3):
                fail.append("step-1 %s %r is not the mean of the shard losses %r (tol %g)" % (k, all_losses[0][0][k], m, tol))
        for name, g, key, cmin in (("G", g1[0], "gG", 0.95), ("D", g1[1], "gD", 0.98)):
            m = np.mean([gathered[r][key] for r in range(world)], axis=0)
            out["checks"]["grad_%s" % name] = [cos(g, m), float(np.linalg.norm(g) / max(np.linalg.norm(m), 1e-30))]
            if cos(g, m) < cmin or abs(np.linalg.norm(g) / np.linalg.norm(m) - 1) > 3e-2:
                fail.append("gradients of %s are not the mean of the shard gradients: %r" % (name, out["checks"]["grad_%s" % name]))
        for i in range(len(batches)):
            if not np.isfinite(list(all_losses[0][i].values())).all():
                fail.append("non-finite loss at step %d" % (i + 1))
        out["ok"] = not fail
        out["failures"] = fail
        print("DP_PARITY " + json.dumps(out))
    api.cenn_dist_shutdown(st)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and fail:
        sys.exit(1)


if __name__ == "__main__":
    main()
