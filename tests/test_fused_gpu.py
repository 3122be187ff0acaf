"""GPU parity of the fused whole-step executor (BF16 tcgen05 path) against the fp64 oracle.
Tolerances: north_star's BF16 bound (2e-2 relative, max|a-b|/max|b|) for per-layer tensors computed from
identical inputs; looser, stated bounds where bf16 rounding compounds through the whole network."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import nets as onets
from oracle import step as ostep

pytestmark = pytest.mark.gpu


def _pair(variant, B=8, nB=256, **extra):
    from video_filler_b200 import models, train
    kw = dict(batchSize=B, nBottleneck=nB, nef=64, ngf=64, ndf=64, **extra)
    if variant == "video":
        kw.setdefault("predLen", 2)
    orc = ostep.StepOracle(onets.default_opt(variant, **kw), seed=1234, dtype=np.float64)
    trn = train.FusedTrainer(models.default_opt(variant, **kw), precision="bf16")
    assert trn.param_count(0) == orc.pG.size and trn.param_count(1) == orc.pD.size
    trn.set_params(0, orc.pG)
    trn.set_params(1, orc.pD)
    return orc, trn


def _g_modules(orc):
    """Oracle modules of netG in execution order, flattened."""
    mods = list(orc.netG.modules[0].modules) + list(orc.netG.modules[1:])
    return mods


def _blocks(mods):
    """Group a flattened module list into (conv, bn or None, activation) blocks."""
    out, i = [], 0
    while i < len(mods):
        m = mods[i]
        if "Convolution" in type(m).__name__:
            bn = mods[i + 1] if i + 1 < len(mods) and "BatchNorm" in type(mods[i + 1]).__name__ else None
            act = mods[i + (2 if bn else 1)]
            out.append((m, bn, act))
        i += 1
    return out


def _flat(net):
    out = []
    for m in net.modules:
        out += _flat(m) if hasattr(m, "modules") else [m]
    return out


def _gate(act, a):
    n = type(act).__name__
    if n == "LeakyReLU":
        return np.where(a > 0, 1.0, 0.2)
    if n == "ReLU":
        return np.where(a > 0, 1.0, 0.0)
    if n == "Tanh":
        return 1.0 - a * a
    raise AssertionError(n)


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.dot(a, b) / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


@pytest.fixture
def fast_oracle():
    """Big shapes (fineSize 256): the oracle's heavy ops on the PyTorch-CPU engine (pinned to the numpy functions by tests/test_oracle_torch_engine.py)."""
    from oracle import torch_engine
    torch_engine.enable()
    yield
    torch_engine.disable()


# the last case is BASELINE.json configs[3]'s shape family: train_deepernet at 256 x 256 (5x5 bottleneck, 8x8 first decoder map, patch head)
@pytest.mark.parametrize("variant,extra", [("image", {}), ("video", {}), ("video", {"wtgdl": 0.5}), ("video", {"weight_nomask": 0.0}),
                                           ("video", {"fineSize": 256, "B": 2, "predLen": 1, "wtgdl": 0.5})])
def test_fused_step_matches_oracle(cenn, fast_oracle, variant, extra):
    """Losses against the fp64 oracle + SELF-CONSISTENCY of every kernel of the generator's forward and backward:
    each block's conv output, activation, weight gradient, and the (dgrad -> BN/activation backward) chain into the
    previous block are recomputed in fp64 with the oracle's formulas FROM THE EXECUTOR'S OWN STORED TENSORS.
    (A direct whole-network comparison is meaningless beyond a few layers: bf16 storage flips ~0.3 % of the
    ReLU/LeakyReLU gates, which makes per-element gradients chaotic -- DESIGN.md "parity".)"""
    from oracle import ops
    orc, trn = _pair(variant, **extra)
    rng = np.random.default_rng(4321)
    batch = orc.synth_batch(rng)
    pG0 = orc.pG.copy()
    lo = orc.step(*batch)
    lg = trn.step_host(*batch)
    for k in ("errD_real", "errG_l2", "errG_total"):          # one network deep
        assert lg[k] == pytest.approx(lo[k], rel=2e-2), k
    for k in ("errD_fake", "errD", "errG"):                   # through G and D (19 bf16 layers)
        assert lg[k] == pytest.approx(lo[k], rel=5e-2), k
    if extra.get("wtgdl"):
        assert lg["errG_gdl"] == pytest.approx(lo["errG_gdl"], rel=2e-2)
    # ---- direct comparisons that are well conditioned
    mods = _flat(orc.netG)
    blocks = _blocks(mods)
    e1 = blocks[0][2].output
    assert rel_err(trn.fetch("G.0.a").reshape(e1.shape), e1) <= 2e-2           # first block: per-layer BF16 bound
    assert rel_err(trn.fetch("fake").reshape(orc.netG.output.shape), orc.netG.output) <= 4e-2   # 12 layers
    gG, gD = trn.get_grads(0), trn.get_grads(1)
    assert _cos(gG, orc.gG) >= 0.9 and _cos(gD, orc.gD) >= 0.9
    assert np.linalg.norm(gG) == pytest.approx(np.linalg.norm(orc.gG), rel=5e-2)
    assert np.linalg.norm(gD) == pytest.approx(np.linalg.norm(orc.gD), rel=5e-2)
    # ---- self-consistency of every generator block (weights used by the step = the initial ones, bf16-rounded)
    offs, off = {}, 0
    for m in mods:
        if getattr(m, "weight", None) is not None:
            offs[id(m)] = off
            off += m.weight.size + m.bias.size
    q = ops.bf16_round
    x_in = q(batch[0].astype(np.float64))
    for bi, (conv, bn, act) in enumerate(blocks):
        full = type(conv).__name__ == "SpatialFullConvolution"
        o = offs[id(conv)]
        w = q(pG0[o:o + conv.weight.size].reshape(conv.weight.shape))
        xin = x_in if bi == 0 else trn.fetch("G.%d.in" % bi).reshape(blocks[bi - 1][2].output.shape).astype(np.float64)
        a = trn.fetch("G.%d.a" % bi).reshape(act.output.shape).astype(np.float64)
        g_y = trn.fetch("G.%d.g" % bi).reshape(conv.output.shape).astype(np.float64)
        geo = (conv.dH, conv.dW, conv.padH, conv.padW)
        # forward: conv (+BN batch statistics) + activation
        y_ref = ops.fullconv_forward(xin, w, None, *geo) if full else ops.conv_forward(xin, w, None, *geo)
        if bn is not None:
            y = trn.fetch("G.%d.y" % bi).reshape(conv.output.shape).astype(np.float64)
            assert rel_err(y, y_ref) <= 5e-3, ("conv output", bi)             # <= 1 bf16 ulp of the largest element
            ob = offs[id(bn)]
            gamma, beta = pG0[ob:ob + bn.weight.size], pG0[ob + bn.weight.size:ob + 2 * bn.weight.size]
            z, mean, invstd = ops.bn_forward(y, gamma, beta, np.zeros_like(gamma), np.ones_like(gamma), True)
        else:
            y, z = None, y_ref
        a_ref = {"LeakyReLU": lambda v: ops.leaky_relu(v, 0.2), "ReLU": ops.relu, "Tanh": np.tanh}[type(act).__name__](z)
        # (2.5 bf16 ulps: the executor takes the batch statistics from the fp32 accumulators, this check from the
        # stored bf16 conv output; the 8-sample bottleneck BN amplifies that difference)
        assert rel_err(a, a_ref) <= 1e-2, ("activation", bi)
        # weight gradient from (input, g_y)
        gw, gb = np.zeros(conv.weight.shape), np.zeros(conv.bias.shape)
        (ops.fullconv_acc_grad if full else ops.conv_acc_grad)(xin, g_y, gw, gb, *geo)
        assert rel_err(gG[o:o + gw.size], gw) <= 1e-4, ("wgrad", bi)
        if bn is None:                                                          # conv bias gradient (non-zero only without BN)
            assert rel_err(gG[o + gw.size:o + gw.size + gb.size], gb) <= 3e-3, ("bias grad", bi)   # the kernel sums fp32 gradients, this check their bf16-rounded copies
        # dgrad into the previous block, then that block's BN / activation backward
        if bi > 0:
            pconv, pbn, pact = blocks[bi - 1]
            g_a = q(ops.fullconv_grad_input(g_y, w, *geo) if full else ops.conv_grad_input(xin.shape, g_y, w, *geo))
            pa = xin                                                            # the previous block's activation output
            dz = g_a * _gate(pact, pa)
            if pbn is not None:
                py = trn.fetch("G.%d.y" % (bi - 1)).reshape(pconv.output.shape).astype(np.float64)
                ob = offs[id(pbn)]
                pgamma = pG0[ob:ob + pbn.weight.size]
                _, pm, pis = ops.bn_forward(py, pgamma, np.zeros_like(pgamma), np.zeros_like(pgamma), np.ones_like(pgamma), True)
                gg, gbt = np.zeros_like(pgamma), np.zeros_like(pgamma)
                exp = ops.bn_backward(py, dz, pgamma, pm, pis, None, None, True, ggamma=gg, gbeta=gbt)
                assert rel_err(gG[ob:ob + gg.size], gg) <= 1e-2, ("BN gamma grad", bi - 1)
                assert rel_err(gG[ob + gg.size:ob + 2 * gg.size], gbt) <= 1e-2, ("BN beta grad", bi - 1)
            else:
                exp = dz
            got = trn.fetch("G.%d.g" % (bi - 1)).reshape(pconv.output.shape).astype(np.float64)
            assert rel_err(got, exp) <= 1.5e-2 and _cos(got, exp) >= 0.9999, ("dgrad + BN/activation backward", bi - 1)
    # Adam moved the parameters in the oracle's direction; BN running statistics (momentum 0.1; D twice, G once)
    dG, dG_ref = trn.get_params(0) - pG0, orc.pG - pG0
    big = np.abs(orc.gG) > 0.1 * np.abs(orc.gG).max()
    assert float(np.mean(np.sign(dG[big]) == np.sign(dG_ref[big]))) >= 0.99
    ref = np.concatenate([np.concatenate([m.running_mean, m.running_var]) for m in mods if hasattr(m, "running_mean")])
    assert rel_err(trn.get_bn_stats(0), ref) <= 2e-2
    ref = np.concatenate([np.concatenate([m.running_mean, m.running_var]) for m in orc.netD.modules if hasattr(m, "running_mean")])
    assert rel_err(trn.get_bn_stats(1), ref) <= 3e-2


def test_fused_losses_track_oracle_over_steps(cenn):
    """north_star: losses within 1 % after many steps (here 40 steps at a reduced batch)."""
    orc, trn = _pair("image", B=8, nB=128)
    rng = np.random.default_rng(77)
    hist_o, hist_g = [], []
    for it in range(40):
        batch = orc.synth_batch(rng)
        lo, lg = orc.step(*batch), trn.step_host(*batch)
        hist_o.append([lo["errD"], lo["errG"], lo["errG_l2"]])
        hist_g.append([lg["errD"], lg["errG"], lg["errG_l2"]])
    ho, hg = np.array(hist_o), np.array(hist_g)
    assert np.all(np.isfinite(hg))
    # the L2 term (what the generator is actually trained on at wtl2 = 0.999): every step within 5 %, and the
    # mean over the last 10 steps within 2 % (trajectories separate through Adam's sign-like updates; the
    # 100-step / 1 % north-star figure is checked at batch 64 by test_losses_after_100_steps_match_oracle_fixture)
    assert np.max(np.abs(hg[:, 2] - ho[:, 2]) / ho[:, 2]) <= 5e-2
    assert abs(hg[-10:, 2].mean() - ho[-10:, 2].mean()) <= 2e-2 * ho[-10:, 2].mean()
    # the adversarial losses are chaotic in the GAN game at batch 8 (who is "winning" flips on tiny perturbations):
    # first step within 5 %, finite afterwards
    assert np.max(np.abs(hg[0, :2] - ho[0, :2]) / ho[0, :2]) <= 5e-2


def test_generator_forward_eval_matches_oracle(cenn):
    orc, trn = _pair("video", B=4, nB=128)
    rng = np.random.default_rng(5)
    for _ in range(2):                       # make the running statistics non-trivial
        batch = orc.synth_batch(rng)
        orc.step(*batch)
        trn.step_host(*batch)
    # evaluate with the oracle's weights and running stats on both sides
    trn.set_params(0, orc.pG)
    ref_stats = np.concatenate([np.concatenate([m.running_mean, m.running_var]) for m in _g_modules(orc) if hasattr(m, "running_mean")])
    trn.set_bn_stats(0, ref_stats)
    orc.netG.evaluate()
    x = rng.uniform(-1, 1, (3, 6, 128, 128))
    y_ref = orc.netG.forward(x)
    orc.netG.training()
    y = trn.generator_forward(x.astype(np.float32))
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) <= 2e-2


def test_fused_rejects_bad_config(cenn):
    from video_filler_b200 import models, train
    from video_filler_b200._lib import CennError
    with pytest.raises(CennError, match="fineSize must be 128"):
        train.FusedTrainer(models.default_opt("image", fineSize=64, batchSize=2))
    with pytest.raises(CennError, match="BF16 tensor-core mode only"):
        train.FusedTrainer(models.default_opt("image", batchSize=2), precision="fp32")


def test_pipelined_host_steps_match_blocking_steps(cenn):
    """cenn_trainer_step_host_async / wait_losses (copy of step k+1 overlapped with step k, losses read one call late)
    against cenn_trainer_step_host on the same inputs."""
    from video_filler_b200 import models, synth, train, util
    kw = dict(batchSize=8, nBottleneck=128, nef=64, ngf=64, ndf=64)
    opt = models.default_opt("image", **kw)
    rng = np.random.default_rng(5)
    pG = util.params_flat(util.weights_init(util.describe_netG(opt), rng))
    pD = util.params_flat(util.weights_init(util.describe_netD(opt), rng))
    blocking, piped = train.FusedTrainer(opt), train.FusedTrainer(opt)
    for t in (blocking, piped):
        t.set_params(0, pG); t.set_params(1, pD)
    batches = [synth.image_batch(8, 128, 4, np.random.default_rng(100 + i)) for i in range(4)]
    ref = [blocking.step_host(*b) for b in batches]
    got = []
    for i, b in enumerate(batches):
        piped.step_host_async(*b)
        if i > 0:
            got.append(piped.wait_losses())
    got.append(piped.wait_losses())
    assert len(got) == len(ref)
    for k in ("errD", "errG", "errG_l2"):                          # step 1: same arithmetic, up to the executor's atomics order
        assert got[0][k] == pytest.approx(ref[0][k], rel=1e-2), k
    for g, r in zip(got, ref):
        assert all(np.isfinite(v) for v in g.values())
        assert g["errG_l2"] == pytest.approx(r["errG_l2"], rel=5e-2)
    with pytest.raises(Exception, match="no step in flight"):
        piped.wait_losses()


@pytest.mark.parametrize("variant,B", [("image", 3), ("video", 5)])
def test_fused_step_ragged_batch(cenn, variant, B):
    """Batch sizes that do not fill the 128-pixel tiles of the small layers (4x4 and 8x8 grids pack several samples per tile):
    out-of-range rows are clipped by TMA and masked out of the BN statistics."""
    orc, trn = _pair(variant, B=B, nB=128)
    batch = orc.synth_batch(np.random.default_rng(9))
    lo, lg = orc.step(*batch), trn.step_host(*batch)
    for k in ("errD_real", "errG_l2", "errG_total"):
        assert lg[k] == pytest.approx(lo[k], rel=2e-2), k
    for k in ("errD_fake", "errD", "errG"):
        assert lg[k] == pytest.approx(lo[k], rel=8e-2), k          # tiny batches: BN over 3-5 samples at the bottleneck
    gG = trn.get_grads(0)
    assert np.all(np.isfinite(gG)) and _cos(gG, orc.gG) >= 0.9


def test_losses_after_100_steps_match_oracle_fixture(cenn):
    """north_star: losses after 100 steps within 1 %.  The oracle's 100 fp32 steps at the CPU config (batch 64, nBottleneck
    4000) are a committed fixture (tests/tools/parity_steps.py --make-golden); the executor replays the same seeded batches."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "tools"))
    import parity_steps
    if not os.path.exists(parity_steps.GOLDEN):
        pytest.skip("tests/golden/losses_100.npz not generated")
    ours, gold = parity_steps.run_executor()
    s = parity_steps.summarize(ours, gold)
    print(s)
    assert len(ours) == 100 and np.all(np.isfinite(ours))
    # the L2 term is what the generator is trained on at wtl2 = 0.999 (errG_total = 0.001 errG + 0.999 errG_l2 up to the edge weighting)
    assert s["errG_l2"]["rel_at_last_step"] <= 1e-2 and s["errG_l2"]["rel_of_mean_last10"] <= 1e-2
    # errG_total = 0.001 * errG + 0.999 * errG_l2: the chaotic adversarial term alone moves it by ~0.6 % (measured 0.81 % in total)
    assert s["errG_total"]["rel_at_last_step"] <= 3e-2 and s["errG_total"]["rel_of_mean_last10"] <= 1e-2
    assert s["errG_l2"]["max_rel_all_steps"] <= 5e-2          # measured 3.2 % at the worst step (profiles/r1_parity_steps.json)
    # adversarial terms: the first step is a pure function of the inputs; later the GAN game amplifies rounding differences.
    # CONTROL (tests/tools/parity_steps.py --control -> tests/golden/parity_control.json): three CPU runs of the same oracle step sequence
    # that differ from the golden run only at rounding level (other summation order, weights perturbed by one fp32 ulp, fp64)
    # drift from it by 13-44 % (errD) and 10-33 % (errG) at step 100 and by 2-33 % in the mean over the last ten steps.  A single step of
    # these terms is a coin flip (errD swings between 0.1 and 6 from step to step on every arm: two builds of the executor measured
    # 20 % and 93 % at step 100), so the executor is held to the control envelope on the ten-step mean (x4), and to the same
    # order of magnitude as the oracle on the medians of the second half of the run.
    assert abs(ours[0, 0] - gold[0, 0]) <= 1e-2 * gold[0, 0] and abs(ours[0, 1] - gold[0, 1]) <= 1e-2 * gold[0, 1]
    import json
    ctl = json.load(open(os.path.join(root, "tests", "golden", "parity_control.json")))["arms"]
    for j, k in enumerate(("errD", "errG")):
        envelope = max(a[k]["rel_of_mean_last10"] for a in ctl.values())
        # (x4: the control arms differ from the golden run at fp32 rounding level, the executor computes in bf16 -- a larger perturbation of the
        # same chaotic game; builds of this round measured 0.10-0.41 for errG against a control maximum of 0.16)
        assert s[k]["rel_of_mean_last10"] <= 4.0 * envelope, (k, s[k]["rel_of_mean_last10"], envelope)
        mo, mg = float(np.median(ours[50:, j])), float(np.median(gold[50:, j]))
        assert mg / 2.0 <= mo <= 2.0 * mg, (k, mo, mg)      # same order of magnitude (control arms: within 18 % of the golden medians; executor r1: 16 %)
    # the reconstruction loss of the control arms: 0.03-0.14 % at step 100, <= 2.1 % at the worst step
    assert max(a["errG_l2"]["rel_at_last_step"] for a in ctl.values()) <= 1e-2


def test_clip_mode_step_equals_three_tensor_step(cenn):
    """cenn_trainer_step_clips_host (device-side maskedFill / expand / hflip / [0,1]->[-1,1], datavid/donkey_folder.lua:161-187)
    against cenn_trainer_step_host fed with the same clips prepared on the host the way the loader does."""
    from video_filler_b200 import models, train
    kw = dict(batchSize=6, nBottleneck=128, nef=64, ngf=64, ndf=64, predLen=2, wtgdl=0.5)
    opt = models.default_opt("video", **kw)
    orc = ostep.StepOracle(onets.default_opt("video", **kw), seed=5, dtype=np.float64)
    rng = np.random.default_rng(9)
    frames = rng.uniform(0, 1, (6, 6, 128, 128)).astype(np.float32)
    mask1 = np.zeros((6, 128, 128), np.uint8)
    for b in range(6):
        y, x = rng.integers(3, 80, 2)
        mask1[b, y:y + 30, x:x + 40] = 1
    flip = np.array([0, 1, 0, 1, 1, 0], np.uint8)
    # the loader's hook on the host: maskedFill in [0,1], hflip of all three, then *2-1
    mv = opt["maskValue"]
    maskx = np.repeat(mask1[:, None], 6, axis=1)
    masked01 = np.where(maskx != 0, np.float32(mv), frames)
    f = flip.astype(bool)
    full, masked, maskh = frames.copy(), masked01.copy(), maskx.copy()
    full[f], masked[f], maskh[f] = full[f][..., ::-1], masked[f][..., ::-1], maskh[f][..., ::-1]
    full, masked = full * 2 - 1, masked * 2 - 1
    res = []
    for mode in ("host", "clips", "clips_noflip"):
        trn = train.FusedTrainer(opt, precision="bf16")
        trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
        if mode == "host":
            losses = trn.step_host(masked, full, maskh)
        elif mode == "clips":
            losses = trn.step_clips_host(frames, mask1, flip)
        else:
            losses = trn.step_clips_host(frames, mask1, None)
        res.append((losses, trn.fetch("ctx"), trn.get_grads(1)))
        if mode == "clips":
            losses2 = trn.step_clips_host(frames, mask1, flip)          # second call: captured graph, updated weights
            assert np.isfinite(list(losses2.values())).all() and losses2["errG_l2"] != losses["errG_l2"]
        trn.close()
    (l0, ctx0, g0), (l1, ctx1, g1), (l2, ctx2, _) = res
    assert np.array_equal(ctx0, ctx1)                                    # identical bf16 inputs on both paths
    assert not np.array_equal(ctx1, ctx2)                                # the flags do something
    # identical inputs: what is left is the run-to-run spread of the executor (order of fp32 atomics); errG is evaluated after
    # D's Adam update and inherits the spread of D's gradients
    # (errD: a handful of samples per BN batch -- one flipped bf16 rounding in a statistic moves the discriminator's loss by up to ~1 %;
    # observed spread between two runs on identical inputs: 0.1-0.8 %)
    for k, tol in (("errG_l2", 2e-3), ("errG_gdl", 2e-3), ("errD", 1.5e-2), ("errG", 4e-2)):
        assert abs(l0[k] - l1[k]) <= tol * max(abs(l0[k]), 1e-3), (k, l0[k], l1[k])
    assert _cos(g0, g1) >= 0.97
    with pytest.raises(Exception, match="video variant"):
        img = train.FusedTrainer(models.default_opt("image", batchSize=2, nBottleneck=128), precision="bf16")
        img.step_clips_host(frames[:2, :3], mask1[:2], None, maskValue=0.4)


def test_fused_wide_net_matches_oracle(cenn):
    """The wide nets of train_wholeim_input.lua:36-48 (nef = ngf = 192, ndf = 128, predLen 1, weight_nomask 1; bottleneck reduced
    from 6400 to 640 to keep the fp64 oracle affordable): channel counts that are multiples of 64 but not powers of two
    (192 / 384 / 768 / 1536) through every tile shape of the executor."""
    from video_filler_b200 import models, train
    kw = dict(batchSize=4, nBottleneck=640, nef=192, ngf=192, ndf=128, predLen=1, weight_nomask=1.0)
    orc = ostep.StepOracle(onets.default_opt("video", **kw), seed=77, dtype=np.float64)
    trn = train.FusedTrainer(models.default_opt("video", **kw), precision="bf16")
    assert trn.param_count(0) == orc.pG.size and trn.param_count(1) == orc.pD.size
    trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
    batch = orc.synth_batch(np.random.default_rng(5))
    lo, lg = orc.step(*batch), trn.step_host(*batch)
    for k in ("errD_real", "errG_l2", "errG_total"):
        assert lg[k] == pytest.approx(lo[k], rel=2e-2), k
    for k in ("errD_fake", "errD", "errG"):
        assert lg[k] == pytest.approx(lo[k], rel=5e-2), k
    # 13 bf16 layers with contractions up to K = 16 * 1536: measured 4.2e-2 of the output range (the 12-layer image net stays under 4e-2)
    assert rel_err(trn.fetch("fake").reshape(orc.netG.output.shape), orc.netG.output) <= 6e-2
    gG, gD = trn.get_grads(0), trn.get_grads(1)
    assert _cos(gG, orc.gG) >= 0.9 and _cos(gD, orc.gD) >= 0.9
    assert np.linalg.norm(gG) == pytest.approx(np.linalg.norm(orc.gG), rel=5e-2)
    trn.close()


# ---------------------------------------------------------------------------------------------------------------------
# netD: block-level self-consistency of all three discriminator sweeps (real, fake, dgrad-only)
# ---------------------------------------------------------------------------------------------------------------------
def _param_offsets(mods):
    offs, off = {}, 0
    for m in mods:
        if getattr(m, "weight", None) is not None:
            offs[id(m)] = off
            off += m.weight.size + m.bias.size
    return offs


def _dev_batch(T, batch):
    import ctypes as C
    a, b = T.CudaTensor.from_numpy(np.ascontiguousarray(batch[0], np.float32)), T.CudaTensor.from_numpy(np.ascontiguousarray(batch[1], np.float32))
    m = None
    if len(batch) > 2:
        mh = np.ascontiguousarray(batch[2]).astype(np.uint8)
        mp = C.c_void_p()
        T.api().cenn_malloc(T.state(), mh.nbytes, C.byref(mp))
        T.api().cenn_copy_h2d(T.state(), mp, mh.ctypes.data_as(C.c_void_p), mh.nbytes)
        m = mp.value
    return a, b, m


def _bce_gpre(sig, label, n):
    """BCECriterion backward followed by Sigmoid backward, per sample."""
    e = 1e-12
    gx = -(label - sig) / ((1 - sig + e) * (sig + e)) / n
    return gx * sig * (1 - sig)


def _check_d_sweep(trn, blocks, offs, p_fwd, p_bwd, grads, label, B, check_fwd_weights=True, what=""):
    """One discriminator sweep recomputed block by block in fp64 from the executor's OWN stored tensors.
    p_fwd: flat parameters the forward used; p_bwd: the ones the backward (dgrad, BN gamma) used; grads: flat gradient vector that
    holds exactly this sweep's parameter gradients (None for the dgrad-only sweep of fGx)."""
    from oracle import ops
    q = ops.bf16_round
    nblk = len(blocks)
    shapes = [blk[0].output.shape for blk in blocks]
    g_next = None                     # gradient w.r.t. the conv output of block bi + 1 (from stored tensors)
    for bi in range(nblk - 1, -1, -1):
        conv, bn, act = blocks[bi]
        head = bi == nblk - 1
        o = offs[id(conv)]
        wf = q(p_fwd[o:o + conv.weight.size].reshape(conv.weight.shape).astype(np.float64))
        wb = q(p_bwd[o:o + conv.weight.size].reshape(conv.weight.shape).astype(np.float64))
        in_shape = (B,) + ((blocks[bi - 1][0].output.shape[1:]) if bi > 0 else tuple(trn_in_shape(trn)))
        xin = trn.fetch("D.%d.in" % bi).reshape(in_shape).astype(np.float64)
        geo = (conv.dH, conv.dW, conv.padH, conv.padW)
        y_ref = ops.conv_forward(xin, wf, None, *geo)
        if head:
            sig = trn.fetch("D.%d.sig" % bi).astype(np.float64)
            sig_ref = 1.0 / (1.0 + np.exp(-y_ref.reshape(-1)))
            assert np.max(np.abs(sig - sig_ref)) <= 2e-3, (what, "head sigmoid")
            g_y = _bce_gpre(sig, label, sig.size).reshape(y_ref.shape)          # one BCE term per patch output (cfg4 convention: 25 per sample at 256 x 256)
        else:
            a = trn.fetch("D.%d.a" % bi).reshape(shapes[bi]).astype(np.float64)
            if bn is not None:
                y = trn.fetch("D.%d.y" % bi).reshape(shapes[bi]).astype(np.float64)
                assert rel_err(y, y_ref) <= 5e-3, (what, "conv output", bi)
                ob = offs[id(bn)]
                gamma, beta = p_fwd[ob:ob + bn.weight.size].astype(np.float64), p_fwd[ob + bn.weight.size:ob + 2 * bn.weight.size].astype(np.float64)
                z, _, _ = ops.bn_forward(y, gamma, beta, np.zeros_like(gamma), np.ones_like(gamma), True)
            else:
                z = y_ref
            assert rel_err(a, ops.leaky_relu(z, 0.2)) <= 1e-2, (what, "activation", bi)
            g_y = trn.fetch("D.%d.g" % bi).reshape(shapes[bi]).astype(np.float64)
        if grads is not None:
            gw, gb = np.zeros(conv.weight.shape), np.zeros(conv.bias.shape)
            ops.conv_acc_grad(xin, g_y, gw, gb, *geo)
            assert rel_err(grads[o:o + gw.size], gw) <= (2e-3 if head else 1e-4), (what, "wgrad", bi)
            if bn is None:
                assert rel_err(grads[o + gw.size:o + gw.size + gb.size], gb) <= 3e-3, (what, "bias grad", bi)
        # dgrad into the previous block (or into D's input: df_dg), then that block's BN / LeakyReLU backward
        g_a = q(ops.conv_grad_input(xin.shape, g_y, wb, *geo))
        if bi == 0:
            if grads is None:          # fGx: df_dg = netD:updateGradInput (train.lua:373)
                got = trn.fetch("df_dg").reshape(xin.shape).astype(np.float64)
                assert rel_err(got, g_a) <= 1.5e-2 and _cos(got, g_a) >= 0.9999, (what, "df_dg")
            continue
        pconv, pbn, pact = blocks[bi - 1]
        dz = g_a * np.where(xin > 0, 1.0, 0.2)
        if pbn is not None:
            py = trn.fetch("D.%d.y" % (bi - 1)).reshape(shapes[bi - 1]).astype(np.float64)
            ob = offs[id(pbn)]
            pgamma = p_bwd[ob:ob + pbn.weight.size].astype(np.float64)
            _, pm, pis = ops.bn_forward(py, pgamma, np.zeros_like(pgamma), np.zeros_like(pgamma), np.ones_like(pgamma), True)
            gg, gbt = np.zeros_like(pgamma), np.zeros_like(pgamma)
            exp = ops.bn_backward(py, dz, pgamma, pm, pis, None, None, True, ggamma=gg, gbeta=gbt)
            if grads is not None:
                assert rel_err(grads[ob:ob + gg.size], gg) <= 1e-2, (what, "BN gamma grad", bi - 1)
                assert rel_err(grads[ob + gg.size:ob + 2 * gg.size], gbt) <= 1e-2, (what, "BN beta grad", bi - 1)
        else:
            exp = dz
        got = trn.fetch("D.%d.g" % (bi - 1)).reshape(shapes[bi - 1]).astype(np.float64)
        assert rel_err(got, exp) <= 1.5e-2 and _cos(got, exp) >= 0.9999, (what, "dgrad + BN/activation backward", bi - 1)


def trn_in_shape(trn):
    o = trn.opt
    if o["variant"] == "image":
        return (o["nc"], o["fineSize"] // 2, o["fineSize"] // 2)
    return (o["nc"] * o["predLen"], o["fineSize"], o["fineSize"])


@pytest.mark.parametrize("variant,extra", [("image", {}), ("video", {}), ("video", {"fineSize": 256, "B": 2, "predLen": 1})])
def test_fused_discriminator_blocks_self_consistent(cenn, fast_oracle, variant, extra):
    """VERDICT r1 (weak 4): the discriminator's kernels were only checked through the losses.  Here every D block of all three
    sweeps is recomputed in fp64 from the executor's own stored tensors:
      * REAL sweep (fDx, train.lua:303-310): the step program is stopped after the sweep (cenn_trainer_step_until) -- forward,
        head + BCE, head wgrad / dgrad, every wgrad, bias / BN affine gradients, dgrad -> BN / LeakyReLU backward chains;
      * FAKE sweep forward (train.lua:337) and the DGRAD-ONLY sweep of fGx (train.lua:373: updated D weights, fake-pass
        activations and BN statistics, label 1) down to df_dg, after a complete step."""
    import video_filler_b200.tensor as T
    orc, trn = _pair(variant, **extra)
    B = orc.opt["batchSize"]
    batch = orc.synth_batch(np.random.default_rng(31))
    pD0 = orc.pD.copy()
    orc.step(*batch)                                          # only to have module shapes / outputs at hand
    mods = _flat(orc.netD)
    blocks = _blocks(mods)
    offs = _param_offsets(mods)
    da, db, dm = _dev_batch(T, batch)
    trn.step_until(da.ptr, db.ptr, dm, "fold_gbias", 0)      # end of the real sweep (gradParametersD holds the real pass only)
    real_in = trn.fetch("D.0.in").reshape((B,) + trn_in_shape(trn))
    from oracle import ops
    assert np.array_equal(real_in, ops.bf16_round(np.asarray(batch[1], np.float32)))
    _check_d_sweep(trn, blocks, offs, pD0, pD0, trn.get_grads(1), 1.0, B, what="real sweep")
    trn.step_device(da.ptr, db.ptr, dm)                       # a complete step from the same (not yet updated) weights
    T.api().cenn_synchronize(T.state())
    pD1 = trn.get_params(1)
    assert not np.array_equal(pD1, pD0.astype(np.float32))
    _check_d_sweep(trn, blocks, offs, pD0, pD1, None, 1.0, B, what="fake forward + dgrad-only sweep")
    trn.close()


@pytest.mark.parametrize("variant,kw", [("image", dict(batchSize=256, nBottleneck=4000)),
                                        ("video", dict(batchSize=64, nBottleneck=4000, predLen=4, wtgdl=0.5))])
def test_fused_step_at_benchmark_shapes(cenn, variant, kw):
    """One step at the SHAPES bench.py times -- BASELINE.json configs[1] (batch 256, nBottleneck 4000) and configs[2] per GPU (64 clips
    of 12 stacked channels, nBottleneck 4000, GDL) -- against the fp32 oracle (its heavy ops on the PyTorch-CPU engine, which
    tests/test_oracle_torch_engine.py pins to the numpy restatement)."""
    from oracle import torch_engine
    from video_filler_b200 import models, train
    torch_engine.enable()
    try:
        full = dict(nef=64, ngf=64, ndf=64, **kw)
        orc = ostep.StepOracle(onets.default_opt(variant, **full), seed=1234, dtype=np.float32)
        trn = train.FusedTrainer(models.default_opt(variant, **full), precision="bf16")
        assert trn.param_count(0) == orc.pG.size and trn.param_count(1) == orc.pD.size
        trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
        batch = orc.synth_batch(np.random.default_rng(2024))
        lo, lg = orc.step(*batch), trn.step_host(*batch)
    finally:
        torch_engine.disable()
    print({k: (lg[k], lo[k]) for k in lo if lo[k] is not None})
    for k in ("errD_real", "errG_l2", "errG_total"):
        assert lg[k] == pytest.approx(lo[k], rel=2e-2), k
    for k in ("errD_fake", "errD", "errG"):
        assert lg[k] == pytest.approx(lo[k], rel=5e-2), k
    if kw.get("wtgdl"):
        assert lg["errG_gdl"] == pytest.approx(lo["errG_gdl"], rel=2e-2)
    e1 = _blocks(_flat(orc.netG))[0][2].output
    assert rel_err(trn.fetch("G.0.a").reshape(e1.shape), e1) <= 2e-2
    assert rel_err(trn.fetch("fake").reshape(orc.netG.output.shape), orc.netG.output) <= 5e-2
    gG, gD = trn.get_grads(0), trn.get_grads(1)
    assert np.all(np.isfinite(gG)) and np.all(np.isfinite(gD))
    assert _cos(gG, orc.gG) >= 0.95 and _cos(gD, orc.gD) >= 0.95
    assert np.linalg.norm(gG) == pytest.approx(np.linalg.norm(orc.gG), rel=5e-2)
    assert np.linalg.norm(gD) == pytest.approx(np.linalg.norm(orc.gD), rel=5e-2)
    trn.close()


def test_frame_mode_step_runs_the_loader_hook_on_the_device(cenn):
    """cenn_trainer_step_frames_host (device-side random crop, mask crop / random-block mask, maskedFill, hflip, rescale:
    datavid/donkey_folder.lua:114-129,138-187) against oracle/loader.py on identical draws: the step inputs the executor derives
    are bit-identical (bf16) to the hook's tensors fed through cenn_trainer_step_host, and the step that follows is the same step."""
    from oracle import loader
    from video_filler_b200 import models, train
    B, F, iH, iW = 6, 128, 180, 240
    kw = dict(batchSize=B, nBottleneck=128, nef=64, ngf=64, ndf=64, predLen=2, wtgdl=0.5)
    opt = models.default_opt("video", **kw)
    orc = ostep.StepOracle(onets.default_opt("video", **kw), seed=5, dtype=np.float64)
    rng = np.random.default_rng(17)
    frames = rng.integers(0, 256, (B, 6, iH, iW)).astype(np.uint8)
    mask_full = np.zeros((iH, iW), np.uint8)
    mask_full[20:70, 150:230] = 255                      # a logo in the upper right corner: some crops miss it -> random blocks
    crop, flip, blocks = loader.draw_hook_params(B, iH, iW, F, rng)
    crop[0] = (0, 0); crop[1] = (iH - F, iW - F)          # the extreme origins; sample 0 misses the logo, sample 1 may hit it
    crop[2] = (10, 100)                                   # certainly overlaps the logo
    masked, full, mask = loader.train_hook(frames, mask_full, crop, flip, blocks, F, opt["maskValue"])
    assert mask[0].max() == 1 and mask[2].max() == 1 and not np.array_equal(mask[0], mask[2])
    res = []
    for mode in ("host", "frames"):
        trn = train.FusedTrainer(opt, precision="bf16")
        trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
        losses = trn.step_host(masked, full, mask) if mode == "host" else trn.step_frames_host(frames, mask_full, crop, flip, blocks)
        res.append((losses, trn.fetch("ctx"), trn.get_grads(1)))
        trn.close()
    (l0, ctx0, g0), (l1, ctx1, g1) = res
    assert np.array_equal(ctx0, ctx1)                     # identical bf16 step inputs on both paths
    for k, tol in (("errG_l2", 2e-3), ("errG_gdl", 2e-3), ("errD", 1.5e-2), ("errG", 4e-2)):
        assert abs(l0[k] - l1[k]) <= tol * max(abs(l0[k]), 1e-3), (k, l0[k], l1[k])
    assert _cos(g0, g1) >= 0.97


def test_fused_wide_net_at_the_real_bottleneck(cenn, fast_oracle):
    """train_wholeim_input.lua:40-43 as shipped: nef = ngf = 192, ndf = 128, nBottleneck 6400 (VERDICT r1 missing 3: the wide net had only
    been run at nBottleneck 640).  fp32 oracle (365 M generator parameters), four samples."""
    from video_filler_b200 import models, train
    kw = dict(batchSize=4, nBottleneck=6400, nef=192, ngf=192, ndf=128, predLen=1, weight_nomask=1.0)
    orc = ostep.StepOracle(onets.default_opt("video", **kw), seed=77, dtype=np.float32)
    trn = train.FusedTrainer(models.default_opt("video", **kw), precision="bf16")
    assert trn.param_count(0) == orc.pG.size and trn.param_count(1) == orc.pD.size
    trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
    batch = orc.synth_batch(np.random.default_rng(5))
    lo, lg = orc.step(*batch), trn.step_host(*batch)
    for k in ("errD_real", "errG_l2", "errG_total"):
        assert lg[k] == pytest.approx(lo[k], rel=2e-2), k
    for k in ("errD_fake", "errD", "errG"):
        assert lg[k] == pytest.approx(lo[k], rel=1e-1), k          # BatchNorm over four values per channel at the 1x1 bottleneck: a flipped bf16 rounding moves a normalised value by O(1)
    gG = trn.get_grads(0)
    assert np.all(np.isfinite(gG)) and _cos(gG, orc.gG) >= 0.9
    trn.close()


def test_byte_image_step_equals_float_step(cenn):
    """cenn_trainer_step_images_u8_host (rescale, centre clone, mean fill on the device: train.lua:286-290) against cenn_trainer_step_host
    fed with the same images prepared on the host the way the script does."""
    from video_filler_b200 import models, train, util
    kw = dict(batchSize=6, nBottleneck=128, nef=64, ngf=64, ndf=64)
    opt = models.default_opt("image", **kw)
    rng = np.random.default_rng(3)
    pG = util.params_flat(util.weights_init(util.describe_netG(opt), rng))
    pD = util.params_flat(util.weights_init(util.describe_netD(opt), rng))
    img = rng.integers(0, 256, (6, 3, 128, 128)).astype(np.uint8)
    real = img.astype(np.float32) / np.float32(255) * np.float32(2) - np.float32(1)
    center = real[:, :, 32:96, 32:96].copy()
    ctx = real.copy()
    for c, v in enumerate(onets.MEAN_FILL):
        ctx[:, c, 36:92, 36:92] = v                         # overlapPred = 4
    res = []
    for mode in ("float", "u8"):
        trn = train.FusedTrainer(opt, precision="bf16")
        trn.set_params(0, pG); trn.set_params(1, pD)
        losses = trn.step_host(ctx, center) if mode == "float" else trn.step_images_u8_host(img)
        res.append((losses, trn.fetch("ctx"), trn.fetch("D.0.in") if False else None, trn.get_grads(1)))
        if mode == "u8":
            trn.step_images_u8_host_async(img); l2 = trn.wait_losses()
            assert np.isfinite(list(l2.values())).all()
        trn.close()
    (l0, c0, _, g0), (l1, c1, _, g1) = res
    assert np.array_equal(c0, c1)                           # identical bf16 real_ctx on both paths
    for k, tol in (("errG_l2", 2e-3), ("errD_real", 2e-3), ("errD", 1.5e-2), ("errG", 4e-2)):
        assert abs(l0[k] - l1[k]) <= tol * max(abs(l0[k]), 1e-3), (k, l0[k], l1[k])
    assert _cos(g0, g1) >= 0.97
