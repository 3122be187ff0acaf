"""GPU parity of train.lua's optional branches inside the fused whole-step executor:

* ``noiseGen``     (train.lua:109-124): 1x1 conv of a noise vector joined to the bottleneck before BN(nBottleneck + nz);
* ``conditionAdv`` (train.lua:158-180): netD takes {context, prediction}; two 5x5 / stride-2 first-layer convs (pad 2 on the
  128 x 128 context, pad 2+32 on the 64 x 64 prediction) joined along the channel axis.

Same method as tests/test_fused_gpu.py: losses against the fp64 oracle, plus SELF-CONSISTENCY of the new kernels -- every tensor
they produce is recomputed in fp64 with the oracle's formulas from the executor's own stored inputs (BF16 per-layer bound 2e-2,
weight gradients 1e-4 because they are fp32 sums of products of stored bf16 values).
"""
import numpy as np
import pytest

from conftest import rel_err
from oracle import nets as onets
from oracle import ops
from oracle import step as ostep

pytestmark = pytest.mark.gpu

B, NB, NZ = 8, 256, 100


def _pair(**extra):
    from video_filler_b200 import models, train
    kw = dict(batchSize=B, nBottleneck=NB, nef=64, ngf=64, ndf=64)
    kw.update(extra)
    orc = ostep.StepOracle(onets.default_opt("image", **kw), seed=1234, dtype=np.float64)
    trn = train.FusedTrainer(models.default_opt("image", **kw), precision="bf16")
    assert trn.param_count(0) == orc.pG.size and trn.param_count(1) == orc.pD.size
    trn.set_params(0, orc.pG)
    trn.set_params(1, orc.pD)
    return orc, trn


def _flat(net):
    out = []
    for m in net.modules:
        out += _flat(m) if hasattr(m, "modules") else [m]
    return out


def _offsets(mods):
    offs, off = {}, 0
    for m in mods:
        if getattr(m, "weight", None) is not None:
            offs[id(m)] = off
            off += m.weight.size + m.bias.size
    return offs


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.dot(a, b) / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def _batch(orc, rng, noise_gen):
    batch = orc.synth_batch(rng)
    noise = rng.uniform(-1, 1, (B, NZ, 1, 1)) if noise_gen else None      # noise:uniform(-1, 1), train.lua:319-320
    return batch, noise


@pytest.fixture
def fast_oracle():
    from oracle import torch_engine
    torch_engine.enable()
    yield
    torch_engine.disable()


# (last case: the shipped defaults nBottleneck 4000 + nz 100 -> 4100 joined channels, pitch 4104: the > 4096-channel BN path)
@pytest.mark.parametrize("extra", [{"noiseGen": 1, "nz": NZ}, {"conditionAdv": 1}, {"noiseGen": 1, "nz": NZ, "conditionAdv": 1},
                                   {"noiseGen": 1, "nz": NZ, "nBottleneck": 4000}])
def test_fused_step_with_optional_branches_matches_oracle(cenn, fast_oracle, extra):
    orc, trn = _pair(**extra)
    batch, noise = _batch(orc, np.random.default_rng(4321), extra.get("noiseGen"))
    args = tuple(batch) + ((noise,) if noise is not None else ())
    # parameter round trip through the padded master layout (Module:getParameters order incl. the extra branches)
    assert np.array_equal(trn.get_params(0), orc.pG.astype(np.float32)) and np.array_equal(trn.get_params(1), orc.pD.astype(np.float32))
    lo = orc.step(*args)
    lg = trn.step_host(*batch, noise=noise)
    for k in ("errD_real", "errG_l2", "errG_total"):
        assert lg[k] == pytest.approx(lo[k], rel=2e-2), k
    for k in ("errD_fake", "errD", "errG"):
        assert lg[k] == pytest.approx(lo[k], rel=5e-2), k
    assert rel_err(trn.fetch("fake").reshape(orc.netG.output.shape), orc.netG.output) <= 4e-2
    gG, gD = trn.get_grads(0), trn.get_grads(1)
    assert np.all(np.isfinite(gG)) and np.all(np.isfinite(gD))
    assert _cos(gG, orc.gG) >= 0.9 and _cos(gD, orc.gD) >= 0.9
    assert np.linalg.norm(gG) == pytest.approx(np.linalg.norm(orc.gG), rel=5e-2)
    assert np.linalg.norm(gD) == pytest.approx(np.linalg.norm(orc.gD), rel=5e-2)
    ref = np.concatenate([np.concatenate([m.running_mean, m.running_var]) for m in _flat(orc.netG) if hasattr(m, "running_mean")])
    assert rel_err(trn.get_bn_stats(0), ref) <= 2e-2
    ref = np.concatenate([np.concatenate([m.running_mean, m.running_var]) for m in _flat(orc.netD) if hasattr(m, "running_mean")])
    assert rel_err(trn.get_bn_stats(1), ref) <= 3e-2
    # a second step runs from the updated weights (graph replay path) and stays finite / close on the L2 term
    batch2, noise2 = _batch(orc, np.random.default_rng(99), extra.get("noiseGen"))
    lo2 = orc.step(*(tuple(batch2) + ((noise2,) if noise2 is not None else ())))
    lg2 = trn.step_host(*batch2, noise=noise2)
    assert lg2["errG_l2"] == pytest.approx(lo2["errG_l2"], rel=3e-2)
    assert all(np.isfinite(v) for v in lg2.values())
    trn.close()


def test_noise_branch_kernels_self_consistent(cenn, fast_oracle):
    """noise_fwd_kernel / noise_wgrad_kernel and the joined BN: recomputed from the executor's stored tensors."""
    orc, trn = _pair(noiseGen=1, nz=NZ)
    batch, noise = _batch(orc, np.random.default_rng(7), True)
    pG0 = orc.pG.copy()
    orc.step(*batch, noise)
    trn.step_host(*batch, noise=noise)
    q = ops.bf16_round
    mods = _flat(orc.netG)
    offs = _offsets(mods)
    nconv = orc.netG.modules[0].modules[1].modules[0]                  # ParallelTable -> netG_noise -> its 1x1 conv
    assert nconv.weight.shape == (NZ, NZ, 1, 1)
    o = offs[id(nconv)]
    w = q(pG0[o:o + NZ * NZ].reshape(NZ, NZ))
    y = trn.fetch("G.5.y").reshape(B, NB + NZ).astype(np.float64)      # the joined tensor [encoder | noise conv] before BN
    y_noise = q(noise.reshape(B, NZ).astype(np.float64)) @ w.T        # conv biases are zero during training (train.lua:279-280)
    assert rel_err(y[:, NB:], y_noise) <= 5e-3                         # one bf16 rounding of the stored value
    # the encoder half against the oracle's encoder output (6 bf16 layers deep)
    enc = orc.netG.modules[0].modules[0].output.reshape(B, NB)
    assert rel_err(y[:, :NB], enc) <= 4e-2
    # joined BN + LeakyReLU over nBottleneck + nz channels from the stored y
    bn = orc.netG.modules[2]
    ob = offs[id(bn)]
    gamma, beta = pG0[ob:ob + NB + NZ], pG0[ob + NB + NZ:ob + 2 * (NB + NZ)]
    z, mean, invstd = ops.bn_forward(y.reshape(B, NB + NZ, 1, 1), gamma, beta, np.zeros(NB + NZ), np.ones(NB + NZ), True)
    a = trn.fetch("G.5.a").reshape(B, NB + NZ).astype(np.float64)
    assert rel_err(a, ops.leaky_relu(z, 0.2).reshape(B, NB + NZ)) <= 1e-2
    # weight / bias gradient of the noise conv from the stored g_y (gradient w.r.t. the joined conv output)
    g_y = trn.fetch("G.5.g").reshape(B, NB + NZ).astype(np.float64)
    gG = trn.get_grads(0)
    gw_ref = g_y[:, NB:].T @ q(noise.reshape(B, NZ).astype(np.float64))
    assert rel_err(gG[o:o + NZ * NZ].reshape(NZ, NZ), gw_ref) <= 1e-4
    # (the bias gradient behind a batch-statistics BN is a sum that cancels to ~0: absolute bound, a bf16 ulp of the summands)
    assert np.abs(gG[o + NZ * NZ:o + NZ * NZ + NZ] - g_y[:, NB:].sum(0)).max() <= 8e-3 * np.abs(g_y[:, NB:]).max()
    # the encoder's bottleneck conv sees only its own columns of g_y (the K extent of its dgrad / wgrad GEMMs is nBottleneck, the pitch NB + NZ)
    e6 = [m for m in _flat(orc.netG.modules[0].modules[0]) if "Convolution" in type(m).__name__][-1]
    oe = offs[id(e6)]
    xin = trn.fetch("G.5.in").reshape(B, -1, 4, 4).astype(np.float64)
    gw = np.zeros(e6.weight.shape)
    ops.conv_acc_grad(xin, g_y[:, :NB].reshape(B, NB, 1, 1), gw, np.zeros(NB), 1, 1, 0, 0)
    assert rel_err(gG[oe:oe + gw.size], gw) <= 1e-4
    # ... and so does its dgrad: E5's stored gradient = BN / LeakyReLU backward of conv_grad_input(g_y[:, :NB]) -- a noise column leaking
    # into the reduction would show here
    w6 = q(pG0[oe:oe + e6.weight.size].reshape(e6.weight.shape))
    g_a = q(ops.conv_grad_input(xin.shape, np.ascontiguousarray(g_y[:, :NB]).reshape(B, NB, 1, 1), w6, 1, 1, 0, 0))
    i6 = [id(m) for m in mods].index(id(e6))
    pbn = mods[i6 - 2]
    assert "BatchNorm" in type(pbn).__name__
    py = trn.fetch("G.4.y").reshape(xin.shape).astype(np.float64)
    ob5 = offs[id(pbn)]
    pgamma = pG0[ob5:ob5 + pbn.weight.size]
    _, pm, pis = ops.bn_forward(py, pgamma, np.zeros_like(pgamma), np.zeros_like(pgamma), np.ones_like(pgamma), True)
    exp = ops.bn_backward(py, g_a * np.where(xin > 0, 1.0, 0.2), pgamma, pm, pis, None, None, True)
    got = trn.fetch("G.4.g").reshape(xin.shape).astype(np.float64)
    assert rel_err(got, exp) <= 1.5e-2 and _cos(got, exp) >= 0.9999
    # the decoder's first layer reads all nBottleneck + nz channels
    g1 = [m for m in mods if type(m).__name__ == "SpatialFullConvolution"][0]
    assert g1.weight.shape[0] == NB + NZ
    og = offs[id(g1)]
    wg1 = q(pG0[og:og + g1.weight.size].reshape(g1.weight.shape))
    y1 = trn.fetch("G.6.y").reshape(g1.output.shape).astype(np.float64)
    assert rel_err(y1, ops.fullconv_forward(a.reshape(B, NB + NZ, 1, 1), wg1, None, 1, 1, 0, 0)) <= 5e-3
    trn.close()


def test_condition_adv_first_layer_self_consistent(cenn, fast_oracle):
    """The joined 5x5 first layer of the conditional discriminator (im2col5 + GEMMs + col2im5): forward of both branches, both
    weight gradients after the REAL sweep, and the gradient w.r.t. the prediction (df_dg[2], train.lua:369-371) after fGx."""
    import video_filler_b200.tensor as T
    orc, trn = _pair(conditionAdv=1)
    ndf = 64
    batch, _ = _batch(orc, np.random.default_rng(11), False)
    pD0 = orc.pD.copy()
    orc.step(*batch)
    q = ops.bf16_round
    mods = _flat(orc.netD)
    offs = _offsets(mods)
    cctx, cpred = mods[0], mods[1]
    assert cctx.weight.shape == (ndf, 3, 5, 5) and (cctx.padH, cpred.padH) == (2, 34)
    oc, op_ = offs[id(cctx)], offs[id(cpred)]
    a_dev = T.CudaTensor.from_numpy(np.ascontiguousarray(batch[0], np.float32))
    b_dev = T.CudaTensor.from_numpy(np.ascontiguousarray(batch[1], np.float32))
    trn.step_until(a_dev.ptr, b_dev.ptr, None, "fold_gbias", 0)        # end of the real sweep
    ctx, real = q(batch[0].astype(np.float64)), q(batch[1].astype(np.float64))
    wc, wp = q(pD0[oc:oc + cctx.weight.size].reshape(cctx.weight.shape)), q(pD0[op_:op_ + cpred.weight.size].reshape(cpred.weight.shape))
    a = trn.fetch("D.0.a").reshape(B, 2 * ndf, 64, 64).astype(np.float64)
    a_ref = ops.leaky_relu(np.concatenate([ops.conv_forward(ctx, wc, None, 2, 2, 2, 2), ops.conv_forward(real, wp, None, 2, 2, 34, 34)], axis=1), 0.2)
    assert rel_err(a, a_ref) <= 5e-3
    g_y = trn.fetch("D.0.g").reshape(B, 2 * ndf, 64, 64).astype(np.float64)
    gD = trn.get_grads(1)
    for conv, o, x, pad, sl in ((cctx, oc, ctx, 2, slice(0, ndf)), (cpred, op_, real, 34, slice(ndf, 2 * ndf))):
        gw, gb = np.zeros(conv.weight.shape), np.zeros(ndf)
        ops.conv_acc_grad(x, np.ascontiguousarray(g_y[:, sl]), gw, gb, 2, 2, pad, pad)
        assert rel_err(gD[o:o + gw.size], gw) <= 1e-4, ("wgrad", pad)
        assert rel_err(gD[o + gw.size:o + gw.size + ndf], gb) <= 3e-3, ("bias grad", pad)
    # complete step: fake sweep forward (context half reused, prediction half recomputed from G's output) and fGx's gradient
    trn.step_device(a_dev.ptr, b_dev.ptr, None)
    T.api().cenn_synchronize(T.state())
    fake = trn.fetch("fake").reshape(B, 3, 64, 64).astype(np.float64)
    a = trn.fetch("D.0.a").reshape(B, 2 * ndf, 64, 64).astype(np.float64)
    a_ref = ops.leaky_relu(np.concatenate([ops.conv_forward(ctx, wc, None, 2, 2, 2, 2), ops.conv_forward(fake, wp, None, 2, 2, 34, 34)], axis=1), 0.2)
    assert rel_err(a, a_ref) <= 5e-3
    pD1 = trn.get_params(1).astype(np.float64)
    wp1 = q(pD1[op_:op_ + cpred.weight.size].reshape(cpred.weight.shape))
    g_y = trn.fetch("D.0.g").reshape(B, 2 * ndf, 64, 64).astype(np.float64)        # fGx sweep: gradient w.r.t. the joined conv output
    df_ref = ops.conv_grad_input(fake.shape, np.ascontiguousarray(g_y[:, ndf:]), wp1, 2, 2, 34, 34)
    df = trn.fetch("df_dg").reshape(B, 3, 64, 64).astype(np.float64)
    assert rel_err(df, df_ref) <= 1e-2 and _cos(df, df_ref) >= 0.9999
    trn.close()
