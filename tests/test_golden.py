"""Golden vectors (tests/golden/ops_golden.npz, made by tests/golden/make_golden.py).
CPU: the oracle reproduces them (guards the restatement against drift).  GPU: the CUDA path, through the C ABI,
reproduces them in fp32 mode (<= 1e-5) -- including a two-step run of the whole G+D step on a miniature network."""
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import nets as onets
from oracle import ops
from oracle import step as ostep

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_golden.npz"))
D = lambda k: G[k].astype(np.float64)
MINI = {"image": dict(batchSize=2, nBottleneck=16, nef=4, ngf=4, ndf=4), "video": dict(batchSize=2, nBottleneck=16, nef=4, ngf=4, ndf=4, predLen=2, wtgdl=0.5)}


def test_oracle_reproduces_golden_ops():
    for tag in ("conv_s2", "conv_v4"):
        k, d, p = [int(v) for v in G[tag + "_geom"]]
        assert np.allclose(ops.conv_forward(D(tag + "_x"), D(tag + "_w"), D(tag + "_b"), d, d, p, p), G[tag + "_y"], atol=1e-12)
        assert np.allclose(ops.conv_grad_input(G[tag + "_x"].shape, D(tag + "_gy"), D(tag + "_w"), d, d, p, p), G[tag + "_gx"], atol=1e-12)
    for tag in ("full_s2", "full_v4"):
        k, d, p = [int(v) for v in G[tag + "_geom"]]
        assert np.allclose(ops.fullconv_forward(D(tag + "_x"), D(tag + "_w"), D(tag + "_b"), d, d, p, p), G[tag + "_y"], atol=1e-12)
        gw, gb = np.zeros(G[tag + "_w"].shape), np.zeros(G[tag + "_b"].shape)
        ops.fullconv_acc_grad(D(tag + "_x"), D(tag + "_gy"), gw, gb, d, d, p, p)
        assert np.allclose(gw, G[tag + "_gw"], atol=1e-11) and np.allclose(gb, G[tag + "_gb"], atol=1e-11)
    assert ops.gdl_forward(D("crit_x"), D("crit_t")) == pytest.approx(float(G["gdl_loss"]), rel=1e-12)
    assert np.allclose(ops.gdl_backward(D("crit_x"), D("crit_t")), G["gdl_grad"], atol=1e-15)
    assert ops.masked_mse_forward(D("crit_x"), D("crit_t"), G["crit_mask"], 0.05) == pytest.approx(float(G["mmse_loss"]), rel=1e-12)
    assert ops.bce_forward(D("bce_x"), D("bce_t")) == pytest.approx(float(G["bce_loss"]), rel=1e-12)


@pytest.mark.parametrize("variant", ["image", "video"])
def test_oracle_reproduces_golden_step(variant):
    orc = ostep.StepOracle(onets.default_opt(variant, **MINI[variant]), seed=77, dtype=np.float64)
    assert np.array_equal(orc.pG, G["step_%s_pG0" % variant])
    for i in range(2):
        batch = [G["step_%s_in%d_%d" % (variant, i, j)] for j in range(2 if variant == "image" else 3)]
        lo = orc.step(*batch)
        ref = G["step_%s_losses" % variant][i]
        assert np.allclose([lo["errD"], lo["errG"], lo["errG_l2"], lo["errG_gdl"] or 0.0, lo["errG_total"]], ref, rtol=1e-10)
    assert np.allclose(orc.pG, G["step_%s_pG2" % variant], atol=1e-12) and np.allclose(orc.pD, G["step_%s_pD2" % variant], atol=1e-12)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_ops_reproduce_golden(cenn):
    import ctypes as C
    from video_filler_b200 import nn, optim
    cenn.set_precision("fp32")
    dev = lambda a: cenn.CudaTensor.from_numpy(np.asarray(a, np.float32))
    for tag, cls in (("conv_s2", nn.SpatialConvolution), ("conv_v4", nn.SpatialConvolution), ("full_s2", nn.SpatialFullConvolution), ("full_v4", nn.SpatialFullConvolution)):
        k, d, p = [int(v) for v in G[tag + "_geom"]]
        w = G[tag + "_w"]
        nin, nout = (w.shape[1], w.shape[0]) if cls is nn.SpatialConvolution else (w.shape[0], w.shape[1])
        m = cls(nin, nout, k, k, d, d, p, p)
        m.weight.copy_(w); m.bias.copy_(G[tag + "_b"])
        x, gy = dev(G[tag + "_x"]), dev(G[tag + "_gy"])
        assert rel_err(m.forward(x).numpy(), G[tag + "_y"]) <= 1e-5, tag
        assert rel_err(m.backward(x, gy).numpy(), G[tag + "_gx"]) <= 1e-5, tag
        assert rel_err(m.gradWeight.numpy(), G[tag + "_gw"]) <= 1e-5 and rel_err(m.gradBias.numpy(), G[tag + "_gb"]) <= 1e-5, tag
    bn = nn.SpatialBatchNormalization(6)
    bn.weight.copy_(G["bn_gamma"]); bn.bias.copy_(G["bn_beta"])
    x, gy = dev(G["bn_x"]), dev(G["bn_gy"])
    assert rel_err(bn.forward(x).numpy(), G["bn_y"]) <= 1e-5
    assert rel_err(bn.running_mean.numpy(), G["bn_running_mean"]) <= 1e-5 and rel_err(bn.running_var.numpy(), G["bn_running_var"]) <= 1e-5
    assert rel_err(bn.backward(x, gy).numpy(), G["bn_gx"]) <= 2e-5
    assert rel_err(bn.gradWeight.numpy(), G["bn_ggamma"]) <= 2e-5 and rel_err(bn.gradBias.numpy(), G["bn_gbeta"]) <= 2e-5
    x, t = dev(G["crit_x"]), dev(G["crit_t"])
    c = nn.GDLCriterion(1)
    assert c.forward(x, t) == pytest.approx(float(G["gdl_loss"]), rel=1e-5)
    assert rel_err(c.backward(x, t).numpy(), G["gdl_grad"]) <= 1e-6
    c = nn.MaskedMSECriterion(0.05); c.setMask(G["crit_mask"])
    assert c.forward(x, t) == pytest.approx(float(G["mmse_loss"]), rel=1e-5)
    assert rel_err(c.backward(x, t).numpy(), G["mmse_grad"]) <= 1e-6
    c = nn.BCECriterion()
    bx, bt = dev(G["bce_x"]), dev(G["bce_t"])
    assert c.forward(bx, bt) == pytest.approx(float(G["bce_loss"]), rel=1e-5)
    assert rel_err(c.backward(bx, bt).numpy(), G["bce_grad"]) <= 1e-5
    xs = dev(G["adam_x0"])
    st = {"learningRate": 2e-3, "beta1": 0.5}
    for i in range(3):
        gi = dev(G["adam_grads"][i])
        optim.adam(lambda _: (0.0, gi), xs, st)
    assert rel_err(xs.numpy(), G["adam_x3"]) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["image", "video"])
def test_cuda_step_reproduces_golden(cenn, variant):
    from video_filler_b200 import models, train
    cenn.set_precision("fp32")
    trn = train.ClosureTrainer(models.default_opt(variant, **MINI[variant]), seed=5)
    trn.parametersG.copy_(G["step_%s_pG0" % variant]); trn.parametersD.copy_(G["step_%s_pD0" % variant])
    for i in range(2):
        batch = [G["step_%s_in%d_%d" % (variant, i, j)] for j in range(2 if variant == "image" else 3)]
        lg = trn.step(*batch)
        ref = G["step_%s_losses" % variant][i]
        got = [lg["errD"], lg["errG"], lg["errG_l2"], lg["errG_gdl"] or 0.0, lg["errG_total"]]
        assert np.allclose(got, ref, rtol=2e-4 if i == 0 else 2e-2), (i, got, ref)
