"""CPU: the PyTorch-CPU engine of the oracle (oracle/torch_engine.py, used by bench.py's cpu_baseline / reference arm and by the
benchmark-shape GPU tests) against the numpy functions it replaces, op by op and over whole steps."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import nets, ops, step, torch_engine


@pytest.fixture(autouse=True)
def _engine_off():
    torch_engine.disable()
    yield
    torch_engine.disable()


def test_engine_ops_match_numpy_ops():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 5, 12, 12)); w = rng.normal(size=(7, 5, 4, 4)); b = rng.normal(size=7)
    gy = rng.normal(size=(3, 7, 6, 6))
    geo = (2, 2, 1, 1)
    assert rel_err(torch_engine.conv_forward(x, w, b, *geo), ops.conv_forward(x, w, b, *geo)) <= 1e-12
    assert rel_err(torch_engine.conv_grad_input(x.shape, gy, w, *geo), ops.conv_grad_input(x.shape, gy, w, *geo)) <= 1e-12
    g1, b1, g2, b2 = np.ones_like(w), np.ones(7), np.ones_like(w), np.ones(7)
    torch_engine.conv_acc_grad(x, gy, g1, b1, *geo, scale=0.5); ops.conv_acc_grad(x, gy, g2, b2, *geo, scale=0.5)
    assert rel_err(g1, g2) <= 1e-12 and rel_err(b1, b2) <= 1e-12
    # transposed conv: weight [Cin, Cout, 4, 4]
    xs = rng.normal(size=(3, 7, 6, 6)); wt = rng.normal(size=(7, 5, 4, 4)); bt = rng.normal(size=5); gyl = rng.normal(size=(3, 5, 12, 12))
    assert rel_err(torch_engine.fullconv_forward(xs, wt, bt, *geo), ops.fullconv_forward(xs, wt, bt, *geo)) <= 1e-12
    assert rel_err(torch_engine.fullconv_grad_input(gyl, wt, *geo), ops.fullconv_grad_input(gyl, wt, *geo)) <= 1e-12
    g1, b1, g2, b2 = np.zeros_like(wt), np.zeros(5), np.zeros_like(wt), np.zeros(5)
    torch_engine.fullconv_acc_grad(xs, gyl, g1, b1, *geo); ops.fullconv_acc_grad(xs, gyl, g2, b2, *geo)
    assert rel_err(g1, g2) <= 1e-12 and rel_err(b1, b2) <= 1e-12
    # 1x1 -> 4x4 bottleneck decoder (stride 1, pad 0)
    x1 = rng.normal(size=(3, 7, 1, 1))
    assert rel_err(torch_engine.fullconv_forward(x1, wt, None, 1, 1, 0, 0), ops.fullconv_forward(x1, wt, None, 1, 1, 0, 0)) <= 1e-12
    # batch norm, train and eval, forward and backward
    gamma, beta = rng.normal(1, 0.1, 5), rng.normal(0, 0.1, 5)
    for train in (True, False):
        rm1, rv1, rm2, rv2 = np.full(5, 0.1), np.full(5, 1.3), np.full(5, 0.1), np.full(5, 1.3)
        y1, m1, s1 = torch_engine.bn_forward(x, gamma, beta, rm1, rv1, train)
        y2, m2, s2 = ops.bn_forward(x, gamma, beta, rm2, rv2, train)
        assert rel_err(y1, y2) <= 1e-12 and rel_err(m1, m2) <= 1e-12 and rel_err(s1, s2) <= 1e-12
        assert rel_err(rm1, rm2) <= 1e-12 and rel_err(rv1, rv2) <= 1e-12
        gyb = rng.normal(size=x.shape)
        gg1, gb1, gg2, gb2 = np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(5)
        gx1 = torch_engine.bn_backward(x, gyb, gamma, m1, s1, rm1, rv1, train, ggamma=gg1, gbeta=gb1)
        gx2 = ops.bn_backward(x, gyb, gamma, m2, s2, rm2, rv2, train, ggamma=gg2, gbeta=gb2)
        assert rel_err(gx1, gx2) <= 1e-11 and rel_err(gg1, gg2) <= 1e-11 and rel_err(gb1, gb2) <= 1e-11
        if train:        # the running statistics are optional in training mode (the block checks of tests/test_fused_gpu.py pass None)
            assert rel_err(torch_engine.bn_backward(x, gyb, gamma, m1, s1, None, None, True), gx2) <= 1e-11
    for f in ("leaky_relu", "relu", "tanh"):
        assert rel_err(getattr(torch_engine, f)(x), getattr(ops, f)(x)) <= 1e-15
    assert rel_err(torch_engine.leaky_relu_grad(x, gyb), ops.leaky_relu_grad(x, gyb)) <= 1e-15
    assert rel_err(torch_engine.relu_grad(x, gyb), ops.relu_grad(x, gyb)) <= 1e-15
    y = np.tanh(x)
    assert rel_err(torch_engine.tanh_grad(y, gyb), ops.tanh_grad(y, gyb)) <= 1e-15
    p1, p2, g = rng.normal(size=100), None, rng.normal(size=100)
    p2 = p1.copy(); s1, s2 = {}, {}
    for _ in range(3):
        torch_engine.adam_step(p1, g, s1, 2e-4, 0.5); ops.adam_step(p2, g, s2, 2e-4, 0.5)
    assert rel_err(p1, p2) <= 1e-14


@pytest.mark.parametrize("variant,kw", [("image", dict(batchSize=4, nBottleneck=64, nef=16, ngf=16, ndf=16)),
                                        ("video", dict(batchSize=3, nBottleneck=64, nef=16, ngf=16, ndf=16, predLen=2, wtgdl=0.5)),
                                        ("video", dict(batchSize=3, nBottleneck=64, nef=16, ngf=16, ndf=16, predLen=1, weight_nomask=0.0))])
def test_engine_steps_match_numpy_steps(variant, kw):
    """Three whole G+D steps (fp64): losses, parameters and gradients of the engine run equal the numpy run."""
    res = []
    for eng in (False, True):
        torch_engine.enable() if eng else torch_engine.disable()
        o = step.StepOracle(nets.default_opt(variant, **kw), seed=1, dtype=np.float64)
        rng = np.random.default_rng(3)
        ls = [o.step(*o.synth_batch(rng)) for _ in range(3)]
        res.append((ls, o.pG.copy(), o.pD.copy(), o.gG.copy(), o.gD.copy()))
    for a, b in zip(res[0][0], res[1][0]):
        for k in a:
            if a[k] is not None:
                assert abs(a[k] - b[k]) <= 1e-9 * max(1.0, abs(a[k])), (k, a[k], b[k])
    for i in (1, 2, 3, 4):
        assert rel_err(res[1][i], res[0][i]) <= 1e-9


def test_enable_is_reversible():
    f = ops.conv_forward
    torch_engine.enable()
    assert ops.conv_forward is torch_engine.conv_forward and torch_engine.enabled()
    torch_engine.disable()
    assert ops.conv_forward is f and not torch_engine.enabled()
