"""CPU: the oracle against an independent engine (PyTorch-CPU autograd, fp64) -- SURVEY.md 4.1."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops

RNG = np.random.default_rng(7)
T = lambda a, g=False: torch.tensor(a, dtype=torch.float64, requires_grad=g)


@pytest.mark.parametrize("k,d,p,H", [(4, 2, 1, 10), (4, 1, 0, 4), (5, 2, 2, 9), (3, 1, 1, 6), (4, 2, 1, 2)])
def test_conv_matches_autograd(k, d, p, H):
    x, w, b = RNG.normal(size=(3, 5, H, H)), RNG.normal(size=(7, 5, k, k)), RNG.normal(size=7)
    tx, tw, tb = T(x, True), T(w, True), T(b, True)
    y = F.conv2d(tx, tw, tb, stride=d, padding=p)
    gy = RNG.normal(size=tuple(y.shape))
    y.backward(T(gy))
    assert np.allclose(ops.conv_forward(x, w, b, d, d, p, p), y.detach().numpy(), atol=1e-12)
    assert np.allclose(ops.conv_grad_input(x.shape, gy, w, d, d, p, p), tx.grad.numpy(), atol=1e-12)
    gw, gb = np.zeros_like(w), np.zeros_like(b)
    ops.conv_acc_grad(x, gy, gw, gb, d, d, p, p, scale=0.5)
    ops.conv_acc_grad(x, gy, gw, gb, d, d, p, p, scale=0.5)   # accumulating
    assert np.allclose(gw, tw.grad.numpy(), atol=1e-11) and np.allclose(gb, tb.grad.numpy(), atol=1e-11)


@pytest.mark.parametrize("k,d,p,H", [(4, 2, 1, 5), (4, 1, 0, 1), (4, 2, 1, 1), (3, 2, 1, 4)])
def test_fullconv_matches_autograd(k, d, p, H):
    x, w, b = RNG.normal(size=(2, 6, H, H)), RNG.normal(size=(6, 4, k, k)), RNG.normal(size=4)
    tx, tw, tb = T(x, True), T(w, True), T(b, True)
    y = F.conv_transpose2d(tx, tw, tb, stride=d, padding=p)
    gy = RNG.normal(size=tuple(y.shape))
    y.backward(T(gy))
    assert np.allclose(ops.fullconv_forward(x, w, b, d, d, p, p), y.detach().numpy(), atol=1e-12)
    assert np.allclose(ops.fullconv_grad_input(gy, w, d, d, p, p), tx.grad.numpy(), atol=1e-12)
    gw, gb = np.zeros_like(w), np.zeros_like(b)
    ops.fullconv_acc_grad(x, gy, gw, gb, d, d, p, p)
    assert np.allclose(gw, tw.grad.numpy(), atol=1e-11) and np.allclose(gb, tb.grad.numpy(), atol=1e-11)


@pytest.mark.parametrize("shape", [(4, 5, 6, 6), (8, 3, 1, 1), (2, 7, 3, 5)])
def test_bn_train_matches_autograd(shape):
    C = shape[1]
    x, g, be = RNG.normal(size=shape), RNG.normal(size=C), RNG.normal(size=C)
    rm, rv = RNG.normal(size=C), RNG.uniform(0.5, 2, size=C)
    trm, trv = T(rm.copy()), T(rv.copy())
    tx, tg, tb = T(x, True), T(g, True), T(be, True)
    y = F.batch_norm(tx, trm, trv, tg, tb, True, 0.1, 1e-5)
    gy = RNG.normal(size=shape)
    y.backward(T(gy))
    yo, sm, si = ops.bn_forward(x, g, be, rm, rv, True)
    assert np.allclose(yo, y.detach().numpy(), atol=1e-10)
    assert np.allclose(rm, trm.numpy(), atol=1e-12) and np.allclose(rv, trv.numpy(), atol=1e-12)
    gg, gb = np.zeros(C), np.zeros(C)
    gx = ops.bn_backward(x, gy, g, sm, si, rm, rv, True, ggamma=gg, gbeta=gb)
    assert np.allclose(gx, tx.grad.numpy(), atol=1e-9)
    assert np.allclose(gg, tg.grad.numpy(), atol=1e-9) and np.allclose(gb, tb.grad.numpy(), atol=1e-9)


def test_bn_eval_matches_autograd():
    x, g, be = RNG.normal(size=(3, 4, 5, 5)), RNG.normal(size=4), RNG.normal(size=4)
    rm, rv = RNG.normal(size=4), RNG.uniform(0.5, 2, size=4)
    tx = T(x, True)
    y = F.batch_norm(tx, T(rm), T(rv), T(g), T(be), False, 0.1, 1e-5)
    gy = RNG.normal(size=x.shape)
    y.backward(T(gy))
    rm0, rv0 = rm.copy(), rv.copy()
    yo, _, _ = ops.bn_forward(x, g, be, rm, rv, False)
    assert np.allclose(yo, y.detach().numpy(), atol=1e-12)
    assert np.array_equal(rm, rm0) and np.array_equal(rv, rv0)
    gx = ops.bn_backward(x, gy, g, None, None, rm, rv, False)
    assert np.allclose(gx, tx.grad.numpy(), atol=1e-12)


def test_bn_single_element_running_var_is_inf():
    # n == 1 in train mode divides by zero in running_var
    rm, rv = np.zeros(2), np.ones(2)
    ops.bn_forward(np.ones((1, 2, 1, 1)), np.ones(2), np.zeros(2), rm, rv, True)
    assert np.all(np.isinf(rv))


def test_activations_and_bce_mse():
    x = RNG.normal(size=(4, 3, 5, 5))
    gy = RNG.normal(size=x.shape)
    tx = T(x, True)
    F.leaky_relu(tx, 0.2).backward(T(gy))
    assert np.allclose(ops.leaky_relu_grad(x, gy, 0.2), tx.grad.numpy())
    assert ops.leaky_relu_grad(np.zeros(1), np.ones(1), 0.2)[0] == pytest.approx(0.2)   # slope 0.2 at x == 0
    assert np.allclose(ops.tanh_grad(np.tanh(x), gy), (T(gy) * (1 - torch.tanh(T(x)) ** 2)).numpy())
    s = ops.sigmoid(x)
    assert np.allclose(ops.sigmoid_grad(s, gy), gy * s * (1 - s))
    p = RNG.uniform(0.01, 0.99, size=(16, 1))
    t = (RNG.uniform(size=16) > 0.5).astype(np.float64)
    tp = T(p, True)
    L = F.binary_cross_entropy(tp.view(-1), T(t))
    L.backward()
    assert ops.bce_forward(p, t) == pytest.approx(L.item(), rel=1e-9)
    assert np.allclose(ops.bce_backward(p, t), tp.grad.numpy(), rtol=1e-8)
    a, b = RNG.normal(size=(2, 3, 4, 4)), RNG.normal(size=(2, 3, 4, 4))
    assert ops.mse_forward(a, b) == pytest.approx(float(((a - b) ** 2).mean()))
    assert np.allclose(ops.mse_backward(a, b), 2 * (a - b) / a.size)


def test_masked_mse_matches_autograd():
    x, t = RNG.normal(size=(2, 3, 6, 6)), RNG.normal(size=(2, 3, 6, 6))
    m = (RNG.uniform(size=x.shape) > 0.7).astype(np.uint8)
    mW = 0.05
    tx = T(x, True)
    wM = T(m.astype(np.float64)) * (1 - mW) + mW
    L = (wM * (tx - T(t)) ** 2).abs().mean()
    L.backward()
    assert ops.masked_mse_forward(x, t, m, mW) == pytest.approx(L.item())
    assert np.allclose(ops.masked_mse_backward(x, t, m, mW), tx.grad.numpy())


def _gdl_torch(inp, tgt):
    def terms(Tn):
        B, C, H, W = Tn.shape
        return (Tn[:, :, :H - 1, :].reshape(B, C, -1), Tn[:, :, 1:, :].reshape(B, C, -1),
                Tn[:, :, :, :W - 1].reshape(B, C, -1), Tn[:, :, :, 1:].reshape(B, C, -1))
    yi1, yj1, yi2, yj2 = terms(tgt)
    hi1, hj1, hi2, hj2 = terms(inp)
    return (((yi2 - yi1).abs() - (hi2 - hi1).abs()).abs().mean() + ((yj2 - yj1).abs() - (hj2 - hj1).abs()).abs().mean())


@pytest.mark.parametrize("H", [2, 5, 8])
def test_gdl_flat_index_matches_autograd(H):
    a, t = RNG.normal(size=(2, 3, H, H)), RNG.normal(size=(2, 3, H, H))
    ta = T(a, True)
    L = _gdl_torch(ta, T(t))
    L.backward()
    assert ops.gdl_forward(a, t) == pytest.approx(L.item())
    assert np.allclose(ops.gdl_backward(a, t), ta.grad.numpy())


def test_gdl_differs_from_textbook_and_rejects_non_square():
    a, t = RNG.normal(size=(1, 1, 6, 6)), RNG.normal(size=(1, 1, 6, 6))
    textbook = (np.abs(np.abs(np.diff(t, axis=2)) - np.abs(np.diff(a, axis=2))).mean()
                + np.abs(np.abs(np.diff(t, axis=3)) - np.abs(np.diff(a, axis=3))).mean())
    assert abs(ops.gdl_forward(a, t) - textbook) > 1e-3       # follow the flat-index text, not the paper
    with pytest.raises(ValueError):
        ops.gdl_forward(np.zeros((1, 1, 4, 6)), np.zeros((1, 1, 4, 6)))


def test_blend_and_composite_and_adam():
    x, t, g = RNG.normal(size=(2, 3, 16, 16)), RNG.normal(size=(2, 3, 16, 16)), RNG.normal(size=(2, 3, 16, 16))
    out = ops.blend_l2_overlap(g, x, t, 0.999, 4)
    W = ops.overlap_weight_matrix(x.shape, 4, 0.999, np.float64)
    assert W[0, 0, 0, 0] == pytest.approx(9.99) and W[0, 0, 8, 8] == pytest.approx(0.999) and W[0, 0, 3, 8] == pytest.approx(9.99)
    assert np.allclose(out, g * 0.001 + W * 2 * (x - t) / x.size)
    assert np.allclose(ops.blend_l2_overlap(g, x, t, 2.0, 0), g + 2.0 * 2 * (x - t) / x.size)
    m = (RNG.uniform(size=x.shape) > 0.5).astype(np.float64)
    out, w = ops.blend_l2_masked(g, x, t, m, 0.999, 0.05)
    assert np.allclose(w, m * 0.95 + 0.05)
    assert np.allclose(out, g * 0.001 + 0.999 * w * 2 * (x - t) / x.size)
    assert np.array_equal(ops.mask_composite(x, m, t), np.where(m != 0, t, x))
    # adam vs torch.optim.Adam (same algorithm up to where eps is added: sqrt(v)+eps unscaled in optim.adam)
    p = RNG.normal(size=50)
    st = {}
    p1 = p.copy()
    m_, v_ = np.zeros(50), np.zeros(50)
    for step in range(1, 4):
        gr = RNG.normal(size=50)
        ops.adam_step(p1, gr, st, 2e-4, 0.5)
        m_ = 0.5 * m_ + 0.5 * gr
        v_ = 0.999 * v_ + 0.001 * gr * gr
        p -= 2e-4 * np.sqrt(1 - 0.999 ** step) / (1 - 0.5 ** step) * m_ / (np.sqrt(v_) + 1e-8)
    assert np.allclose(p1, p, atol=1e-15)


def test_parallel_join_tables_match_autograd():
    """nn.ParallelTable + nn.JoinTable(2) as used by conditionAdv (train.lua:158-180): two 5x5/s2 convs (pad 2 on the 16x16
    'context', pad 2+4 on the 8x8 'prediction' so that both give 8x8 maps), joined along the channel axis, LeakyReLU."""
    from oracle import nn as onn
    net = onn.Sequential()
    a = onn.SpatialConvolution(3, 4, 5, 5, 2, 2, 2, 2, dtype=np.float64)
    b = onn.SpatialConvolution(3, 4, 5, 5, 2, 2, 2 + 4, 2 + 4, dtype=np.float64)
    net.add(onn.ParallelTable().add(onn.Sequential().add(a)).add(onn.Sequential().add(b))).add(onn.JoinTable(2)).add(onn.LeakyReLU(0.2, False))
    for m in (a, b):
        m.weight[...] = RNG.normal(size=m.weight.shape)
        m.bias[...] = RNG.normal(size=m.bias.shape)
    xa, xb = RNG.normal(size=(2, 3, 16, 16)), RNG.normal(size=(2, 3, 8, 8))
    ta, tb = T(xa, True), T(xb, True)
    wa, ba, wb, bb = T(a.weight, True), T(a.bias, True), T(b.weight, True), T(b.bias, True)
    y = F.leaky_relu(torch.cat([F.conv2d(ta, wa, ba, stride=2, padding=2), F.conv2d(tb, wb, bb, stride=2, padding=6)], dim=1), 0.2)
    out = net.forward([xa, xb])
    assert out.shape == (2, 8, 8, 8) and np.allclose(out, y.detach().numpy(), atol=1e-12)
    gy = RNG.normal(size=out.shape)
    y.backward(T(gy))
    net.zeroGradParameters()
    gin = net.backward([xa, xb], gy)
    assert np.allclose(gin[0], ta.grad.numpy(), atol=1e-11) and np.allclose(gin[1], tb.grad.numpy(), atol=1e-11)
    assert np.allclose(a.gradWeight, wa.grad.numpy(), atol=1e-10) and np.allclose(b.gradWeight, wb.grad.numpy(), atol=1e-10)
    assert np.allclose(a.gradBias, ba.grad.numpy(), atol=1e-10) and np.allclose(b.gradBias, bb.grad.numpy(), atol=1e-10)
    # getParameters walks the table members in order (train.lua:262-263)
    flat, _ = net.getParameters()
    assert flat.size == 2 * (4 * 3 * 25 + 4)
