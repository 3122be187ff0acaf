"""GPU parity: every THNN-shaped entry point of libcenn (through the nn.* mirror and the C ABI)
against the CPU oracle on the same seeded inputs.  fp32 mode: <= 1e-5 (north_star); bf16 mode: <= 2e-2.
Metric: max|a-b| / max|b| per tensor."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from oracle import ops

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "bf16": 2e-2}


@pytest.fixture(params=["fp32", "bf16"])
def mode(request, cenn):
    cenn.set_precision(request.param)
    yield request.param
    cenn.set_precision("fp32")


def dev(cenn, a):
    return cenn.CudaTensor.from_numpy(np.asarray(a, np.float32))


# layer shapes of the nets at reduced batch + odd shapes of the upstream options
CONV_CASES = [
    # N, Cin, H, Cout, k, d, p
    (4, 3, 32, 64, 4, 2, 1),      # E1 / D1 (thin input)
    (4, 64, 32, 64, 4, 2, 1),     # E2
    (4, 64, 16, 128, 4, 2, 1),    # E3 / D2
    (2, 128, 16, 256, 4, 2, 1),   # E4 / D3
    (8, 256, 8, 512, 4, 2, 1),    # E5 / D4
    (8, 512, 4, 200, 4, 1, 0),    # E6 bottleneck (4x4 valid -> 1x1), non-multiple-of-64 Cout
    (8, 512, 4, 1, 4, 1, 0),      # D5 head (Cout = 1)
    (2, 12, 32, 32, 4, 2, 1),     # video D0 (12 -> 32)
    (2, 3, 16, 8, 5, 2, 2),       # conditionAdv 5x5 (train.lua:161)
    (1, 5, 9, 7, 3, 1, 1),        # ragged everything
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_spatial_convolution(cenn, mode, case):
    from video_filler_b200 import nn
    N, Ci, H, Co, k, d, p = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = rng.uniform(-1, 1, (N, Ci, H, H)).astype(np.float32)
    w = rng.normal(0, 0.05, (Co, Ci, k, k)).astype(np.float32)
    b = rng.normal(0, 0.1, Co).astype(np.float32)
    m = nn.SpatialConvolution(Ci, Co, k, k, d, d, p, p)
    m.weight.copy_(w); m.bias.copy_(b)
    y = m.forward(dev(cenn, x)).numpy()
    y_ref = ops.conv_forward(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), d, d, p, p)
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) <= TOL[mode]
    if mode == "bf16" and (k, d, p) == (4, 2, 1):
        assert rel_err(y, y_ref) > 1e-5, "bf16 mode must run the tcgen05 path, not the fp32 kernel"
    gy = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    gw0 = rng.normal(0, 0.01, w.shape).astype(np.float32)   # accumulate on top of existing grads
    gb0 = rng.normal(0, 0.01, b.shape).astype(np.float32)
    m.gradWeight.copy_(gw0); m.gradBias.copy_(gb0)
    gx = m.backward(dev(cenn, x), dev(cenn, gy), 0.5).numpy()
    gx_ref = ops.conv_grad_input(x.shape, gy.astype(np.float64), w.astype(np.float64), d, d, p, p)
    gw_ref, gb_ref = gw0.astype(np.float64), gb0.astype(np.float64)
    ops.conv_acc_grad(x.astype(np.float64), gy.astype(np.float64), gw_ref, gb_ref, d, d, p, p, 0.5)
    assert rel_err(gx, gx_ref) <= TOL[mode]
    assert rel_err(m.gradWeight.numpy(), gw_ref) <= TOL[mode]
    assert rel_err(m.gradBias.numpy(), gb_ref) <= TOL[mode]


FULL_CASES = [
    # N, Cin, H, Cout, k, d, p
    (8, 200, 1, 512, 4, 1, 0),    # G1 (1x1 -> 4x4)
    (8, 512, 4, 256, 4, 2, 1),    # G2
    (4, 256, 8, 128, 4, 2, 1),    # G3
    (4, 128, 16, 64, 4, 2, 1),    # G4
    (4, 64, 32, 3, 4, 2, 1),      # G5 image (thin output)
    (2, 64, 32, 64, 4, 2, 1),     # G5 video
    (2, 64, 32, 12, 4, 2, 1),     # G6 video
    (1, 5, 3, 4, 3, 2, 1),        # ragged
]


@pytest.mark.parametrize("case", FULL_CASES)
def test_spatial_full_convolution(cenn, mode, case):
    from video_filler_b200 import nn
    N, Ci, H, Co, k, d, p = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = rng.uniform(-1, 1, (N, Ci, H, H)).astype(np.float32)
    w = rng.normal(0, 0.05, (Ci, Co, k, k)).astype(np.float32)
    b = rng.normal(0, 0.1, Co).astype(np.float32)
    m = nn.SpatialFullConvolution(Ci, Co, k, k, d, d, p, p)
    m.weight.copy_(w); m.bias.copy_(b)
    y = m.forward(dev(cenn, x)).numpy()
    y_ref = ops.fullconv_forward(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), d, d, p, p)
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) <= TOL[mode]
    gy = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    m.gradWeight.zero(); m.gradBias.zero()
    gx = m.backward(dev(cenn, x), dev(cenn, gy)).numpy()
    gx_ref = ops.fullconv_grad_input(gy.astype(np.float64), w.astype(np.float64), d, d, p, p)
    gw_ref, gb_ref = np.zeros(w.shape), np.zeros(b.shape)
    ops.fullconv_acc_grad(x.astype(np.float64), gy.astype(np.float64), gw_ref, gb_ref, d, d, p, p)
    assert rel_err(gx, gx_ref) <= TOL[mode]
    assert rel_err(m.gradWeight.numpy(), gw_ref) <= TOL[mode]
    assert rel_err(m.gradBias.numpy(), gb_ref) <= TOL[mode]


@pytest.mark.parametrize("shape", [(8, 64, 32, 32), (16, 200, 1, 1), (3, 5, 7, 9), (4, 512, 4, 4)])
@pytest.mark.parametrize("train", [True, False])
def test_batchnorm(cenn, shape, train):
    from video_filler_b200 import nn
    Cn = shape[1]
    rng = np.random.default_rng(11)
    x = (rng.normal(0.3, 1.5, shape)).astype(np.float32)
    g, be = rng.normal(1, 0.02, Cn).astype(np.float32), rng.normal(0, 0.1, Cn).astype(np.float32)
    rm, rv = rng.normal(0, 0.1, Cn).astype(np.float32), rng.uniform(0.5, 2, Cn).astype(np.float32)
    m = nn.SpatialBatchNormalization(Cn)
    m.weight.copy_(g); m.bias.copy_(be); m.running_mean.copy_(rm); m.running_var.copy_(rv)
    if not train:
        m.evaluate()
    y = m.forward(dev(cenn, x)).numpy()
    rm_ref, rv_ref = rm.astype(np.float64), rv.astype(np.float64)
    y_ref, sm, si = ops.bn_forward(x.astype(np.float64), g.astype(np.float64), be.astype(np.float64), rm_ref, rv_ref, train)
    assert rel_err(y, y_ref) <= 1e-5
    assert rel_err(m.running_mean.numpy(), rm_ref) <= 1e-5 and rel_err(m.running_var.numpy(), rv_ref) <= 1e-5
    gy = rng.normal(0, 1, shape).astype(np.float32)
    m.gradWeight.fill(0.25); m.gradBias.fill(-0.5)
    gx = m.backward(dev(cenn, x), dev(cenn, gy), 2.0).numpy()
    gg, gb = np.full(Cn, 0.25), np.full(Cn, -0.5)
    gx_ref = ops.bn_backward(x.astype(np.float64), gy.astype(np.float64), g.astype(np.float64), sm, si, rm_ref, rv_ref, train,
                             ggamma=gg, gbeta=gb, scale=2.0)
    assert rel_err(gx, gx_ref) <= 2e-5
    assert rel_err(m.gradWeight.numpy(), gg) <= 2e-5 and rel_err(m.gradBias.numpy(), gb) <= 2e-5
    # updateGradInput alone must not touch the parameter gradients (netD:updateGradInput, train.lua:373)
    before = m.gradWeight.numpy().copy()
    gx2 = m.updateGradInput(dev(cenn, x), dev(cenn, gy)).numpy()
    assert np.array_equal(before, m.gradWeight.numpy()) and rel_err(gx2, gx_ref) <= 2e-5


def test_activations(cenn):
    from video_filler_b200 import nn
    rng = np.random.default_rng(3)
    x = rng.normal(0, 1, (3, 5, 7, 3)).astype(np.float32)   # odd length -> exercises the unaligned tail
    x.flat[::17] = 0.0
    gy = rng.normal(0, 1, x.shape).astype(np.float32)
    for mod, f, fg in [(nn.LeakyReLU(0.2), lambda v: ops.leaky_relu(v, 0.2), lambda v, y, g: ops.leaky_relu_grad(v, g, 0.2)),
                       (nn.ReLU(), ops.relu, lambda v, y, g: ops.relu_grad(v, g)),
                       (nn.Tanh(), ops.tanh, lambda v, y, g: ops.tanh_grad(y, g)),
                       (nn.Sigmoid(), ops.sigmoid, lambda v, y, g: ops.sigmoid_grad(y, g))]:
        y = mod.forward(dev(cenn, x)).numpy()
        y_ref = f(x.astype(np.float64))
        assert rel_err(y, y_ref) <= 1e-6, type(mod).__name__
        gx = mod.updateGradInput(dev(cenn, x), dev(cenn, gy)).numpy()
        assert rel_err(gx, fg(x.astype(np.float64), y_ref, gy.astype(np.float64))) <= 1e-5, type(mod).__name__
    # in-place flavour overwrites its input and evaluates the gradient mask on the overwritten tensor
    xi = dev(cenn, x)
    m = nn.LeakyReLU(0.2, True)
    out = m.forward(xi)
    assert out is xi and rel_err(xi.numpy(), ops.leaky_relu(x, 0.2)) <= 1e-6
    gi = dev(cenn, gy)
    assert m.updateGradInput(xi, gi) is gi
    assert rel_err(gi.numpy(), ops.leaky_relu_grad(x, gy, 0.2)) <= 1e-6


def test_criteria(cenn):
    from video_filler_b200 import nn
    rng = np.random.default_rng(5)
    p = rng.uniform(0.001, 0.999, (64, 1)).astype(np.float32)
    p[0, 0], p[1, 0] = 1.0, 0.0        # saturated D outputs
    for label in (1.0, 0.0):
        t = np.full(64, label, np.float32)
        c = nn.BCECriterion()
        loss = c.forward(dev(cenn, p), dev(cenn, t))
        assert loss == pytest.approx(ops.bce_forward(p, t), rel=1e-5)
        g = c.backward(dev(cenn, p), dev(cenn, t)).numpy()
        assert rel_err(g, ops.bce_backward(p.astype(np.float64), t.astype(np.float64))) <= 1e-4
    a = rng.normal(0, 1, (4, 3, 16, 16)).astype(np.float32)
    b = rng.normal(0, 1, a.shape).astype(np.float32)
    for crit, f, fb in [(nn.MSECriterion(), ops.mse_forward, ops.mse_backward),
                        (nn.AbsCriterion(), ops.abs_criterion_forward, ops.abs_criterion_backward)]:
        assert crit.forward(dev(cenn, a), dev(cenn, b)) == pytest.approx(f(a, b), rel=1e-5)
        assert rel_err(crit.backward(dev(cenn, a), dev(cenn, b)).numpy(), fb(a.astype(np.float64), b.astype(np.float64))) <= 1e-6
    with pytest.raises(ValueError):
        nn.MSECriterion().forward(dev(cenn, a), dev(cenn, b[:2]))


def test_masked_mse_criterion(cenn):
    from video_filler_b200 import nn
    rng = np.random.default_rng(9)
    x = rng.normal(0, 1, (3, 12, 16, 16)).astype(np.float32)
    t = rng.normal(0, 1, x.shape).astype(np.float32)
    mask = (rng.uniform(size=x.shape) > 0.8).astype(np.uint8)
    with pytest.raises(TypeError):
        nn.MaskedMSECriterion()          # MaskedMSECriterion.lua:15: nil mWeight errors
    c = nn.MaskedMSECriterion(0.05)
    with pytest.raises(AssertionError):
        c.setMask(mask.astype(np.float32))   # :25 asserts a ByteTensor
    c.setMask(mask)
    assert c.forward(dev(cenn, x), dev(cenn, t)) == pytest.approx(ops.masked_mse_forward(x, t, mask, 0.05), rel=1e-5)
    g = c.backward(dev(cenn, x), dev(cenn, t)).numpy()
    assert rel_err(g, ops.masked_mse_backward(x.astype(np.float64), t.astype(np.float64), mask, 0.05)) <= 1e-6


@pytest.mark.parametrize("shape", [(2, 3, 2, 2), (2, 12, 16, 16), (1, 1, 33, 33), (4, 12, 128, 128)])
def test_gdl_criterion(cenn, shape):
    from video_filler_b200 import nn
    rng = np.random.default_rng(13)
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    t = rng.uniform(-1, 1, shape).astype(np.float32)
    c = nn.GDLCriterion(1)
    assert c.forward(dev(cenn, x), dev(cenn, t)) == pytest.approx(ops.gdl_forward(x.astype(np.float64), t.astype(np.float64)), rel=1e-5)
    g = c.backward(dev(cenn, x), dev(cenn, t)).numpy()
    assert rel_err(g, ops.gdl_backward(x.astype(np.float64), t.astype(np.float64))) <= 1e-6


def test_gdl_rejects_non_square_and_alpha(cenn):
    from video_filler_b200 import nn
    from video_filler_b200._lib import CennError
    with pytest.raises(AssertionError):
        nn.GDLCriterion(2)
    x = dev(cenn, np.zeros((1, 1, 4, 6), np.float32))
    with pytest.raises(CennError, match="inconsistent tensor size"):
        nn.GDLCriterion(1).forward(x, x)


def test_blends_composite_adam_and_tensor_math(cenn):
    T = cenn
    api, st = T.api(), T.state()
    rng = np.random.default_rng(17)
    shape = (3, 3, 64, 64)
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    t = rng.uniform(-1, 1, shape).astype(np.float32)
    g = rng.normal(0, 1e-3, shape).astype(np.float32)
    dx, dt = dev(T, x), dev(T, t)     # keep the device tensors alive across the raw-pointer calls
    for wtl2, ov in [(0.999, 4), (0.999, 0), (2.0, 4)]:
        d = dev(T, g)
        loss = C.c_float()
        api.cenn_WeightedMSEBlend_overlap(st, C.c_void_p(d.ptr), C.c_void_p(dx.ptr), C.c_void_p(dt.ptr),
                                          *shape, wtl2, ov, C.byref(loss))
        ref = ops.blend_l2_overlap(g.astype(np.float64), x.astype(np.float64), t.astype(np.float64), wtl2, ov)
        assert rel_err(d.numpy(), ref) <= 1e-5
        assert loss.value == pytest.approx(ops.mse_forward(x, t), rel=1e-5)
    mask = (rng.uniform(size=shape) > 0.85).astype(np.float32)
    for lam, wtgdl in [(0.05, 0.0), (0.0, 0.0), (0.05, 0.5)]:
        d, dm = dev(T, g), dev(T, mask)
        loss = C.c_float()
        api.cenn_WeightedMSEBlend_masked(st, C.c_void_p(d.ptr), C.c_void_p(dx.ptr), C.c_void_p(dt.ptr),
                                         C.c_void_p(dm.ptr), x.size, 0.999, lam, wtgdl, C.byref(loss))
        ref, w = ops.blend_l2_masked(g.astype(np.float64), x.astype(np.float64), t.astype(np.float64),
                                     mask.astype(np.float64), 0.999, lam)
        ref = ref + wtgdl * ops.mse_backward(x.astype(np.float64), t.astype(np.float64))
        assert rel_err(d.numpy(), ref) <= 1e-5
        if lam != 0:
            assert rel_err(dm.numpy(), w) <= 1e-6     # weights written in place over input_mask (:494)
        else:
            assert np.array_equal(dm.numpy(), mask)
    d, dmask = dev(T, x), dev(T, mask)
    api.cenn_MaskComposite(st, C.c_void_p(d.ptr), C.c_void_p(dmask.ptr), C.c_void_p(dt.ptr), x.size)
    assert np.array_equal(d.numpy(), ops.mask_composite(x, mask, t))
    # adam, 3 steps on an odd-length vector
    from video_filler_b200 import optim
    n = 100003
    p = rng.normal(0, 1, n).astype(np.float32)
    pd = dev(T, p)
    st_o, st_d = {}, {"learningRate": 2e-3, "beta1": 0.5}
    p_ref = p.astype(np.float64)
    for _ in range(3):
        gr = rng.normal(0, 1, n).astype(np.float32)
        gd = dev(T, gr)
        optim.adam(lambda xx: (0.0, gd), pd, st_d)
        ops.adam_step(p_ref, gr.astype(np.float64), st_o, 2e-3, 0.5)
    assert rel_err(pd.numpy(), p_ref) <= 1e-6
    # tensor math used by the scripts
    a = dev(T, x)
    a.mul(0.5).add(0.25).add(2.0, dt).cmul(dt)
    assert rel_err(a.numpy(), (x * 0.5 + 0.25 + 2.0 * t) * t) <= 1e-6
    a = dev(T, np.abs(x) + 1).sqrt()
    assert rel_err(a.numpy(), np.sqrt(np.abs(x) + 1)) <= 1e-6
    dg, dden = dev(T, g), dev(T, np.abs(x) + 1)
    a = dev(T, x).addcmul(0.5, dt, dg).addcdiv(-2.0, dt, dden)
    assert rel_err(a.numpy(), x + 0.5 * t * g - 2.0 * t / (np.abs(x) + 1)) <= 1e-6
    a = dev(T, x)
    a.fill_box(1, 2, 8, 56, 8, 56, -0.1843)   # train.lua:289 mean fill of one channel
    ref = x.copy(); ref[:, 1, 8:56, 8:56] = -0.1843
    assert np.array_equal(a.numpy(), ref)
    assert np.array_equal(dev(T, x).crop(16, 16, 32, 32).numpy(), x[:, :, 16:48, 16:48])
    u8 = (rng.uniform(size=shape) > 0.5).astype(np.uint8)
    assert np.array_equal(T.CudaTensor(shape).copy_(u8).numpy(), u8.astype(np.float32))
    r = T.CudaTensor(1 << 20).normal(0.0, 0.02, seed=42).numpy()
    assert abs(r.mean()) < 2e-4 and abs(r.std() - 0.02) < 2e-4
    r = T.CudaTensor(1 << 20).uniform(-1, 1, seed=43).numpy()
    assert r.min() >= -1 and r.max() <= 1 and abs(r.mean()) < 5e-3


def test_errors_do_not_abort(cenn):
    from video_filler_b200 import nn
    from video_filler_b200._lib import CennError
    m = nn.SpatialConvolution(3, 8, 4, 4, 2, 2, 1, 1)
    with pytest.raises(ValueError, match="invalid number of input planes"):
        m.forward(cenn.CudaTensor(2, 5, 8, 8))
    with pytest.raises(CennError, match="output size is too small"):
        nn.SpatialConvolution(3, 8, 4, 4).forward(cenn.CudaTensor(1, 3, 2, 2).zero())
    # the state survives errors
    assert cenn.CudaTensor(4).fill(2).numpy().tolist() == [2, 2, 2, 2]
