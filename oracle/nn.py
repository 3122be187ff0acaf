"""Minimal numpy mirror of the Torch7 ``nn`` module/criterion objects the scripts build.

Same surface as the reference uses: ``forward/backward/updateGradInput/
zeroGradParameters/parameters/getParameters/training/evaluate/apply`` and the fields
``weight bias gradWeight gradBias output gradInput running_mean running_var modules train``.
Test infrastructure only -- see ``oracle/__init__.py``.
"""
import numpy as np

from . import ops


# --------------------------------------------------------------------------------------------
# Optional storage-precision emulation.  When QUANT is a callable (e.g. ops.bf16_round) the oracle rounds
#   * conv / full-conv weights as seen by the three conv phases (the fp32 master copy stays exact for Adam),
#   * the outputs of conv, full-conv and activation modules (BN output is not rounded: the product fuses
#     BN-apply with the activation and stores once),
#   * the gradInput of conv, full-conv (dgrad outputs) and BN modules (its backward output),
# i.e. exactly the tensors the BF16 executor stores in bf16.  Used by tests/test_fused_gpu.py so that ReLU /
# LeakyReLU gates agree between oracle and product and whole-network gradients can be compared tightly.
QUANT = None
_Q_OUT = ("SpatialConvolution", "SpatialFullConvolution", "LeakyReLU", "ReLU", "Tanh")
_Q_GIN = ("SpatialConvolution", "SpatialFullConvolution", "SpatialBatchNormalization")


def _q(x):
    return QUANT(x) if QUANT is not None and x is not None else x


def _qw(w):
    return QUANT(w) if QUANT is not None else w


class Module:
    def __init__(self):
        self.train = True
        self.output = None
        self.gradInput = None
        self.dtype = np.float32

    # -- nn.Module API ---------------------------------------------------
    def forward(self, x):
        return self.updateOutput(x)

    def backward(self, x, gy, scale=1.0):
        self.updateGradInput(x, gy)
        self.accGradParameters(x, gy, scale)
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        pass

    def parameters(self):
        return [], []

    def zeroGradParameters(self):
        for g in self.parameters()[1]:
            g[...] = 0

    def training(self):
        self.apply(lambda m: setattr(m, 'train', True))

    def evaluate(self):
        self.apply(lambda m: setattr(m, 'train', False))

    def apply(self, fn):
        fn(self)

    def type_name(self):
        return 'nn.' + type(self).__name__

    def getParameters(self):
        """Flatten (weight, bias) of every module, in module order, into one vector each for
        params and grads; module fields become views (Module:getParameters, train.lua:262-263)."""
        ps, gs = self.parameters()
        n = sum(p.size for p in ps)
        flat_p = np.empty(n, self.dtype_of(ps))
        flat_g = np.zeros(n, self.dtype_of(ps))
        holders = []
        self._collect_holders(holders)
        off = 0
        for mod, pname, gname in holders:
            p = getattr(mod, pname)
            g = getattr(mod, gname)
            k = p.size
            flat_p[off:off + k] = p.reshape(-1)
            flat_g[off:off + k] = g.reshape(-1)
            setattr(mod, pname, flat_p[off:off + k].reshape(p.shape))
            setattr(mod, gname, flat_g[off:off + k].reshape(p.shape))
            off += k
        return flat_p, flat_g

    @staticmethod
    def dtype_of(ps):
        return ps[0].dtype if ps else np.float32

    def _collect_holders(self, out):
        if getattr(self, 'weight', None) is not None:
            out.append((self, 'weight', 'gradWeight'))
        if getattr(self, 'bias', None) is not None:
            out.append((self, 'bias', 'gradBias'))


class Sequential(Module):
    def __init__(self):
        super().__init__()
        self.modules = []

    def add(self, m):
        self.modules.append(m)
        return self

    def updateOutput(self, x):
        for m in self.modules:
            x = m.updateOutput(x)
            if QUANT is not None and type(m).__name__ in _Q_OUT:
                x = m.output = _q(x)
        self.output = x
        return x

    def _inputs(self, x):
        ins = [x]
        for m in self.modules[:-1]:
            ins.append(m.output)
        return ins

    def updateGradInput(self, x, gy):
        ins = self._inputs(x)
        for m, xi in zip(reversed(self.modules), reversed(ins)):
            gy = m.updateGradInput(xi, gy)
            if QUANT is not None and type(m).__name__ in _Q_GIN:
                gy = m.gradInput = _q(gy)
        self.gradInput = gy
        return gy

    def accGradParameters(self, x, gy, scale=1.0):
        ins = self._inputs(x)
        for m, xi in zip(reversed(self.modules), reversed(ins)):
            m.accGradParameters(xi, gy, scale)
            gy = m.gradInput

    def backward(self, x, gy, scale=1.0):
        ins = self._inputs(x)
        for m, xi in zip(reversed(self.modules), reversed(ins)):
            gy = m.backward(xi, gy, scale)
            if QUANT is not None and type(m).__name__ in _Q_GIN:
                gy = m.gradInput = _q(gy)
        self.gradInput = gy
        return gy

    def parameters(self):
        ps, gs = [], []
        for m in self.modules:
            p, g = m.parameters()
            ps += p
            gs += g
        return ps, gs

    def apply(self, fn):
        fn(self)
        for m in self.modules:
            m.apply(fn)

    def _collect_holders(self, out):
        for m in self.modules:
            m._collect_holders(out)


class ParallelTable(Module):
    """nn.ParallelTable: the i-th member module is applied to the i-th element of the input table
    (train.lua:115-118 noiseGen, :170-174 conditionAdv)."""

    def __init__(self):
        super().__init__()
        self.modules = []

    def add(self, m):
        self.modules.append(m)
        return self

    def updateOutput(self, xs):
        self.output = [m.updateOutput(x) for m, x in zip(self.modules, xs)]
        return self.output

    def updateGradInput(self, xs, gys):
        self.gradInput = [m.updateGradInput(x, gy) for m, x, gy in zip(self.modules, xs, gys)]
        return self.gradInput

    def accGradParameters(self, xs, gys, scale=1.0):
        for m, x, gy in zip(self.modules, xs, gys):
            m.accGradParameters(x, gy, scale)

    def backward(self, xs, gys, scale=1.0):
        self.gradInput = [m.backward(x, gy, scale) for m, x, gy in zip(self.modules, xs, gys)]
        return self.gradInput

    def parameters(self):
        ps, gs = [], []
        for m in self.modules:
            p, g = m.parameters()
            ps += p
            gs += g
        return ps, gs

    def apply(self, fn):
        fn(self)
        for m in self.modules:
            m.apply(fn)

    def _collect_holders(self, out):
        for m in self.modules:
            m._collect_holders(out)


class JoinTable(Module):
    """nn.JoinTable(dimension): concatenates the tensors of the input table along `dimension` (1-based; the scripts use
    JoinTable(2) on batch-mode 4-D tensors = the channel axis, train.lua:120,177); gradInput = the matching narrows."""

    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def updateOutput(self, xs):
        self.output = np.concatenate(xs, axis=self.dimension - 1)
        return self.output

    def updateGradInput(self, xs, gy):
        sizes = [x.shape[self.dimension - 1] for x in xs]
        self.gradInput = [g.copy() for g in np.split(gy, np.cumsum(sizes)[:-1], axis=self.dimension - 1)]
        return self.gradInput


class SpatialConvolution(Module):
    def __init__(self, nIn, nOut, kW, kH, dW=1, dH=1, padW=0, padH=None, dtype=np.float32):
        super().__init__()
        self.nInputPlane, self.nOutputPlane = nIn, nOut
        self.kW, self.kH, self.dW, self.dH = kW, kH, dW, dH
        self.padW = padW
        self.padH = padW if padH is None else padH
        self.weight = np.zeros((nOut, nIn, kH, kW), dtype)
        self.bias = np.zeros(nOut, dtype)
        self.gradWeight = np.zeros_like(self.weight)
        self.gradBias = np.zeros_like(self.bias)

    def updateOutput(self, x):
        self.output = ops.conv_forward(x, _qw(self.weight), self.bias, self.dH, self.dW, self.padH, self.padW)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = ops.conv_grad_input(x.shape, gy, _qw(self.weight), self.dH, self.dW, self.padH, self.padW)
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        ops.conv_acc_grad(x, gy, self.gradWeight, self.gradBias, self.dH, self.dW, self.padH, self.padW, scale)

    def parameters(self):
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]


class SpatialFullConvolution(Module):
    def __init__(self, nIn, nOut, kW, kH, dW=1, dH=1, padW=0, padH=None, adjW=0, adjH=0, dtype=np.float32):
        super().__init__()
        self.nInputPlane, self.nOutputPlane = nIn, nOut
        self.kW, self.kH, self.dW, self.dH = kW, kH, dW, dH
        self.padW = padW
        self.padH = padW if padH is None else padH
        self.adjW, self.adjH = adjW, adjH
        self.weight = np.zeros((nIn, nOut, kH, kW), dtype)
        self.bias = np.zeros(nOut, dtype)
        self.gradWeight = np.zeros_like(self.weight)
        self.gradBias = np.zeros_like(self.bias)

    def updateOutput(self, x):
        self.output = ops.fullconv_forward(x, _qw(self.weight), self.bias, self.dH, self.dW, self.padH, self.padW,
                                           self.adjH, self.adjW)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = ops.fullconv_grad_input(gy, _qw(self.weight), self.dH, self.dW, self.padH, self.padW)
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        ops.fullconv_acc_grad(x, gy, self.gradWeight, self.gradBias, self.dH, self.dW, self.padH, self.padW, scale)

    def parameters(self):
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]


class SpatialBatchNormalization(Module):
    def __init__(self, C, eps=1e-5, momentum=0.1, affine=True, dtype=np.float32):
        super().__init__()
        self.eps, self.momentum, self.affine = eps, momentum, affine
        self.weight = np.ones(C, dtype) if affine else None
        self.bias = np.zeros(C, dtype) if affine else None
        self.gradWeight = np.zeros(C, dtype) if affine else None
        self.gradBias = np.zeros(C, dtype) if affine else None
        self.running_mean = np.zeros(C, dtype)
        self.running_var = np.ones(C, dtype)
        self.save_mean = None
        self.save_std = None

    def updateOutput(self, x):
        self.output, self.save_mean, self.save_std = ops.bn_forward(
            x, self.weight, self.bias, self.running_mean, self.running_var, self.train, self.momentum, self.eps)
        return self.output

    def _bwd(self, x, gy, want_gx, gg, gb, scale):
        return ops.bn_backward(x, gy, self.weight, self.save_mean, self.save_std, self.running_mean,
                               self.running_var, self.train, self.eps, gg, gb, scale, want_gx)

    def updateGradInput(self, x, gy):
        self.gradInput = self._bwd(x, gy, True, None, None, 1.0)
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        self._bwd(x, gy, False, self.gradWeight, self.gradBias, scale)

    def parameters(self):
        if not self.affine:
            return [], []
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]


class LeakyReLU(Module):
    def __init__(self, negval=0.2, inplace=False):
        super().__init__()
        self.negval, self.inplace = negval, inplace

    def updateOutput(self, x):
        self.output = ops.leaky_relu(x, self.negval)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = ops.leaky_relu_grad(x, gy, self.negval)
        return self.gradInput


class ReLU(Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def updateOutput(self, x):
        self.output = ops.relu(x)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = ops.relu_grad(x, gy)
        return self.gradInput


class Tanh(Module):
    def updateOutput(self, x):
        self.output = ops.tanh(x)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = ops.tanh_grad(self.output, gy)
        return self.gradInput


class Sigmoid(Module):
    def updateOutput(self, x):
        self.output = ops.sigmoid(x)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = ops.sigmoid_grad(self.output, gy)
        return self.gradInput


class View(Module):
    """nn.View(1):setNumInputDims(3): [B,1,h,w] -> [B*h*w, 1] (train.lua:199)."""

    def __init__(self, size):
        super().__init__()
        self.size = size
        self.numInputDims = None

    def setNumInputDims(self, n):
        self.numInputDims = n
        return self

    def updateOutput(self, x):
        self.output = x.reshape(-1, self.size)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = gy.reshape(x.shape)
        return self.gradInput


# ---------------------------------------------------------------------------
class Criterion:
    def __init__(self):
        self.output = 0.0
        self.gradInput = None

    def forward(self, x, t):
        return self.updateOutput(x, t)

    def backward(self, x, t):
        return self.updateGradInput(x, t)


class BCECriterion(Criterion):
    def updateOutput(self, x, t):
        self.output = ops.bce_forward(x, t)
        return self.output

    def updateGradInput(self, x, t):
        self.gradInput = ops.bce_backward(x, t)
        return self.gradInput


class MSECriterion(Criterion):
    def updateOutput(self, x, t):
        self.output = ops.mse_forward(x, t)
        return self.output

    def updateGradInput(self, x, t):
        self.gradInput = ops.mse_backward(x, t)
        return self.gradInput


class AbsCriterion(Criterion):
    def updateOutput(self, x, t):
        self.output = ops.abs_criterion_forward(x, t)
        return self.output

    def updateGradInput(self, x, t):
        self.gradInput = ops.abs_criterion_backward(x, t)
        return self.gradInput


class MaskedMSECriterion(Criterion):
    """MaskedMSECriterion.lua:4-42.  ``mWeight`` is mandatory (nil errors in the reference)."""

    def __init__(self, mWeight):
        super().__init__()
        if mWeight is None:
            raise TypeError("attempt to perform arithmetic on a nil value (mWeight)")
        self.mWeight = mWeight
        self.mask = None

    def setMask(self, m):
        assert m.dtype == np.uint8, "setMask expects a ByteTensor"
        self.mask = m.astype(np.float64)

    def updateOutput(self, x, t):
        self.output = ops.masked_mse_forward(x, t, self.mask, self.mWeight)
        return self.output

    def updateGradInput(self, x, t):
        self.gradInput = ops.masked_mse_backward(x, t, self.mask, self.mWeight)
        return self.gradInput


class GDLCriterion(Criterion):
    """gdl_criterion.lua:4-53 (alpha must be 1)."""

    def __init__(self, alpha=1):
        super().__init__()
        assert alpha == 1
        self.alpha = alpha

    def updateOutput(self, x, t):
        self.output = ops.gdl_forward(x, t)
        return self.output

    def updateGradInput(self, x, t):
        self.gradInput = ops.gdl_backward(x, t)
        return self.gradInput
