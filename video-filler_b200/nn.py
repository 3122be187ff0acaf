"""nn.* modules and criteria with the Torch7 surface the reference scripts use, over libcenn.

Mirrors (same names, argument order and error behaviour) the modules built at
``train.lua:79-202`` / ``train_vid_weighted.lua:112-239`` and the repo-local criteria
``MaskedMSECriterion.lua`` / ``gdl_criterion.lua``.  Each method forwards to the THNN-shaped
C entry point it replaces (``include/cenn.h``), exactly as the Lua methods forward to
``input.THNN.<Op>_<phase>``.  Module-owned ``output`` / ``gradInput`` buffers are reused across
calls and parameters may be views into a flat storage (``getParameters``).
"""
import ctypes as C

from .tensor import CudaTensor, api, state


def _p(t):
    return C.c_void_p(t.ptr) if t is not None else None


def _buf(cur, shape):
    shape = tuple(int(s) for s in shape)
    if cur is None or cur.shape != shape:
        return CudaTensor(shape)
    return cur


class Module:
    def __init__(self):
        self.train = True
        self.output = None
        self.gradInput = None

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
            g.zero()

    def training(self):
        self.apply(lambda m: setattr(m, "train", True))
        return self

    def evaluate(self):
        self.apply(lambda m: setattr(m, "train", False))
        return self

    def apply(self, fn):
        fn(self)

    def type_name(self):
        return "nn." + type(self).__name__

    def cuda(self):
        return self

    def _holders(self, out):
        if getattr(self, "weight", None) is not None:
            out.append((self, "weight", "gradWeight"))
        if getattr(self, "bias", None) is not None:
            out.append((self, "bias", "gradBias"))

    def getParameters(self):
        """Module:getParameters (train.lua:262-263): one flat storage for params, one for grads;
        weight then bias per module in module order; module fields become views."""
        holders = []
        self._holders(holders)
        n = sum(getattr(m, p).nelement() for m, p, _ in holders)
        flat_p, flat_g = CudaTensor(n), CudaTensor(n)
        flat_g.zero()
        off = 0
        for m, pn, gn in holders:
            p, g = getattr(m, pn), getattr(m, gn)
            k = p.nelement()
            vp = flat_p.narrow_flat(off, k, p.shape)
            vg = flat_g.narrow_flat(off, k, p.shape)
            vp.copy_(p)
            vg.copy_(g)
            setattr(m, pn, vp)
            setattr(m, gn, vg)
            off += k
        return flat_p, flat_g


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
        self.output = x
        return x

    def _inputs(self, x):
        return [x] + [m.output for m in self.modules[:-1]]

    def updateGradInput(self, x, gy):
        for m, xi in zip(reversed(self.modules), reversed(self._inputs(x))):
            gy = m.updateGradInput(xi, gy)
        self.gradInput = gy
        return gy

    def accGradParameters(self, x, gy, scale=1.0):
        for m, xi in zip(reversed(self.modules), reversed(self._inputs(x))):
            m.accGradParameters(xi, gy, scale)
            gy = m.gradInput

    def backward(self, x, gy, scale=1.0):
        for m, xi in zip(reversed(self.modules), reversed(self._inputs(x))):
            gy = m.backward(xi, gy, scale)
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

    def _holders(self, out):
        for m in self.modules:
            m._holders(out)


class ParallelTable(Module):
    """nn.ParallelTable: member i is applied to element i of the input table (train.lua:115-118, 170-174)."""

    def __init__(self):
        super().__init__()
        self.modules = []

    def add(self, m):
        self.modules.append(m)
        return self

    def _check(self, xs):
        if not isinstance(xs, (list, tuple)) or len(xs) != len(self.modules):
            raise ValueError("ParallelTable: table of %d tensors expected" % len(self.modules))

    def updateOutput(self, xs):
        self._check(xs)
        self.output = [m.updateOutput(x) for m, x in zip(self.modules, xs)]
        return self.output

    def updateGradInput(self, xs, gys):
        self._check(xs)
        self.gradInput = [m.updateGradInput(x, gy) for m, x, gy in zip(self.modules, xs, gys)]
        return self.gradInput

    def accGradParameters(self, xs, gys, scale=1.0):
        for m, x, gy in zip(self.modules, xs, gys):
            m.accGradParameters(x, gy, scale)

    def backward(self, xs, gys, scale=1.0):
        self._check(xs)
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

    def _holders(self, out):
        for m in self.modules:
            m._holders(out)


class JoinTable(Module):
    """nn.JoinTable(dimension) for batch-mode tensors; the scripts use dimension 2 = the channel axis (train.lua:120,177)."""

    def __init__(self, dimension):
        super().__init__()
        if dimension != 2:
            raise ValueError("JoinTable: only dimension 2 (channels of batch-mode tensors) is used by the reference scripts")
        self.dimension = dimension
        self._grads = None

    @staticmethod
    def _per_sample(x):
        n = 1
        for d in x.shape[1:]:
            n *= d
        return n

    def updateOutput(self, xs):
        N, rest = xs[0].shape[0], xs[0].shape[2:]
        for x in xs:
            if x.shape[0] != N or x.shape[2:] != rest:
                raise ValueError("JoinTable: inconsistent tensor sizes %s vs %s" % (xs[0].shape, x.shape))
        self.output = _buf(self.output, (N, sum(x.shape[1] for x in xs)) + tuple(rest))
        total, off = self._per_sample(self.output), 0
        for x in xs:
            k = self._per_sample(x)
            api().cenn_JoinTable_updateOutput(state(), _p(self.output), _p(x), N, total, off, k)
            off += k
        return self.output

    def updateGradInput(self, xs, gy):
        if self._grads is None or [g.shape for g in self._grads] != [x.shape for x in xs]:
            self._grads = [CudaTensor(x.shape) for x in xs]
        total, off = self._per_sample(gy), 0
        for x, g in zip(xs, self._grads):
            k = self._per_sample(x)
            api().cenn_JoinTable_updateGradInput(state(), _p(gy), _p(g), x.shape[0], total, off, k)
            off += k
        self.gradInput = self._grads
        return self.gradInput


def _check4(x, planes, what):
    if x.dim() != 4:
        raise ValueError("%s: 4D (batch mode) tensor expected, got %dD" % (what, x.dim()))
    if x.shape[1] != planes:
        raise ValueError("%s: invalid number of input planes: expected %d, got %d" % (what, planes, x.shape[1]))


class SpatialConvolution(Module):
    """nn.SpatialConvolution(nIn, nOut, kW, kH, dW, dH, padW, padH) -> THNN SpatialConvolutionMM_*."""

    def __init__(self, nIn, nOut, kW, kH, dW=1, dH=1, padW=0, padH=None):
        super().__init__()
        self.nInputPlane, self.nOutputPlane = nIn, nOut
        self.kW, self.kH, self.dW, self.dH = kW, kH, dW, dH
        self.padW = padW
        self.padH = padW if padH is None else padH
        self.weight = CudaTensor(nOut, nIn, kH, kW).zero()
        self.bias = CudaTensor(nOut).zero()
        self.gradWeight = CudaTensor(nOut, nIn, kH, kW).zero()
        self.gradBias = CudaTensor(nOut).zero()

    def _geom(self, x):
        _check4(x, self.nInputPlane, "SpatialConvolution")
        N, Cn, H, W = x.shape
        oH = (H + 2 * self.padH - self.kH) // self.dH + 1
        oW = (W + 2 * self.padW - self.kW) // self.dW + 1
        return N, Cn, H, W, oH, oW

    def _args(self, N, Cn, H, W):
        return (N, Cn, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH)

    def updateOutput(self, x):
        N, Cn, H, W, oH, oW = self._geom(x)
        self.output = _buf(self.output, (N, self.nOutputPlane, max(oH, 0), max(oW, 0)))
        api().cenn_SpatialConvolutionMM_updateOutput(state(), _p(x), _p(self.output), _p(self.weight), _p(self.bias),
                                                     *self._args(N, Cn, H, W))
        return self.output

    def updateGradInput(self, x, gy):
        N, Cn, H, W, oH, oW = self._geom(x)
        self.gradInput = _buf(self.gradInput, x.shape)
        api().cenn_SpatialConvolutionMM_updateGradInput(state(), _p(gy), _p(self.gradInput), _p(self.weight),
                                                        *self._args(N, Cn, H, W))
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        N, Cn, H, W, oH, oW = self._geom(x)
        api().cenn_SpatialConvolutionMM_accGradParameters(state(), _p(x), _p(gy), _p(self.gradWeight), _p(self.gradBias),
                                                          *self._args(N, Cn, H, W), float(scale))

    def parameters(self):
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]


class SpatialFullConvolution(Module):
    """nn.SpatialFullConvolution(nIn, nOut, kW, kH, dW, dH, padW, padH, adjW, adjH)."""

    def __init__(self, nIn, nOut, kW, kH, dW=1, dH=1, padW=0, padH=None, adjW=0, adjH=0):
        super().__init__()
        self.nInputPlane, self.nOutputPlane = nIn, nOut
        self.kW, self.kH, self.dW, self.dH = kW, kH, dW, dH
        self.padW = padW
        self.padH = padW if padH is None else padH
        self.adjW, self.adjH = adjW, adjH
        if self.adjW > self.dW - 1 or self.adjH > self.dH - 1:
            raise ValueError("adjW and adjH must be smaller than self.dW - 1 and self.dH - 1 respectively")
        self.weight = CudaTensor(nIn, nOut, kH, kW).zero()
        self.bias = CudaTensor(nOut).zero()
        self.gradWeight = CudaTensor(nIn, nOut, kH, kW).zero()
        self.gradBias = CudaTensor(nOut).zero()

    def _args(self, N, Cn, H, W):
        return (N, Cn, H, W, self.nOutputPlane, self.kW, self.kH, self.dW, self.dH, self.padW, self.padH, self.adjW, self.adjH)

    def updateOutput(self, x):
        _check4(x, self.nInputPlane, "SpatialFullConvolution")
        N, Cn, H, W = x.shape
        oH = (H - 1) * self.dH - 2 * self.padH + self.kH + self.adjH
        oW = (W - 1) * self.dW - 2 * self.padW + self.kW + self.adjW
        self.output = _buf(self.output, (N, self.nOutputPlane, oH, oW))
        api().cenn_SpatialFullConvolution_updateOutput(state(), _p(x), _p(self.output), _p(self.weight), _p(self.bias),
                                                       *self._args(N, Cn, H, W))
        return self.output

    def updateGradInput(self, x, gy):
        _check4(x, self.nInputPlane, "SpatialFullConvolution")
        N, Cn, H, W = x.shape
        self.gradInput = _buf(self.gradInput, x.shape)
        api().cenn_SpatialFullConvolution_updateGradInput(state(), _p(gy), _p(self.gradInput), _p(self.weight),
                                                          *self._args(N, Cn, H, W))
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        N, Cn, H, W = x.shape
        api().cenn_SpatialFullConvolution_accGradParameters(state(), _p(x), _p(gy), _p(self.gradWeight), _p(self.gradBias),
                                                            *self._args(N, Cn, H, W), float(scale))

    def parameters(self):
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]


class SpatialBatchNormalization(Module):
    """nn.SpatialBatchNormalization(C, eps=1e-5, momentum=0.1, affine=true) -> THNN BatchNormalization_*."""

    def __init__(self, Cn, eps=1e-5, momentum=0.1, affine=True):
        super().__init__()
        self.eps, self.momentum, self.affine = eps, momentum, affine
        self.nFeature = Cn
        self.weight = CudaTensor(Cn).fill(1.0) if affine else None
        self.bias = CudaTensor(Cn).zero() if affine else None
        self.gradWeight = CudaTensor(Cn).zero() if affine else None
        self.gradBias = CudaTensor(Cn).zero() if affine else None
        self.running_mean = CudaTensor(Cn).zero()
        self.running_var = CudaTensor(Cn).fill(1.0)
        self.save_mean = CudaTensor(Cn).zero()
        self.save_std = CudaTensor(Cn).zero()

    def _dims(self, x):
        _check4(x, self.nFeature, "SpatialBatchNormalization")
        N, Cn, H, W = x.shape
        return N, Cn, H * W

    def updateOutput(self, x):
        N, Cn, sp = self._dims(x)
        self.output = _buf(self.output, x.shape)
        api().cenn_BatchNormalization_updateOutput(state(), _p(x), _p(self.output), _p(self.weight), _p(self.bias),
                                                   _p(self.running_mean), _p(self.running_var), _p(self.save_mean),
                                                   _p(self.save_std), N, Cn, sp, int(self.train), self.momentum, self.eps)
        return self.output

    def _backward(self, x, gy, gx, gw, gb, scale):
        N, Cn, sp = self._dims(x)
        api().cenn_BatchNormalization_backward(state(), _p(x), _p(gy), _p(gx), _p(gw), _p(gb), _p(self.weight),
                                               _p(self.running_mean), _p(self.running_var), _p(self.save_mean),
                                               _p(self.save_std), N, Cn, sp, int(self.train), float(scale), self.eps)

    def backward(self, x, gy, scale=1.0):
        self.gradInput = _buf(self.gradInput, x.shape)
        self._backward(x, gy, self.gradInput, self.gradWeight, self.gradBias, scale)
        return self.gradInput

    def updateGradInput(self, x, gy):
        self.gradInput = _buf(self.gradInput, x.shape)
        self._backward(x, gy, self.gradInput, None, None, 1.0)
        return self.gradInput

    def accGradParameters(self, x, gy, scale=1.0):
        self._backward(x, gy, None, self.gradWeight, self.gradBias, scale)

    def parameters(self):
        if not self.affine:
            return [], []
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]


class _Pointwise(Module):
    inplace = False

    def _out(self, x):
        if self.inplace:
            self.output = x
        else:
            self.output = _buf(self.output, x.shape)
        return self.output

    def _gin(self, gy):
        if self.inplace:
            self.gradInput = gy
        else:
            self.gradInput = _buf(self.gradInput, gy.shape)
        return self.gradInput


class LeakyReLU(_Pointwise):
    def __init__(self, negval=1.0 / 100, inplace=False):
        super().__init__()
        self.negval, self.inplace = negval, inplace

    def updateOutput(self, x):
        y = self._out(x)
        api().cenn_LeakyReLU_updateOutput(state(), _p(x), _p(y), x.nelement(), self.negval, int(self.inplace))
        return y

    def updateGradInput(self, x, gy):
        gx = self._gin(gy)
        api().cenn_LeakyReLU_updateGradInput(state(), _p(x), _p(gy), _p(gx), x.nelement(), self.negval, int(self.inplace))
        return gx


class ReLU(_Pointwise):
    """nn.ReLU(inplace) = nn.Threshold(0, 0, inplace)."""

    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace
        self.threshold, self.val = 0.0, 0.0

    def updateOutput(self, x):
        y = self._out(x)
        api().cenn_Threshold_updateOutput(state(), _p(x), _p(y), x.nelement(), self.threshold, self.val, int(self.inplace))
        return y

    def updateGradInput(self, x, gy):
        gx = self._gin(gy)
        api().cenn_Threshold_updateGradInput(state(), _p(x), _p(gy), _p(gx), x.nelement(), self.threshold, int(self.inplace))
        return gx


class Tanh(_Pointwise):
    def updateOutput(self, x):
        y = self._out(x)
        api().cenn_Tanh_updateOutput(state(), _p(x), _p(y), x.nelement())
        return y

    def updateGradInput(self, x, gy):
        gx = self._gin(gy)
        api().cenn_Tanh_updateGradInput(state(), _p(gy), _p(gx), _p(self.output), gy.nelement())
        return gx


class Sigmoid(_Pointwise):
    def updateOutput(self, x):
        y = self._out(x)
        api().cenn_Sigmoid_updateOutput(state(), _p(x), _p(y), x.nelement())
        return y

    def updateGradInput(self, x, gy):
        gx = self._gin(gy)
        api().cenn_Sigmoid_updateGradInput(state(), _p(gy), _p(gx), _p(self.output), gy.nelement())
        return gx


class View(Module):
    """nn.View(size):setNumInputDims(n) -- [B,1,h,w] -> [B*h*w, size] (train.lua:199)."""

    def __init__(self, size):
        super().__init__()
        self.size = size
        self.numInputDims = None

    def setNumInputDims(self, n):
        self.numInputDims = n
        return self

    def updateOutput(self, x):
        self.output = x.view(-1, self.size)
        return self.output

    def updateGradInput(self, x, gy):
        self.gradInput = gy.view(x.shape)
        return self.gradInput


# ------------------------------------------------------------------------------ criteria
class Criterion:
    def __init__(self):
        self.output = 0.0
        self.gradInput = None

    def forward(self, x, t):
        return self.updateOutput(x, t)

    def backward(self, x, t):
        return self.updateGradInput(x, t)

    def cuda(self):
        return self


def _same_n(x, t, what):
    if x.nelement() != t.nelement():
        raise ValueError("%s: input and target size mismatch (%d vs %d elements)" % (what, x.nelement(), t.nelement()))


class _SimpleCriterion(Criterion):
    fwd = bwd = None
    sizeAverage = True

    def updateOutput(self, x, t):
        _same_n(x, t, type(self).__name__)
        loss = C.c_float()
        getattr(api(), self.fwd)(state(), _p(x), _p(t), x.nelement(), int(self.sizeAverage), C.byref(loss))
        self.output = loss.value
        return self.output

    def updateGradInput(self, x, t):
        _same_n(x, t, type(self).__name__)
        self.gradInput = _buf(self.gradInput, x.shape)
        getattr(api(), self.bwd)(state(), _p(x), _p(t), _p(self.gradInput), x.nelement(), int(self.sizeAverage))
        return self.gradInput


class BCECriterion(_SimpleCriterion):
    fwd, bwd = "cenn_BCECriterion_updateOutput", "cenn_BCECriterion_updateGradInput"


class MSECriterion(_SimpleCriterion):
    fwd, bwd = "cenn_MSECriterion_updateOutput", "cenn_MSECriterion_updateGradInput"


class AbsCriterion(_SimpleCriterion):
    fwd, bwd = "cenn_AbsCriterion_updateOutput", "cenn_AbsCriterion_updateGradInput"


class MaskedMSECriterion(Criterion):
    """nn.MaskedMSECriterion(mWeight) (MaskedMSECriterion.lua:4-42); fused forward+backward kernel."""

    def __init__(self, mWeight=None):
        super().__init__()
        if mWeight is None:
            # MaskedMSECriterion.lua:15 uses the raw argument: nil -> "attempt to perform arithmetic on a nil value"
            raise TypeError("attempt to perform arithmetic on a nil value (mWeight)")
        self.mWeight = mWeight
        self.mask = None

    def setMask(self, m):
        import numpy as np
        if not (isinstance(m, np.ndarray) and m.dtype == np.uint8):
            raise AssertionError("setMask expects a ByteTensor (uint8 array)")   # MaskedMSECriterion.lua:25
        self.mask = CudaTensor(m.shape).copy_(m)

    def updateOutput(self, x, t):
        _same_n(x, t, "MaskedMSECriterion")
        if self.mask is None or self.mask.nelement() != x.nelement():
            raise ValueError("MaskedMSECriterion: mask not set or of the wrong size")
        loss = C.c_float()
        self.gradInput = _buf(self.gradInput, x.shape)
        api().cenn_MaskedMSECriterion_forward_backward(state(), _p(x), _p(t), _p(self.mask), _p(self.gradInput),
                                                       x.nelement(), float(self.mWeight), C.byref(loss))
        self.output = loss.value
        return self.output

    def updateGradInput(self, x, t):
        self.gradInput = _buf(self.gradInput, x.shape)
        api().cenn_MaskedMSECriterion_forward_backward(state(), _p(x), _p(t), _p(self.mask), _p(self.gradInput),
                                                       x.nelement(), float(self.mWeight), None)
        return self.gradInput


class GDLCriterion(Criterion):
    """nn.GDLCriterion(alpha) (gdl_criterion.lua:4-53); alpha must be 1; fused stencil kernel."""

    def __init__(self, alpha=None):
        super().__init__()
        assert alpha == 1, "assertion failed!"   # gdl_criterion.lua:9 (nil alpha also fails there)
        self.alpha = alpha

    def _dims(self, x, t):
        _same_n(x, t, "GDLCriterion")
        if x.dim() != 4:
            raise ValueError("GDLCriterion: 4D tensor expected")
        return x.shape

    def updateOutput(self, x, t):
        N, Cn, H, W = self._dims(x, t)
        loss = C.c_float()
        api().cenn_GDLCriterion_forward_backward(state(), _p(x), _p(t), None, N, Cn, H, W, C.byref(loss))
        self.output = loss.value
        return self.output

    def updateGradInput(self, x, t):
        N, Cn, H, W = self._dims(x, t)
        self.gradInput = _buf(self.gradInput, x.shape)
        api().cenn_GDLCriterion_forward_backward(state(), _p(x), _p(t), _p(self.gradInput), N, Cn, H, W, None)
        return self.gradInput
