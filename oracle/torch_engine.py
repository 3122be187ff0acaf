"""PyTorch-CPU engine for the oracle's heavy ops (TEST INFRASTRUCTURE, NOT PRODUCT).

``enable()`` swaps the convolution / batch-norm / activation / Adam functions of ``oracle.ops`` for PyTorch-CPU
(oneDNN / MKL) implementations with identical signatures and semantics, so that ``oracle.nn`` and ``oracle.step`` --
the module objects and the fDx / fGx step sequence of ``train.lua:278-410`` -- run unchanged on a multi-threaded
engine.  This is the CPU baseline BASELINE.md section 4 plans (the reference's ``gpu=0`` path cannot run: no Torch7):
``F.conv2d`` / ``F.conv_transpose2d`` / ``native_batch_norm`` have the weight layouts and formulas of SURVEY 9.1-9.3
(``[Cout,Cin,kH,kW]`` / ``[Cin,Cout,kH,kW]``, cross-correlation, biased variance for normalisation, unbiased variance
into running_var, double accumulators for float tensors on the CPU).  Checked against the numpy functions it replaces in
``tests/test_oracle_torch_engine.py``.

Everything takes and returns numpy arrays (``torch.from_numpy`` shares memory: no copies besides the results).
"""
import numpy as np

from . import ops

_saved = {}


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))


def conv_forward(x, w, b, dH, dW, pH, pW):
    import torch.nn.functional as F
    return F.conv2d(_t(x), _t(w), None if b is None else _t(b), (dH, dW), (pH, pW)).numpy()


def conv_grad_input(x_shape, gy, w, dH, dW, pH, pW):
    import torch
    return torch.nn.grad.conv2d_input(tuple(x_shape), _t(w), _t(gy), (dH, dW), (pH, pW)).numpy()


def conv_acc_grad(x, gy, gw, gb, dH, dW, pH, pW, scale=1.0):
    import torch
    g = torch.nn.grad.conv2d_weight(_t(x), tuple(gw.shape), _t(gy), (dH, dW), (pH, pW)).numpy()
    gw += (scale * g).astype(gw.dtype)
    if gb is not None:
        gb += (scale * _t(gy).sum(dim=(0, 2, 3)).numpy()).astype(gb.dtype)


def fullconv_forward(x, w, b, dH, dW, pH, pW, adjH=0, adjW=0):
    import torch.nn.functional as F
    return F.conv_transpose2d(_t(x), _t(w), None if b is None else _t(b), (dH, dW), (pH, pW), (adjH, adjW)).numpy()


def fullconv_grad_input(gy, w, dH, dW, pH, pW):
    """dgrad of the transposed conv = ordinary conv of gy with w ([Cin,Cout,k,k] read as [out=Cin, in=Cout])."""
    import torch.nn.functional as F
    return F.conv2d(_t(gy), _t(w), None, (dH, dW), (pH, pW)).numpy()


def fullconv_acc_grad(x, gy, gw, gb, dH, dW, pH, pW, scale=1.0):
    """gradWeight[c,o,u,v] += sum x[n,c,i,j] gy[n,o,i*d-p+u,j*d-p+v]: the weight gradient of conv2d(gy -> x)."""
    import torch
    g = torch.nn.grad.conv2d_weight(_t(gy), tuple(gw.shape), _t(x), (dH, dW), (pH, pW)).numpy()
    gw += (scale * g).astype(gw.dtype)
    if gb is not None:
        gb += (scale * _t(gy).sum(dim=(0, 2, 3)).numpy()).astype(gb.dtype)


def bn_forward(x, gamma, beta, running_mean, running_var, train, momentum=0.1, eps=1e-5):
    import torch
    C = x.shape[1]
    g = _t(gamma) if gamma is not None else torch.ones(C, dtype=_t(x).dtype)
    b = _t(beta) if beta is not None else torch.zeros(C, dtype=_t(x).dtype)
    # running_* are updated in place (shared memory) in training mode
    y, mean, invstd = torch.native_batch_norm(_t(x), g, b, _t(running_mean), _t(running_var), bool(train), momentum, eps)
    if not train:
        mean = _t(running_mean).clone()
        invstd = 1.0 / torch.sqrt(_t(running_var).double() + eps)
    return y.numpy(), mean.to(y.dtype).numpy(), invstd.to(y.dtype).numpy()


def bn_backward(x, gy, gamma, save_mean, save_invstd, running_mean, running_var, train, eps=1e-5,
                ggamma=None, gbeta=None, scale=1.0, want_gx=True):
    import torch
    C = x.shape[1]
    tx = _t(x)
    g = _t(gamma) if gamma is not None else torch.ones(C, dtype=tx.dtype)
    want_p = ggamma is not None or gbeta is not None
    gx, gg, gb = torch.ops.aten.native_batch_norm_backward(
        _t(gy), tx, g, None if running_mean is None else _t(running_mean), None if running_var is None else _t(running_var),
        None if save_mean is None else _t(save_mean),
        None if save_invstd is None else _t(save_invstd), bool(train), eps, [bool(want_gx), want_p, want_p])
    if ggamma is not None:
        ggamma += (scale * gg.numpy()).astype(ggamma.dtype)
    if gbeta is not None:
        gbeta += (scale * gb.numpy()).astype(gbeta.dtype)
    return gx.numpy() if want_gx else None


def leaky_relu(x, negval=0.2):
    import torch.nn.functional as F
    return F.leaky_relu(_t(x), negval).numpy()


def leaky_relu_grad(x_or_y, gy, negval=0.2):
    import torch
    tg = _t(gy)
    return torch.where(_t(x_or_y) > 0, tg, tg * negval).numpy()


def relu(x):
    import torch
    return torch.relu(_t(x)).numpy()


def relu_grad(x_or_y, gy):
    import torch
    tg = _t(gy)
    return torch.where(_t(x_or_y) > 0, tg, torch.zeros((), dtype=tg.dtype)).numpy()


def tanh(x):
    import torch
    return torch.tanh(_t(x)).numpy()


def tanh_grad(y, gy):
    ty = _t(y)
    return (_t(gy) * (1 - ty * ty)).numpy()


def adam_step(x, g, state, lr, beta1, beta2=0.999, eps=1e-8):
    """optim.adam, in place on x / state, multi-threaded."""
    import torch
    if 't' not in state:
        state['t'] = 0
        state['m'] = np.zeros_like(x)
        state['v'] = np.zeros_like(x)
    state['t'] += 1
    t = state['t']
    tx, tg, m, v = _t(x), _t(g), _t(state['m']), _t(state['v'])
    m.mul_(beta1).add_(tg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(tg, tg, value=1 - beta2)
    denom = v.sqrt().add_(eps)
    step = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    tx.addcdiv_(m, denom, value=-step)


_NAMES = ("conv_forward", "conv_grad_input", "conv_acc_grad", "fullconv_forward", "fullconv_grad_input", "fullconv_acc_grad",
          "bn_forward", "bn_backward", "leaky_relu", "leaky_relu_grad", "relu", "relu_grad", "tanh", "tanh_grad", "adam_step")


def enable(threads=None):
    """Route oracle.ops' heavy functions through PyTorch-CPU.  Returns the thread count in use."""
    import torch
    if threads:
        torch.set_num_threads(int(threads))
    if not _saved:
        for n in _NAMES:
            _saved[n] = getattr(ops, n)
            setattr(ops, n, globals()[n])
    return torch.get_num_threads()


def disable():
    for n, f in _saved.items():
        setattr(ops, n, f)
    _saved.clear()


def enabled():
    return bool(_saved)
