"""optim.adam(opfunc, x, state) over a flat CudaTensor (train.lua:421,424).

The five Torch tensor passes (mul/add, mul/addcmul, copy/sqrt/add, addcdiv) are one fused kernel
(``cenn_AdamFlat``); state keys match optim.adam: t, m, v (denom is not materialised).
"""
import ctypes as C

from .tensor import CudaTensor, api, state as _state


def adam(opfunc, x, config, state=None):
    state = config if state is None else state
    lr = config.get("learningRate", 0.001)
    beta1 = config.get("beta1", 0.9)
    beta2 = config.get("beta2", 0.999)
    eps = config.get("epsilon", 1e-8)
    fx, dfdx = opfunc(x)
    if "t" not in state:
        state["t"] = 0
        state["m"] = CudaTensor(x.shape).zero()
        state["v"] = CudaTensor(x.shape).zero()
    state["t"] += 1
    api().cenn_AdamFlat(_state(), C.c_void_p(x.ptr), C.c_void_p(dfdx.ptr), C.c_void_p(state["m"].ptr),
                        C.c_void_p(state["v"].ptr), x.nelement(), lr, beta1, beta2, eps, state["t"])
    return x, [fx]
