"""CudaTensor: a device-storage handle over the C ABI (stand-in for cutorch's torch.CudaTensor).

Only what the reference scripts do to GPU tensors is provided: copy to/from host,
fill/zero/mul/add/cmul/addcmul/addcdiv/sqrt, views into a flat storage, clone.  fp32, contiguous.
"""
import ctypes as C

import numpy as np

from . import _lib

_state = None
_api = None


def api():
    global _api
    if _api is None:
        _api = _lib.Api()
    return _api


def state(device=0):
    """The process-wide cenn_state (``require 'cunn'; cutorch.setDevice(opt.gpu)``, train.lua:249-250)."""
    global _state
    if _state is None:
        out = C.c_void_p()
        api().cenn_init(int(device), C.byref(out))
        _state = out
    return _state


def set_precision(mode):
    api().cenn_set_precision(state(), {"fp32": 0, "bf16": 1}.get(mode, mode))


def synchronize():
    api().cenn_synchronize(state())


class CudaTensor:
    def __init__(self, *shape, ptr=None, base=None):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        self.shape = tuple(int(s) for s in shape)
        self.base = base
        if ptr is None:
            p = C.c_void_p()
            api().cenn_malloc(state(), max(self.nelement(), 1) * 4, C.byref(p))
            self.ptr = p.value
            self._own = True
        else:
            self.ptr = int(ptr)
            self._own = False

    # -- construction -------------------------------------------------------
    @staticmethod
    def from_numpy(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        t = CudaTensor(a.shape)
        t.copy_(a)
        return t

    def __del__(self):
        try:
            if getattr(self, "_own", False) and self.ptr and _state is not None:
                api().cenn_free(_state, C.c_void_p(self.ptr))
        except Exception:
            pass

    # -- shape ----------------------------------------------------------------
    def nelement(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def size(self, d=None):
        return self.shape if d is None else self.shape[d]

    def dim(self):
        return len(self.shape)

    def view(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            k = shape.index(-1)
            known = 1
            for i, s in enumerate(shape):
                if i != k:
                    known *= s
            shape[k] = self.nelement() // known
        t = CudaTensor(shape, ptr=self.ptr, base=self.base or self)
        assert t.nelement() == self.nelement(), "view: size mismatch"
        return t

    def narrow_flat(self, offset, n, shape=None):
        """View of n elements starting at flat offset (how getParameters aliases module fields)."""
        assert 0 <= offset and offset + n <= self.nelement()
        return CudaTensor(shape if shape is not None else (n,), ptr=self.ptr + 4 * offset, base=self.base or self)

    # -- data movement ------------------------------------------------------------
    def copy_(self, src):
        if isinstance(src, CudaTensor):
            assert src.nelement() == self.nelement(), "copy: nElement mismatch"
            api().cenn_copy_d2d(state(), C.c_void_p(self.ptr), C.c_void_p(src.ptr), self.nelement() * 4)
        else:
            a = np.ascontiguousarray(src)
            assert a.size == self.nelement(), "copy: nElement mismatch"
            if a.dtype == np.uint8:
                tmp = C.c_void_p()
                api().cenn_malloc(state(), max(a.size, 1), C.byref(tmp))
                api().cenn_copy_h2d(state(), tmp, a.ctypes.data_as(C.c_void_p), a.size)
                api().cenn_u8_to_float(state(), C.c_void_p(self.ptr), tmp, a.size)
                api().cenn_free(state(), tmp)
            else:
                a = np.ascontiguousarray(a, dtype=np.float32)
                api().cenn_copy_h2d(state(), C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p), a.size * 4)
        return self

    copy = copy_

    def numpy(self):
        out = np.empty(self.shape, np.float32)
        if out.size:
            api().cenn_copy_d2h(state(), out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), out.size * 4)
        return out

    float = numpy

    def clone(self):
        t = CudaTensor(self.shape)
        t.copy_(self)
        return t

    def new(self, *shape):
        return CudaTensor(*shape)

    # -- math (in place, like Torch) --------------------------------------------------
    def fill(self, v):
        api().cenn_fill(state(), C.c_void_p(self.ptr), self.nelement(), float(v))
        return self

    def zero(self):
        return self.fill(0.0)

    def mul(self, a):
        api().cenn_mul(state(), C.c_void_p(self.ptr), self.nelement(), float(a))
        return self

    def add(self, a, x=None):
        """t:add(scalar) or t:add(scalar, tensor) or t:add(tensor)."""
        if x is None and isinstance(a, CudaTensor):
            a, x = 1.0, a
        if x is None:
            api().cenn_add_scalar(state(), C.c_void_p(self.ptr), self.nelement(), float(a))
        else:
            assert x.nelement() == self.nelement()
            api().cenn_axpy(state(), C.c_void_p(self.ptr), C.c_void_p(x.ptr), self.nelement(), float(a))
        return self

    def cmul(self, x):
        assert x.nelement() == self.nelement()
        api().cenn_cmul(state(), C.c_void_p(self.ptr), C.c_void_p(x.ptr), self.nelement())
        return self

    def addcmul(self, a, p, q):
        api().cenn_addcmul(state(), C.c_void_p(self.ptr), float(a), C.c_void_p(p.ptr), C.c_void_p(q.ptr), self.nelement())
        return self

    def addcdiv(self, a, p, q):
        api().cenn_addcdiv(state(), C.c_void_p(self.ptr), float(a), C.c_void_p(p.ptr), C.c_void_p(q.ptr), self.nelement())
        return self

    def sqrt(self):
        api().cenn_sqrt(state(), C.c_void_p(self.ptr), self.nelement())
        return self

    def normal(self, mean=0.0, std=1.0, seed=0):
        api().cenn_normal(state(), C.c_void_p(self.ptr), self.nelement(), float(mean), float(std), int(seed))
        return self

    def uniform(self, a=0.0, b=1.0, seed=0):
        api().cenn_uniform(state(), C.c_void_p(self.ptr), self.nelement(), float(a), float(b), int(seed))
        return self

    def maskedFill(self, mask, v):
        api().cenn_masked_fill(state(), C.c_void_p(self.ptr), C.c_void_p(mask.ptr), self.nelement(), float(v))
        return self

    def fill_box(self, c0, c1, y0, y1, x0, x1, v):
        """t[{{},{c0+1,c1},{y0+1,y1},{x0+1,x1}}] = v with 0-based half-open bounds."""
        N, Cn, H, W = self.shape
        api().cenn_fill_box(state(), C.c_void_p(self.ptr), N, Cn, H, W, c0, c1, y0, y1, x0, x1, float(v))
        return self

    def crop(self, y0, x0, h, w):
        N, Cn, H, W = self.shape
        out = CudaTensor(N, Cn, h, w)
        api().cenn_crop(state(), C.c_void_p(out.ptr), C.c_void_p(self.ptr), N, Cn, H, W, y0, x0, h, w)
        return out
