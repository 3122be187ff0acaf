"""video-filler_b200: B200-native (sm_100a) context-encoder G+D step behind the Torch7 nn surface.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (``include/cenn.h``) -> ``csrc/libcenn.so``
  _lib.py      ctypes binding generated from ``include/cenn.h`` (stand-in for the LuaJIT FFI cdef)
  tensor.py    CudaTensor: device storage handle (cutorch stand-in)
  nn.py        nn.* modules / criteria with the reference's :forward/:backward surface
  optim.py     optim.adam
  models.py    network builders of train.lua / train_vid_weighted.lua
  train.py     the fDx / fGx closures (op-by-op drop-in path) and the fused whole-step Trainer
  t7.py        Torch7 .t7 reader / writer (util.save / util.load)

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but creating a
state without a B200 and a built ``libcenn.so`` raises.
"""
__version__ = "0.1"
