"""CPU oracle for the context-encoder G+D step (TEST INFRASTRUCTURE, NOT PRODUCT).

This package is a plain-numpy restatement of the Torch7 operator semantics the
reference scripts rely on and of the step sequence of
``train.lua:278-410`` / ``train_vid_weighted.lua:373-537``.

PARITY UNPINNED: the reference (/root/reference) is pure Lua on top of
un-vendored, un-versioned Torch7 packages (``nn``, ``nngraph``, ``optim``,
``cunn``); it ships no tests, golden vectors or fixtures, and neither LuaJIT nor
Torch7 exists in this environment.  The oracle therefore pins *our restatement*
of upstream torch/nn (late 2016).  It is cross-checked against an independent
engine (PyTorch-CPU autograd, fp64) in ``tests/test_oracle_*.py`` and against
committed golden vectors in ``tests/golden/`` produced by
``tests/golden/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``video_filler_b200``) never does, and fails loudly without its CUDA library.
"""
