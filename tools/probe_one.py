"""Run one gather-GEMM probe case (for ncu captures): python tools/probe_one.py <kind> <N> <h> <w> <Cs> <Cl> <stats> <act> [iters]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_filler_b200.tensor as T
from video_filler_b200 import _lib
lib = _lib.load(); st = T.state(0)
fn = lib.cenn_debug_gemm_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p]
a = [int(x) for x in sys.argv[1:9]]
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 3
ms = C.c_float(); dbg = (C.c_uint64 * 16)()
os.environ.setdefault("PROBE_NO_DBG", "1")
rc = fn(st, *a, iters, C.byref(ms), dbg)
print("rc", rc, "ms", ms.value, _lib.last_error() if rc else "")
