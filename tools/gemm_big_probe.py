"""Large plain GEMMs through the production 1-CTA gather kernel (cenn_debug_gemm_probe kind 2), for comparison with the CTA-pair probe."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_filler_b200.tensor as T
from video_filler_b200 import _lib
lib = _lib.load(); st = T.state(0)
fn = lib.cenn_debug_gemm_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
os.environ["PROBE_NO_DBG"] = "1"
for M, Nc, K in ((8192, 8192, 4096), (4096, 4096, 4096), (16384, 256, 2048), (65536, 128, 1024)):
    ms = C.c_float(); dbg = (C.c_uint64 * 16)()
    rc = fn(st, 2, M, 1, 1, Nc, K, 0, 0, 10, C.byref(ms), dbg)
    if rc:
        print(M, Nc, K, "FAILED", _lib.last_error()); continue
    d = list(dbg); grid, bn, stg = d[15] >> 32, (d[15] >> 8) & 0xffff, d[15] & 0xff
    print("gemm M %d N %d K %d: %.1f us, %.1f TFLOP/s (grid %d BN %d stages %d)" % (M, Nc, K, ms.value * 1e3, 2.0 * M * Nc * K / ms.value / 1e9, grid, bn, stg))
