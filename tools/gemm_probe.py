"""Probe the persistent gather-GEMM on representative layer shapes: duration, TFLOP/s and CTA 0's per-role cycle counters."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import video_filler_b200.tensor as T
from video_filler_b200 import _lib
lib = _lib.load(); st = T.state(0)
fn = lib.cenn_debug_gemm_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
B = 256
cases = [  # name, kind, N, h, w, Cs, Cl, stats, act, flops
    ("E1 gemm M=1M N=64 K=64 leaky", 2, B * 64 * 64, 1, 1, 64, 64, 0, 1, 2.0 * B * 64 * 64 * 64 * 64),
    ("E2 fprop 64->64 @32x32 +stats", 0, B, 32, 32, 64, 64, 1, 0, 2.0 * B * 32 * 32 * 64 * 1024),
    ("E3 fprop 64->128 @16x16 +stats", 0, B, 16, 16, 128, 64, 1, 0, 2.0 * B * 16 * 16 * 128 * 1024),
    ("E4 fprop 128->256 @8x8 +stats", 0, B, 8, 8, 256, 128, 1, 0, 2.0 * B * 8 * 8 * 256 * 2048),
    ("E5 fprop 256->512 @4x4 +stats", 0, B, 4, 4, 512, 256, 1, 0, 2.0 * B * 4 * 4 * 512 * 4096),
    ("E6 gemm M=256 N=4000 K=8192", 2, B, 1, 1, 4000, 8192, 1, 0, 2.0 * B * 4000 * 8192),
    ("G2 dgrad-type 512->256 @4x4->8x8", 1, B, 4, 4, 512, 256, 1, 0, 2.0 * B * 4 * 4 * 256 * 16 * 512),
    ("G4 dgrad-type 128->64 @16->32", 1, B, 16, 16, 128, 64, 1, 0, 2.0 * B * 16 * 16 * 64 * 16 * 128),
    ("G5 dgrad-type 64->4 @32->64 tanh", 1, B, 32, 32, 64, 4, 0, 3, 2.0 * B * 32 * 32 * 4 * 16 * 64),
]
variants = [(0, "full"), (1, "no MMA"), (3, "B only, no MMA"), (5, "A only, no MMA"), (6, "MMA only")] if "--variants" in sys.argv else [(0, "full")]
if "--stages" in sys.argv:
    os.environ["PROBE_STAGES"] = sys.argv[sys.argv.index("--stages") + 1]
if "--commit" in sys.argv:
    variants = [(7, "empty loop, commit"), (7 + 32, "empty loop, sw arrive"), (6, "MMA only, commit"), (6 + 32, "MMA only, sw arrive")]
if "--nodbg" in sys.argv:
    os.environ["PROBE_NO_DBG"] = "1"
if "--mma" in sys.argv:
    variants = [(6, "MMA only"), (7, "no MMA no TMA"), (7 + 8, "4 MMA alternating acc"), (7 + 16, "8 MMA same acc"), (7 + 24, "8 MMA alternating acc")]
for name, kind, N, h, w, Cs, Cl, stats, act, flops in [c for c in cases for _ in variants]:
    flag, vname = variants[0]; variants = variants[1:] + variants[:1]
    os.environ["PROBE_FLAGS"] = str(flag)
    name = "%s [%s]" % (name, vname)
    ms = C.c_float(); dbg = (C.c_uint64 * 16)()
    rc = fn(st, kind, N, h, w, Cs, Cl, stats, act, 10, C.byref(ms), dbg)
    if rc:
        print(name, "FAILED", _lib.last_error()); continue
    d = list(dbg)
    grid, bn, stg = d[15] >> 32, (d[15] >> 8) & 0xffff, d[15] & 0xff
    print("%-56s %8.1f us %7.1f TF/s grid=%d BN=%d stages=%d tiles/cta0=%d | prod wait %5.1f%% | mma wait full %5.1f%% tempty %5.1f%% | epi wait %5.1f%% (%d cyc/tile) | mma thread per k-block: wait %d fence %d issue %d commit %d" % (
        name, ms.value * 1e3, flops / ms.value / 1e9, grid, bn, stg, d[7], 100.0 * d[0] / max(d[1], 1), 100.0 * d[2] / max(d[4], 1), 100.0 * d[3] / max(d[4], 1),
        100.0 * d[5] / max(d[6], 1), d[6] // max(d[7], 1), d[2] // max(d[11], 1), d[10] // max(d[11], 1), d[9] // max(d[11], 1), d[8] // max(d[11], 1)))
