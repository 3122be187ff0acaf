"""UMMA smem-descriptor semantics on real hardware: which rows does a K-major SWIZZLE_128B descriptor read when the
start address is `start_row` 128-byte rows into a TMA-written patch, with a given base_offset field and SBO?"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import video_filler_b200.tensor as T
from video_filler_b200 import _lib
lib = _lib.load(); st = T.state(0)
fn = lib.cenn_debug_desc_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]


def bf16_bits(x):
    return (np.asarray(x, np.float32).view(np.uint32) >> 16).astype(np.uint16)


rows = np.repeat(np.arange(160, dtype=np.float32)[:, None], 64, 1)
cols = np.repeat(np.arange(64, dtype=np.float32)[None, :], 160, 0)
for start, bo, sbo in [(0, 0, 1024), (1, 0, 1024), (1, 1, 1024), (3, 0, 1024), (3, 3, 1024), (8, 0, 1024), (9, 1, 1024), (9, 0, 1024),
                       (0, 0, 1152), (1, 0, 1152), (1, 1, 1152), (0, 0, 2048), (1, 0, 2048), (1, 1, 2048), (0, 0, 1280), (2, 0, 1280), (2, 2, 1280)]:
    outs = []
    for src in (rows, cols):
        a = np.ascontiguousarray(bf16_bits(src))
        out = np.zeros((128, 64), np.float32)
        rc = fn(st, a.ctypes.data_as(C.c_void_p), start, bo, sbo, out.ctypes.data_as(C.c_void_p))
        if rc:
            print("FAILED", _lib.last_error()); sys.exit(1)
        outs.append(out)
    r = np.arange(128)
    exp_row = start + (r // 8) * (sbo // 128) + r % 8
    got_row, got_col = outs[0], outs[1]
    ok_row = np.all(got_row == exp_row[:, None])
    ok_col = np.all(got_col == np.arange(64)[None, :])
    print("start_row=%d base_off=%d SBO=%d : rows %s cols %s" % (start, bo, sbo, "OK" if ok_row else "WRONG", "OK" if ok_col else "WRONG"))
    if not (ok_row and ok_col):
        for rr in (0, 1, 7, 8, 9, 17):
            print("    r=%3d expect row %3d: got rows %s cols %s" % (rr, exp_row[rr], got_row[rr, ::8].astype(int).tolist(), got_col[rr, ::8].astype(int).tolist()))
