"""One TMA box of the implicit-im2col map (conv_tc.cu: map_thin_gather) against the im2col rows numpy expects, for the fprop (128 px)
and wgrad (64 px) box shapes of the thin first layers.  Usage: python tools/tma_thin_probe.py [Cp]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_filler_b200.tensor as T
from video_filler_b200 import _lib
lib = _lib.load(); st = T.state(0)
fn = lib.cenn_debug_tma_thin_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 11 + [C.c_void_p]


def run(Cp, N, H2, W2, bw, bh, bn, c1, x0, y0, n0):
    rng = np.random.default_rng(0)
    # bordered tensor of distinct small integers (exact in bf16): value = unique id mod 251
    pad = np.zeros((N, H2 + 2, W2 + 2, Cp), np.float32)
    pad[:, 1:-1, 1:-1, :] = rng.integers(1, 250, (N, H2, W2, Cp))
    bits = (pad.view(np.uint32) >> 16).astype(np.uint16)
    nbytes = bw * bh * bn * 128
    out = np.zeros(nbytes, np.uint8)
    rc = fn(st, bits.ctypes.data, N, H2, W2, Cp, bw, bh, bn, c1, x0, y0, n0, out.ctypes.data)
    if rc:
        print("FAILED", _lib.last_error()); return False
    got = out.view(np.uint16).reshape(-1, 64)                     # [pixel row][64 elements], 16-byte chunks XOR-swizzled by (row & 7)
    rows = got.shape[0]
    unsw = np.empty_like(got)
    for r in range(rows):
        for ch in range(8):
            unsw[r, ch * 8:(ch + 1) * 8] = got[r, (ch ^ (r & 7)) * 8:((ch ^ (r & 7)) + 1) * 8]
    vals = (unsw.astype(np.uint32) << 16).view(np.float32)
    exp = np.zeros((rows, 64), np.float32)
    r = 0
    for n in range(bn):
        for y in range(bh):
            for x in range(bw):
                oy, ox, nn = y0 + y, x0 + x, n0 + n
                if Cp == 4:      # row = (u, v, c)
                    win = pad[nn, 2 * oy:2 * oy + 4, 2 * ox:2 * ox + 4, :] if (oy < H2 // 2 and ox < W2 // 2 and nn < N) else np.zeros((4, 4, 4))
                    exp[r] = win.reshape(-1)
                else:            # one window row u = c1: (v, c)
                    win = pad[nn, 2 * oy + c1, 2 * ox:2 * ox + 4, :] if (oy < H2 // 2 and ox < W2 // 2 and nn < N) else np.zeros((4, 16))
                    exp[r] = win.reshape(-1)
                r += 1
    ok = np.array_equal(vals, exp)
    print("Cp %d N %d %dx%d box (%d,%d,%d) at (c1 %d, x %d, y %d, n %d): %s" % (Cp, N, H2, W2, bw, bh, bn, c1, x0, y0, n0, "OK" if ok else "MISMATCH"))
    if not ok:
        bad = np.argwhere(vals != exp)
        print("  first mismatches (row, k):", bad[:6].tolist(), "got", vals[tuple(bad[0])], "expected", exp[tuple(bad[0])])
    return ok


Cp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for args in ((Cp, 4, 64, 64, 32, 4, 1, 0, 0, 4, 1), (Cp, 4, 64, 64, 32, 2, 1, 0, 0, 30, 3), (Cp, 4, 64, 64, 32, 2, 1, 0, 0, 0, 0),
             (Cp, 2, 128, 128, 64, 2, 1, 0, 0, 62, 1), (Cp, 2, 128, 128, 64, 1, 1, 0, 0, 63, 1), (Cp, 3, 8, 8, 4, 4, 4, 0, 0, 0, 0)):
    if Cp == 16:
        for u in (0, 3):
            run(*args[:7], u, *args[8:])
    else:
        run(*args)
