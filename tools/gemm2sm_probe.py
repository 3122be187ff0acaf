"""Bring-up probe for CTA pairs (tcgen05.mma.cta_group::2): C = A * B^T through cenn_debug_gemm2sm_probe vs numpy.

EXPERIMENTAL: the kernel had only been through ptxas when round 1 ended.  Run it under a hard timeout,
smallest shape first -- a protocol bug traps after ~2 s of mbarrier waiting instead of hanging the GPU:

    timeout 60 python tools/gemm2sm_probe.py 256 256 64          # one cluster, one k-block
    timeout 60 python tools/gemm2sm_probe.py 256 256 512         # pipeline wrap-around (8 k-blocks through 4 stages)
    timeout 60 python tools/gemm2sm_probe.py 8192 8192 4096 20   # throughput: 550 GFLOP per launch
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def bf16_bits(x):
    """float32 array -> (uint16 bf16 bit patterns, the rounded values as float32); round to nearest even."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r, (r.astype(np.uint32) << 16).view(np.float32)


def main():
    M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 256, 64)
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    import video_filler_b200.tensor as T
    from video_filler_b200 import _lib
    T.state(0)
    lib = _lib.load()
    fn = lib.cenn_debug_gemm2sm_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    a_bits, a = bf16_bits(rng.normal(0, 1, (M, K)))
    b_bits, b = bf16_bits(rng.normal(0, 1, (N, K)))
    c = np.zeros((M, N), np.float32)
    ms = C.c_float()
    rc = fn(T.state(), a_bits.ctypes.data, b_bits.ctypes.data, M, N, K, iters, c.ctypes.data, C.byref(ms))
    if rc:
        print("FAILED:", _lib.last_error())
        sys.exit(1)
    check_rows = slice(None) if M * N * K <= (1 << 31) else slice(0, 512)
    ref = a[check_rows].astype(np.float64) @ b.T.astype(np.float64)
    err = float(np.max(np.abs(c[check_rows] - ref)) / np.max(np.abs(ref)))
    print("M %d N %d K %d: max rel err %.3e" % (M, N, K, err), end="")
    if iters:
        print("; %.3f ms per launch = %.1f TFLOP/s" % (ms.value, 2.0 * M * N * K / (ms.value * 1e-3) / 1e12), end="")
    print()
    if err > 1e-5:
        bad = np.argwhere(np.abs(c[check_rows] - ref) > 1e-3 * np.max(np.abs(ref)))
        print("first mismatches (row, col):", bad[:8].tolist(), "rows with errors:", sorted(set(bad[:, 0].tolist()))[:16])
        sys.exit(2)


if __name__ == "__main__":
    main()
