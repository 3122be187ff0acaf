"""Per-op CUDA-event profile of one fused G+D step (B=256, cfg2): writes a table to gpurun_out/ (copy to profiles/)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import video_filler_b200.tensor as T
from video_filler_b200 import models, train
from video_filler_b200 import synth, util

B = int(os.environ.get("B", "256"))
variant = os.environ.get("VARIANT", "image")
T.state(0)
opt = models.default_opt(variant, batchSize=B)
trn = train.FusedTrainer(opt)
rng = np.random.default_rng(1234)
for idx, describe in ((0, util.describe_netG), (1, util.describe_netD)):
    trn.set_params(idx, util.params_flat(util.weights_init(describe(opt), rng)))
if variant == "image":
    a, b = synth.image_batch(B, 128, 4, rng); m = None
else:
    a, b, m = synth.video_batch(B, 12, 128, 110 / 255.0, rng)
da, db = T.CudaTensor.from_numpy(a), T.CudaTensor.from_numpy(b)
dm = None
if m is not None:
    import ctypes as C
    dm = C.c_void_p(); T.api().cenn_malloc(T.state(), m.size, C.byref(dm)); T.api().cenn_copy_h2d(T.state(), dm, m.ctypes.data_as(C.c_void_p), m.size); dm = dm.value
for _ in range(int(os.environ.get("STEPS", "3"))):
    trn.step_device(da.ptr, db.ptr, dm)
prof = trn.profile_step(da.ptr, db.ptr, dm, repeats=5)
lines = ["%4s %-16s %9s %10s %9s" % ("#", "op", "ms", "GFLOP", "TFLOP/s")]
for i, (n, t, f) in enumerate(zip(prof["names"], prof["ms"], prof["flops"])):
    lines.append("%4d %-16s %9.4f %10.2f %9.1f" % (i, n, t, f / 1e9, f / 1e12 / (t / 1e3) if f > 0 and t > 0 else 0))
lines.append("total %.3f ms ; tensor-core ops %.3f ms, %.1f GFLOP -> %.1f TFLOP/s" % (prof["total_ms"], prof["tc_ms"], prof["tc_flops"] / 1e9, prof["tc_flops"] / 1e12 / (prof["tc_ms"] / 1e3)))
lines.append(json.dumps(prof["by_op"]))
out = os.path.join("gpurun_out", "ops_%s_b%d.txt" % (variant, B))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-2:]))
