"""Two eagerly launched G+D steps (BASELINE.json configs[1]: batch 256) and nothing else: the command ncu wraps for the launch list
and the `--set full` captures of profiles/ (eager so that every kernel appears by name; CENN_NO_GRAPH=1 is set here)."""
import os, sys
os.environ["CENN_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import video_filler_b200.tensor as T
from video_filler_b200 import models, synth, train, util
T.state(0)
B = int(os.environ.get("B", "256"))
opt = models.default_opt("image", batchSize=B)
trn = train.FusedTrainer(opt, precision="bf16")
rng = np.random.default_rng(1234)
trn.set_params(0, util.params_flat(util.weights_init(util.describe_netG(opt), rng)))
trn.set_params(1, util.params_flat(util.weights_init(util.describe_netD(opt), rng)))
a, b = synth.image_batch(B, 128, 4, rng)
da, db = T.CudaTensor.from_numpy(a), T.CudaTensor.from_numpy(b)
for _ in range(int(os.environ.get("STEPS", "2"))):
    trn.step_device(da.ptr, db.ptr, None)
T.synchronize()
print("losses", trn.read_losses())
trn.close()
