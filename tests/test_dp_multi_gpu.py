"""The real multi-GPU data plane (NCCL communicators, peer-mailbox BN exchanges, bf16 gradient buckets, Adam on bucket sums) against
one global-batch executor: tools/dp_parity.py under torchrun, one rank per GPU.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on the
single-GPU box of the round-end run (tests/test_dp_gpu.py covers the decomposition there, tests/test_dp_gloo.py the host protocol)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# ("--branches": train.lua's noiseGen + conditionAdv options, "--bn-local": per-rank BN statistics; flags of tools/dp_parity.py, not
# environment variables: stripped below)
@pytest.mark.parametrize("variant,env", [("image", {}), ("video", {}), ("image", {"CENN_NO_XR": "1"}), ("image", {"CENN_FP32_BUCKETS": "1"}),
                                         ("image", {"CENN_XR_PULL": "1"}), ("image", {"--branches": "1"}), ("image", {"--bn-local": "1"}), ("video", {"--bn-local": "1"}),
                                         # nBottleneck 4000: E6 / G1 are 32.8 M elements each -> the sharded reduce + Adam over peer memory; and the same on NCCL buckets
                                         ("image", {"--nB": "4000", "--per-rank": "4"}), ("image", {"--nB": "4000", "--per-rank": "4", "CENN_NO_SHARD_ADAM": "1"})])
def test_data_parallel_step_equals_global_batch_step(variant, env):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else (4 if variant == "video" else 2)
    port = 29500 + (hash((variant, tuple(sorted(env)))) % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_parity.py"), "--variant", variant]
    env = dict(env)
    for flag in ("--branches", "--bn-local"):
        if env.pop(flag, None):
            cmd.append(flag)
    for flag in ("--nB", "--per-rank"):
        if flag in env:
            cmd += [flag, env.pop(flag)]
    e = dict(os.environ); e.update(env)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DP_PARITY ")]
    assert lines, "no result line\nstdout:\n%s\nstderr:\n%s" % (r.stdout[-3000:], r.stderr[-3000:])
    out = json.loads(lines[-1][len("DP_PARITY "):])
    print(json.dumps(out["checks"]))
    assert r.returncode == 0 and out["ok"], out.get("failures")
    assert out["replicas_bit_identical"]
