"""GPU parity of the whole G+D step through the drop-in (op-by-op) path against the oracle."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import nets as onets
from oracle import step as ostep

pytestmark = pytest.mark.gpu


def _copy_params(src_oracle, dst_trainer):
    dst_trainer.parametersG.copy_(src_oracle.pG)
    dst_trainer.parametersD.copy_(src_oracle.pD)


@pytest.mark.parametrize("variant,extra", [("image", {}), ("video", {}), ("video", {"wtgdl": 0.5}), ("video", {"weight_nomask": 0.0}),
                                           ("video", {"wtl2": 0.0, "wtgdl": 0.5}),       # the GDL gradient term must survive wtl2 == 0
                                           ("image", {"noiseGen": 1, "nz": 12}), ("image", {"conditionAdv": 1}),
                                           ("image", {"noiseGen": 1, "nz": 8, "conditionAdv": 1})])
def test_closure_step_matches_oracle_fp32(cenn, variant, extra):
    from video_filler_b200 import models, train
    cenn.set_precision("fp32")
    kw = dict(batchSize=4, nBottleneck=96, nef=16, ngf=16, ndf=16, **extra)
    if variant == "video":
        kw["predLen"] = 2
    opt = models.default_opt(variant, **kw)
    oopt = onets.default_opt(variant, **kw)
    orc = ostep.StepOracle(oopt, seed=1234, dtype=np.float64)
    trn = train.ClosureTrainer(opt, seed=99)
    _copy_params(orc, trn)
    rng = np.random.default_rng(4321)
    for it in range(3):
        batch = orc.synth_batch(rng)
        if extra.get("noiseGen"):          # train.lua:319-323 redraws the noise inside fDx; both sides get the same draw
            batch = tuple(batch) + (rng.normal(0, 1, (4, extra["nz"], 1, 1)),)
        lo = orc.step(*batch)
        lg = trn.step(*batch)
        # step 1 is a pure kernel-parity check; later steps inherit Adam's sign-like first updates
        # (m/sqrt(v) = +-1), which amplify fp32-vs-fp64 noise, so they get the north-star 1 % bound
        # (and the adversarial losses of a GAN at batch 4 are chaotic, so after step 1 only the L2 term is bounded)
        tol = 2e-4 if it == 0 else 2e-2
        for k in (("errD", "errG", "errG_l2", "errG_total") if it == 0 else ("errG_l2", "errG_total")):
            if lo[k] is None:                  # wtl2 == 0: no L2 term is computed (train_vid_weighted.lua:485)
                assert lg[k] is None, (it, k)
                continue
            if extra.get("wtl2") == 0.0 and it > 0:
                continue                       # only adversarial terms left: chaotic after the first step
            assert lg[k] == pytest.approx(lo[k], rel=tol), (it, k)
        assert all(np.isfinite(v) for v in lg.values() if v is not None)
        if extra.get("wtgdl") and it == 0:
            assert lg["errG_gdl"] == pytest.approx(lo["errG_gdl"], rel=tol)
        if it == 0:
            # whole-network fp32 gradients vs fp64: BN over 4 samples is ill-conditioned (1/sqrt(var+eps) amplifies
            # rounding), so the bound is looser than the 1e-5 per-layer bound checked in test_ops_gpu.py
            assert rel_err(trn.gradParametersG.numpy(), orc.gG) <= 3e-3
            assert rel_err(trn.gradParametersD.numpy(), orc.gD) <= 3e-3
            assert rel_err(trn.parametersG.numpy(), orc.pG) <= 5e-3   # |update| = lr for every weight
            assert rel_err(trn.parametersD.numpy(), orc.pD) <= 5e-3
    # reference invariants (SURVEY 9.9 i): Adam moves G's conv biases after fGx; D's were re-zeroed by fGx
    first_G = trn.netG.modules[0].modules[0]
    first_G = first_G.modules[0] if extra.get("noiseGen") else first_G          # ParallelTable -> netE -> first conv
    first_D = trn.netD.modules[0].modules[0].modules[0] if extra.get("conditionAdv") else trn.netD.modules[0]
    assert float(np.abs(first_G.bias.numpy()).max()) > 0
    assert float(np.abs(first_D.bias.numpy()).max()) == 0
