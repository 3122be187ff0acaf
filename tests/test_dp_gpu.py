"""Data parallelism on ONE GPU: two half-batch executors (world_size 2, ranks 0/1) driven phase by phase, their sync
buffers summed by hand (what the library's NCCL all-reduce does on a multi-GPU box), against one executor at the global
batch.  Checks that every cross-sample quantity of the step (BN batch statistics, BN backward sums, gradients, losses)
goes through a sync point, i.e. that sharding the batch does not change the step."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _sum_buffers(T, ptrs, count, is_double):
    dt, nb = (np.float64, 8) if is_double else (np.float32, 4)
    acc = None
    bufs = []
    for p in ptrs:
        a = np.empty(count, dt)
        T.api().cenn_copy_d2h(T.state(), a.ctypes.data_as(C.c_void_p), C.c_void_p(p), count * nb)
        bufs.append(a)
        acc = a.astype(np.float64) if acc is None else acc + a
    out = acc.astype(dt)
    for p in ptrs:
        T.api().cenn_copy_h2d(T.state(), C.c_void_p(p), out.ctypes.data_as(C.c_void_p), count * nb)


@pytest.mark.parametrize("variant", ["image", "video"])
def test_two_rank_step_equals_global_batch_step(cenn, variant):
    import video_filler_b200.tensor as T
    from video_filler_b200 import models, synth, train, util
    kw = dict(batchSize=8, nBottleneck=128, nef=64, ngf=64, ndf=64)
    if variant == "video":
        kw["predLen"] = 2
    opt_full = models.default_opt(variant, **kw)
    opt_half = models.default_opt(variant, **dict(kw, batchSize=4))
    rng = np.random.default_rng(7)
    pG = util.params_flat(util.weights_init(util.describe_netG(opt_full), rng))
    pD = util.params_flat(util.weights_init(util.describe_netD(opt_full), rng))
    full = train.FusedTrainer(opt_full, precision="bf16")
    halves = [train.FusedTrainer(opt_half, precision="bf16", world_size=2, rank=r) for r in (0, 1)]
    for t in [full] + halves:
        t.set_params(0, pG); t.set_params(1, pD)
    drng = np.random.default_rng(11)
    if variant == "image":
        batch = synth.image_batch(8, 128, 4, drng)
    else:
        batch = synth.video_batch(8, 6, 128, 0.43, drng)
    losses_full = full.step_host(*batch)
    full2 = train.FusedTrainer(opt_full, precision="bf16")
    full2.set_params(0, pG); full2.set_params(1, pD)
    full2.step_host(*batch)
    print("run-to-run (same executor config twice): rel G", rel_err(full2.get_grads(0), full.get_grads(0)), "rel D", rel_err(full2.get_grads(1), full.get_grads(1)))
    dev = []
    for r in (0, 1):
        sl = slice(4 * r, 4 * r + 4)
        a, b = T.CudaTensor.from_numpy(np.ascontiguousarray(batch[0][sl])), T.CudaTensor.from_numpy(np.ascontiguousarray(batch[1][sl]))
        m = None
        if variant == "video":
            mh = np.ascontiguousarray(batch[2][sl]).astype(np.uint8)
            mp = C.c_void_p()
            T.api().cenn_malloc(T.state(), mh.nbytes, C.byref(mp))
            T.api().cenn_copy_h2d(T.state(), mp, mh.ctypes.data_as(C.c_void_p), mh.nbytes)
            m = mp.value
        dev.append((a, b, m))
    for r in (0, 1):
        halves[r].step_phase(-1, dev[r][0].ptr, dev[r][1].ptr, dev[r][2])
    nsync = 0
    while True:
        infos = [h.sync_info() for h in halves]
        assert infos[0][1:] == infos[1][1:]
        if infos[0][3]:
            break
        if infos[0][0] and infos[0][1]:
            _sum_buffers(T, [i[0] for i in infos], infos[0][1], infos[0][2])
            nsync += 1
        for r in (0, 1):
            halves[r].step_phase(0, dev[r][0].ptr, dev[r][1].ptr, dev[r][2])
    assert nsync >= 30                                    # BN forward + backward sweeps, two gradient vectors, the losses
    for r in (0, 1):
        lr = halves[r].read_losses()
        print(r, lr, losses_full)
        # D on the real half-batches sees exactly the global-batch statistics: tight.  Everything downstream of the
        # generator passes the 8-sample bottleneck BatchNorm (8 values per channel), which amplifies the bf16-level
        # differences of the summation order: looser.
        assert lr["errD_real"] == pytest.approx(losses_full["errD_real"], rel=1e-3), r
        for k in ("errG_l2", "errG_total"):
            assert lr[k] == pytest.approx(losses_full[k], rel=1e-2), (r, k)
        for k in ("errD", "errG"):                            # adversarial terms: run-to-run spread of the same executor is ~1 %
            assert lr[k] == pytest.approx(losses_full[k], rel=4e-2), (r, k)
    # replicas hold the same parameters after the step, equal to the global-batch step (bf16 storage + different
    # summation order: Adam's first update is +-lr per weight, so compare the update direction and the gradients)
    gG = [h.get_grads(0) for h in halves]
    assert np.array_equal(gG[0], gG[1])
    gGf, gDh, gDf = full.get_grads(0), halves[0].get_grads(1), full.get_grads(1)
    cosG = float(np.dot(gG[0], gGf) / (np.linalg.norm(gG[0]) * np.linalg.norm(gGf)))
    cosD = float(np.dot(gDh, gDf) / (np.linalg.norm(gDh) * np.linalg.norm(gDf)))
    print("cos G", cosG, "cos D", cosD, "rel G", rel_err(gG[0], gGf), "rel D", rel_err(gDh, gDf))
    # The executor is not bit-reproducible run to run (fp32 atomics in the statistics / split-K reductions change the
    # summation order, a flipped bf16 rounding then moves LeakyReLU gates); at batch 8 that alone moves gradients by
    # several per cent (printed above), so the sharded step is held to the same bound: same direction, same size.
    assert cosG >= 0.97 and cosD >= 0.97
    assert np.linalg.norm(gG[0]) == pytest.approx(np.linalg.norm(gGf), rel=5e-2)
    assert np.linalg.norm(gDh) == pytest.approx(np.linalg.norm(gDf), rel=5e-2)
    assert np.array_equal(halves[0].get_params(0), halves[1].get_params(0))
    assert rel_err(halves[0].get_bn_stats(0), full.get_bn_stats(0)) <= 5e-3
