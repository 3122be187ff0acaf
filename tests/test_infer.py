"""Whole-frame inference sweep (test_vid_wholeim.lua:98-226): host logic on CPU against a literal tile-by-tile
restatement of the Lua loop, and on the GPU through the executor's eval-mode generator against the oracle network."""
import numpy as np
import pytest


def _lua_loop(forward, frames01, mask_hw, maskValue, ncimage, inputLen, F):
    """Tile by tile, 1-based like the script: the oracle for infer.inpaint_wholeim."""
    P, _, inh, inw = frames01.shape
    outh, outw = -(-inh // F) * F, -(-inw // F) * F
    im = frames01.astype(np.float32).copy()
    im[:, :, mask_hw] = maskValue
    images = np.zeros((P, ncimage, outh, outw), np.float32)
    images[:, :, :inh, :inw] = im
    full = (images * 2 - 1).reshape(P * ncimage, outh, outw)
    ncinput, nc = ncimage * inputLen, ncimage * P
    B = P // inputLen
    outImages = np.zeros((P, ncimage, outh, outw), np.float32)
    for h in range(1, outh + 1, F):
        for w in range(1, outw + 1, F):
            flip = h == 1 and w in (1, F + 1, 2 * F + 1)
            patch = np.zeros((B, ncinput, F, F), np.float32)
            for fr in range(1, nc + 1, ncinput):
                idx = (fr + ncinput - 1) // ncinput
                t = full[fr - 1:fr + ncinput - 1, h - 1:h + F - 1, w - 1:w + F - 1]
                patch[idx - 1] = t[:, ::-1, :] if flip else t
            out = forward(patch)
            if flip:
                out = out[:, :, ::-1, :]
            outImages[:, :, h - 1:h + F - 1, w - 1:w + F - 1] = out.reshape(P, ncimage, F, F)
    pad = np.zeros((ncimage, outh, outw), bool)
    pad[:, :inh, :inw] = mask_hw[None]
    inpaint = np.where(pad[None], outImages, full.reshape(P, ncimage, outh, outw))
    return (outImages + 1) / 2, (inpaint + 1) / 2


def _case(rng, P=4, inh=150, inw=200):
    frames = rng.uniform(0, 1, (P, 3, inh, inw)).astype(np.float32)
    mask = np.zeros((inh, inw), bool)
    mask[20:60, 130:190] = True
    mask[100:140, 10:50] = True
    return frames, mask


@pytest.mark.parametrize("inputLen", [1, 2])
def test_sweep_matches_tile_by_tile_loop(inputLen):
    from video_filler_b200 import infer
    rng = np.random.default_rng(0)
    frames, mask = _case(rng)
    w = rng.normal(0, 0.3, (3 * inputLen, 3 * inputLen)).astype(np.float32)

    def fwd(x):   # any per-sample map stands in for the eval-mode generator; not flip-equivariant on purpose
        y = np.tanh(np.einsum("oc,nchw->nohw", w, x))
        return y * np.linspace(0.5, 1.0, x.shape[2], dtype=np.float32)[None, None, :, None]

    out, full, inp = infer.inpaint_wholeim(fwd, frames, mask, 110 / 255.0, ncimage=3, inputLen=inputLen, max_batch=3)
    ref_out, ref_inp = _lua_loop(fwd, frames, mask, 110 / 255.0, 3, inputLen, 128)
    assert out.shape == (4, 3, 256, 256)
    np.testing.assert_allclose(out, ref_out, rtol=0, atol=1e-6)
    np.testing.assert_allclose(inp, ref_inp, rtol=0, atol=1e-6)
    # outside the mask the composite is the (masked, padded) input itself
    assert np.array_equal(inp[:, :, :150, :200][:, :, ~mask], full[:, :, :150, :200][:, :, ~mask])


@pytest.mark.gpu
def test_wholeim_sweep_on_executor_matches_oracle(cenn):
    from conftest import rel_err
    from oracle import nets as onets
    from oracle import step as ostep
    from video_filler_b200 import infer, models, train
    kw = dict(batchSize=8, nBottleneck=128, nef=64, ngf=64, ndf=64, predLen=1)
    orc = ostep.StepOracle(onets.default_opt("video", **kw), seed=3, dtype=np.float64)
    trn = train.FusedTrainer(models.default_opt("video", **kw), precision="bf16")
    trn.set_params(0, orc.pG)
    orc.netG.evaluate()
    rng = np.random.default_rng(1)
    frames, mask = _case(rng, P=2)
    out, full, inp = infer.inpaint_wholeim(trn.generator_forward, frames, mask, 110 / 255.0, ncimage=3, inputLen=1, max_batch=8)
    ref_out, ref_inp = _lua_loop(lambda x: orc.netG.forward(x.astype(np.float64)).astype(np.float32), frames, mask, 110 / 255.0, 3, 1, 128)
    assert rel_err(out, ref_out) <= 2e-2 and rel_err(inp, ref_inp) <= 2e-2


def _oracle_generator(variant, seed, **kw):
    """Oracle generator in eval mode with non-trivial biases and running statistics (what a trained checkpoint holds)."""
    from oracle import nets as onets
    from oracle import step as ostep
    orc = ostep.StepOracle(onets.default_opt(variant, **kw), seed=seed, dtype=np.float64)
    rng = np.random.default_rng(seed + 1)
    orc.pG += rng.normal(0, 0.01, orc.pG.size)            # flat storage: conv biases and BN shifts become non-zero too
    mods = list(orc.netG.modules[0].modules) + list(orc.netG.modules[1:])
    stats = []
    for m in mods:
        if hasattr(m, "running_mean"):
            m.running_mean[:] = rng.normal(0, 0.1, m.running_mean.shape)
            m.running_var[:] = rng.uniform(0.5, 1.5, m.running_var.shape)
            stats += [m.running_mean.copy(), m.running_var.copy()]
    orc.netG.evaluate()
    return orc, np.concatenate(stats)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["video", "image"])
def test_inpainter_forward_matches_oracle_eval(cenn, variant):
    """cenn_inpainter_forward_*: BN folded into the operands == eval-mode netG:forward (test.lua:92, test_vid_wholeim.lua:180)."""
    from conftest import rel_err
    from video_filler_b200 import infer, models
    kw = dict(batchSize=4, nBottleneck=128, nef=64, ngf=64, ndf=64)
    if variant == "video":
        kw["predLen"] = 2
    orc, stats = _oracle_generator(variant, 11, **kw)
    eng = infer.Inpainter(models.default_opt(variant, **kw), batch=4)
    assert eng.counts() == (orc.pG.size, stats.size)
    eng.load(orc.pG, stats)
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (6, eng.ncin, 128, 128)).astype(np.float32)       # 6 tiles through a 4-tile engine: one full + one ragged forward
    y = eng.forward(x)
    y_ref = orc.netG.forward(x.astype(np.float64))
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) <= 2e-2
    y1 = eng.forward(x[:1])                                                   # tiles are independent in eval mode
    assert rel_err(y1, y[:1]) <= 1e-6
    with pytest.raises(Exception, match="tiles outside"):
        eng._api.cenn_inpainter_forward_host(eng.h, x.ctypes.data, y.ctypes.data, 5)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("inputLen,with_init", [(1, False), (2, False), (1, True)])
def test_device_sweep_matches_host_sweep_and_oracle(cenn, inputLen, with_init):
    """cenn_inpainter_sweep_host (everything on the device) vs the numpy sweep around the same engine (identical network
    outputs -> identical images) and vs the literal Lua loop around the fp64 oracle generator."""
    from conftest import rel_err
    from video_filler_b200 import infer, models
    kw = dict(batchSize=5, nBottleneck=128, nef=64, ngf=64, ndf=64, predLen=inputLen)
    orc, stats = _oracle_generator("video", 21, **kw)
    eng = infer.Inpainter(models.default_opt("video", **kw), batch=5)         # 4 tiles x (4/inputLen) groups: ragged last chunk
    eng.load(orc.pG, stats)
    init = None
    if with_init:
        orc_i, stats_i = _oracle_generator("video", 31, **kw)
        init = infer.Inpainter(models.default_opt("video", **kw), batch=5)
        init.load(orc_i.pG, stats_i)
    rng = np.random.default_rng(1)
    frames, mask = _case(rng, P=4)
    mv = 110 / 255.0
    out, full, inp = eng.sweep(frames, mask, mv, init=init)
    h_out, h_full, h_inp = infer.inpaint_wholeim(eng.forward, frames, mask, mv, ncimage=3, inputLen=inputLen, max_batch=5,
                                                 forward_init=init.forward if init else None)
    assert out.shape == (4, 3, 256, 256)
    np.testing.assert_allclose(full, h_full, rtol=0, atol=1e-6)
    # same engine, same tiles: the device sweep differs from the host sweep only by the bf16 rounding of the tile inputs
    # it writes directly (the host path rounds the same values inside forward) -> identical up to fp32 rescaling
    np.testing.assert_allclose(out, h_out, rtol=0, atol=1e-6)
    np.testing.assert_allclose(inp, h_inp, rtol=0, atol=1e-6)
    if not with_init:
        ref_out, ref_inp = _lua_loop(lambda x: orc.netG.forward(x.astype(np.float64)).astype(np.float32), frames, mask, mv, 3, inputLen, 128)
        assert rel_err(out, ref_out) <= 2e-2 and rel_err(inp, ref_inp) <= 2e-2
    assert np.array_equal(inp[:, :, :150, :200][:, :, ~mask], full[:, :, :150, :200][:, :, ~mask])
    if inputLen == 2:
        with pytest.raises(Exception, match="padding in time"):      # test_vid_wholeim.lua:41
            eng.sweep(frames[:3], mask, mv)
    eng.close()
    if init:
        init.close()
