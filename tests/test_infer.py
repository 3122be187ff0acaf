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
