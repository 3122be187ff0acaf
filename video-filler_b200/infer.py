"""Full-frame video inpainting sweep of ``test_vid_wholeim.lua:98-226`` on the executor's eval-mode generator.

The reference pads the (inh x inw) frames bottom-right to a multiple of fineSize (:109-111,139), walks the 128x128 tiles
one by one (:159-205: gather the tile of every frame group, vertical flip for the first three tiles of the top row
(:167-170,194-200), one ``net:forward`` per tile, optional initializer net + ``inpaint_utils.fillIn`` (:179-190)), writes
the outputs back and composites them into the input under the padded mask (:207-220).  Eval-mode BatchNorm makes the
tiles independent, so here ALL tiles of all frame groups go through the generator as one batch (or as few batches as
the executor's batchSize allows): same arithmetic per tile, one launch sequence instead of one per tile.
"""
import numpy as np

FLIPPED_TOP_TILES = 3          # test_vid_wholeim.lua:167: h == 1 and w in {1, fineSize+1, 2*fineSize+1}


def pad_frames(frames01, mask_hw, maskValue, fineSize=128):
    """frames01 [predLen, nc, inh, inw] in [0,1]; mask_hw [inh, inw] bool (already scaled / thresholded, :53-54).
    Returns fullImages [predLen*nc, outh, outw] in [-1,1] (masked pixels = maskValue, padding = 0 -> -1), :60-72."""
    P, nc, inh, inw = frames01.shape
    outh, outw = -(-inh // fineSize) * fineSize, -(-inw // fineSize) * fineSize
    im = frames01.astype(np.float32).copy()
    im[:, :, mask_hw] = maskValue
    images = np.zeros((P, nc, outh, outw), np.float32)
    images[:, :, :inh, :inw] = im
    return (images * 2 - 1).reshape(P * nc, outh, outw)


def tile_batch(fullImages, nc_total, ncinput, fineSize=128):
    """Gather every (tile, frame group) as one generator input [T*B, ncinput, F, F]; returns (batch, tile list)."""
    _, outh, outw = fullImages.shape
    B = nc_total // ncinput
    tiles, out = [], []
    for h in range(0, outh, fineSize):
        for w in range(0, outw, fineSize):
            flip = h == 0 and w in tuple(i * fineSize for i in range(FLIPPED_TOP_TILES))
            tiles.append((h, w, flip))
            for g in range(B):
                patch = fullImages[g * ncinput:(g + 1) * ncinput, h:h + fineSize, w:w + fineSize]
                out.append(patch[:, ::-1, :] if flip else patch)          # image.vflip (:168)
    return np.ascontiguousarray(np.stack(out), np.float32), tiles


def untile(outputs, tiles, predLen, ncimage, ncinput, outh, outw, fineSize=128):
    """outputs [T*B, ncinput, F, F] -> outImages [predLen, ncimage, outh, outw] (:194-204)."""
    B = predLen * ncimage // ncinput
    outImages = np.zeros((predLen, ncimage, outh, outw), np.float32)
    flat = outImages.reshape(predLen * ncimage, outh, outw)
    for ti, (h, w, flip) in enumerate(tiles):
        for g in range(B):
            o = outputs[ti * B + g]
            flat[g * ncinput:(g + 1) * ncinput, h:h + fineSize, w:w + fineSize] = o[:, ::-1, :] if flip else o
    return outImages


def composite(outImages, fullImages, mask_hw):
    """inpaintImages[i] = where(padmask, outImages[i], fullImages[i]); all three rescaled to [0,1] (:207-224)."""
    P, nc, outh, outw = outImages.shape
    pad = np.zeros((nc, outh, outw), bool)
    pad[:, :mask_hw.shape[0], :mask_hw.shape[1]] = mask_hw[None]
    full = fullImages.reshape(P, nc, outh, outw)
    inpaint = np.where(pad[None], outImages, full)
    return (outImages + 1) * 0.5, (full + 1) * 0.5, (inpaint + 1) * 0.5


def inpaint_wholeim(forward, frames01, mask_hw, maskValue, ncimage=3, inputLen=1, fineSize=128, max_batch=None,
                    forward_init=None):
    """``forward(x [n, ncinput, F, F]) -> [n, ncinput, F, F]`` is the eval-mode generator (e.g.
    ``FusedTrainer.generator_forward``); ``forward_init`` the optional initializer net (withInit, :179-190)."""
    P = frames01.shape[0]
    assert P % inputLen == 0, "I don't do padding in time dim (test_vid_wholeim.lua:41)"
    ncinput, nc_total = ncimage * inputLen, ncimage * P
    full = pad_frames(frames01, mask_hw, maskValue, fineSize)
    _, outh, outw = full.shape
    x, tiles = tile_batch(full, nc_total, ncinput, fineSize)
    if forward_init is not None:
        # fillIn(input, tile mask, netI(input)) (:179-188): the tile's slice of the padded mask, un-flipped like the reference
        padm = np.zeros((outh, outw), bool)
        padm[:mask_hw.shape[0], :mask_hw.shape[1]] = mask_hw
        B = nc_total // ncinput
        mid = _run(forward_init, x, max_batch)
        for ti, (h, w, _) in enumerate(tiles):
            m = padm[h:h + fineSize, w:w + fineSize]
            for g in range(B):
                x[ti * B + g][:, m] = mid[ti * B + g][:, m]
    y = _run(forward, x, max_batch)
    outImages = untile(y, tiles, P, ncimage, ncinput, outh, outw, fineSize)
    return composite(outImages, full, mask_hw)


def _run(forward, x, max_batch):
    if max_batch is None or x.shape[0] <= max_batch:
        return forward(x)
    return np.concatenate([forward(x[i:i + max_batch]) for i in range(0, x.shape[0], max_batch)])
