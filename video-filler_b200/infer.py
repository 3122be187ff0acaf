"""Full-frame video inpainting sweep of ``test_vid_wholeim.lua:98-226`` on the executor's eval-mode generator.

The reference pads the (inh x inw) frames bottom-right to a multiple of fineSize (:109-111,139), walks the 128x128 tiles
one by one (:159-205: gather the tile of every frame group, vertical flip for the first three tiles of the top row
(:167-170,194-200), one ``net:forward`` per tile, optional initializer net + ``inpaint_utils.fillIn`` (:179-190)), writes
the outputs back and composites them into the input under the padded mask (:207-220).  Eval-mode BatchNorm makes the
tiles independent, so here ALL tiles of all frame groups go through the generator as one batch (or as few batches as
the executor's batchSize allows): same arithmetic per tile, one launch sequence instead of one per tile.
"""
import ctypes as C

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


class Inpainter:
    """Device-side inference engine (cenn_inpainter_*): eval-mode generator with BatchNorm folded into the operands, one
    tensor-core GEMM per layer, CUDA-graph replay; ``sweep`` runs the whole of test_vid_wholeim.lua:98-226 on the device
    (pad, tile gather + flip, forward, write-back, composite).  ``opt`` is the option table of the training script
    (models.default_opt); ``batch`` = tiles per forward."""

    def __init__(self, opt, batch):
        from . import _lib
        from .tensor import api, state
        self.opt, self.batch = opt, int(batch)
        self._api, self._state = api(), state()
        video = opt["variant"] == "video"
        self.cfg = _lib.InpainterConfig(variant=1 if video else 0, batch=self.batch, fineSize=opt["fineSize"],
                                        nBottleneck=opt["nBottleneck"], nef=opt["nef"], ngf=opt["ngf"], nc=opt["nc"],
                                        inputLen=opt.get("predLen", 1) if video else 1)
        self.ncin = opt["nc"] * (opt.get("predLen", 1) if video else 1)
        self.out_size = opt["fineSize"] if video else opt["fineSize"] // 2
        h = C.c_void_p()
        self._api.cenn_inpainter_create(self._state, C.byref(self.cfg), C.byref(h))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._api.cenn_inpainter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counts(self):
        a, b = C.c_int64(), C.c_int64()
        self._api.cenn_inpainter_param_count(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def load(self, flat_G, bn_stats):
        """flat_G: netG:getParameters() vector; bn_stats: [running_mean, running_var] per BN layer in module order."""
        flat_G = np.ascontiguousarray(flat_G, np.float32)
        bn_stats = np.ascontiguousarray(bn_stats, np.float32)
        n, m = self.counts()
        assert flat_G.size == n and bn_stats.size == m, (flat_G.size, n, bn_stats.size, m)
        self._api.cenn_inpainter_load_host(self.h, flat_G.ctypes.data_as(C.c_void_p), bn_stats.ctypes.data_as(C.c_void_p))

    def forward(self, x):
        """x [n, ncin, F, F] host fp32 -> [n, ncin, out, out]; n may exceed the engine's batch (chunked)."""
        x = np.ascontiguousarray(x, np.float32)
        assert x.ndim == 4 and x.shape[1] == self.ncin
        out = np.empty((x.shape[0], self.ncin, self.out_size, self.out_size), np.float32)
        for i in range(0, x.shape[0], self.batch):
            xi, oi = x[i:i + self.batch], out[i:i + self.batch]
            self._api.cenn_inpainter_forward_host(self.h, xi.ctypes.data_as(C.c_void_p), oi.ctypes.data_as(C.c_void_p), xi.shape[0])
        return out

    def forward_device(self, in_ptr, out_ptr, n):
        self._api.cenn_inpainter_forward_device(self.h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), int(n))

    def sweep(self, frames01, mask_hw, maskValue, init=None, want=("out", "full", "inpaint"), buffers=None):
        """frames01 [P, nc, inh, inw] in [0,1], mask_hw [inh, inw] bool -> (outImages, fullImages, inpaintImages) in [0,1],
        each [P, nc, outh, outw] -- what inpaint_wholeim returns, computed on the device.  Images not named in ``want``
        are not copied back (None in the result); ``buffers`` may hold preallocated (e.g. pinned) result arrays by name."""
        frames01 = np.ascontiguousarray(frames01, np.float32)
        mask = np.ascontiguousarray(mask_hw, np.uint8)
        P, nc, inh, inw = frames01.shape
        F = self.opt["fineSize"]
        outh, outw = -(-inh // F) * F, -(-inw // F) * F
        res = []
        for name in ("out", "full", "inpaint"):
            if name not in want:
                res.append(None)
            elif buffers and name in buffers:
                assert buffers[name].size == P * nc * outh * outw and buffers[name].dtype == np.float32
                res.append(buffers[name].reshape(P, nc, outh, outw))
            else:
                res.append(np.empty((P, nc, outh, outw), np.float32))
        self._api.cenn_inpainter_sweep_host(self.h, init.h if init is not None else None, frames01.ctypes.data_as(C.c_void_p),
                                            mask.ctypes.data_as(C.c_void_p), P, inh, inw, float(maskValue),
                                            *[r.ctypes.data_as(C.c_void_p) if r is not None else None for r in res])
        return tuple(res)
