"""Synthetic batches with the data-loader's output contract -- used by benchmarks and examples.

image variant (data/donkey_folder.lua:70-88 + train.lua:286-290): ``real_ctx`` [B,3,F,F] in [-1,1] with the inner
(F/2 - 2*overlapPred)^2 centre filled with the mean colour, and ``real_center`` [B,3,F/2,F/2] (the cloned centre).
video variant (datavid/dataset.lua:426): ``masked``, ``full`` [B,nc,F,F] and the uint8 ``mask`` (shared over channels;
the random-block rule of datavid/donkey_folder.lua:114-129), masked = 2*maskValue-1 under the mask.
"""
import numpy as np

MEAN_FILL = (2 * 117.0 / 255.0 - 1.0, 2 * 104.0 / 255.0 - 1.0, 2 * 123.0 / 255.0 - 1.0)


def image_batch(B, fineSize, overlapPred, rng):
    real = rng.uniform(-1.0, 1.0, (B, 3, fineSize, fineSize)).astype(np.float32)
    q, h = fineSize // 4, fineSize // 2
    center = np.ascontiguousarray(real[:, :, q:q + h, q:q + h])
    ctx = real
    for c in range(3):
        ctx[:, c, q + overlapPred:q + h - overlapPred, q + overlapPred:q + h - overlapPred] = MEAN_FILL[c]
    return ctx, center


def block_mask(fineSize, rng):
    m = np.zeros((fineSize, fineSize), np.uint8)
    blk = fineSize // 6
    for _ in range(int(rng.integers(2, 11))):
        y, x = int(rng.integers(0, fineSize - blk + 1)), int(rng.integers(0, fineSize - blk + 1))
        m[y:y + blk, x:x + blk] = 1
    return m


def video_batch(B, nc, fineSize, maskValue, rng):
    full = rng.uniform(-1.0, 1.0, (B, nc, fineSize, fineSize)).astype(np.float32)
    mask = np.empty((B, nc, fineSize, fineSize), np.uint8)
    for b in range(B):
        mask[b] = block_mask(fineSize, rng)[None]
    masked = np.where(mask != 0, np.float32(2 * maskValue - 1), full).astype(np.float32)
    return masked, full, mask
