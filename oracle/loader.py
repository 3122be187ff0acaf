"""The video loader's per-sample hook restated (TEST INFRASTRUCTURE): ``datavid/donkey_folder.lua:114-129`` (randomBlockMask)
and ``:138-187`` (trainHook with a mask).  The random draws are explicit arguments so that the device-side hook
(``cenn_trainer_step_frames_host``) can be checked against it on identical draws; ``draw_hook_params`` restates how the
script draws them (torch.uniform / torch.random ranges), with a numpy generator standing in for Torch's Mersenne twister.
"""
import numpy as np


def draw_hook_params(B, iH, iW, F, rng):
    """crop origins (0-based), hflip flags and random-block tables for B samples.

    :149-150  h1 = ceil(uniform(1e-2, iH - oH)), w1 likewise; image.crop(input, w1, h1, w1 + oW, h1 + oH) takes columns [w1, w1 + oW)
    :172      hflip when uniform() > 0.5
    :118-123  blockSize = floor(h / 6); nBlocks = random(2, 10); tlx = random(3, w - blockSize - 2) (1-based, inclusive), tly likewise
    """
    crop = np.zeros((B, 2), np.int32)
    flip = np.zeros(B, np.uint8)
    blocks = np.zeros((B, 21), np.int32)
    bs = F // 6
    for b in range(B):
        crop[b, 0] = min(max(int(np.ceil(rng.uniform(1e-2, max(iH - F, 1e-2)))), 0), iH - F)
        crop[b, 1] = min(max(int(np.ceil(rng.uniform(1e-2, max(iW - F, 1e-2)))), 0), iW - F)
        flip[b] = 1 if rng.uniform() > 0.5 else 0
        n = int(rng.integers(2, 11))
        blocks[b, 0] = n
        for k in range(n):
            blocks[b, 1 + 2 * k] = int(rng.integers(3, F - bs - 2 + 1)) - 1      # tlx, 1-based in the script
            blocks[b, 2 + 2 * k] = int(rng.integers(3, F - bs - 2 + 1)) - 1      # tly
    return crop, flip, blocks


def train_hook(frames_u8, mask_full, crop, flip, blocks, F, maskValue):
    """One batch through trainHook(path, withMask=true): returns (masked, full, mask) = the loader's (masked, out, maskout),
    float32 [B,C,F,F] in [-1,1] and uint8 [B,C,F,F] (datavid/dataset.lua:426 order)."""
    B, C = frames_u8.shape[:2]
    bs = F // 6
    full = np.empty((B, C, F, F), np.float32)
    masked = np.empty((B, C, F, F), np.float32)
    mask = np.empty((B, C, F, F), np.uint8)
    for b in range(B):
        h1, w1 = int(crop[b, 0]), int(crop[b, 1])
        out = frames_u8[b, :, h1:h1 + F, w1:w1 + F].astype(np.float32) / np.float32(255.0)          # image.load: byte / 255; image.crop (:151)
        m = (mask_full[h1:h1 + F, w1:w1 + F] != 0)
        if m.max() <= 0:                                                                              # mask crop all black -> randomBlockMask (:163-168)
            m = np.zeros((F, F), bool)
            for k in range(int(blocks[b, 0])):
                tlx, tly = int(blocks[b, 1 + 2 * k]), int(blocks[b, 2 + 2 * k])
                m[tly:tly + bs, tlx:tlx + bs] = True
        mk = np.where(m[None], np.float32(maskValue), out)                                            # maskedFill in [0,1] (:166 / :124)
        mx = np.broadcast_to(m[None], (C, F, F)).astype(np.uint8)
        if flip is not None and flip[b]:                                                              # image.hflip of all three (:172-178)
            out, mk, mx = out[..., ::-1], mk[..., ::-1], mx[..., ::-1]
        full[b] = out * np.float32(2) - np.float32(1)                                                 # out:mul(2):add(-1) (:180-183)
        masked[b] = mk * np.float32(2) - np.float32(1)
        mask[b] = mx
    return masked, full, mask
