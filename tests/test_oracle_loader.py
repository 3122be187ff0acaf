"""CPU: the loader-hook restatement (oracle/loader.py, datavid/donkey_folder.lua:114-129,138-187)."""
import numpy as np

from oracle import loader


def test_hook_crops_masks_flips_and_rescales():
    rng = np.random.default_rng(3)
    B, C, iH, iW, F = 4, 6, 90, 120, 48
    frames = rng.integers(0, 256, (B, C, iH, iW)).astype(np.uint8)
    mask_full = np.zeros((iH, iW), np.uint8); mask_full[5:20, 60:110] = 1
    crop, flip, blocks = loader.draw_hook_params(B, iH, iW, F, rng)
    assert crop[:, 0].min() >= 0 and crop[:, 0].max() <= iH - F and crop[:, 1].max() <= iW - F
    assert 2 <= blocks[:, 0].min() and blocks[:, 0].max() <= 10
    bs = F // 6
    for b in range(B):
        for k in range(blocks[b, 0]):      # 1-based corners in [3, F - bs - 2] keep every block two pixels inside the crop
            assert 2 <= blocks[b, 1 + 2 * k] <= F - bs - 3 and 2 <= blocks[b, 2 + 2 * k] <= F - bs - 3
    crop[0] = (0, 0); flip[0] = 0          # misses the logo -> random blocks
    crop[1] = (0, 60); flip[1] = 1         # contains the logo, flipped
    mv = 110.0 / 255.0
    masked, full, mask = loader.train_hook(frames, mask_full, crop, flip, blocks, F, mv)
    assert full.min() >= -1 and full.max() <= 1 and mask.dtype == np.uint8
    # sample 0: union of the drawn blocks, same on every channel; masked = fill inside, full outside
    exp = np.zeros((F, F), bool)
    for k in range(blocks[0, 0]):
        exp[blocks[0, 2 + 2 * k]:blocks[0, 2 + 2 * k] + bs, blocks[0, 1 + 2 * k]:blocks[0, 1 + 2 * k] + bs] = True
    assert np.array_equal(mask[0, 0].astype(bool), exp) and np.array_equal(mask[0, 0], mask[0, 5])
    assert np.allclose(masked[0][:, exp], 2 * mv - 1) and np.array_equal(masked[0][:, ~exp], full[0][:, ~exp])
    assert np.array_equal(full[0], frames[0, :, :F, :F].astype(np.float32) / np.float32(255) * 2 - 1)
    # sample 1: the logo crop, mirrored
    m1 = (mask_full[0:F, 60:60 + F] != 0)[:, ::-1]
    assert np.array_equal(mask[1, 2].astype(bool), m1)
    assert np.array_equal(full[1], (frames[1, :, :F, 60:60 + F].astype(np.float32) / np.float32(255))[..., ::-1] * 2 - 1)
