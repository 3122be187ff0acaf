"""Generates tests/golden/ops_golden.npz: seeded inputs and the oracle's outputs (float64 arithmetic, stored as float32/64)
for every op on the hot path plus a two-step run of the whole G+D step on a miniature network.

The reference ships no golden vectors and cannot run here (no Torch7 / LuaJIT), so these vectors pin OUR restatement
(oracle/) against drift and give the CUDA path a fixed target; they do not pin Torch7 itself ("parity unpinned").
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nets, ops, step  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    g = {}
    f32 = lambda a: np.asarray(a, np.float32)
    # SpatialConvolution 4x4 / s2 / p1 (E3-like, reduced) and the 4x4 valid bottleneck conv
    for tag, (N, Ci, H, Co, k, d, p) in {"conv_s2": (2, 8, 8, 16, 4, 2, 1), "conv_v4": (3, 8, 4, 5, 4, 1, 0)}.items():
        x, w, b = f32(rng.uniform(-1, 1, (N, Ci, H, H))), f32(rng.normal(0, 0.05, (Co, Ci, k, k))), f32(rng.normal(0, 0.1, Co))
        y = ops.conv_forward(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), d, d, p, p)
        gy = f32(rng.normal(0, 1, y.shape))
        gw, gb = np.zeros(w.shape), np.zeros(b.shape)
        ops.conv_acc_grad(x.astype(np.float64), gy.astype(np.float64), gw, gb, d, d, p, p, 1.0)
        gx = ops.conv_grad_input(x.shape, gy.astype(np.float64), w.astype(np.float64), d, d, p, p)
        g.update({tag + "_x": x, tag + "_w": w, tag + "_b": b, tag + "_gy": gy, tag + "_y": y, tag + "_gx": gx, tag + "_gw": gw, tag + "_gb": gb,
                  tag + "_geom": np.array([k, d, p])})
    # SpatialFullConvolution 4x4 / s2 / p1 and 1x1 -> 4x4
    for tag, (N, Ci, H, Co, k, d, p) in {"full_s2": (2, 16, 4, 8, 4, 2, 1), "full_v4": (3, 6, 1, 8, 4, 1, 0)}.items():
        x, w, b = f32(rng.uniform(-1, 1, (N, Ci, H, H))), f32(rng.normal(0, 0.05, (Ci, Co, k, k))), f32(rng.normal(0, 0.1, Co))
        y = ops.fullconv_forward(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), d, d, p, p)
        gy = f32(rng.normal(0, 1, y.shape))
        gw, gb = np.zeros(w.shape), np.zeros(b.shape)
        ops.fullconv_acc_grad(x.astype(np.float64), gy.astype(np.float64), gw, gb, d, d, p, p, 1.0)
        gx = ops.fullconv_grad_input(gy.astype(np.float64), w.astype(np.float64), d, d, p, p)
        g.update({tag + "_x": x, tag + "_w": w, tag + "_b": b, tag + "_gy": gy, tag + "_y": y, tag + "_gx": gx, tag + "_gw": gw, tag + "_gb": gb,
                  tag + "_geom": np.array([k, d, p])})
    # SpatialBatchNormalization, training mode
    x, gam, bet = f32(rng.normal(0.3, 1.5, (4, 6, 5, 5))), f32(rng.normal(1, 0.02, 6)), f32(rng.normal(0, 0.1, 6))
    rm, rv = np.zeros(6), np.ones(6)
    y, sm, si = ops.bn_forward(x.astype(np.float64), gam.astype(np.float64), bet.astype(np.float64), rm, rv, True)
    gy = f32(rng.normal(0, 1, x.shape))
    gg, gb = np.zeros(6), np.zeros(6)
    gx = ops.bn_backward(x.astype(np.float64), gy.astype(np.float64), gam.astype(np.float64), sm, si, rm, rv, True, ggamma=gg, gbeta=gb)
    g.update(bn_x=x, bn_gamma=gam, bn_beta=bet, bn_gy=gy, bn_y=y, bn_running_mean=rm, bn_running_var=rv, bn_gx=gx, bn_ggamma=gg, bn_gbeta=gb)
    # criteria
    p_, t_ = f32(rng.uniform(0.02, 0.98, (16, 1))), f32(rng.uniform(size=16) > 0.5)
    g.update(bce_x=p_, bce_t=t_, bce_loss=ops.bce_forward(p_.astype(np.float64), t_.astype(np.float64)),
             bce_grad=ops.bce_backward(p_.astype(np.float64), t_.astype(np.float64)))
    a, b = f32(rng.uniform(-1, 1, (2, 3, 8, 8))), f32(rng.uniform(-1, 1, (2, 3, 8, 8)))
    m = (rng.uniform(size=a.shape) > 0.7).astype(np.uint8)
    g.update(crit_x=a, crit_t=b, crit_mask=m,
             mse_loss=ops.mse_forward(a.astype(np.float64), b.astype(np.float64)), mse_grad=ops.mse_backward(a.astype(np.float64), b.astype(np.float64)),
             mmse_loss=ops.masked_mse_forward(a.astype(np.float64), b.astype(np.float64), m, 0.05),
             mmse_grad=ops.masked_mse_backward(a.astype(np.float64), b.astype(np.float64), m, 0.05),
             gdl_loss=ops.gdl_forward(a.astype(np.float64), b.astype(np.float64)), gdl_grad=ops.gdl_backward(a.astype(np.float64), b.astype(np.float64)))
    df = f32(rng.normal(0, 1e-3, a.shape))
    g.update(blend_df=df, blend_overlap=ops.blend_l2_overlap(df.astype(np.float64), a.astype(np.float64), b.astype(np.float64), 0.999, 2),
             blend_masked=ops.blend_l2_masked(df.astype(np.float64), a.astype(np.float64), b.astype(np.float64), m.astype(np.float64), 0.999, 0.05)[0])
    # optim.adam, 3 steps
    x0 = f32(rng.normal(0, 1, 257))
    grads = f32(rng.normal(0, 1, (3, 257)))
    xs, st = x0.astype(np.float64), {}
    for i in range(3):
        ops.adam_step(xs, grads[i].astype(np.float64), st, 2e-3, 0.5)
    g.update(adam_x0=x0, adam_grads=grads, adam_x3=xs)
    # the whole step, two iterations, miniature nets, both variants
    for variant, kw in (("image", dict(batchSize=2, nBottleneck=16, nef=4, ngf=4, ndf=4)),
                        ("video", dict(batchSize=2, nBottleneck=16, nef=4, ngf=4, ndf=4, predLen=2, wtgdl=0.5))):
        orc = step.StepOracle(nets.default_opt(variant, **kw), seed=77, dtype=np.float64)
        g["step_%s_pG0" % variant], g["step_%s_pD0" % variant] = orc.pG.copy(), orc.pD.copy()
        srng = np.random.default_rng(78)
        hist = []
        for i in range(2):
            batch = orc.synth_batch(srng)
            for j, arr in enumerate(batch):
                g["step_%s_in%d_%d" % (variant, i, j)] = arr
            lo = orc.step(*batch)
            hist.append([lo["errD"], lo["errG"], lo["errG_l2"], lo["errG_gdl"] or 0.0, lo["errG_total"]])
        g["step_%s_losses" % variant] = np.array(hist)
        g["step_%s_pG2" % variant], g["step_%s_pD2" % variant] = orc.pG.copy(), orc.pD.copy()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ops_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
