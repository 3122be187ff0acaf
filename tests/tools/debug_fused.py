"""Per-layer diagnostics of the fused executor against the oracle (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import nets as onets, step as ostep
import video_filler_b200.tensor as T
from video_filler_b200 import models, train

def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

def flat_modules(net):
    out = []
    for m in net.modules:
        if hasattr(m, "modules"): out += flat_modules(m)
        else: out.append(m)
    return out

variant = sys.argv[1] if len(sys.argv) > 1 else "image"
if len(sys.argv) > 2 and sys.argv[2] == "quant":
    from oracle import nn as onn, ops as oops
    onn.QUANT = oops.bf16_round
kw = dict(batchSize=int(os.environ.get("DBG_B", "8")), nBottleneck=256, nef=64, ngf=64, ndf=64)
if variant == "video": kw["predLen"] = 2
T.state(0)
orc = ostep.StepOracle(onets.default_opt(variant, **kw), seed=1234, dtype=np.float64)
trn = train.FusedTrainer(models.default_opt(variant, **kw))
trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
batch = orc.synth_batch(np.random.default_rng(4321))
lo = orc.step(*batch); lg = trn.step_host(*batch)
print({k: (round(lg[k], 5), round(lo[k], 5)) for k in lo if lo[k] is not None})
for name, net, gref, idx in (("G", orc.netG, orc.gG, 0), ("D", orc.netD, orc.gD, 1)):
    g = trn.get_grads(idx)
    print(f"{name} overall grads rel={rel(g, gref):9.3e} cos={float(np.dot(g, gref) / (np.linalg.norm(g) * np.linalg.norm(gref))):+.5f} norm ratio={np.linalg.norm(g) / np.linalg.norm(gref):.4f}")
    mods = flat_modules(net)
    off = 0; blk = -1
    for m in mods:
        tn = type(m).__name__
        if "Convolution" in tn: blk += 1
        if getattr(m, "weight", None) is None: continue
        for pname in ("weight", "bias"):
            n = getattr(m, pname).size
            a, b = g[off:off + n], gref[off:off + n]
            cos = float(np.dot(a, b) / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
            print(f"{name} blk{blk} {tn:26s} {pname:6s} n={n:8d} rel={rel(a, b):9.3e} cos={cos:+.5f} |ref|max={np.abs(b).max():.3e} |got|max={np.abs(a).max():.3e}")
            off += n
    # activations (fake-pass state for D, G forward)
    blk = -1
    for i, m in enumerate(mods):
        tn = type(m).__name__
        if "Convolution" in tn:
            blk += 1
            # block activation = output of the last pointwise module before the next conv
            j = i + 1
            while j < len(mods) and "Convolution" not in type(mods[j]).__name__ and type(mods[j]).__name__ != "View": j += 1
            act = mods[j - 1].output
            try:
                if blk < len([x for x in mods if "Convolution" in type(x).__name__]) - (1 if name == "D" else 0):
                    got = trn.fetch(f"{name}.{blk}.a").reshape(act.shape)
                    print(f"{name} blk{blk} act   rel={rel(got, act):9.3e}")
                    gi = m.gradInput
            except Exception as e:
                print("fetch failed", name, blk, e)
