"""Isolate the executor's backward kernels: gate agreement with the (bf16-emulating) oracle, and weight gradients
recomputed in fp64 from the executor's OWN stored tensors (so upstream differences cancel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import nets as onets, step as ostep, ops, nn as onn
import video_filler_b200.tensor as T
from video_filler_b200 import models, train

def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
def cos(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.dot(a, b) / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))
def flat_modules(net):
    out = []
    for m in net.modules:
        out += flat_modules(m) if hasattr(m, "modules") else [m]
    return out

onn.QUANT = ops.bf16_round
kw = dict(batchSize=int(os.environ.get("DBG_B", "8")), nBottleneck=256, nef=64, ngf=64, ndf=64)
T.state(0)
orc = ostep.StepOracle(onets.default_opt("image", **kw), seed=1234, dtype=np.float64)
trn = train.FusedTrainer(models.default_opt("image", **kw))
trn.set_params(0, orc.pG); trn.set_params(1, orc.pD)
batch = orc.synth_batch(np.random.default_rng(4321))
wG0 = orc.pG.copy()
lo = orc.step(*batch); lg = trn.step_host(*batch)
mods = flat_modules(orc.netG)
convs = [i for i, m in enumerate(mods) if "Convolution" in type(m).__name__]
g = trn.get_grads(0)
off = 0
offs = {}
for m in mods:
    if getattr(m, "weight", None) is None: continue
    offs[id(m)] = off
    off += m.weight.size + m.bias.size
for blk, i in enumerate(convs):
    m = mods[i]
    j = i + 1
    while j < len(mods) and "Convolution" not in type(mods[j]).__name__: j += 1
    act_or = mods[j - 1].output
    a = trn.fetch(f"G.{blk}.a").reshape(act_or.shape)
    mism = float(np.mean((a > 0) != (act_or > 0)))
    line = f"G blk{blk} {type(m).__name__:24s} act rel={rel(a, act_or):.2e} gate mismatch={mism:.2e}"
    has_bn = "BatchNorm" in type(mods[i + 1]).__name__
    if has_bn:
        y = trn.fetch(f"G.{blk}.y").reshape(m.output.shape)
        line += f" y rel={rel(y, m.output):.2e}"
    # weight gradient recomputed from the executor's own tensors
    xin = trn.fetch(f"G.{blk}.in").reshape(mods[i - 1].output.shape if i > 0 else batch[0].shape) if blk > 0 else ops.bf16_round(batch[0].astype(np.float64))
    gy = trn.fetch(f"G.{blk}.g").reshape(m.output.shape)
    gw = np.zeros(m.weight.shape); gb = np.zeros(m.bias.shape)
    if type(m).__name__ == "SpatialConvolution":
        ops.conv_acc_grad(xin.astype(np.float64), gy.astype(np.float64), gw, gb, m.dH, m.dW, m.padH, m.padW)
    else:
        ops.fullconv_acc_grad(xin.astype(np.float64), gy.astype(np.float64), gw, gb, m.dH, m.dW, m.padH, m.padW)
    o = offs[id(m)]
    got_w = g[o:o + gw.size]; got_b = g[o + gw.size:o + gw.size + gb.size]
    line += f" | wgrad(self-consistent) rel={rel(got_w, gw):.2e} cos={cos(got_w, gw):.6f}; oracle g_y rel={rel(gy, m.gradInput if False else gy):.0e}"
    # compare executor g_y with the oracle's gradient w.r.t. the conv output
    if has_bn:
        gor = mods[i + 1].gradInput
    else:
        gor = mods[i + 1].gradInput
    line += f" g_y vs oracle rel={rel(gy, gor):.2e} cos={cos(gy, gor):.5f}"
    print(line)
