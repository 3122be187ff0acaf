"""util.save / util.load of the reference (``util.lua:25-105``) for host-side module trees.

A network is described on the host as a tree of ``HostModule`` (class name, scalar attributes, numpy
parameters, children) -- the same shape as the Torch7 ``nn`` object graph.  ``save`` follows
``util.save``: buffers (``output``, ``gradInput``, ``finput``, ``fgradInput``) are emptied, parameter
gradients dropped, ``cudnn.SpatialConvolution`` / ``fbnn.SpatialBatchNormalization`` written as their
``nn.*`` equivalents; ``load`` accepts those class names (and the pre-rename ``running_std`` field).
Everything here is pure numpy, so checkpoints can be produced and inspected without a GPU.
"""
import numpy as np

from . import t7

PARAM_FIELDS = ("weight", "bias")
BUFFER_FIELDS = ("running_mean", "running_var")


class HostModule:
    def __init__(self, classname, attrs=None, tensors=None, modules=None):
        self.classname = classname
        self.attrs = dict(attrs or {})
        self.tensors = dict(tensors or {})      # weight, bias, running_mean, running_var (numpy fp32)
        self.modules = modules                  # list for containers, None otherwise

    def walk(self):
        yield self
        for m in self.modules or []:
            yield from m.walk()


def _conv(cls, nin, nout, k, d, p, full=False):
    a = dict(nInputPlane=nin, nOutputPlane=nout, kW=k, kH=k, dW=d, dH=d, padW=p, padH=p)
    if full:
        a.update(adjW=0, adjH=0)
        w = np.zeros((nin, nout, k, k), np.float32)
    else:
        w = np.zeros((nout, nin, k, k), np.float32)
    return HostModule(cls, a, dict(weight=w, bias=np.zeros(nout, np.float32)))


def _bn(c):
    return HostModule("nn.SpatialBatchNormalization", dict(eps=1e-5, momentum=0.1, affine=True, nDim=4),
                      dict(weight=np.ones(c, np.float32), bias=np.zeros(c, np.float32),
                           running_mean=np.zeros(c, np.float32), running_var=np.ones(c, np.float32)))


def _leaky():
    return HostModule("nn.LeakyReLU", dict(negval=0.2, inplace=True))


def _relu():
    return HostModule("nn.ReLU", dict(threshold=0, val=0, inplace=True))


def describe_netG(opt):
    """train.lua:87-150 / train_vid_weighted.lua:112-176 as a host-side tree."""
    video = opt["variant"] == "video"
    nc = opt["nc"] * opt["predLen"] if video else opt["nc"]
    nef, ngf, nB = opt["nef"], opt["ngf"], opt["nBottleneck"]
    C, F = "nn.SpatialConvolution", "nn.SpatialFullConvolution"
    e = [_conv(C, nc, nef, 4, 2, 1), _leaky()]
    for cin, cout in ((nef, nef), (nef, nef * 2), (nef * 2, nef * 4), (nef * 4, nef * 8)):
        e += [_conv(C, cin, cout, 4, 2, 1), _bn(cout), _leaky()]
    e.append(_conv(C, nef * 8, nB, 4, 1, 0))
    netE, nz_size = HostModule("nn.Sequential", modules=e), nB
    if opt.get("noiseGen"):               # train.lua:109-124
        nz = opt.get("nz", 100)
        noise = HostModule("nn.Sequential", modules=[_conv(C, nz, nz, 1, 1, 0)])
        head = [HostModule("nn.ParallelTable", modules=[netE, noise]), HostModule("nn.JoinTable", dict(dimension=2, size=t7.Storage([])))]
        nz_size = nB + nz
    else:
        head = [netE]
    g = head + [_bn(nz_size), _leaky(), _conv(F, nz_size, ngf * 8, 4, 1, 0, True), _bn(ngf * 8), _relu()]
    chain = [(ngf * 8, ngf * 4), (ngf * 4, ngf * 2), (ngf * 2, ngf)] + ([(ngf, ngf)] if video else [])
    for cin, cout in chain:
        g += [_conv(F, cin, cout, 4, 2, 1, True), _bn(cout), _relu()]
    g += [_conv(F, ngf, nc, 4, 2, 1, True), HostModule("nn.Tanh")]
    return HostModule("nn.Sequential", modules=g)


def describe_netD(opt):
    """train.lua:157-202 / train_vid_weighted.lua:183-239."""
    video = opt["variant"] == "video"
    nc = opt["nc"] * opt["predLen"] if video else opt["nc"]
    ndf = opt["ndf"]
    C = "nn.SpatialConvolution"
    d = []
    if video:
        d += [_conv(C, nc, ndf // 2, 4, 2, 1), _leaky(), _conv(C, ndf // 2, ndf, 4, 2, 1), _leaky()]
    elif opt.get("conditionAdv"):         # train.lua:158-180
        ctx = HostModule("nn.Sequential", modules=[_conv(C, nc, ndf, 5, 2, 2)])
        pred = HostModule("nn.Sequential", modules=[_conv(C, nc, ndf, 5, 2, 2 + 32)])
        d += [HostModule("nn.ParallelTable", modules=[ctx, pred]), HostModule("nn.JoinTable", dict(dimension=2, size=t7.Storage([]))), _leaky(),
              _conv(C, ndf * 2, ndf, 4, 2, 1), _bn(ndf), _leaky()]
    else:
        d += [_conv(C, nc, ndf, 4, 2, 1), _leaky()]
    for cin, cout in ((ndf, ndf * 2), (ndf * 2, ndf * 4), (ndf * 4, ndf * 8)):
        d += [_conv(C, cin, cout, 4, 2, 1), _bn(cout), _leaky()]
    d += [_conv(C, ndf * 8, 1, 4, 1, 0), HostModule("nn.Sigmoid"),
          HostModule("nn.View", dict(size=t7.Storage([1]), numElements=1, numInputDims=3))]
    return HostModule("nn.Sequential", modules=d)


def weights_init(net, rng):
    """weights_init of train.lua:58-67 on a host tree: conv / full-conv weight N(0, 0.02), bias 0; BN gamma N(1, 0.02), beta 0."""
    for m in net.walk():
        if "Convolution" in m.classname:
            m.tensors["weight"] = rng.normal(0.0, 0.02, m.tensors["weight"].shape).astype(np.float32)
            m.tensors["bias"][...] = 0
        elif "BatchNormalization" in m.classname:
            m.tensors["weight"] = rng.normal(1.0, 0.02, m.tensors["weight"].shape).astype(np.float32)
            m.tensors["bias"][...] = 0
    return net


# ---- flat vectors in Module:getParameters order (weight, bias per module, module order) ---------------------------
def params_flat(net):
    return np.concatenate([m.tensors[f].ravel() for m in net.walk() for f in PARAM_FIELDS if f in m.tensors]).astype(np.float32)


def set_params_flat(net, flat):
    off = 0
    for m in net.walk():
        for f in PARAM_FIELDS:
            if f in m.tensors:
                n = m.tensors[f].size
                m.tensors[f] = np.asarray(flat[off:off + n], np.float32).reshape(m.tensors[f].shape).copy()
                off += n
    assert off == len(flat), "flat parameter vector has %d elements, network needs %d" % (len(flat), off)


def bn_stats_flat(net):
    return np.concatenate([m.tensors[f].ravel() for m in net.walk() if "running_mean" in m.tensors for f in BUFFER_FIELDS]).astype(np.float32)


def set_bn_stats_flat(net, flat):
    off = 0
    for m in net.walk():
        if "running_mean" in m.tensors:
            for f in BUFFER_FIELDS:
                n = m.tensors[f].size
                m.tensors[f] = np.asarray(flat[off:off + n], np.float32).copy()
                off += n
    assert off == len(flat)


# ---- util.save / util.load ------------------------------------------------------------------------------------
_EMPTY = lambda: np.zeros([0], np.float32)


def _to_t7(m):
    cls = m.classname
    fields = dict(m.attrs)
    if cls == "cudnn.SpatialConvolution":            # util.lua:33-39
        cls = "nn.SpatialConvolution"
    if cls == "fbnn.SpatialBatchNormalization":      # util.lua:40-49
        cls = "nn.SpatialBatchNormalization"
    fields["train"] = fields.get("train", True)
    fields["_type"] = "torch.FloatTensor"
    fields["output"] = _EMPTY()                      # recursiveTableClear (util.lua:54-57)
    fields["gradInput"] = _EMPTY()
    for k, v in m.tensors.items():
        fields[k] = np.ascontiguousarray(v, np.float32).copy()      # clone: no storage offsets (util.lua:64-69)
    if "Convolution" in cls:
        fields["finput"] = _EMPTY()
        fields["fgradInput"] = _EMPTY()
    if "BatchNormalization" in cls:
        fields["save_mean"] = _EMPTY()
        fields["save_std"] = _EMPTY()
    if m.modules is not None:
        fields["modules"] = [_to_t7(c) for c in m.modules]
    # gradWeight / gradBias are nil'ed (util.lua:83): simply not written
    return t7.TorchObject(cls, fields)


def save(filename, net, gpu=0):
    """util.save(filename, net, gpu)."""
    t7.save(filename, _to_t7(net))


def _from_t7(o):
    if not isinstance(o, t7.TorchObject):
        raise ValueError("not a Torch7 module: %r" % (o,))
    cls = {"cudnn.SpatialConvolution": "nn.SpatialConvolution", "fbnn.SpatialBatchNormalization": "nn.SpatialBatchNormalization"}.get(o.classname, o.classname)
    attrs, tensors, modules = {}, {}, None
    for k, v in o.fields.items():
        if k == "modules":
            modules = [_from_t7(c) for c in (v if isinstance(v, list) else [v[i] for i in sorted(v)])]
        elif k in ("weight", "bias", "running_mean", "running_var"):
            if v is not None:
                tensors[k] = np.asarray(v, np.float32)
        elif k == "running_std":                     # pre-rename nn: running_std holds 1/sqrt(var + eps)
            eps = o.fields.get("eps", 1e-5)
            tensors["running_var"] = (1.0 / np.square(np.asarray(v, np.float64)) - eps).astype(np.float32)
        elif k in ("output", "gradInput", "finput", "fgradInput", "gradWeight", "gradBias", "save_mean", "save_std", "_type",
                   "buffer", "buffer2", "centered", "std", "normalized"):
            continue
        else:
            attrs[k] = v
    return HostModule(cls, attrs, tensors, modules)


def load(filename, gpu=0):
    """util.load(filename, gpu): the module tree; zero gradient tensors are implied (util.lua:99-105)."""
    return _from_t7(t7.load(filename))
