"""CPU: the Torch7 .t7 format and the util.save / util.load semantics (util.lua:25-105)."""
import io
import struct

import numpy as np
import pytest

from video_filler_b200 import models, t7, util


def _i(v): return struct.pack("<i", v)
def _l(v): return struct.pack("<q", v)
def _s(s): return _i(len(s)) + s.encode()


def test_reader_on_hand_built_bytes(tmp_path):
    """A file assembled by hand from the format description: {weight = FloatTensor(2,3), n = 4, ok = true, name = 'x'}
    wrapped in an nn.Identity-like torch object; the tensor is referenced twice (second time by index only)."""
    data = np.arange(6, dtype="<f4")
    tensor = (_i(4) + _i(3) + _s("V 1") + _s("torch.FloatTensor") + _i(2) + _l(2) + _l(3) + _l(3) + _l(1) + _l(1)
              + _i(4) + _i(4) + _s("V 1") + _s("torch.FloatStorage") + _l(6) + data.tobytes())
    table = (_i(3) + _i(2) + _i(5)
             + _i(2) + _s("weight") + tensor
             + _i(2) + _s("n") + _i(1) + struct.pack("<d", 4.0)
             + _i(2) + _s("ok") + _i(5) + _i(1)
             + _i(2) + _s("name") + _i(2) + _s("x")
             + _i(2) + _s("alias") + _i(4) + _i(3))
    blob = _i(4) + _i(1) + _s("V 1") + _s("nn.Identity") + table
    p = tmp_path / "hand.t7"
    p.write_bytes(blob)
    o = t7.load(str(p))
    assert isinstance(o, t7.TorchObject) and o.classname == "nn.Identity"
    assert np.array_equal(o["weight"], np.arange(6, dtype=np.float32).reshape(2, 3))
    assert o["n"] == 4 and o["ok"] is True and o["name"] == "x"
    assert o["alias"] is o["weight"]                      # shared reference index


def test_writer_bytes_of_small_objects(tmp_path):
    p = tmp_path / "n.t7"
    t7.save(str(p), 2.5)
    assert p.read_bytes() == _i(1) + struct.pack("<d", 2.5)
    t7.save(str(p), "ab")
    assert p.read_bytes() == _i(2) + _s("ab")
    t7.save(str(p), [True, None])
    assert p.read_bytes() == _i(3) + _i(1) + _i(2) + _i(1) + struct.pack("<d", 1.0) + _i(5) + _i(1) + _i(1) + struct.pack("<d", 2.0) + _i(0)
    a = np.arange(4, dtype=np.float32).reshape(2, 2)
    t7.save(str(p), a)
    assert p.read_bytes() == (_i(4) + _i(1) + _s("V 1") + _s("torch.FloatTensor") + _i(2) + _l(2) + _l(2) + _l(2) + _l(1) + _l(1)
                              + _i(4) + _i(2) + _s("V 1") + _s("torch.FloatStorage") + _l(4) + a.tobytes())


def test_roundtrip_types(tmp_path):
    obj = {"a": [1, 2.5, "s", False, None], "t": np.random.default_rng(0).normal(size=(3, 1, 2)).astype(np.float32),
           "b": np.array([1, 0, 1], np.uint8), "empty": np.zeros([0], np.float32), "size": t7.Storage([1]),
           "o": t7.TorchObject("nn.Tanh", {"train": True})}
    p = tmp_path / "r.t7"
    t7.save(str(p), obj)
    r = t7.load(str(p))
    assert r["a"] == [1, 2.5, "s", False, None]
    assert np.array_equal(r["t"], obj["t"]) and r["t"].dtype == np.float32
    assert np.array_equal(r["b"], obj["b"]) and r["b"].dtype == np.uint8
    assert r["empty"].size == 0
    assert isinstance(r["size"], t7.Storage) and list(r["size"]) == [1]
    assert r["o"].classname == "nn.Tanh" and r["o"]["train"] is True


@pytest.mark.parametrize("variant", ["image", "video"])
def test_util_save_load_roundtrip(tmp_path, variant):
    opt = models.default_opt(variant, nBottleneck=32, nef=8, ngf=8, ndf=8)
    rng = np.random.default_rng(1)
    for describe in (util.describe_netG, util.describe_netD):
        net = describe(opt)
        flat = rng.normal(0, 0.02, util.params_flat(net).size).astype(np.float32)
        util.set_params_flat(net, flat)
        stats = rng.uniform(0.5, 1.5, util.bn_stats_flat(net).size).astype(np.float32)
        util.set_bn_stats_flat(net, stats)
        p = tmp_path / "net.t7"
        util.save(str(p), net)
        raw = t7.load(str(p))
        # util.save semantics: nn.* class names, emptied buffers, no gradient tensors
        assert raw.classname == "nn.Sequential" and raw["output"].size == 0
        first = raw["modules"][0]
        conv = first["modules"][0] if first.classname == "nn.Sequential" else first
        assert conv.classname == "nn.SpatialConvolution" and "gradWeight" not in conv and conv["finput"].size == 0
        assert conv["weight"].dtype == np.float32 and conv["weight"].shape[2:] == (4, 4)
        back = util.load(str(p))
        assert np.array_equal(util.params_flat(back), flat)            # tensor-bit-identical after load
        assert np.array_equal(util.bn_stats_flat(back), stats)
        assert [m.classname for m in back.walk()] == [m.classname for m in net.walk()]


def test_param_counts_match_reference_nets():
    # cfg1/2 conv parameters G 71,118,691, D 2,764,737 (+ 2C per BN layer)
    opt = models.default_opt("image")
    g = util.params_flat(util.describe_netG(opt)).size
    d = util.params_flat(util.describe_netD(opt)).size
    assert g == 71118691 + 2 * (64 + 128 + 256 + 512 + 4000 + 512 + 256 + 128 + 64)
    assert d == 2764737 + 2 * (128 + 256 + 512)
    opt = models.default_opt("video")
    assert util.params_flat(util.describe_netG(opt)).size == 71202732 + 2 * (64 + 128 + 256 + 512 + 4000 + 512 + 256 + 128 + 64 + 64)
    assert util.params_flat(util.describe_netD(opt)).size == 2800609 + 2 * (128 + 256 + 512)


def test_load_converts_legacy_classes(tmp_path):
    old = t7.TorchObject("nn.Sequential", {"modules": [
        t7.TorchObject("cudnn.SpatialConvolution", dict(nInputPlane=3, nOutputPlane=2, kW=4, kH=4, dW=2, dH=2, padW=1, padH=1,
                                                         weight=np.ones((2, 3, 4, 4), np.float32), bias=np.zeros(2, np.float32))),
        t7.TorchObject("fbnn.SpatialBatchNormalization", dict(eps=1e-5, momentum=0.1, affine=True, weight=np.ones(2, np.float32),
                                                              bias=np.zeros(2, np.float32), running_mean=np.zeros(2, np.float32),
                                                              running_std=np.full(2, 0.5, np.float32)))]})
    p = tmp_path / "old.t7"
    t7.save(str(p), old)
    net = util.load(str(p))
    assert [m.classname for m in net.modules] == ["nn.SpatialConvolution", "nn.SpatialBatchNormalization"]
    assert np.allclose(net.modules[1].tensors["running_var"], 1 / 0.25 - 1e-5)


def test_noisegen_conditionadv_trees_roundtrip_and_match_oracle_layout(tmp_path):
    """train.lua:109-124,158-180: ParallelTable / JoinTable(2) containers in the host tree -- same getParameters order and
    size as the oracle nets, and a util.save / util.load round trip that keeps the container classes."""
    from oracle import nets as onets
    kw = dict(nBottleneck=32, nef=8, ngf=8, ndf=8, noiseGen=1, nz=12, conditionAdv=1)
    opt = models.default_opt("image", **kw)
    oG, oD = onets.build_netG(onets.default_opt("image", **kw)), onets.build_netD(onets.default_opt("image", **kw))
    rng = np.random.default_rng(3)
    for describe, onet in ((util.describe_netG, oG), (util.describe_netD, oD)):
        net = describe(opt)
        oflat, _ = onet.getParameters()
        assert util.params_flat(net).size == oflat.size
        # same order: fill the oracle's flat vector with its own indices, copy into the host tree, compare per-tensor shapes
        oflat[...] = np.arange(oflat.size)
        util.set_params_flat(net, oflat.astype(np.float32))
        holders = []
        onet._collect_holders(holders)
        host_tensors = [m.tensors[f] for m in net.walk() for f in util.PARAM_FIELDS if f in m.tensors]
        assert len(holders) == len(host_tensors)
        for (m, pn, _), ht in zip(holders, host_tensors):
            assert getattr(m, pn).shape == ht.shape and np.array_equal(getattr(m, pn).astype(np.float32), ht)
        flat = rng.normal(0, 0.02, oflat.size).astype(np.float32)
        util.set_params_flat(net, flat)
        p = tmp_path / "net.t7"
        util.save(str(p), net)
        back = util.load(str(p))
        names = [m.classname for m in back.walk()]
        assert "nn.ParallelTable" in names and "nn.JoinTable" in names and names == [m.classname for m in net.walk()]
        assert np.array_equal(util.params_flat(back), flat)
        join = [m for m in back.walk() if m.classname == "nn.JoinTable"][0]
        assert join.attrs["dimension"] == 2


def test_cyclic_tables_keep_identity_and_empty_tables_map_to_lists(tmp_path):
    """ADVICE r1 (t7 reader): a back-reference taken while a list-like table is still being read (nngraph gModules are cyclic: node ->
    children -> node) must end up pointing at the SAME object the table became, and the 0-entry table has one fixed mapping."""
    node = t7.TorchObject("nngraph.Node", {"id": 1})
    children = [node, t7.TorchObject("nngraph.Node", {"id": 2})]
    node.fields["children"] = children          # cycle: children[0].children is children
    root = {"nodes": children, "empty": [], "first": node}
    p = str(tmp_path / "cyc.t7")
    t7.save(p, root)
    back = t7.load(p)
    assert isinstance(back["nodes"], list) and len(back["nodes"]) == 2
    assert back["nodes"][0] is back["first"]
    assert back["first"].fields["children"] is back["nodes"]          # not a stale dict copy
    assert back["nodes"][0].fields["children"][1].fields["id"] == 2
    assert back["empty"] == [] and isinstance(back["empty"], list)
    # an object whose field table is empty keeps an empty dict of fields
    p2 = str(tmp_path / "e.t7")
    t7.save(p2, t7.TorchObject("nn.Identity", {}))
    assert t7.load(p2).fields == {}
