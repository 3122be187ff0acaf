"""CPU: the C-ABI library loads without a GPU, exports every symbol include/cenn.h declares, and fails loudly
(no CPU fallback) when asked to run."""
import ctypes as C
import os
import re

import pytest

from video_filler_b200 import _lib


def test_header_declares_the_thnn_table():
    names = set(_lib.PROTOS)
    for op in ("SpatialConvolutionMM", "SpatialFullConvolution"):
        for phase in ("updateOutput", "updateGradInput", "accGradParameters"):
            assert "cenn_%s_%s" % (op, phase) in names
    for n in ("cenn_BatchNormalization_updateOutput", "cenn_BatchNormalization_backward", "cenn_LeakyReLU_updateOutput",
              "cenn_Threshold_updateGradInput", "cenn_Tanh_updateOutput", "cenn_Sigmoid_updateGradInput", "cenn_BCECriterion_updateOutput",
              "cenn_MSECriterion_updateGradInput", "cenn_AbsCriterion_updateOutput", "cenn_MaskedMSECriterion_forward_backward",
              "cenn_GDLCriterion_forward_backward", "cenn_WeightedMSEBlend_overlap", "cenn_WeightedMSEBlend_masked", "cenn_MaskComposite",
              "cenn_AdamFlat", "cenn_trainer_create", "cenn_trainer_step_host", "cenn_trainer_step_phase", "cenn_trainer_sync_info"):
        assert n in names, n
    assert len(names) >= 80


def test_library_exports_every_declared_symbol():
    lib = _lib.load()                 # getattr on every prototype: AttributeError if one is missing
    for name in _lib.PROTOS:
        assert getattr(lib, name) is not None
    assert lib.cenn_version().startswith(b"cenn")


def test_no_torch_types_in_signatures():
    src = open(_lib.HEADER).read()
    assert "at::" not in src and "torch::" not in src and "THCudaTensor *" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def test_trainer_config_layout_matches_header():
    src = open(_lib.HEADER).read()
    body = re.search(r"typedef struct cenn_trainer_config \{(.*?)\} cenn_trainer_config;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, names = decl.split(None, 1)
        fields += [(n.strip(), typ) for n in names.split(",")]
    assert [f[0] for f in fields] == [f[0] for f in _lib.TrainerConfig._fields_]
    for (n, typ), (_, ct) in zip(fields, _lib.TrainerConfig._fields_):
        assert {"int": C.c_int, "float": C.c_float}[typ] is ct, n
    assert C.sizeof(_lib.TrainerConfig) == 4 * len(fields)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_fails_loudly_without_gpu():
    lib = _lib.load()
    out = C.c_void_p()
    assert lib.cenn_init(0, C.byref(out)) != 0
    assert b"CUDA" in lib.cenn_last_error() or b"device" in lib.cenn_last_error()
    import video_filler_b200.tensor as T
    with pytest.raises(_lib.CennError):
        T.CudaTensor(4)


def _struct_fields(src, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            typ, names = decl.split(None, 1)
            fields += [(n.strip(), typ) for n in names.split(",")]
    return fields


def test_inpainter_config_layout_matches_header():
    fields = _struct_fields(open(_lib.HEADER).read(), "cenn_inpainter_config")
    assert [f[0] for f in fields] == [f[0] for f in _lib.InpainterConfig._fields_]
    assert all(t == "int" for _, t in fields) and C.sizeof(_lib.InpainterConfig) == 4 * len(fields)


def test_lua_shim_prototypes_match_header():
    """lua/cenn.lua cannot be executed here (no LuaJIT / Torch7): check statically that every prototype in its ffi.cdef
    blocks names an exported function with the same parameter list, and that its struct declarations match the header."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lua = open(os.path.join(root, "lua", "cenn.lua")).read()
    cdefs = "\n".join(re.findall(r"ffi\.cdef\[\[(.*?)\]\]", lua, re.S))
    header = re.sub(r"/\*.*?\*/", " ", open(_lib.HEADER).read(), flags=re.S)

    def norm(args):
        out = []
        for a in args.split(","):
            a = re.sub(r"\s+", " ", a.strip())
            m = re.match(r"(.*?)(\w+)$", a)
            out.append(re.sub(r"\s*\*\s*", "*", m.group(1).strip()))
        return out

    hdr = {m.group(1): norm(m.group(2)) for m in re.finditer(r"CENN_API\s+[\w\s\*]+?\b(cenn_\w+)\s*\(([^;]*?)\)\s*;", header, re.S)}
    seen = 0
    for m in re.finditer(r"\b(?:int|const char \*)\s*(cenn_\w+)\s*\(([^;]*?)\)\s*;", cdefs, re.S):
        name, args = m.group(1), m.group(2)
        assert name in hdr, "lua/cenn.lua declares %s, which include/cenn.h does not export" % name
        if args.strip() != "void":
            assert norm(args) == hdr[name], name
        seen += 1
    assert seen >= 50
    for struct in ("cenn_trainer_config", "cenn_inpainter_config"):
        lua_fields = []
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), cdefs, re.S).group(1)
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                typ, names = decl.split(None, 1)
                lua_fields += [(n.strip(), typ) for n in names.split(",")]
        assert lua_fields == _struct_fields(open(_lib.HEADER).read(), struct), struct
    # every lib.cenn_* call in the shim is declared in its cdef
    declared = set(re.findall(r"\b(cenn_\w+)\s*\(", cdefs))
    for call in set(re.findall(r"\blib\.(cenn_\w+)", lua)):
        assert call in declared, "lua/cenn.lua calls %s without declaring it" % call


def _shim_tensor_methods():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lua = open(os.path.join(root, "lua", "cenn.lua")).read()
    meths = set(re.findall(r"function CudaTensor[:.](\w+)", lua))
    if "CudaTensor.__index = function" in lua:
        meths.add("__index")
    if "CudaTensor.__newindex = function" in lua:
        meths.add("__newindex")
    return lua, meths


def test_lua_shim_implements_every_tensor_method_the_scripts_call():
    """`train*.lua run unchanged` needs every method the scripts call on GPU tensors.  The call sites are a committed
    fixture (tools/scan_lua_methods.py over the reference's Lua files); optim.adam's tensor calls and SURVEY 9.11's list are added."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fx = json.load(open(os.path.join(root, "tests", "golden", "lua_tensor_methods.json")))
    lua, have = _shim_tensor_methods()
    need = set(fx["methods"])
    need |= {"mul", "add", "addcmul", "addcdiv", "sqrt", "copy", "clone", "new", "resizeAs", "zero"}             # optim.adam
    need |= {"float", "fill", "cmul", "normal", "uniform", "maskedSelect", "maskedCopy", "maskedFill", "view", "size", "min", "max", "mean", "std",
             "nElement", "dim", "resize", "type", "typeAs", "contiguous", "narrow"}                               # SURVEY 9.11 + nn containers
    missing = sorted(m for m in need if m not in have)
    assert not missing, "lua/cenn.lua CudaTensor lacks: %s" % missing
    for c in fx["calls"]:
        # a 4-D box index yields a Window (fill / clone / copy); an index on the first dimension only yields a narrow() view (every tensor method)
        if c["indexed"] and c["method"] in ("fill", "clone", "copy"):
            assert re.search(r"function Window:%s\b" % c["method"], lua), "range-indexed %s:%s (%s:%d) has no Window method" % (c["receiver"], c["method"], c["file"], c["line"])
    assert "first_dim_only" in lua and "t:narrow(1, a, b - a + 1)" in lua
    # module:cuda() / :float() and criterion:cuda() (train.lua:254-256; util.lua:76,81)
    for fn in ("function nn.Module:cuda()", "function nn.Module:float()", "function nn.Criterion:cuda()"):
        assert fn in lua, fn
    # the fixture is current when the reference is at hand (it is not on the GPU box)
    if os.path.isdir("/root/reference"):
        import sys
        sys.path.insert(0, os.path.join(root, "tools"))
        import scan_lua_methods
        assert sorted({c["method"] for c in scan_lua_methods.scan("/root/reference")}) == fx["methods"]


def test_debug_header_symbols_are_exported_and_kept_out_of_the_boundary():
    """The probes live in csrc/probes.cu and are declared in include/cenn_debug.h, NOT in the drop-in boundary include/cenn.h."""
    import ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dbg = open(os.path.join(root, "include", "cenn_debug.h")).read()
    names = re.findall(r"CENN_API\s+int\s+(cenn_debug_\w+)\s*\(", dbg)
    assert len(names) >= 4
    assert "cenn_debug_" not in open(os.path.join(root, "include", "cenn.h")).read()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        getattr(lib, n)
    assert "cenn_debug_" not in open(os.path.join(root, "video-filler_b200", "csrc", "conv_tc.cu")).read()
