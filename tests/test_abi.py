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
