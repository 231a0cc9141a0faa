"""CPU: the C-ABI library loads and exports every symbol include/vc_b200.h declares; argument
validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

import video_captioning_b200 as vc
from video_captioning_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "vc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vc_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_functions() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = vc.load_library()
    for name in _header_functions():
        assert getattr(lib, name) is not None
    assert lib.vc_version() >= 100


def test_create_rejects_bad_descriptors():
    lib = vc.load_library()
    h = ctypes.c_void_p()
    good = dict(feature_dim=256, hidden_dim=128, embed_dim=128, attn_dim=128, vocab_size=1000, enc_layers=2,
                dec_layers=2, attention=0, num_heads=8, precision=0)
    d = _native.ModelDesc(**good)
    assert lib.vc_model_create(ctypes.byref(d), ctypes.byref(h)) == 0
    assert lib.vc_workspace_bytes(h, 4, 16, 5, 12) > 0
    # not finalized -> decode entry points refuse with a message, no GPU touched
    p = _native.DecodeParams(0, 1, 4, 1, 2, 1.0, 1.0, 0)
    st = lib.vc_decode_greedy(h, 1, 16, None, ctypes.byref(p), ctypes.c_void_p(8), None, ctypes.c_void_p(8), 1 << 40, None)
    assert st == 1 and b"finalized" in lib.vc_last_error()
    lib.vc_model_destroy(h)
    # any vocabulary size is accepted (padded internally): reference checkpoints have V = len(vocabulary)
    d = _native.ModelDesc(**{**good, "vocab_size": 1001})
    h = ctypes.c_void_p()
    assert lib.vc_model_create(ctypes.byref(d), ctypes.byref(h)) == 0
    lib.vc_model_destroy(h)
    for bad in (dict(hidden_dim=100), dict(vocab_size=3), dict(enc_layers=9), dict(precision=7),
                dict(attention=4, num_heads=7), dict(precision=1, embed_dim=40)):
        d = _native.ModelDesc(**{**good, **bad})
        h = ctypes.c_void_p()
        assert lib.vc_model_create(ctypes.byref(d), ctypes.byref(h)) == 1, bad
        assert len(lib.vc_last_error()) > 0


def test_product_has_no_cpu_path():
    import torch
    from oracle import synth
    cfg = synth.make_config("tiny")
    m = vc.VideoCaptioningModel(cfg, 1000)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.generate(torch.zeros(1, 16, 256), 1, 2)
    with pytest.raises(ValueError, match="Unsupported generation method"):
        m.generate(torch.zeros(1, 16, 256), 1, 2, method="sample")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "video-captioning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
