"""The C-ABI library loads on a CPU-only box and exports exactly what include/stz.h declares; the
weight-blob layout agrees between spec.py and csrc/stz_layout.h; the product path fails loudly
without a GPU (no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import pytest
import torch

import styletts_zs_b200 as stz
from styletts_zs_b200 import path as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return stz.load_library()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "stz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = _header_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"libstz.so lacks {name}"
    assert sorted(P.EXPORTED_SYMBOLS) == declared, "path.py binds a different set than include/stz.h declares"
    assert lib.stz_abi_version() == stz.ABI_VERSION


@pytest.mark.parametrize("cfg", [stz.DEFAULT, stz.TINY], ids=["default", "tiny"])
def test_weight_layout_matches_spec(lib, cfg):
    cc = P._cconfig(cfg)
    offs = stz.weight_offsets(cfg)
    assert lib.stz_weights_nfloats(C.byref(cc)) == offs["__total__"][0]
    for name, (off, shape) in offs.items():
        if name == "__total__":
            continue
        assert lib.stz_weight_offset(C.byref(cc), name.encode()) == off, name
        assert off % 64 == 0
    assert lib.stz_weight_offset(C.byref(cc), b"no.such.entry") == -1
    blob = stz.init_weights(cfg, 0)
    assert blob.numel() == offs["__total__"][0] and bool(torch.isfinite(blob).all())
    assert torch.equal(blob, stz.init_weights(cfg, 0))            # deterministic
    assert not torch.equal(blob, stz.init_weights(cfg, 1))


def test_config_struct_matches_dataclass():
    names = [f[0] for f in P._CConfig._fields_]
    assert names == list(stz.DEFAULT.as_dict().keys())
    hdr = open(os.path.join(ROOT, "include", "stz.h")).read()
    body = re.search(r"typedef struct stz_config \{(.*?)\} stz_config;", hdr, re.S).group(1)
    hdr_names = re.findall(r"\b([a-z_]+)\s*[,;]", body)
    assert hdr_names == names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    with pytest.raises(stz.StzError):
        stz.StyleTTSZSPath(stz.TINY, stz.init_weights(stz.TINY, 0))
    cc = P._cconfig(stz.DEFAULT)
    w = stz.init_weights(stz.DEFAULT, 0)
    h = C.c_void_p()
    rc = lib.stz_create(C.byref(cc), C.c_void_p(w.data_ptr()), w.numel(), 0, C.byref(h))
    assert rc == -3 and not h.value                                # STZ_E_DEVICE
    assert b"no CPU fallback" in lib.stz_last_error(None)
    with pytest.raises(stz.StzError):
        stz.StyleTTSZSPath(stz.DEFAULT, backend="oracle")


def test_create_rejects_bad_arguments(lib):
    cc = P._cconfig(stz.DEFAULT)
    h = C.c_void_p()
    assert lib.stz_create(None, None, 0, 0, C.byref(h)) == -1       # STZ_E_ARG
    bad = P._cconfig(stz.StzConfig(n_heads=7))                       # d_model != 64 * heads
    w = torch.zeros(16)
    assert lib.stz_create(C.byref(bad), C.c_void_p(w.data_ptr()), 16, 0, C.byref(h)) == -2   # STZ_E_SHAPE
    assert lib.stz_sample_style(None, None, None, None, None, None, 1, 1, 1, 1, 1.0, 0, None, None) == -1
    assert lib.stz_predict_duration(None, None, None, None, 1, 1, None, None, None) == -1


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "styletts-zs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
