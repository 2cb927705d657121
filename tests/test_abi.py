"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol declared in
include/ppg.h; config validation mirrors the reference's errors. No compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from predpreygrass_b200 import _lib
from predpreygrass_b200.config import BASE_CONFIG, PpgConfig, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    from predpreygrass_b200.build import build

    build()


def test_library_exports_every_declared_symbol():
    _build()
    hdr = open(os.path.join(ROOT, "include", "ppg.h")).read()
    declared = set(re.findall(r"\b(ppg_[a-z_]+)\s*\(", hdr))
    declared -= {"ppg_handle"}
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.SYMBOLS) == declared
    assert _lib.load().ppg_abi_version() == 10


def test_struct_layout_matches_header():
    _build()
    L = _lib.load()
    c = PpgConfig()
    L.ppg_default_config(C.byref(c))
    assert c.struct_size == C.sizeof(PpgConfig)
    ref = make_config(BASE_CONFIG, cap_live=(64, 192))
    for f, _ in PpgConfig._fields_:
        a, b = getattr(c, f), getattr(ref, f)
        if hasattr(a, "__len__"):
            assert bytes(a) == bytes(b), f
        else:
            assert a == b, f


def test_create_rejects_bad_config_like_the_reference():
    _build()
    L = _lib.load()
    h = C.c_void_p()
    # "Cannot place more unique positions than grid cells." (BASE:167-168)
    bad = make_config(dict(BASE_CONFIG, grid_size=5, initial_num_grass=100), cap_live=(32, 32))
    assert L.ppg_create(C.byref(bad), 4, 0, C.byref(h)) == 1
    assert b"unique positions" in L.ppg_last_error(None)
    bad = make_config(BASE_CONFIG, cap_live=(64, 192))
    bad.struct_size = 8
    assert L.ppg_create(C.byref(bad), 4, 0, C.byref(h)) == 1
    ok = make_config(BASE_CONFIG, cap_live=(64, 192))
    rc = L.ppg_create(C.byref(ok), 4, 0, C.byref(h))
    import torch

    if not torch.cuda.is_available():
        assert rc == 3  # PPG_ERR_NO_DEVICE: fails loudly, no CPU fallback
    else:
        assert rc == 0
        L.ppg_destroy(h)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "predpreygrass_b200")
    banned = re.compile(r"(^|\n)\s*(from|import)\s+oracle\b|libppg_oracle|ppgo_|oracle/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not banned.search(src), (f, "references the oracle")


def test_batched_env_needs_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from predpreygrass_b200.batched import BatchedPredPreyGrass

    with pytest.raises(_lib.PpgError):
        BatchedPredPreyGrass(make_config(BASE_CONFIG, cap_live=(64, 192)), 4)
