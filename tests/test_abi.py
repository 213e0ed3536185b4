"""The C-ABI library loads and exports every symbol include/oarfish_em.h declares
(no compute calls here: this file runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oarfish_em.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oar_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("oar_store_create", "oar_store_destroy", "oar_em", "oar_bootstrap", "oar_bootstrap_weights",
                 "oar_bootstrap_sample_weights", "oar_em_batched", "oar_last_error", "oar_version"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from oarfish_b200 import _lib
    lib = ctypes.CDLL(_lib.EM_LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"


def test_python_binding_table_matches_header():
    from oarfish_b200 import _lib
    assert sorted(_lib.ABI) == declared_symbols()
    lib = _lib.load_em_lib()
    assert lib.oar_version() >= 1000


def test_signatures_use_no_torch_types():
    text = open(HEADER).read()
    assert "torch" not in text and "at::" not in text and "Tensor" not in text


def test_product_does_not_import_the_oracle():
    # the oracle is test infrastructure: nothing under oarfish_b200/ may reference it
    pkg = os.path.join(ROOT, "oarfish_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboarfish_oracle" not in src, f


def test_compute_fails_loudly_without_a_device():
    import numpy as np
    from oarfish_b200 import _lib, DeviceStore
    lib = _lib.load_em_lib()
    if lib.oar_device_count() > 0:
        pytest.skip("a CUDA device is present")
    rp = np.array([0, 1], dtype=np.uint64)
    with pytest.raises(_lib.OarfishError) as ei:
        DeviceStore(rp, np.zeros(1, np.uint32), np.ones(1, np.float32), 1)
    assert ei.value.code in (_lib.OAR_ERR_CUDA,)
