"""CPU: the C-ABI library loads and exports every symbol include/cupoch_b200.h declares; the product
path fails loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cupoch_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cphb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from cupoch_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS), set(syms) ^ set(_lib.EXPORTED_SYMBOLS)
    assert L.cphb_version() == 100


def test_no_cpu_fallback():
    import cupoch_b200 as cph
    from cupoch_b200 import _lib
    if _lib.lib().cphb_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.CphbError):
        cph.geometry.PointCloud(np.zeros((4, 3), np.float32))
    with pytest.raises(_lib.CphbError):
        cph.geometry.KDTreeFlann().set_geometry(None)


def test_product_never_imports_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "cupoch_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "orc_" not in txt, f
