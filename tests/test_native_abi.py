"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) when there is no device."""
import os
import re

import pytest

from multimodal_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "klnmf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(klnmf_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    lib = _native.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libklnmf.so lacks %s declared in include/klnmf.h" % n
    assert lib.klnmf_abi_version() == _native.ABI_VERSION == 3


def test_python_prototypes_cover_the_header():
    assert sorted(_native.PROTOTYPES) == _declared()


def test_no_cpu_fallback():
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_native.KlnmfError) as e:
        _native.Engine(10, 5, 2)
    assert "no CPU path" in str(e.value)
    import numpy as np
    from multimodal_b200.lib.nmf import KLdivNMF
    with pytest.raises(_native.KlnmfError):
        KLdivNMF(n_components=2, max_iter=3).fit(np.ones((4, 3)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py") or fn.endswith(".cu") or fn.endswith(".cuh"):
                src = open(os.path.join(base, fn)).read()
                assert "oracle" not in src.replace("NMF oracle", ""), fn


def test_bad_mode_rejected():
    with pytest.raises(ValueError):
        _native.resolve_mode("fp8")
    assert _native.resolve_mode("tf32x3") == 1 and _native.resolve_mode("fp64") == 2
