"""CPU-only: the C-ABI library loads and exports every symbol include/css_b200.h
declares; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "css_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(css_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_builds_and_loads():
    from claude_semantic_search_b200 import build
    lib = build.build_native()
    assert lib.exists()
    ctypes.CDLL(str(lib))


def test_every_declared_symbol_is_exported():
    from claude_semantic_search_b200 import _native
    lib = ctypes.CDLL(str(_native.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in css_b200.h but not exported: {missing}"
    # and the ctypes table binds exactly the declared set
    assert sorted(_native.SIGNATURES) == declared


def test_abi_version_and_error_string():
    from claude_semantic_search_b200 import _native
    lib = _native.load()
    assert lib.css_abi_version() == 1
    assert isinstance(lib.css_last_error(), bytes)


def test_compute_fails_loudly_without_gpu(gpu_available):
    if gpu_available:
        pytest.skip("GPU present")
    from claude_semantic_search_b200 import _native
    with pytest.raises(_native.NoDeviceError):
        _native.Index(768)
    with pytest.raises(_native.NoDeviceError):
        _native.device_count()
    from claude_semantic_search_b200 import HybridStorage, StorageConfig
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        st = HybridStorage(StorageConfig(data_dir=d))   # constructing is allowed
        with pytest.raises(_native.NoDeviceError):
            st.initialize()                              # touching the device is not


def test_product_never_imports_oracle():
    pkg = ROOT / "claude_semantic_search_b200"
    for p in pkg.rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
