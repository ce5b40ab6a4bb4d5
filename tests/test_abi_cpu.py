"""The C-ABI library loads on a CPU-only box, exports every symbol include/kws.h declares,
and the product path fails loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "kws.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kws_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header():
    from speech_recognition_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from speech_recognition_b200.build import build
        build()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libkws.so does not export {s}"
        assert s in _lib.SIGNATURES, f"ctypes binding missing for {s}"
    assert lib.kws_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speech_recognition_b200 import Engine, KwsError
    with pytest.raises(KwsError, match="no CUDA device"):
        Engine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "speech_recognition_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
