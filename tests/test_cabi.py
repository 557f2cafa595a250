"""CPU: the C-ABI library loads and exports exactly what include/gavatar.h declares (no compute calls here)."""
import os
import re

from gaussianavatar_b200 import _lib
from gaussianavatar_b200.build import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gavatar.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ga_[a-z0-9_]+)\s*\(", txt)))


def test_library_exists_and_loads():
    assert os.path.exists(LIB_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    assert L.ga_version() >= 100
    assert L.ga_last_error() is not None


def test_header_and_binding_agree():
    declared = _declared()
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in gavatar.h but not exported"
        assert name in _lib._SIGNATURES, f"{name} declared in gavatar.h but not bound in _lib.py"
    for name in _lib._SIGNATURES:
        assert name in declared, f"{name} bound in _lib.py but not declared in gavatar.h"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
