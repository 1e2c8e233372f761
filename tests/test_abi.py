"""The C-ABI library loads on a GPU-less host and exports every symbol include/jamie_b200.h declares (no compute)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'jamie_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(jb_[a-z_0-9]+)\s*\(', src)))


def test_header_symbols_exported_and_bound():
    from jamie_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    for n in names:
        assert getattr(lib, n) is not None
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (jb_[a-z_0-9]+)', out))
    assert set(names) <= exported, set(names) - exported
    assert lib.jb_version() >= 100


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from jamie_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError):
        _lib.load()


def test_product_never_imports_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, 'jamie_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(base, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M) or '/root/reference' in txt:
                    bad.append(f)
    assert not bad, bad
