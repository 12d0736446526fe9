"""CPU: the oracle is test infrastructure -- nothing in the product package may import or execute it."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for dp, _, fs in os.walk(d):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                yield os.path.join(dp, f)


def test_product_never_touches_the_oracle_or_the_reference():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|/root/reference|oracle/_ref', re.M)
    bad = [p for p in _py_files(os.path.join(ROOT, 'evreal_b200')) if pat.search(open(p).read())]
    assert bad == []


def test_no_compat_layers_in_product():
    pat = re.compile(r'^\s*(from|import)\s+(triton|tilelang)\b|torch\.compile\(', re.M)
    bad = [p for p in _py_files(os.path.join(ROOT, 'evreal_b200')) if pat.search(open(p).read())]
    assert bad == []


def test_oracle_headers_say_test_infrastructure():
    for f in os.listdir(os.path.join(ROOT, 'oracle')):
        if f.endswith('.py'):
            assert 'TEST INFRASTRUCTURE ONLY' in open(os.path.join(ROOT, 'oracle', f)).read(), f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    from evreal_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError):
        _lib.load()
