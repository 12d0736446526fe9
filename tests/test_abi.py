"""CPU: the C-ABI library loads and exports every symbol include/evreal_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'evreal_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(evk_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ('evk_voxelize', 'evk_voxelize_raw', 'evk_normalize_pad', 'evk_model_create', 'evk_model_forward',
                 'evk_model_reset_states', 'evk_mse_ssim', 'evk_percentile_normalize', 'evk_crop', 'evk_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from evreal_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from evreal_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_ctypes_signature_table_covers_the_header():
    from evreal_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.evk_version() >= 100
    assert isinstance(lib.evk_last_error(), bytes)


def test_bad_arguments_fail_without_a_gpu():
    from evreal_b200 import _lib
    import pytest
    lib = _lib.load()
    rc = lib.evk_voxelize(None, None, None, None, 0, 5, 8, 8, None, None, None)
    assert rc == _lib.EVK_ERR_ARG
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert b'null' in lib.evk_last_error()
