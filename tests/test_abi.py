"""CPU: the C-ABI shared library loads and exports every symbol include/homan_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes

from homan_b200 import _lib, build


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    handle = _lib.lib()
    assert handle.hm_version() >= 100
    assert len(_lib.SIGNATURES) > 5
    missing = [name for name in _lib.SIGNATURES if not hasattr(handle, name)]
    assert not missing, missing
    assert isinstance(handle.hm_last_error(), bytes)


def test_invalid_arguments_return_error_codes_not_crashes():
    handle = _lib.lib()
    rc = handle.hm_raster_setup(None, None, 1, 1, 1, 1, 256, 1, 1, None, None, None)
    assert rc == -1
    assert b"null" in handle.hm_last_error()
