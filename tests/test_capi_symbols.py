"""The C-ABI library loads without a GPU and exports every symbol include/gridforce_b200.h declares; the ctypes
binding lists exactly those symbols; compute entry points fail loudly (no fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gridforce_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"GFB_API\s+[\w\s\*]+?\b(gfb_\w+)\s*\(", text)))


def test_header_declares_symbols():
    names = _declared()
    assert "gfb_kernel_execute_host" in names and "gfb_kernel_execute_device" in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    import openmmgridforce_b200 as gf
    lib = ctypes.CDLL(gf.library_path())
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/gridforce_b200.h but not exported"


def test_binding_covers_header():
    from openmmgridforce_b200 import capi
    assert sorted(capi.SIGNATURES) == _declared()


def test_version_and_no_silent_fallback():
    import openmmgridforce_b200 as gf
    lib = gf.load_library()
    assert lib.gfb_version() == 200
    try:
        n = gf.Device.count()
    except gf.GridForceB200Error:
        n = 0
    if n == 0:
        with pytest.raises(gf.GridForceB200Error):
            gf.Device(0)
