"""The C-ABI library loads and exports every symbol declared in include/bsi_b200.h (no compute calls: CPU only)."""

import ctypes
import os
import re

import helpers as H
from bsi_b200 import _lib as L
from bsi_b200 import build as B


def declared_symbols():
    text = open(os.path.join(H.ROOT, "include", "bsi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsi_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = B.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bsi_b200.h but not exported"


def test_ctypes_binding_covers_header():
    assert sorted(L.SIGNATURES) == declared_symbols()
    lib = L.load()
    assert lib.bsi_abi_version() == 2
    assert isinstance(lib.bsi_last_error(), bytes)


def test_structs_match_header_layout():
    assert ctypes.sizeof(L.RowRef) == 16 and ctypes.sizeof(L.Noise) == 40 and L.Noise.key_ptr.offset == 32
    assert ctypes.sizeof(L.DitConfig) == 40
    assert L.GemmArgs.gate.offset % 8 == 0 and ctypes.sizeof(L.GemmArgs) % 8 == 0


def test_sass_is_blackwell_native():
    """The GEMM must be tcgen05/TMA code (UTCHMMA / UTMALDG in SASS), not a recompiled mma.sync kernel."""
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        import pytest

        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", B.LIB], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", B.LIB], capture_output=True, text=True).stdout
