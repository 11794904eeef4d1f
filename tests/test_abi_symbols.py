"""The C-ABI library loads and exports every symbol include/rc_api.h declares (no compute)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rc_api.h")


def declared_symbols():
    src = subprocess.run(["gcc", "-E", "-P", HEADER], capture_output=True, text=True, check=True).stdout
    names = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", src))
    assert len(names) > 80
    return names


def test_header_is_plain_c():
    # compiles as C (no C++ / torch types in the signatures)
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_library_exports_every_declared_symbol():
    from rusty_compression_b200 import _lib
    from rusty_compression_b200.build import build
    build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = declared_symbols()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes table binds exactly the declared surface
    assert set(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ declared
    assert lib.rc_version() == 100


def test_library_has_blackwell_native_code():
    """SASS evidence that the hot kernels are what DESIGN.md says: TMA loads feeding the FP64
    tensor pipe (no tcgen05 kind exists for FP64)."""
    from rusty_compression_b200 import _lib
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in out and "UTMALDG" in out and "SYNCS" in out


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from rusty_compression_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")
