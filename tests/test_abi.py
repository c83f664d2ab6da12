"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol the header
declares, and fails loudly (B200_ENODEV, "No CUDA devices found") instead of falling back when no
GPU is present. No compute call is made without a GPU."""
import ctypes as C
import re
import subprocess

import pytest

from phantomsdr_b200 import _ffi


def header_symbols():
    # the drop-in boundary plus the tuning / profiling header: every include/*.h declaration must be exported
    text = _ffi.HEADER.read_text() + _ffi.DEBUG_HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_tuning_knobs_are_not_part_of_the_boundary():
    public = re.sub(r"/\*.*?\*/", "", _ffi.HEADER.read_text(), flags=re.S)
    opts = set(re.findall(r"#define (B200_OPT_[A-Z0-9_]+)", public))
    assert opts == {"B200_OPT_RELOAD_BOTH", "B200_OPT_HOST_MIRROR", "B200_OPT_INPUT_FORMAT", "B200_OPT_PEER_STORES", "B200_OPT_PCM16"}
    assert "b200_debug" not in public


def test_library_built_in_tree():
    assert _ffi.LIB_PATH.exists(), "run __graft_entry__.build()"
    assert _ffi.LIB_PATH.parent.name == "lib" and _ffi.LIB_PATH.parents[1].name == "phantomsdr_b200"


def test_exports_every_declared_symbol():
    L = _ffi.lib()
    declared = header_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/phantomsdr_b200.h but not exported"
    # and the ctypes table covers the whole header (no entry point is unreachable from Python)
    assert sorted(_ffi.SIGNATURES) == declared


def test_only_c_abi_symbols_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", str(_ffi.LIB_PATH)], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith("b200_") for n in names), [n for n in names if not n.startswith("b200_")][:5]


def test_abi_version_and_error_convention():
    L = _ffi.lib()
    assert L.b200_abi_version() == 1
    h = C.c_void_p()
    assert L.b200_engine_create(C.byref(h), 1000, 1, 1, 0, 0) == -95  # not a power of two -> B200_ENOTSUP
    assert b"power of two" in L.b200_last_error()
    assert L.b200_engine_create(None, 1 << 16, 1, 1, 0, 0) == -22


def test_no_cpu_fallback_without_gpu():
    L = _ffi.lib()
    if L.b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.b200_engine_create(C.byref(h), 1 << 16, 1, 7, 0, 0)
    assert rc == -19 and not h.value
    assert L.b200_last_error() == b"No CUDA devices found"  # the reference's cuFFT ctor message, fft_cuda.cu:12
    from phantomsdr_b200.backend import B200FFT, B200Error

    with pytest.raises(B200Error):
        B200FFT(1 << 16, 1, 7, 0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under phantomsdr_b200/ may import or load it."""
    pkg = _ffi.PKG
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, path
