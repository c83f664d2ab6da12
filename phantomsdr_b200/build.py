"""Build the CUDA engine in-tree: phantomsdr_b200/lib/libphantomsdr_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libphantomsdr_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + list((PKG.parent / "include").glob("*.h")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


HOST_SRC = PKG / "host" / "spectrum_loop.cpp"
HOST_BIN = PKG / "lib" / "spectrum_loop"


def build_host(force: bool = False) -> Path:
    """C++ host driver above the C ABI (mirrors the reference's fft_task); links against the in-tree .so."""
    deps = [HOST_SRC, PKG.parent / "include" / "b200_fft.hpp", PKG.parent / "include" / "phantomsdr_b200.h"]
    if not force and HOST_BIN.exists() and all(d.stat().st_mtime <= HOST_BIN.stat().st_mtime for d in deps) \
            and LIB.stat().st_mtime <= HOST_BIN.stat().st_mtime:
        return HOST_BIN
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-pthread", "-o", str(HOST_BIN), str(HOST_SRC),
           f"-L{LIB.parent}", "-lphantomsdr_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + str(LIB.parent),
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return HOST_BIN


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        build_host()
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc(), *NVCC_FLAGS, "-o", str(LIB), *map(str, sources()), "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc only needs a host g++
    env = dict(os.environ)
    subprocess.check_call(cmd, env=env)
    build_host(force=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
