"""Builds the in-tree native library `rvpt_b200/librvpt_b200.so` for sm_100a.

    python -m rvpt_b200.build            # build if sources are newer
    python -m rvpt_b200.build --force

The flags are part of the arithmetic contract (include/rvpt_math.h):
`-fmad=false` on the device and `-ffp-contract=off` on the host keep every
float32 operation separately rounded, which is what makes the GPU result
bit-identical to the CPU oracle.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
CSRC = HERE / "csrc"
LIB = HERE / "librvpt_b200.so"

CUDA_SOURCES = ["kernels.cu", "engine.cu", "bvh_gpu.cu"]
HOST_SOURCES = ["bvh_build.cpp", "camera.cpp"]
HEADLESS = HERE / "rvpt_headless"
HEADLESS_SOURCES = [CSRC / "host" / "rvpt_host.cpp", CSRC / "host" / "rvpt_headless.cpp"]
HEADERS = [CSRC / "device_scene.h", CSRC / "kernels.h", ROOT / "include" / "rvpt_abi.h",
           ROOT / "include" / "rvpt_math.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden,-Wall",
    "-cudart", "static",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the rvpt_b200 CUDA library cannot be built")
    return nvcc


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = ([CSRC / s for s in CUDA_SOURCES + HOST_SOURCES] + HEADERS + [Path(__file__)]
            + HEADLESS_SOURCES + [CSRC / "host" / "rvpt_host.h"])
    if not HEADLESS.exists():
        return True
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    for src in CUDA_SOURCES + HOST_SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(str(obj))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", str(tmp), *objs]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    # the C++ host mirror (RVPT / Camera / load_model) + headless driver, linked against the C ABI
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-o", str(HEADLESS),
           *[str(s) for s in HEADLESS_SOURCES], str(LIB), f"-Wl,-rpath,{HERE}", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
