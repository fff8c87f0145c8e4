"""Build libgymrs_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU container and on the GPU box alike.
The translation units compile in parallel and are linked into one shared library.  The .so and
objects are git-ignored but travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shlex
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libgymrs_b200.so")
SOURCES = ["kernels_cartpole.cu", "kernels_mountain_car.cu", "kernels_pendulum.cu", "capi.cu"]
HEADERS = ["kernels_impl.cuh", "kernels.hpp", "envs.cuh", "lanes.cuh", "philox.cuh",
           os.path.join(ROOT, "include", "gymrs_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # no -use_fast_math: transcendental and division accuracy is chosen per call site (envs.cuh)
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: cannot build the CUDA extension")


def _extra() -> list:
    """Extra nvcc flags for experiments, e.g. GYMRS_NVCC_EXTRA='-DGYMRS_STEP_MIN_CTAS=4'."""
    return shlex.split(os.environ.get("GYMRS_NVCC_EXTRA", ""))


STAMP = LIB + ".flags"


def _flag_stamp() -> str:
    """The effective nvcc command line: a library built with other flags (an experiment through
    GYMRS_NVCC_EXTRA, say) is stale for a plain build and the other way round."""
    return " ".join(NVCC_FLAGS + _extra())


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    if open(STAMP).read() != _flag_stamp():
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [
        h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    flags = NVCC_FLAGS + _extra() + (["-Xptxas", "-v"] if verbose else [])
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    cmds = [[nvcc] + flags + ["-c", os.path.join(CSRC, s), "-o", o] for s, o in zip(SOURCES, objs)]
    with ThreadPoolExecutor(max_workers=len(cmds)) as ex:
        logs = list(ex.map(_run, cmds))
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
          "-o", LIB] + objs)
    with open(STAMP, "w") as f:
        f.write(_flag_stamp())
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
