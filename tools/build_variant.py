#!/usr/bin/env python
"""Build an experimental variant of the library next to the real one:
    python tools/build_variant.py scalar -DGYMRS_FAST_LANE=1
writes gym_rs_b200/variants/libgymrs_b200_scalar.so; run anything with
    GYMRS_LIB_PATH=gym_rs_b200/variants/libgymrs_b200_scalar.so python bench.py ...
The real library is rebuilt afterwards with the default flags (the flag stamp makes sure of it)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
name, flags = sys.argv[1], sys.argv[2:]
os.environ["GYMRS_NVCC_EXTRA"] = " ".join(flags)
from gym_rs_b200 import build  # noqa: E402

lib = build.build(force=True)
out = os.path.join(ROOT, "gym_rs_b200", "variants")
os.makedirs(out, exist_ok=True)
dst = os.path.join(out, f"libgymrs_b200_{name}.so")
shutil.copy(lib, dst)
os.environ["GYMRS_NVCC_EXTRA"] = ""
build.build(force=True)
print(dst)
