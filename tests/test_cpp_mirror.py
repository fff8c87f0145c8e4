"""The C++ host mirror (include/gym_rs.hpp) of the reference's Env surface: compiles everywhere;
on a GPU box the test binary also runs the reference's unit tests + the config-1 plumbing run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "env_test")


def build():
    from gym_rs_b200 import _capi
    import oracle
    _capi.load()
    oracle.build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "env_test.cpp"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
           "-L" + os.path.join(ROOT, "gym_rs_b200"), "-lgymrs_b200",
           "-L" + os.path.join(ROOT, "oracle"), "-lgymrs_oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "gym_rs_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-o", BIN]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return BIN


def test_cpp_mirror_compiles_and_unit_tests_pass_without_gpu():
    """clip / Discrete::contains / seed echo run anywhere; env construction must fail loudly on a
    box without a device (the binary checks that itself)."""
    r = subprocess.run([build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "failures=0" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_plumbing_on_gpu():
    r = subprocess.run([build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "env_test: failures=0" in r.stdout
