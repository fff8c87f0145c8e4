"""The example programs (examples/*.py, the counterparts of the reference's examples/*.rs) run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,expect", [("cartpole.py", "episode returns:"), ("mountain_car.py", "reward of the last 200: -200.0"),
                                         ("batched_rollout.py", "gymrs_rollout")])
def test_example_runs(name, expect):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", name)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert expect in r.stdout, r.stdout
