"""bench.py on a GPU: one JSON line with every key the measurement contract names (small sizes so
it runs in seconds; the numbers themselves are not checked here)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", ["cartpole", "pendulum"])
def test_bench_line_on_gpu(env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--env", env, "--steps", "40", "--warmup", "3",
                        "--envs", "131072", "--ring", "4", "--burn-in", "10", "--e2e-steps", "6",
                        "--rollout-steps", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 40 and d["warmup"] >= 3 and d["scaling"] == "weak"
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 40 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "l2_policy" in d["config"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e2e = d["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 4 * 131072
    assert e2e["d2h_bytes_per_step"] == (21 if env == "cartpole" else 17) * 131072
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    ck = d["clocks"]
    assert set(ck) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert e2e["pcie_ceiling_gbs"] > 0 and 0 < e2e["frac_of_pcie"] < 1.5
    if env == "cartpole":
        assert e2e["compact"]["value"] > 0 and e2e["compact"]["bytes_per_step"] < (4 + 21) * 131072
    for extra in ("single_stream_chained", "l2_resident", "rollout"):
        assert d[extra]["value"] > 0
    for mode in ("cold_ring", "l2_resident"):
        assert d["single_stream_default"][mode]["value"] > 0
    assert d["timing_sanity"]["regions"] > 0 and "method" in d
    assert d["rollout"]["frac_of_write_peak"] > 0


def test_reference_arm_line(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5",
                        "--warmup", "2", "--envs", "65536", "--ref-repeats", "3"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
