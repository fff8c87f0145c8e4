"""Host-side logic of the multi-GPU path, on CPU with gloo (world size 2): global env-id sharding,
max-over-ranks timing reduction, rank-0-only reference arm.  The device work itself is covered by
test_parity_gpu.py::test_sharding_invariance_at_full_size (same global ids -> same bits)."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 4096                      # envs per rank ("weak scaling": fixed per-GPU work)
    offset = rank * n             # bench.py: rank r owns global ids [r * n, (r + 1) * n)
    # each rank resets its shard with the same seed, keyed by GLOBAL env id (oracle = device stream)
    shard = oracle.reset_batch(oracle.CARTPOLE, n, seed=7, global_env_offset=offset)
    gathered = [torch.zeros(4, n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(shard))
    whole = torch.cat(gathered, dim=1).numpy()
    ref = oracle.reset_batch(oracle.CARTPOLE, n * world, seed=7)
    ok_shard = bool(np.array_equal(whole, ref))
    # timing reduction: the job's time is the MAX over ranks
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = world * n * 10 / float(t.item())
    if rank == 0:
        print(json.dumps({{"ok_shard": ok_shard, "max_t": float(t.item()), "value": value, "world": world}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_sharding_and_max_reduction_with_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok_shard"] and out["world"] == 2
    assert out["max_t"] == 2.0 and out["value"] == 2 * 4096 * 10 / 2.0


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """bench.py --impl reference under torchrun: rank 0 prints the line, the others exit 0 silently."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29572", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "1", "--envs", "8192"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"


def test_bench_line_contract_fields():
    """The keys bench.py promises (checked on the reference arm, which needs no GPU)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--envs", "4096", "--env", "mountain_car"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "MountainCar" in d["config"]["workload"]


def test_shard_range_partitions_exactly():
    from gym_rs_b200.sharding import shard_range
    for total, world in ((1 << 23, 8), (1 << 20, 1), (1000003, 8), (5, 8), (0, 3)):
        ranges = [shard_range(r, world, total) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
            assert e0 == b1 and e0 >= b0
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(3, 8, 1 << 23) == (3 << 20, 4 << 20)   # BASELINE config 5: 1M envs per GPU


def test_device_for_rank_spreads_ranks_over_the_box():
    from gym_rs_b200.sharding import device_for_rank
    # as many ranks as GPUs, or one visible GPU per rank (CUDA_VISIBLE_DEVICES): the identity
    assert [device_for_rank(r, 8, 8) for r in range(8)] == list(range(8))
    assert [device_for_rank(r, 2, 2) for r in range(2)] == [0, 1]
    # fewer ranks than GPUs: spread over the box (different PCIe root complexes), never two ranks on one GPU
    assert [device_for_rank(r, 2, 8) for r in range(2)] == [0, 4]
    assert [device_for_rank(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]
    assert device_for_rank(0, 1, 8) == 0
    assert [device_for_rank(r, 3, 8) for r in range(3)] == [0, 2, 4]
    assert [device_for_rank(r, 2, 8, spread=False) for r in range(2)] == [0, 1]
    for world, visible in ((1, 1), (2, 3), (3, 8), (5, 8), (8, 8)):
        devs = [device_for_rank(r, world, visible) for r in range(world)]
        assert len(set(devs)) == world and max(devs) < visible
    with pytest.raises(ValueError):
        device_for_rank(0, 4, 2)
    with pytest.raises(ValueError):
        device_for_rank(2, 2, 8)
