"""The reset stream is "in-kernel curand" (BASELINE.json north_star): the hand-written Philox4x32-10 of
csrc/philox.cuh -- restated in oracle/ and pinned there by the Random123 known answers -- returns
exactly what cuRAND's device generator returns for
    curand_init(seed, subsequence = epoch, offset = 4 * global_env_id);  curand4()
and a device reset is those words mapped to U[low, high).  tests/cuda/curand_check is a test helper
compiled by __graft_entry__.build(); it links cuRAND's header-only device API, not this library.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cuda", "curand_check")


def curand_words(triples):
    if not os.path.exists(BIN):
        import __graft_entry__
        __graft_entry__.build_test_helpers()
    text = "".join(f"{s} {g} {e}\n" for s, g, e in triples)
    r = subprocess.run([BIN], input=text, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return [[int(x) for x in line.split()] for line in r.stdout.splitlines()]


def test_philox_equals_curand_philox4_32_10():
    rng = np.random.default_rng(0)
    triples = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (2 ** 64 - 1, 2 ** 40 - 1, 2 ** 32 + 5)]
    triples += [(int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 40)), int(rng.integers(0, 2 ** 33)))
                for _ in range(200)]
    got = curand_words(triples)
    for (seed, gid, epoch), words in zip(triples, got):
        ctr = [gid & 0xFFFFFFFF, gid >> 32, epoch & 0xFFFFFFFF, epoch >> 32]
        key = [seed & 0xFFFFFFFF, seed >> 32]
        assert oracle.philox4x32_10(ctr, key) == words, (seed, gid, epoch)


def test_device_reset_is_curand_words_mapped_to_the_reset_box():
    import gym_rs_b200 as g
    n, seed, off = 4096, 2024, 777
    env = g.CartPoleEnv(num_envs=n, global_env_offset=off)
    env.reset(seed=seed)
    st = env.get_state()
    words = np.array(curand_words([(seed, off + i, 0) for i in range(n)]), dtype=np.uint64)
    # U[-0.05, 0.05) on a 2^-24 grid: low + (w >> 8) * 2^-24 * (high - low)   (csrc/philox.cuh)
    lo, hi = np.float32(-0.05), np.float32(0.05)
    scale24 = np.float32((hi - lo) * np.float32(2.0 ** -24))
    want = (words >> np.uint64(8)).astype(np.float32).T * scale24 + lo
    want = np.minimum(want, np.nextafter(hi, lo))
    # the device does it in one FMA; the f32 two-step above can differ by one ulp of 0.05
    assert np.abs(st - want).max() <= 4e-9
    assert (st >= lo).all() and (st < hi).all()
    env.close()
