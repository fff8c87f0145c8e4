"""Checkpoint / resume (`Env: Clone + Serialize`, core.rs:25).

CPU part: the blob format is pinned by building one by hand (header + payload + checksum restated
in Python) and having the library validate it -- no device needed.  GPU part: a handle restored
from a blob continues bit-identically to the handle that was saved, auto-resets, time limit and
steps_beyond_terminated included.
"""
import struct

import numpy as np
import pytest

MASK = (1 << 64) - 1


def blob_hash(buf: bytes, skip_off: int) -> int:
    """the library's checksum: multiply-xorshift over little-endian 8-byte words"""
    h = 0x9E3779B97F4A7C15 ^ len(buf)
    for i in range(0, len(buf) - 7, 8):
        w = 0 if i == skip_off else int.from_bytes(buf[i:i + 8], "little")
        h = ((h ^ w) * 0xFF51AFD7ED558CCD) & MASK
        h ^= h >> 29
    return h


def pad8(b):
    return (b + 7) & ~7


def handmade_cartpole_blob(n=5, seed=77, step_count=9, global_off=1000):
    """256-byte header + state[4][n] f32 + reward[n] f32 + done[n] + truncated[n] + sbt[n] i32"""
    params = struct.pack("<8d2i", 9.8, 1.0, 0.1, 0.5, 10.0, 0.02, 0.20943951023931953, 2.4, 0, 500)
    payload = b""
    for arr in (np.arange(4 * n, dtype=np.float32), np.ones(n, np.float32), np.zeros(n, np.uint8),
                np.zeros(n, np.uint8), np.full(n, -1, np.int32)):
        raw = arr.tobytes()
        payload += raw + b"\0" * (pad8(len(raw)) - len(raw))
    total = 256 + len(payload)
    head = struct.pack("<8sIIQQIIQQQQ", b"GYMRSCKP", 1, 0, n, global_off, 0, 0, seed, step_count, total, 0)
    head += struct.pack("<8f", *([-0.05] * 4 + [0.05] * 4)) + params.ljust(96, b"\0") + b"\0" * 56
    assert len(head) == 256
    blob = bytearray(head + payload)
    blob[64:72] = blob_hash(bytes(blob), 64).to_bytes(8, "little")
    return bytes(blob)


def test_blob_format_is_pinned_and_validated_on_the_host():
    from gym_rs_b200 import _capi
    from gym_rs_b200.core import checkpoint_info
    blob = handmade_cartpole_blob()
    info = checkpoint_info(blob)
    assert info == {"kind": 0, "flags": 0, "num_envs": 5, "global_env_offset": 1000, "seed": 77,
                    "step_count": 9, "bytes": len(blob)}
    # trailing bytes after the blob are ignored; any damage inside it is caught
    assert checkpoint_info(blob + b"xx")["num_envs"] == 5
    for pos in (3, 20, 70, 100, 200, 260, len(blob) - 1):
        bad = bytearray(blob)
        bad[pos] ^= 0x40
        with pytest.raises(_capi.GymrsError):
            checkpoint_info(bytes(bad))
    with pytest.raises(_capi.GymrsError, match="truncated"):
        checkpoint_info(blob[:-8])
    with pytest.raises(_capi.GymrsError, match="shorter"):
        checkpoint_info(blob[:100])
    # without a device a blob cannot become a handle: there is no CPU path
    import ctypes as C
    L = _capi.load()
    if L.gymrs_device_count() == 0:
        h = C.c_void_p()
        buf = np.frombuffer(blob, dtype=np.uint8)
        rc = L.gymrs_checkpoint_create(buf.ctypes.data_as(C.c_void_p), buf.nbytes, 0, C.byref(h))
        assert rc == _capi.ERR_NO_DEVICE and not h.value


# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def g():
    import gym_rs_b200
    return gym_rs_b200


def _actions(torch, g, kind, steps, n, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    if kind == "pendulum":
        return (torch.rand((steps, n), device="cuda", generator=gen) * 4 - 2).contiguous()
    hi = 2 if kind == "cartpole" else 3
    return torch.randint(0, hi, (steps, n), device="cuda", dtype=torch.int32, generator=gen)


def _snapshot(env):
    env.sync()
    out = [env.get_state(), env._t_obs.cpu().numpy().copy(), env._t_reward.cpu().numpy().copy(),
           env._t_done.cpu().numpy().copy(), env._t_truncated.cpu().numpy().copy()]
    if env._t_sbt is not None:
        out.append(env._t_sbt.cpu().numpy().copy())
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind,time_limit", [("cartpole", False), ("cartpole", True), ("mountain_car", True),
                                             ("pendulum", False), ("pendulum", True)])
def test_restored_handle_continues_bit_identically(torch, g, kind, time_limit):
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind]
    n = 100_003  # ragged on purpose
    env = cls(num_envs=n, time_limit=time_limit, global_env_offset=12345)
    env.reset(seed=2024)
    acts = _actions(torch, g, kind, 90, n, 1)
    for k in range(40):
        env.step(acts[k], autoreset=True)
    blob = env.checkpoint()
    info = g.core.checkpoint_info(blob)
    assert info["num_envs"] == n and info["seed"] == 2024 and info["step_count"] == 40
    assert info["global_env_offset"] == 12345 and info["flags"] == (1 if time_limit else 0)
    at_save = _snapshot(env)

    for k in range(40, 90):
        env.step(acts[k], autoreset=True)
    want = _snapshot(env)

    # (1) a new handle from the blob alone
    twin = cls.from_checkpoint(blob)
    for a, b in zip(_snapshot(twin), at_save):
        assert np.array_equal(a, b, equal_nan=True)
    for k in range(40, 90):
        twin.step(acts[k], autoreset=True)
    for a, b in zip(_snapshot(twin), want):
        assert np.array_equal(a, b, equal_nan=True)
    # the fused rollout continues from a checkpoint the same way
    third = cls.from_checkpoint(blob)
    third.rollout(acts[40:90].contiguous(), autoreset=True)
    assert np.array_equal(third.get_state(), want[0])

    # (2) load into an existing, unrelated handle
    other = cls(num_envs=n, time_limit=time_limit)
    other.reset(seed=1)
    other.restore(blob)
    for k in range(40, 90):
        other.step(acts[k], autoreset=True)
    for a, b in zip(_snapshot(other), want):
        assert np.array_equal(a, b, equal_nan=True)
    for e in (env, twin, third, other):
        e.close()


@pytest.mark.gpu
def test_checkpoint_keeps_steps_beyond_terminated_and_mutated_fields(torch, g):
    n = 4096
    env = g.CartPoleEnv(num_envs=n)
    p = env.params
    p.force_mag = 7.5
    p.kinematics_integrator = 1
    env.params = p
    env.reset(seed=3)
    ones = torch.ones(n, device="cuda", dtype=torch.int32)
    for _ in range(60):  # no auto-reset: every pole falls, steps_beyond_terminated counts up
        env.step(ones)
    blob = env.checkpoint()
    twin = g.CartPoleEnv.from_checkpoint(blob)
    assert twin.params.force_mag == 7.5 and twin.params.kinematics_integrator == 1
    _, sbt = twin.get_state(with_sbt=True)
    assert (sbt >= 0).all() and sbt.max() > 10
    for _ in range(5):
        env.step(ones)
        twin.step(ones)
    for a, b in zip(_snapshot(env), _snapshot(twin)):
        assert np.array_equal(a, b, equal_nan=True)
    assert float(twin._t_reward.max()) == 0.0  # cartpole.rs:455-464: 0.0 once past termination
    env.close()
    twin.close()


@pytest.mark.gpu
def test_checkpoint_load_rejects_mismatched_handles(torch, g):
    from gym_rs_b200 import _capi
    env = g.CartPoleEnv(num_envs=1000)
    blob = env.checkpoint()
    for other in (g.CartPoleEnv(num_envs=999), g.MountainCarEnv(num_envs=1000),
                  g.CartPoleEnv(num_envs=1000, time_limit=True)):
        with pytest.raises(_capi.GymrsError):
            other.restore(blob)
        other.close()
    bad = blob.copy()
    bad[300] ^= 1
    with pytest.raises(_capi.GymrsError, match="checksum"):
        env.restore(bad)
    with pytest.raises(ValueError):
        g.MountainCarEnv.from_checkpoint(blob)
    env.close()
