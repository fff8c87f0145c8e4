"""A small workload touching every kernel (step_kernel, step_stream_kernel, rollout_kernel,
reset_kernel; plain, chained and host paths) -- sized so that compute-sanitizer's slow tools
(racecheck, initcheck, synccheck) finish in seconds.  Usage:
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_rs_b200 as g  # noqa: E402

n = 8192 + 6  # not a multiple of 4: the last thread takes the scalar out-of-line path
gen = torch.Generator(device="cuda").manual_seed(0)
for cls, hi in ((g.CartPoleEnv, 2), (g.MountainCarEnv, 3), (g.PendulumEnv, 0)):
    # last entry: the high-occupancy build (gymrs_set_launch_occupancy), plain and chained
    for vec, pdl, wide in ((0, 1, False), (4, 2, False), (8, 2, False), (8, 0, False), (1, 2, False), (0, 1, True), (4, 2, True)):
        env = cls(num_envs=n, time_limit=(vec == 4))
        env.set_launch_config(vec=vec, block=0, pdl=pdl)
        env.set_launch_occupancy(wide)
        env.reset(seed=1)
        if hi:
            acts = [torch.randint(0, hi, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(4)]
        else:
            acts = [torch.rand((n,), generator=gen, device="cuda") * 4 - 2 for _ in range(4)]
        for t in range(12):
            env.step(acts[t % 4], autoreset=True)
        env.rollout(torch.stack(acts), autoreset=True)
        for t in range(3):
            env.step(acts[t % 4], autoreset=False)
        h = [a.cpu() for a in acts[:2]]
        obs = np.empty((env.obs_dim, n), dtype=np.float32)
        rew = np.empty(n, dtype=np.float32)
        done = np.empty(n, dtype=np.uint8)
        env.step_host(h[0], obs, rew, done, None, autoreset=True)
        env.sync()
        assert np.isfinite(env.get_state()).all()
        # far-out states: whole threads leave the straight-line path (out-of-line scalar steps, libm trig)
        st = env.get_state()
        st[0, ::7] = 3000.0
        st[-1, ::5] = -50.0
        env.set_state(st)
        env.step(acts[0], autoreset=False)
        env.rollout(torch.stack(acts), autoreset=False)
        # pipelined host loop with the compact wire formats (widen / pack kernels)
        S = 2
        hobs = torch.empty((S, env.obs_dim, n)).pin_memory()
        hrew = torch.empty((S, n)).pin_memory()
        if hi:
            hbits = torch.empty((S, (n + 7) // 8), dtype=torch.uint8).pin_memory()
            hact = torch.stack(h).to(torch.uint8).pin_memory()
            env.rollout_host(hact, hobs, hrew, hbits, None, n_steps=5, u8_actions=True, packed_done=True)
        else:
            hdone = torch.empty((S, n), dtype=torch.uint8).pin_memory()
            env.rollout_host(torch.stack(h).pin_memory(), hobs, hrew, hdone, None, n_steps=5)
        env.sync()
        env.close()
# gymrs_step_pass: a two-stream pass bracketed by events
import ctypes as C  # noqa: E402
from gym_rs_b200 import _capi  # noqa: E402
_L = _capi.load()
_s1, _s2 = torch.cuda.Stream(), torch.cuda.Stream()
_envs = [g.MountainCarEnv(num_envs=n, global_env_offset=i * n) for i in range(2)]
for _i, _e in enumerate(_envs):
    _e.reset(seed=2)
    _e.sync()
    _e.set_stream((_s1 if _i == 0 else _s2).cuda_stream)
_a = torch.randint(0, 3, (n,), generator=gen, device="cuda", dtype=torch.int32)
_b, _d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_b.record(_s1)
_d.record(_s1)
torch.cuda.synchronize()
_hs = (C.c_void_p * 4)(*[_envs[i % 2].handle.value for i in range(4)])
_ap = (C.c_void_p * 4)(*[_a.data_ptr()] * 4)
_capi.check(_L.gymrs_step_pass(_hs, _ap, 4, _capi.STEP_AUTORESET, C.c_void_p(_b.cuda_event), C.c_void_p(_d.cuda_event), None))
_d.synchronize()
for _e in _envs:
    _e.sync()
    _e.close()
# device-counted kernel variants (CUDA-graph capture): captured steps + rollout + seeded reset,
# replayed, then an eager step, a host step, a checkpoint round trip and a clone on the same handle
side = torch.cuda.Stream()
for cls, hi in ((g.CartPoleEnv, 2), (g.PendulumEnv, 0)):
    env = cls(num_envs=n, time_limit=True)
    env.reset(seed=3)
    if hi:
        acts = torch.randint(0, hi, (4, n), generator=gen, device="cuda", dtype=torch.int32)
    else:
        acts = torch.rand((4, n), generator=gen, device="cuda") * 4 - 2
    env.sync()
    env.set_stream(side.cuda_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        env.reset(seed=3)
        for t in range(4):
            env.step(acts[t], autoreset=True)
        env.rollout(acts, autoreset=True)
    with torch.cuda.stream(side):  # the handle's stream
        for _ in range(3):
            graph.replay()
    env.step(acts[0], autoreset=True)
    obs = np.empty((env.obs_dim, n), dtype=np.float32)
    rew = np.empty(n, dtype=np.float32)
    done = np.empty(n, dtype=np.uint8)
    env.step_host(acts[1].cpu(), obs, rew, done, None, autoreset=True)
    blob = env.checkpoint()
    count = g.core.checkpoint_info(blob)["step_count"]
    assert count == 10, count
    twin = cls.from_checkpoint(blob)
    other = env.clone()
    twin.step(acts[2], autoreset=True)
    other.step(acts[2], autoreset=True)
    assert np.array_equal(twin.get_state(), other.get_state())
    for e in (env, twin, other):
        e.close()
print("sanitize_small ok")
