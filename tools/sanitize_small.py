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

n = 8192 + 4
gen = torch.Generator(device="cuda").manual_seed(0)
for cls, hi in ((g.CartPoleEnv, 2), (g.MountainCarEnv, 3), (g.PendulumEnv, 0)):
    for vec, pdl in ((0, 1), (4, 2), (8, 2), (8, 0), (1, 2)):
        env = cls(num_envs=n, time_limit=(vec == 4))
        env.set_launch_config(vec=vec, block=0, pdl=pdl)
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
        env.close()
print("sanitize_small ok")
