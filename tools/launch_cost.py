"""Host cost of one gymrs_step call (ctypes + library + cudaLaunchKernelEx), measured on a tiny batch so
the GPU is never the bottleneck, and the same through a C loop-free path for comparison."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gym_rs_b200 as g
from gym_rs_b200 import _capi

L = _capi.load()
for n in (1024, 1 << 20):
    env = g.CartPoleEnv(num_envs=n)
    env.reset(seed=0)
    a = torch.zeros(n, dtype=torch.int32, device="cuda")
    h, p = env.handle, a.data_ptr()
    for pdl in (0, 1, 2):
        env.set_launch_config(0, 0, pdl)
        for _ in range(200):
            L.gymrs_step(h, p, 1)
        torch.cuda.synchronize()
        k = 400 if n == 1024 else 4000
        t0 = time.perf_counter()
        for _ in range(k):
            L.gymrs_step(h, p, 1)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"n={n:8d} pdl={pdl}: host {1e6 * (t1 - t0) / k:5.2f} us/call, until drained {1e6 * (t2 - t0) / k:5.2f} us/step")
    env.close()
