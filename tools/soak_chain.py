"""Soak test of chained (pdl = 2) launches: many thousands of back-to-back steps on interleaved
handles, compared bit for bit with plain stream-ordered launches (a rare ordering bug would show
up as a state mismatch).  Usage: python tools/soak_chain.py [steps] [vec]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_rs_b200 as g  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
vec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << 20
gen = torch.Generator(device="cuda").manual_seed(3)
acts = [torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(8)]
torch.cuda.synchronize()
finals = {}
for pdl in (0, 2):
    envs = [g.CartPoleEnv(num_envs=n, global_env_offset=k * n) for k in range(2)]
    for e in envs:
        e.set_launch_config(vec=vec, block=0, pdl=pdl)
        e.reset(seed=5)
    for t in range(steps):
        for k, e in enumerate(envs):
            e.step(acts[(t + 3 * k) % 8], autoreset=True)
    for e in envs:
        e.sync()
    finals[pdl] = [e.get_state() for e in envs]
    for e in envs:
        e.close()
ok = all(np.array_equal(a, b) for a, b in zip(finals[0], finals[2]))
print("soak", steps, "steps x 2 handles, vec", vec, "->", "IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
