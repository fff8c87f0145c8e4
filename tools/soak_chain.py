"""Soak test of chained (pdl = 2) launches: many thousands of back-to-back steps on interleaved
handles, compared bit for bit with plain stream-ordered launches (a rare ordering bug would show
up as a state mismatch).  Usage: python tools/soak_chain.py [steps] [vec] [wide] [env]
wide = 1: the chained run uses the high-occupancy build (every CTA of a step resident at once, so every CTA of
the next step spins on its predecessor's flag: the hardest case for the per-CTA release / acquire protocol);
env = cartpole | mountain_car | pendulum."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_rs_b200 as g  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
vec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
wide = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
env_name = sys.argv[4] if len(sys.argv) > 4 else "cartpole"
cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[env_name]
n = 1 << 20
gen = torch.Generator(device="cuda").manual_seed(3)
if env_name == "pendulum":
    acts = [torch.rand((n,), generator=gen, device="cuda") * 4 - 2 for _ in range(8)]
else:
    acts = [torch.randint(0, 2 if env_name == "cartpole" else 3, (n,), generator=gen, device="cuda", dtype=torch.int32)
            for _ in range(8)]
torch.cuda.synchronize()
finals = {}
for pdl in (0, 2):
    envs = [cls(num_envs=n, global_env_offset=k * n) for k in range(2)]
    for e in envs:
        e.set_launch_config(vec=vec, block=0, pdl=pdl)
        e.set_launch_occupancy(wide and pdl == 2)
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
print("soak", env_name, steps, "steps x 2 handles, vec", vec, "wide", int(wide), "->", "IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
