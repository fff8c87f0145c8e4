"""Experiment: us/step of one CartPole/MountainCar/Pendulum step under different launch schemes on ONE stream.
   pool = 'ring' (16 cold 1M-env batches round-robin) or 'resident' (one batch, L2-resident)
   scheme = eager pdl 0/1/2, or a CUDA graph of `span` captured steps (device-counted, pdl 1 edges)
Usage: python tools/exp_modes.py [env] [n_envs]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import gym_rs_b200 as g
from gym_rs_b200 import _capi

env_name = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[env_name]
L = _capi.load()
RING, STEPS, REPS = 16, 1600, 7
stream = torch.cuda.Stream()


def actions(k):
    gen = torch.Generator(device="cuda").manual_seed(k)
    if env_name == "pendulum":
        return (torch.rand(n, device="cuda", generator=gen) * 4 - 2).contiguous()
    return torch.randint(0, 2 if env_name == "cartpole" else 3, (n,), device="cuda", dtype=torch.int32, generator=gen)


def build(count, pdl):
    envs = []
    for i in range(count):
        e = cls(num_envs=n, global_env_offset=i * n)
        e.reset(seed=0)
        e.sync()
        e.set_stream(stream.cuda_stream)
        e.set_launch_config(0, 0, pdl)
        envs.append((e, [actions(100 * i + v) for v in range(4)]))
    return envs


def time_region(fn):
    ts = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / STEPS)
    return statistics.median(ts)


for pool_name, count in (("ring", RING), ("resident", 1)):
    for pdl in (0, 1, 2):
        envs = build(count, pdl)
        sched = [(e.handle, a[v].data_ptr()) for v in range(4) for e, a in envs]

        def eager():
            for i in range(STEPS):
                h, a = sched[i % len(sched)]
                L.gymrs_step(h, a, 1)
        for _ in range(2):
            eager()  # burn-in
        print(f"{env_name} {pool_name:8s} eager pdl={pdl}: {time_region(eager):6.2f} us/step", flush=True)
        if pdl == 1:
            for span in (len(sched), 4 * len(sched)):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    for i in range(span):
                        h, a = sched[i % len(sched)]
                        _capi.check(L.gymrs_step(h, a, 1))

                def replay():
                    with torch.cuda.stream(stream):
                        for _ in range(STEPS // span):
                            graph.replay()
                replay()
                us = time_region(replay) * STEPS / (STEPS // span * span)
                print(f"{env_name} {pool_name:8s} graph of {span:3d} steps:   {us:6.2f} us/step", flush=True)
                del graph
            # the handles are device-counted now: the same eager loop runs the DEVC kernel variants
            print(f"{env_name} {pool_name:8s} eager pdl=1, device-counted kernels: {time_region(eager):6.2f} us/step", flush=True)
            for e, _ in envs:
                e.set_launch_config(0, 0, 0)
            print(f"{env_name} {pool_name:8s} eager pdl=0, device-counted kernels: {time_region(eager):6.2f} us/step", flush=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for i in range(len(sched)):
                    _capi.check(L.gymrs_step(sched[i][0], sched[i][1], 1))

            def replay0():
                with torch.cuda.stream(stream):
                    for _ in range(STEPS // len(sched)):
                        graph.replay()
            replay0()
            print(f"{env_name} {pool_name:8s} graph captured with pdl=0:            {time_region(replay0) * STEPS / (STEPS // len(sched) * len(sched)):6.2f} us/step", flush=True)
            del graph
        for e, _ in envs:
            e.close()
