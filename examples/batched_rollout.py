"""The shape the GPU path is built for: 1,048,576 CartPole instances per handle, random actions,
auto-reset, (a) one launch per step with the results left on the device, (b) 64 steps fused into one
launch, (c) host buffers with pipelined delivery.  Prints env-steps/s for each."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_rs_b200 as g  # noqa: E402

N, STEPS = 1 << 20, 256


def main():
    env = g.CartPoleEnv(num_envs=N)
    env.reset(seed=0)
    actions = torch.randint(0, 2, (64, N), device="cuda", dtype=torch.int32)

    env.step(actions[0], autoreset=True)
    env.sync()
    t0 = time.perf_counter()
    for k in range(STEPS):
        out = env.step(actions[k % 64], autoreset=True)     # ActionReward of zero-copy device views
    env.sync()
    dt = time.perf_counter() - t0
    print(f"gymrs_step    : {N * STEPS / dt / 1e9:7.1f} G env-steps/s, mean reward {float(out.reward.mean()):.3f}")

    obs = torch.empty((64, 4, N), device="cuda")
    env.rollout(actions, obs_out=obs, autoreset=True)
    env.sync()
    t0 = time.perf_counter()
    for _ in range(STEPS // 64):
        env.rollout(actions, obs_out=obs, autoreset=True)
    env.sync()
    dt = time.perf_counter() - t0
    print(f"gymrs_rollout : {N * STEPS / dt / 1e9:7.1f} G env-steps/s")



    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory()
    bufs = [(actions[i].cpu().pin_memory(), pinned((4, N), torch.float32), pinned((N,), torch.float32),
             pinned((N,), torch.uint8)) for i in range(2)]

    def host_steps(count):
        tickets = [None, None]
        for k in range(count):
            slot = k & 1
            if tickets[slot] is not None:
                env.host_wait(tickets[slot])                 # results of step k - 2 are in bufs[slot]
            a, o, r, d = bufs[slot]
            tickets[slot] = env.step_host_async(a, o, r, d, None, autoreset=True)
        for t in tickets:
            if t is not None:
                env.host_wait(t)

    host_steps(8)                                            # first use sets up staging buffers and events
    t0 = time.perf_counter()
    host_steps(128)
    dt = time.perf_counter() - t0
    print(f"host buffers  : {N * 128 / dt / 1e9:7.2f} G env-steps/s (PCIe-bound: 25 B per env-step)")
    env.close()


if __name__ == "__main__":
    main()
