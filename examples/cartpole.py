"""BASELINE config 1 (plumbing, one env): random-action CartPole episodes through the GPU-backed Env
surface -- the protocol of the reference's example program (15 episodes of at most 475 steps, a reset
after each), with RenderMode.NONE because the B200 path draws nothing."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_rs_b200.envs.classical_control.cartpole import CartPoleEnv  # noqa: E402
from gym_rs_b200.utils.renderer import RenderMode  # noqa: E402

EPISODES, MAX_STEPS = 15, 475


def run_episode(env, rng):
    total = 0.0
    for _ in range(MAX_STEPS):
        outcome = env.step(rng.randrange(2))
        total += outcome.reward
        if outcome.done:
            break
    return total


def main():
    rng = random.Random()
    env = CartPoleEnv(RenderMode.NONE)
    returns = []
    for _ in range(EPISODES):
        env.reset(None, False, None)
        returns.append(run_episode(env, rng))
    print("episode returns:", returns)
    env.close()


if __name__ == "__main__":
    main()
