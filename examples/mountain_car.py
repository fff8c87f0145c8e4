"""One MountainCar-v0 env driven with random actions: stop at the goal or after 200 steps, then keep
stepping past the end of the episode (the env has no truncation and never refuses a step), as the
reference's example program does -- here with RenderMode.NONE."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_rs_b200.envs.classical_control.mountain_car import MountainCarEnv  # noqa: E402
from gym_rs_b200.utils.renderer import RenderMode  # noqa: E402


def main():
    rng = random.Random()
    car = MountainCarEnv(RenderMode.NONE)
    car.reset(None, False, None)
    steps, reached_goal = 0, False
    while not reached_goal and steps <= 200:
        reached_goal = car.step(rng.randrange(3)).done
        steps += 1
    print("first phase:", steps, "steps, goal reached:", reached_goal)
    total_reward = 0.0
    for _ in range(200):
        total_reward += car.step(rng.randrange(3)).reward
        steps += 1
    print("after", steps, "steps; reward of the last 200:", total_reward)
    car.close()


if __name__ == "__main__":
    main()
