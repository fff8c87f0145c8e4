"""gymrs_step / gymrs_rollout / seeded gymrs_reset captured into a CUDA graph.

Kernel arguments are frozen at capture, but the auto-reset stream is keyed by the step index, so
a captured step must not reuse the epoch it was captured with: once a handle's step has been
captured the DEVICE counts its steps (BatchArgs::epoch_dev) and every replay is a new step.  The
property checked here: capture K steps, replay R times == K * R eager steps, bit for bit.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K, R = 8, 3


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


def make(g, kind, n, **kw):
    return {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind](num_envs=n, **kw)


def random_actions(torch, kind, steps, n, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    if kind == "pendulum":
        return (torch.rand((steps, n), device="cuda", generator=gen) * 4 - 2).contiguous()
    return torch.randint(0, 2 if kind == "cartpole" else 3, (steps, n), device="cuda", dtype=torch.int32,
                         generator=gen)


def outputs(env):
    env.sync()
    return [env.get_state(), env._t_obs.cpu().numpy().copy(), env._t_reward.cpu().numpy().copy(),
            env._t_done.cpu().numpy().copy(), env._t_truncated.cpu().numpy().copy()]


@pytest.mark.parametrize("kind,time_limit,config", [
    ("cartpole", False, None), ("cartpole", True, None), ("mountain_car", True, None), ("pendulum", False, None),
    # opt-in launch schemes are demoted for captured steps: chained launches (pdl = 2) to a grid-wide
    # dependency, the persistent TMA-staged kernel (vec = 8) to step_kernel; one lane per env stays
    ("cartpole", False, (8, 0, 2)), ("pendulum", True, (1, 64, 2)),
    # the high-occupancy build exists for host-counted steps only: a captured step of such a handle uses
    # the device-counted default build (4th entry = gymrs_set_launch_occupancy)
    ("mountain_car", False, (0, 0, 1, True))])
def test_replayed_graph_equals_eager_steps(torch, g, kind, time_limit, config):
    from gym_rs_b200 import _capi
    n = 200_000
    all_actions = random_actions(torch, kind, K * R + 4, n, 7)
    eager = make(g, kind, n, time_limit=time_limit)
    eager.reset(seed=11)
    env = make(g, kind, n, time_limit=time_limit)
    if config:
        env.set_launch_config(*config[:3])
        env.set_launch_occupancy(len(config) > 3 and config[3])
    env.reset(seed=11)
    # three eager steps first: the device counter must pick up where the host count stands
    for k in range(3):
        eager.step(all_actions[K * R + k], autoreset=True)
        env.step(all_actions[K * R + k], autoreset=True)

    side = torch.cuda.Stream()
    slots = torch.empty_like(all_actions[:K])
    env.sync()
    env.set_stream(side.cuda_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for k in range(K):
            env.step(slots[k], autoreset=True)
        # calls that synchronise cannot be recorded; they refuse without breaking the capture
        with pytest.raises(_capi.GymrsError, match="cannot be captured"):
            env.get_state()
    # capture records, it does not run: the handle is where the three eager steps left it
    assert np.array_equal(env.get_state(), eager.get_state())

    for r in range(R):
        slots.copy_(all_actions[r * K:(r + 1) * K])
        torch.cuda.synchronize()
        with torch.cuda.stream(side):      # replay on the handle's stream: env.sync() then covers it
            graph.replay()
        for k in range(K):
            eager.step(all_actions[r * K + k], autoreset=True)
        for a, b in zip(outputs(env), outputs(eager)):
            assert np.array_equal(a, b, equal_nan=True)

    # the host's view of the counter follows the device: checkpoint and eager steps carry on
    info = g.core.checkpoint_info(env.checkpoint())
    assert info["step_count"] == 3 + K * R and info["seed"] == 11
    with torch.cuda.stream(side):
        env.step(all_actions[K * R + 3], autoreset=True)
    eager.step(all_actions[K * R + 3], autoreset=True)
    for a, b in zip(outputs(env), outputs(eager)):
        assert np.array_equal(a, b, equal_nan=True)
    # host-buffer steps on a device-counted handle take their epoch from the device count too
    host_act = all_actions[0].cpu().numpy()
    bufs = [np.zeros((env.obs_dim, n), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)]
    env.step_host(host_act, *bufs, autoreset=True)
    eager.step(all_actions[0], autoreset=True)
    assert np.array_equal(bufs[0], outputs(eager)[1], equal_nan=True)
    twin = env.clone()
    twin.step(all_actions[1], autoreset=True)
    eager.step(all_actions[1], autoreset=True)
    assert np.array_equal(twin.get_state(), eager.get_state())
    for e in (env, eager, twin):
        e.close()


def test_captured_reset_and_rollout(torch, g):
    from gym_rs_b200 import _capi
    n, steps = 65_536, 16
    acts = random_actions(torch, "cartpole", steps, n, 3)
    env = make(g, "cartpole", n)
    env.reset(seed=5)
    side = torch.cuda.Stream()
    env.sync()
    env.set_stream(side.cuda_stream)
    obs_out = torch.empty((steps, 4, n), device="cuda")
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        env.reset(seed=5)                  # a whole episode batch per replay: reset + fused rollout
        env.rollout(acts, obs_out=obs_out, autoreset=True)
        with pytest.raises(_capi.GymrsError, match="seeded full"):
            env.reset()                    # entropy-seeded: the seed would be frozen into the graph
    ref = make(g, "cartpole", n)
    ref.reset(seed=5)
    ref_obs = torch.empty_like(obs_out)
    ref.rollout(acts, obs_out=ref_obs, autoreset=True)
    ref.sync()
    for _ in range(2):                     # every replay restarts from the same seeded reset
        obs_out.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(obs_out, ref_obs)
        assert np.array_equal(env.get_state(), ref.get_state())
    assert g.core.checkpoint_info(env.checkpoint())["step_count"] == steps
    env.close()
    ref.close()


def test_replay_after_reseeding_is_reported_not_silently_wrong(torch, g):
    from gym_rs_b200 import _capi
    n = 4096
    acts = random_actions(torch, "cartpole", 1, n, 1)
    env = make(g, "cartpole", n)
    env.reset(seed=1)
    side = torch.cuda.Stream()
    env.sync()
    env.set_stream(side.cuda_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        env.step(acts[0], autoreset=True)
    with torch.cuda.stream(side):
        graph.replay()
        env.sync()
        env.reset(seed=2)       # the captured step still holds seed 1's Philox keys
        graph.replay()
        with pytest.raises(_capi.GymrsError, match="different seed"):
            env.sync()
        env.reset(seed=1)
        graph.replay()
        env.sync()              # same seed again: fine
    env.close()


def test_replay_after_a_parameter_or_variant_change_is_reported(torch, g):
    """A captured step freezes what the host decided at capture: the folded parameter block and the kernel
    variant (an auto-reset step of a handle whose steps_beyond_terminated is all-None skips that row).
    Changing either afterwards bumps the handle's parameter generation; a stale replay raises a sticky
    error instead of silently stepping with the old constants / the wrong variant."""
    from gym_rs_b200 import _capi
    n = 4096
    acts = random_actions(torch, "cartpole", 2, n, 1)
    side = torch.cuda.Stream()

    def captured_step(env):
        env.sync()
        env.set_stream(side.cuda_stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            env.step(acts[0], autoreset=True)
        return graph

    # 1. gymrs_set_params after capture
    env = make(g, "cartpole", n)
    env.reset(seed=1)
    graph = captured_step(env)
    with torch.cuda.stream(side):
        graph.replay()
        env.sync()
        p = env.params
        p.gravity = 3.7
        env.params = p
        graph.replay()
        with pytest.raises(_capi.GymrsError, match="re-capture"):
            env.sync()
        fresh = captured_step(env)          # recorded under the new constants: fine
    with torch.cuda.stream(side):
        fresh.replay()
        env.sync()
    env.close()

    # 2. the first step WITHOUT auto-reset makes steps_beyond_terminated matter: the graph that recorded
    #    the variant without that row is stale; one recorded afterwards clears the row and matches eager steps
    env = make(g, "cartpole", n)
    ref = make(g, "cartpole", n)
    for e in (env, ref):
        e.reset(seed=2)
    graph = captured_step(env)
    with torch.cuda.stream(side):
        graph.replay()
        ref_out = None
    ref.step(acts[0], autoreset=True)
    with torch.cuda.stream(side):
        env.step(acts[1], autoreset=False)
        graph.replay()
        with pytest.raises(_capi.GymrsError, match="re-capture"):
            env.sync()
    env.close()
    ref.close()
    env = make(g, "cartpole", n)
    ref = make(g, "cartpole", n)
    for e in (env, ref):
        e.reset(seed=2)
        for _ in range(60):                 # without resets nearly every pole falls: sbt = Some(k) everywhere
            e.step(acts[1], autoreset=False)
    graph = captured_step(env)              # recorded with the steps_beyond_terminated row
    for _ in range(3):
        with torch.cuda.stream(side):
            graph.replay()
        ref.step(acts[0], autoreset=True)
    env.sync()
    ref.sync()
    assert np.array_equal(env.get_state(), ref.get_state())
    assert np.array_equal(env.steps_beyond_terminated.cpu().numpy(), ref.steps_beyond_terminated.cpu().numpy())
    assert np.array_equal(env._t_reward.cpu().numpy(), ref._t_reward.cpu().numpy())
    env.close()
    ref.close()
