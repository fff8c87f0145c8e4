"""Multi-GPU helpers: the env batch shards trivially (no env reads another env's state,
cartpole.rs:408-448, mountain_car.rs:408-425), so the step path needs NO collective.

* `shard_range` -- contiguous global env-id range of a rank; reset sampling is keyed by the global
  id, so per-env results do not depend on the number of shards.
* `device_for_rank` -- which GPU of the box a rank drives (ranks are spread over the box's PCIe root complexes).
* `make_sharded_env` -- one handle per process/GPU holding this rank's range.
* `gather_observations` -- the only (optional) collective: an NCCL all-gather of the observation
  rows for callers that want one contiguous [obs_dim, total_envs] view on every rank
  (SURVEY.md section 8e).  Not on the step path; bench.py never calls it inside a timed region.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(rank: int, world_size: int, total_envs: int) -> Tuple[int, int]:
    """[begin, end) of global env ids owned by `rank`: sizes differ by at most one."""
    if not (0 <= rank < world_size) or total_envs < 0:
        raise ValueError("bad rank / world_size / total_envs")
    base, extra = divmod(total_envs, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def device_for_rank(local_rank: int, local_world_size: int, visible_devices: int, spread: bool = True) -> int:
    """CUDA device ordinal of a rank on one box.  With fewer ranks than visible GPUs the ranks are spread over
    the whole box (rank r -> GPU r * (visible // ranks)) instead of packed onto GPUs 0..ranks-1: the step path
    itself does not care, but host delivery does -- neighbouring GPUs usually hang off the same PCIe root complex
    / socket and share its device-to-host write bandwidth (measured on an 8 x B200 host, plain pinned-memory
    copies: GPUs {0, 1} 72 GB/s, {0, 4} 96 GB/s; {0, 1, 2, 3} 76 GB/s, {0, 2, 4, 6} 114 GB/s;
    tools/pcie_probe_multi.py).  `spread=False`, or as many ranks as GPUs, gives the identity mapping."""
    if not (0 <= local_rank < local_world_size) or visible_devices < 1:
        raise ValueError("bad local_rank / local_world_size / visible_devices")
    if local_world_size > visible_devices:
        raise ValueError("more ranks than visible GPUs")
    stride = visible_devices // local_world_size if spread else 1
    return local_rank * stride


def make_sharded_env(env_cls, total_envs: int, rank: int, world_size: int, device: int, **kw):
    """Create this rank's shard of a `total_envs`-instance batch on `device`."""
    begin, end = shard_range(rank, world_size, total_envs)
    if end == begin:
        raise ValueError("more ranks than env instances")
    return env_cls(num_envs=end - begin, device=device, global_env_offset=begin, **kw)


def gather_observations(env, total_envs: int, out=None):
    """All-gather the observation rows of every rank's shard into [obs_dim, total_envs] (same on all
    ranks).  Requires an initialised torch.distributed process group (NCCL) and equal shard sizes
    or sizes given by `shard_range`."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    obs = env._t_obs  # [obs_dim, n_local] view of the handle's device rows (row stride ld)
    if out is None:
        out = torch.empty((env.obs_dim, total_envs), dtype=obs.dtype, device=obs.device)
    sizes = [shard_range(r, world, total_envs) for r in range(world)]
    assert sizes[rank][1] - sizes[rank][0] == env.num_envs, "handle does not hold this rank's shard_range"
    env.sync()
    if all(e - b == sizes[0][1] - sizes[0][0] for b, e in sizes):
        # equal shards: one all_gather_into_tensor per observation row (rows are contiguous in `out`
        # only per rank, so gather into [world, n_local] and copy each row into place)
        tmp = torch.empty((world, env.obs_dim, env.num_envs), dtype=obs.dtype, device=obs.device)
        dist.all_gather_into_tensor(tmp, obs.contiguous())
        for r, (b, e) in enumerate(sizes):
            out[:, b:e] = tmp[r]
    else:
        parts = [torch.empty((env.obs_dim, e - b), dtype=obs.dtype, device=obs.device) for b, e in sizes]
        dist.all_gather(parts, obs.contiguous())
        for (b, e), p in zip(sizes, parts):
            out[:, b:e] = p
    return out
