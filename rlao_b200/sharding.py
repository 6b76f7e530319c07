"""Multi-GPU plumbing: environments are independent, so they are sharded across ranks (one process per GPU) with
no collective inside `env.step`; NCCL (NVLink 5 / NVSwitch) is used only between rollouts, to reduce rollout
statistics and to gather the transition windows sampled for the PO4AO dynamics update
(MAIN_CODE/PO4AO/mbrl.py:94-142 trains on `replay.sample_contiguous` windows)."""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun environment."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_envs(total_envs, rank=None, world_size=None):
    """Contiguous block partition of `total_envs` environments: returns (n_local, env_offset).  env_offset is what
    `OOPAO.set_params(..., env_offset=...)` needs to keep the Philox streams of different ranks disjoint."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(total_envs, world_size)
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return n_local, offset


def reduce_rollout_stats(strehl_sum, reward_sum, count):
    """Sum of per-rank rollout statistics (tensors or floats) -> (mean Strehl, mean reward, total count)."""
    t = torch.as_tensor([float(strehl_sum), float(reward_sum), float(count)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        t = t.cpu()
    return float(t[0] / t[2]), float(t[1] / t[2]), int(t[2])


def gather_windows(*tensors):
    """All-gather of equally shaped per-rank tensors along dim 0 (sampled transition windows: states, actions,
    rewards, next states).  Returns a tuple of tensors of world_size x the local leading dimension."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tuple(tensors)
    out = []
    for t in tensors:
        parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, t.contiguous())
        out.append(torch.cat(parts, dim=0))
    return tuple(out)
