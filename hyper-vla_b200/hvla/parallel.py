"""Multi-GPU plumbing: environments shard across ranks, parameters are replicated, and the
inference path has NO collective.  The only exchange is optional: gathering the
(B/n, 4, 7) actions when one host consumer needs all of them (reference equivalent: none -- the
reference's data parallelism exists only in training, scripts/train.py:405, 460)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(num_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of the env axis owned by `rank`; remainders go to the low ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(num_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_task_table(task_index: np.ndarray, lo: int, hi: int):
    """Tasks used by envs [lo, hi): (sorted unique global task ids, local index per env).  Each rank
    generates weights only for the tasks its envs use instead of receiving a broadcast."""
    ti = np.asarray(task_index)[lo:hi]
    uniq, local = np.unique(ti, return_inverse=True)
    return uniq.astype(np.int32), local.astype(np.int32)


def gather_actions(actions, num_envs: int, group=None):
    """All-gather per-rank action shards (torch tensor (b_r, 4, 7), ragged allowed) into (num_envs, 4, 7)
    on every rank.  NCCL on GPU tensors, gloo on CPU tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(num_envs, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(actions.shape[1:]), dtype=actions.dtype, device=actions.device)
    pad[: actions.shape[0]] = actions
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    assert actions.shape[0] == sizes[rank][1] - sizes[rank][0]
    return torch.cat([out[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
