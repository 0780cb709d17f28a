"""world_size-2 gloo test of the multi-GPU plumbing (CPU): env sharding + the optional action gather."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _worker(rank, world, port, num_envs, out_dir):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hyper-vla_b200"))
    import torch.distributed as dist
    from hvla import parallel as PL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = PL.shard_range(num_envs, rank, world)
    # each rank "acts" on its own shard: action[e] = e (stand-in for the GPU path, which needs no exchange)
    mine = torch.arange(lo, hi, dtype=torch.float32)[:, None, None].expand(hi - lo, 4, 7).contiguous()
    full = PL.gather_actions(mine, num_envs)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("num_envs", [10, 7])
def test_env_sharding_and_action_gather_world2(tmp_path, num_envs):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + num_envs
    mp.spawn(_worker, args=(2, port, num_envs, str(tmp_path)), nprocs=2, join=True)
    expect = np.broadcast_to(np.arange(num_envs, dtype=np.float32)[:, None, None], (num_envs, 4, 7))
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), expect)
