"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes partition a batch, each runs the (CPU
oracle) forward on its shard, the gathered result must equal the single-process result.  The GPU path
uses the same partition / gather code with NCCL (bench.py --gpus N)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_partitions():
    from hvi_cidnet_b200.dist import shard_range
    for n in (0, 1, 5, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == n
            pos = 0
            for s, c in parts:
                assert s == pos and c >= 0
                pos += c
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hvi_cidnet_b200  # noqa: F401
    from hvi_cidnet_b200.dist import forward_sharded
    from oracle import cidnet_oracle as O
    torch.set_grad_enabled(False)
    torch.set_num_threads(2)
    sd = O.make_state_dict(3, True)
    x = O.make_input("uniform", 3, 16, 24, seed=8)          # ragged: 3 images over 2 ranks
    model = lambda t: O.forward(t, sd)                       # stands in for the CUDA module on CPU
    y = forward_sharded(model, x, gather=True)
    ref = O.forward(x, sd)
    ret[rank] = float((y - ref).abs().max())
    dist.destroy_process_group()


def test_forward_sharded_gloo_world2():
    world = 2
    port = 29000 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] <= 1e-6, ret[r]
