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


# ---------------------------------------------------------------------------------------------
# Row-strip sharding of ONE image (cfg 5): the library's exchange schedule, dry-run on CPU.
# cidnet_forward_sharded_dry walks the same schedule as the CUDA forward (no kernels) and calls the
# halo / all-reduce callbacks with pointers into a host workspace; StripComm moves the rows over gloo.
# ---------------------------------------------------------------------------------------------
def test_strip_plan_covers_image():
    import hvi_cidnet_b200  # noqa: F401
    from hvi_cidnet_b200.dist import strip_plan, strip_local_range
    from hvi_cidnet_b200 import _lib
    for H, world in ((2160, 8), (2160, 1), (64, 3), (400, 2), (640, 4)):
        pos = 0
        for r in range(world):
            sh = strip_plan(H, world, r, 16)
            assert sh.row_begin == pos and sh.row_begin % 8 == 0 and sh.row_end % 8 == 0
            a, b = strip_local_range(sh)
            assert a == max(0, sh.row_begin - (16 if world > 1 else 0)) and b == min(H, sh.row_end + (16 if world > 1 else 0))
            import ctypes
            assert _lib.lib().cidnet_shard_local_rows(ctypes.byref(sh)) == b - a
            pos = sh.row_end
        assert pos == H
    with pytest.raises(_lib.CidnetError):
        strip_plan(64, 16, 0, 16)            # more ranks than coarsest rows
    with pytest.raises(_lib.CidnetError):
        strip_plan(64, 4, 0, 32)             # halo larger than the owned rows


def _pattern(rank, nfloats):
    return (torch.arange(nfloats, dtype=torch.float32) % 997.0) + 1000.0 * (rank + 1)


def _dry_worker(rank, world, port, H, W, ret, variant=0):
    import ctypes
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hvi_cidnet_b200  # noqa: F401
    from hvi_cidnet_b200 import _lib
    from hvi_cidnet_b200.dist import StripComm, strip_plan
    L = _lib.lib()
    sh = strip_plan(H, world, rank, 16)
    rows = L.cidnet_shard_local_rows(ctypes.byref(sh))
    nbytes = L.cidnet_workspace_bytes(1, rows, W)
    raw = torch.zeros(nbytes + 2048, dtype=torch.uint8)
    off = (-raw.data_ptr()) % 1024
    ws = raw[off:off + (nbytes // 4) * 4]
    ws.view(torch.float32).copy_(_pattern(rank, ws.numel() // 4))
    comm = StripComm(ws)
    comm.trace = []
    nh, na = ctypes.c_int(), ctypes.c_int()
    rc = L.cidnet_forward_sharded_dry_variant(variant, W, ctypes.byref(sh), ws.data_ptr(), ws.numel(), comm.halo_cb,
                                              comm.allreduce_cb, None, ctypes.byref(nh), ctypes.byref(na))
    assert rc == 0 and comm.error is None, (rc, comm.error, L.cidnet_last_error())
    logs, traces = [None] * world, [None] * world
    dist.all_gather_object(logs, comm.log)
    dist.all_gather_object(traces, comm.trace)
    bad = 0
    kinds = [[e[0] for e in lg] for lg in logs]
    assert all(k == kinds[0] for k in kinds), "ranks disagree on the callback sequence"
    # The workspace is re-used by liveness (a later tensor may occupy an earlier one's bytes), so every exchange is checked
    # AT THE TIME IT HAPPENS: the block a rank received from a neighbour must be the block that neighbour sent in the same
    # call, the geometry of the request must mirror the neighbour's, and an all-reduce must leave the sum of all ranks' inputs.
    for i, e in enumerate(comm.log):
        tr = comm.trace[i]
        if e[0] == "halo":
            assert all(len(lg[i][1]) == len(e[1]) for lg in logs)
            for j, (o, rb, nr, top, bot) in enumerate(e[1]):
                if top:
                    po, prb, pnr, ptop, pbot = logs[rank - 1][i][1][j]
                    assert prb == rb and pbot == top
                if bot:
                    po, prb, pnr, ptop, pbot = logs[rank + 1][i][1][j]
                    assert prb == rb and ptop == bot
            # received blocks, in request order, against what the peer sent TO THIS RANK in the same call (same order)
            for peer in {p for p, _ in tr[2]}:
                got = [fp for p, fp in tr[2] if p == peer]
                sent = [fp for p, fp in traces[peer][i][1] if p == rank]
                bad += int(got != sent)
        else:
            want = sum(traces[r][i][1] for r in range(world))
            bad += int(not torch.equal(tr[2], want))
            assert all(lg[i][2] == e[2] for lg in logs)
    ret[rank] = (bad, nh.value, na.value, len(comm.log), comm.bytes_sent)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,variant", [(2, 64, 16, 0), (3, 64, 24, 0), (2, 64, 16, 1)])
def test_strip_schedule_dry_run_gloo(world, H, W, variant):
    """variant 1 = MSSA: its 7x7 spatial-attention gates add halo refreshes (three valid rows per up-block pair)"""
    port = 31000 + (os.getpid() % 2000) + world + 5 * variant
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dry_worker, args=(world, port, H, W, ret, variant), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        bad, nh, na, nlog, sent = ret[r]
        assert bad == 0, f"rank {r}: {bad} halo / all-reduce regions hold the wrong data"
        assert na == 6 and nh >= 6 and nlog == nh + na     # one all-reduce per LCA stage
        assert sent > 0
