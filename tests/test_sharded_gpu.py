"""Row-strip sharded forward of ONE image (BASELINE.json configs[4]; include/cidnet_b200.h
cidnet_forward_sharded) against the unsharded CUDA forward and the fp32 oracle.

The ranks are real processes.  With >= world GPUs every rank takes its own GPU and the halos /
Gram sums travel over NCCL (NVLink); on a single-GPU box the ranks share cuda:0 and the very same
callbacks are staged through gloo -- the schedule, the kernels and the row arithmetic under test
are identical.  Tolerances: owned rows equal the unsharded forward up to the summation order of the
Gram (each rank sums its own rows, then the ranks are summed), amplified by the fp16 roundings downstream -> 5e-4 (typical
1.5e-4); repeated sharded forwards are bit-equal; against the oracle the usual 2e-3 / 50 dB contract."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, parity_error, psnr_kept

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, H, W, use_nccl, ret, mssa=False, transport=None):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if use_nccl else 0)
    torch.cuda.set_device(dev)
    if use_nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import hvi_cidnet_b200  # noqa: F401
    if mssa:
        from hvi_cidnet_b200.net.CIDNet_MSSA import CIDNet
    else:
        from hvi_cidnet_b200.net.CIDNet import CIDNet
    from hvi_cidnet_b200.dist import RowShardedCIDNet
    from oracle import cidnet_oracle as O
    torch.set_grad_enabled(False)
    torch.set_num_threads(4)
    sd = O.make_state_dict(5, True, mssa=mssa)
    model = CIDNet().to(dev).eval()
    model.load_state_dict(sd, strict=True)
    x = O.make_input("uniform", 1, H, W, seed=33)
    net = RowShardedCIDNet(model, halo=16, transport=transport)
    y = net(x.pin_memory(), gather=True).cpu()           # full image on every rank
    y2 = net(x.to(dev), gather=True).cpu()               # second call: same workspace, device input
    y3 = net(x.to(dev), gather=True).cpu()               # third / fourth call: CUDA-graph capture and replay of the
    y4 = net(x.to(dev), gather=True).cpu()               # kernels + NCCL exchanges (NCCL transport only; eager otherwise)
    full = model(x.to(dev)).cpu()                        # unsharded CUDA forward on this rank's GPU
    # sharded vs unsharded: different split of the same fp32 sums -> small differences; a pixel beyond 5e-4 must be the
    # reference's black-pixel hole (conftest.parity_error).  Repeats of the SAME sharded forward are bit-equal (no atomics).
    out = {"vs_full": parity_error(y, full, 5e-4)[0], "repeat_equal": bool(torch.equal(y, y2) and torch.equal(y, y3) and torch.equal(y, y4)),
           "replays": net.replays, "graph_error": getattr(net, "graph_error", None), "transport": net.transport}
    if net.transport == "peer":
        out.update(peer_forwards=net.peer_forwards, peer_error=net.peer_error())
    else:
        out.update(halo_calls=sum(1 for e in net.comm.log if e[0] == "halo"),
                   allreduce_calls=sum(1 for e in net.comm.log if e[0] == "allreduce"), direct=bool(net.comm.direct))
    if rank == 0:
        taps = {}
        ref = O.forward(x, sd, mssa=mssa, taps=taps)
        err, _, keep = parity_error(y.clamp(0, 1), ref.clamp(0, 1), 2e-3, taps["out_hvi"], float(sd["trans.density_k"].reshape(-1)[0]))
        out["vs_oracle"] = err
        out["psnr"] = psnr_kept(y.clamp(0, 1), ref.clamp(0, 1), keep)
    ret[rank] = out
    net.close()                                          # captured graphs must go before the communicator
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["callbacks", "peer"])
@pytest.mark.parametrize("world,H,W,mssa", [(2, 128, 64, False), (3, 192, 40, False), (2, 400, 600, False),
                                            (2, 128, 64, True), (3, 240, 40, True), (2, 2160, 3840, False)])
def test_row_sharded_forward_matches_unsharded(world, H, W, mssa, transport):
    """mssa=True: the MSSA variant -- its 7x7 spatial-attention gates need three valid halo rows per up block.
    transport "peer" (CUDA-IPC mapped workspaces, halo / statistics kernels reading the neighbours' memory over NVLink,
    one CUDA graph) needs one GPU per rank: ranks that share a GPU must not spin on one another's flags.  The 4K case
    (BASELINE.json configs[4] at full size) runs where every rank has its own GPU."""
    use_nccl = torch.cuda.device_count() >= world
    if transport == "peer" and not use_nccl:
        pytest.skip("peer-memory transport needs one GPU per rank")
    if H >= 2160 and not use_nccl:
        pytest.skip("the 4K strip case needs one GPU per rank")
    port = 33000 + (os.getpid() % 2000) + world + 7 * int(mssa) + 13 * (transport == "peer") + (H // 100)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, H, W, use_nccl, ret, mssa, transport), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        o = ret[r]
        assert o["vs_full"] <= 5e-4, (r, o)
        assert o["repeat_equal"], (r, o)
        assert o["transport"] == transport
        if transport == "peer":
            assert o["peer_forwards"] == 4 and o["peer_error"] == 0, (r, o)
        else:
            assert o["allreduce_calls"] == 6 and o["halo_calls"] >= 6, (r, o)
            assert o["direct"] == use_nccl
            # graph replay of the callback transport is opt-in (CIDNET_SHARD_GRAPH=1) and needs NCCL
            assert o["replays"] == (2 if (use_nccl and os.environ.get("CIDNET_SHARD_GRAPH") == "1") else 0), (r, o)
    assert ret[0]["vs_oracle"] <= 2e-3 and ret[0]["psnr"] >= 50.0, ret[0]


def test_sharded_entry_with_one_rank_is_the_plain_forward():
    """nranks == 1: no callbacks, same kernels -> bit-identical to cidnet_forward."""
    from hvi_cidnet_b200.net.CIDNet import CIDNet
    from hvi_cidnet_b200.dist import RowShardedCIDNet
    from oracle import cidnet_oracle as O
    torch.set_grad_enabled(False)
    sd = O.make_state_dict(2, True)
    model = CIDNet().cuda().eval()
    model.load_state_dict(sd, strict=True)
    x = O.make_input("uniform", 1, 64, 96, seed=4).cuda()
    y = RowShardedCIDNet(model)(x)
    full = model(x)
    assert torch.equal(y, full)
