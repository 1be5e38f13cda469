"""2-rank probe of the sharded forward's CUDA-graph capture (kernels + NCCL halo / all-reduce calls).
    CIDNET_SHARD_GRAPH=1 python -X faulthandler -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/shard_graph_probe.py [H W steps]
Prints, per rank, eager vs replay step times and the max difference between the two outputs."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import cidnet_oracle as O
from hvi_cidnet_b200.net.CIDNet import CIDNet
from hvi_cidnet_b200.dist import RowShardedCIDNet, strip_plan, strip_local_range

H, W, steps = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (2160, 3840, 10)))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.set_grad_enabled(False)
def log(*a):
    print(f"[rank {rank}]", *a, flush=True)
model = CIDNet().to(dev).eval(); model.load_state_dict(O.make_state_dict(0, False))
sh = strip_plan(H, world, rank, 16); a, b = strip_local_range(sh)
x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(1))[:, :, a:b, :].contiguous().to(dev)
own = slice(sh.row_begin - a, sh.row_end - a)        # only the owned rows of the local output are meaningful
def timed(net, n):
    """device time of n steps (CUDA events on the launching stream), max over ranks"""
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y, _ = net.forward_strip(x, H)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), y[:, :, own].clone()
eager = RowShardedCIDNet(model, halo=16, graph=False)
ms_e, y_e = timed(eager, 3); ms_e, y_e = timed(eager, steps)
log(f"eager  {ms_e:.3f} ms/step")
net = RowShardedCIDNet(model, halo=16, graph=True)
for i in range(3):
    log("graph path call", i); y, _ = net.forward_strip(x, H); torch.cuda.synchronize()
    log("  done; replays", net.replays, "error", getattr(net, "graph_error", None))
ms_g, y_g = timed(net, steps)
d = (y_e - y_g).abs().amax(dim=1).flatten()
log(f"replay {ms_g:.3f} ms/step, replays {net.replays}, owned rows: max |eager - replay| = {float(d.max()):.3e}, "
    f"pixels > 5e-4: {int((d > 5e-4).sum())} of {d.numel()}")
if rank == 0:
    mp = H * W / 1e6
    print(f'{{"probe": "cfg5 row-sharded forward", "H": {H}, "W": {W}, "n_gpus": {world}, "steps": {steps}, '
          f'"eager_ms": {ms_e:.4f}, "eager_MPs": {mp / ms_e * 1e3:.1f}, "graph_ms": {ms_g:.4f}, "graph_MPs": {mp / ms_g * 1e3:.1f}}}', flush=True)
net.close(); eager.close(); del net, eager
torch.cuda.synchronize(); dist.barrier()
faulthandler.cancel_dump_traceback_later(); faulthandler.dump_traceback_later(20, exit=True)
dist.destroy_process_group()
log("clean exit")
