"""Per-kernel CUDA-event times of a few forwards (library profiling marks), for A/B runs of env-selected variants.
    CIDNET_DW_VARIANT=7 python scripts/time_kernels.py 1 640 1120 [filter]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.cidnet_oracle import make_state_dict
from hvi_cidnet_b200.net.CIDNet import CIDNet

B, H, W = (int(v) for v in sys.argv[1:4])
flt = sys.argv[4] if len(sys.argv) > 4 else ""
torch.set_grad_enabled(False)
m = CIDNet().cuda().eval(); m.load_state_dict(make_state_dict(0, False))
xs = [torch.rand(B, 3, H, W, device="cuda") for _ in range(4)]
for i in range(6):
    y = m(xs[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    y = m(xs[i % 4])
e1.record(); torch.cuda.synchronize()
m.set_profiling(True)
agg = {}
for i in range(5):
    m(xs[i % 4])
    for name, ms, by, fl in m.read_profile():
        a = agg.setdefault(name, [0.0, 0.0]); a[0] += ms / 5; a[1] += by / 5
tot = sum(a[0] for a in agg.values())
print(f"step {e0.elapsed_time(e1) / 10:.4f} ms (graph), kernel sum {tot:.4f} ms, checksum {float(y.double().mean()):.6f}")
for k, (ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if flt in k:
        print(f"  {k:36s} {ms * 1e3:8.1f} us {by / ms / 1e6:7.0f} GB/s")
