"""Run a few forwards of one workload (target of the ncu captures; no timing here).
    python scripts/prof_forward.py B H W n [marks.txt]
With a 5th argument the names of the forward's launches (in launch order, from the library's own
profiling marks) are written to that file, so an ncu launch list can be labelled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.cidnet_oracle import make_state_dict
from hvi_cidnet_b200.net.CIDNet import CIDNet

B, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (1, 640, 1120)))
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.set_grad_enabled(False)
m = CIDNet().cuda().eval()
m.load_state_dict(make_state_dict(0, False))
x = torch.rand(B, 3, H, W, device="cuda")
for _ in range(n):
    y = m(x)
torch.cuda.synchronize()
print("ok", float(y.mean()), m.num_launches())
if len(sys.argv) > 5:
    m.set_profiling(True)
    m(x)
    names = [r[0] + "\t%.1f\t%.0f" % (r[1] * 1e3, r[2]) for r in m.read_profile()]
    m.set_profiling(False)
    open(sys.argv[5], "w").write("\n".join(names) + "\n")
