"""Run-to-run reproducibility of the forward: N replays on the same input, max |y_i - y_0| per run.
The only run-dependent arithmetic is the order of the Gram's fp32 atomics; anything beyond ~1e-4 is a race."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cidnet_oracle as O
from hvi_cidnet_b200.net.CIDNet import CIDNet
torch.set_grad_enabled(False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
shapes = (((1, 48, 64), 7), ((1, 400, 600), 5), ((3, 48, 64), 7), ((2, 200, 304), 3))
if len(sys.argv) > 2:
    shapes = (((1, 400, 600), 5),)
for (B, H, W), seed in shapes:
    m = CIDNet().cuda().eval()
    sd = O.make_state_dict(seed, True)
    m.load_state_dict(sd, strict=True)
    x = O.make_input("uniform", B, H, W, seed=4).cuda()
    y0 = m(x).clone()
    worst, big, first_bad = 0.0, 0, None
    for i in range(n):
        y = m(x)
        d = float((y - y0).abs().max())
        if d > 1e-3 and first_bad is None:
            bad = ((y - y0).abs() > 1e-3)
            idx = bad.nonzero()
            first_bad = (i, int(bad.sum()), idx.min(0).values.tolist(), idx.max(0).values.tolist())
        worst = max(worst, d); big += d > 1e-3
    ref = O.forward(x.cpu(), sd).clamp(0, 1)
    print(f"{B}x{H}x{W}: worst run-to-run diff {worst:.3e}, runs above 1e-3: {big}/{n}, first bad {first_bad}, vs oracle {float((y0.cpu().clamp(0,1) - ref).abs().max()):.3e}")
