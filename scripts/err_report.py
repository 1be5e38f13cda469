"""Print the end-to-end error of the CUDA forward against the fp32 oracle for a few inputs (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cidnet_oracle as O
from hvi_cidnet_b200.net.CIDNet import CIDNet
torch.set_grad_enabled(False)
m = CIDNet().cuda().eval()
for seed, pert in ((0, False), (5, True), (3, True)):
    sd = O.make_state_dict(seed, pert)
    m.load_state_dict(sd, strict=True)
    for kind in ("uniform", "dark"):
        x = O.make_input(kind, 1, 200, 304, seed=21)
        ref = O.forward(x, sd).clamp(0, 1)
        y = m(x.cuda()).cpu().clamp(0, 1)
        print(f"seed {seed} {kind:8s} max-abs {float((y - ref).abs().max()):.3e}  PSNR {O.psnr(y, ref):.1f} dB")
