"""GPU PyTorch-eager bar: the oracle port of the reference forward (identical op
sequence to net/CIDNet.py, verified bit-exact on CPU) run with torch on the B200, in PyTorch's default
numeric mode (cuDNN convs in TF32) and in strict fp32.  Not part of bench.py's line; results go to
profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cidnet_oracle as O

torch.set_grad_enabled(False)
O.FAST_BILINEAR = True
dev = torch.device("cuda")
sd = {k: v.to(dev) for k, v in O.make_state_dict(0, False).items()}
out = {}
for name, (B, H, W) in {"cfg1": (1, 400, 600), "cfg2": (1, 640, 1120), "cfg4_b16": (16, 400, 600)}.items():
    x = torch.rand(B, 3, H, W, device=dev)
    for mode, tf32 in (("tf32_default", True), ("strict_fp32", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        for _ in range(3):
            O.forward(x, sd, run_dead_block=True)
        torch.cuda.synchronize()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            O.forward(x, sd, run_dead_block=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"{name}/{mode}"] = {"ms": ms, "MPps": B * H * W / ms / 1e3}
print(json.dumps(out))
