"""Quick bandwidth probe of the standalone HVIT/PHVIT kernels (cfg 3 shape)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hvi_cidnet_b200.net.HVI_transform import RGB_HVI

B, H, W = 32, 1080, 1920
t = RGB_HVI().cuda()
x = torch.rand(B, 3, H, W, device="cuda")
hvi = t.HVIT(x)
res = {}
for name, fn, arg in (("hvit", t.HVIT, x), ("phvit", t.PHVIT, hvi)):
    for _ in range(3):
        fn(arg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        fn(arg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = {"ms": ms, "GBps": 24.0 * B * H * W / ms / 1e6}
print(json.dumps(res))
