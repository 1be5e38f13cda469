"""Per-stage parity report: run the CUDA forward and the CPU oracle on the same input and
weights and print max-abs / relative error of every tap in graph order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cidnet_oracle as O
from hvi_cidnet_b200.net.CIDNet import CIDNet

torch.set_grad_enabled(False)
B, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (2, 64, 96)))
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sd = O.make_state_dict(seed, True)
x = O.make_input("uniform", B, H, W, seed=3)
taps = {}
ref = O.forward(x, sd, taps=taps)
m = CIDNet().cuda().eval()
m.load_state_dict(sd)
xc = x.cuda()
y = m(xc).cpu()            # eager (first call with this shape)
y2 = m(xc).cpu()           # captured into a CUDA graph
y3 = m(xc.clone()).cpu()   # graph replay with a patched input pointer
torch.cuda.synchronize()
print('graph replay vs eager max diff:', float((y2 - y).abs().max()), float((y3 - y).abs().max()))
order = ["hvi", "i_enc0", "hv_0", "i_enc1", "hv_1", "I_LCA1", "HV_LCA1",
         "i_enc2", "hv_2", "I_LCA2", "HV_LCA2", "i_enc3", "hv_3", "I_LCA3", "HV_LCA3", "I_LCA4", "HV_LCA4",
         "hvd3", "id3", "HV_LCA5", "hvd2", "id2", "I_LCA6", "HV_LCA6", "id1", "hvd1", "out_hvi"]
for name in order:
    a = m.read_tap(name).cpu()
    r = taps[name]
    d = (a - r).abs()
    print(f"{name:22s} shape={tuple(r.shape)} max|ref|={float(r.abs().max()):9.4f} maxabs={float(d.max()):.3e} "
          f"rel={float(d.max() / r.abs().max().clamp_min(1e-9)):.3e} nan={int(torch.isnan(a).sum())}")
d = (y.clamp(0, 1) - ref.clamp(0, 1)).abs()
print(f"OUTPUT maxabs={float(d.max()):.3e} psnr={O.psnr(y.clamp(0,1), ref.clamp(0,1)):.2f} dB launches={m.num_launches()}")
