"""Per-tap error report of one forward against the oracle: python scripts/diag_taps.py B H W [seed perturb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cidnet_oracle as O
from hvi_cidnet_b200.net.CIDNet import CIDNet
B, H, W = (int(v) for v in sys.argv[1:4])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
perturb = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
torch.set_grad_enabled(False)
sd = O.make_state_dict(seed, perturb)
m = CIDNet().cuda().eval(); m.load_state_dict(sd, strict=True)
x = O.make_input("uniform", B, H, W, seed=1234)
taps = {}
ref = O.forward(x, sd, taps=taps)
for run in range(3):
    y = m(x.cuda()).cpu()
    print(f"run {run} ({'eager' if run == 0 else 'graph'}): out max-abs {float((y.clamp(0,1)-ref.clamp(0,1)).abs().max()):.3e}")
    if run in (0, 2):
        for name in ("hvi", "i_enc0", "hv_0", "i_enc1", "hv_1", "I_LCA1.after_cab", "HV_LCA1.after_cab", "I_LCA1", "HV_LCA1", "i_enc2", "hv_2",
                     "I_LCA2.after_cab", "I_LCA2", "HV_LCA2", "i_enc3", "hv_3", "I_LCA3", "HV_LCA3", "I_LCA4", "HV_LCA4", "hvd3", "id3", "HV_LCA5", "hvd2", "id2",
                     "I_LCA6", "HV_LCA6", "id1", "hvd1", "out_hvi"):
            try:
                t = m.read_tap(name).cpu()
            except Exception as e:
                print("  ", name, "unavailable"); continue
            r = taps[name]
            print(f"   {name:20s} rel {float((t-r).abs().max()/r.abs().max().clamp_min(1e-6)):.3e}  scale {float(r.abs().max()):.3f}")
