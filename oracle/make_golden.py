"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python oracle/make_golden.py
The reference is imported from where it lies (never copied).  Fixtures are kept
small; the weights are re-created deterministically by
`oracle.cidnet_oracle.make_state_dict(seed)` (numpy PCG64) and a checksum of
them is stored so drift is detected.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import cidnet_oracle as O  # noqa: E402
from net.CIDNet import CIDNet          # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_grad_enabled(False)


def checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def hvi_cases():
    """HVIT / PHVIT vectors, incl. ties, greys, black, white and `gated` flags."""
    out = {}
    kinds = ["uniform", "dark", "grid8", "grey8", "onehot", "const:0", "const:0.5", "const:1"]
    for k_val in (0.2, 0.37):
        model = CIDNet().eval()
        model.trans.density_k.data.fill_(k_val)
        for kind in kinds:
            x = O.make_input(kind, 1, 24, 40, seed=7)
            hvi = model.trans.HVIT(x)
            rgb = model.trans.PHVIT(hvi)
            tag = f"{kind}|k={k_val}"
            out[tag + "|x"] = x.numpy()
            out[tag + "|hvi"] = hvi.numpy()
            out[tag + "|rgb"] = rgb.numpy()
    # PHVIT on out-of-range HVI input (what the network actually feeds it) + gates
    model = CIDNet().eval()
    model.trans.density_k.data.fill_(0.2)
    rng = np.random.default_rng(11)
    hv = torch.from_numpy(rng.uniform(-1.3, 1.3, (1, 3, 24, 40)).astype(np.float32))
    model.trans.HVIT(torch.rand(1, 3, 8, 8))       # sets this_k
    out["phvit_wild|in"] = hv.numpy()
    out["phvit_wild|plain"] = model.trans.PHVIT(hv).numpy()
    model.trans.gated, model.trans.gated2 = True, True
    model.trans.alpha_s, model.trans.alpha = 1.3, 0.8
    out["phvit_wild|gated"] = model.trans.PHVIT(hv).numpy()
    fresh = CIDNet().eval()                          # this_k == 0 path (:14)
    out["phvit_wild|k0"] = fresh.trans.PHVIT(hv).numpy()
    np.savez_compressed(os.path.join(OUT, "hvi_cases.npz"), **out)
    print("hvi_cases", len(out))


def forward_cases():
    taps_wanted = ["hvi", "i_enc0", "hv_0", "i_enc1", "hv_1", "I_LCA1", "HV_LCA1", "i_enc2", "hv_2",
                   "I_LCA3", "HV_LCA4", "hvd3", "id3", "HV_LCA5", "id1", "hvd1", "out_hvi"]
    for seed, perturb, kind, (B, H, W) in [(0, True, "uniform", (2, 32, 48)),
                                           (1, False, "dark", (1, 40, 24)),
                                           (2, True, "grid8", (1, 16, 16))]:
        sd = O.make_state_dict(seed, perturb)
        model = CIDNet().eval()
        model.load_state_dict(sd, strict=True)
        x = O.make_input(kind, B, H, W, seed=100 + seed)
        y = model(x)
        # intermediates through forward hooks on the reference's own sub-modules
        caps = {}
        hooks = []
        for name in ["IE_block0", "HVE_block0", "IE_block1", "HVE_block1", "I_LCA1", "HV_LCA1", "IE_block2",
                     "HVE_block2", "I_LCA3", "HV_LCA4", "HVD_block3", "ID_block3", "HV_LCA5", "ID_block1",
                     "HVD_block1"]:
            hooks.append(getattr(model, name).register_forward_hook(
                lambda m, i, o, name=name: caps.__setitem__(name, o.detach().clone())))
        y2 = model(x)
        assert torch.equal(y, y2)
        for h in hooks:
            h.remove()
        model.trans.gated, model.trans.gated2 = True, True
        model.trans.alpha_s, model.trans.alpha = 1.3, 0.9
        yg = model(x)
        out = {"x": x.numpy(), "y": y.numpy(), "y_gated": yg.numpy(),
               "seed": np.int64(seed), "perturb": np.int64(perturb),
               "weights_abs_sum": np.float64(checksum(sd))}
        ren = {"IE_block0": "i_enc0", "HVE_block0": "hv_0", "IE_block1": "i_enc1", "HVE_block1": "hv_1",
               "IE_block2": "i_enc2", "HVE_block2": "hv_2", "HVD_block3": "hvd3", "ID_block3": "id3",
               "ID_block1": "id1", "HVD_block1": "hvd1"}
        for k_, v in caps.items():
            out["tap|" + ren.get(k_, k_)] = v.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, f"forward_s{seed}.npz"), **out)
        print("forward", seed, y.shape, float(y.mean()))
    del taps_wanted


def forward_mssa_cases():
    """The fork's MSSA variant (net/CIDNet_MSSA.py): outputs + the tensors after each SpatialAttention gate."""
    from net.CIDNet_MSSA import CIDNet as CIDNetMSSA
    for seed, perturb, kind, (B, H, W) in [(3, True, "uniform", (2, 32, 48)), (4, False, "dark", (1, 40, 24))]:
        sd = O.make_state_dict(seed, perturb, mssa=True)
        model = CIDNetMSSA().eval()
        model.load_state_dict(sd, strict=True)
        x = O.make_input(kind, B, H, W, seed=100 + seed)
        caps, hooks = {}, []
        for name, tapname in [("sa_hv3", "hvd3"), ("sa_i3", "id3"), ("I_LCA5", "I_LCA5"), ("HV_LCA5", "HV_LCA5"),
                              ("sa_hv2", "hvd2"), ("sa_i2", "id2"), ("sa_i1", "id1"), ("sa_hv1", "hvd1")]:
            hooks.append(getattr(model, name).register_forward_hook(
                lambda m, i, o, tapname=tapname: caps.__setitem__(tapname, o.detach().clone())))
        y = model(x)
        for h in hooks:
            h.remove()
        out = {"x": x.numpy(), "y": y.numpy(), "seed": np.int64(seed), "perturb": np.int64(perturb),
               "weights_abs_sum": np.float64(checksum(sd))}
        for k_, v in caps.items():
            out["tap|" + k_] = v.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, f"forward_mssa_s{seed}.npz"), **out)
        print("forward_mssa", seed, y.shape, float(y.mean()))
    with open(os.path.join(OUT, "state_dict_keys_mssa.txt"), "w") as f:
        for k_, v in CIDNetMSSA().state_dict().items():
            f.write(f"{k_} {list(v.shape)}\n")


def hvi_backward_cases():
    """Gradients of HVIT / PHVIT as the reference's training loop sees them (train.py:61-62, CIDNet.py:121): the
    UNMODIFIED reference under autograd.  Inputs, upstream gradients, input gradients and d/d density_k."""
    out = {}
    with torch.enable_grad():
        for k_val in (0.2, 0.37, 1.0):
            for kind in ["uniform", "dark", "grid8", "grey8", "onehot", "const:0", "const:0.5", "const:1"]:
                model = CIDNet().eval()
                model.trans.density_k.data.fill_(k_val)
                x = O.make_input(kind, 1, 24, 40, seed=7).requires_grad_(True)
                go = torch.randn(1, 3, 24, 40, generator=torch.Generator().manual_seed(21))
                model.trans.HVIT(x).backward(go)
                tag = f"hvit|{kind}|k={k_val}"
                out[tag + "|x"] = x.detach().numpy()
                out[tag + "|go"] = go.numpy()
                out[tag + "|gx"] = x.grad.numpy()
                out[tag + "|gk"] = model.trans.density_k.grad.numpy()
        rng = np.random.default_rng(11)
        for k_val in (0.2, 0.0, 1.0):
            for gated in (False, True):
                for kind in ("wild", "hvit"):
                    model = CIDNet().eval()
                    model.trans.this_k = k_val
                    model.trans.gated, model.trans.gated2 = gated, gated
                    model.trans.alpha_s, model.trans.alpha = 1.3, 0.8
                    if kind == "wild":
                        hv = torch.from_numpy(rng.uniform(-1.3, 1.3, (1, 3, 24, 40)).astype(np.float32))
                    else:
                        hv = O.hvit(O.make_input("uniform", 1, 24, 40, seed=3), 0.2)
                    hv = hv.clone().requires_grad_(True)
                    go = torch.randn(1, 3, 24, 40, generator=torch.Generator().manual_seed(22))
                    model.trans.PHVIT(hv).backward(go)
                    tag = f"phvit|{kind}|k={k_val}|gated={int(gated)}"
                    out[tag + "|x"] = hv.detach().numpy()
                    out[tag + "|go"] = go.numpy()
                    out[tag + "|gx"] = hv.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "hvi_backward.npz"), **out)
    print("hvi_backward", len(out))


def state_dict_keys():
    model = CIDNet()
    with open(os.path.join(OUT, "state_dict_keys.txt"), "w") as f:
        for k_, v in model.state_dict().items():
            f.write(f"{k_} {list(v.shape)}\n")


if __name__ == "__main__":
    torch.manual_seed(0)
    only = sys.argv[1:]          # e.g. `python oracle/make_golden.py hvi_backward_cases` regenerates one fixture
    for fn in (hvi_cases, forward_cases, forward_mssa_cases, hvi_backward_cases, state_dict_keys):
        if not only or fn.__name__ in only:
            fn()
