"""GPU parity of the MSSA variant (hvi-cidnet_b200/net/CIDNet_MSSA.py -> CIDNET_VARIANT_MSSA) against outputs of
the unmodified /root/reference/net/CIDNet_MSSA.py (tests/golden/forward_mssa_s*.npz) and the CPU oracle
(oracle.forward(mssa=True)).  Same contract as the base model: max-abs <= 2e-3, PSNR >= 50 dB."""
import os

import numpy as np
import pytest
import torch

from oracle import cidnet_oracle as O
from conftest import GOLDEN, parity_error, psnr_kept

pytestmark = pytest.mark.gpu
MAXABS, PSNR = 2e-3, 50.0


@pytest.fixture(scope="module")
def model():
    from hvi_cidnet_b200.net.CIDNet_MSSA import CIDNet
    return CIDNet().cuda().eval()


def _check(y, ref, x, sd):
    taps = {}
    O.forward(x, sd, taps=taps, mssa=True)
    y, ref = y.clamp(0, 1), ref.clamp(0, 1)
    err, excused, keep = parity_error(y, ref, MAXABS, taps["out_hvi"], float(sd["trans.density_k"].reshape(-1)[0]))
    ps = psnr_kept(y, ref, keep)
    assert err <= MAXABS and ps >= PSNR, f"max-abs {err:.3e}, PSNR {ps:.1f} dB, {excused} excused"


@pytest.mark.parametrize("seed", [3, 4])
def test_golden_reference_outputs(model, seed):
    g = np.load(os.path.join(GOLDEN, f"forward_mssa_s{seed}.npz"))
    sd = O.make_state_dict(int(g["seed"]), bool(g["perturb"]), mssa=True)
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        y = model(torch.from_numpy(g["x"]).cuda()).cpu()
    _check(y, torch.from_numpy(g["y"]), torch.from_numpy(g["x"]), sd)
    for key in g.files:                         # tensors right after each SpatialAttention gate, I_LCA5 (live here)
        if key.startswith("tap|"):
            ref = torch.from_numpy(g[key])
            got = model.read_tap(key[4:]).cpu()
            assert float((got - ref).abs().max()) <= 1.5e-2 * max(float(ref.abs().max()), 1.0), key


@pytest.mark.parametrize("kind,shape", [("uniform", (1, 400, 600)), ("dark", (2, 200, 304)), ("grid8", (1, 104, 72)),
                                        ("const:0.5", (1, 32, 32))])
def test_against_oracle(model, kind, shape):
    sd = O.make_state_dict(6, True, mssa=True)
    model.load_state_dict(sd, strict=True)
    x = O.make_input(kind, *shape, seed=23)
    with torch.no_grad():
        ref = O.forward(x, sd, mssa=True)
        y = model(x.cuda()).cpu()
    _check(y, ref, x, sd)


def test_differs_from_base_graph(model):
    """the gates and the live I_LCA5 must actually be on the path: same shared weights, different result"""
    from hvi_cidnet_b200.net.CIDNet import CIDNet as Base
    sd = O.make_state_dict(6, True, mssa=True)
    model.load_state_dict(sd, strict=True)
    base = Base().cuda().eval()
    base.load_state_dict({k: v for k, v in sd.items() if not k.startswith("sa_")}, strict=True)
    x = O.make_input("uniform", 1, 64, 96, seed=3).cuda()
    with torch.no_grad():
        assert float((model(x) - base(x)).abs().max()) > 1e-2
    # 3 up-block pairs x one gate launch (I and HV share it; the channel mean / max come from the up blocks' epilogues)
    # + the 3 GEMM launches a live I_LCA5 adds
    assert model.num_launches() == base.num_launches() + 3 + 3


def test_strict_load_needs_the_gate_weights(model):
    sd = O.make_state_dict(6, True, mssa=False)
    with pytest.raises(RuntimeError):
        model.load_state_dict(sd, strict=True)
