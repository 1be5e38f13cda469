"""End-to-end GPU parity of CIDNet.forward (C ABI -> sm_100a kernels) against the CPU oracle and
the reference-generated golden outputs.  Contract (BASELINE.json north_star): max-abs <= 2e-3 and
PSNR >= 50 dB on the clamped RGB output; operands/activations are 16-bit (fp16 by default, fp32
accumulate), the oracle is strict fp32."""
import os

import numpy as np
import pytest
import torch

from oracle import cidnet_oracle as O
from conftest import GOLDEN, max_err_robust

pytestmark = pytest.mark.gpu
MAXABS, PSNR = 2e-3, 50.0


@pytest.fixture(scope="module")
def model():
    from hvi_cidnet_b200.net.CIDNet import CIDNet
    return CIDNet().cuda().eval()


def _check(y, ref):
    y, ref = y.clamp(0, 1), ref.clamp(0, 1)
    err = max_err_robust(y, ref)             # all but <= 3 pixels (PHVIT's black-pixel discontinuity, see conftest)
    ps = O.psnr(y, ref)
    assert err <= MAXABS and ps >= PSNR, f"max-abs {err:.3e}, PSNR {ps:.1f} dB"
    return err, ps


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_golden_reference_outputs(model, seed):
    g = np.load(os.path.join(GOLDEN, f"forward_s{seed}.npz"))
    sd = O.make_state_dict(int(g["seed"]), bool(g["perturb"]))
    model.load_state_dict(sd, strict=True)
    model.trans.gated = model.trans.gated2 = False
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        y = model(x).cpu()
    _check(y, torch.from_numpy(g["y"]))
    # per-stage taps against the reference's own sub-module outputs
    for key in g.files:
        if key.startswith("tap|"):
            ref = torch.from_numpy(g[key])
            got = model.read_tap(key[4:]).cpu()
            scale = float(ref.abs().max())
            assert float((got - ref).abs().max()) <= 1.5e-2 * max(scale, 1.0), key
    model.trans.gated, model.trans.gated2, model.trans.alpha_s, model.trans.alpha = True, True, 1.3, 0.9
    with torch.no_grad():
        yg = model(x).cpu()
    _check(yg, torch.from_numpy(g["y_gated"]))
    model.trans.gated = model.trans.gated2 = False
    model.trans.alpha = 1.0


@pytest.mark.parametrize("kind,shape", [("uniform", (1, 400, 600)), ("dark", (1, 400, 600)), ("grid8", (2, 200, 304)),
                                        ("uniform", (3, 64, 72)), ("const:0.5", (1, 32, 32)), ("const:0", (1, 16, 24))])
def test_against_oracle(model, kind, shape):
    sd = O.make_state_dict(5, True)
    model.load_state_dict(sd, strict=True)
    x = O.make_input(kind, *shape, seed=21)
    with torch.no_grad():
        ref = O.forward(x, sd)
        y = model(x.cuda()).cpu()
    _check(y, ref)


def test_cfg2_full_size_against_oracle_and_batch_property(model):
    """BASELINE.json configs[1] at its full size (1x3x640x1120): direct parity against the fp32 oracle, plus the
    size-independent property that a batch of two copies gives two equal results equal to the single-image run."""
    sd = O.make_state_dict(0, False)                 # the bench's weights (default init values)
    model.load_state_dict(sd, strict=True)
    x = O.make_input("uniform", 1, 640, 1120, seed=1234)
    with torch.no_grad():
        ref = O.forward(x, sd)
        xd = x.cuda()
        y = model(xd)
        y2 = model(torch.cat([xd, xd]))
    _check(y.cpu(), ref)
    assert max_err_robust(y2[0:1], y2[1:2]) <= 5e-4 and max_err_robust(y2[0:1], y) <= 5e-4


def test_default_init_state_dict_round_trip(model, tmp_path):
    """torch.save(state_dict) -> torch.load(map_location=cpu) -> strict load (eval_SID_blur.py:22)."""
    from hvi_cidnet_b200.net.CIDNet import CIDNet
    torch.manual_seed(0)
    fresh = CIDNet()
    path = os.path.join(tmp_path, "w.pth")
    torch.save(fresh.state_dict(), path)
    sd = torch.load(path, map_location=lambda storage, loc: storage)
    model.load_state_dict(sd, strict=True)
    x = O.make_input("uniform", 1, 64, 96, seed=2)
    with torch.no_grad():
        y = model(x.cuda()).cpu()
        ref = O.forward(x, {k: v.float() for k, v in sd.items()})
    _check(y, ref)
    assert abs(model.trans.this_k - 0.2) < 1e-6        # set by forward's HVIT (HVI_transform.py:38)


def test_shape_errors(model):
    with pytest.raises(RuntimeError):
        model(torch.rand(1, 3, 20, 24, device="cuda"))      # H % 8 != 0: reference dies in torch.cat
    with pytest.raises(RuntimeError):
        model(torch.rand(1, 3, 16, 16))                      # CPU tensor: no fallback
    assert model(torch.empty(0, 3, 16, 16, device="cuda")).shape == (0, 3, 16, 16)


def test_batch_independence(model):
    """images are independent units (attention is per image): a batch equals its images run alone."""
    sd = O.make_state_dict(7, True)
    model.load_state_dict(sd, strict=True)
    x = O.make_input("uniform", 3, 48, 64, seed=4).cuda()
    with torch.no_grad():
        yb = model(x)
        ys = torch.cat([model(x[i:i + 1]) for i in range(3)])
    # fp32 atomics accumulate the Gram partial sums in a run-dependent order; downstream fp16 roundings amplify
    # the last-bit differences to ~1.5e-4 on the output (measured over 300 replays: scripts/stress_repeat.py)
    assert max_err_robust(yb, ys) <= 5e-4


def test_streamed_driver_matches_plain_forward(model):
    """hvi-cidnet_b200/stream.py: 3-stream pipelined host->device->host loop == model(x) per batch, in order."""
    from hvi_cidnet_b200.stream import StreamedCIDNet
    sd = O.make_state_dict(5, True)
    model.load_state_dict(sd, strict=True)
    xs = [O.make_input("uniform", 2, 64, 96, seed=100 + i).pin_memory() for i in range(7)]
    with torch.no_grad():
        want = [model(x.cuda()).cpu() for x in xs]
        got = [y.clone() for y in StreamedCIDNet(model, depth=3).run(iter(xs))]
    assert len(got) == len(want)
    for g, w in zip(got, want):        # not bit-equal: the Gram's fp32 atomics are order dependent run to run
        assert max_err_robust(g, w) <= 5e-4
    with pytest.raises(RuntimeError):
        model(xs[0].cuda(), out=torch.empty(1, 3, 8, 8, device="cuda"))


@pytest.mark.parametrize("h,w,gamma", [(61, 93, 1.0), (64, 96, 1.0), (50, 77, 0.8)])
def test_u8_pre_post_and_enhance(model, h, w, gamma):
    """8-bit I/O path (cidnet_pre_u8 / cidnet_post_u8 / CIDNet.enhance_u8) against the oracle's restatement
    of the reference's caller code (ToTensor, reflect pad to x8, **gamma; clamp, crop, ToPILImage)."""
    import ctypes as C
    from hvi_cidnet_b200 import _lib
    sd = O.make_state_dict(5, True)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(7)
    img = torch.randint(0, 256, (2, h, w, 3), generator=g, dtype=torch.uint8)
    ref_x = O.pre_u8(img, gamma)
    H, W = ref_x.shape[2], ref_x.shape[3]
    lib = _lib.lib()
    d_img = img.cuda()
    x = torch.empty(2, 3, H, W, device="cuda")
    _lib.check(lib.cidnet_pre_u8(d_img.data_ptr(), x.data_ptr(), 2, h, w, H, W, float(gamma), _lib.stream_ptr(x.device)))
    if gamma == 1.0:
        assert torch.equal(x.cpu(), ref_x)                        # bit exact: same division, pure data movement
    else:
        assert float((x.cpu() - ref_x).abs().max()) <= 2e-6       # powf vs torch.pow
    # post on identical fp32 input: exact (integer result of the same fp32 multiply + truncation)
    yy = (torch.rand(2, 3, H, W, generator=g) * 1.4 - 0.2)
    ref_o = O.post_u8(yy, h, w)
    o = torch.empty(2, h, w, 3, dtype=torch.uint8, device="cuda")
    _lib.check(lib.cidnet_post_u8(yy.cuda().data_ptr(), o.data_ptr(), 2, h, w, H, W, _lib.stream_ptr(o.device)))
    assert torch.equal(o.cpu(), ref_o)
    # end to end: 8-bit in, 8-bit out; the forward's 2e-3 tolerance is < 1 LSB (1/255), truncation can flip one level
    with torch.no_grad():
        got = model.enhance_u8(d_img, gamma).cpu()
        want = O.post_u8(O.forward(ref_x, sd), h, w)
    diff = (got.int() - want.int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 0.1
    with pytest.raises(RuntimeError):
        model.enhance_u8(img)                                     # CPU tensor: no fallback
