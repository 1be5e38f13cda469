"""GPU parity of the HVIT / PHVIT backward kernels (csrc/hvi_bwd.cu, through the C ABI and the autograd Functions of the
`RGB_HVI` mirror) against (1) gradients the UNMODIFIED reference produced under autograd (tests/golden/hvi_backward.npz,
oracle/make_golden.py), (2) the closed-form CPU oracle (oracle/hvi_backward.py) at larger and ragged sizes, and
(3) size-independent properties at the cfg-3 frame size.

Tolerance: a gradient element may differ by 2e-4 of its own magnitude + 2e-5 of the tensor's largest magnitude (fp32
evaluation order differs from autograd's op-by-op chain; the reference's own fp32 gradients differ from an fp64 evaluation
of the same formulas by up to 8e-5 relative, see the oracle pinning test); d/dk, a sum over every pixel, by 1e-5 of sum|g|."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import cidnet_oracle as O
from oracle import hvi_backward as HB
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def grad_close(a, ref, rtol=2e-4, atol_rel=2e-5, keep=None):
    a, ref = torch.as_tensor(a).cpu(), torch.as_tensor(ref).cpu()
    bound = rtol * ref.abs() + atol_rel * max(float(ref.abs().max()), 1e-6)
    ratio = (a - ref).abs() / bound
    if keep is not None:
        ratio = ratio * keep[:, None].to(ratio.dtype)
    return float(ratio.max())


def off_sextant_boundary(x, k, max_excused=6):
    """bool [B,H,W]: pixels whose 6*h the ORACLE puts at least 2e-6 away from an integer (10x the error of either atan2).  The gradient of PHVIT jumps
    across a sextant boundary (the colours do not), so a pixel ON a boundary may be routed differently by two correct
    fp32 evaluations; at most `max_excused` such pixels are taken out of a comparison, nothing else is."""
    keep = HB.phvit_sextant_margin(x, k) >= 2e-6
    assert int((~keep).sum()) <= max_excused, int((~keep).sum())
    return keep


def make_trans(k):
    from hvi_cidnet_b200.net.HVI_transform import RGB_HVI
    t = RGB_HVI().cuda()
    t.density_k.data.fill_(k)
    return t


def run_hvit(trans, x, go):
    with torch.enable_grad():
        xg = x.cuda().requires_grad_(True)
        trans.density_k.grad = None
        out = trans.HVIT(xg)
        assert out.requires_grad and out.grad_fn is not None
        out.backward(go.cuda())
    return xg.grad.cpu(), trans.density_k.grad.cpu()


def run_phvit(trans, x, go):
    with torch.enable_grad():
        xg = x.cuda().requires_grad_(True)
        trans.density_k.grad = None
        trans.PHVIT(xg).backward(go.cuda())
    assert trans.density_k.grad is None          # this_k is a python float in the reference: no gradient through PHVIT
    return xg.grad.cpu()


def test_golden_gradients_of_the_reference():
    g = np.load(os.path.join(GOLDEN, "hvi_backward.npz"))
    tags = sorted({k.rsplit("|", 1)[0] for k in g.files})
    worst_h = worst_p = worst_k = 0.0
    for tag in tags:
        parts = tag.split("|")
        k = float(parts[2].split("=")[1])
        x, go, gx = (torch.from_numpy(g[tag + "|" + s]) for s in ("x", "go", "gx"))
        trans = make_trans(k)
        if parts[0] == "hvit":
            ours, gk = run_hvit(trans, x, go)
            worst_h = max(worst_h, grad_close(ours, gx))
            assert grad_close(ours, gx) <= 1.0, tag
            ek = abs(float(gk) - float(g[tag + "|gk"][0])) / (1e-5 * float(go.abs().sum()))
            worst_k = max(worst_k, ek)
            assert ek <= 1.0, (tag, float(gk), float(g[tag + "|gk"][0]))
        else:
            gated = parts[3] == "gated=1"
            trans.this_k = k
            trans.gated, trans.gated2, trans.alpha_s, trans.alpha = gated, gated, 1.3, 0.8
            ours = run_phvit(trans, x, go)
            keep = off_sextant_boundary(x, k)
            worst_p = max(worst_p, grad_close(ours, gx, keep=keep))
            assert grad_close(ours, gx, keep=keep) <= 1.0, tag
    print(f"\n[hvi backward margin] golden: HVIT {worst_h:.3f}, PHVIT {worst_p:.3f}, d/dk {worst_k:.3f} of the tolerance")


@pytest.mark.parametrize("kind", ["uniform", "dark", "grid8", "grey8", "onehot"])
@pytest.mark.parametrize("shape", [(2, 200, 304), (1, 37, 53), (3, 1, 1)])
def test_hvit_backward_against_oracle(kind, shape):
    B, H, W = shape
    for k in (0.2, 1.0):
        kf = np.float32(k).item()
        x = O.make_input(kind, B, H, W, seed=5)
        go = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(9))
        ref, ref_k = HB.hvit_backward(x, kf, go)
        ours, gk = run_hvit(make_trans(k), x, go)
        assert grad_close(ours, ref) <= 1.0, (kind, shape, k)
        assert abs(float(gk) - float(ref_k)) <= 1e-5 * float(go.abs().sum()), (kind, shape, k, float(gk), float(ref_k))


@pytest.mark.parametrize("shape", [(2, 200, 304), (1, 37, 53), (3, 1, 1)])
@pytest.mark.parametrize("gated", [False, True])
def test_phvit_backward_against_oracle(shape, gated):
    B, H, W = shape
    rng = np.random.default_rng(17)
    for k in (0.2, 0.0, 2.5):
        for kind in ("wild", "hvit"):
            if kind == "wild":       # what the network feeds PHVIT: out of range in every channel
                x = torch.from_numpy(rng.uniform(-1.3, 1.3, (B, 3, H, W)).astype(np.float32))
            else:
                x = O.hvit(O.make_input("uniform", B, H, W, seed=3), 0.2)
            go = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(10))
            ref = HB.phvit_backward(x, k, go, gated, 1.3, gated, 0.8)
            trans = make_trans(0.2)
            trans.this_k = k
            trans.gated, trans.gated2, trans.alpha_s, trans.alpha = gated, gated, 1.3, 0.8
            ours = run_phvit(trans, x, go)
            assert grad_close(ours, ref, keep=off_sextant_boundary(x, k)) <= 1.0, (shape, gated, k, kind)


def test_phvit_uses_the_k_of_the_last_hvit():
    """PHVIT's backward reads `this_k` the way its forward does: the device snapshot taken by the last HVIT."""
    trans = make_trans(0.37)
    with torch.no_grad():
        trans.HVIT(torch.rand(1, 3, 8, 8, device="cuda"))
    trans.density_k.data.fill_(0.9)                  # must NOT be what PHVIT differentiates with
    x = O.hvit(O.make_input("uniform", 1, 40, 56, seed=4), 0.37)
    go = torch.randn(1, 3, 40, 56, generator=torch.Generator().manual_seed(2))
    ours = run_phvit(trans, x, go)
    kf = np.float32(0.37).item()
    assert grad_close(ours, HB.phvit_backward(x, kf, go), keep=off_sextant_boundary(x, kf)) <= 1.0


def test_full_hd_properties_and_determinism():
    """cfg-3 frame size (4 x 1080 x 1920): (a) two runs are bit-equal, d/dk included (fixed-order reduction, no atomics);
    (b) with only the I channel's gradient set, the result is that gradient moved to ONE channel per pixel (I = max:
    the channel sum reproduces it exactly) and d/dk is 0; (c) the backward is linear in the upstream gradient."""
    trans = make_trans(0.2)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(4, 3, 1080, 1920, device="cuda", generator=gen)
    g1 = torch.randn(4, 3, 1080, 1920, device="cuda", generator=gen)
    g2 = torch.randn(4, 3, 1080, 1920, device="cuda", generator=gen)

    def bwd(g):
        with torch.enable_grad():
            xg = x.clone().requires_grad_(True)
            trans.density_k.grad = None
            trans.HVIT(xg).backward(g)
        return xg.grad, trans.density_k.grad.clone()
    a1, k1 = bwd(g1)
    a1b, k1b = bwd(g1)
    assert torch.equal(a1, a1b) and torch.equal(k1, k1b)
    gi = torch.zeros_like(g1)
    gi[:, 2] = g1[:, 2]
    ai, ki = bwd(gi)
    assert torch.equal(ai.sum(dim=1), g1[:, 2]) and float(ki) == 0.0
    assert int((ai != 0).sum(dim=1).max()) <= 1
    a2, k2 = bwd(g2)
    a12, k12 = bwd(2.0 * g1 + g2)
    assert grad_close(a12, 2.0 * a1 + a2, rtol=1e-5, atol_rel=1e-6) <= 1.0
    assert abs(float(k12) - (2.0 * float(k1) + float(k2))) <= 1e-5 * float((2.0 * g1 + g2).abs().sum())
    # PHVIT: same two checks
    hv = trans.HVIT(x).detach()

    def pbwd(g):
        with torch.enable_grad():
            hg = hv.clone().requires_grad_(True)
            trans.PHVIT(hg).backward(g)
        return hg.grad
    p1, p1b, p2 = pbwd(g1), pbwd(g1), pbwd(g2)
    assert torch.equal(p1, p1b)
    assert grad_close(pbwd(2.0 * g1 + g2), 2.0 * p1 + p2, rtol=1e-5, atol_rel=1e-6) <= 1.0


def test_no_grad_records_nothing_and_cabi_without_dk():
    from hvi_cidnet_b200 import _lib
    trans = make_trans(0.2)
    x = torch.rand(1, 3, 16, 24, device="cuda", requires_grad=True)
    with torch.no_grad():
        assert trans.HVIT(x).grad_fn is None and trans.PHVIT(x).grad_fn is None
    # frozen density_k + image without gradient: plain launch as well
    trans.density_k.requires_grad_(False)
    with torch.enable_grad():
        assert trans.HVIT(x.detach()).grad_fn is None
        out = trans.HVIT(x)                      # image gradient only
        out.backward(torch.ones_like(out))
    assert trans.density_k.grad is None and x.grad is not None
    # C ABI directly: grad_k / scratch may be NULL; grad_k without scratch is refused
    L = _lib.lib()
    xs, go, gi = x.detach().contiguous(), torch.randn(1, 3, 16, 24, device="cuda"), torch.empty(1, 3, 16, 24, device="cuda")
    s = _lib.stream_ptr(xs.device)
    _lib.check(L.cidnet_hvit_backward(xs.data_ptr(), go.data_ptr(), gi.data_ptr(), None, None, 1, 16, 24, 0.2, None, s))
    ref, _ = HB.hvit_backward(xs.cpu(), np.float32(0.2).item(), go.cpu())
    assert grad_close(gi, ref) <= 1.0
    gk = torch.empty(1, device="cuda")
    assert L.cidnet_hvit_backward(xs.data_ptr(), go.data_ptr(), gi.data_ptr(), gk.data_ptr(), None, 1, 16, 24, 0.2, None, s) == _lib.ERR_INVALID
    assert int(L.cidnet_hvi_backward_scratch_bytes()) >= 4 * 148 * 8
    # empty batch: d/dk is defined (0) and nothing is launched
    scratch = torch.empty(int(L.cidnet_hvi_backward_scratch_bytes()) // 4, device="cuda")
    gk.fill_(7.0)
    _lib.check(L.cidnet_hvit_backward(None, None, None, gk.data_ptr(), scratch.data_ptr(), 0, 16, 24, 0.2, None, s))
    assert float(gk) == 0.0
