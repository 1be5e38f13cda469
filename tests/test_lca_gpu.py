"""Isolated GPU parity of ONE LCA stage pair (cidnet_test_lca_stage -> CAB kernels: q|k|v depthwise, tcgen05 Gram with
the fused sum q^2 / sum k^2, fixed-order slab reduction + softmax + fold, per-image attn.v projection; IEL kernels:
LayerNorm + project_in, gate chain, project_out) against the oracle's `lca` (net/LCA.py:19-41, 60-67, 78-93) on the
same random inputs, at C = 36 / 72 / 144, including the restricted-rows statistics the row-strip sharded forward uses.

Tolerances: the kernels store every intermediate (q|k|v, x', t, g) in 16 bits; the oracle runs fp32 on the same
fp16-rounded inputs.  The measured errors are printed; the bounds are relative to each tensor's max-abs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cidnet_oracle as O

pytestmark = pytest.mark.gpu
HEADS = {1: 2, 2: 4, 3: 8}
CH = {1: 36, 2: 72, 3: 144}


@pytest.fixture(scope="module")
def model():
    from hvi_cidnet_b200.net.CIDNet import CIDNet
    m = CIDNet().cuda().eval()
    m.load_state_dict(O.make_state_dict(11, True), strict=True)
    return m


def _act_round(t):
    from hvi_cidnet_b200 import _lib
    dt = torch.float16 if _lib.lib().cidnet_act_dtype() == 0 else torch.bfloat16
    return t.to(dt).float()


def _cab_rows(x, y, sd, pfx, heads, y0, y1):
    """oracle.cab with the normalisation sums and the Gram restricted to image rows [y0, y1) (what a rank of the
    row-strip sharded forward contributes before the all-reduce; with one rank it is the whole statistic)."""
    b, c, h, w = x.shape
    q = F.conv2d(F.conv2d(x, sd[pfx + ".q.weight"]), sd[pfx + ".q_dwconv.weight"], padding=1, groups=c)
    kv = F.conv2d(F.conv2d(y, sd[pfx + ".kv.weight"]), sd[pfx + ".kv_dwconv.weight"], padding=1, groups=2 * c)
    k, v = kv.chunk(2, dim=1)
    qs = q[:, :, y0:y1].reshape(b, heads, c // heads, -1)
    ks = k[:, :, y0:y1].reshape(b, heads, c // heads, -1)
    attn = (F.normalize(qs, dim=-1) @ F.normalize(ks, dim=-1).transpose(-2, -1)) * sd[pfx + ".temperature"]
    attn = attn.softmax(dim=-1)
    out = (attn @ v.reshape(b, heads, c // heads, h * w)).reshape(b, c, h, w)
    return F.conv2d(out, sd[pfx + ".project_out.weight"])


def _lca_ref(x, y, sd, pfx, heads, residual, rows):
    nw, nb = sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"]
    xa = x + _cab_rows(O.layer_norm_cf(x, nw, nb), O.layer_norm_cf(y, nw, nb), sd, pfx + ".ffn", heads, *rows)
    g = O.iel(O.layer_norm_cf(xa, nw, nb), sd, pfx + ".gdfn")
    return xa, (xa + g if residual else g)


@pytest.mark.parametrize("n,B,H,W,rows", [(1, 2, 40, 56, None), (2, 1, 24, 40, None), (3, 2, 16, 24, None), (4, 1, 10, 14, None),
                                          (6, 1, 33, 47, None), (1, 1, 48, 40, (8, 40)), (3, 1, 16, 24, (2, 14))])
def test_lca_stage_against_oracle(model, n, B, H, W, rows):
    lvl = n if n <= 3 else 7 - n
    C, heads = CH[lvl], HEADS[lvl]
    sd = O.make_state_dict(11, True)
    g = torch.Generator().manual_seed(100 * n + H)
    x_i = _act_round(torch.randn(B, C, H, W, generator=g) * 0.7 + 0.1)
    x_hv = _act_round(torch.randn(B, C, H, W, generator=g) * 0.5)
    got = model.run_lca_stage(n, x_i.cuda(), x_hv.cuda(), rows)
    r = rows if rows is not None else (0, H)
    ref_ai, ref_oi = _lca_ref(x_i, x_hv, sd, f"I_LCA{n}", heads, True, r)
    ref_ah, ref_oh = _lca_ref(x_hv, x_i, sd, f"HV_LCA{n}", heads, False, r)
    if rows is None:            # the helper above IS the oracle when all rows count
        assert torch.equal(ref_oi, O.lca(x_i, x_hv, sd, f"I_LCA{n}", heads, True))
    report = {}
    for name, ref, tol in (("after_cab_i", ref_ai, 4e-3), ("after_cab_hv", ref_ah, 4e-3), ("out_i", ref_oi, 8e-3), ("out_hv", ref_oh, 8e-3)):
        out = got[name].cpu()
        assert torch.isfinite(out).all(), name
        rel = float((out - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
        report[name] = rel
        assert rel <= tol, (name, rel, report)
    print(f"LCA{n} C={C} {B}x{H}x{W} rows={rows}: " + ", ".join(f"{k} {v:.2e}" for k, v in report.items()))


def test_lca_stage_is_bit_reproducible(model):
    """no atomics anywhere in the attention: the same inputs give the same bits, call after call"""
    g = torch.Generator().manual_seed(5)
    x_i = torch.randn(2, 36, 64, 80, generator=g).cuda()
    x_hv = torch.randn(2, 36, 64, 80, generator=g).cuda()
    first = model.run_lca_stage(1, x_i, x_hv)
    for _ in range(5):
        again = model.run_lca_stage(1, x_i, x_hv)
        for k in first:
            assert torch.equal(first[k], again[k]), k


@pytest.mark.parametrize("H,W", [(320, 560), (104, 64), (200, 300)])
def test_lca_stage_full_size_split_k_partitions(model, H, W):
    """L1 at BASELINE configs[1] size (320x560 = 2800 pixel chunks -> 13 chunks per Gram CTA, the case where the last
    chunk lives in pipeline stage 0) and two sizes with a ragged last chunk: the split-K partition must not matter."""
    sd = O.make_state_dict(11, True)
    g = torch.Generator().manual_seed(H)
    x_i = _act_round(torch.randn(1, 36, H, W, generator=g) * 0.7 + 0.1)
    x_hv = _act_round(torch.randn(1, 36, H, W, generator=g) * 0.5)
    got = model.run_lca_stage(1, x_i.cuda(), x_hv.cuda())
    for pfx, xa, xb, res, key in (("I_LCA1", x_i, x_hv, True, "i"), ("HV_LCA1", x_hv, x_i, False, "hv")):
        ref_a, ref_o = _lca_ref(xa, xb, sd, pfx, 2, res, (0, H))
        for name, ref, tol in ((f"after_cab_{key}", ref_a, 4e-3), (f"out_{key}", ref_o, 8e-3)):
            rel = float((got[name].cpu() - ref).abs().max() / ref.abs().max())
            assert rel <= tol, (name, rel)
