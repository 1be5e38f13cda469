"""CPU oracle for the CIDNet inference forward path -- TEST INFRASTRUCTURE ONLY.

This file is a from-scratch fp32 restatement (functional torch + explicit index
math) of the reference algorithm.  It is NOT the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / CPU baseline.  The product path
(`hvi-cidnet_b200/`) never imports anything from `oracle/`.

Parity pinning: the reference ships no golden vectors or tests (SURVEY.md §4,
§8c).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF,
generated in the build container by importing `/root/reference/net` unmodified
(`oracle/make_golden.py`) and committed under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against them on CPU.

Reference lines followed (relative to /root/reference):
  hvit            net/HVI_transform.py:16-47
  phvit           net/HVI_transform.py:49-122
  layer_norm_cf   net/transformer_utils.py:21-29
  bilinear_ac     torch UpsamplingBilinear2d (align_corners=True), used at
                  net/transformer_utils.py:40,59
  norm_downsample net/transformer_utils.py:31-48
  norm_upsample   net/transformer_utils.py:50-70
  cab             net/LCA.py:19-41
  iel             net/LCA.py:60-67
  hv_lca / i_lca  net/LCA.py:78-81 / :90-93
  forward         net/CIDNet.py:71-122
  spatial_attention / forward(mssa=True)
                  net/CIDNet_MSSA.py:10-25 / :100-161 (the fork's MSSA variant)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

PI = 3.141592653589793  # net/HVI_transform.py:4
EPS = 1e-8

FAST_BILINEAR = False   # bench.py's CPU baseline sets this; tests use the explicit restatement

CHANNELS = (36, 36, 72, 144)
HEADS = (1, 2, 4, 8)


# --------------------------------------------------------------------------- #
# HVI transform
# --------------------------------------------------------------------------- #
def hvit(img: torch.Tensor, k: float) -> torch.Tensor:
    """RGB -> HVI.  net/HVI_transform.py:16-47.

    Mask priority follows the assignment order of the reference (:23-27):
    (min==max) > (r==max) > (g==max) > (b==max).  `k` is density_k as a python
    float (the reference uses the 1-element parameter tensor; identical in fp32).
    """
    r, g, b = img[:, 0], img[:, 1], img[:, 2]
    value = torch.maximum(torch.maximum(r, g), b)
    vmin = torch.minimum(torch.minimum(r, g), b)
    d = value - vmin + EPS
    hue_b = 4.0 + (r - g) / d
    hue_g = 2.0 + (b - r) / d
    hue_r = torch.remainder((g - b) / d, 6)          # python-style %, :25
    hue = torch.where(b == value, hue_b, torch.zeros_like(value))
    hue = torch.where(g == value, hue_g, hue)
    hue = torch.where(r == value, hue_r, hue)
    hue = torch.where(vmin == value, torch.zeros_like(hue), hue)
    hue = hue / 6.0
    sat = (value - vmin) / (value + EPS)
    sat = torch.where(value == 0, torch.zeros_like(sat), sat)
    kt = torch.full([1], float(k), dtype=img.dtype, device=img.device)
    cs = ((value * 0.5 * PI).sin() + EPS).pow(kt)
    ch = (2.0 * PI * hue).cos()
    cv = (2.0 * PI * hue).sin()
    return torch.stack([cs * sat * ch, cs * sat * cv, value], dim=1)


def phvit(img: torch.Tensor, k: float, gated: bool = False, alpha_s: float = 1.3,
          gated2: bool = False, alpha: float = 1.0) -> torch.Tensor:
    """HVI -> RGB.  net/HVI_transform.py:49-122.  `k` is the python float
    `this_k` (0 when HVIT was never called, :14)."""
    H = torch.clamp(img[:, 0], -1, 1)
    V = torch.clamp(img[:, 1], -1, 1)
    I = torch.clamp(img[:, 2], 0, 1)
    v = I
    cs = ((v * 0.5 * PI).sin() + EPS).pow(k)
    H = torch.clamp(H / (cs + EPS), -1, 1)
    V = torch.clamp(V / (cs + EPS), -1, 1)
    h = torch.atan2(V + EPS, H + EPS) / (2 * PI)
    h = h % 1
    s = torch.sqrt(H ** 2 + V ** 2 + EPS)
    if gated:
        s = s * alpha_s
    s = torch.clamp(s, 0, 1)
    v = torch.clamp(v, 0, 1)
    hi = torch.floor(h * 6.0)
    f = h * 6.0 - hi
    p = v * (1.0 - s)
    q = v * (1.0 - (f * s))
    t = v * (1.0 - ((1.0 - f) * s))
    zero = torch.zeros_like(h)
    # sextant table :92-114; a pixel whose hi is outside 0..5 stays black.
    table_r = (v, q, p, p, t, v)
    table_g = (t, v, v, q, p, p)
    table_b = (p, p, t, v, v, q)
    r, g, b = zero, zero, zero
    for n in range(6):
        m = hi == n
        r = torch.where(m, table_r[n], r)
        g = torch.where(m, table_g[n], g)
        b = torch.where(m, table_b[n], b)
    rgb = torch.stack([r, g, b], dim=1)
    if gated2:
        rgb = rgb * alpha
    return rgb


# --------------------------------------------------------------------------- #
# blocks
# --------------------------------------------------------------------------- #
def layer_norm_cf(x, w, b, eps=1e-6):
    """channels_first LayerNorm, net/transformer_utils.py:25-28 (biased var)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[None, :, None, None] * x + b[None, :, None, None]


def bilinear_ac(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """Bilinear resampling with align_corners=True, written out explicitly
    (ratio (in-1)/(out-1) in fp32, i0=(int)src, i1=i0+(i0<in-1), lam=src-i0),
    the semantics of nn.UpsamplingBilinear2d at transformer_utils.py:40,59."""
    if FAST_BILINEAR:   # timing runs only: the ATen kernel the reference itself calls (equal to 5e-7)
        return F.interpolate(x, size=(out_h, out_w), mode="bilinear", align_corners=True)
    B, C, H, W = x.shape

    def axis(n_in, n_out):
        r = torch.tensor((n_in - 1) / (n_out - 1) if n_out > 1 else 0.0, dtype=torch.float32)
        dst = torch.arange(n_out, dtype=torch.float32)
        src = r * dst
        i0 = src.to(torch.int64)
        i1 = i0 + (i0 < n_in - 1).to(torch.int64)
        lam = src - i0.to(torch.float32)
        return i0, i1, lam

    y0, y1, ly = axis(H, out_h)
    x0, x1, lx = axis(W, out_w)
    ly = ly.to(x.dtype)[None, None, :, None]
    lx = lx.to(x.dtype)[None, None, None, :]
    top = x[:, :, y0][:, :, :, x0] * (1 - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (1 - lx) + x[:, :, y1][:, :, :, x1] * lx
    return top * (1 - ly) + bot * ly


def prelu(x, w):
    return torch.where(x >= 0, x, w.reshape(1, 1, 1, 1) * x)


def norm_downsample(x, sd, pfx):
    """conv3x3 (zero pad) at input resolution -> bilinear x0.5 -> PReLU.
    net/transformer_utils.py:38-43 (use_norm=False)."""
    x = F.conv2d(x, sd[pfx + ".down.0.weight"], padding=1)
    x = bilinear_ac(x, int(math.floor(x.shape[2] * 0.5)), int(math.floor(x.shape[3] * 0.5)))
    return prelu(x, sd[pfx + ".prelu.weight"])


def norm_upsample(x, y, sd, pfx):
    """conv3x3 -> bilinear x2 -> cat skip -> 1x1 -> PReLU.
    net/transformer_utils.py:57-66 (use_norm=False)."""
    x = F.conv2d(x, sd[pfx + ".up_scale.0.weight"], padding=1)
    x = bilinear_ac(x, x.shape[2] * 2, x.shape[3] * 2)
    x = torch.cat([x, y], dim=1)
    x = F.conv2d(x, sd[pfx + ".up.weight"])
    return prelu(x, sd[pfx + ".prelu.weight"])


def block0(x, w):
    """ReplicationPad2d(1) + conv3x3 no bias.  net/CIDNet.py:21-24 etc."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w)


def cab(x, y, sd, pfx, heads):
    """Channel cross-attention.  net/LCA.py:19-41."""
    b, c, h, w = x.shape
    q = F.conv2d(F.conv2d(x, sd[pfx + ".q.weight"]), sd[pfx + ".q_dwconv.weight"], padding=1, groups=c)
    kv = F.conv2d(F.conv2d(y, sd[pfx + ".kv.weight"]), sd[pfx + ".kv_dwconv.weight"], padding=1, groups=2 * c)
    k, v = kv.chunk(2, dim=1)
    q = q.reshape(b, heads, c // heads, h * w)
    k = k.reshape(b, heads, c // heads, h * w)
    v = v.reshape(b, heads, c // heads, h * w)
    q = F.normalize(q, dim=-1)
    k = F.normalize(k, dim=-1)
    attn = (q @ k.transpose(-2, -1)) * sd[pfx + ".temperature"]
    attn = attn.softmax(dim=-1)
    out = (attn @ v).reshape(b, c, h, w)
    return F.conv2d(out, sd[pfx + ".project_out.weight"])


def iel(x, sd, pfx):
    """Gated depthwise FFN.  net/LCA.py:60-67."""
    x = F.conv2d(x, sd[pfx + ".project_in.weight"])
    c2 = x.shape[1]
    x1, x2 = F.conv2d(x, sd[pfx + ".dwconv.weight"], padding=1, groups=c2).chunk(2, dim=1)
    hdim = c2 // 2
    x1 = torch.tanh(F.conv2d(x1, sd[pfx + ".dwconv1.weight"], padding=1, groups=hdim)) + x1
    x2 = torch.tanh(F.conv2d(x2, sd[pfx + ".dwconv2.weight"], padding=1, groups=hdim)) + x2
    return F.conv2d(x1 * x2, sd[pfx + ".project_out.weight"])


def lca(x, y, sd, pfx, heads, residual_ffn: bool, taps: Optional[dict] = None):
    """HV_LCA (residual_ffn=False, net/LCA.py:78-81) / I_LCA (True, :90-93)."""
    nw, nb = sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"]
    x = x + cab(layer_norm_cf(x, nw, nb), layer_norm_cf(y, nw, nb), sd, pfx + ".ffn", heads)
    if taps is not None:
        taps[pfx + ".after_cab"] = x
    g = iel(layer_norm_cf(x, nw, nb), sd, pfx + ".gdfn")
    return x + g if residual_ffn else g


def spatial_attention(x, w):
    """SpatialAttention (net/CIDNet_MSSA.py:10-25): channel mean and max -> 7x7 conv (2 -> 1, zero
    padding 3, no bias) -> sigmoid -> gate on every channel."""
    avg = x.mean(dim=1, keepdim=True)
    mx = x.max(dim=1, keepdim=True)[0]
    y = F.conv2d(torch.cat([avg, mx], dim=1), w, padding=w.shape[-1] // 2)
    return x * torch.sigmoid(y)


# --------------------------------------------------------------------------- #
# whole network
# --------------------------------------------------------------------------- #
def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], gated=False, alpha_s=1.3,
            gated2=False, alpha=1.0, taps: Optional[dict] = None,
            run_dead_block: bool = False, mssa: bool = False) -> torch.Tensor:
    """net/CIDNet.py:71-122.  `taps`, when given, receives named intermediates
    (NCHW fp32) for per-kernel parity tests.  I_LCA5 (:105) is dead in the
    reference (its result is overwritten at :109) and is skipped unless
    `run_dead_block` is set (used only for timing the reference's full work).
    `mssa=True` follows net/CIDNet_MSSA.py:100-161 instead: a SpatialAttention gate after each of
    the six up blocks, I_LCA5 live, and ID_block2 fed by I_LCA5's output (:143-144)."""
    h2, h3, h4 = HEADS[1], HEADS[2], HEADS[3]
    k = float(sd["trans.density_k"].reshape(-1)[0])
    hvi = hvit(x, k)
    i = hvi[:, 2:3]

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    tap("hvi", hvi)
    i_enc0 = tap("i_enc0", block0(i, sd["IE_block0.1.weight"]))
    i_enc1 = tap("i_enc1", norm_downsample(i_enc0, sd, "IE_block1"))
    hv_0 = tap("hv_0", block0(hvi, sd["HVE_block0.1.weight"]))
    hv_1 = tap("hv_1", norm_downsample(hv_0, sd, "HVE_block1"))
    i_jump0, hv_jump0 = i_enc0, hv_0

    i_enc2 = tap("I_LCA1", lca(i_enc1, hv_1, sd, "I_LCA1", h2, True, taps))
    hv_2 = tap("HV_LCA1", lca(hv_1, i_enc1, sd, "HV_LCA1", h2, False, taps))
    v_jump1, hv_jump1 = i_enc2, hv_2
    i_enc2 = tap("i_enc2", norm_downsample(i_enc2, sd, "IE_block2"))
    hv_2 = tap("hv_2", norm_downsample(hv_2, sd, "HVE_block2"))

    i_enc3 = tap("I_LCA2", lca(i_enc2, hv_2, sd, "I_LCA2", h3, True, taps))
    hv_3 = tap("HV_LCA2", lca(hv_2, i_enc2, sd, "HV_LCA2", h3, False, taps))
    v_jump2, hv_jump2 = i_enc3, hv_3
    i_enc3 = tap("i_enc3", norm_downsample(i_enc2, sd, "IE_block3"))   # pre-LCA2 tensors, :94-95
    hv_3 = tap("hv_3", norm_downsample(hv_2, sd, "HVE_block3"))

    i_enc4 = tap("I_LCA3", lca(i_enc3, hv_3, sd, "I_LCA3", h4, True, taps))
    hv_4 = tap("HV_LCA3", lca(hv_3, i_enc3, sd, "HV_LCA3", h4, False, taps))

    i_dec4 = tap("I_LCA4", lca(i_enc4, hv_4, sd, "I_LCA4", h4, True, taps))
    hv_4 = tap("HV_LCA4", lca(hv_4, i_enc4, sd, "HV_LCA4", h4, False, taps))

    def sa(name, t):
        return spatial_attention(t, sd[name + ".conv1.weight"]) if mssa else t

    hv_3 = tap("hvd3", sa("sa_hv3", norm_upsample(hv_4, hv_jump2, sd, "HVD_block3")))
    i_dec3 = tap("id3", sa("sa_i3", norm_upsample(i_dec4, v_jump2, sd, "ID_block3")))
    i_dec2 = i_dec3
    if mssa:
        i_dec2 = tap("I_LCA5", lca(i_dec3, hv_3, sd, "I_LCA5", h3, True, taps))   # live, CIDNet_MSSA.py:139
    elif run_dead_block:
        lca(i_dec3, hv_3, sd, "I_LCA5", h3, True)                      # dead, :105
    hv_2 = tap("HV_LCA5", lca(hv_3, i_dec3, sd, "HV_LCA5", h3, False, taps))

    hv_2 = tap("hvd2", sa("sa_hv2", norm_upsample(hv_2, hv_jump1, sd, "HVD_block2")))
    i_dec2 = tap("id2", sa("sa_i2", norm_upsample(i_dec2, v_jump1, sd, "ID_block2")))

    i_dec1 = tap("I_LCA6", lca(i_dec2, hv_2, sd, "I_LCA6", h2, True, taps))
    hv_1 = tap("HV_LCA6", lca(hv_2, i_dec2, sd, "HV_LCA6", h2, False, taps))

    i_dec1 = tap("id1", sa("sa_i1", norm_upsample(i_dec1, i_jump0, sd, "ID_block1")))
    i_dec0 = tap("i_dec0", block0(i_dec1, sd["ID_block0.1.weight"]))
    hv_1 = tap("hvd1", sa("sa_hv1", norm_upsample(hv_1, hv_jump0, sd, "HVD_block1")))
    hv_0 = tap("hv_dec0", block0(hv_1, sd["HVD_block0.1.weight"]))

    out_hvi = tap("out_hvi", torch.cat([hv_0, i_dec0], dim=1) + hvi)
    return phvit(out_hvi, k, gated, alpha_s, gated2, alpha)


# --------------------------------------------------------------------------- #
# state_dict surface and deterministic weights
# --------------------------------------------------------------------------- #
def state_dict_spec(mssa: bool = False):
    """Ordered {key: shape} for the 191 fp32 tensors of net.CIDNet.CIDNet (+ the six
    `sa_*.conv1.weight [1,2,7,7]` of net.CIDNet_MSSA.CIDNet when `mssa`)."""
    c1, c2, c3, c4 = CHANNELS
    spec = {}
    spec["HVE_block0.1.weight"] = (c1, 3, 3, 3)
    for pfx in ("HVE", "IE"):
        pass
    def down(p, ci, co):
        spec[p + ".prelu.weight"] = (1,)
        spec[p + ".down.0.weight"] = (co, ci, 3, 3)
    def up(p, ci, co):
        spec[p + ".prelu.weight"] = (1,)
        spec[p + ".up_scale.0.weight"] = (co, ci, 3, 3)
        spec[p + ".up.weight"] = (co, 2 * co, 1, 1)
    def lca_spec(p, c, heads):
        h = int(c * 2.66)
        spec[p + ".norm.weight"] = (c,)
        spec[p + ".norm.bias"] = (c,)
        spec[p + ".gdfn.project_in.weight"] = (2 * h, c, 1, 1)
        spec[p + ".gdfn.dwconv.weight"] = (2 * h, 1, 3, 3)
        spec[p + ".gdfn.dwconv1.weight"] = (h, 1, 3, 3)
        spec[p + ".gdfn.dwconv2.weight"] = (h, 1, 3, 3)
        spec[p + ".gdfn.project_out.weight"] = (c, h, 1, 1)
        spec[p + ".ffn.temperature"] = (heads, 1, 1)
        spec[p + ".ffn.q.weight"] = (c, c, 1, 1)
        spec[p + ".ffn.q_dwconv.weight"] = (c, 1, 3, 3)
        spec[p + ".ffn.kv.weight"] = (2 * c, c, 1, 1)
        spec[p + ".ffn.kv_dwconv.weight"] = (2 * c, 1, 3, 3)
        spec[p + ".ffn.project_out.weight"] = (c, c, 1, 1)
    down("HVE_block1", c1, c2); down("HVE_block2", c2, c3); down("HVE_block3", c3, c4)
    up("HVD_block3", c4, c3); up("HVD_block2", c3, c2); up("HVD_block1", c2, c1)
    spec["HVD_block0.1.weight"] = (2, c1, 3, 3)
    spec["IE_block0.1.weight"] = (c1, 1, 3, 3)
    down("IE_block1", c1, c2); down("IE_block2", c2, c3); down("IE_block3", c3, c4)
    up("ID_block3", c4, c3); up("ID_block2", c3, c2); up("ID_block1", c2, c1)
    spec["ID_block0.1.weight"] = (1, c1, 3, 3)
    lvl = {1: (c2, HEADS[1]), 2: (c3, HEADS[2]), 3: (c4, HEADS[3]),
           4: (c4, HEADS[3]), 5: (c3, HEADS[2]), 6: (c2, HEADS[1])}
    for br in ("HV", "I"):
        for n in range(1, 7):
            lca_spec(f"{br}_LCA{n}", *lvl[n])
    spec["trans.density_k"] = (1,)
    assert len(spec) == 191
    if mssa:
        for name in ("sa_hv3", "sa_i3", "sa_hv2", "sa_i2", "sa_hv1", "sa_i1"):   # CIDNet_MSSA.py:93-98
            spec[name + ".conv1.weight"] = (1, 2, 7, 7)
    return spec


def make_state_dict(seed: int = 0, perturb: bool = True, mssa: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights that do not depend on torch's RNG (numpy
    PCG64, identical on every host).  Conv weights ~ U(-1/sqrt(fan_in), +) like
    PyTorch's default init.  With `perturb`, the parameters whose defaults are
    trivial (LN 1/0, temperature 1, PReLU 0.25, k 0.2) are moved off their
    defaults so the tests exercise them."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in state_dict_spec(mssa).items():
        if key.endswith("norm.weight"):
            a = 1.0 + (rng.uniform(-0.3, 0.3, shape) if perturb else 0.0) * np.ones(shape)
        elif key.endswith("norm.bias"):
            a = (rng.uniform(-0.2, 0.2, shape) if perturb else 0.0) * np.ones(shape)
        elif key.endswith("temperature"):
            a = rng.uniform(0.5, 3.0, shape) if perturb else np.ones(shape)
        elif key.endswith("prelu.weight"):
            a = rng.uniform(0.05, 0.4, shape) if perturb else np.full(shape, 0.25)
        elif key == "trans.density_k":
            a = np.full(shape, 0.37 if perturb else 0.2)
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, shape)
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape).copy())
    return sd


def make_input(kind: str, B: int, H: int, W: int, seed: int = 1234) -> torch.Tensor:
    """Synthetic inputs of SURVEY §8d: uniform / dark / 8-bit grid / constants."""
    import numpy as np
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        a = rng.random((B, 3, H, W), dtype=np.float32)
    elif kind == "dark":
        a = (rng.random((B, 3, H, W), dtype=np.float32) ** 3) * 0.3
    elif kind == "grid8":
        a = rng.integers(0, 256, (B, 3, H, W)).astype(np.float32) / 255.0
    elif kind == "grey8":       # exact greys and many channel ties
        g = rng.integers(0, 8, (B, 1, H, W)).astype(np.float32) / 7.0
        a = np.repeat(g, 3, axis=1)
        a[:, 1] = np.where(rng.random((B, H, W)) < 0.5, a[:, 1], rng.integers(0, 8, (B, H, W)) / 7.0)
    elif kind.startswith("const:"):
        a = np.full((B, 3, H, W), float(kind.split(":")[1]), dtype=np.float32)
    elif kind == "onehot":
        a = np.zeros((B, 3, H, W), dtype=np.float32)
        idx = rng.integers(0, 3, (B, H, W))
        for c in range(3):
            a[:, c] = (idx == c)
    else:
        raise ValueError(kind)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def pre_u8(img_u8_hwc: torch.Tensor, gamma: float = 1.0) -> torch.Tensor:
    """Caller-side pre-processing of the reference's eval loops, restated with the same torch calls:
    transforms.ToTensor (uint8 HWC -> float CHW / 255), reflect padding of the bottom / right edge to a
    multiple of 8 (data/eval_sets.py:22-27, demo.py:47-52) and `input ** gamma` (eval.py:64, demo.py:57).
    img_u8_hwc: [B,h,w,3] uint8 -> [B,3,H,W] float32."""
    x = img_u8_hwc.permute(0, 3, 1, 2).to(torch.float32).div(255)
    factor = 8
    h, w = x.shape[2], x.shape[3]
    H, W = ((h + factor) // factor) * factor, ((w + factor) // factor) * factor
    padh = H - h if h % factor != 0 else 0
    padw = W - w if w % factor != 0 else 0
    x = F.pad(x, (0, padw, 0, padh), "reflect")
    return x ** gamma


def post_u8(out: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """clamp(0,1) (eval.py:69), crop (eval.py:71), transforms.ToPILImage on a float tensor
    (= mul(255).byte(), truncation).  [B,3,H,W] float32 -> [B,h,w,3] uint8."""
    o = torch.clamp(out, 0, 1)[:, :, :h, :w]
    return o.mul(255).byte().permute(0, 2, 3, 1).contiguous()


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)
