"""CPU oracle for the BACKWARD of the HVI transform -- TEST INFRASTRUCTURE ONLY.

What the reference's training loop differentiates (`train.py:61-62`: `model.HVIT(output_rgb)` inside the loss;
`net/CIDNet.py:121`: `self.trans.PHVIT(output_hvi)` at the end of the forward) is whatever autograd derives from the
tensor program of `net/HVI_transform.py:16-47` (HVIT) and `:49-122` (PHVIT).  This file restates those vector-Jacobian
products in closed form, op by op, with the autograd conventions of the ops the reference uses:

  * `img.max(1)[0]`, `img.min(1)[0]` (:20-21)  -> the gradient goes to ONE channel, the first arg-max / arg-min (r before g
    before b on ties: torch's reduction returns the first extremal index on CPU);
  * `hue[mask] = expr[mask]` (:22-24)          -> only the LAST assignment that hits a pixel carries gradient
    (priority r==max > g==max > b==max), `hue[min==max] = 0` (:26) and `saturation[value==0] = 0` (:30) cut it;
  * `% 6`, `% 1` (:24, :66)                   -> derivative 1 w.r.t. the dividend;
  * `.pow(k)` with the PARAMETER k (:40)      -> d/dbase = k * base^(k-1), d/dk = result * log(base), summed over all pixels;
    `.pow(this_k)` with the python float (:60) -> d/dbase only, zero when k == 0;
  * `torch.clamp` (:55-57, :63-64, :72-73)    -> passes the gradient where min <= x <= max (inclusive);
  * `floor` (:80)                             -> zero gradient; the sextant masks (:85-114) route r, g, b to v / p / q / t.

Parity pinning: `oracle/make_golden.py::hvi_backward_cases` runs the UNMODIFIED reference under autograd in the build
container and commits inputs, upstream gradients and the resulting input / `density_k` gradients as
`tests/golden/hvi_backward.npz`; `tests/test_oracle_golden.py` holds this restatement to them on CPU.
Only `tests/` (and `__graft_entry__.smoke()`) may import this file.
"""
from __future__ import annotations

import torch

PI = 3.141592653589793
EPS = 1e-8


def hvit_backward(img: torch.Tensor, k: float, grad_hvi: torch.Tensor):
    """VJP of RGB_HVI.HVIT (net/HVI_transform.py:16-47).  Returns (grad_img [B,3,H,W], grad_k scalar tensor)."""
    dt = img.dtype
    r, g, b = img[:, 0], img[:, 1], img[:, 2]
    gH, gV, gI = grad_hvi[:, 0], grad_hvi[:, 1], grad_hvi[:, 2]
    value = torch.maximum(torch.maximum(r, g), b)
    vmin = torch.minimum(torch.minimum(r, g), b)
    d = value - vmin + EPS

    grey = vmin == value
    br = (r == value) & ~grey                      # branch that wrote the hue last (:22-26)
    bg = (g == value) & ~grey & ~br
    bb = ~grey & ~br & ~bg
    num = torch.where(br, g - b, torch.where(bg, b - r, r - g))
    base = torch.where(br, torch.zeros_like(r), torch.where(bg, torch.full_like(r, 2.0), torch.full_like(r, 4.0)))
    hue = base + num / d
    hue = torch.where(br, torch.remainder(hue, 6), hue)
    hue = torch.where(grey, torch.zeros_like(hue), hue) / 6.0

    sat = (value - vmin) / (value + EPS)
    sat_live = value != 0
    sat = torch.where(sat_live, sat, torch.zeros_like(sat))
    ang = value * 0.5 * PI
    sbase = ang.sin() + EPS
    kt = torch.tensor(float(k), dtype=dt)
    cs = sbase.pow(kt)
    ch = (2.0 * PI * hue).cos()
    cv = (2.0 * PI * hue).sin()

    # H = (cs * sat) * ch ; V = (cs * sat) * cv ; I = value          (:43-45)
    css = cs * sat
    g_css = gH * ch + gV * cv
    g_cs = g_css * sat
    g_sat = torch.where(sat_live, g_css * cs, torch.zeros_like(cs))
    g_hue = (gV * ch - gH * cv) * css * (2.0 * PI) / 6.0
    g_hue = torch.where(grey, torch.zeros_like(g_hue), g_hue)

    # hue = base + num / d
    g_num = g_hue / d
    g_d = -g_hue * num / (d * d)
    g_value = gI + g_d
    g_min = -g_d
    # sat = (value - vmin) / (value + eps)
    ve = value + EPS
    g_value = g_value + g_sat / ve - g_sat * (value - vmin) / (ve * ve)
    g_min = g_min - g_sat / ve
    # cs = (sin(value * pi/2) + eps) ** k
    g_value = g_value + g_cs * kt * sbase.pow(kt - 1.0) * ang.cos() * (0.5 * PI)
    g_k = (g_cs * cs * sbase.log()).sum()

    zero = torch.zeros_like(r)
    gr = torch.where(bg, -g_num, torch.where(bb, g_num, zero))        # bg: num = b - r ; bb: num = r - g
    gg = torch.where(br, g_num, torch.where(bb, -g_num, zero))        # br: num = g - b
    gb = torch.where(br, -g_num, torch.where(bg, g_num, zero))
    # value / vmin -> first extremal channel
    amax_r = r == value
    amax_g = ~amax_r & (g == value)
    amax_b = ~amax_r & ~amax_g
    amin_r = r == vmin
    amin_g = ~amin_r & (g == vmin)
    amin_b = ~amin_r & ~amin_g
    gr = gr + torch.where(amax_r, g_value, zero) + torch.where(amin_r, g_min, zero)
    gg = gg + torch.where(amax_g, g_value, zero) + torch.where(amin_g, g_min, zero)
    gb = gb + torch.where(amax_b, g_value, zero) + torch.where(amin_b, g_min, zero)
    return torch.stack([gr, gg, gb], dim=1), g_k


def phvit_backward(img: torch.Tensor, k: float, grad_rgb: torch.Tensor, gated: bool = False, alpha_s: float = 1.3,
                   gated2: bool = False, alpha: float = 1.0) -> torch.Tensor:
    """VJP of RGB_HVI.PHVIT (net/HVI_transform.py:49-122) w.r.t. its input; `k` = this_k, a python float (no gradient)."""
    H0, V0, I0 = img[:, 0], img[:, 1], img[:, 2]
    gr, gg, gb = grad_rgb[:, 0], grad_rgb[:, 1], grad_rgb[:, 2]
    if gated2:
        gr, gg, gb = gr * alpha, gg * alpha, gb * alpha
    Hc, Vc, v = H0.clamp(-1, 1), V0.clamp(-1, 1), I0.clamp(0, 1)
    ang = v * 0.5 * PI
    sbase = ang.sin() + EPS
    cs = sbase.pow(k)
    den = cs + EPS
    H2, V2 = Hc / den, Vc / den
    H3, V3 = H2.clamp(-1, 1), V2.clamp(-1, 1)
    x, y = H3 + EPS, V3 + EPS
    h = torch.atan2(y, x) / (2 * PI)
    h = h % 1
    s_raw = torch.sqrt(H3 ** 2 + V3 ** 2 + EPS)
    s_pre = s_raw * alpha_s if gated else s_raw
    s = s_pre.clamp(0, 1)
    hi = torch.floor(h * 6.0)
    f = h * 6.0 - hi

    zero = torch.zeros_like(h)
    # sextant table :92-114: which of (v, p, q, t) each colour channel was copied from
    #            r  g  b
    table = {0: ("v", "t", "p"), 1: ("q", "v", "p"), 2: ("p", "v", "t"), 3: ("p", "q", "v"), 4: ("t", "p", "v"),
             5: ("v", "p", "q")}
    acc = {"v": zero, "p": zero, "q": zero, "t": zero}
    for n, (sr, sg, sb) in table.items():
        m = hi == n
        acc[sr] = acc[sr] + torch.where(m, gr, zero)
        acc[sg] = acc[sg] + torch.where(m, gg, zero)
        acc[sb] = acc[sb] + torch.where(m, gb, zero)
    g_v, g_p, g_q, g_t = acc["v"], acc["p"], acc["q"], acc["t"]
    # p = v (1 - s) ; q = v (1 - f s) ; t = v (1 - (1 - f) s)         (:82-84)
    g_v = g_v + g_p * (1.0 - s) + g_q * (1.0 - f * s) + g_t * (1.0 - (1.0 - f) * s)
    g_s = -g_p * v - g_q * v * f - g_t * v * (1.0 - f)
    g_f = (g_t - g_q) * v * s
    # s = clamp(s_pre, 0, 1) ; s_pre = s_raw [* alpha_s] ; s_raw = sqrt(H3^2 + V3^2 + eps)
    g_s = torch.where((s_pre >= 0) & (s_pre <= 1), g_s, zero)
    if gated:
        g_s = g_s * alpha_s
    g_H3 = g_s * H3 / s_raw
    g_V3 = g_s * V3 / s_raw
    # f = 6 h - hi ; h = atan2(y, x) / 2pi (the % 1 passes the gradient)
    g_h = 6.0 * g_f / (2 * PI)
    r2 = x * x + y * y
    g_V3 = g_V3 + g_h * x / r2
    g_H3 = g_H3 - g_h * y / r2
    # H3 = clamp(H2) ; H2 = Hc / den
    g_H2 = torch.where((H2 >= -1) & (H2 <= 1), g_H3, zero)
    g_V2 = torch.where((V2 >= -1) & (V2 <= 1), g_V3, zero)
    g_Hc = g_H2 / den
    g_Vc = g_V2 / den
    g_den = -(g_H2 * Hc + g_V2 * Vc) / (den * den)
    # den = cs + eps ; cs = sbase ** k (python float k: zero gradient when k == 0)
    if float(k) != 0.0:
        g_v = g_v + g_den * k * sbase.pow(k - 1.0) * ang.cos() * (0.5 * PI)
    gH = torch.where((H0 >= -1) & (H0 <= 1), g_Hc, zero)
    gV = torch.where((V0 >= -1) & (V0 <= 1), g_Vc, zero)
    gI = torch.where((I0 >= 0) & (I0 <= 1), g_v, zero)
    return torch.stack([gH, gV, gI], dim=1)


def phvit_sextant_margin(img: torch.Tensor, k: float) -> torch.Tensor:
    """Distance of 6*h from the nearest integer, per pixel [B,H,W] (net/HVI_transform.py:66,80): r, g, b are continuous
    across a sextant boundary but their DERIVATIVES are not, so two correct fp32 evaluations that round `h` one ulp apart
    may legitimately disagree on the gradient of a pixel that sits on a boundary.  Tests excuse such pixels only when
    this margin (computed here, by the oracle) is below 2e-6."""
    Hc, Vc, v = img[:, 0].clamp(-1, 1), img[:, 1].clamp(-1, 1), img[:, 2].clamp(0, 1)
    den = ((v * 0.5 * PI).sin() + EPS).pow(k) + EPS
    H3, V3 = (Hc / den).clamp(-1, 1), (Vc / den).clamp(-1, 1)
    h6 = (torch.atan2(V3 + EPS, H3 + EPS) / (2 * PI) % 1) * 6.0
    return (h6 - torch.round(h6)).abs()
