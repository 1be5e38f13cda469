"""The -DCIDNET_ACT_BF16 build (libcidnet_b200_bf16.so: bf16 activations / tensor-core operands, fp32-accumulating
depthwise and IEL kernels) run through the same forward parity checks in a fresh interpreter (the library is chosen
at import time by CIDNET_LIB).  bf16 has fp32's range (no 65504 ceiling) but only 8 bits of mantissa: measured on B200
2.1e-3 max-abs at 200x304 (SURVEY App. E predicted 1.5e-3), i.e. it does NOT meet the 2e-3 contract the default fp16
build meets with a 10x margin.  Its stated tolerance is therefore 4e-3 max-abs / 50 dB; it exists for weights whose
activations would exceed fp16's range, not as the default."""
BF16_MAXABS = 4e-3
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r'''
import json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import torch
from oracle import cidnet_oracle as O
from conftest import parity_error, psnr_kept
import hvi_cidnet_b200
from hvi_cidnet_b200 import _lib
from hvi_cidnet_b200.net.CIDNet import CIDNet
assert _lib.lib().cidnet_act_dtype() == 1, "not the bf16 build"
torch.set_grad_enabled(False)
out = []
m = CIDNet().cuda().eval()
for seed, kind, shape in ((5, "uniform", (1, 200, 304)), (0, "dark", (2, 64, 96)), (7, "grid8", (1, 104, 72))):
    sd = O.make_state_dict(seed, True)
    m.load_state_dict(sd, strict=True)
    x = O.make_input(kind, *shape, seed=21)
    taps = {}
    ref = O.forward(x, sd, taps=taps).clamp(0, 1)
    y = m(x.cuda()).cpu().clamp(0, 1)
    y2 = m(x.cuda()).cpu().clamp(0, 1)
    err, excused, keep = parity_error(y, ref, 4e-3, taps["out_hvi"], float(sd["trans.density_k"][0]))
    out.append({"kind": kind, "err": err, "psnr": psnr_kept(y, ref, keep), "excused": excused, "repeat_equal": bool(torch.equal(y, y2))})
print("RESULT " + json.dumps(out))
'''


def test_bf16_library_meets_the_contract():
    lib = os.path.join(ROOT, "hvi-cidnet_b200", "libcidnet_b200_bf16.so")
    if not os.path.exists(lib):
        from hvi_cidnet_b200 import build
        build.build(bf16=True)
    env = dict(os.environ, CIDNET_LIB=lib)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    print("bf16 build:", res)
    for e in res:
        assert e["err"] <= BF16_MAXABS and e["psnr"] >= 50.0 and e["repeat_equal"], e
