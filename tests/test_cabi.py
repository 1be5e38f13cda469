"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/cidnet_b200.h declares (no compute calls here)."""
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    import hvi_cidnet_b200  # noqa: F401
    from hvi_cidnet_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "cidnet_b200.h")).read()
    declared = sorted(set(re.findall(r"CIDNET_API\s+[\w\s\*]+?\b(cidnet_\w+)\s*\(", hdr)))
    assert len(declared) >= 13
    h = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/cidnet_b200.h but not exported"


def test_ctypes_signatures_cover_header(built_lib):
    from hvi_cidnet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cidnet_b200.h")).read()
    declared = set(re.findall(r"CIDNET_API\s+[\w\s\*]+?\b(cidnet_\w+)\s*\(", hdr))
    assert declared <= set(_lib.SIGNATURES)
    h = _lib.lib()
    assert h.cidnet_abi_version() == 1
    assert h.cidnet_act_dtype() in (0, 1)


def test_ops_refuse_cpu_tensors(built_lib):
    import torch
    from hvi_cidnet_b200.net.HVI_transform import RGB_HVI
    m = RGB_HVI()
    assert list(m.state_dict().keys()) == ["density_k"]
    assert (m.gated, m.gated2, m.alpha, m.alpha_s, m.this_k) == (False, False, 1.0, 1.3, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.HVIT(torch.rand(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.PHVIT(torch.rand(1, 3, 8, 8))


def test_mssa_mirror_state_dict_surface():
    """net/CIDNet_MSSA.py mirror: exactly the 197 keys / shapes of the reference's MSSA class (golden key list
    dumped from /root/reference/net/CIDNet_MSSA.py by oracle/make_golden.py)."""
    import os
    from conftest import GOLDEN
    from hvi_cidnet_b200.net.CIDNet_MSSA import CIDNet
    lines = open(os.path.join(GOLDEN, "state_dict_keys_mssa.txt")).read().strip().splitlines()
    ref = {l.split(" ", 1)[0]: eval(l.split(" ", 1)[1]) for l in lines}
    sd = CIDNet().state_dict()
    assert set(sd.keys()) == set(ref.keys()) and len(sd) == 197      # order is irrelevant for dict loading
    for k, v in sd.items():
        assert list(v.shape) == ref[k], k
