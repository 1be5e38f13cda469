import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


def max_err_robust(y, ref, max_outlier_px=3):
    """max |y - ref| over all pixels except at most `max_outlier_px` of them.

    The reference's PHVIT (net/HVI_transform.py:63-107) is discontinuous on a measure-zero set: a hue that
    lands exactly on the wrap (h*6 == 6 after the `% 1`) matches none of the six sextant masks and the pixel
    comes out BLACK.  A last-bit difference upstream (e.g. the order of the Gram's fp32 atomics, which is run
    dependent) can move a pixel on or off that set, changing it by ~1.0 while every other pixel agrees to 1e-4.
    Observed about once per hundred 400x600 forwards.  Such pixels are excluded (and counted) here."""
    import torch
    d = (y - ref).abs()
    if d.dim() == 4:
        d = d.amax(dim=1)                       # per pixel, over channels
    flat = d.flatten()
    k = min(max_outlier_px, flat.numel() - 1)
    if k <= 0:
        return float(flat.max())
    top = torch.topk(flat, k + 1).values
    return float(top[-1])
