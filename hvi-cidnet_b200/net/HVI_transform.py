"""Host-side mirror of the reference's `net.HVI_transform.RGB_HVI`
(/root/reference/net/HVI_transform.py:6-122): same attributes (`density_k`,
`gated`, `gated2`, `alpha`, `alpha_s`, `this_k`), same `HVIT` / `PHVIT` methods,
same statefulness (`PHVIT` uses the `this_k` cached by the last `HVIT`, 0 on a
fresh module).  The arithmetic is one CUDA kernel per call in libcidnet_b200.so.

The reference reads `density_k` with `k.item()` on every HVIT call (a host sync,
:38).  Here the kernels read k straight from device memory; `this_k` is kept as a
device-side snapshot and only turned into a Python float when somebody reads it.
"""
import torch
import torch.nn as nn

from .. import _lib

pi = 3.141592653589793


class RGB_HVI(nn.Module):
    def __init__(self):
        super(RGB_HVI, self).__init__()
        self.density_k = torch.nn.Parameter(torch.full([1], 0.2))  # reference :9
        self.gated = False
        self.gated2 = False
        self.alpha = 1.0
        self.alpha_s = 1.3
        self._this_k_value = 0      # python number assigned by the user / initial 0 (reference :14)
        self._this_k_dev = None     # device snapshot taken by the last HVIT

    # `this_k`: float, set by HVIT, read by PHVIT (reference :38,:59)
    @property
    def this_k(self):
        if self._this_k_dev is not None:
            self._this_k_value = float(self._this_k_dev.item())
            self._this_k_dev = None
        return self._this_k_value

    @this_k.setter
    def this_k(self, v):
        self._this_k_value = v
        self._this_k_dev = None

    def _note_hvit_called(self):
        self._this_k_dev = self.density_k.detach().float().clone()

    def _check(self, img, who):
        x = _lib.require_cuda_f32(img, "img")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"{who} expects [B,3,H,W], got {tuple(x.shape)}")
        return x

    def HVIT(self, img):
        x = self._check(img, "HVIT")
        k = self.density_k.detach()
        if k.device != x.device or k.dtype != torch.float32:
            k = k.to(x.device, torch.float32)
        self._this_k_dev = k.clone()                               # reference :38
        out = torch.empty_like(x)
        B, _, H, W = x.shape
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cidnet_hvit(x.data_ptr(), out.data_ptr(), B, H, W, 0.0, self._this_k_dev.data_ptr(),
                                              _lib.stream_ptr(x.device)))
        return out

    def PHVIT(self, img):
        x = self._check(img, "PHVIT")
        out = torch.empty_like(x)
        B, _, H, W = x.shape
        kd = self._this_k_dev
        if kd is not None and kd.device != x.device:
            kd = kd.to(x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cidnet_phvit(x.data_ptr(), out.data_ptr(), B, H, W,
                                               float(self._this_k_value), kd.data_ptr() if kd is not None else None,
                                               int(bool(self.gated)), float(self.alpha_s),
                                               int(bool(self.gated2)), float(self.alpha),
                                               _lib.stream_ptr(x.device)))
        return out
