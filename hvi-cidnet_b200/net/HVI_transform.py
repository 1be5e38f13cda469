"""Host-side mirror of the reference's `net.HVI_transform.RGB_HVI`
(/root/reference/net/HVI_transform.py:6-122): same attributes (`density_k`,
`gated`, `gated2`, `alpha`, `alpha_s`, `this_k`), same `HVIT` / `PHVIT` methods,
same statefulness (`PHVIT` uses the `this_k` cached by the last `HVIT`, 0 on a
fresh module).  The arithmetic is one CUDA kernel per call in libcidnet_b200.so.

The reference reads `density_k` with `k.item()` on every HVIT call (a host sync,
:38).  Here the kernels read k straight from device memory; `this_k` is kept as a
device-side snapshot and only turned into a Python float when somebody reads it.

Training-side use (train.py:61-62: `model.HVIT(output_rgb)` inside the loss;
net/CIDNet.py:121: PHVIT at the end of the forward): when autograd is recording and
the image or `density_k` requires a gradient, HVIT / PHVIT run as
`torch.autograd.Function`s whose backward is ONE kernel each
(`cidnet_hvit_backward` / `cidnet_phvit_backward`, csrc/hvi_bwd.cu) with the
conventions autograd applies to the reference's tensor program (first arg-max /
arg-min channel, last masked hue assignment wins, closed-interval clamps, d/dk of
`pow` summed over every pixel).  Under `torch.no_grad()` nothing is recorded.
"""
import torch
import torch.nn as nn

from .. import _lib

pi = 3.141592653589793


def _hvit_launch(x, k_dev):
    out = torch.empty_like(x)
    B, _, H, W = x.shape
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cidnet_hvit(x.data_ptr(), out.data_ptr(), B, H, W, 0.0, k_dev.data_ptr(),
                                          _lib.stream_ptr(x.device)))
    return out


def _phvit_launch(x, k_val, k_dev, gated, alpha_s, gated2, alpha):
    out = torch.empty_like(x)
    B, _, H, W = x.shape
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cidnet_phvit(x.data_ptr(), out.data_ptr(), B, H, W, k_val,
                                           k_dev.data_ptr() if k_dev is not None else None,
                                           gated, alpha_s, gated2, alpha, _lib.stream_ptr(x.device)))
    return out


class _HVITFunction(torch.autograd.Function):
    """HVIT with the backward autograd derives from net/HVI_transform.py:16-47 (one kernel; the forward saves only its input)."""

    @staticmethod
    def forward(ctx, img, density_k, k_dev):
        ctx.save_for_backward(img, k_dev)
        ctx.k_like = (density_k.shape, density_k.dtype, density_k.device)
        return _hvit_launch(img, k_dev)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_hvi):
        img, k_dev = ctx.saved_tensors
        g = _lib.require_cuda_f32(grad_hvi, "grad_hvi")
        grad_img = torch.empty_like(img)
        want_k = ctx.needs_input_grad[1]
        B, _, H, W = img.shape
        L = _lib.lib()
        gk = scratch = None
        if want_k:
            gk = torch.empty(1, dtype=torch.float32, device=img.device)
            scratch = torch.empty(int(L.cidnet_hvi_backward_scratch_bytes()) // 4, dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            _lib.check(L.cidnet_hvit_backward(img.data_ptr(), g.data_ptr(), grad_img.data_ptr(),
                                              gk.data_ptr() if want_k else None, scratch.data_ptr() if want_k else None,
                                              B, H, W, 0.0, k_dev.data_ptr(), _lib.stream_ptr(img.device)))
        if want_k:
            shape, dtype, device = ctx.k_like
            gk = gk.reshape(shape).to(device=device, dtype=dtype)
        return (grad_img if ctx.needs_input_grad[0] else None), gk, None


class _PHVITFunction(torch.autograd.Function):
    """PHVIT with the backward autograd derives from net/HVI_transform.py:49-122 (`this_k` is a python float there: no
    gradient reaches density_k through PHVIT)."""

    @staticmethod
    def forward(ctx, img, k_val, k_dev, gated, alpha_s, gated2, alpha):
        ctx.save_for_backward(img) if k_dev is None else ctx.save_for_backward(img, k_dev)
        ctx.cfg = (k_val, gated, alpha_s, gated2, alpha)
        return _phvit_launch(img, k_val, k_dev, gated, alpha_s, gated2, alpha)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_rgb):
        saved = ctx.saved_tensors
        img, k_dev = saved[0], (saved[1] if len(saved) > 1 else None)
        k_val, gated, alpha_s, gated2, alpha = ctx.cfg
        g = _lib.require_cuda_f32(grad_rgb, "grad_rgb")
        grad_img = torch.empty_like(img)
        B, _, H, W = img.shape
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().cidnet_phvit_backward(img.data_ptr(), g.data_ptr(), grad_img.data_ptr(), B, H, W, k_val,
                                                        k_dev.data_ptr() if k_dev is not None else None,
                                                        gated, alpha_s, gated2, alpha, _lib.stream_ptr(img.device)))
        return grad_img, None, None, None, None, None, None


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
        if torch.is_grad_enabled() and (img.requires_grad or self.density_k.requires_grad):
            return _HVITFunction.apply(x, self.density_k, self._this_k_dev)
        return _hvit_launch(x, self._this_k_dev)

    def PHVIT(self, img):
        x = self._check(img, "PHVIT")
        kd = self._this_k_dev
        if kd is not None and kd.device != x.device:
            kd = kd.to(x.device)
        args = (float(self._this_k_value), kd, int(bool(self.gated)), float(self.alpha_s), int(bool(self.gated2)),
                float(self.alpha))
        if torch.is_grad_enabled() and img.requires_grad:
            return _PHVITFunction.apply(x, *args)
        return _phvit_launch(x, *args)
