"""Host-side mirror of the reference's `net.CIDNet.CIDNet` (/root/reference/net/CIDNet.py:8-126).

Drop-in surface kept: constructor `CIDNet(channels=[36,36,72,144], heads=[1,2,4,8], norm=False)`,
`forward(x)`, `HVIT(x)`, the `trans` sub-module with its mutable knobs, the 191-tensor fp32
`state_dict` (strict-loadable from the reference's `.pth` files) and the
`PyTorchModelHubMixin` methods.  The sub-modules below are PARAMETER CONTAINERS only -- they
give the parameters the reference's names, shapes and default initialisation, and are never
called.  All arithmetic of `forward` runs in libcidnet_b200.so (sm_100a CUDA, C ABI in
include/cidnet_b200.h); there is no PyTorch / CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from .HVI_transform import RGB_HVI

try:  # the reference mixes this in (CIDNet.py:6,8); keep from_pretrained/save_pretrained when available
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

_CHANNELS = [36, 36, 72, 144]
_HEADS = [1, 2, 4, 8]


class _Norm(nn.Module):
    """parameters of transformer_utils.LayerNorm (:14-15)"""
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


def _conv(ci, co, k, groups=1):
    return nn.Conv2d(ci, co, kernel_size=k, stride=1, padding=k // 2, groups=groups, bias=False)


class _CAB(nn.Module):
    """parameters of LCA.CAB (:8-17)"""
    def __init__(self, dim, heads):
        super().__init__()
        self.temperature = nn.Parameter(torch.ones(heads, 1, 1))
        self.q, self.q_dwconv = _conv(dim, dim, 1), _conv(dim, dim, 3, groups=dim)
        self.kv, self.kv_dwconv = _conv(dim, dim * 2, 1), _conv(dim * 2, dim * 2, 3, groups=dim * 2)
        self.project_out = _conv(dim, dim, 1)


class _IEL(nn.Module):
    """parameters of LCA.IEL (:46-57)"""
    def __init__(self, dim):
        super().__init__()
        h = int(dim * 2.66)
        self.project_in = _conv(dim, h * 2, 1)
        self.dwconv = _conv(h * 2, h * 2, 3, groups=h * 2)
        self.dwconv1, self.dwconv2 = _conv(h, h, 3, groups=h), _conv(h, h, 3, groups=h)
        self.project_out = _conv(h, dim, 1)


class _LCA(nn.Module):
    """HV_LCA / I_LCA (:71-93): `ffn` is the attention, `gdfn` the IEL (reference naming)."""
    def __init__(self, dim, heads):
        super().__init__()
        self.gdfn, self.norm, self.ffn = _IEL(dim), _Norm(dim), _CAB(dim, heads)


class _Down(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.prelu = nn.PReLU()
        self.down = nn.Sequential(_conv(ci, co, 3), nn.Identity())


class _Up(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.prelu = nn.PReLU()
        self.up_scale = nn.Sequential(_conv(ci, co, 3), nn.Identity())
        self.up = _conv(co * 2, co, 1)


def _block0(ci, co):
    return nn.Sequential(nn.Identity(), nn.Conv2d(ci, co, 3, stride=1, padding=0, bias=False))


class CIDNet(nn.Module, PyTorchModelHubMixin):
    _variant = 0          # CIDNET_VARIANT_BASE; net/CIDNet_MSSA.py's mirror overrides it

    def __init__(self, channels=[36, 36, 72, 144], heads=[1, 2, 4, 8], norm=False):
        super(CIDNet, self).__init__()
        if list(channels) != _CHANNELS or list(heads) != _HEADS:
            raise NotImplementedError("the B200 kernels are specialised for channels=[36,36,72,144], heads=[1,2,4,8] "
                                      "(the only configuration the reference's call sites and weights use)")
        if norm:
            raise NotImplementedError("norm=True (extra LayerNorms, transformer_utils.py:35-36,55) has no shipped weights")
        c1, c2, c3, c4 = channels
        for br, cin0, cout0 in (("HV", 3, 2), ("I", 1, 1)):
            setattr(self, f"{br}E_block0", _block0(cin0, c1))
            for n, (ci, co) in enumerate(((c1, c2), (c2, c3), (c3, c4)), start=1):
                setattr(self, f"{br}E_block{n}", _Down(ci, co))
            for n, (ci, co) in ((3, (c4, c3)), (2, (c3, c2)), (1, (c2, c1))):
                setattr(self, f"{br}D_block{n}", _Up(ci, co))
            setattr(self, f"{br}D_block0", _block0(c1, cout0))
        for br in ("HV", "I"):
            for n, lvl in ((1, 1), (2, 2), (3, 3), (4, 3), (5, 2), (6, 1)):
                setattr(self, f"{br}_LCA{n}", _LCA(channels[lvl], heads[lvl]))
        self.trans = RGB_HVI()
        # ---- native state (not part of the state_dict) ----
        self._ctx = None
        self._ctx_device = None
        self._synced = None
        self._workspaces = {}

    # ------------------------------------------------------------------ native context
    def _weights_signature(self):
        """Changes whenever a packed parameter may have changed: in-place updates bump `_version`, re-assignment
        (load_state_dict(assign=True), .to(), .half()) changes `data_ptr`.  density_k is excluded on purpose: the
        kernels read it from device memory on every call.  The parameter list is cached (walking named_parameters()
        costs more than the 190 attribute reads below)."""
        ps = self.__dict__.get("_packed_params")
        if ps is None or len(ps) != sum(1 for _ in self.parameters()) - 1:
            ps = [p for n, p in self.named_parameters() if n != "trans.density_k"]
            self.__dict__["_packed_params"] = ps
        return hash(tuple((p._version, p.data_ptr()) for p in ps))

    def _ensure_ctx(self, device):
        lib = _lib.lib()
        if self._ctx is None or self._ctx_device != device:
            self._release()
            ctx = C.c_void_p()
            _lib.check(lib.cidnet_create(C.byref(ctx), device.index if device.index is not None else torch.cuda.current_device()))
            _lib.check(lib.cidnet_set_variant(ctx, self._variant))
            self._ctx, self._ctx_device, self._synced = ctx, device, None
        sig = self._weights_signature()
        if self._synced != sig:
            self.sync_weights()
            self._synced = sig
        return self._ctx

    def sync_weights(self):
        """(re)pack the current parameters into the kernels' device layouts.  Called automatically
        when a parameter's version changes (load_state_dict, optimizer steps); call it by hand after
        editing a parameter through `.data`."""
        lib = _lib.lib()
        if self._ctx is None:
            return
        for name, p in self.state_dict().items():
            t = p.detach().to("cpu", torch.float32).contiguous()
            _lib.check(lib.cidnet_set_weight(self._ctx, name.encode(), t.data_ptr(), t.numel()))
        with torch.cuda.device(self._ctx_device):
            _lib.check(lib.cidnet_finalize_weights(self._ctx))

    # copy.deepcopy / pickle / torch.save(model): the native handle and the workspaces stay behind; the copy lazily
    # creates its own context on its first forward (the reference nn.Module supports all three)
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_ctx"], st["_ctx_device"], st["_synced"], st["_workspaces"] = None, None, None, {}
        st.pop("_packed_params", None)
        return st

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _apply(self, fn, *a, **kw):
        # .to() / .cuda() / .half() replace parameter storage: forget the cached parameter list
        self.__dict__.pop("_packed_params", None)
        return super()._apply(fn, *a, **kw)

    def _release(self):
        ctx = self.__dict__.get("_ctx")
        if ctx is not None:
            try:
                _lib.lib().cidnet_destroy(ctx)
            except Exception:
                pass
        self.__dict__["_ctx"] = None
        self.__dict__["_workspaces"] = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _workspace(self, B, H, W, device):
        key = (B, H, W, str(device))
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = _lib.lib().cidnet_workspace_bytes(B, H, W)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._workspaces = {key: ws}        # keep only the latest shape
        off = (-ws.data_ptr()) % 1024
        return ws, ws.data_ptr() + off, ws.numel() - off

    # ------------------------------------------------------------------ reference surface
    def forward(self, x, out=None):
        """`out` (optional, extension): a CUDA fp32 tensor of x's shape to write the result into (the
        streamed driver keeps a ring of them); by default a fresh tensor is returned like the reference."""
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError(f"CIDNet.forward: input is on {getattr(x, 'device', None)}; the B200-native path has no "
                               "CPU fallback -- move the model and the input to an sm_100 CUDA device")
        dtypes = x.dtype
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"CIDNet.forward expects [B,3,H,W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise RuntimeError(f"Sizes of tensors must match: H and W must be multiples of 8, got {H}x{W} "
                               "(the reference fails in torch.cat at transformer_utils.py:64; pad first)")
        k = self.trans.density_k
        if k.device != x.device:
            raise RuntimeError(f"model parameters are on {k.device} but the input is on {x.device}")
        xin = x.contiguous() if dtypes == torch.float32 else x.float().contiguous()
        if out is None:
            out = torch.empty_like(xin)
        elif (out.shape != xin.shape or out.dtype != torch.float32 or out.device != xin.device
              or not out.is_contiguous() or dtypes != torch.float32):
            raise RuntimeError("CIDNet.forward: `out` must be a contiguous fp32 tensor of the input's shape and device")
        if B == 0:
            return out.to(dtypes)
        with torch.cuda.device(x.device):
            ctx = self._ensure_ctx(x.device)
            ws, ws_ptr, ws_bytes = self._workspace(B, H, W, x.device)
            t = self.trans
            t._note_hvit_called()                                   # this_k = k.item(), lazily (HVI_transform.py:38)
            # the kernels read k from device memory: the live parameter when it is fp32, else the fp32 snapshot just taken
            # (model.half() / .double()), never a value cached at the last weight sync
            kd = k.detach() if k.dtype == torch.float32 else t._this_k_dev
            kptr = kd.data_ptr()
            _lib.check(_lib.lib().cidnet_forward(ctx, xin.data_ptr(), out.data_ptr(), B, H, W, ws_ptr, ws_bytes, kptr,
                                                 int(bool(t.gated)), float(t.alpha_s), int(bool(t.gated2)),
                                                 float(t.alpha), _lib.stream_ptr(x.device)))
        return out if dtypes == torch.float32 else out.to(dtypes)

    def enhance_u8(self, img, gamma=1.0, out=None):
        """8-bit path (extension): img CUDA uint8 [B,h,w,3] (HWC, any h,w) -> enhanced uint8 [B,h,w,3].
        Equivalent to the reference's eval loop around the model (ToTensor, reflect-pad to a multiple of 8, **gamma,
        forward, clamp(0,1), crop, ToPILImage: eval.py:56-73, data/eval_sets.py:22-27); the conversions are fused into
        the first and the last kernel of the forward (cidnet_forward_u8), so the padded fp32 image never exists."""
        if not isinstance(img, torch.Tensor) or not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] != 3:
            raise RuntimeError("enhance_u8 expects a CUDA uint8 tensor [B,h,w,3] (no CPU fallback)")
        B, h, w, _ = img.shape
        H, W = -(-h // 8) * 8, -(-w // 8) * 8
        img = img.contiguous()
        if out is None:
            out = torch.empty_like(img)
        elif out.shape != img.shape or out.dtype != torch.uint8 or out.device != img.device or not out.is_contiguous():
            raise RuntimeError("enhance_u8: `out` must be a contiguous uint8 tensor of the input's shape and device")
        if B == 0:
            return out
        with torch.cuda.device(img.device):
            ctx = self._ensure_ctx(img.device)
            ws, ws_ptr, ws_bytes = self._workspace(B, H, W, img.device)
            t = self.trans
            t._note_hvit_called()
            kd = t.density_k.detach() if t.density_k.dtype == torch.float32 else t._this_k_dev
            _lib.check(_lib.lib().cidnet_forward_u8(ctx, img.data_ptr(), out.data_ptr(), B, h, w, float(gamma), ws_ptr, ws_bytes,
                                                    kd.data_ptr(), int(bool(t.gated)), float(t.alpha_s), int(bool(t.gated2)),
                                                    float(t.alpha), _lib.stream_ptr(img.device)))
        return out

    def HVIT(self, x):
        hvi = self.trans.HVIT(x)
        return hvi

    # ------------------------------------------------------------------ test / profiling helpers
    def read_tap(self, name):
        """fp32 NCHW copy of a named internal activation of the LAST forward (parity tests)."""
        lib = _lib.lib()
        dims = (C.c_int * 3)()
        _lib.check(lib.cidnet_read_tap(self._ctx, name.encode(), None, 0, dims, None))
        Cc, H, W = dims[0], dims[1], dims[2]
        B = next(iter(self._workspaces))[0]
        out = torch.empty(B, Cc, H, W, device=self._ctx_device)
        with torch.cuda.device(self._ctx_device):
            _lib.check(lib.cidnet_read_tap(self._ctx, name.encode(), out.data_ptr(), out.numel(), dims,
                                           _lib.stream_ptr(self._ctx_device)))
        return out

    def run_lca_stage(self, n, x_i, x_hv, stat_rows=None):
        """Unit-test helper (cidnet_test_lca_stage): LCA stage n (1..6) alone on fp32 CUDA tensors [B,C,H,W] at the
        stage's own resolution.  Returns dict(after_cab_i, after_cab_hv, out_i, out_hv); the I entries are None for
        the base graph's dead I_LCA5.  stat_rows=(y0, y1) restricts the attention statistics to those rows."""
        lib = _lib.lib()
        dev = x_i.device
        x_i, x_hv = _lib.require_cuda_f32(x_i, "x_i"), _lib.require_cuda_f32(x_hv, "x_hv")
        with torch.cuda.device(dev):
            ctx = self._ensure_ctx(dev)
            B, _, H, W = x_i.shape
            outs = [torch.empty_like(x_i) for _ in range(4)]
            y0, y1 = stat_rows if stat_rows is not None else (0, 0)
            _lib.check(lib.cidnet_test_lca_stage(ctx, int(n), x_i.data_ptr(), x_hv.data_ptr(), outs[0].data_ptr(),
                                                 outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), B, H, W,
                                                 int(y0), int(y1), _lib.stream_ptr(dev)))
        i_live = not (n == 5 and self._variant == 0)
        return {"after_cab_i": outs[0] if i_live else None, "after_cab_hv": outs[1],
                "out_i": outs[2] if i_live else None, "out_hv": outs[3]}

    def set_profiling(self, enable):
        """record a CUDA event before every kernel launch of forward() (see cidnet_profile_*)."""
        if self._ctx is None:
            raise RuntimeError("run one forward first")
        _lib.check(_lib.lib().cidnet_profile_enable(self._ctx, int(bool(enable))))

    def read_profile(self, spans=False):
        """[(name, ms, algorithmic_bytes, flops)] of the last profiled forward (synchronises); with `spans` each tuple
        also carries (start_ms, end_ms) relative to the forward's first launch (pairs overlap on two streams)."""
        lib = _lib.lib()
        torch.cuda.synchronize(self._ctx_device)
        out = []
        buf = C.create_string_buffer(96)
        ms, by, fl, t0, t1 = C.c_float(), C.c_double(), C.c_double(), C.c_float(), C.c_float()
        for i in range(lib.cidnet_profile_count(self._ctx)):
            _lib.check(lib.cidnet_profile_get(self._ctx, i, buf, 96, C.byref(ms), C.byref(by), C.byref(fl)))
            rec = (buf.value.decode(), float(ms.value), float(by.value), float(fl.value))
            if spans:
                _lib.check(lib.cidnet_profile_get_span(self._ctx, i, C.byref(t0), C.byref(t1)))
                rec += (float(t0.value), float(t1.value))
            out.append(rec)
        return out

    def num_launches(self):
        return int(_lib.lib().cidnet_forward_launches(self._ctx)) if self._ctx is not None else 0
