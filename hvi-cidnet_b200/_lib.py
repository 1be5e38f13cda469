"""ctypes binding of libcidnet_b200.so (C ABI: include/cidnet_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CIDNET_LIB selects another build of the same ABI (e.g. libcidnet_b200_bf16.so, `build.py --bf16`)
LIB_PATH = os.environ.get("CIDNET_LIB") or os.path.join(_HERE, "libcidnet_b200.so")

OK, ERR_INVALID, ERR_ARCH, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4


class CidnetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcidnet_b200 error {code}: {msg}")
        self.code = code


class Shard(C.Structure):
    """cidnet_shard (include/cidnet_b200.h)"""
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("H_global", C.c_int32),
                ("row_begin", C.c_int32), ("row_end", C.c_int32), ("halo", C.c_int32)]


class HaloReq(C.Structure):
    """cidnet_halo_req"""
    _fields_ = [("base", C.c_void_p), ("row_bytes", C.c_int64), ("rows", C.c_int32),
                ("halo_top", C.c_int32), ("halo_bot", C.c_int32), ("reserved", C.c_int32)]


HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(HaloReq), C.c_int)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)

_lib = None

# name -> (restype, argtypes); must list every symbol include/cidnet_b200.h declares
SIGNATURES = {
    "cidnet_last_error": (C.c_char_p, []),
    "cidnet_abi_version": (C.c_int, []),
    "cidnet_act_dtype": (C.c_int, []),
    "cidnet_hvit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "cidnet_phvit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                               C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "cidnet_hvi_backward_scratch_bytes": (C.c_int64, []),
    "cidnet_hvit_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_void_p, C.c_void_p]),
    "cidnet_phvit_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                        C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "cidnet_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "cidnet_destroy": (C.c_int, [C.c_void_p]),
    "cidnet_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "cidnet_finalize_weights": (C.c_int, [C.c_void_p]),
    "cidnet_set_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "cidnet_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "cidnet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "cidnet_forward_launches": (C.c_int, [C.c_void_p]),
    "cidnet_shard_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Shard)]),
    "cidnet_shard_local_rows": (C.c_int, [C.POINTER(Shard)]),
    "cidnet_forward_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Shard), C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float,
                                         HALO_FN, ALLREDUCE_FN, C.c_void_p, C.c_void_p]),
    "cidnet_forward_sharded_dry": (C.c_int, [C.c_int, C.POINTER(Shard), C.c_void_p, C.c_int64, HALO_FN, ALLREDUCE_FN,
                                             C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cidnet_forward_sharded_dry_variant": (C.c_int, [C.c_int, C.c_int, C.POINTER(Shard), C.c_void_p, C.c_int64, HALO_FN,
                                                     ALLREDUCE_FN, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cidnet_peer_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cidnet_peer_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "cidnet_peer_open": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cidnet_peer_close": (C.c_int, [C.c_void_p]),
    "cidnet_peer_free": (C.c_int, [C.c_void_p]),
    "cidnet_peer_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "cidnet_forward_sharded_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Shard), C.POINTER(C.c_void_p),
                                              C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "cidnet_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "cidnet_pre_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "cidnet_post_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cidnet_read_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.c_void_p]),
    "cidnet_set_graphs": (C.c_int, [C.c_void_p, C.c_int]),
    "cidnet_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "cidnet_profile_count": (C.c_int, [C.c_void_p]),
    "cidnet_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_float),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cidnet_profile_get_span": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    # unit-test hook (csrc/test_hooks.cu): x, w_host, aux, ln_host, out, B, Cin, H, W, Cout, ksize, mode, flat, prelu, stream
    "cidnet_test_conv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_void_p]),
    # unit-test hook (csrc/api.cu): ctx, n, x_i, x_hv, after_cab_i, after_cab_hv, out_i, out_hv, B, H, W, stat_y0, stat_y1, stream
    "cidnet_test_lca_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}


def lib():
    """Load the shared library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python hvi-cidnet_b200/build.py` "
                "(there is no CPU/PyTorch fallback for the CIDNet hot path)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(code):
    if code != OK:
        raise CidnetError(code, lib().cidnet_last_error().decode("utf-8", "replace"))


def require_cuda_f32(t, name):
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200-native CIDNet path has no CPU fallback; move it to an sm_100 CUDA device")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
