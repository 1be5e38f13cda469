"""hvi-cidnet_b200: B200-native (sm_100a) implementation of the HVI-CIDNet
inference forward path behind the reference's own module surface.

    from hvi_cidnet_b200.net.CIDNet import CIDNet      # drop-in for net.CIDNet.CIDNet

The arithmetic lives in libcidnet_b200.so (hand-written CUDA, C ABI declared in
include/cidnet_b200.h); this package is the thin Python host side.  There is no
CPU or PyTorch fallback: importing the ops without the built library, or
calling them without an sm_100 device, raises.
"""
__version__ = "0.1.0"
