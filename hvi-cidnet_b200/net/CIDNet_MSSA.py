"""Host-side mirror of the fork's MSSA variant, `net.CIDNet_MSSA.CIDNet`
(/root/reference/net/CIDNet_MSSA.py:28-166) -- the class the fork's `eval.py` / `train.py` import.

Differences from `net.CIDNet.CIDNet`, all executed by libcidnet_b200.so (CIDNET_VARIANT_MSSA):
  * six `SpatialAttention` gates (:10-25) -- channel mean and max -> 7x7 conv -> sigmoid -> product --
    after the up blocks (`sa_hv3, sa_i3, sa_hv2, sa_i2, sa_hv1, sa_i1`, :132-153); six more
    state_dict tensors `sa_*.conv1.weight [1,2,7,7]` (197 in total);
  * `I_LCA5` is live and `ID_block2` consumes its output (:139-146).
Same surface otherwise: `forward(x)`, `HVIT(x)`, `trans`, strict `load_state_dict`.
"""
import torch.nn as nn

from .CIDNet import CIDNet as _BaseCIDNet


class SpatialAttention(nn.Module):
    """parameter container of CIDNet_MSSA.SpatialAttention (:11-18); never called"""
    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size == 7, "the B200 kernel is specialised for the 7x7 gate every call site uses"
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=3, bias=False)
        self.sigmoid = nn.Sigmoid()


class CIDNet(_BaseCIDNet):
    _variant = 1          # CIDNET_VARIANT_MSSA

    def __init__(self, channels=[36, 36, 72, 144], heads=[1, 2, 4, 8], norm=False):
        super().__init__(channels, heads, norm)
        for name in ("sa_hv3", "sa_i3", "sa_hv2", "sa_i2", "sa_hv1", "sa_i1"):     # order of :93-98
            setattr(self, name, SpatialAttention())
