"""Import the UNMODIFIED reference STA model from /root/reference (this container only).

xformers is not installed here, so `xformers.ops.memory_efficient_attention` is shimmed in
sys.modules with torch SDPA using the same (B, N, H, K) layout as the call site
vista_slam/sta_model/blocks/sta_blocks.py:139-143.  Used only by the golden-vector
generators under tools/ -- never at test/bench run time (the GPU box has no /root/reference).
"""
import sys
import types

import torch
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def _install_xformers_shim():
    if "xformers" in sys.modules:
        return
    xf = types.ModuleType("xformers")
    ops = types.ModuleType("xformers.ops")

    def memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
        assert attn_bias is None and p == 0.0
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=scale)
        return o.transpose(1, 2)

    ops.memory_efficient_attention = memory_efficient_attention
    xf.ops = ops
    sys.modules["xformers"] = xf
    sys.modules["xformers.ops"] = ops


def import_reference_sta():
    _install_xformers_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from vista_slam.sta_model.sta_model import SymmetricTwoViewAssociation  # noqa
    return SymmetricTwoViewAssociation
