"""One cfg-2 forward (after N warm-ups) for ncu captures:  python tools/one_forward.py [warmups] [pairs] [H] [W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vista_slam_b200.sta_model.sta_model import SymmetricTwoViewAssociation as STA  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
P = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H = int(sys.argv[3]) if len(sys.argv) > 3 else 384
W = int(sys.argv[4]) if len(sys.argv) > 4 else 512
m = STA().eval()
g = torch.Generator().manual_seed(0)
a = (torch.rand(P, 3, H, W, generator=g) * 2 - 1).bfloat16().cuda()
b = (torch.rand(P, 3, H, W, generator=g) * 2 - 1).bfloat16().cuda()
for _ in range(warm):
    m.forward_pairs(a, b)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed_forward")
m.forward_pairs(a, b)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("launches", m.launch_count)
