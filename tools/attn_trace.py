"""clock64 trace of one CTA of the attention kernel (run under gpurun): STA_ATTN_TRACE hands the kernel a device buffer.
    STA_ATTN_FEAT=15 python tools/attn_trace.py [n] [heads]

The stamps exist only in the experimental builds of csrc/attention.cu (git history of round 2: the feature-template
versions); the production kernel carries no trace code.  The traces that drove the round-2 decisions are kept in
profiles/r02_attn_trace_*.log.
"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
os.environ["STA_ATTN_TRACE"] = str(buf.data_ptr())
from vista_slam_b200._lib import check, cur_stream, lib, ptr
L = lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = 32
C = heads * 64
qkv = torch.randn(batch, n, 3 * C, device="cuda").bfloat16()
out = torch.zeros(batch, n, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    check(L.sta_op_attention(ptr(qkv), 3 * C, 0, ptr(qkv), 3 * C, C, ptr(qkv), 3 * C, 2 * C, ptr(out), C, batch, heads, n, n, 0, 0.125, 0, cur_stream()))
torch.cuda.synchronize()
b = buf.cpu().tolist()
if not any(x > 0 for x in b):
    sys.exit("this build of libsta_b200.so has no attention trace stamps (see the docstring)")
t0 = min(x for x in b if x > 0)
f = lambda e: [x - t0 if x > 0 else None for x in e]
print("feat", os.environ.get("STA_ATTN_FEAT"), "n", n, "heads", heads)
print("MMA thread step n: [A: before_sfree, after_sfree, after_issue_S] [B: ...] [A: before_pfull, after_pfull, after_issue_PV] [B: ...]")
for n_ in range(2, 16):
    e = b[16 * n_: 16 * n_ + 16]
    print(n_, f(e[0:3]), f(e[4:7]), f(e[8:11]), f(e[12:15]))
for grp in range(2):
    print("softmax group", grp, ": n: before_sfull, after_sfull, after_load+sfree, after_max, after_ofull/rescale, after_exp+arrive")
    for u in range(2, 16):
        print(u, f(b[512 + grp * 256 + 8 * u: 512 + grp * 256 + 8 * u + 6]))
for grp in range(2):
    print("epilogue group", grp, ": item: before_ofull, after_ofull, after_store")
    for it in range(4):
        print(it, f(b[1100 + grp * 32 + it * 4: 1100 + grp * 32 + it * 4 + 3]))
