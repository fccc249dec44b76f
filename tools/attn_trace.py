"""clock64 trace of CTA 0 of the one-query-tile attention kernels (run under gpurun).  The stamps are compiled in only with
-DSTA_ATTN_TRACE_BUILD (ATTN_STAMP in csrc/attention.cu); build that variant as tools/ab/libsta_attn_trace.so:

    nvcc <flags of vista_slam_b200/build.py> -DSTA_ATTN_TRACE_BUILD -c csrc/attention.cu -o /tmp/attention_trace.o
    nvcc -shared -o tools/ab/libsta_attn_trace.so /tmp/attention_trace.o csrc/build/{gemm,host_util,...}.o -cudart static

    STA_ATTN_FEAT=192 python tools/attn_trace.py [n] [heads]      # 192: attention_1q_kernel (448: the streamed experiment,
                                                                  # tools/experiments/attention_1qs_kernel.cu.txt)
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.environ.setdefault("STA_B200_LIB", os.path.join(ROOT, "tools", "ab", "libsta_attn_trace.so"))
import torch  # noqa: E402

buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
os.environ["STA_ATTN_TRACE"] = str(buf.data_ptr())
from vista_slam_b200._lib import check, cur_stream, lib, ptr  # noqa: E402

L = lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = 32
C = heads * 64
qkv = torch.randn(batch, n, 3 * C, device="cuda").bfloat16()
out = torch.zeros(batch, n, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    check(L.sta_op_attention(ptr(qkv), 3 * C, 0, ptr(qkv), 3 * C, C, ptr(qkv), 3 * C, 2 * C, ptr(out), C, batch, heads, n, n, 0,
                             0.125, 0, cur_stream()))
torch.cuda.synchronize()
b = buf.cpu().tolist()
if not any(x > 0 for x in b):
    sys.exit("this build of libsta_b200.so has no attention trace stamps (see the docstring)")
t0 = min(x for x in b if x > 0)
f = lambda e: " ".join("%6s" % (x - t0 if x > 0 else "-") for x in e)
feat = os.environ.get("STA_ATTN_FEAT")
print("feat", feat, "n", n, "heads", heads, "(clk since the first stamp; key-tile step g of CTA 0)")
print("MMA thread   g: loop_top  s_free_ok  S_issued  p_full_ok  v_full_ok  PV_issued")
for g in range(24):
    print("  %2d: %s" % (g, f(b[8 * g: 8 * g + 6])))
if feat == "448":
    print("softmax row 0  g: tile_top  s_full(g+1)_ok  before_o_full  o_full_ok  loop_end  before_st_wait  p_full_arrived")
else:
    print("softmax row 0  g: before_s_full  s_full_ok  max_done  o_full_ok  -  before_st_wait  p_full_arrived")
for g in range(24):
    print("  %2d: %s" % (g, f(b[256 + 8 * g: 256 + 8 * g + 7])))
