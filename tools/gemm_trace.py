import math, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
os.environ["STA_GEMM_TRACE"] = str(buf.data_ptr())
import tools.bringup as bu
from vista_slam_b200._lib import EPI_BF16, EPI_F32, EPI_GELU, EPI_ROPE
M = 24576
which = sys.argv[1] if len(sys.argv) > 1 else "rope"
N, K, epi = {"rope": (3072, 1024, EPI_ROPE), "bf16": (3072, 1024, EPI_BF16), "f32": (1024, 1024, EPI_F32),
             "gelu": (4096, 1024, EPI_GELU)}[which]
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda")
kw = dict(epi=epi, A=A, lda=K, W=W, ldw=K, M=M, N=N, K=K, bias=bias, ldo=N)
if epi == EPI_F32:
    out = torch.zeros(M, N, device="cuda"); kw.update(out=out, resid=out)
elif epi == EPI_ROPE:
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    pos = torch.randint(0, 32, (M, 2), device="cuda", dtype=torch.int32); kw.update(out=out, pos=pos, rope_cols=2 * N // 3)
else:
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); kw.update(out=out)
d = bu.gemm_desc(**kw)
for _ in range(3):
    bu.run_gemm(d)
torch.cuda.synchronize()
b = buf.cpu().tolist()
t0 = min(x for x in b if x > 0)
print(which, "epilogue warp 4: tile: before_tfull_wait, after_wait, after_epilogue   |  MMA: before_tempty_wait, after_wait, after_issue")
for t in range(16):
    e = [x - t0 for x in b[4 * t:4 * t + 3]]
    m = [x - t0 for x in b[256 + 4 * t:256 + 4 * t + 3]]
    print(t, e, " | ", m)
print("TMA epilogue fine trace (warp 4): per tile: t(tfull) then per chunk [tmem+bias ready, math done, store issued] relative to tfull")
for t in range(2, 8):
    r = b[512 + 16 * t: 512 + 16 * t + 13]
    if r[0] <= 0:
        continue
    print(t, r[0] - t0, [x - r[0] for x in r[1:] if x > 0])
