"""Same-box A/B of the attention kernel's feature sets (STA_ATTN_FEAT, csrc/attention.cu::AttnFeat): correctness against
torch SDPA on identical bf16 operands and stand-alone throughput at the cfg-2 shapes (run under gpurun).

    python tools/attn_ab.py [feat ...]        # default: every compiled instance
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
FEATS = [16, 144, 192]
REF_LIB = os.path.join(ROOT, "tools", "ab", "libsta_r01attn.so")  # optional: a build with the round-1 attention kernel


def child():
    import torch
    import torch.nn.functional as F
    from vista_slam_b200._lib import check, cur_stream, lib, ptr
    L = lib()
    dev = "cuda"
    torch.manual_seed(0)
    worst = 0.0
    for (batch, heads, n, shift) in [(2, 3, 196, 0), (4, 12, 769, 2), (3, 16, 768, 0), (2, 2, 901, 1), (2, 3, 130, 0), (2, 2, 1, 0)]:
        C = heads * 64
        qkv = torch.randn(batch, n, 3 * C, device=dev).bfloat16()
        out = torch.zeros(batch, n, C, device=dev, dtype=torch.bfloat16)
        check(L.sta_op_attention(ptr(qkv), 3 * C, 0, ptr(qkv), 3 * C, C, ptr(qkv), 3 * C, 2 * C, ptr(out), C, batch, heads, n, n,
                                 shift, 0.125, 0, cur_stream()), "attention")
        torch.cuda.synchronize()
        q, k, v = [qkv[..., i * C:(i + 1) * C].float().view(batch, n, heads, 64).transpose(1, 2) for i in range(3)]
        if shift:
            idx = [(b + shift) % batch for b in range(batch)]
            k, v = k[idx], v[idx]
        ref = F.scaled_dot_product_attention(q, k, v, scale=0.125).transpose(1, 2).reshape(batch, n, C)
        e = float((out.float() - ref).abs().max() / ref.abs().max())
        worst = max(worst, e)
    # growing maxima (lazy rescale path, incl. the row-sum accumulator)
    batch, heads, n = 2, 3, 600
    C = heads * 64
    qkv = torch.randn(batch, n, 3 * C, device=dev)
    qkv[:, :, C:2 * C] *= (1.0 + 5.0 * (torch.arange(n, device=dev) // 128).float())[None, :, None]
    qkv = qkv.bfloat16()
    out = torch.zeros(batch, n, C, device=dev, dtype=torch.bfloat16)
    check(L.sta_op_attention(ptr(qkv), 3 * C, 0, ptr(qkv), 3 * C, C, ptr(qkv), 3 * C, 2 * C, ptr(out), C, batch, heads, n, n, 0,
                             0.125, 0, cur_stream()), "attention")
    q, k, v = [qkv[..., i * C:(i + 1) * C].float().view(batch, n, heads, 64).transpose(1, 2) for i in range(3)]
    ref = F.scaled_dot_product_attention(q, k, v, scale=0.125).transpose(1, 2).reshape(batch, n, C)
    e_resc = float((out.float() - ref).abs().max() / ref.abs().max())

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    res = []
    for (batch, heads, n, shift) in [(32, 16, 768, 0), (32, 12, 769, 0), (32, 12, 769, 16)]:
        C = heads * 64
        qkv = torch.randn(batch, n, 3 * C, device=dev).bfloat16()
        out = torch.zeros(batch, n, C, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: check(L.sta_op_attention(ptr(qkv), 3 * C, 0, ptr(qkv), 3 * C, C, ptr(qkv), 3 * C, 2 * C, ptr(out), C,
                                                     batch, heads, n, n, shift, 0.125, 0, cur_stream())))
        res.append("n%d%s %.3f ms %4.0f TF/s" % (n, "x" if shift else " ", ms, 4.0 * batch * heads * n * n * 64 / ms / 1e9))
    print("feat %2s  maxerr %.2e  rescale-case %.2e | %s" % (os.environ.get("STA_ATTN_FEAT"), worst, e_resc, " | ".join(res)), flush=True)


if __name__ == "__main__":
    if os.environ.get("ATTN_AB_CHILD"):
        child()
    else:
        feats = [int(a) for a in sys.argv[1:]] or FEATS
        for rep in range(2):
            if os.path.exists(REF_LIB):
                env = dict(os.environ, STA_B200_LIB=REF_LIB, ATTN_AB_CHILD="1", STA_ATTN_FEAT="r01")
                r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=300)
                print(r.stdout.strip() or ("r01 lib FAILED: %s" % r.stderr[-400:]), flush=True)
            for f in feats:
                for pp in ("1",):
                    env = dict(os.environ, STA_ATTN_FEAT=str(f), ATTN_AB_CHILD="1", STA_ATTN_PINGPONG=pp)
                    r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True,
                                       timeout=300)
                    print(("pingpong %s " % pp) + (r.stdout.strip() or ("feat %d FAILED: %s" % (f, r.stderr[-400:]))), flush=True)
