"""Same-box A/B of the two A-operand staging schemes of the implicit-GEMM 3x3 convolution (csrc/gemm.cuh: A_CONV3 = one TMA
box per filter tap, A_CONV3H = one halo tile per channel chunk feeding all nine taps; env STA_CONV_HALO=0/1): correctness
against torch conv2d on identical bf16 operands (ragged sizes) and stand-alone throughput at the DPT-head shapes of cfg-2
(16 views per head half).  Run under gpurun:

    python tools/conv_ab.py
"""
import math
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def child():
    import torch
    import torch.nn.functional as F
    from bringup import gemm_desc, run_gemm
    from vista_slam_b200._lib import EPI_BF16, EPI_HEAD
    dev = "cuda"
    torch.manual_seed(0)
    worst = 0.0
    for (nimg, H, W, Cin, Cout) in [(2, 8, 16, 64, 256), (3, 14, 14, 128, 256), (1, 20, 36, 256, 128), (2, 33, 23, 192, 256),
                                    (1, 48, 64, 256, 256)]:
        x = torch.randn(nimg, H, W, Cin, device=dev).bfloat16()
        wp = (torch.randn(Cout, 9 * Cin, device=dev) / math.sqrt(9 * Cin)).bfloat16()
        b = torch.randn(Cout, device=dev)
        out = torch.zeros(nimg, H, W, Cout, device=dev, dtype=torch.bfloat16)
        out2 = torch.zeros_like(out)
        r1 = torch.randn(nimg, H, W, Cout, device=dev).bfloat16()
        run_gemm(gemm_desc(conv3x3=1, epi=EPI_BF16, A=x, W=wp, ldw=9 * Cin, N=Cout, K=9 * Cin, nimg=nimg, H=H, Wd=W, Cin=Cin,
                           bias=b, out=out, ldo=Cout, out2=out2, resid=r1))
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2), b,
                       padding=1).permute(0, 2, 3, 1) + r1.float()
        worst = max(worst, float((out.float() - ref).abs().max() / ref.abs().max()),
                    float((out2.float() - ref.relu()).abs().max() / ref.abs().max()))
    # which filter taps are wrong (diagnostic for the descriptor arithmetic of the halo path)
    taps = []
    nimg, H, W, Cin, Cout = 1, 32, 16, 64, 128
    x = torch.randn(nimg, H, W, Cin, device=dev).bfloat16()
    for t in range(9):
        w4 = torch.zeros(Cout, 9, Cin, device=dev)
        w4[:, t] = torch.randn(Cout, Cin, device=dev) / 8
        wp = w4.reshape(Cout, 9 * Cin).bfloat16()
        out = torch.zeros(nimg, H, W, Cout, device=dev, dtype=torch.bfloat16)
        run_gemm(gemm_desc(conv3x3=1, epi=EPI_BF16, A=x, W=wp, ldw=9 * Cin, N=Cout, K=9 * Cin, nimg=nimg, H=H, Wd=W, Cin=Cin,
                           out=out, ldo=Cout))
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2),
                       padding=1).permute(0, 2, 3, 1)
        taps.append("%.0e" % float((out.float() - ref).abs().max() / ref.abs().max()))
    print("halo %s per-tap error: %s" % (os.environ.get("STA_CONV_HALO", "1"), " ".join(taps)), flush=True)

    def timeit(fn, n=10):
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
    for (name, nimg, H, W, Cin, Cout, epi, skip) in [("refine 96x128 256->256", 16, 96, 128, 256, 256, EPI_BF16, 0),
                                                      ("same + skip + relu copy", 16, 96, 128, 256, 256, EPI_BF16, 1),
                                                      ("head.0 192x256 256->128", 16, 192, 256, 256, 128, EPI_BF16, 0),
                                                      ("head.2 384x512 128->128 (+1x1)", 16, 384, 512, 128, 128, EPI_HEAD, 0)]:
        x = torch.randn(nimg, H, W, Cin, device=dev).bfloat16()
        wp = (torch.randn(Cout, 9 * Cin, device=dev) / math.sqrt(9 * Cin)).bfloat16()
        b = torch.randn(Cout, device=dev) * 0.1
        kw = dict(conv3x3=1, epi=epi, A=x, W=wp, ldw=9 * Cin, N=Cout, K=9 * Cin, nimg=nimg, H=H, Wd=W, Cin=Cin, bias=b)
        keep = []
        if epi == EPI_HEAD:
            w4t = (torch.randn(128, 4, device=dev) / math.sqrt(128)).contiguous()
            b4 = torch.randn(4, device=dev) * 0.1
            pts = torch.zeros(nimg, H, W, 3, device=dev)
            conf = torch.zeros(nimg, H, W, device=dev)
            kw.update(head_w=w4t, head_b=b4, pts3d=pts, conf=conf)
            keep += [w4t, b4, pts, conf]
        else:
            out = torch.zeros(nimg, H, W, Cout, device=dev, dtype=torch.bfloat16)
            kw.update(out=out, ldo=Cout)
            keep.append(out)
            if skip:
                r1 = torch.randn(nimg, H, W, Cout, device=dev).bfloat16()
                out2 = torch.zeros_like(out)
                kw.update(resid=r1, out2=out2)
                keep += [r1, out2]
        d = gemm_desc(**kw)
        ms = timeit(lambda: run_gemm(d))
        res.append("%s %.3f ms %4.0f TF/s" % (name, ms, 2.0 * nimg * H * W * 9 * Cin * Cout / ms / 1e9))
    print("halo %s ew16 %s  maxerr %.2e | %s" % (os.environ.get("STA_CONV_HALO", "1"), os.environ.get("STA_CONV_EW16", "1"), worst,
                                                 " | ".join(res)), flush=True)


if __name__ == "__main__":
    if os.environ.get("CONV_AB_CHILD"):
        child()
    else:
        for rep in range(2):
            for h, ew in (("0", "0"), ("1", "0"), ("1", "1")):
                env = dict(os.environ, STA_CONV_HALO=h, STA_CONV_EW16=ew, CONV_AB_CHILD="1")
                r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=300)
                print(r.stdout.strip() or ("halo %s FAILED: %s" % (h, r.stderr[-600:])), flush=True)
