"""GPU bring-up of the whole model against the oracle and the golden fixtures (run under gpurun).

    python tools/model_check.py [--skip-big]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

from oracle.sta_oracle import usable_cpus, StaOracle, flops_per_pair, make_images, make_state_dict  # noqa: E402
from vista_slam_b200 import _lib  # noqa: E402
from vista_slam_b200.sta_model.sta_model import SymmetricTwoViewAssociation as STA  # noqa: E402


def nrm(a, b):
    """max |a-b| / max |b| and median elementwise relative error."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    d = (a - b).abs()
    return "maxn=%.2e medrel=%.2e" % (float(d.max() / b.abs().max().clamp_min(1e-30)),
                                      float((d / b.abs().clamp_min(1e-6)).median()))


def main():
    torch.set_num_threads(usable_cpus())
    t0 = time.time()
    sd = make_state_dict(0)
    print("weights %.1fs, cpu threads %d" % (time.time() - t0, torch.get_num_threads()), flush=True)
    model = STA()
    model.load_state_dict(sd, strict=True)
    model.eval()
    dev = torch.device("cuda")
    t0 = time.time()
    model._ready(torch.empty(1, device=dev))
    torch.cuda.synchronize()
    print("upload %.1fs, device bytes %.1f MB" % (time.time() - t0, model.device_bytes / 1e6), flush=True)
    orc32 = StaOracle(sd, emulate_bf16=False)
    orcbf = StaOracle(sd, emulate_bf16=True)

    for case in ("pair_64x80", "pair_b2_48x64"):
        g = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
        meta = json.loads(str(g["meta"]))
        B, H, W = meta["B"], meta["H"], meta["W"]
        img1, img2 = make_images(B, H, W, meta["image_seed"])
        m, s = model.forward_pairs(img1.to(dev), img2.to(dev))
        torch.cuda.synchronize()
        with torch.no_grad():
            om, os_ = orcbf.forward_pair(img1, img2)
        print("== %s (fused forward_pairs) ==" % case)
        for k, gk in (("pts3d_pred", "pts3d"), ("conf", "conf"), ("relative_pose", "pose"), ("relative_pose_conf", "pose_conf")):
            print("  main.%-18s vs reference-golden: %s | vs oracle-bf16emu: %s" %
                  (k, nrm(m[k], torch.from_numpy(g["main_" + gk])), nrm(m[k], om[k])))
            print("  supp.%-18s vs reference-golden: %s | vs oracle-bf16emu: %s" %
                  (k, nrm(s[k], torch.from_numpy(g["support_" + gk])), nrm(s[k], os_[k])))
        # sub-entry points
        ts = torch.tensor([[H, W]] * B)
        f1, p1 = model._encode_image(img1.to(dev), ts, normalize=False)
        f2, p2 = model._encode_image(img2.to(dev), ts, normalize=False)
        d1, d2 = model._decode_stereo(f1, f2, p1, p2)
        print("  enc_feat1 vs golden: %s ; pos equal: %s" % (nrm(f1, torch.from_numpy(g["enc_feat1"])),
                                                            bool((p1.cpu() == torch.from_numpy(g["pos1"])).all())))
        for k, t in (("dec1_6", d1[6]), ("dec1_9", d1[9]), ("dec1_12", d1[12]), ("dec2_12", d2[12])):
            print("  %s vs golden: %s" % (k, nrm(t, torch.from_numpy(g[k]))))
        pts = model.head_pts([f1] + [t[:, 1:, :] for t in d1], ts)
        pose = model.head_pose_s(d1[-1][:, 0, :])
        print("  sub-entry pts3d vs fused: %s ; pose vs fused: %s" % (nrm(pts["pts3d"], m["pts3d_pred"]),
                                                                     nrm(pose["pose"], m["relative_pose"])))
        R = m["relative_pose"][:, :3, :3].double().cpu()
        print("  R orthonormality %.2e det %s" % (float((R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max()),
                                                  torch.det(R).tolist()), flush=True)

    if "--skip-big" in sys.argv:
        return
    # cfg-1: single 224x224 pair
    img1, img2 = make_images(1, 224, 224, 1234)
    m, s = model.forward_pairs(img1.to(dev), img2.to(dev))
    torch.cuda.synchronize()
    t0 = time.time()
    with torch.no_grad():
        o32m, o32s = orc32.forward_pair(img1, img2)
    t32 = time.time() - t0
    with torch.no_grad():
        obm, obs = orcbf.forward_pair(img1, img2)
    print("== 224x224 pair (oracle fp32 CPU %.2fs) ==" % t32)
    for k in ("pts3d_pred", "conf", "relative_pose", "relative_pose_conf"):
        print("  main.%-18s vs oracle-fp32: %s | vs oracle-bf16emu: %s | emu vs fp32: %s" %
              (k, nrm(m[k], o32m[k]), nrm(m[k], obm[k]), nrm(obm[k], o32m[k])))
        print("  supp.%-18s vs oracle-fp32: %s | vs oracle-bf16emu: %s" % (k, nrm(s[k], o32s[k]), nrm(s[k], obs[k])))
    print("  value ranges: pts3d |max| %.3f conf [%.3f, %.3f] pose_conf %.3f" %
          (float(m["pts3d_pred"].abs().max()), float(m["conf"].min()), float(m["conf"].max()),
           float(m["relative_pose_conf"][0])), flush=True)

    # cfg-2 timing: 16 pairs 384x512 bf16
    B, H, W = 16, 384, 512
    g = torch.Generator().manual_seed(1)
    i1 = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).bfloat16().to(dev)
    i2 = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).bfloat16().to(dev)
    for _ in range(2):
        model.forward_pairs(i1, i2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    l0 = model.launch_count
    e0.record()
    for _ in range(n):
        out = model.forward_pairs(i1, i2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = flops_per_pair(H, W)
    print("== cfg-2: %d pairs %dx%d bf16: %.2f ms/step, %.1f pairs/s, %.0f TFLOP/s algorithmic, %d launches/step, ws %.1f GB ==" %
          (B, W, H, ms, B / ms * 1e3, B * fl / ms / 1e9, (model.launch_count - l0) // n, model.device_bytes / 1e9))
    print("  finite:", bool(torch.isfinite(out[0]["pts3d_pred"]).all()), bool(torch.isfinite(out[1]["relative_pose"]).all()))
    L = _lib.lib()
    L.sta_profile(model._handle, 1)
    model.forward_pairs(i1, i2)
    ms4 = (ctypes.c_double * 4)()
    cnt4 = (ctypes.c_int64 * 4)()
    fl4 = (ctypes.c_double * 4)()
    _lib.check(L.sta_profile_read(model._handle, ms4, cnt4, fl4))
    L.sta_profile(model._handle, 0)
    for i, nm in enumerate(("gemm-linear", "gemm-conv3x3", "attention", "layernorm")):
        print("  family %-12s %4d launches %8.3f ms %8.1f TFLOP/s" % (nm, cnt4[i], ms4[i], fl4[i] / max(ms4[i], 1e-9) / 1e9))
    print("  profiled sum %.2f ms of %.2f ms step" % (sum(ms4), ms), flush=True)


if __name__ == "__main__":
    main()
