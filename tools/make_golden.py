"""Generate tests/golden/*.npz by running the UNMODIFIED reference STA model (this container only).

    python tools/make_golden.py

For each golden case the deterministic synthetic checkpoint oracle.make_state_dict(seed) is loaded
into the reference module with strict=True, the reference forward() (sta_model.py:247-291) and its
sub-entry points are run in fp32 on CPU, and the outputs are stored (float32, compressed).  The
oracle restatement is run on the same inputs and its deviation from the reference is printed and
stored in the fixture's metadata.  The GPU box never runs this script (no /root/reference there).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle.sta_oracle import usable_cpus, StaOracle, make_images, make_state_dict  # noqa: E402
from ref_import import import_reference_sta  # noqa: E402

CASES = [
    # name, B, H, W, weight seed, image seed
    ("pair_64x80", 1, 64, 80, 0, 1234),     # odd token-grid width (5): exercises the refinenet4 crop
    ("pair_b2_48x64", 2, 48, 64, 0, 77),    # batch 2, 3x4 token grid
    ("pair_224x224", 1, 224, 224, 0, 1234),  # cfg-1: the reference's native size (sta_model.py:34)
    ("pair_384x512", 1, 384, 512, 0, 1234),  # one cfg-2 pair (H = 384, W = 512); bf16-representable images
]
# cases whose images are rounded to bf16 first (cfg-2 feeds bf16 images; the reference consumes them upcast to fp32)
BF16_IMAGES = {"pair_384x512"}
# the large case stores only what the GPU tests compare (outputs + one feature per stage) to keep the fixture small
SLIM = {"pair_384x512"}


def maxrel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    only = set(sys.argv[1:])  # optional: regenerate only the named cases
    torch.set_num_threads(usable_cpus())
    STA = import_reference_sta()
    t0 = time.time()
    sd = make_state_dict(0)
    print("state dict: %d keys, %.1fs" % (len(sd), time.time() - t0))
    ref = STA()
    missing = ref.load_state_dict(sd, strict=True)
    print("reference load_state_dict(strict=True):", missing)
    ref.eval()
    orc = StaOracle(sd, emulate_bf16=False)
    outdir = os.path.join(ROOT, "tests", "golden")
    for name, B, H, W, wseed, iseed in CASES:
        assert wseed == 0
        if only and name not in only:
            continue
        img1, img2 = make_images(B, H, W, iseed)
        if name in BF16_IMAGES:
            img1, img2 = img1.bfloat16().float(), img2.bfloat16().float()
        ts = torch.tensor([[H, W]] * B)
        views = {"main_view": {"img": img1, "true_shape": ts},
                 "neighbor_views": [{"img": img2, "true_shape": ts}], "loop_views": []}
        with torch.no_grad():
            out = ref(views)
            f1, pos1 = ref._encode_image(img1, ts, normalize=False)
            f2, pos2 = ref._encode_image(img2, ts, normalize=False)
            d1, d2 = ref._decode_stereo(f1, f2, pos1, pos2)
            o_main, o_sup = orc.forward_pair(img1, img2)
            of1, _ = orc.encode_image(img1)
        mv, sv = out["main_views"][0], out["support_views"][0]
        dev = {}
        for k in ("pts3d_pred", "conf", "relative_pose", "relative_pose_conf"):
            dev["main." + k] = maxrel(o_main[k], mv[k])
            dev["support." + k] = maxrel(o_sup[k], sv[k])
        dev["enc_feat"] = maxrel(of1, f1)
        print(name, json.dumps(dev, indent=1))
        arrays = {
            "main_pts3d": mv["pts3d_pred"], "main_conf": mv["conf"], "main_pose": mv["relative_pose"],
            "main_pose_conf": mv["relative_pose_conf"],
            "support_pts3d": sv["pts3d_pred"], "support_conf": sv["conf"], "support_pose": sv["relative_pose"],
            "support_pose_conf": sv["relative_pose_conf"],
            "enc_feat1": f1, "dec1_6": d1[6], "dec1_9": d1[9], "dec1_12": d1[12], "dec2_12": d2[12],
            "pos1": pos1,
        }
        if name in SLIM:
            for k in ("dec1_6", "dec1_9", "dec2_12", "pos1"):
                arrays.pop(k)
        meta = {"case": name, "B": B, "H": H, "W": W, "weight_seed": wseed, "image_seed": iseed,
                "bf16_images": name in BF16_IMAGES, "torch": torch.__version__, "reference_commit": "b13ac44", "precision": "fp32 CPU",
                "oracle_vs_reference_maxrel": dev}
        np.savez_compressed(os.path.join(outdir, name + ".npz"), meta=json.dumps(meta),
                            **{k: v.detach().cpu().numpy() for k, v in arrays.items()})
        print("wrote", name, {k: tuple(v.shape) for k, v in arrays.items()})


if __name__ == "__main__":
    main()
