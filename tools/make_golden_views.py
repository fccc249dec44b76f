"""Golden vectors for the general forward(views) loop of the UNMODIFIED reference (this container only):
several support views (neighbour + loop), an all-portrait batch and a mixed landscape/portrait batch -- the
`transpose_to_landscape` branches of vista_slam/utils/misc.py:36-82 that a single landscape pair never takes.

    python tools/make_golden_views.py

Writes tests/golden/views_*.npz (outputs only; the inputs are regenerated from the seeds in the metadata).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle.sta_oracle import StaOracle, make_images, make_state_dict, usable_cpus  # noqa: E402
from ref_import import import_reference_sta  # noqa: E402

# name, B, H, W (tensor shape), per-sample true shapes, number of neighbour views, number of loop views, image seed
CASES = [
    ("views_portrait_80x48_s3", 1, 80, 48, [[80, 48]], 2, 1, 21),       # all-portrait, 3 support views (5x3 token grid)
    ("views_mixed_b2_64x80_s2", 2, 64, 80, [[64, 80], [80, 64]], 1, 1, 22),  # sample 1 declares a portrait true_shape
]
KEYS = ("pts3d_pred", "conf", "relative_pose", "relative_pose_conf")


def case_inputs(B, H, W, n_support, seed):
    imgs = [make_images(B, H, W, seed + k)[0] for k in range(n_support + 1)]
    return imgs[0], imgs[1:]


def maxrel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    torch.set_num_threads(usable_cpus())
    STA = import_reference_sta()
    sd = make_state_dict(0)
    ref = STA()
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    orc = StaOracle(sd, emulate_bf16=False)
    for name, B, H, W, shapes, n_nb, n_loop, seed in CASES:
        ts = torch.tensor(shapes)
        main_img, sup = case_inputs(B, H, W, n_nb + n_loop, seed)
        views = {"main_view": {"img": main_img, "true_shape": ts},
                 "neighbor_views": [{"img": im, "true_shape": ts} for im in sup[:n_nb]],
                 "loop_views": [{"img": im, "true_shape": ts} for im in sup[n_nb:]]}
        with torch.no_grad():
            out = ref(views)
            o_main, o_sup = orc.forward_views(main_img, ts, [(im, ts) for im in sup])
        arrays, dev = {}, {}
        for side, res, ores in (("main", out["main_views"], o_main), ("support", out["support_views"], o_sup)):
            assert len(res) == n_nb + n_loop
            for i, (r, o) in enumerate(zip(res, ores)):
                for k in KEYS:
                    arrays["%s%d_%s" % (side, i, k)] = r[k].detach().cpu().numpy()
                    dev["%s%d.%s" % (side, i, k)] = maxrel(o[k], r[k])
        meta = {"case": name, "B": B, "H": H, "W": W, "true_shape": shapes, "neighbors": n_nb, "loops": n_loop,
                "image_seed": seed, "weight_seed": 0, "torch": torch.__version__, "reference_commit": "b13ac44",
                "precision": "fp32 CPU", "oracle_vs_reference_maxrel_worst": max(dev.values())}
        print(name, "oracle vs reference worst %.2e" % max(dev.values()), {k: tuple(v.shape) for k, v in list(arrays.items())[:4]})
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=json.dumps(meta), **arrays)


if __name__ == "__main__":
    main()
