"""GPU: the general `forward(views)` loop (sta_model.py:247-291) -- several support views (neighbour + loop), an
all-portrait batch and a mixed landscape / portrait batch (transpose_to_landscape, utils/misc.py:36-82) -- against golden
vectors produced by the unmodified reference (tools/make_golden_views.py), in production precision (measured-bound
tolerances of tests/test_model_gpu.py) and in the split-precision parity mode (north_star's 1e-3 / 1e-4)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.sta_oracle import make_images
from test_model_gpu import TOL_FP32, TOL_X3, maxn

pytestmark = pytest.mark.gpu

CASES = ["views_portrait_80x48_s3", "views_mixed_b2_64x80_s2"]


def load_case(case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(g["meta"]))
    n_sup = meta["neighbors"] + meta["loops"]
    imgs = [make_images(meta["B"], meta["H"], meta["W"], meta["image_seed"] + k)[0] for k in range(n_sup + 1)]
    ts = torch.tensor(meta["true_shape"])
    views = {"main_view": {"img": imgs[0].cuda(), "true_shape": ts},
             "neighbor_views": [{"img": im.cuda(), "true_shape": ts} for im in imgs[1:1 + meta["neighbors"]]],
             "loop_views": [{"img": im.cuda(), "true_shape": ts} for im in imgs[1 + meta["neighbors"]:]]}
    return g, meta, views, n_sup


def check(out, g, n_sup, tol):
    assert len(out["main_views"]) == n_sup and len(out["support_views"]) == n_sup
    for side in ("main", "support"):
        for i in range(n_sup):
            res = out[side + "_views"][i]
            for k in tol:
                gold = torch.from_numpy(g["%s%d_%s" % (side, i, k)])
                assert tuple(res[k].shape) == tuple(gold.shape), (side, i, k, res[k].shape, gold.shape)
                e = maxn(res[k], gold)
                assert e < tol[k], (side, i, k, e)


@pytest.mark.parametrize("case", CASES)
def test_forward_views_vs_reference_golden(cuda_model, case):
    g, meta, views, n_sup = load_case(case)
    check(cuda_model(views), g, n_sup, TOL_FP32)


@pytest.mark.parametrize("case", CASES)
def test_forward_views_split_precision_meets_north_star(cuda_model_x3, case):
    g, meta, views, n_sup = load_case(case)
    check(cuda_model_x3(views), g, n_sup, TOL_X3)


def test_forward_pairs_on_a_portrait_batch_returns_the_landscape_maps(cuda_model):
    """forward_pairs (fused path) must agree with the reference semantics the general loop implements: an all-portrait
    batch comes back transposed to landscape (utils/misc.py:58-61)."""
    g, meta, views, n_sup = load_case("views_portrait_80x48_s3")
    main, sup = cuda_model.forward_pairs(views["main_view"]["img"], views["neighbor_views"][0]["img"])
    gold = torch.from_numpy(g["main0_pts3d_pred"])
    assert tuple(main["pts3d_pred"].shape) == tuple(gold.shape) == (1, 48, 80, 3)
    assert maxn(main["pts3d_pred"], gold) < TOL_FP32["pts3d_pred"]
    assert maxn(sup["conf"], torch.from_numpy(g["support0_conf"])) < TOL_FP32["conf"]
