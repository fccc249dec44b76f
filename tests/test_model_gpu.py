"""GPU: the whole STA path through the module surface -> C ABI -> sm_100a kernels, against the oracle
(same seeded inputs) and the golden vectors produced by the unmodified reference.

Tolerance protocol: bf16 MMA operands cannot meet rtol 1e-3 against an fp32 oracle --
the reference's own fp32 -> bf16-autocast deviation is 5e-3 (pts3d) / 3.5e-3 (pose), max-normalised.  So:
  (a) vs the oracle in bf16-operand emulation (same operand precision): tight bounds, stated per output;
  (b) vs the fp32 reference golden / fp32 oracle: the bf16 noise floor, stated per output.
Max-normalised error = max|a-b| / max|b|.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.sta_oracle import StaOracle, make_images

pytestmark = pytest.mark.gpu

# Max-normalised bounds, set at ~3x the deviations measured on B200 in round 1 (profiles/r01_model_check*.log:
# trunk features 6e-3..9e-3, pose 2e-3..1.8e-2, pts3d/conf 1e-2..2e-2 at |xyz| <= 3, pose_conf <= 2e-3).
# (a) vs oracle-bf16emu, (b) vs fp32 reference golden / fp32 oracle.
TOL_EMU = {"pts3d_pred": 5e-2, "conf": 5e-2, "relative_pose": 4e-2, "relative_pose_conf": 5e-3}
TOL_FP32 = {"pts3d_pred": 1e-1, "conf": 1e-1, "relative_pose": 5e-2, "relative_pose_conf": 1e-2}


def maxn(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("case", ["pair_64x80", "pair_b2_48x64"])
def test_forward_pairs_vs_oracle_and_reference_golden(cuda_model, state_dict, case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(g["meta"]))
    img1, img2 = make_images(meta["B"], meta["H"], meta["W"], meta["image_seed"])
    main, sup = cuda_model.forward_pairs(img1.cuda(), img2.cuda())
    with torch.no_grad():
        o_main, o_sup = StaOracle(state_dict, emulate_bf16=True).forward_pair(img1, img2)
    gk = {"pts3d_pred": "pts3d", "conf": "conf", "relative_pose": "pose", "relative_pose_conf": "pose_conf"}
    for res, orc, pre in ((main, o_main, "main_"), (sup, o_sup, "support_")):
        for k in TOL_EMU:
            assert maxn(res[k], orc[k]) < TOL_EMU[k], (pre, k, maxn(res[k], orc[k]))
            assert maxn(res[k], torch.from_numpy(g[pre + gk[k]])) < TOL_FP32[k], (pre, k)


def test_sub_entry_points_match_reference_golden_and_fused_path(cuda_model):
    """the slam.py call sequence (slam.py:144,162-180): _encode_image -> _decode_stereo -> head_pose_s / head_pts."""
    g = np.load(os.path.join(GOLDEN_DIR, "pair_64x80.npz"))
    meta = json.loads(str(g["meta"]))
    B, H, W = meta["B"], meta["H"], meta["W"]
    img1, img2 = make_images(B, H, W, meta["image_seed"])
    ts = torch.tensor([[H, W]] * B)
    f1, p1 = cuda_model._encode_image(img1.cuda(), ts, normalize=False)
    f2, p2 = cuda_model._encode_image(img2.cuda(), ts, normalize=False)
    assert f1.shape == (B, (H // 16) * (W // 16), 1024) and p1.dtype == torch.int64
    assert torch.equal(p1.cpu(), torch.from_numpy(g["pos1"]))
    assert maxn(f1, torch.from_numpy(g["enc_feat1"])) < 5e-2
    d1, d2 = cuda_model._decode_stereo(f1, f2, p1, p2)
    assert len(d1) == 13 and len(d2) == 13 and d1[0].shape == (B, (H // 16) * (W // 16) + 1, 768)
    for k, t in (("dec1_6", d1[6]), ("dec1_9", d1[9]), ("dec1_12", d1[12]), ("dec2_12", d2[12])):
        assert maxn(t, torch.from_numpy(g[k])) < 5e-2, k
    pts = cuda_model.head_pts([f1] + [t[:, 1:, :] for t in d1], ts)
    pose = cuda_model.head_pose_s(d1[-1][:, 0, :])
    main, _ = cuda_model.forward_pairs(img1.cuda(), img2.cuda())
    # same kernels, same order -> the decomposed path reproduces the fused one up to the fp32->bf16 hand-over
    assert maxn(pts["pts3d"], main["pts3d_pred"]) < 2e-2
    assert maxn(pose["pose"], main["relative_pose"]) < 5e-3
    out = cuda_model({"main_view": {"img": img1.cuda(), "true_shape": ts},
                      "neighbor_views": [{"img": img2.cuda(), "true_shape": ts}], "loop_views": []})
    assert set(out) == {"main_views", "support_views"}
    assert set(out["main_views"][0]) == {"pts3d_pred", "conf", "relative_pose", "relative_pose_conf"}
    assert out["main_views"][0]["pts3d_pred"].shape == (B, H, W, 3)


def test_cfg1_single_224_pair_vs_oracle(cuda_model, state_dict):
    img1, img2 = make_images(1, 224, 224, 1234)
    main, sup = cuda_model.forward_pairs(img1.cuda(), img2.cuda())
    with torch.no_grad():
        o_main, o_sup = StaOracle(state_dict, emulate_bf16=True).forward_pair(img1, img2)
    for res, orc in ((main, o_main), (sup, o_sup)):
        for k in TOL_EMU:
            assert maxn(res[k], orc[k]) < TOL_EMU[k], (k, maxn(res[k], orc[k]))


def test_pose_is_a_rigid_transform(cuda_model):
    img1, img2 = make_images(3, 64, 96, 5)
    main, sup = cuda_model.forward_pairs(img1.cuda(), img2.cuda())
    for res in (main, sup):
        P = res["relative_pose"].double().cpu()
        R = P[:, :3, :3]
        assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-5  # pp.mat2SE3 atol 1e-3
        assert (torch.det(R) - 1).abs().max() < 1e-5
        assert torch.equal(P[:, 3], torch.tensor([[0., 0., 0., 1.]], dtype=torch.float64).expand(3, 4))
        assert ((res["relative_pose_conf"] > 0) & (res["relative_pose_conf"] < 1)).all()
        assert (res["conf"] >= 1).all()


def test_full_size_properties_cfg2(cuda_model):
    """cfg-2 size (16 pairs, 512x384 bf16): size-independent properties -- finiteness, batch invariance (a pair's
    result does not depend on its batch neighbours) and view symmetry (swapping the views swaps the outputs)."""
    B, H, W = 16, 384, 512
    g = torch.Generator().manual_seed(3)
    i1 = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).bfloat16().cuda()
    i2 = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).bfloat16().cuda()
    main, sup = cuda_model.forward_pairs(i1, i2)
    for res in (main, sup):
        assert res["pts3d_pred"].shape == (B, H, W, 3) and torch.isfinite(res["pts3d_pred"]).all()
        assert torch.isfinite(res["relative_pose"]).all()
    again, _ = cuda_model.forward_pairs(i1, i2)
    assert torch.equal(again["pts3d_pred"], main["pts3d_pred"])         # deterministic kernels (no atomics)
    # no cross-pair op: a pair's result does not depend on its batch neighbours.  Bit-identical as long as the same
    # tile route is taken (8 vs 16 pairs); a single pair takes the small-problem route (128-wide tiles, split-K with a
    # fixed but different fp32 summation order), so it agrees to the bf16 rounding level instead.
    m8, s8 = cuda_model.forward_pairs(i1[4:12], i2[4:12])
    assert torch.equal(m8["pts3d_pred"][1], main["pts3d_pred"][5])
    assert torch.equal(s8["relative_pose"][1], sup["relative_pose"][5])
    m1, s1 = cuda_model.forward_pairs(i1[5:6], i2[5:6])
    assert maxn(m1["pts3d_pred"][0], main["pts3d_pred"][5]) < 2e-2
    assert maxn(s1["relative_pose"][0], sup["relative_pose"][5]) < 5e-3
    m2, s2 = cuda_model.forward_pairs(i2, i1)                             # swapped views
    assert torch.equal(m2["pts3d_pred"], sup["pts3d_pred"]) and torch.equal(s2["conf"], main["conf"])


def test_host_entry_point_matches_device_entry_point(cuda_model):
    img1, img2 = make_images(2, 64, 80, 9)
    h1, h2 = img1.bfloat16().pin_memory(), img2.bfloat16().pin_memory()
    out = cuda_model.forward_pairs_host(h1, h2)
    main, sup = cuda_model.forward_pairs(h1.cuda(), h2.cuda())
    assert torch.equal(out["pts3d"][0], main["pts3d_pred"].cpu()) and torch.equal(out["pose"][1], sup["relative_pose"].cpu())


def test_ragged_and_edge_shapes(cuda_model):
    # smallest image (1 token), odd token grids (crop path of refinenet4), portrait batch, B chunking (> 16 pairs)
    for (B, H, W) in ((1, 16, 16), (2, 48, 80), (1, 80, 48), (17, 32, 32)):
        img1, img2 = make_images(B, H, W, 11)
        main, sup = cuda_model.forward_pairs(img1.cuda(), img2.cuda())
        assert main["pts3d_pred"].shape == (B, H, W, 3) and torch.isfinite(main["pts3d_pred"]).all(), (B, H, W)
    with pytest.raises(AssertionError):
        cuda_model._encode_image(torch.zeros(1, 3, 30, 32, device="cuda"), None, normalize=False)
