"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(tools/make_golden.py).  This is what pins the oracle (the reference ships no tests of its own)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.sta_oracle import StaOracle, flops_per_pair, make_images, rope2d, state_dict_spec, token_positions


def test_state_dict_spec_matches_reference_dump():
    spec = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_spec.json")))
    mine = [[k, list(s)] for k, s in state_dict_spec()]
    assert mine == spec
    assert len(spec) == 665
    n = 0
    seen = set()
    for k, s in spec:
        if "scratch.layer_rn." in k:
            continue  # aliases of scratch.layerN_rn (same Parameter)
        n += int(np.prod(s))
        seen.add(k)
    assert n == 438455505


def test_flop_model_matches_survey():
    # 435.8 GF @224^2, 1857.5 GF @512x384 (FlopCounterMode on the reference)
    assert abs(flops_per_pair(224, 224) / 1e9 - 435.8) < 0.1
    assert abs(flops_per_pair(384, 512) / 1e9 - 1857.5) < 0.1


def _maxn(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("case", ["pair_64x80", "pair_b2_48x64", "pair_224x224", "pair_384x512"])
def test_oracle_fp32_matches_reference_golden(state_dict, case):
    """incl. cfg-1 (224x224, the reference's native size) and one cfg-2 pair (512x384)."""
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(g["meta"]))
    img1, img2 = make_images(meta["B"], meta["H"], meta["W"], meta["image_seed"])
    if meta.get("bf16_images"):
        img1, img2 = img1.bfloat16().float(), img2.bfloat16().float()
    orc = StaOracle(state_dict, emulate_bf16=False)
    with torch.no_grad():
        f1, pos1 = orc.encode_image(img1)
        f2, pos2 = orc.encode_image(img2)
        d1, d2 = orc.decode_stereo(f1, f2, pos1, pos2)
        main, sup = orc.forward_pair(img1, img2)
    if "pos1" in g:
        assert torch.equal(pos1, torch.from_numpy(g["pos1"]))
    tol = 2e-4  # fp32 summation-order noise through 36 blocks; measured 1e-6 .. 3e-5
    assert _maxn(f1, torch.from_numpy(g["enc_feat1"])) < tol
    for k, t in (("dec1_6", d1[6]), ("dec1_9", d1[9]), ("dec1_12", d1[12]), ("dec2_12", d2[12])):
        if k in g:  # the large fixture keeps one feature per stage
            assert _maxn(t, torch.from_numpy(g[k])) < tol, k
    for pre, res in (("main_", main), ("support_", sup)):
        assert _maxn(res["pts3d_pred"], torch.from_numpy(g[pre + "pts3d"])) < tol
        assert _maxn(res["conf"], torch.from_numpy(g[pre + "conf"])) < tol
        assert _maxn(res["relative_pose"], torch.from_numpy(g[pre + "pose"])) < tol
        assert _maxn(res["relative_pose_conf"], torch.from_numpy(g[pre + "pose_conf"])) < tol


@pytest.mark.parametrize("case", ["views_portrait_80x48_s3", "views_mixed_b2_64x80_s2"])
def test_oracle_forward_views_matches_reference_golden(state_dict, case):
    """general forward(views) loop: several support views, all-portrait and mixed-orientation batches
    (tools/make_golden_views.py ran the unmodified reference)."""
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(g["meta"]))
    n_sup = meta["neighbors"] + meta["loops"]
    imgs = [make_images(meta["B"], meta["H"], meta["W"], meta["image_seed"] + k)[0] for k in range(n_sup + 1)]
    ts = torch.tensor(meta["true_shape"])
    with torch.no_grad():
        main, sup = StaOracle(state_dict, emulate_bf16=False).forward_views(imgs[0], ts, [(im, ts) for im in imgs[1:]])
    for side, res in (("main", main), ("support", sup)):
        assert len(res) == n_sup
        for i in range(n_sup):
            for k in ("pts3d_pred", "conf", "relative_pose", "relative_pose_conf"):
                gold = torch.from_numpy(g["%s%d_%s" % (side, i, k)])
                assert res[i][k].shape == gold.shape
                assert _maxn(res[i][k], gold) < 2e-4, (side, i, k)


def test_oracle_bf16_emulation_stays_near_fp32(state_dict):
    img1, img2 = make_images(1, 64, 80, 1234)
    with torch.no_grad():
        a, _ = StaOracle(state_dict, emulate_bf16=False).forward_pair(img1, img2)
        b, _ = StaOracle(state_dict, emulate_bf16=True).forward_pair(img1, img2)
    # bf16 operand noise floor: ~1e-2 on the trunk features; pts3d = dir * expm1(d) amplifies an absolute error
    # in d (up to ~3 here) by e^d/(e^d - 1) * d, hence the looser bound on the pointmap
    assert _maxn(b["pts3d_pred"], a["pts3d_pred"]) < 1e-1
    assert _maxn(b["relative_pose"], a["relative_pose"]) < 3e-2
    print("bf16-emulation vs fp32: pts3d %.3e pose %.3e conf %.3e" % (
        _maxn(b["pts3d_pred"], a["pts3d_pred"]), _maxn(b["relative_pose"], a["relative_pose"]), _maxn(b["conf"], a["conf"])))


def test_rope_negative_positions_and_norm():
    # pose token sits at (-1,-1) (sta_model.py:214-219); RoPE is a rotation -> preserves pair norms
    t = torch.randn(1, 2, 3, 64)
    pos = torch.tensor([[[-1, -1], [0, 0], [5, 7]]])
    r = rope2d(t, pos)
    assert torch.allclose(r[:, :, 1], t[:, :, 1])  # position 0 = identity
    assert torch.allclose(r.norm(dim=-1), t.norm(dim=-1), atol=1e-5)
    assert not torch.allclose(r[:, :, 0], t[:, :, 0])


def test_token_positions_order():
    pos = token_positions(2, 3, 4)
    assert pos.shape == (2, 12, 2) and pos.dtype == torch.int64
    assert pos[0, 5].tolist() == [1, 1] and pos[1, 11].tolist() == [2, 3]
