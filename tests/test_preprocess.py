"""SLAM image preprocessing -- byte work, so the bar is BIT-EXACT.

CPU: the numpy oracle (oracle/preprocess_oracle.py: crop geometry + PIL's 8-bit Lanczos resampler + ToTensor /
Normalize / Grayscale) against golden outputs of the UNMODIFIED reference pipeline (PIL + torchvision,
tools/make_golden_preprocess.py): landscape, portrait, non-square resolution, odd sizes, up-scaling.
GPU: the CUDA kernels through the C ABI against the same golden vectors and against the oracle on further shapes
(identity passes, tiny frames), plus the module-level contract (shape query, no CPU path)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import preprocess_oracle as orc

G = np.load(os.path.join(GOLDEN_DIR, "preprocess.npz"))
CASES = sorted(set(k.rsplit("_", 1)[0] for k in G.files))


@pytest.mark.parametrize("case", CASES)
def test_oracle_is_bit_exact_with_reference_golden(case):
    H, W = [int(v) for v in G[case + "_hw"]]
    rgb, gray, _ = orc.process_image(orc.synthetic_frame(H, W), tuple(int(v) for v in G[case + "_res"]))
    assert np.array_equal(rgb, G[case + "_rgb"]) and np.array_equal(gray, G[case + "_gray"])


def test_geometry_and_coefficients_known_answers():
    g = orc.crop_resize_geometry(480, 640, (224, 224))
    assert g["crop"] == (10, 10, 630, 470) and g["resized"] == (301, 224) and g["final"] == (38, 0)   # 38.5 -> 38 (half to even)
    ksize, bounds, kk = orc.precompute_coeffs_8bpc(620, 301)
    assert ksize == 15 and kk.shape == (301, 15)
    assert np.all(kk.sum(axis=1) - (1 << 22) <= 8) and np.all((1 << 22) - kk.sum(axis=1) <= 8)   # rows sum to ~1.0 in fixed point
    with pytest.raises(NotImplementedError):
        orc.crop_resize_geometry(300, 300, (512, 384))   # the reference would draw a random orientation


def test_shape_query_and_no_cpu_path():
    from vista_slam_b200.datasets.slam_images_only import SLAM_image_only
    ds = SLAM_image_only([], resolution=(512, 384), device="cpu")
    assert ds.output_shape(480, 640) == (384, 512) and ds.output_shape(1280, 720) == (512, 384)   # portrait -> transposed
    with pytest.raises(RuntimeError):
        ds.process_image(np.zeros((480, 640, 3), dtype=np.uint8))


def test_host_geometry_and_coefficient_tables_match_oracle_exactly():
    """The library's host side (crop geometry, PIL coefficient windows in double precision) against the oracle over
    many frame sizes -- no GPU needed: these tables are all the kernels consume besides the pixels."""
    import ctypes
    from vista_slam_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(5)
    sizes = [(480, 640), (720, 1280), (1280, 720), (1080, 1920), (357, 491), (150, 200), (244, 500), (468, 468)]
    sizes += [(int(rs.randint(120, 1400)), int(rs.randint(120, 2000))) for _ in range(60)]
    checked = 0
    for (H, W) in sizes:
        for res in ((224, 224), (512, 384)):
            try:
                g = orc.crop_resize_geometry(H, W, res)
            except (NotImplementedError, AssertionError):
                out = (ctypes.c_int * 10)()
                assert L.sta_preprocess_geometry(H, W, res[0], res[1], 10, 10, out) != 0   # refused alike
                continue
            out = (ctypes.c_int * 10)()
            _lib.check(L.sta_preprocess_geometry(H, W, res[0], res[1], 10, 10, out))
            assert tuple(out) == g["crop"] + g["resized"] + g["final"] + tuple(int(v) for v in g["out"]), (H, W, res)
            l, t, r, b = g["crop"]
            for in_size, out_size in ((r - l, g["resized"][0]), (b - t, g["resized"][1])):
                ksize, bounds, kk = orc.precompute_coeffs_8bpc(in_size, out_size)
                cb = (ctypes.c_int * (2 * out_size))()
                ck = (ctypes.c_int * (out_size * ksize))()
                cks = ctypes.c_int()
                _lib.check(L.sta_preprocess_coeffs(in_size, out_size, ctypes.byref(cks), cb, ck, out_size * ksize))
                assert cks.value == ksize
                assert np.array_equal(np.frombuffer(cb, dtype=np.int32).reshape(out_size, 2), bounds)
                assert np.array_equal(np.frombuffer(ck, dtype=np.int32).reshape(out_size, ksize), kk), (in_size, out_size)
                checked += 1
    assert checked > 100


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_kernels_are_bit_exact_with_reference_golden(case):
    from vista_slam_b200.datasets.slam_images_only import SLAM_image_only
    H, W = [int(v) for v in G[case + "_hw"]]
    ds = SLAM_image_only([], resolution=tuple(int(v) for v in G[case + "_res"]))
    v = ds.process_image(orc.synthetic_frame(H, W), "dir/" + case + ".png")
    assert v["img_name"] == case + ".png"
    assert np.array_equal(v["rgb"].cpu().numpy(), G[case + "_rgb"])
    assert np.array_equal(v["gray"].cpu().numpy(), G[case + "_gray"])


@pytest.mark.gpu
def test_kernels_match_oracle_on_more_shapes():
    from vista_slam_b200.datasets.slam_images_only import SLAM_image_only
    # camera-size frame, identity vertical pass (crop height == resolution), exact 2x, tiny frame, full-HD portrait
    for (H, W), res in (((480, 640), (224, 224)), ((244, 500), (224, 224)), ((468, 468), (224, 224)),
                        ((64, 90), (32, 32)), ((1080, 1920), (512, 384)), ((1920, 1080), (512, 384))):
        frame = orc.synthetic_frame(H, W, seed=H + W)
        ds = SLAM_image_only([], resolution=res)
        v = ds.process_image(torch.from_numpy(frame).cuda())           # device-resident input
        rgb, gray, _ = orc.process_image(frame, res)
        assert v["rgb"].shape == rgb.shape, (H, W, res)
        assert np.array_equal(v["rgb"].cpu().numpy(), rgb) and np.array_equal(v["gray"].cpu().numpy(), gray), (H, W, res)
        v2 = ds.process_image(frame)                                   # second call: cached plan
        assert torch.equal(v2["rgb"], v["rgb"])
