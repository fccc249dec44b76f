"""Pointmap consumers.

CPU: the numpy oracle (oracle/slam_utils_oracle.py) against golden vectors produced by the UNMODIFIED reference
functions (tools/make_golden_slam_utils.py), incl. the x/0, 0/0 and conf < 1e-6 pixels; the GPU mirror refuses CPU
tensors.  GPU: the CUDA kernels (through the C ABI) against the golden vectors, against the oracle at cfg-2 size, and
size-independent properties (scale equivariance, planted intrinsics).  Tolerance: fp32 sums of <= 4e5 terms, 1e-4
relative (the reference itself sums in fp32; oracle and kernels combine partials in fp64)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import slam_utils_oracle as orc

G = np.load(os.path.join(GOLDEN_DIR, "slam_utils.npz"))
RTOL = 1e-4


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b))), (a, b)


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_matches_reference_golden(case):
    p, c = G[case + "_pts3d"], G[case + "_conf"]
    _close(orc.estimate_intrinsic_from_pts3d(p, c, True), G[case + "_K_shared"])
    _close(orc.estimate_intrinsic_from_pts3d(p, c, False), G[case + "_K_each"])
    d, m = orc.depth_and_mean_conf(p, c)
    assert np.array_equal(d, G[case + "_depth"])
    _close(m, G[case + "_conf_mean"])
    _close(orc.estimate_scale_with_depth_and_confidence(p[0, ..., 2], G[case + "_Dj"], c[0], c[1]), G[case + "_scale"])
    _close(orc.scale_confidence(c[0], c[1]), G[case + "_scale_conf"])


def test_gpu_mirror_has_no_cpu_path():
    from vista_slam_b200.utils import slam_utils as su
    with pytest.raises(RuntimeError):
        su.estimate_intrinsic_from_pts3d(torch.zeros(1, 4, 4, 3), torch.ones(1, 4, 4))
    with pytest.raises(RuntimeError):
        su.estimate_scale_with_depth_and_confidence(*[torch.ones(4)] * 4)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_kernels_match_reference_golden(case):
    from vista_slam_b200.utils import slam_utils as su
    p = torch.from_numpy(G[case + "_pts3d"]).cuda()
    c = torch.from_numpy(G[case + "_conf"]).cuda()
    K, depth, cmean = su.pointmap_consumers(p, c, shared_intrinsic=True)
    _close(K.cpu().numpy(), G[case + "_K_shared"])
    assert np.array_equal(depth.cpu().numpy(), G[case + "_depth"])  # a copy: bit exact
    _close(cmean.cpu().numpy(), G[case + "_conf_mean"])
    _close(su.estimate_intrinsic_from_pts3d(p, c, shared_intrinsic=False).cpu().numpy(), G[case + "_K_each"])
    Dj = torch.from_numpy(G[case + "_Dj"]).cuda()
    s, sc = su.scale_and_confidence(p[0, ..., 2], Dj, c[0], c[1])
    _close(s.item(), G[case + "_scale"])
    _close(sc.item(), G[case + "_scale_conf"])
    _close(su.estimate_scale_with_depth_and_confidence(p[0, ..., 2], Dj, c[0], c[1]).item(), G[case + "_scale"])


@pytest.mark.gpu
def test_kernels_match_oracle_at_cfg2_size_and_properties():
    from vista_slam_b200.utils import slam_utils as su
    g = torch.Generator().manual_seed(7)
    B, H, W = 4, 384, 512
    fx, fy = 410.0, 395.0
    jj, ii = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    z = 0.5 + 3.0 * torch.rand(B, H, W, generator=g)
    pts = torch.stack([(ii.float() - W / 2.0) / fx * z, (jj.float() - H / 2.0) / fy * z, z], dim=-1).contiguous()
    conf = 1.0 + torch.rand(B, H, W, generator=g)
    K, depth, cmean = su.pointmap_consumers(pts.cuda(), conf.cuda(), shared_intrinsic=False)
    # planted pinhole camera is recovered (noise-free rays): property, independent of any reference
    _close(K[:, 0, 0].cpu().numpy(), np.full(B, fx), 1e-3)
    _close(K[:, 1, 1].cpu().numpy(), np.full(B, fy), 1e-3)
    # oracle parity at full size
    _close(K.cpu().numpy(), orc.estimate_intrinsic_from_pts3d(pts.numpy(), conf.numpy(), False))
    _close(su.estimate_intrinsic_from_pts3d(pts.cuda(), conf.cuda(), True).cpu().numpy(),
           orc.estimate_intrinsic_from_pts3d(pts.numpy(), conf.numpy(), True))
    d_ref, m_ref = orc.depth_and_mean_conf(pts.numpy(), conf.numpy())
    assert np.array_equal(depth.cpu().numpy(), d_ref)
    _close(cmean.cpu().numpy(), m_ref)
    # scale: Dj = 2.5 * Di exactly -> s = 2.5; scaling the points does not change K (X/Z, Y/Z are ratios)
    Di = pts[0, ..., 2].cuda()
    s, sc = su.scale_and_confidence(Di, 2.5 * Di, conf[0].cuda(), conf[1].cuda())
    _close(s.item(), 2.5, 1e-5)
    _close(sc.item(), orc.scale_confidence(conf[0].numpy(), conf[1].numpy()))
    K2 = su.estimate_intrinsic_from_pts3d((3.0 * pts).cuda(), conf.cuda(), False)
    _close(K2.cpu().numpy(), K.cpu().numpy(), 1e-5)
    # noisy depths against the oracle
    Dj = (1.7 * pts[1, ..., 2] + 0.05 * torch.randn(H, W, generator=g))
    s, _ = su.scale_and_confidence(Di, Dj.cuda(), conf[0].cuda(), conf[1].cuda())
    _close(s.item(), orc.estimate_scale_with_depth_and_confidence(pts[0, ..., 2].numpy(), Dj.numpy(), conf[0].numpy(),
                                                                  conf[1].numpy()))
