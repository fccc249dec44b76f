"""TEST INFRASTRUCTURE -- CPU restatement of the pointmap consumers that follow the STA heads in OnlineSLAM
(SURVEY.md section 8(f) rank 2).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this
module; the product path is the CUDA library.

Restates, in numpy with float64 accumulation:
  * estimate_intrinsic_from_pts3d        vista_slam/utils/slam_utils.py:8-79   (call site slam.py:184)
  * depths = pcls[..., 2]                vista_slam/slam.py:185
  * conf.mean()                          vista_slam/pose_graph.py:41
  * estimate_scale_with_depth_and_confidence   vista_slam/utils/slam_utils.py:168-190 (call site slam.py:224-226)
  * scale_conf = sqrt(ci * cj).mean()    vista_slam/slam.py:227

Pinned against the unmodified reference functions by tools/make_golden_slam_utils.py ->
tests/golden/slam_utils.npz (tests/test_oracle_golden.py)."""
import numpy as np


def _ratio_nan_to_zero(a, b):
    """torch.nan_to_num(a / b, nan=0, posinf=0, neginf=0) in float32 (slam_utils.py:39-40)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        q = a.astype(np.float32) / b.astype(np.float32)
    q[~np.isfinite(q)] = 0.0
    return q


def estimate_intrinsic_from_pts3d(pts3d, confidence, shared_intrinsic=False):
    """pts3d [B,H,W,3] fp32, confidence [B,H,W] fp32 -> K [3,3] (shared) or [B,3,3], float32.

    Weighted least squares of u = fx * X/Z and v = fy * Y/Z around the image centre (slam_utils.py:20-79)."""
    pts3d = np.asarray(pts3d, dtype=np.float32)
    confidence = np.asarray(confidence, dtype=np.float32)
    B, H, W, _ = pts3d.shape
    cx, cy = W / 2.0, H / 2.0
    u = (np.arange(W, dtype=np.float32) - np.float32(cx))[None, :].repeat(H, 0).reshape(1, -1)   # slam_utils.py:25-31
    v = (np.arange(H, dtype=np.float32) - np.float32(cy))[:, None].repeat(W, 1).reshape(1, -1)
    X = pts3d[..., 0].reshape(B, -1)
    Y = pts3d[..., 1].reshape(B, -1)
    Z = pts3d[..., 2].reshape(B, -1)
    w = np.maximum(confidence.reshape(B, -1), np.float32(1e-6))                                     # :37
    xz = _ratio_nan_to_zero(X, Z)
    yz = _ratio_nan_to_zero(Y, Z)
    f64 = np.float64
    fx_num = (w.astype(f64) * xz * u).sum(axis=1)
    fx_den = (w.astype(f64) * xz * xz).sum(axis=1)
    fy_num = (w.astype(f64) * yz * v).sum(axis=1)
    fy_den = (w.astype(f64) * yz * yz).sum(axis=1)
    if shared_intrinsic:                                                                            # :42-60
        with np.errstate(divide="ignore", invalid="ignore"):
            fx = fx_num.sum() / fx_den.sum()
            fy = fy_num.sum() / fy_den.sum()
        return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        fx = fx_num / fx_den                                                                        # :62-79
        fy = fy_num / fy_den
    K = np.zeros((B, 3, 3), dtype=np.float32)
    K[:, 0, 0] = fx
    K[:, 1, 1] = fy
    K[:, 0, 2] = cx
    K[:, 1, 2] = cy
    K[:, 2, 2] = 1.0
    return K


def depth_and_mean_conf(pts3d, confidence):
    """depths = pts3d[..., 2] (slam.py:185) and the per-view confidence mean used to pick a view's best node
    (pose_graph.py:41)."""
    pts3d = np.asarray(pts3d, dtype=np.float32)
    confidence = np.asarray(confidence, dtype=np.float32)
    B = pts3d.shape[0]
    return np.ascontiguousarray(pts3d[..., 2]), confidence.reshape(B, -1).astype(np.float64).mean(axis=1).astype(np.float32)


def estimate_scale_with_depth_and_confidence(Di, Dj, ci, cj):
    """s with Dj ~ s * Di under weights w = max(ci * cj, 1e-6) (slam_utils.py:168-190), float32 scalar."""
    Di, Dj, ci, cj = [np.asarray(a, dtype=np.float32).reshape(-1) for a in (Di, Dj, ci, cj)]
    w = np.maximum(ci * cj, np.float32(1e-6)).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float32((w * Di * Dj).sum() / (w * Di * Di).sum())


def scale_confidence(ci, cj):
    """(ci * cj).sqrt().mean() (slam.py:227), float32 scalar."""
    ci, cj = [np.asarray(a, dtype=np.float32).reshape(-1) for a in (ci, cj)]
    return np.float32(np.sqrt(ci * cj).astype(np.float64).mean())
