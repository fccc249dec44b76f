"""Golden vectors for the pointmap consumers: runs the UNMODIFIED reference functions of
/root/reference/vista_slam/utils/slam_utils.py (this container only; colorama is absent and is shimmed with an
empty module -- it only provides print colours) on small seeded inputs and stores inputs + outputs in
tests/golden/slam_utils.npz."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference/vista_slam/utils/slam_utils.py"


def load_reference():
    if "colorama" not in sys.modules:
        col = types.ModuleType("colorama")

        class _Any:
            def __getattr__(self, k):
                return ""
        col.Fore = _Any()
        col.Style = _Any()
        sys.modules["colorama"] = col
    spec = importlib.util.spec_from_file_location("ref_slam_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    g = torch.Generator().manual_seed(4321)
    out = {}
    for name, (B, H, W) in {"a": (2, 24, 32), "b": (3, 17, 41)}.items():
        # a plausible camera: points on rays through the pixel grid, noisy depth, a few degenerate pixels
        fx, fy = 0.9 * W, 1.1 * H
        jj, ii = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        z = 1.0 + 2.0 * torch.rand(B, H, W, generator=g)
        x = (ii.float() - W / 2.0) / fx * z + 0.01 * torch.randn(B, H, W, generator=g)
        y = (jj.float() - H / 2.0) / fy * z + 0.01 * torch.randn(B, H, W, generator=g)
        pts = torch.stack([x, y, z], dim=-1).contiguous()
        pts[0, 0, 0] = torch.tensor([0.3, -0.2, 0.0])   # X/Z = inf  -> 0
        pts[0, 1, 1] = torch.tensor([0.0, 0.0, 0.0])    # 0/0 = nan  -> 0
        pts[B - 1, 2, 3] = torch.tensor([-0.5, 0.1, -0.0])
        conf = 1.0 + torch.rand(B, H, W, generator=g) * 3.0
        conf[0, 3, 3] = 0.0                               # clamped to 1e-6
        out[name + "_pts3d"] = pts.numpy()
        out[name + "_conf"] = conf.numpy()
        out[name + "_K_shared"] = ref.estimate_intrinsic_from_pts3d(pts, conf, shared_intrinsic=True).numpy()
        out[name + "_K_each"] = ref.estimate_intrinsic_from_pts3d(pts, conf, shared_intrinsic=False).numpy()
        out[name + "_depth"] = pts[..., 2].contiguous().numpy()                      # slam.py:185
        out[name + "_conf_mean"] = np.array([conf[b].mean().item() for b in range(B)], dtype=np.float32)  # pose_graph.py:41
        Di, Dj = pts[0, ..., 2], pts[1, ..., 2] * 1.7
        ci, cj = conf[0], conf[1]
        out[name + "_scale"] = np.float32(ref.estimate_scale_with_depth_and_confidence(Di, Dj, ci, cj).item())
        out[name + "_scale_conf"] = np.float32((ci * cj).sqrt().mean().item())       # slam.py:227
        out[name + "_Dj"] = Dj.numpy()
    path = os.path.join(ROOT, "tests", "golden", "slam_utils.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
