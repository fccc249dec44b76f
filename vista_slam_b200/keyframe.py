"""Keyframe step of the SLAM frontend.

`OnlineSLAM` (vista_slam/slam.py) calls the STA model once per keyframe for the encoder (add_view, :142-151) and once
per candidate edge for decoder + heads (regress_two_views, :153-189), each edge with its own host synchronisation
(`rel_pose_conf_ij < thres`, :169).  `KeyframeFrontend` keeps the same per-view cache (encoder features on the GPU)
but regresses ALL edges of a keyframe in one batched call (sta_regress_pairs): one decoder pass over K feature pairs,
both pose heads, both DPT heads and the pointmap consumers (per-edge shared intrinsics, depth maps, confidence means),
with no host synchronisation inside -- the caller reads `pose_conf` once per keyframe.

Results are numerically the per-edge results of the reference sequence (the decoder has no cross-sample op), so
`regress_two_views(i, j)` is also offered as the K = 1 special case with the reference's return convention."""
import torch

from . import _lib
from .utils.slam_utils import _scratch_for


class KeyframeFrontend:
    def __init__(self, sta):
        self.frontend = sta
        self.enc_features = []   # like OnlineSLAM.enc_features (slam.py:145)
        self.img_shapes = []

    @property
    def view_num(self):
        return len(self.enc_features)

    @torch.no_grad()
    def add_view(self, image, image_shape=None):
        """slam.py:142-151: encode one view (B = 1) and cache its features; returns the view index."""
        feat, _ = self.frontend._encode_image(image, image_shape, normalize=False)
        self.enc_features.append(feat)
        self.img_shapes.append((int(image.shape[-2]), int(image.shape[-1])))
        return len(self.enc_features) - 1

    @torch.no_grad()
    def regress_views(self, i, js):
        """All edges (i, j) for j in js in one batch.  Returns a dict of CUDA tensors, edge-major:
        pose [K,4,4] and pose_conf [K] of the i -> j direction (slam.py:164-167), pts3d [2,K,H,W,3], conf [2,K,H,W]
        ([0] = ij, [1] = ji as in the torch.cat of slam.py:181-182), intri [K,3,3], depths [2,K,H,W], conf_mean [2,K]."""
        js = list(js)
        K = len(js)
        if K == 0:
            raise ValueError("no edges")
        H, W = self.img_shapes[i]
        if any(self.img_shapes[j] != (H, W) for j in js):
            raise ValueError("all views of a batch must have the same size")
        m = self.frontend
        fi = self.enc_features[i].expand(K, -1, -1).contiguous() if K > 1 else self.enc_features[i]
        fj = torch.cat([self.enc_features[j] for j in js], dim=0) if K > 1 else self.enc_features[js[0]]
        L = m._ready(fi)
        dev = fi.device
        f32 = dict(device=dev, dtype=torch.float32)
        pose = torch.empty(2, K, 4, 4, **f32)
        pconf = torch.empty(2, K, **f32)
        pts = torch.empty(2, K, H, W, 3, **f32)
        conf = torch.empty(2, K, H, W, **f32)
        intri = torch.empty(K, 3, 3, **f32)
        depth = torch.empty(2, K, H, W, **f32)
        cmean = torch.empty(2, K, **f32)
        with torch.cuda.device(dev):
            _lib.check(L.sta_regress_pairs(m._handle, _lib.ptr(fi), _lib.ptr(fj), K, H, W, _lib.ptr(pose), _lib.ptr(pconf),
                                           _lib.ptr(pts), _lib.ptr(conf), _lib.ptr(intri), _lib.ptr(depth), _lib.ptr(cmean),
                                           _lib.ptr(_scratch_for(dev, 2 * K)), _lib.cur_stream()), "sta_regress_pairs")
        return {"pose": pose[0], "pose_conf": pconf[0], "pose_ji": pose[1], "pose_conf_ji": pconf[1], "pts3d": pts,
                "conf": conf, "intri": intri, "depths": depth, "conf_mean": cmean}

    @torch.no_grad()
    def regress_views_gated(self, i, js, rel_pose_thres):
        """The keyframe step with the reference's early-out (slam.py:169-170) kept: ONE decoder + pose-head pass over all
        candidate edges (i, j), ONE device-to-host copy of the K pose confidences, then the DPT heads and the pointmap
        consumers only for the edges the reference would keep (`conf >= thres or i - j == 1`).

        Returns a list with one entry per candidate, in the return convention of OnlineSLAM.regress_two_views
        (slam.py:153-189) minus the pypose conversion, which stays with the caller:
            (pose_ij [1,4,4], rel_pose_conf_ij [1], confs [2,H,W] | None, intri [3,3] | None, depths [2,H,W] | None)
        -- the last three are None for rejected edges, exactly where the reference returns None."""
        import ctypes
        js = list(js)
        K = len(js)
        if K == 0:
            return []
        H, W = self.img_shapes[i]
        if any(self.img_shapes[j] != (H, W) for j in js):
            raise ValueError("all views of a batch must have the same size")
        m = self.frontend
        fi = self.enc_features[i].expand(K, -1, -1).contiguous() if K > 1 else self.enc_features[i]
        fj = torch.cat([self.enc_features[j] for j in js], dim=0) if K > 1 else self.enc_features[js[0]]
        L = m._ready(fi)
        dev = fi.device
        f32 = dict(device=dev, dtype=torch.float32)
        pose = torch.empty(2, K, 4, 4, **f32)
        pconf = torch.empty(2, K, **f32)
        with torch.cuda.device(dev):
            _lib.check(L.sta_regress_pairs_begin(m._handle, _lib.ptr(fi), _lib.ptr(fj), K, H, W, _lib.ptr(pose), _lib.ptr(pconf),
                                                 _lib.cur_stream()), "sta_regress_pairs_begin")
            conf_host = pconf[0].cpu()  # the one host synchronisation of the keyframe (slam.py:169 has one per edge)
            keep = [k for k, j in enumerate(js) if not (float(conf_host[k]) < rel_pose_thres and i - j != 1)]
            n = len(keep)
            out = [(pose[0, k:k + 1], pconf[0, k:k + 1], None, None, None) for k in range(K)]
            if n == 0:
                return out
            pts = torch.empty(2, n, H, W, 3, **f32)
            conf = torch.empty(2, n, H, W, **f32)
            intri = torch.empty(n, 3, 3, **f32)
            depth = torch.empty(2, n, H, W, **f32)
            idx = (ctypes.c_int * n)(*keep)
            _lib.check(L.sta_regress_pairs_finish(m._handle, idx, n, _lib.ptr(pts), _lib.ptr(conf), _lib.ptr(intri),
                                                  _lib.ptr(depth), None, _lib.ptr(_scratch_for(dev, 2 * n)), _lib.cur_stream()),
                       "sta_regress_pairs_finish")
        for pos, k in enumerate(keep):
            out[k] = (pose[0, k:k + 1], pconf[0, k:k + 1], conf[:, pos], intri[pos], depth[:, pos])
        self.last_pts3d = pts  # [2, n, H, W, 3] of the kept edges (diagnostics / tests)
        self.last_kept = keep
        return out

    @torch.no_grad()
    def regress_two_views(self, i, j):
        """K = 1 with the return convention of slam.py:153-189 minus the pypose conversion:
        (pose_ij [1,4,4], rel_pose_conf_ij [1], confs [2,H,W], intri [3,3], depths [2,H,W])."""
        r = self.regress_views(i, [j])
        return r["pose"], r["pose_conf"], r["conf"][:, 0], r["intri"][0], r["depths"][:, 0]
