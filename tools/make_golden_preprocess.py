"""Golden vectors for the SLAM image preprocessing: runs the UNMODIFIED reference
(vista_slam/datasets/slam_images_only.py::SLAM_image_only.process_image -> PIL LANCZOS + torchvision transforms) in
this container on seeded synthetic frames and stores inputs + outputs in tests/golden/preprocess.npz.  munch and
colorama are absent here and are shimmed (a dict with attribute access; print colours) -- neither touches the pixels."""
import os
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def load_reference():
    sys.path.insert(0, "/root/reference")
    for name in ("munch", "colorama", "networkx"):
        try:
            __import__(name)
        except ImportError:
            m = types.ModuleType(name)
            if name == "colorama":
                class _Any:
                    def __getattr__(self, k):
                        return ""
                m.Fore = _Any()
                m.Style = _Any()
            if name == "munch":
                class Munch(dict):
                    __getattr__ = dict.get

                    def __setattr__(self, k, v):
                        self[k] = v
                m.Munch = Munch
            sys.modules[name] = m
    from vista_slam.datasets.slam_images_only import SLAM_image_only
    return SLAM_image_only


def main():
    SLAM_image_only = load_reference()
    from oracle.preprocess_oracle import synthetic_frame   # shared, seeded input generator (inputs are not stored)
    out = {}
    # (H, W) of the frame, (res_w, res_h): landscape downscale, portrait, non-square resolution, odd sizes, upscale
    cases = {"qvga_224": ((300, 400), (224, 224)), "portrait_224": ((640, 360), (224, 224)),
             "wide_256x192": ((420, 630), (256, 192)), "odd_224": ((357, 491), (224, 224)),
             "upscale_224": ((150, 200), (224, 224))}
    for name, ((H, W), res) in cases.items():
        frame = synthetic_frame(H, W, seed=77)
        ds = SLAM_image_only([], resolution=res)
        v = ds.process_image(frame, name + ".png")
        out[name + "_hw"] = np.array([H, W])
        out[name + "_res"] = np.array(res)
        out[name + "_rgb"] = v["rgb"].numpy()
        out[name + "_gray"] = v["gray"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
