"""TEST INFRASTRUCTURE -- CPU restatement of the SLAM image preprocessing.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path is the CUDA library.

Restates, bit-exactly (integer / byte work; the float outputs are single correctly-rounded fp32 operations):
  * SLAM_image_only.process_image                   vista_slam/datasets/slam_images_only.py:22-34
  * _crop_resize_if_necessary_image_only            vista_slam/datasets/base/base_view_graph_dataset.py:171-225
  * rescale_image_depthmap / crop_image_depthmap    vista_slam/utils/cropping.py:54-84,102-118
  * PIL.Image.resize(..., resample=LANCZOS) on 8-bit RGB -- third-party: Pillow (unpinned in the reference's
    requirements; 12.2.0 in this container), src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc
    (PRECISION_BITS = 22), ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc, clip8.  The published
    algorithm is restated here; parity is anchored on the reference's own call site by tools/make_golden_preprocess.py
  * ImgNorm = ToTensor + Normalize(0.5, 0.5)        vista_slam/utils/image.py:13   (x / 255 - 0.5) / 0.5
  * ImgGray = ToTensor + Grayscale(1)               slam_images_only.py:20         0.2989 r + 0.587 g + 0.114 b

Pinned against the unmodified reference (PIL + torchvision) by tests/golden/preprocess.npz."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def crop_resize_geometry(H, W, resolution, w_edge=10, h_edge=10):
    """-> dict(crop=(l,t,r,b), resized=(rw_full, rh_full), final=(l2,t2), out=(out_w,out_h)).
    base_view_graph_dataset.py:181-223 with aug_crop <= 1 (the SLAM datasets' setting)."""
    cx, cy = int(W / 2), int(H / 2)
    mx, my = min(cx, W - cx), min(cy, H - cy)
    assert mx > W / 5 and my > H / 5
    l, t, r, b = cx - mx, cy - my, cx + mx, cy + my
    l, t = max(l, w_edge), max(t, h_edge)
    r, b = min(r, W - w_edge), min(b, H - h_edge)
    W1, H1 = r - l, b - t
    res = tuple(resolution)
    assert res[0] >= res[1]
    if H1 > 1.1 * W1:
        res = res[::-1]
    elif 0.9 < H1 / W1 < 1.1 and res[0] != res[1]:
        raise NotImplementedError("square image with a non-square resolution: the reference picks an orientation at random")
    scale_final = max(res[0] / W1, res[1] / H1) + 1e-8                       # cropping.py:68
    full = np.floor(np.array([W1, H1]) * scale_final).astype(int)            # cropping.py:69
    l2 = int(np.int32(np.round(full[0] / 2 - res[0] / 2)))                   # base_view_graph_dataset.py:220 (half to even)
    t2 = int(np.int32(np.round(full[1] / 2 - res[1] / 2)))
    return {"crop": (l, t, r, b), "resized": (int(full[0]), int(full[1])), "final": (l2, t2), "out": res}


def _lanczos(x):
    if -3.0 <= x < 3.0:
        def sinc(v):
            if v == 0.0:
                return 1.0
            v = v * math.pi
            return math.sin(v) / v
        return sinc(x) * sinc(x / 3)
    return 0.0


def precompute_coeffs_8bpc(in_size, out_size):
    """Resample.c precompute_coeffs (box = whole axis) + normalize_coeffs_8bpc -> (ksize, bounds [out][2], kk int32 [out][ksize])."""
    in0, in1 = np.float32(0.0), np.float32(in_size)
    scale = float(in1 - in0) / out_size
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = float(in0) + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img, out_size):
    """8bpc resampling of axis 0 of an array [n][...] uint8 (vertical pass; the horizontal pass is the same on a
    transposed view).  int64 accumulation equals PIL's int32 (no overflow: sum |k| < 2 for Lanczos)."""
    _, bounds, kk = precompute_coeffs_8bpc(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for yy in range(out_size):
        ymin, ymax = bounds[yy]
        k = kk[yy, :ymax].astype(np.int64).reshape((-1,) + (1,) * (img.ndim - 1))
        ss = (1 << (PRECISION_BITS - 1)) + (src[ymin:ymin + ymax] * k).sum(axis=0)
        out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_lanczos_resize_rgb8(img, out_w, out_h):
    """PIL.Image.resize((out_w, out_h), LANCZOS) for an HxWx3 uint8 array: horizontal pass, then vertical pass
    (ImagingResample, Resample.c); a pass whose size does not change is skipped."""
    H, W, _ = img.shape
    if out_w != W:
        img = _resample_axis0(img.transpose(1, 0, 2), out_w).transpose(1, 0, 2)
    if out_h != H:
        img = _resample_axis0(img, out_h)
    return np.ascontiguousarray(img)


def process_image(rgb, resolution=(224, 224), w_edge=10, h_edge=10):
    """slam_images_only.py:22-34 -> (rgb fp32 [3][h][w] in [-1, 1], gray fp32 [1][h][w] in [0, 1], uint8 [h][w][3])."""
    rgb = np.asarray(rgb, dtype=np.uint8)
    H, W, _ = rgb.shape
    g = crop_resize_geometry(H, W, resolution, w_edge, h_edge)
    l, t, r, b = g["crop"]
    img = pil_lanczos_resize_rgb8(rgb[t:b, l:r], *g["resized"])
    l2, t2 = g["final"]
    ow, oh = g["out"]
    img = img[t2:t2 + oh, l2:l2 + ow]
    f = img.astype(np.float32) / np.float32(255)                               # ToTensor
    f = f.transpose(2, 0, 1)
    norm = (f - np.float32(0.5)) / np.float32(0.5)                             # Normalize(0.5, 0.5)
    gray = (np.float32(0.2989) * f[0] + np.float32(0.587) * f[1] + np.float32(0.114) * f[2])[None]  # rgb_to_grayscale
    return np.ascontiguousarray(norm), np.ascontiguousarray(gray.astype(np.float32)), img


def synthetic_frame(H, W, seed=77):
    """Seeded test frame shared by the golden generator and the tests (inputs are regenerated, not stored): smooth
    gradients + texture + hard edges + saturated patches (exercises the negative Lanczos lobes and clip8)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.stack([127 + 120 * np.sin(xx / 17.0 + yy / 31.0), 127 + 120 * np.cos(xx / 7.0) * np.sin(yy / 11.0),
                    255.0 * ((xx // 16 + yy // 16) % 2)], axis=-1)
    img += rs.randn(H, W, 3) * 12
    img[H // 4:H // 4 + 20, W // 3:W // 3 + 40] = 255
    img[H // 2:H // 2 + 9, :] = 0
    return np.clip(img, 0, 255).astype(np.uint8)
