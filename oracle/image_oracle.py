"""TEST INFRASTRUCTURE -- CPU restatement of the reference's image-loading arithmetic.

Path restated: ``soft_intro_vae/dataset.py:12-47`` ``load_image`` as the image configs call it
(``train_soft_intro_vae.py:388-392, 400-404, 415-417``: ``input_height=None, crop_height=None, output_height=S,
is_mirror=True``) = ``ImageOps.mirror`` (dataset.py:26-27) -> ``img.resize((S, S), Image.BICUBIC)`` (dataset.py:46)
followed by ``transforms.ToTensor()`` (dataset.py:66-68, 75).

The arithmetic lives in a third-party dependency that is not under /root/reference: **Pillow**
(``environment.yml`` pins ``pillow=8.0.1``; the container has Pillow 12.2.0 -- the resampling code restated here,
``src/libImaging/Resample.c``, is unchanged between the two: ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``, ``bicubic_filter`` with a = -0.5, PRECISION_BITS = 32-8-2) and
``torchvision.transforms.functional.to_tensor`` (uint8 HWC -> float32 CHW, ``.div(255)``).

Pinned: ``tests/test_image_oracle.py`` holds this file bit-exact against Pillow itself (the library the reference
calls) over a sweep of sizes, and against the committed fixture ``tests/golden/image_pipeline.npz`` produced by the
unmodified reference ``load_image`` + ``ToTensor`` (``oracle/make_image_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: #define PRECISION_BITS (32 - 8 - 2)


def bicubic_filter(x):
    """Resample.c bicubic_filter, a = -0.5 (Keys)"""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


BICUBIC_SUPPORT = 2.0


def precompute_coeffs(in_size, out_size, in0=0.0, in1=None):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc.  Returns (ksize, bounds[out,2] (xmin, count), kk[out,ksize]
    int32 fixed point).  Same operation order in IEEE doubles as the C code."""
    if in1 is None:
        in1 = float(in_size)
    scale = (in1 - in0) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)            # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            # normalize_coeffs_8bpc
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx, 0] = xmin
        bounds[xx, 1] = xmax
    return ksize, bounds, kk


def _clip8(acc):
    """Resample.c clip8: lookup of (acc >> PRECISION_BITS) in a table that saturates to [0, 255]"""
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img, out_size, axis):
    """one pass of ImagingResampleHorizontal_8bpc (axis=1) / Vertical_8bpc (axis=0) on a uint8 [H,W,C] array"""
    in_size = img.shape[axis]
    _, bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, out_h, out_w):
    """Image.resize((out_w, out_h), Image.BICUBIC) of an 8-bit image [H,W,C]: ImagingResample's two passes, horizontal
    first, each skipped when that axis keeps its size (Resample.c need_horizontal / need_vertical)."""
    x = img
    if out_w != x.shape[1]:
        x = resample_axis(x, out_w, 1)
    if out_h != x.shape[0]:
        x = resample_axis(x, out_h, 0)
    return np.ascontiguousarray(x)


def load_image_tensor(img_u8, out_h, out_w, mirror):
    """dataset.py:26-27 (mirror) + :46 (resize) + ToTensor: uint8 [H,W,C] -> float32 [C,out_h,out_w] in [0,1]"""
    x = img_u8[:, ::-1] if mirror else img_u8
    y = resize_bicubic_u8(np.ascontiguousarray(x), out_h, out_w)
    return (y.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1).copy()


def crop_margins(w, h, crop_w, crop_h, is_random_crop, rnd):
    """dataset.py:32-43: the (left, top, right, bottom) borders ImageOps.crop removes.  rnd: Python's `random` module (or a
    random.Random) -- the random crop draws cx1 then cy1 from it, after the mirror coin (:26)."""
    if is_random_crop:
        cx1 = rnd.randint(0, w - crop_w)
        cx2 = w - crop_w - cx1
        cy1 = rnd.randint(0, h - crop_h)
        cy2 = h - crop_h - cy1
    else:
        cx2 = cx1 = int(round((w - crop_w) / 2.))
        cy2 = cy1 = int(round((h - crop_h) / 2.))
    return cx1, cy1, cx2, cy2


def load_image_u8(img_u8, mirror, out_h, out_w, input_hw=None, margins=None):
    """load_image in full (dataset.py:12-47) on decoded pixels: mirror (:26-27) -> optional resize to (input_h, input_w)
    (:29-30) -> optional ImageOps.crop of the borders `margins` = (left, top, right, bottom) (:32-44) -> resize to
    (out_h, out_w) (:46).  Every intermediate is an 8-bit image, as in Pillow.  uint8 [H,W,C] -> uint8 [out_h,out_w,C]"""
    x = np.ascontiguousarray(img_u8[:, ::-1] if mirror else img_u8)
    if input_hw is not None:
        x = resize_bicubic_u8(x, int(input_hw[0]), int(input_hw[1]))
    if margins is not None:
        l, t, r, b = (int(v) for v in margins)
        x = np.ascontiguousarray(x[t:x.shape[0] - b, l:x.shape[1] - r])
    return resize_bicubic_u8(x, out_h, out_w)


def to_tensor(img_u8):
    """transforms.ToTensor (dataset.py:66-68): uint8 HWC -> float32 CHW / 255"""
    return (img_u8.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1).copy()


def batch(images_u8, mirror_flags, out_h, out_w):
    """images_u8: [B,H,W,C] uint8 -> [B,C,out_h,out_w] float32"""
    return np.stack([load_image_tensor(im, out_h, out_w, bool(m)) for im, m in zip(images_u8, mirror_flags)])
