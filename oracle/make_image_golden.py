"""
ORACLE TOOLING -- TEST INFRASTRUCTURE ONLY (build container only; needs /root/reference and Pillow).

Golden vectors for the image-loading path: runs the UNMODIFIED reference
``soft_intro_vae/dataset.py`` -- ``ImageDatasetFromFile.__getitem__`` (:74-78) = ``load_image`` (:12-47) + ``ToTensor``
(:66-68) -- with the arguments the image configs use (train_soft_intro_vae.py:388-392, 400-404: input_height=None,
crop_height=None, output_height=S, is_mirror=True) on synthetic PNG files, and records at its boundary: the decoded
source pixels, whether ``random.randint(0, 1) is 0`` mirrored the image (dataset.py:26), and the returned float32 tensor.

No reference source is copied; the module is imported from /root/reference.
Usage:  python oracle/make_image_golden.py    (writes tests/golden/image_pipeline.npz, < 300 kB; outputs stored as the uint8 v with tensor == v / 255)
"""
import os
import random
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "image_pipeline.npz")


def synth(rng, h, w):
    """decoded-photo-like content: smooth gradients + texture + saturated patches (exercises the clip8 saturation)"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(xx / (3.0 + c) + yy / 7.0) * np.cos(yy / (5.0 + 2 * c)) for c in range(3)], -1)
    img += rng.normal(0, 25, img.shape)
    img = np.clip(img, 0, 255).astype(np.uint8)
    img[: h // 5, : w // 4] = 255
    img[-(h // 6):, -(w // 3):] = 0
    img[h // 2, :] = 255
    return img


def main():
    from PIL import Image
    warnings.simplefilter("ignore", SyntaxWarning)          # dataset.py:22-27 `is not 'RGB'`
    sys.path.insert(0, os.path.join(REF, "soft_intro_vae"))
    sys.modules.pop("dataset", None)
    import dataset as ref_dataset                            # the reference module, unmodified
    sys.path.pop(0)
    rng = np.random.default_rng(7)
    cases = [("celeba_like", 109, 89, 64, 4), ("hq_like", 96, 96, 24, 4), ("up_odd", 37, 45, 64, 3), ("same", 32, 32, 32, 2)]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, h, w, size, n in cases:
            files, srcs = [], []
            for i in range(n):
                a = synth(rng, h, w)
                fn = "%s_%d.png" % (name, i)
                Image.fromarray(a, "RGB").save(os.path.join(d, fn))          # PNG: lossless, so the decoded pixels == a
                files.append(fn)
                srcs.append(a)
            ds = ref_dataset.ImageDatasetFromFile(files, d, input_height=None, crop_height=None, output_height=size,
                                                  is_mirror=True)
            random.seed(1234)
            got = np.stack([ds[i].numpy() for i in range(n)])
            random.seed(1234)
            flags = np.array([1 if random.randint(0, 1) == 0 else 0 for _ in range(n)], dtype=np.uint8)   # dataset.py:26
            out[name + "/src"] = np.stack(srcs)
            out[name + "/mirror"] = flags
            # ToTensor is exactly uint8 / 255 in float32: store the bytes (4x smaller) after checking that claim
            out_u8 = np.rint(got.astype(np.float64) * 255.0).astype(np.uint8)
            assert np.array_equal(out_u8.astype(np.float32) / np.float32(255.0), got.astype(np.float32))
            out[name + "/out_u8"] = out_u8
            out[name + "/size"] = np.array([size], dtype=np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def replay_draws(n, is_mirror, crop, is_random_crop, wh_of):
    """the draws load_image takes from Python's `random` per image, in its order: mirror coin (:26), then cx1, cy1 of a random
    crop (:35-38).  wh_of(i) = size of image i when the crop happens.  -> flags [n], margins [n,4] (left, top, right, bottom)"""
    flags, margins = [], []
    for i in range(n):
        flags.append(1 if (is_mirror and random.randint(0, 1) == 0) else 0)
        if crop is not None:
            w, h = wh_of(i)
            ch, cw = crop
            if is_random_crop:
                cx1 = random.randint(0, w - cw); cx2 = w - cw - cx1
                cy1 = random.randint(0, h - ch); cy2 = h - ch - cy1
            else:
                cx2 = cx1 = int(round((w - cw) / 2.))
                cy2 = cy1 = int(round((h - ch) / 2.))
            margins.append((cx1, cy1, cx2, cy2))
        else:
            margins.append((0, 0, 0, 0))
    return np.array(flags, dtype=np.uint8), np.array(margins, dtype=np.int32)


def main_crop():
    """the other branches of load_image (dataset.py:29-44): two-stage resize, random / centre crops, grey images -- again the
    UNMODIFIED reference ImageDatasetFromFile on synthetic PNGs -> tests/golden/image_pipeline_crop.npz"""
    from PIL import Image
    warnings.simplefilter("ignore", SyntaxWarning)
    sys.path.insert(0, os.path.join(REF, "soft_intro_vae"))
    sys.modules.pop("dataset", None)
    import dataset as ref_dataset
    sys.path.pop(0)
    rng = np.random.default_rng(11)
    # name, source (h, w), n, kwargs of ImageDatasetFromFile
    cases = [
        ("two_stage", (109, 89), 3, dict(input_height=80, output_height=64, is_mirror=True)),
        ("two_stage_rect", (70, 95), 3, dict(input_height=56, input_width=72, output_height=40, output_width=48, is_mirror=True)),
        ("random_crop", (90, 77), 4, dict(input_height=None, crop_height=60, crop_width=50, output_height=32, is_random_crop=True, is_mirror=True)),
        ("center_crop_odd", (61, 64), 3, dict(input_height=None, crop_height=48, output_height=24, is_random_crop=False, is_mirror=True)),
        ("two_stage_random_crop", (100, 100), 4, dict(input_height=72, input_width=88, crop_height=64, output_height=48, is_random_crop=True, is_mirror=True)),
        ("two_stage_center_crop_gray", (75, 83), 3, dict(input_height=64, crop_height=51, crop_width=40, output_height=32, is_random_crop=False, is_mirror=True, is_gray=True)),
    ]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, (h, w), n, kw in cases:
            files, srcs = [], []
            for i in range(n):
                a = synth(rng, h, w)
                fn = "%s_%d.png" % (name, i)
                Image.fromarray(a, "RGB").save(os.path.join(d, fn))
                files.append(fn)
                srcs.append(a)
            ds = ref_dataset.ImageDatasetFromFile(files, d, **kw)
            random.seed(4321)
            got = np.stack([ds[i].numpy() for i in range(n)])
            ih = kw.get("input_height")
            iw = kw.get("input_width") if kw.get("input_width") is not None else ih
            crop = None
            if kw.get("crop_height") is not None:
                crop = (kw["crop_height"], kw.get("crop_width") if kw.get("crop_width") is not None else kw["crop_height"])
            random.seed(4321)
            flags, margins = replay_draws(n, kw.get("is_mirror", True), crop, kw.get("is_random_crop", False),
                                          lambda i: (iw, ih) if ih is not None else (w, h))
            out[name + "/src"] = np.stack(srcs)
            out[name + "/mirror"] = flags
            out[name + "/margins"] = margins
            out_u8 = np.rint(got.astype(np.float64) * 255.0).astype(np.uint8)
            assert np.array_equal(out_u8.astype(np.float32) / np.float32(255.0), got.astype(np.float32))
            out[name + "/out_u8"] = out_u8                                       # [n, C, oh, ow]
            oh = kw["output_height"]
            ow = kw.get("output_width") if kw.get("output_width") is not None else oh
            out[name + "/args"] = np.array([ih if ih is not None else -1, iw if iw is not None else -1,
                                            crop[0] if crop else -1, crop[1] if crop else -1, oh, ow,
                                            int(kw.get("is_random_crop", False)), int(kw.get("is_gray", False))], dtype=np.int32)
    dst = os.path.join(ROOT, "tests", "golden", "image_pipeline_crop.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    if "--crop" in sys.argv:
        main_crop()
    else:
        main()
        main_crop()
