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


if __name__ == "__main__":
    main()
