"""CPU: the image-pipeline oracle (oracle/image_oracle.py) against (1) the golden vectors recorded from the unmodified
reference dataset.py (tests/golden/image_pipeline.npz, oracle/make_image_golden.py) and (2) Pillow itself -- the
library the reference calls -- over a sweep of sizes; and the C-ABI's host-side coefficient routine against the oracle's
(integer tables: bit-exact)."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

from oracle import image_oracle as IO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "image_pipeline.npz")
PKG = "soft-intro-vae-pytorch_b200"


def golden_cases():
    z = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in z.files})
    return [(n, z[n + "/src"], z[n + "/mirror"], z[n + "/out_u8"], int(z[n + "/size"][0])) for n in names]


def test_oracle_matches_reference_golden():
    for name, src, mirror, out_u8, size in golden_cases():
        got = IO.batch(src, mirror, size, size)
        want = out_u8.astype(np.float32) / np.float32(255.0)
        assert got.dtype == np.float32 and got.shape == want.shape, name
        assert np.array_equal(got, want), name
    assert any(m.any() for _, _, m, _, _ in golden_cases()) and any((1 - m).any() for _, _, m, _, _ in golden_cases())


@pytest.mark.parametrize("shape", [((218, 178), (128, 128)), ((64, 64), (32, 32)), ((100, 37), (13, 91)), ((33, 47), (47, 33)),
                                   ((40, 40), (40, 17)), ((17, 29), (64, 128)), ((300, 200), (9, 7)), ((5, 5), (64, 64)),
                                   ((1, 7), (4, 4)), ((31, 31), (31, 31))])
@pytest.mark.parametrize("ch", [3, 1])
def test_oracle_matches_pillow(shape, ch):
    from PIL import Image, ImageOps
    (h, w), (oh, ow) = shape
    rng = np.random.default_rng(h * 1000 + w + ch)
    a = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
    a[: h // 2, : w // 2] = 255
    a[h // 2:, w // 2:] = 0
    im = Image.fromarray(a if ch == 3 else a[..., 0], "RGB" if ch == 3 else "L")
    for mirror in (False, True):
        ref = np.asarray((ImageOps.mirror(im) if mirror else im).resize((ow, oh), Image.BICUBIC))
        ref = ref if ch == 3 else ref[..., None]
        got = IO.load_image_tensor(a, oh, ow, mirror)
        assert np.array_equal(got, (ref.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1))


@pytest.mark.parametrize("sizes", [(1024, 256), (178, 256), (218, 256), (256, 256), (37, 64), (300, 7), (5, 64), (1, 4), (1000, 33)])
def test_cabi_coefficients_match_oracle(sizes):
    """sivae_resample_coeffs is host-only arithmetic (IEEE doubles in Pillow's operation order): runs without a GPU"""
    L = importlib.import_module(PKG + ".lib")
    lib = L.load()
    n_in, n_out = sizes
    ksize, bounds, kk = IO.precompute_coeffs(n_in, n_out)
    ks = C.c_int(0)
    b = np.zeros((n_out, 2), dtype=np.int32)
    k = np.zeros((n_out, ksize), dtype=np.int32)
    rc = lib.sivae_resample_coeffs(n_in, n_out, C.byref(ks), b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), k.size)
    assert rc == 0 and ks.value == ksize
    assert np.array_equal(b, bounds) and np.array_equal(k, kk)
    rc = lib.sivae_resample_coeffs(n_in, n_out, C.byref(ks), b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), 1)
    assert rc != 0          # capacity too small: error, nothing written past the buffer


def test_dataset_decodes_and_draws_the_mirror_coin_like_the_reference(tmp_path):
    """host half of gpu_dataset.ImageDatasetFromFile: decoded pixels == the file's, and the mirror coin consumes Python's
    `random` stream exactly like dataset.py:26 (flags recorded from the unmodified reference with the same seed)"""
    import random

    import torch
    from PIL import Image
    M = importlib.import_module(PKG + ".gpu_dataset")
    z = np.load(GOLD)
    src, mirror = z["hq_like/src"], z["hq_like/mirror"]
    names = []
    for i, a in enumerate(src):
        names.append("im_%d.png" % i)
        Image.fromarray(a, "RGB").save(tmp_path / names[-1])
    ds = M.ImageDatasetFromFile(names, str(tmp_path), input_height=None, crop_height=None, output_height=24, is_mirror=True)
    random.seed(1234)
    items = [ds[i] for i in range(len(ds))]
    assert [f for _, f in items] == [int(m) for m in mirror]
    for (img, _), a in zip(items, src):
        assert img.dtype == torch.uint8 and np.array_equal(img.numpy(), a)
    groups = M.collate_decoded(items)
    assert len(groups) == 1 and groups[0][1].shape == src.shape and groups[0][2].tolist() == [int(m) for m in mirror]
    with pytest.raises(NotImplementedError):
        M.ImageDatasetFromFile(names, str(tmp_path), input_height=128)          # two-stage resize: not on this path
    with pytest.raises(RuntimeError):
        M.ImageBatcher(24, 24, "cpu")                                           # no CPU fallback
