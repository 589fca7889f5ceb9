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


GOLD_CROP = os.path.join(ROOT, "tests", "golden", "image_pipeline_crop.npz")


def crop_cases():
    """(name, src [n,H,W,3], mirror [n], margins [n,4], out_u8 [n,C,oh,ow], args) recorded from the unmodified reference
    ImageDatasetFromFile with input_height / crop_* / is_random_crop / is_gray set (oracle/make_image_golden.py --crop)"""
    z = np.load(GOLD_CROP)
    out = []
    for n in sorted({k.split("/")[0] for k in z.files}):
        ih, iw, ch, cw, oh, ow, rnd, gray = (int(v) for v in z[n + "/args"])
        out.append((n, z[n + "/src"], z[n + "/mirror"], z[n + "/margins"], z[n + "/out_u8"],
                    dict(input_hw=(ih, iw) if ih > 0 else None, crop=(ch, cw) if ch > 0 else None, out_hw=(oh, ow),
                         is_random_crop=bool(rnd), is_gray=bool(gray))))
    return out


def _gray(a):
    from PIL import Image
    return np.asarray(Image.fromarray(a, "RGB").convert("L"))[..., None]


def test_oracle_load_image_branches_match_reference_golden():
    """two-stage resize (dataset.py:29-30), random / centre ImageOps.crop (:32-44), grey conversion: the oracle's load_image_u8
    on the recorded source pixels, mirror coins and crop margins == the tensors the unmodified reference dataset returned"""
    cases = crop_cases()
    assert len(cases) == 6 and any(c[5]["is_random_crop"] for c in cases) and any(c[5]["is_gray"] for c in cases)
    for name, src, mirror, margins, out_u8, a in cases:
        for i in range(len(src)):
            img = _gray(src[i]) if a["is_gray"] else src[i]
            got = IO.load_image_u8(img, bool(mirror[i]), a["out_hw"][0], a["out_hw"][1], input_hw=a["input_hw"],
                                   margins=margins[i] if a["crop"] else None)
            assert np.array_equal(got.transpose(2, 0, 1), out_u8[i]), (name, i)
            assert np.array_equal(IO.to_tensor(got), out_u8[i].astype(np.float32) / np.float32(255.0))
    # the centre crop keeps w - 2 round((w - crop) / 2) columns: 61 rows, crop 48 -> round(6.5) = 6 (banker's) -> 49 rows
    n, _, _, margins, _, a = [c for c in cases if c[0] == "center_crop_odd"][0]
    assert tuple(margins[0]) == (8, 6, 8, 6)


def test_oracle_crop_margins_replay_the_reference_draws():
    """IO.crop_margins consumes Python's `random` exactly like dataset.py:35-38 (after the mirror coin of :26)"""
    import random
    for name, src, mirror, margins, _, a in crop_cases():
        if not a["crop"]:
            continue
        random.seed(4321)
        h, w = src.shape[1:3]
        if a["input_hw"]:
            h, w = a["input_hw"]
        for i in range(len(src)):
            flag = 1 if random.randint(0, 1) == 0 else 0
            m = IO.crop_margins(w, h, a["crop"][1], a["crop"][0], a["is_random_crop"], random)
            assert flag == int(mirror[i]) and tuple(m) == tuple(int(v) for v in margins[i]), (name, i)


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
    with pytest.raises(RuntimeError):
        M.ImageBatcher(24, 24, "cpu")                                           # no CPU fallback


def test_dataset_draws_crop_offsets_like_the_reference(tmp_path):
    """host half of the crop / two-stage branches: per item the mirror coin and then cx1, cy1 of a random crop come out of
    Python's `random` stream exactly as in the unmodified reference run that recorded the golden (dataset.py:26, :35-38)"""
    import random

    from PIL import Image
    M = importlib.import_module(PKG + ".gpu_dataset")
    for name, src, mirror, margins, _, a in crop_cases():
        names = []
        for i, im in enumerate(src):
            names.append("%s_%d.png" % (name, i))
            Image.fromarray(im, "RGB").save(tmp_path / names[-1])
        kw = dict(input_height=a["input_hw"][0] if a["input_hw"] else None, input_width=a["input_hw"][1] if a["input_hw"] else None,
                  output_height=a["out_hw"][0], output_width=a["out_hw"][1], crop_height=a["crop"][0] if a["crop"] else None,
                  crop_width=a["crop"][1] if a["crop"] else None, is_random_crop=a["is_random_crop"], is_mirror=True,
                  is_gray=a["is_gray"])
        ds = M.ImageDatasetFromFile(names, str(tmp_path), **kw)
        random.seed(4321)
        items = [ds[i] for i in range(len(ds))]
        assert [it[1] for it in items] == [int(m) for m in mirror], name
        if a["crop"]:
            assert [tuple(it[2]) for it in items] == [tuple(int(v) for v in m) for m in margins], name
        else:
            assert all(len(it) == 2 for it in items)
        assert items[0][0].shape[2] == (1 if a["is_gray"] else 3)
        groups = M.collate_decoded(items)
        assert sum(g[0].numel() for g in groups) == len(items)


def test_gpu_jpeg_opt_in_keeps_files_compressed_and_draws_unchanged(tmp_path, monkeypatch):
    """SIVAE_GPU_JPEG=1 (host half): JPEG files stay compressed until the batch reaches the GPU (EncodedJpeg with the decoded
    geometry from the header), other files are decoded as before, the `random` draws do not change, groups split by kind;
    the C ABI reports nvJPEG as unavailable without a GPU instead of crashing"""
    import ctypes as C
    import pickle
    import random

    import torch
    from PIL import Image
    M = importlib.import_module(PKG + ".gpu_dataset")
    L = importlib.import_module(PKG + ".lib")
    rng = np.random.default_rng(5)
    names = []
    for i in range(3):
        names.append("j%d.jpg" % i)
        Image.fromarray(rng.integers(0, 256, (40, 56, 3), dtype=np.uint8), "RGB").save(tmp_path / names[-1], quality=90)
    names.append("p.png")
    Image.fromarray(rng.integers(0, 256, (40, 56, 3), dtype=np.uint8), "RGB").save(tmp_path / "p.png")
    kw = dict(input_height=None, crop_height=32, crop_width=30, is_random_crop=True, output_height=16, is_mirror=True)
    ds = M.ImageDatasetFromFile(names, str(tmp_path), **kw)
    random.seed(9)
    plain = [ds[i] for i in range(4)]
    monkeypatch.setenv("SIVAE_GPU_JPEG", "1")
    random.seed(9)
    items = [ds[i] for i in range(4)]
    assert [type(it[0]).__name__ for it in items] == ["EncodedJpeg"] * 3 + ["Tensor"]
    assert all(tuple(a[0].shape) == tuple(b[0].shape) and a[1:] == b[1:] for a, b in zip(items, plain))
    raw = open(tmp_path / "j0.jpg", "rb").read()
    assert bytes(items[0][0].data.numpy().tobytes()) == raw
    groups = M.collate_decoded(items)
    assert sorted(type(g[1]).__name__ for g in groups) == ["EncodedBatch", "Tensor"]
    pickle.loads(pickle.dumps(groups))                       # DataLoader workers ship the groups between processes
    with pytest.raises(RuntimeError):
        M.decode_jpeg_batch([items[0][0].data], items[0][0].shape, "cpu")
    if not torch.cuda.is_available():
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        assert L.load().sivae_jpeg_info(buf, len(raw), C.byref(h), C.byref(w), C.byref(c)) == -9


def test_load_spec_orchestration_on_an_oracle_backed_batcher(monkeypatch):
    """LoadSpec.assemble (which launches, with which windows / flags, in which order) checked WITHOUT a GPU: ImageBatcher is
    replaced by a stand-in that computes each launch with the numpy oracle, and the assembled batches must equal the tensors the
    unmodified reference dataset returned (image_pipeline_crop.npz) -- in particular a crop of a MIRRORED image (dataset.py:26-27
    precede :32-44) must come out as the mirrored window whose left border is the crop's right border"""
    import torch
    M = importlib.import_module(PKG + ".gpu_dataset")
    calls = []

    class FakeBatcher:
        def __init__(self, out_h, out_w=None, device="cuda:0"):
            self.out_h, self.out_w, self.device = int(out_h), int(out_w if out_w is not None else out_h), device

        def __call__(self, images_u8, mirror=None, out=None, window=None, origins=None, as_u8=False):
            imgs = images_u8.numpy()
            B = imgs.shape[0]
            flags = np.zeros(B, np.uint8) if mirror is None else np.asarray(torch.as_tensor(mirror)).astype(np.uint8)
            calls.append((tuple(imgs.shape[1:3]), (self.out_h, self.out_w), window, bool(flags.any()), as_u8))
            res = []
            for i in range(B):
                a = imgs[i]
                if window is not None:
                    x, y = (int(v) for v in torch.as_tensor(origins)[i])
                    a = a[y:y + window[0], x:x + window[1]]
                res.append(IO.load_image_u8(np.ascontiguousarray(a), bool(flags[i]), self.out_h, self.out_w))
            res = np.stack(res)
            return torch.from_numpy(res) if as_u8 else torch.from_numpy(np.stack([IO.to_tensor(r) for r in res]))

    monkeypatch.setattr(M, "ImageBatcher", FakeBatcher)
    for name, src, mirror, margins, out_u8, a in crop_cases():
        imgs = np.stack([_gray(s_) for s_ in src]) if a["is_gray"] else src
        spec = M.LoadSpec(a["input_hw"][0] if a["input_hw"] else None, a["input_hw"][1] if a["input_hw"] else None,
                          a["out_hw"][0], a["out_hw"][1], a["crop"][0] if a["crop"] else None, a["crop"][1] if a["crop"] else None,
                          a["is_random_crop"], True)
        calls.clear()
        got = spec.assemble({"device": "cpu"}, torch.from_numpy(imgs), torch.from_numpy(mirror),
                            torch.from_numpy(margins.astype(np.int64)) if a["crop"] else None)
        assert np.array_equal(got.numpy(), out_u8.astype(np.float32) / np.float32(255.0)), name
        assert len(calls) == (2 if a["input_hw"] else 1), (name, calls)
        if a["input_hw"]:
            assert calls[0][4] and not calls[1][3]            # stage 1 writes the 8-bit image and carries the mirror; stage 2 none
        assert (calls[-1][2] is not None) == bool(a["crop"])


def test_monsters_dataset_equals_the_reference_class(tmp_path, monkeypatch):
    """`monsters128` (dataset.py:96-149): the drop-in DigitalMonstersDataset against the UNMODIFIED reference class, item by item
    under the same torch seed (RandomAffine / ColorJitter / RandomHorizontalFlip draw from torch's generator).  The reference
    class passes `fillcolor=` to RandomAffine, which torchvision renamed to `fill=` in 0.13 -- the only shim applied to it."""
    import sys
    import warnings

    import torch
    import torchvision.transforms as T
    from PIL import Image
    if not os.path.isdir("/root/reference/soft_intro_vae"):
        pytest.skip("needs the reference tree (build container only)")
    rng = np.random.default_rng(3)
    for sub, n in (("pokemon", 3), (os.path.join("digimon", "200"), 2), ("nexomon", 2)):
        os.makedirs(tmp_path / sub)
        for i in range(n):
            a = rng.integers(0, 256, (40 + 3 * i, 36 + 5 * i, 3), dtype=np.uint8)
            Image.fromarray(a, "RGB").save(tmp_path / sub / ("m%d.png" % i))
    (tmp_path / "pokemon" / "notes.txt").write_text("not an image")

    class Affine(T.RandomAffine):
        def __init__(self, degrees, translate=None, scale=None, shear=None, fillcolor=0, **kw):
            super().__init__(degrees, translate=translate, scale=scale, shear=shear, fill=fillcolor, **kw)

    warnings.simplefilter("ignore", SyntaxWarning)
    sys.path.insert(0, "/root/reference/soft_intro_vae")
    sys.modules.pop("dataset", None)
    try:
        import dataset as ref_dataset
    finally:
        sys.path.pop(0)
    M = importlib.import_module(PKG + ".gpu_dataset")
    root = str(tmp_path) + os.sep
    ours = M.DigitalMonstersDataset(root_path=root, output_height=32)           # built on the unpatched torchvision
    monkeypatch.setattr(ref_dataset.transforms, "RandomAffine", Affine)        # (ref_dataset.transforms IS torchvision.transforms)
    ref = ref_dataset.DigitalMonstersDataset(root_path=root, output_height=32)
    assert len(ref) == len(ours) == 7 and ref.image_filenames == ours.image_filenames
    for i in range(len(ref)):
        torch.manual_seed(100 + i)
        want = ref[i]
        torch.manual_seed(100 + i)
        got = ours[i]
        assert got.dtype == torch.float32 and got.shape == (3, 32, 32) and torch.equal(got, want), i
    sys.modules.pop("dataset", None)
