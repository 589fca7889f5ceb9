"""-m gpu: the image batch-assembly kernel (csrc/image.cu) through the C ABI against the oracle (oracle/image_oracle.py,
pinned to Pillow and to the reference's dataset.py in tests/test_image_oracle.py).  Integer / correctly-rounded work:
the bar is bit-exact."""
import importlib
import os
import random

import numpy as np
import pytest
import torch

from oracle import image_oracle as IO

pytestmark = pytest.mark.gpu
PKG = "soft-intro-vae-pytorch_b200"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "image_pipeline.npz")


def _mod():
    return importlib.import_module(PKG + ".gpu_dataset")


def test_golden_vectors_of_the_reference_loader():
    z = np.load(GOLD)
    for name in sorted({k.split("/")[0] for k in z.files}):
        src, mirror, out_u8, size = z[name + "/src"], z[name + "/mirror"], z[name + "/out_u8"], int(z[name + "/size"][0])
        got = _mod().ImageBatcher(size, size, "cuda:0")(torch.from_numpy(src), torch.from_numpy(mirror)).cpu().numpy()
        assert np.array_equal(got, out_u8.astype(np.float32) / np.float32(255.0)), name


@pytest.mark.parametrize("shape", [((218, 178), (128, 128)), ((64, 64), (32, 32)), ((100, 37), (13, 91)), ((33, 47), (47, 33)),
                                   ((40, 40), (40, 17)), ((17, 29), (64, 128)), ((300, 200), (9, 7)), ((5, 5), (64, 64)),
                                   ((1, 7), (4, 4)), ((31, 31), (31, 31)), ((257, 131), (96, 200))])
@pytest.mark.parametrize("ch", [3, 1])
def test_matches_oracle(shape, ch):
    """up- and down-scaling, identity axes, ragged tiles, rows whose byte length is not a multiple of 4 (unaligned row
    starts in the word staging), mixed mirror flags, saturating content"""
    (h, w), (oh, ow) = shape
    B = 5
    rng = np.random.default_rng(h * 131 + w * 7 + ch)
    src = rng.integers(0, 256, (B, h, w, ch), dtype=np.uint8)
    src[:, : h // 2, : w // 2] = 255
    src[:, h // 2:, w // 2:] = 0
    mirror = np.array([0, 1, 1, 0, 1], dtype=np.uint8)
    got = _mod().ImageBatcher(oh, ow, "cuda:0")(torch.from_numpy(src), torch.from_numpy(mirror)).cpu().numpy()
    want = IO.batch(src, mirror, oh, ow)
    assert got.shape == want.shape and np.array_equal(got, want)
    # no mirror array at all == all zeros
    got0 = _mod().ImageBatcher(oh, ow, "cuda:0")(torch.from_numpy(src), None).cpu().numpy()
    assert np.array_equal(got0, IO.batch(src, np.zeros(B, np.uint8), oh, ow))


def test_full_size_properties():
    """BASELINE sizes (1024x1024 -> 256x256, batch 32): two images against the oracle, the rest through size-independent
    properties: mirror flag == mirrored source, constant images stay constant, identity geometry == source / 255."""
    rng = np.random.default_rng(3)
    B, S, T = 32, 1024, 256
    src = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    src[5] = 255
    src[6] = 0
    src[7] = 77
    mirror = (np.arange(B) % 2).astype(np.uint8)
    bt = _mod().ImageBatcher(T, T, "cuda:0")
    got = bt(torch.from_numpy(src), torch.from_numpy(mirror)).cpu().numpy()
    for i in (0, 1):
        assert np.array_equal(got[i], IO.load_image_tensor(src[i], T, T, bool(mirror[i])))
    flipped = np.ascontiguousarray(src[:, :, ::-1])
    got_f = bt(torch.from_numpy(flipped), torch.from_numpy(1 - mirror)).cpu().numpy()
    assert np.array_equal(got, got_f)
    assert (got[5] == 1.0).all() and (got[6] == 0.0).all() and (got[7] == np.float32(77) / np.float32(255)).all()
    ident = _mod().ImageBatcher(S, S, "cuda:0")(torch.from_numpy(src[:2]), None).cpu().numpy()
    assert np.array_equal(ident, (src[:2].astype(np.float32) / np.float32(255)).transpose(0, 3, 1, 2))


GOLD_CROP = os.path.join(ROOT, "tests", "golden", "image_pipeline_crop.npz")


def test_load_image_branches_vs_reference_golden(tmp_path):
    """two-stage resize (dataset.py:29-30), random / centre crops (:32-44), grey images: files -> ImageDatasetFromFile with those
    arguments -> GpuImageLoader == the tensors the UNMODIFIED reference dataset returned for the same `random` seed
    (tests/golden/image_pipeline_crop.npz, oracle/make_image_golden.py --crop), bit for bit; load_image() returns the same
    PIL image the reference's does"""
    from PIL import Image
    M = _mod()
    z = np.load(GOLD_CROP)
    for name in sorted({k.split("/")[0] for k in z.files}):
        src, out_u8 = z[name + "/src"], z[name + "/out_u8"]
        ih, iw, ch, cw, oh, ow, rnd, gray = (int(v) for v in z[name + "/args"])
        names = []
        for i, a in enumerate(src):
            names.append("%s_%d.png" % (name, i))
            Image.fromarray(a, "RGB").save(tmp_path / names[-1])
        kw = dict(input_height=ih if ih > 0 else None, input_width=iw if iw > 0 else None, output_height=oh, output_width=ow,
                  crop_height=ch if ch > 0 else None, crop_width=cw if cw > 0 else None, is_random_crop=bool(rnd), is_mirror=True,
                  is_gray=bool(gray))
        ds = M.ImageDatasetFromFile(names, str(tmp_path), **kw)
        loader = M.GpuImageLoader(torch.utils.data.DataLoader(ds, batch_size=len(names), shuffle=False, num_workers=0,
                                                              collate_fn=M.collate_decoded), "cuda:0")
        random.seed(4321)
        (got,) = list(loader)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), out_u8.astype(np.float32) / np.float32(255.0)), name
        random.seed(4321)
        im = M.load_image(str(tmp_path / names[0]), **kw)
        a0 = np.asarray(im)
        a0 = a0[..., None] if a0.ndim == 2 else a0
        assert np.array_equal(a0.transpose(2, 0, 1), out_u8[0]), name


@pytest.mark.parametrize("ch", [3, 1])
def test_window_and_u8_output_match_oracle(ch):
    """the two generalisations of the kernel behind those branches, directly: per-image crop windows (corners, unaligned
    origins, windows as large as the image) with and without mirroring, and the 8-bit NHWC output"""
    rng = np.random.default_rng(77 + ch)
    B, H, W = 6, 53, 71
    src = rng.integers(0, 256, (B, H, W, ch), dtype=np.uint8)
    src[:, :20, :30] = 255
    src[:, 30:, 40:] = 0
    wh, ww, oh, ow = 31, 45, 24, 56
    origins = np.array([[0, 0], [W - ww, H - wh], [1, 2], [13, 7], [W - ww, 0], [3, H - wh]], dtype=np.int32)
    mirror = np.array([0, 1, 1, 0, 1, 0], dtype=np.uint8)
    bt = _mod().ImageBatcher(oh, ow, "cuda:0")
    want_u8 = np.stack([IO.resize_bicubic_u8(np.ascontiguousarray(
        (src[i, y:y + wh, x:x + ww][:, ::-1] if mirror[i] else src[i, y:y + wh, x:x + ww])), oh, ow)
        for i, (x, y) in enumerate(origins)])
    got = bt(torch.from_numpy(src), torch.from_numpy(mirror), window=(wh, ww), origins=torch.from_numpy(origins)).cpu().numpy()
    assert np.array_equal(got, np.stack([IO.to_tensor(w) for w in want_u8]))
    got8 = bt(torch.from_numpy(src), torch.from_numpy(mirror), window=(wh, ww), origins=torch.from_numpy(origins), as_u8=True)
    assert got8.dtype == torch.uint8 and tuple(got8.shape) == (B, oh, ow, ch) and np.array_equal(got8.cpu().numpy(), want_u8)
    # whole image as the window == the plain call; u8 output of the plain geometry
    full = bt(torch.from_numpy(src), torch.from_numpy(mirror), window=(H, W), origins=np.zeros((B, 2), np.int32)).cpu().numpy()
    assert np.array_equal(full, IO.batch(src, mirror, oh, ow))
    full8 = bt(torch.from_numpy(src), torch.from_numpy(mirror), as_u8=True).cpu().numpy()
    assert np.array_equal(full8, np.stack([IO.load_image_u8(src[i], bool(mirror[i]), oh, ow) for i in range(B)]))
    with pytest.raises(ValueError, match="outside"):
        bt(torch.from_numpy(src), None, window=(wh, ww), origins=np.array([[W - ww + 1, 0]] * B, np.int32))


def test_gpu_jpeg_decode_opt_in(tmp_path, monkeypatch):
    """SIVAE_GPU_JPEG=1: JPEG files decoded by nvJPEG straight into the device batch the resize kernel reads.
    nvJPEG is a library and NOT bit-exact with Pillow's libjpeg-turbo (IDCT / colour rounding, 4:2:0 chroma up-sampling), which
    is why the path is opt-in: held here to a few grey levels on average against Pillow's decode of the same files; the
    measured differences go to gpurun_out/jpeg_report.json"""
    import ctypes as C
    import json

    from PIL import Image
    M = _mod()
    L = importlib.import_module(PKG + ".lib")
    yy, xx = np.mgrid[0:96, 0:80].astype(np.float64)
    smooth = [np.clip(np.stack([127 + 110 * np.sin(xx / (6.0 + c + i) + yy / 9.0) * np.cos(yy / (7.0 + 2 * c)) for c in range(3)], -1),
                      0, 255).astype(np.uint8) for i in range(4)]
    files = {"s444": [], "s420": []}
    for i, a in enumerate(smooth):
        for tag, sub in (("s444", 0), ("s420", 2)):
            fn = "%s_%d.jpg" % (tag, i)
            Image.fromarray(a, "RGB").save(tmp_path / fn, quality=92, subsampling=sub)
            files[tag].append(fn)
    raw0 = open(tmp_path / files["s444"][0], "rb").read()
    h, w, c = C.c_int(), C.c_int(), C.c_int()
    rc = L.load().sivae_jpeg_info((C.c_ubyte * len(raw0)).from_buffer_copy(raw0), len(raw0), C.byref(h), C.byref(w), C.byref(c))
    if rc == -9:
        pytest.skip("nvJPEG not available on this box: %s" % L.load().sivae_last_error().decode())
    assert rc == 0 and (h.value, w.value, c.value) == (96, 80, 3)
    report = {}
    for tag, names in files.items():
        datas = [torch.from_numpy(np.fromfile(tmp_path / n, dtype=np.uint8)) for n in names]
        got = M.decode_jpeg_batch(datas, (96, 80, 3), "cuda:0")
        assert got.dtype == torch.uint8 and tuple(got.shape) == (4, 96, 80, 3) and got.is_cuda
        ref = np.stack([np.asarray(Image.open(tmp_path / n).convert("RGB")) for n in names])
        d = np.abs(got.cpu().numpy().astype(np.int32) - ref.astype(np.int32))
        report[tag] = dict(mean_abs=float(d.mean()), max_abs=int(d.max()), frac_equal=float((d == 0).mean()))
        assert d.mean() < (3.0 if tag == "s444" else 10.0), (tag, report[tag])      # measured: see profiles/ jpeg report
    with pytest.raises(RuntimeError, match="expected"):
        M.decode_jpeg_batch([torch.from_numpy(np.fromfile(tmp_path / files["s444"][0], dtype=np.uint8))], (90, 80, 3), "cuda:0")
    # the loader end to end: compressed files -> nvJPEG -> mirror / resize / ToTensor kernel, against the default (Pillow) decode
    kw = dict(input_height=None, crop_height=None, output_height=32, is_mirror=True)
    ds = M.ImageDatasetFromFile(files["s444"], str(tmp_path), **kw)
    mk = lambda: M.GpuImageLoader(torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, num_workers=0,
                                                              collate_fn=M.collate_decoded), "cuda:0")
    random.seed(2)
    (base,) = list(mk())
    monkeypatch.setenv("SIVAE_GPU_JPEG", "1")
    random.seed(2)
    (fast,) = list(mk())
    assert fast.shape == base.shape == (4, 3, 32, 32)
    report["loader_mean_abs"] = float((fast - base).abs().mean())
    assert report["loader_mean_abs"] < 3.0 / 255
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "jpeg_report.json"), "w") as f:
        json.dump(report, f)


def test_downscale_beyond_staging_capacity_fails_loudly():
    src = torch.zeros(1, 4096, 8, 3, dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="down-scaling factor too large"):
        _mod().ImageBatcher(2, 8, "cuda:0")(src, None)
    with pytest.raises(RuntimeError):
        _mod().ImageBatcher(8, 8, "cuda:0")(torch.zeros(1, 8, 8, 2, dtype=torch.uint8), None)      # 2 channels


def test_loader_end_to_end(tmp_path):
    """files -> ImageDatasetFromFile (decode + the reference's mirror coin) -> DataLoader -> GpuImageLoader == the tensors
    the reference's dataset returns for the same seed (golden), including a batch with two source geometries"""
    from PIL import Image
    M = _mod()
    z = np.load(GOLD)
    src, mirror, out_u8 = z["celeba_like/src"], z["celeba_like/mirror"], z["celeba_like/out_u8"]
    size = int(z["celeba_like/size"][0])
    names = []
    for i, a in enumerate(src):
        names.append("img_%d.png" % i)
        Image.fromarray(a, "RGB").save(tmp_path / names[-1])
    ds = M.ImageDatasetFromFile(names, str(tmp_path), input_height=None, crop_height=None, output_height=size, is_mirror=True)
    loader = M.GpuImageLoader(torch.utils.data.DataLoader(ds, batch_size=len(names), shuffle=False, num_workers=0,
                                                          collate_fn=M.collate_decoded), "cuda:0")
    random.seed(1234)                                   # the seed the golden was recorded with (oracle/make_image_golden.py)
    batches = list(loader)
    assert len(batches) == 1 and batches[0].is_cuda
    assert np.array_equal(batches[0].cpu().numpy(), out_u8.astype(np.float32) / np.float32(255.0))
    # mixed geometries in one batch
    odd = np.ascontiguousarray(src[0][:50, :41])
    Image.fromarray(odd, "RGB").save(tmp_path / "odd.png")
    ds2 = M.ImageDatasetFromFile(["img_0.png", "odd.png", "img_1.png"], str(tmp_path), input_height=None, crop_height=None,
                                 output_height=size, is_mirror=False)
    (b2,) = list(M.GpuImageLoader(torch.utils.data.DataLoader(ds2, batch_size=3, collate_fn=M.collate_decoded), "cuda:0"))
    want = np.stack([IO.load_image_tensor(a, size, size, False) for a in (src[0], odd, src[1])])
    assert np.array_equal(b2.cpu().numpy(), want)


def test_device_prefetcher_keeps_values_and_order(tmp_path):
    """DevicePrefetcher (batch i+1 moves / is assembled on a side stream under step i): same batches, same order, on the
    device -- for a plain float loader, a labelled one, and the GPU image loader"""
    from PIL import Image
    T = importlib.import_module(PKG + ".train_soft_intro_vae")
    g = torch.Generator().manual_seed(2)
    data = torch.rand(37, 3, 8, 8, generator=g)
    plain = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(data), batch_size=8, shuffle=False, pin_memory=True)
    got = [b[0] for b in T.DevicePrefetcher(plain, "cuda:0")]
    assert all(t.is_cuda for t in got) and len(got) == 5
    # consume with work on the main stream in between, as the trainer does
    acc = torch.zeros(3, 8, 8, device="cuda:0")
    for b in T.DevicePrefetcher(plain, "cuda:0"):
        acc += b[0].sum(0) * 1.0
        torch.cuda.current_stream().synchronize()
    assert torch.equal(torch.cat(got).cpu(), data)
    assert torch.allclose(acc.cpu(), data.sum(0), rtol=1e-5, atol=1e-5)
    M = _mod()
    z = np.load(GOLD)
    src, out_u8, size = z["celeba_like/src"], z["celeba_like/out_u8"], int(z["celeba_like/size"][0])
    names = []
    for i, a in enumerate(src):
        names.append("p_%d.png" % i)
        Image.fromarray(a, "RGB").save(tmp_path / names[-1])
    ds = M.ImageDatasetFromFile(names, str(tmp_path), input_height=None, crop_height=None, output_height=size, is_mirror=False)
    loader = M.GpuImageLoader(torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, collate_fn=M.collate_decoded), "cuda:0")
    want = np.stack([IO.load_image_tensor(a, size, size, False) for a in src])
    got = torch.cat(list(T.DevicePrefetcher(loader, "cuda:0"))).cpu().numpy()
    assert np.array_equal(got, want)
