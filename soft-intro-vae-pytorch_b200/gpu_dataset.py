"""
GPU batch assembly for the file datasets of ``soft_intro_vae/dataset.py``.

The reference's ``ImageDatasetFromFile.__getitem__`` (dataset.py:74-78) does, per image and on one CPU core (``main.py:47``
passes ``num_workers=0``): ``Image.open`` -> ``convert('RGB')`` -> random ``ImageOps.mirror`` -> ``resize((S, S),
Image.BICUBIC)`` -> ``ToTensor``.  With the step at several hundred images/s per GPU that loader is the bottleneck.
Here the dataset only *decodes* (``Image.open`` / ``convert``: entropy decoding is not on this path) and draws the mirror
coin exactly like dataset.py:26 (``random.randint(0, 1) is 0``, same consumption of Python's ``random`` stream); mirror,
bicubic resize and ToTensor run for the whole batch in one CUDA kernel (``csrc/image.cu`` through
``sivae_image_batch_u8``), bit-exact with Pillow's fixed-point resampler and therefore with the tensor the reference's
dataset returns (``tests/test_gpu_image.py``; oracle pinned in ``tests/test_image_oracle.py``).

Same constructor signature as the reference class, and all of ``load_image`` (dataset.py:12-47): what the image configs pass
(train_soft_intro_vae.py:388-392, 400-404, 415-417: ``input_height=None, crop_height=None``) is ONE kernel launch per batch;
``input_height`` (two-stage resize, :29-30) adds a first launch that writes the 8-bit intermediate image, ``crop_height``
(random / centre ``ImageOps.crop``, :32-44) makes the last resize read a per-image window -- the random crop draws cx1, cy1
from Python's ``random`` stream right after the mirror coin, as the reference does.  There is no CPU fallback on this path.
"""
import ctypes as C
import os
import random

import numpy as np
import torch
import torch.utils.data as data

from . import lib as L


class EncodedJpeg:
    """a JPEG file kept COMPRESSED by the dataset (SIVAE_GPU_JPEG=1): `data` = its bytes (1-D uint8 tensor), `shape` = the
    (H, W, C) its decoded pixels will have.  Decoded per batch on the GPU by nvJPEG (decode_jpeg_batch)."""
    __slots__ = ("data", "shape")

    def __init__(self, data, shape):
        self.data, self.shape = data, tuple(int(v) for v in shape)


class EncodedBatch:
    """same-sized EncodedJpeg items of one collated group"""
    __slots__ = ("datas", "shape")

    def __init__(self, datas, shape):
        self.datas, self.shape = list(datas), tuple(shape)


def gpu_jpeg_enabled():
    """opt-in: JPEG files are decoded on the GPU by nvJPEG instead of by Pillow on the host.  NOT bit-exact with the reference's
    loader (libjpeg-turbo and nvJPEG round the IDCT / colour conversion differently and up-sample 4:2:0 chroma differently; a few
    grey levels) -- hence off by default."""
    return os.environ.get("SIVAE_GPU_JPEG", "0") == "1"


def encoded_jpeg(file_path, is_gray=False):
    """-> EncodedJpeg if `file_path` is an RGB JPEG wanted as RGB (the `Image.open` of dataset.py:20-24 without a mode
    conversion), else None: grey / CMYK files and is_gray data sets stay with Pillow.  Reads only the header."""
    if is_gray or not file_path.lower().endswith((".jpg", ".jpeg")):
        return None
    from PIL import Image
    with Image.open(file_path) as im:
        if im.format != "JPEG" or im.mode != "RGB":
            return None
        w, h = im.size
    return EncodedJpeg(torch.from_numpy(np.fromfile(file_path, dtype=np.uint8)), (h, w, 3))


def decode_jpeg_batch(datas, shape, device):
    """compressed JPEG files (1-D uint8 host tensors) of one size -> uint8 [B,H,W,C] on `device` (sivae_jpeg_decode_batch)"""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("decode_jpeg_batch runs on CUDA devices only; got %s" % (device,))
    h, w, ch = shape
    if ch != 3:
        raise ValueError("decode_jpeg_batch decodes RGB JPEG files to [B,H,W,3]")
    B = len(datas)
    datas = [d.contiguous() for d in datas]
    out = torch.empty(B, h, w, ch, dtype=torch.uint8, device=device)
    ptrs = (C.c_void_p * B)(*[d.data_ptr() for d in datas])
    lens = (C.c_longlong * B)(*[d.numel() for d in datas])
    with torch.cuda.device(device):
        L.check(L.load().sivae_jpeg_decode_batch(ptrs, lens, B, h, w, L.ptr(out),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sivae_jpeg_decode_batch")
    return out


def decode_image(file_path, is_gray=False):
    """dataset.py:20-24: Image.open + mode conversion -> uint8 [H,W,C] (C = 3, or 1 for is_gray)"""
    from PIL import Image
    img = Image.open(file_path)
    if is_gray is False and img.mode != 'RGB':
        img = img.convert('RGB')
    if is_gray and img.mode != 'L':
        img = img.convert('L')
    a = np.array(img)              # owned, writable copy of the decoded pixels
    if a.ndim == 2:
        a = a[:, :, None]
    return np.ascontiguousarray(a)


class ImageBatcher:
    """mirror + Image.BICUBIC resize to (out_h, out_w) + ToTensor of a batch of decoded images on `device`.
    Coefficient tables ("plans", include/sivae.h) are cached per source geometry in torch-owned device memory."""

    def __init__(self, out_h, out_w=None, device="cuda:0"):
        self.out_h, self.out_w = int(out_h), int(out_w if out_w is not None else out_h)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ImageBatcher runs on CUDA devices only (no CPU fallback); got %s" % (device,))
        self._plans = {}

    def _plan(self, in_h, in_w):
        key = (in_h, in_w)
        p = self._plans.get(key)
        if p is None:
            lib = L.load()
            nbytes = lib.sivae_image_plan_bytes(in_h, in_w, self.out_h, self.out_w)
            p = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                L.check(lib.sivae_image_plan_init(in_h, in_w, self.out_h, self.out_w, L.ptr(p), nbytes,
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sivae_image_plan_init")
            self._plans[key] = p
        return p

    def __call__(self, images_u8, mirror=None, out=None, window=None, origins=None, as_u8=False):
        """images_u8: uint8 [B,H,W,C] (host or device); mirror: [B] flags or None -> float32 [B,C,out_h,out_w] on device.
        window = (win_h, win_w) + origins = int [B,2] (x, y): resize only that window of every image (ImageOps.crop,
        dataset.py:32-44; the mirror flag then mirrors the WINDOW).  as_u8: return the resized 8-bit image itself, uint8
        [B,out_h,out_w,C] (first stage of the two-stage resize, :29-30) instead of its ToTensor."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
            raise ValueError("expected a uint8 [B,H,W,C] batch of decoded images")
        B, H, W, ch = images_u8.shape
        src = images_u8.to(self.device, non_blocking=True).contiguous()
        flags = None
        if mirror is not None:
            flags = torch.as_tensor(mirror).to(dtype=torch.uint8).to(self.device, non_blocking=True).contiguous()
        if out is None:
            out = (torch.empty(B, self.out_h, self.out_w, ch, dtype=torch.uint8, device=self.device) if as_u8 else
                   torch.empty(B, ch, self.out_h, self.out_w, dtype=torch.float32, device=self.device))
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if window is None and not as_u8:
            plan = self._plan(H, W)
            with torch.cuda.device(self.device):
                L.check(L.load().sivae_image_batch_u8(L.ptr(src), L.ptr(flags), B, H, W, ch, self.out_h, self.out_w, L.ptr(plan),
                                                      L.ptr(out), stream), "sivae_image_batch_u8")
            return out
        wh, ww = (H, W) if window is None else (int(window[0]), int(window[1]))
        xy = None
        if window is not None:
            o = torch.as_tensor(origins, dtype=torch.int32).reshape(B, 2)
            if int(o.min()) < 0 or int(o[:, 0].max()) + ww > W or int(o[:, 1].max()) + wh > H:
                raise ValueError("crop window outside the image")
            xy = o.to(self.device, non_blocking=True).contiguous()
        plan = self._plan(wh, ww)
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_image_batch_u8_ex(L.ptr(src), L.ptr(flags), L.ptr(xy), B, H, W, ch, wh, ww, self.out_h, self.out_w,
                                                     L.ptr(plan), None if as_u8 else L.ptr(out), L.ptr(out) if as_u8 else None, stream),
                    "sivae_image_batch_u8_ex")
        return out


def _crop_margins(w, h, crop_w, crop_h, is_random_crop):
    """dataset.py:32-43: borders (left, top, right, bottom) that ImageOps.crop removes; a random crop draws cx1 then cy1"""
    if is_random_crop:
        cx1 = random.randint(0, w - crop_w)
        cx2 = w - crop_w - cx1
        cy1 = random.randint(0, h - crop_h)
        cy2 = h - crop_h - cy1
    else:
        cx2 = cx1 = int(round((w - crop_w) / 2.))
        cy2 = cy1 = int(round((h - crop_h) / 2.))
    return cx1, cy1, cx2, cy2


class LoadSpec:
    """the geometry arguments of load_image (dataset.py:12-19 incl. the width defaults) + the per-image random draws"""

    def __init__(self, input_height, input_width, output_height, output_width, crop_height, crop_width, is_random_crop, is_mirror):
        self.input_height = input_height
        self.input_width = input_width if input_width is not None else input_height
        self.output_height = output_height
        self.output_width = output_width if output_width is not None else output_height
        self.crop_height = crop_height
        self.crop_width = crop_width if crop_width is not None else crop_height
        self.is_random_crop, self.is_mirror = is_random_crop, is_mirror

    def draw(self, h, w):
        """-> (mirror flag, (left, top, right, bottom) borders or None) for a decoded image of h x w pixels, consuming Python's
        `random` stream in load_image's order: mirror coin (:26), then the crop offsets (:35-38)"""
        mirror = 1 if (self.is_mirror and random.randint(0, 1) == 0) else 0
        margins = None
        if self.crop_height is not None:
            if self.input_height is not None:
                h, w = self.input_height, self.input_width
            margins = _crop_margins(w, h, self.crop_width, self.crop_height, self.is_random_crop)
        return mirror, margins

    def assemble(self, batchers, images_u8, mirror, margins, as_u8=False):
        """batch of same-sized decoded images -> float32 [B,C,oh,ow] on the device (uint8 [B,oh,ow,C] with as_u8).
        batchers: dict cache of ImageBatcher per output size."""
        def bt(h, w):
            key = (int(h), int(w))
            if key not in batchers:
                batchers[key] = ImageBatcher(h, w, batchers["device"])
            return batchers[key]
        x, flags = images_u8, mirror
        if self.input_height is not None:                            # :29-30 -- the mirror (:26-27) happens before it
            x = bt(self.input_height, self.input_width)(x, flags, as_u8=True)
            flags = None
        last = bt(self.output_height, self.output_width)
        if margins is None:
            return last(x, flags, as_u8=as_u8)
        m = torch.as_tensor(margins, dtype=torch.int64).reshape(-1, 4)
        H, W = x.shape[1], x.shape[2]
        wh, ww = H - int(m[0, 1]) - int(m[0, 3]), W - int(m[0, 0]) - int(m[0, 2])
        if not (bool((H - m[:, 1] - m[:, 3] == wh).all()) and bool((W - m[:, 0] - m[:, 2] == ww).all())):
            raise ValueError("images of one group must share the crop size")
        left = m[:, 0].clone()
        if flags is not None:                                        # crop of the MIRRORED image = mirrored window whose
            f = torch.as_tensor(flags).to(torch.bool).cpu()          # left border is the right border of the crop
            left[f] = m[:, 2][f]
        return last(x, flags, window=(wh, ww), origins=torch.stack([left, m[:, 1]], 1), as_u8=as_u8)


class ImageDatasetFromFile(data.Dataset):
    """Reference signature (dataset.py:50-53).  __getitem__ -> (uint8 [H,W,C] decoded pixels, mirror flag, crop borders): the
    rest of load_image + ToTensor happens per batch on the GPU (LoadSpec.assemble / GpuImageLoader)."""

    def __init__(self, image_list, root_path,
                 input_height=128, input_width=None, output_height=128, output_width=None,
                 crop_height=None, crop_width=None, is_random_crop=False, is_mirror=True, is_gray=False):
        super(ImageDatasetFromFile, self).__init__()
        self.image_filenames = image_list
        self.is_random_crop = is_random_crop
        self.is_mirror = is_mirror
        self.input_height = input_height
        self.input_width = input_width
        self.output_height = output_height
        self.output_width = output_width if output_width is not None else output_height
        self.root_path = root_path
        self.crop_height = crop_height
        self.crop_width = crop_width
        self.is_gray = is_gray
        self.spec = LoadSpec(input_height, input_width, output_height, output_width, crop_height, crop_width, is_random_crop,
                             is_mirror)

    def __getitem__(self, index):
        path = os.path.join(self.root_path, self.image_filenames[index])
        img = encoded_jpeg(path, self.is_gray) if gpu_jpeg_enabled() else None      # opt-in: stays compressed until the GPU
        if img is None:
            img = torch.from_numpy(decode_image(path, self.is_gray))
        mirror, margins = self.spec.draw(img.shape[0], img.shape[1])           # dataset.py:26, :35-38: same draws, same order
        if margins is None:
            return img, mirror
        return img, mirror, margins

    def __len__(self):
        return len(self.image_filenames)


def collate_decoded(samples):
    """-> list of (indices, uint8 [b,H,W,C], flags [b], borders [b,4] or None) groups, one per source geometry and crop size
    (usually exactly one)"""
    groups = {}
    for i, smp in enumerate(samples):
        img, flag = smp[0], smp[1]
        m = smp[2] if len(smp) > 2 else None
        key = (isinstance(img, EncodedJpeg),) + tuple(img.shape) + ((m[0] + m[2], m[1] + m[3]) if m is not None else ())
        groups.setdefault(key, []).append((i, img, flag, m))
    out = []
    for items in groups.values():
        idx = torch.tensor([i for i, _, _, _ in items], dtype=torch.long)
        borders = None if items[0][3] is None else torch.tensor([list(m) for _, _, _, m in items], dtype=torch.int64)
        if isinstance(items[0][1], EncodedJpeg):
            imgs = EncodedBatch([im.data for _, im, _, _ in items], items[0][1].shape)
        else:
            imgs = torch.stack([im for _, im, _, _ in items])
        out.append((idx, imgs, torch.tensor([f for _, _, f, _ in items], dtype=torch.uint8), borders))
    return out


class GpuImageLoader:
    """Iterates a DataLoader over an ImageDatasetFromFile (collate_fn=collate_decoded) and yields what the reference's
    loader yields -- float32 [B,C,S,S] batches in [0,1] -- already on `device`."""

    def __init__(self, loader, device):
        self.loader = loader
        ds = loader.dataset
        self.spec = ds.spec
        self.batchers = {"device": torch.device(device)}
        self.batcher = ImageBatcher(ds.output_height, ds.output_width, device)
        self.batchers[(int(ds.output_height), int(ds.output_width))] = self.batcher
        self.dataset, self.batch_size = ds, loader.batch_size

    def __len__(self):
        return len(self.loader)

    def _pixels(self, imgs):
        """decoded uint8 [b,H,W,C] pixels of a group (compressed groups: nvJPEG on the device)"""
        return decode_jpeg_batch(imgs.datas, imgs.shape, self.batcher.device) if isinstance(imgs, EncodedBatch) else imgs

    def __iter__(self):
        for groups in self.loader:
            if len(groups) == 1:
                yield self.spec.assemble(self.batchers, self._pixels(groups[0][1]), groups[0][2], groups[0][3])
                continue
            n = sum(g[0].numel() for g in groups)
            ch = groups[0][1].shape[-1]
            out = torch.empty(n, ch, self.batcher.out_h, self.batcher.out_w, dtype=torch.float32, device=self.batcher.device)
            for idx, imgs, flags, borders in groups:
                out[idx.to(out.device)] = self.spec.assemble(self.batchers, self._pixels(imgs), flags, borders)
            yield out


def list_images_in_dir(path):
    """dataset.py:84-93: the .jpg / .gif / .png files of a directory, in os.listdir order"""
    return [os.path.join(path, f) for f in os.listdir(path) if os.path.splitext(f)[1].lower() in (".jpg", ".gif", ".png")]


class DigitalMonstersDataset(data.Dataset):
    """`monsters128` (dataset.py:96-149, train_soft_intro_vae.py:418-423): pokemon / digimon / nexomon sprites, resized by
    load_image (no mirror, no crop) and augmented per item with torchvision's PIL transforms -- RandomAffine(0, translate = 5 px,
    white fill), ColorJitter(hue=0.5), RandomHorizontalFlip -- before ToTensor.  Those augmentations draw from torch's CPU
    generator and work on PIL images, so this data set stays on the host (a few thousand 128x128 sprites: not a loader
    bottleneck); it exists here because the reference's own class no longer constructs under current torchvision
    (`RandomAffine(fillcolor=...)` became `fill=` in 0.13) -- same transforms, same order, same draws."""

    def __init__(self, root_path, input_height=None, input_width=None, output_height=128, output_width=None, is_gray=False,
                 pokemon=True, digimon=True, nexomon=True):
        super(DigitalMonstersDataset, self).__init__()
        import torchvision.transforms as transforms
        image_list = []
        for wanted, label, sub in ((pokemon, "pokemon", ("pokemon",)), (digimon, "digimon", ("digimon", "200")),
                                   (nexomon, "nexomon", ("nexomon",))):
            if wanted:
                print("collecting %s..." % label)
                image_list.extend(list_images_in_dir(os.path.join(root_path, *sub)))
        print(f'total images: {len(image_list)}')
        self.image_filenames = image_list
        self.input_height, self.input_width = input_height, input_width
        self.output_height, self.output_width = output_height, output_width
        self.root_path, self.is_gray = root_path, is_gray
        self.input_transform = transforms.Compose([
            transforms.RandomAffine(0, translate=(5 / output_height, 5 / output_height), fill=(255, 255, 255)),
            transforms.ColorJitter(hue=0.5),
            transforms.RandomHorizontalFlip(p=0.5),
            transforms.ToTensor()])

    def __getitem__(self, index):
        from PIL import Image
        # load_image(..., crop_height=None, is_mirror=False, is_gray=False), dataset.py:140-142, on the host
        img = Image.open(self.image_filenames[index])
        if img.mode != 'RGB':
            img = img.convert('RGB')
        ow = self.output_width if self.output_width is not None else self.output_height
        if self.input_height is not None:
            iw = self.input_width if self.input_width is not None else self.input_height
            img = img.resize((iw, self.input_height), Image.BICUBIC)
        img = img.resize((ow, self.output_height), Image.BICUBIC)
        return self.input_transform(img)

    def __len__(self):
        return len(self.image_filenames)


def load_image(file_path, input_height=128, input_width=None, output_height=128, output_width=None,
               crop_height=None, crop_width=None, is_random_crop=True, is_mirror=True, is_gray=False, device="cuda:0"):
    """Reference signature (dataset.py:12-13) for single images; returns the PIL image the reference returns, computed on
    the GPU (same draws from Python's `random`, same pixels)."""
    from PIL import Image
    a = decode_image(file_path, is_gray)
    spec = LoadSpec(input_height, input_width, output_height, output_width, crop_height, crop_width, is_random_crop, is_mirror)
    mirror, margins = spec.draw(a.shape[0], a.shape[1])
    u8 = spec.assemble({"device": torch.device(device)}, torch.from_numpy(a)[None], [mirror],
                       None if margins is None else [list(margins)], as_u8=True)[0].cpu().numpy()
    return Image.fromarray(u8[:, :, 0], 'L') if u8.shape[2] == 1 else Image.fromarray(u8, 'RGB')
