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

Same constructor signature as the reference class.  Supported: what the image configs pass
(train_soft_intro_vae.py:388-392, 400-404, 415-417): ``input_height=None, crop_height=None``.  The two-stage resize and
the crops of ``load_image`` raise ``NotImplementedError`` -- there is no CPU fallback on this path.
"""
import ctypes as C
import os
import random

import numpy as np
import torch
import torch.utils.data as data

from . import lib as L


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

    def __call__(self, images_u8, mirror=None, out=None):
        """images_u8: uint8 [B,H,W,C] (host or device); mirror: [B] flags or None -> float32 [B,C,out_h,out_w] on device"""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
            raise ValueError("expected a uint8 [B,H,W,C] batch of decoded images")
        B, H, W, ch = images_u8.shape
        src = images_u8.to(self.device, non_blocking=True).contiguous()
        flags = None
        if mirror is not None:
            flags = torch.as_tensor(mirror).to(dtype=torch.uint8).to(self.device, non_blocking=True).contiguous()
        if out is None:
            out = torch.empty(B, ch, self.out_h, self.out_w, dtype=torch.float32, device=self.device)
        plan = self._plan(H, W)
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_image_batch_u8(L.ptr(src), L.ptr(flags), B, H, W, ch, self.out_h, self.out_w, L.ptr(plan),
                                                  L.ptr(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "sivae_image_batch_u8")
        return out


class ImageDatasetFromFile(data.Dataset):
    """Reference signature (dataset.py:50-53).  __getitem__ -> (uint8 [H,W,C] decoded pixels, mirror flag): the rest of
    load_image + ToTensor happens per batch on the GPU (ImageBatcher / GpuImageLoader)."""

    def __init__(self, image_list, root_path,
                 input_height=128, input_width=None, output_height=128, output_width=None,
                 crop_height=None, crop_width=None, is_random_crop=False, is_mirror=True, is_gray=False):
        super(ImageDatasetFromFile, self).__init__()
        if input_height is not None or input_width is not None or crop_height is not None or crop_width is not None:
            raise NotImplementedError("GPU batch assembly covers input_height=None, crop_height=None (what the image "
                                      "configs pass); the two-stage resize / crops of load_image are not implemented")
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

    def __getitem__(self, index):
        a = decode_image(os.path.join(self.root_path, self.image_filenames[index]), self.is_gray)
        mirror = 1 if (self.is_mirror and random.randint(0, 1) == 0) else 0      # dataset.py:26, same draw
        return torch.from_numpy(a), mirror

    def __len__(self):
        return len(self.image_filenames)


def collate_decoded(samples):
    """-> list of (indices, uint8 [b,H,W,C], flags [b]) groups, one per source geometry (usually exactly one)"""
    groups = {}
    for i, (img, flag) in enumerate(samples):
        groups.setdefault(tuple(img.shape), []).append((i, img, flag))
    out = []
    for items in groups.values():
        idx = torch.tensor([i for i, _, _ in items], dtype=torch.long)
        out.append((idx, torch.stack([im for _, im, _ in items]), torch.tensor([f for _, _, f in items], dtype=torch.uint8)))
    return out


class GpuImageLoader:
    """Iterates a DataLoader over an ImageDatasetFromFile (collate_fn=collate_decoded) and yields what the reference's
    loader yields -- float32 [B,C,S,S] batches in [0,1] -- already on `device`."""

    def __init__(self, loader, device):
        self.loader = loader
        ds = loader.dataset
        self.batcher = ImageBatcher(ds.output_height, ds.output_width, device)
        self.dataset, self.batch_size = ds, loader.batch_size

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for groups in self.loader:
            if len(groups) == 1:
                yield self.batcher(groups[0][1], groups[0][2])
                continue
            n = sum(g[0].numel() for g in groups)
            ch = groups[0][1].shape[-1]
            out = torch.empty(n, ch, self.batcher.out_h, self.batcher.out_w, dtype=torch.float32, device=self.batcher.device)
            for idx, imgs, flags in groups:
                out[idx.to(out.device)] = self.batcher(imgs, flags)
            yield out


def load_image(file_path, input_height=128, input_width=None, output_height=128, output_width=None,
               crop_height=None, crop_width=None, is_random_crop=True, is_mirror=True, is_gray=False, device="cuda:0"):
    """Reference signature (dataset.py:12-13) for single images; returns the PIL image the reference returns, computed on
    the GPU.  Only the configuration the image configs use (input_height=None, crop_height=None) is implemented."""
    from PIL import Image
    if input_height is not None or crop_height is not None:
        raise NotImplementedError("load_image on the GPU path: input_height=None and crop_height=None only")
    a = decode_image(file_path, is_gray)
    mirror = 1 if (is_mirror and random.randint(0, 1) == 0) else 0
    t = ImageBatcher(output_height, output_width, device)(torch.from_numpy(a)[None], [mirror])[0]
    u8 = torch.round(t * 255.0).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
    return Image.fromarray(u8[:, :, 0], 'L') if u8.shape[2] == 1 else Image.fromarray(u8, 'RGB')
