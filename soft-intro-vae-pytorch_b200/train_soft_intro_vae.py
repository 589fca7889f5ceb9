"""
Drop-in replacement of the reference module ``soft_intro_vae/train_soft_intro_vae.py`` whose compute runs in the
B200-native engine (libsivae_b200.so, hand-written sm_100a CUDA kernels) instead of torch.nn / autograd.

Same public names, signatures and defaults as the reference (file:line of the original in brackets):
  ResidualBlock [:38-75]  Encoder [:78-122]  Decoder [:125-169]  SoftIntroVAE [:172-223]
  calc_kl [:231-251]  reparameterize [:254-265]  calc_reconstruction_loss [:268-294]
  str_to_list / is_image_file / record_scalar / record_image [:297-313]  load_model / save_checkpoint [:316-329]
  train_soft_intro_vae [:337-702]
so that ``soft_intro_vae/main.py`` (``from train_soft_intro_vae import train_soft_intro_vae``) runs unmodified
when this directory precedes the reference's on ``sys.path``.

What is different by construction
  * the modules are containers of parameters (same state_dict schema / init / RNG consumption as the reference);
    their ``forward`` dispatches to the engine and returns plain tensors without an autograd graph -- the backward
    of the introspective step is hand-written inside the engine (``sivae_e_step`` / ``sivae_d_step``).
  * CUDA only.  There is no CPU fallback: on a CPU device the image model raises.
  * extra ``dataset`` values ``synthetic32 / synthetic128 / synthetic256`` (uniform [0,1) images, no files), used by
    bench.py; data-parallel when ``torch.distributed`` is initialised (one process per GPU, NCCL all-reduce of the
    flat encoder / decoder gradient buffers before each Adam step, BN statistics stay per rank).
"""
import contextlib
import importlib
import os
import pickle
import random
import sys
import time
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))
_PKG = os.path.basename(_HERE)
_L = importlib.import_module(_PKG + ".lib")
_E = importlib.import_module(_PKG + ".engine")
if os.environ.get("SIVAE_ANNOUNCE") == "1":      # which file served `import train_soft_intro_vae` (tests/test_gpu_boundary.py)
    print("SIVAE_DROPIN %s %s" % (__name__, os.path.abspath(__file__)), file=sys.stderr)

__all__ = ["ResidualBlock", "Encoder", "Decoder", "SoftIntroVAE", "calc_kl", "reparameterize",
           "calc_reconstruction_loss", "str_to_list", "is_image_file", "record_scalar", "record_image", "load_model",
           "save_checkpoint", "train_soft_intro_vae"]


# ======================================================================================================
# model containers
# ======================================================================================================
def _engine_only(what):
    raise RuntimeError(what + ": this module is a parameter container of the B200 engine; call it through a "
                       "SoftIntroVAE placed on a CUDA device (no CPU / autograd fallback exists).")


class ResidualBlock(nn.Module):
    """Parameter container of one residual block (reference :38-75): conv_expand (1x1, only if inc != outc), conv1
    3x3, bn1, conv2 3x3, bn2.  Built from the same torch.nn constructors in the same order, hence the same init."""

    def __init__(self, inc=64, outc=64, groups=1, scale=1.0):
        super().__init__()
        if groups != 1 or int(outc * scale) != outc:
            raise NotImplementedError("the B200 engine implements groups=1, scale=1.0 (all the reference driver uses)")
        self.conv_expand = nn.Conv2d(inc, outc, 1, 1, 0, groups=1, bias=False) if inc != outc else None
        self.conv1 = nn.Conv2d(inc, outc, 3, 1, 1, groups=groups, bias=False)
        self.bn1 = nn.BatchNorm2d(outc)
        self.relu1 = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = nn.Conv2d(outc, outc, 3, 1, 1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(outc)
        self.relu2 = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        _engine_only("ResidualBlock.forward")


def _conv_output_size(channels, image_size):
    sz = image_size // 2
    for _ in channels[1:]:
        sz //= 2
    return torch.Size([channels[-1], sz, sz])


class _NetBase(nn.Module):
    _owner = None

    def _eng(self, batch):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            _engine_only(type(self).__name__ + ".forward")
        return owner, owner._ensure_engine(batch)

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            owner.params_changed()
        return out


class Encoder(_NetBase):
    def __init__(self, cdim=3, zdim=512, channels=(64, 128, 256, 512, 512, 512), image_size=256, conditional=False,
                 cond_dim=10):
        super().__init__()
        self.zdim, self.cdim, self.image_size = zdim, cdim, image_size
        self.conditional, self.cond_dim = conditional, cond_dim
        channels = list(channels)
        cc = channels[0]
        self.main = nn.Sequential(nn.Conv2d(cdim, cc, 5, 1, 2, bias=False), nn.BatchNorm2d(cc), nn.LeakyReLU(0.2),
                                  nn.AvgPool2d(2))
        sz = image_size // 2
        for ch in channels[1:]:
            self.main.add_module("res_in_{}".format(sz), ResidualBlock(cc, ch, scale=1.0))
            self.main.add_module("down_to_{}".format(sz // 2), nn.AvgPool2d(2))
            cc, sz = ch, sz // 2
        self.main.add_module("res_in_{}".format(sz), ResidualBlock(cc, cc, scale=1.0))
        self._channels = channels
        self.conv_output_size = self.calc_conv_output_size()
        num_fc_features = int(np.prod(self.conv_output_size))
        print("conv shape: ", self.conv_output_size)
        print("num fc features: ", num_fc_features)
        # conditional: the condition is concatenated to the flattened features in forward (:106-109, :118-119)
        self.fc = nn.Linear(num_fc_features + (self.cond_dim if self.conditional else 0), 2 * zdim)

    def calc_conv_output_size(self):
        """The reference pushes a zero image through `main` in train mode here (:111-114).  The shape is computed
        arithmetically instead; the side effect on the BN buffers of that all-zero batch (running_mean stays 0,
        running_var -> 0.9, num_batches_tracked -> 1) is reproduced exactly."""
        for m in self.main.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_var.fill_(0.9)
                m.num_batches_tracked.fill_(1)
        return _conv_output_size(self._channels, self.image_size)

    def forward(self, x, o_cond=None):
        owner, eng = self._eng(x.size(0))
        # :118-119 -- a conditional encoder called WITHOUT a condition fails in its fc layer in the reference (matmul shape
        # error); the engine refuses the call the same way (RuntimeError)
        cond = o_cond if (self.conditional and o_cond is not None) else None
        return eng.encode(owner._as_input(x), self.training, cond=cond)


class Decoder(_NetBase):
    def __init__(self, cdim=3, zdim=512, channels=(64, 128, 256, 512, 512, 512), image_size=256, conditional=False,
                 conv_input_size=None, cond_dim=10):
        super().__init__()
        self.cdim, self.image_size, self.conditional, self.cond_dim = cdim, image_size, conditional, cond_dim
        channels = list(channels)
        cc = channels[-1]
        self.conv_input_size = conv_input_size
        num_fc_features = cc * 4 * 4 if conv_input_size is None else int(np.prod(conv_input_size))
        # conditional: z is concatenated with the condition in forward (:139-143, :163-165)
        self.fc = nn.Sequential(nn.Linear(zdim + (cond_dim if conditional else 0), num_fc_features), nn.ReLU(True))
        sz = 4
        self.main = nn.Sequential()
        for ch in channels[::-1]:
            self.main.add_module("res_in_{}".format(sz), ResidualBlock(cc, ch, scale=1.0))
            self.main.add_module("up_to_{}".format(sz * 2), nn.Upsample(scale_factor=2, mode="nearest"))
            cc, sz = ch, sz * 2
        self.main.add_module("res_in_{}".format(sz), ResidualBlock(cc, cc, scale=1.0))
        self.main.add_module("predict", nn.Conv2d(cc, cdim, 5, 1, 2))
        self._net_id = _L.NET_DECODER

    def forward(self, z, y_cond=None):
        owner, eng = self._eng(z.size(0))
        z = z.reshape(z.size(0), -1).to(device=eng.device, dtype=torch.float32).contiguous()
        cond = y_cond if (self.conditional and y_cond is not None) else None
        return eng.decode(z, self.training, net=self._net_id, cond=cond)


class SoftIntroVAE(nn.Module):
    """SoftIntroVAE container (reference :172-223).  Owns the native engine once placed on a CUDA device."""
    _bootstrap = False

    def __init__(self, cdim=3, zdim=512, channels=(64, 128, 256, 512, 512, 512), image_size=256, conditional=False,
                 cond_dim=10):
        super().__init__()
        self.zdim, self.conditional, self.cond_dim = zdim, conditional, cond_dim
        self._arch = dict(cdim=cdim, zdim=zdim, channels=list(channels), image_size=image_size,
                          cond_dim=int(cond_dim) if conditional else 0)
        self._engine = None
        self._conv_backend = int(os.environ.get("SIVAE_CONV_BACKEND", _L.CONV_AUTO))
        self.encoder = Encoder(cdim, zdim, channels, image_size, conditional=conditional, cond_dim=cond_dim)
        self.decoder = Decoder(cdim, zdim, channels, image_size, conditional=conditional,
                               conv_input_size=self.encoder.conv_output_size, cond_dim=cond_dim)
        self._wire()

    def _wire(self):
        ref = weakref.ref(self)
        for name, net_id in (("encoder", _L.NET_ENCODER), ("decoder", _L.NET_DECODER), ("target_decoder", _L.NET_TARGET)):
            m = getattr(self, name, None)
            if m is not None:
                m._owner = ref
                m._net_id = net_id

    def _nets(self):
        out = {_L.NET_ENCODER: self.encoder, _L.NET_DECODER: self.decoder}
        if self._bootstrap:
            out[_L.NET_TARGET] = self.target_decoder
        return out

    # ---- engine ownership ---------------------------------------------------------------------------------
    def _device(self):
        return next(self.parameters()).device

    def _ensure_engine(self, batch):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("SoftIntroVAE (B200 engine) must be on a CUDA device; there is no CPU fallback")
        if self._engine is None or self._engine.device != dev:
            self._attach(dev, max(int(batch), 1))
        elif batch > self._engine.max_batch:
            self._engine.grow(int(batch))
        return self._engine

    def reserve(self, batch):
        """size the engine workspace for `batch` images per step (avoids a re-allocation on the first step)"""
        return self._ensure_engine(batch)

    def _attach(self, dev, batch):
        old = self._engine
        a = self._arch
        eng = _E.Engine(a["cdim"], a["zdim"], a["channels"], a["image_size"], batch, dev, bootstrap=self._bootstrap,
                        conv_backend=self._conv_backend, cond_dim=a["cond_dim"])
        with torch.no_grad():
            for net_id, mod in self._nets().items():
                mem = eng.mem[net_id]
                params = dict(mod.named_parameters())
                if len(params) != len(mem.tensors):
                    raise RuntimeError("parameter schema mismatch between module and engine")
                for name, kind, off, numel, shape in mem.tensors:
                    p = params[name]
                    if tuple(p.shape) != tuple(shape):
                        raise RuntimeError("shape mismatch for %s: %s vs %s" % (name, tuple(p.shape), shape))
                    v = _E.NetMemory.view(mem.params, kind, off, numel, shape)
                    v.copy_(p.data.to(dev))
                    p.data = v
                    if mem.grads is not None:
                        p.grad = _E.NetMemory.view(mem.grads, kind, off, numel, shape)
                for name, c, off, idx in mem.bns:
                    bn = mod.get_submodule(name)
                    mem.bn[off:off + c].copy_(bn.running_mean.to(dev))
                    mem.bn[off + c:off + 2 * c].copy_(bn.running_var.to(dev))
                    mem.nbt[idx] = int(bn.num_batches_tracked)
                    bn.running_mean = mem.bn[off:off + c]
                    bn.running_var = mem.bn[off + c:off + 2 * c]
                    bn.num_batches_tracked = mem.nbt[idx]
                if old is not None and old.mem[net_id].m is not None:
                    mem.m.copy_(old.mem[net_id].m.to(dev))
                    mem.v.copy_(old.mem[net_id].v.to(dev))
                    _L.load().sivae_adam_set_step(eng.handle, net_id, _L.load().sivae_adam_get_step(old.handle, net_id))
        if old is not None:
            eng.recon_loss = old.recon_loss
            old.close()
        self._engine = eng
        eng.params_changed()

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        eng = self._engine
        if eng is not None:
            p = next(self.parameters())
            lo, hi = eng.mem[_L.NET_ENCODER].params.data_ptr(), eng.mem[_L.NET_ENCODER].params.data_ptr() + eng.mem[_L.NET_ENCODER].params.numel() * 4
            if not (p.is_cuda and lo <= p.data_ptr() < hi):
                if p.is_cuda:
                    self._attach(p.device, eng.max_batch)
                else:
                    self._engine = None
                    eng.close()
        return self

    def params_changed(self):
        if self._engine is not None:
            self._engine.params_changed()

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.params_changed()
        return out

    def _as_input(self, x):
        a = self._arch
        if x.dim() != 4 or tuple(x.shape[1:]) != (a["cdim"], a["image_size"], a["image_size"]):
            raise ValueError("expected input of shape [B,%d,%d,%d], got %s" % (a["cdim"], a["image_size"], a["image_size"], tuple(x.shape)))
        return x.to(device=self._device(), dtype=torch.float32).contiguous()

    # ---- reference API ------------------------------------------------------------------------------------
    def forward(self, x, o_cond=None, deterministic=False):
        mu, logvar = self.encode(x, o_cond=o_cond)            # :186-201: the condition reaches both fc layers
        z = mu if deterministic else reparameterize(mu, logvar)
        y = self.decode(z, y_cond=o_cond)
        return mu, logvar, z, y

    def sample(self, z, y_cond=None):
        return self.decode(z, y_cond=y_cond)

    def sample_with_noise(self, num_samples=1, device=torch.device("cpu"), y_cond=None):
        # the reference reads the non-existent attribute `self.z_dim` here (:208, AttributeError); `zdim` is meant
        z = torch.randn(num_samples, self.zdim).to(device)
        return self.decode(z, y_cond=y_cond)

    def encode(self, x, o_cond=None):
        return self.encoder(x, o_cond=o_cond)

    def decode(self, z, y_cond=None):
        return self.decoder(z, y_cond=y_cond)


# ======================================================================================================
# helpers (tensor-level API kept for callers of the reference helpers; plain torch, not on the hot path)
# ======================================================================================================
def calc_kl(logvar, mu, mu_o=0.0, logvar_o=0.0, reduce='sum'):
    """KL( N(mu, e^logvar) || N(mu_o, e^logvar_o) ) per sample, then 'sum' | 'mean' | anything else = none."""
    mu_o = torch.as_tensor(mu_o, dtype=mu.dtype, device=mu.device)
    logvar_o = torch.as_tensor(logvar_o, dtype=mu.dtype, device=mu.device)
    inner = 1 + logvar - logvar_o - (logvar - logvar_o).exp() - (mu - mu_o).pow(2) * torch.exp(-logvar_o)
    kl = -0.5 * inner.sum(1)
    if reduce == 'sum':
        return kl.sum()
    if reduce == 'mean':
        return kl.mean()
    return kl


def reparameterize(mu, logvar):
    """z = mu + eps * exp(logvar / 2), eps ~ N(0, I) drawn on mu's device."""
    return torch.addcmul(mu, torch.randn_like(logvar), torch.exp(0.5 * logvar))


def calc_reconstruction_loss(x, recon_x, loss_type='mse', reduction='sum'):
    """per-sample reconstruction error; argument order (x, recon_x) as in the reference."""
    if reduction not in ('sum', 'mean', 'none'):
        raise NotImplementedError
    b = recon_x.size(0)
    recon_x, x = recon_x.reshape(b, -1), x.reshape(b, -1)
    if loss_type == 'mse':
        err = (recon_x - x).square().sum(1)
        return err.sum() if reduction == 'sum' else err.mean() if reduction == 'mean' else err
    if loss_type == 'l1':
        return F.l1_loss(recon_x, x, reduction=reduction)
    if loss_type == 'bce':
        return F.binary_cross_entropy(recon_x, x, reduction=reduction)
    raise NotImplementedError


def str_to_list(x):
    return [int(v) for v in x.split(',')]


def is_image_file(filename):
    return filename.endswith((".jpg", ".png", ".jpeg", ".bmp"))


def record_scalar(writer, scalar_list, scalar_name_list, cur_iter):
    names = [n.strip(' ') for n in scalar_name_list[1:-1].split(',')]
    for name, item in zip(names, scalar_list):
        writer.add_scalar(name, item, cur_iter)


def record_image(writer, image_list, cur_iter, num_rows=8):
    from torchvision.utils import make_grid
    writer.add_image('visualization', make_grid(torch.cat(image_list, dim=0), nrow=num_rows), cur_iter)


def load_model(model, pretrained, device):
    """reference :316-318.  A checkpoint written by this module also carries `sivae_train_state` (see save_checkpoint); it is
    kept on the model and applied by the trainer once the engine exists, so that training RESUMES (optimiser moments and step
    counters, RNG streams, iteration counter) instead of restarting Adam from zero like the reference does."""
    weights = torch.load(pretrained, map_location=device)
    model.load_state_dict(weights['model'], strict=False)
    model._resume_state = weights.get("sivae_train_state")


def _train_state(model, iteration):
    """what the reference's checkpoints lack: Adam moments / step counters per parameter (reference key names
    and logical shapes), the four RNG streams the trainer draws from, the iteration counter"""
    eng = getattr(model, "_engine", None)
    if eng is None:
        return None
    lib = _L.load()
    adam = {}
    for net_id, mod in model._nets().items():
        mem = eng.mem[net_id]
        if mem.m is None:
            continue
        pref = {_L.NET_ENCODER: "encoder.", _L.NET_DECODER: "decoder."}[net_id]
        ent = {"step": int(lib.sivae_adam_get_step(eng.handle, net_id)), "m": {}, "v": {}}
        for name, kind, off, numel, shape in mem.tensors:
            ent["m"][pref + name] = _E.NetMemory.view(mem.m, kind, off, numel, shape).detach().cpu().contiguous()
            ent["v"][pref + name] = _E.NetMemory.view(mem.v, kind, off, numel, shape).detach().cpu().contiguous()
        adam[net_id] = ent
    # plain tensors / ints / floats / strings only, so that torch.load(weights_only=True) -- the default the reference's own
    # load_model runs into under current torch -- still accepts the file
    pv, pstate, pgauss = random.getstate()
    nname, nkey, npos, nhas, ncached = np.random.get_state()
    rng = {"python": {"version": int(pv), "state": torch.tensor(list(pstate), dtype=torch.int64), "gauss": pgauss},
           "numpy": {"name": str(nname), "key": torch.from_numpy(nkey.astype(np.int64)), "pos": int(npos), "has_gauss": int(nhas),
                     "cached": float(ncached)},
           "torch_cpu": torch.get_rng_state(), "torch_cuda": torch.cuda.get_rng_state(eng.device)}
    return {"version": 1, "adam": adam, "rng": rng, "cur_iter": int(iteration)}


def _restore_train_state(model, st, restore_rng=True):
    eng = model._engine
    lib = _L.load()
    with torch.no_grad():
        for net_id, ent in st["adam"].items():
            mem = eng.mem[net_id]
            pref = {_L.NET_ENCODER: "encoder.", _L.NET_DECODER: "decoder."}[net_id]
            for name, kind, off, numel, shape in mem.tensors:          # (map_location may have put the saved tensors anywhere)
                _E.NetMemory.view(mem.m, kind, off, numel, shape).copy_(ent["m"][pref + name].to(mem.m.device))
                _E.NetMemory.view(mem.v, kind, off, numel, shape).copy_(ent["v"][pref + name].to(mem.v.device))
            lib.sivae_adam_set_step(eng.handle, net_id, int(ent["step"]))
    if restore_rng:
        r = st["rng"]
        random.setstate((r["python"]["version"], tuple(int(x) for x in r["python"]["state"].tolist()), r["python"]["gauss"]))
        np.random.set_state((r["numpy"]["name"], r["numpy"]["key"].cpu().numpy().astype(np.uint32), r["numpy"]["pos"], r["numpy"]["has_gauss"],
                             r["numpy"]["cached"]))
        torch.set_rng_state(r["torch_cpu"].cpu())
        torch.cuda.set_rng_state(r["torch_cuda"].cpu(), eng.device)
    return int(st.get("cur_iter", 0))


_save_thread = None


def _join_async_save():
    global _save_thread
    if _save_thread is not None:
        _save_thread.join()
        _save_thread = None


def save_checkpoint(model, epoch, iteration, prefix=""):
    """reference :321-329: `{"epoch", "model": state_dict}` at ./saves/<prefix>model_epoch_{e}_iter_{i}.pth -- loadable by the
    reference.  Extensions a reference loader ignores (it reads only ['model']): the key `sivae_train_state` (optimiser + RNG
    state for an exact resume; SIVAE_CKPT_STATE=0 omits it) and SIVAE_ASYNC_SAVE=1 = the file is written by a background thread
    from a host snapshot taken here (the pattern of style_soft_intro_vae/checkpointer.py:60-67); joined before the next save and
    at the end of training."""
    global _save_thread
    os.makedirs("./saves/", exist_ok=True)
    path = "./saves/" + prefix + "model_epoch_{}_iter_{}.pth".format(epoch, iteration)
    # contiguous host copies in the reference layout, so the file is loadable by the reference and small
    state = {k: v.detach().cpu().clone().contiguous() for k, v in model.state_dict().items()}
    payload = {"epoch": epoch, "model": state}
    if os.environ.get("SIVAE_CKPT_STATE", "1") != "0":
        ts = _train_state(model, iteration)
        if ts is not None:
            payload["sivae_train_state"] = ts
    _join_async_save()
    if os.environ.get("SIVAE_ASYNC_SAVE", "0") == "1":
        import threading
        _save_thread = threading.Thread(target=torch.save, args=(payload, path), daemon=False)
        _save_thread.start()
    else:
        torch.save(payload, path)
    print("model checkpoint saved @ {}".format(path))


# ======================================================================================================
# data sets
# ======================================================================================================
_ARCH = {32: [64, 128, 256], 128: [64, 128, 256, 512, 512], 256: [64, 128, 256, 512, 512, 512]}


class _Synthetic(torch.utils.data.Dataset):
    """uniform [0,1) images (the ToTensor range) -- benchmark / smoke input, no files."""

    def __init__(self, n, ch, size, seed=1234):
        self.x = torch.rand(n, ch, size, size, generator=torch.Generator().manual_seed(seed))

    def __len__(self):
        return self.x.size(0)

    def __getitem__(self, i):
        return self.x[i]


def _reference_dataset_module():
    try:
        return importlib.import_module("dataset")
    except ImportError as e:      # the file loaders are the reference's (out of the hot-path scope)
        raise ImportError("dataset '%s' needs the reference's dataset.py on sys.path (soft_intro_vae/dataset.py)") from e


def _build_dataset(dataset):
    """-> (train_set, image_size, channels, ch, labelled)   (dataset switch of the reference, :376-440)"""
    if dataset.startswith("synthetic"):
        size = int(dataset[len("synthetic"):].split(":")[0] or 32)
        n = int(dataset.split(":")[1]) if ":" in dataset else 512
        if size not in _ARCH:
            raise NotImplementedError("synthetic sizes: 32, 128, 256")
        return _Synthetic(n, 3, size), size, _ARCH[size], 3, False
    if dataset in ("cifar10", "svhn", "mnist", "fmnist"):
        from torchvision import transforms
        from torchvision import datasets as tvd
        tt = transforms.ToTensor()
        if dataset == "cifar10":
            return tvd.CIFAR10(root='./cifar10_ds', train=True, download=True, transform=tt), 32, [64, 128, 256], 3, True
        if dataset == "svhn":
            return tvd.SVHN(root='./svhn', split='train', transform=tt, download=True), 32, [64, 128, 256], 3, True
        if dataset == "fmnist":
            return tvd.FashionMNIST(root='./fmnist_ds', train=True, download=True, transform=tt), 28, [64, 128], 1, True
        return tvd.MNIST(root='./mnist_ds', train=True, download=True, transform=tt), 28, [64, 128], 1, True
    files = {"celeb128": (128, [64, 128, 256, 512, 512], '../data/celeb256/img_align_celeba', 162770),
             "celeb256": (256, [64, 128, 256, 512, 512, 512], '../data/celeb256/img_align_celeba', 162770),
             "celeb1024": (1024, [16, 32, 64, 128, 256, 512, 512, 512], './celeb1024', 29000)}
    if dataset in files:
        size, channels, root, n_train = files[dataset]
        names = [f for f in os.listdir(root) if is_image_file(f)][:n_train]
        assert len(names) > 0
        # decode on the host, mirror + bicubic resize + ToTensor per batch on the GPU (gpu_dataset.py; bit-exact with the
        # reference's loader).  SIVAE_GPU_LOADER=0: the reference's own dataset.py (PIL on the DataLoader workers).
        mod = _reference_dataset_module() if os.environ.get("SIVAE_GPU_LOADER", "1") == "0" else \
            importlib.import_module(_PKG + ".gpu_dataset")
        ds = mod.ImageDatasetFromFile(names, root, input_height=None, crop_height=None, output_height=size, is_mirror=True)
        return ds, size, channels, 3, False
    if dataset == "monsters128":
        # host-side data set (torchvision PIL augmentations); the reference's own class does not construct under torchvision >= 0.13
        ds = importlib.import_module(_PKG + ".gpu_dataset").DigitalMonstersDataset(root_path='./monsters_ds/', output_height=128)
        return ds, 128, [64, 128, 256, 512, 512], 3, False
    raise NotImplementedError("dataset is not supported")


# ======================================================================================================
# training driver
# ======================================================================================================
class _Tracker:
    """running per-epoch means of the logged statistics (:498-504, :632-639)"""
    KEYS = ("kl_real", "kl_fake", "kl_rec", "rec_err", "exp_elbo_f", "exp_elbo_r", "diff_kl")

    def __init__(self):
        self.hist = {k: [] for k in self.KEYS}
        self.reset()

    def reset(self):
        self.cur = {k: [] for k in self.KEYS}

    def add(self, **kw):
        for k, v in kw.items():
            self.cur[k].append(v)

    def close_epoch(self):
        for k in self.KEYS:
            if k != "diff_kl":
                self.hist[k].append(float(np.mean(self.cur[k])))


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def _graph_mode():
    """SIVAE_CUDA_GRAPH: 0 = never; 1 (default) = replay the whole iteration from ONE CUDA graph -- under torch.distributed
    too: the two gradient all-reduces are raw ncclAllReduce calls of the library's own communicator on the step's stream
    (sivae_comm_init), captured with the kernels around them."""
    try:
        return int(os.environ.get("SIVAE_CUDA_GRAPH", "1"))
    except ValueError:
        return 1


def _dp_mode():
    """SIVAE_DP_COMM: 'lib' (default) = the library owns an NCCL communicator (one graph per iteration); 'torch' = round-1
    path: torch.distributed all-reduces launched eagerly between three graph segments (also what a non-NCCL backend gets)"""
    return os.environ.get("SIVAE_DP_COMM", "lib")


def _lib_comm(eng, dist):
    """the engine's own communicator, created on first use (None: keep using torch.distributed collectives)"""
    if _dp_mode() != "lib" or dist.get_backend() != "nccl":
        return False
    if eng.comm_world != dist.get_world_size():
        eng.comm_init(dist)
    return True


def _segmented_step(eng, dist, real, noise, eps, hp, lr_e, lr_d, inv_world, key):
    """data-parallel iteration as three replayed graph segments cut at the two all-reduces (which run eagerly on the
    current stream, ordered with the replays): [E half] AR(enc grads) [Adam(E) + D half] AR(dec grads) [Adam(D)]"""
    enc, dec = _L.NET_ENCODER, _L.NET_DECODER
    eng.graphed(("introspective/E",) + key, [real.contiguous(), noise.contiguous(), eps[:3].contiguous()],
                lambda r_, n_, e_: eng.e_step(r_, n_, e_, hp))
    dist.all_reduce(eng.mem[enc].grads)

    def d_half(e_):
        eng.adam(enc, lr_e, inv_world)
        eng.d_step(e_, hp)
    eng.graphed(("introspective/D",) + key, [eps[3:].contiguous()], d_half)
    dist.all_reduce(eng.mem[dec].grads)
    eng.graphed(("introspective/A",) + key, [], lambda: eng.adam(dec, lr_d, inv_world))


def introspective_iteration(model, real, noise, eps, hp, lr_e, lr_d, use_graph=None, reuse_decoder_passes=None):
    """One E-step + D-step through the engine (reference :551-624).  real: [B,C,S,S] on the model's device; noise:
    [B,z]; eps: [5,B,z].  Returns the 16-float device statistics tensor (see include/sivae.h).
    The ~2000 kernel launches of a step are replayed from CUDA graphs after the first two calls with the same batch
    size and hyper-parameters (Engine.graphed); use_graph=False (or SIVAE_CUDA_GRAPH=0) keeps every call eager.
    Data parallel (torch.distributed initialised): the flat encoder / decoder gradient buffers are
    sum-all-reduced once each, 1/world folded into the Adam kernel.  The graph is then cut at the two collectives --
    [E half] all-reduce [Adam(E) + D half] all-reduce [Adam(D)] -- so NCCL is never captured.
    reuse_decoder_passes (None: keep the engine's setting, default off / SIVAE_REUSE_DEC): the D half re-uses the E half's
    fake / rec decoder passes instead of recomputing them (bit-identical results, see Engine.reuse_decoder_passes)."""
    eng = model._ensure_engine(real.size(0))
    if reuse_decoder_passes is not None and bool(reuse_decoder_passes) != eng.reuse_decoder_passes:
        eng.reuse_decoder_passes = reuse_decoder_passes
    dist = _dist()
    inv_world = 1.0 / dist.get_world_size() if dist else 1.0
    enc, dec = _L.NET_ENCODER, _L.NET_DECODER
    own = bool(dist) and _lib_comm(eng, dist)          # collectives inside the library (raw NCCL on the step's stream)

    def step(real_, noise_, eps_):
        if own or not dist:
            return eng.iteration(real_, noise_, eps_, hp, lr_e, lr_d)       # one call: both halves, collectives, both Adam steps
        eng.e_step(real_, noise_, eps_[:3], hp)
        dist.all_reduce(eng.mem[enc].grads)
        eng.adam(enc, lr_e, inv_world)
        eng.d_step(eps_[3:], hp)
        dist.all_reduce(eng.mem[dec].grads)
        eng.adam(dec, lr_d, inv_world)

    mode = _graph_mode()
    if use_graph is None:
        use_graph = mode >= 1
    if use_graph and real.is_cuda and not torch.cuda.is_current_stream_capturing():
        key = (tuple(real.shape), bytes(hp), float(lr_e), float(lr_d), inv_world, eng.reuse_decoder_passes, own)
        if dist and not own:
            _segmented_step(eng, dist, real, noise, eps, hp, lr_e, lr_d, inv_world, key)
        else:
            eng.graphed(("introspective",) + key, [real.contiguous(), noise.contiguous(), eps.contiguous()], step)
        eng.note_batch(real.size(0))
    else:
        step(real.contiguous(), noise.contiguous(), eps.contiguous())
    return eng.stats


def vae_iteration(model, real, eps, hp, lr_e, lr_d):
    """vanilla VAE warm-up step (reference :512-536)"""
    eng = model._ensure_engine(real.size(0))
    dist = _dist()
    inv_world = 1.0 / dist.get_world_size() if dist else 1.0
    own = bool(dist) and _lib_comm(eng, dist)
    # bootstrap variant: model(real_batch) reconstructs through the frozen TARGET decoder (bootstrap :196-217, target=True by
    # default; warm-up at :546), so the trainable decoder receives no gradient and optimizer_d.step() moves nothing
    boot = bool(getattr(model, "_bootstrap", False))
    nets = (_L.NET_ENCODER,) if boot else (_L.NET_ENCODER, _L.NET_DECODER)
    eng.vae_step(real, eps, hp)
    for net in nets:
        if own:
            eng.allreduce_grads(net)
        elif dist:
            dist.all_reduce(eng.mem[net].grads)
    eng.adam(_L.NET_ENCODER, lr_e, inv_world)
    if not boot:
        eng.adam(_L.NET_DECODER, lr_d, inv_world)
    return eng.stats


class DevicePrefetcher:
    """Iterates `loader` one batch ahead: while the step on batch i runs, batch i+1 is fetched and moved to `device` on a side
    stream (for a GpuImageLoader that includes its resize kernel), so the copy engine works under the step instead of in
    front of it.  Yields what the loader yields, with every tensor on `device`; values and order are unchanged
    (the reference does `batch.to(device)` synchronously at :549)."""

    _streams = {}           # one side stream per device, shared by all prefetchers of the process

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.dataset = getattr(loader, "dataset", None)
        self.batch_size = getattr(loader, "batch_size", None)
        self._stream = DevicePrefetcher._streams.get(self.device)

    def __len__(self):
        return len(self.loader)

    def _move(self, b, main):
        if torch.is_tensor(b):
            t = b.to(self.device, non_blocking=True)
            t.record_stream(main)                 # allocated on the side stream, consumed on the main one
            return t
        if isinstance(b, (list, tuple)):
            return type(b)(self._move(x, main) for x in b)
        return b

    def __iter__(self):
        if self._stream is None:
            self._stream = DevicePrefetcher._streams.setdefault(self.device, torch.cuda.Stream(self.device))
        side, it = self._stream, iter(self.loader)

        def fetch():
            main = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(side):
                try:
                    b = next(it)
                except StopIteration:
                    return None
                b = self._move(b, main)
                ev = torch.cuda.Event()
                ev.record(side)
            return b, ev
        nxt = fetch()
        while nxt is not None:
            cur, ev = nxt
            nxt = fetch()                         # enqueue batch i+1 before the consumer launches (and waits on) step i
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur


def _milestone_lr(lr, epoch_steps, milestones=(350,), gamma=0.1):
    """MultiStepLR(milestones=(350,), gamma=0.1) stepped once per epoch (:453-454, :649-650)"""
    return lr * (gamma ** sum(1 for m in milestones if epoch_steps >= m))


def _save_grid(tensors, path, nrow):
    import torchvision.utils as vutils
    vutils.save_image(torch.cat(tensors, dim=0).data.cpu(), path, nrow=nrow)


def train_soft_intro_vae(dataset='cifar10', z_dim=128, lr_e=2e-4, lr_d=2e-4, batch_size=128, num_workers=4,
                         start_epoch=0, exit_on_negative_diff=False,
                         num_epochs=250, num_vae=0, save_interval=50, recon_loss_type="mse",
                         beta_kl=1.0, beta_rec=1.0, beta_neg=1.0, test_iter=1000, seed=-1, pretrained=None,
                         device=torch.device("cpu"), num_row=8, gamma_r=1e-8, with_fid=False):
    """Train a Soft-IntroVAE on the B200 engine.  Arguments, defaults, printed lines, output files
    (./figures_<dataset>/image_<iter>.jpg, ./saves/*.pth, ./soft_intro_train_graphs.jpg + _data.pickle) and the
    SystemError contract (NaN loss; negative KL difference) follow the reference function of the same name."""
    return _run_training(SoftIntroVAE, None, dataset, z_dim, lr_e, lr_d, batch_size, num_workers, start_epoch,
                         exit_on_negative_diff, num_epochs, num_vae, save_interval, recon_loss_type, beta_kl, beta_rec,
                         beta_neg, test_iter, seed, pretrained, device, num_row, gamma_r, with_fid)


def _run_training(model_cls, copy_to_target_freq, dataset, z_dim, lr_e, lr_d, batch_size, num_workers, start_epoch,
                  exit_on_negative_diff, num_epochs, num_vae, save_interval, recon_loss_type, beta_kl, beta_rec,
                  beta_neg, test_iter, seed, pretrained, device, num_row, gamma_r, with_fid):
    """shared driver of the standard and the bootstrap (copy_to_target_freq is not None) trainers"""
    if recon_loss_type not in _L.LOSS_TYPES:
        raise NotImplementedError                 # calc_reconstruction_loss :292-293
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("train_soft_intro_vae (B200 engine) needs device=torch.device('cuda:N'); no CPU fallback")
    if seed != -1:
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        torch.backends.cudnn.deterministic = True
        print("random seed: ", seed)

    train_set, image_size, channels, ch, labelled = _build_dataset(dataset)
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    main = rank == 0                 # under data parallelism only rank 0 prints, saves figures / checkpoints / the pickle
    with contextlib.redirect_stdout(None) if not main else contextlib.nullcontext():
        model = model_cls(cdim=ch, zdim=z_dim, channels=channels, image_size=image_size).to(device)
    if pretrained is not None:
        load_model(model, pretrained, device)
    if main:
        print(model)
    model.reserve(batch_size).recon_loss = recon_loss_type      # 'mse' | 'l1' | 'bce' (:268-294), in every loss term of the step
    resumed_iter = 0
    if getattr(model, "_resume_state", None) is not None:
        # checkpoint written by this module: continue Adam, the RNG streams and the iteration counter where they stopped
        resumed_iter = _restore_train_state(model, model._resume_state, restore_rng=os.environ.get("SIVAE_RESUME_RNG", "1") != "0")
        model._resume_state = None
    if dist:
        # every replica starts from rank 0's state (weights, BN buffers, Adam state): with the default seed = -1 each rank
        # would otherwise build different random weights and only the gradients would ever be averaged
        model._engine.broadcast_state(dist, src=0)
        # ... and draws its OWN noise / eps / shuffling from here on, also when a fixed seed made the generators identical
        torch.manual_seed(torch.initial_seed() + 7919 * rank)
        torch.cuda.manual_seed(torch.initial_seed())

    fig_dir = './figures_' + dataset.replace(":", "_")
    os.makedirs(fig_dir, exist_ok=True)
    scale = 1 / (ch * image_size ** 2)
    hp = _E.make_hyper(beta_kl, beta_rec, beta_neg, gamma_r, scale)
    hp_vae = hp

    sampler = None
    if dist:
        sampler = torch.utils.data.distributed.DistributedSampler(train_set, shuffle=True, seed=0 if seed == -1 else int(seed))
    gpu_ds = importlib.import_module(_PKG + ".gpu_dataset")
    on_gpu = isinstance(train_set, gpu_ds.ImageDatasetFromFile)
    loader = torch.utils.data.DataLoader(train_set, batch_size=batch_size, shuffle=sampler is None, sampler=sampler,
                                         num_workers=num_workers, pin_memory=True,
                                         collate_fn=gpu_ds.collate_decoded if on_gpu else None)
    if on_gpu:
        loader = gpu_ds.GpuImageLoader(loader, device)      # yields the float32 [B,C,S,S] batches the reference's loader yields
    fid_loader = loader
    if os.environ.get("SIVAE_PREFETCH", "1") != "0":
        loader = DevicePrefetcher(loader, device)           # host->device copy of batch i+1 under step i
    from tqdm import tqdm
    start_time = time.time()
    cur_iter = resumed_iter
    track = _Tracker()
    best_fid = None
    eps = torch.empty(5, batch_size, z_dim, device=device)
    prefix = "{}_soft_intro_betas_{}_{}_{}_".format(dataset, beta_kl, beta_neg, beta_rec)
    real_batch = None
    noise_host = None
    stats_host = [torch.empty(16).pin_memory() for _ in range(2)]
    async_stats = os.environ.get("SIVAE_ASYNC_STATS", "0") == "1"
    pending = None

    def consume_stats(p):
        sh_, ev_, epoch_, pbar_ = p
        ev_.synchronize()
        st = sh_.clone()
        model._engine.check_loss_domain(st)                                 # bce outside [0, 1]: RuntimeError like F.binary_cross_entropy
        if bool(st[15] != 0):                                               # isnan(lossD) or isnan(lossE) (:625-626)
            raise SystemError
        kl_real, kl_fake, kl_rec, rec_err = st[1].item(), st[7].item(), st[6].item(), st[5].item()
        pbar_.set_description_str('epoch #{}'.format(epoch_))
        pbar_.set_postfix(r_loss=rec_err, kl=kl_real, diff_kl=kl_fake - kl_real, expelbo_f=st[3].item())
        track.add(diff_kl=kl_fake - kl_real, kl_real=kl_real, kl_fake=kl_fake, kl_rec=kl_rec, rec_err=rec_err,
                  exp_elbo_f=st[3].item(), exp_elbo_r=st[2].item())

    for epoch in range(start_epoch, num_epochs):
        if with_fid and ((epoch == 0) or (epoch >= 100 and epoch % 20 == 0) or epoch == num_epochs - 1):
            from metrics.fid_score import calculate_fid_given_dataset     # the reference's evaluation package
            with torch.no_grad():
                print("calculating fid...")
                fid = calculate_fid_given_dataset(fid_loader, model, batch_size, cuda=True, dims=2048, device=device,
                                                  num_images=50000)
                print("fid:", fid)
                if best_fid is None:
                    best_fid = fid
                elif best_fid > fid:
                    print("best fid updated: {} -> {}".format(best_fid, fid))
                    best_fid = fid
                    if main:
                        save_checkpoint(model, epoch, cur_iter, prefix + "fid_" + str(fid) + "_")
        if epoch % save_interval == 0 and epoch > 0 and main:
            save_checkpoint(model, (epoch // save_interval) * save_interval, cur_iter, prefix)
        model.train()
        if sampler is not None:
            sampler.set_epoch(epoch)
        track.reset()
        cur_lr_e = _milestone_lr(lr_e, epoch - start_epoch)
        cur_lr_d = _milestone_lr(lr_d, epoch - start_epoch)
        pbar = tqdm(iterable=loader, disable=not main)
        for batch in pbar:
            if labelled:
                batch = batch[0]
            if batch.dim() == 3:
                batch = batch.unsqueeze(0)
            b_size = batch.size(0)
            if epoch < num_vae:
                real_batch = batch.to(device, non_blocking=True)
                e = torch.randn((b_size, z_dim), device=device)                     # the draw of reparameterize()
                st = vae_iteration(model, real_batch, e, hp_vae, cur_lr_e, cur_lr_d).cpu()
                model._engine.check_loss_domain(st)
                if bool(st[15] != 0):
                    raise SystemError
                pbar.set_description_str('epoch #{}'.format(epoch))
                pbar.set_postfix(r_loss=st[11].item(), kl=st[12].item())
                if cur_iter % test_iter == 0 and main:
                    _, _, _, rec = model(real_batch)
                    _save_grid([real_batch, rec], '{}/image_{}.jpg'.format(fig_dir, cur_iter), num_row)
            else:
                if noise_host is None or noise_host[0].size(0) != b_size:
                    noise_host = [torch.empty(b_size, z_dim).pin_memory() for _ in range(2)]   # persistent pinned staging
                nh = noise_host[cur_iter % 2]            # last used two steps ago: that step's statistics have been read
                torch.randn((b_size, z_dim), out=nh)                                # CPU generator, like the reference (:547)
                noise_batch = nh.to(device, non_blocking=True)
                real_batch = batch.to(device, non_blocking=True)
                e5 = eps if b_size == batch_size else torch.empty(5, b_size, z_dim, device=device)
                for i in range(5):                                                  # device generator, draw order of
                    torch.randn((b_size, z_dim), out=e5[i])                         # :560,567,568,602,605
                st_dev = introspective_iteration(model, real_batch, noise_batch, e5, hp, cur_lr_e, cur_lr_d)
                sh = stats_host[cur_iter % 2]
                sh.copy_(st_dev, non_blocking=True)                                 # the logged scalars (:628-639) + NaN flag
                ev = torch.cuda.Event()
                ev.record()
                # Default: the statistics are read right after the step, like the reference's .item() calls (:625-639).
                # SIVAE_ASYNC_STATS=1 reads those of step i once step i+1 is queued, so a slow host never exposes the next
                # launch behind the device->host sync; the NaN guard and the tqdm postfix then lag by one iteration, epoch
                # statistics stay complete (flushed at the end of the epoch).  Measured equal on this pool's hosts
                # (profiles/r01y_e2e_sweep.json), hence off.
                if pending is not None:
                    consume_stats(pending)
                pending = (sh, ev, epoch, pbar)
                if not async_stats:
                    consume_stats(pending)
                    pending = None
                if cur_iter % test_iter == 0 and main:
                    _, _, _, rec_det = model(real_batch, deterministic=True)
                    fake = model._engine.last_image(0)            # `fake` of the D half (:597), not a new forward
                    k = min(b_size, 16)
                    _save_grid([real_batch[:k], rec_det[:k], fake[:k]], '{}/image_{}.jpg'.format(fig_dir, cur_iter), num_row)
            cur_iter += 1
        if pending is not None:
            consume_stats(pending)
            pending = None
        pbar.close()
        if copy_to_target_freq is not None and epoch % copy_to_target_freq == 0:
            # bootstrap: the frozen target decoder follows the decoder, lagging by up to copy_to_target_freq epochs
            model.target_decoder.load_state_dict(model.decoder.state_dict())
        diff_kls = track.cur["diff_kl"]
        if exit_on_negative_diff and epoch > 50 and np.mean(diff_kls) < -1.0:
            print(f'the kl difference [{np.mean(diff_kls):.3f}] between fake and real is negative (no sampling improvement)')
            print("try to lower beta_neg hyperparameter")
            print("exiting...")
            raise SystemError("Negative KL Difference")
        if epoch > num_vae - 1:
            track.close_epoch()
        if epoch > num_vae - 1 and main:
            h = track.hist
            print('#' * 50)
            print(f'Epoch {epoch} Summary:')
            print(f'beta_rec: {beta_rec}, beta_kl: {beta_kl}, beta_neg: {beta_neg}')
            print(f'rec: {h["rec_err"][-1]:.3f}, kl: {h["kl_real"][-1]:.3f}, kl_fake: {h["kl_fake"][-1]:.3f}, kl_rec: {h["kl_rec"][-1]:.3f}')
            print(f'diff_kl: {np.mean(diff_kls):.3f}, exp_elbo_f: {h["exp_elbo_f"][-1]:.4e}, exp_elbo_r: {h["exp_elbo_r"][-1]:.4e}')
            print(f'time: {time.time() - start_time}')
            print('#' * 50)
        if epoch == num_epochs - 1 and real_batch is not None and main:
            with torch.no_grad():
                _, _, _, rec_det = model(real_batch, deterministic=True)
                noise_batch = torch.randn(size=(real_batch.size(0), z_dim)).to(device)
                fake = model.sample(noise_batch)
                k = min(real_batch.size(0), 16)
                _save_grid([real_batch[:k], rec_det[:k], fake[:k]], '{}/image_{}.jpg'.format(fig_dir, cur_iter), num_row)
            h = track.hist
            try:
                import matplotlib
                matplotlib.use('Agg')
                import matplotlib.pyplot as plt
                fig = plt.figure()
                ax = fig.add_subplot(1, 1, 1)
                for key, label in (("kl_real", "kl_real"), ("kl_fake", "kl_fake"), ("kl_rec", "kl_rec"), ("rec_err", "rec_err")):
                    ax.plot(np.arange(len(h[key])), h[key], label=label)
                ax.legend()
                plt.savefig('./soft_intro_train_graphs.jpg')
            except ImportError:
                print("matplotlib not available: skipping ./soft_intro_train_graphs.jpg")
            with open('./soft_intro_train_graphs_data.pickle', 'wb') as fp:
                pickle.dump({"kl_real": h["kl_real"], "kl_fake": h["kl_fake"], "kl_rec": h["kl_rec"], "rec_err": h["rec_err"]}, fp)
            save_checkpoint(model, epoch, cur_iter, prefix)
            model.train()
    _join_async_save()


if __name__ == '__main__':
    dev = torch.device("cuda:0")
    try:
        train_soft_intro_vae(dataset="synthetic32", z_dim=128, batch_size=32, num_workers=0, num_epochs=2, num_vae=0,
                             beta_kl=1.0, beta_neg=256, beta_rec=1.0, device=dev, save_interval=50, start_epoch=0,
                             lr_e=2e-4, lr_d=2e-4, pretrained=None, test_iter=1000, with_fid=False)
    except SystemError:
        print("Error, probably loss is NaN, try again...")
