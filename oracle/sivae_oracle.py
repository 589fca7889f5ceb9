"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement of the Soft-IntroVAE image hot path of the reference
(`soft_intro_vae/train_soft_intro_vae.py`, `soft_intro_vae_bootstrap/...`).  Every function
cites the reference lines it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this file.

The arithmetic of the reference lives in PyTorch (third party, pinned by the reference at
pytorch=1.7.0, `environment.yml:81`; runnable here only as torch 2.11).  This restatement
therefore uses plain `torch.nn.functional` CPU ops and autograd, written functionally over a
reference-format ``state_dict``, so it can run in fp32 (bit-comparable with the
reference module on the same thread count) or fp64 (the "exact" value the CUDA path is
measured against).

Parity pin: the reference has NO tests / golden vectors of its own, so the
pin is outputs of the unmodified reference run in the build container:
`oracle/make_golden.py` drives the reference's own `train_soft_intro_vae()` for one
teacher-forced iteration and commits the results under `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5          # nn.BatchNorm2d default, train_soft_intro_vae.py:58,62,90
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.2      # train_soft_intro_vae.py:59,63,91


# --------------------------------------------------------------------------------------
# architecture bookkeeping (index/shape work: must be bit-exact)
# --------------------------------------------------------------------------------------
@dataclass
class Arch:
    """Shape logic of Encoder/Decoder.__init__ (train_soft_intro_vae.py:79-109, 126-159)."""
    cdim: int = 3
    zdim: int = 512
    channels: Sequence[int] = (64, 128, 256, 512, 512, 512)
    image_size: int = 256

    def enc_blocks(self) -> List[Tuple[str, int, int, int]]:
        """[(module name, inc, outc, spatial size the block runs at)] -- :94-101."""
        out = []
        cc = self.channels[0]
        sz = self.image_size // 2
        for ch in self.channels[1:]:
            out.append(("res_in_{}".format(sz), cc, ch, sz))
            cc, sz = ch, sz // 2
        out.append(("res_in_{}".format(sz), cc, cc, sz))
        return out

    def conv_output_size(self) -> Tuple[int, int, int]:
        """calc_conv_output_size, :111-114 (shape of one sample after `main`)."""
        sz = self.image_size // 2
        for _ in self.channels[1:]:
            sz = sz // 2
        return (self.channels[-1], sz, sz)

    def dec_blocks(self) -> List[Tuple[str, int, int, int]]:
        """[(name, inc, outc, real spatial size)] -- :150-158.  Names count from 4 (`sz = 4`)
        regardless of the real size; the real size starts at conv_output_size."""
        out = []
        cc = self.channels[-1]
        name_sz = 4
        real = self.conv_output_size()[1]
        for ch in list(self.channels)[::-1]:
            out.append(("res_in_{}".format(name_sz), cc, ch, real))
            cc, name_sz, real = ch, name_sz * 2, real * 2
        out.append(("res_in_{}".format(name_sz), cc, cc, real))
        return out


def param_keys(sd: Dict[str, Tensor], prefix: str) -> List[str]:
    """Trainable tensors of one net in registration order (what optim.Adam(model.X.parameters())
    sees, :450-451): everything that is not a BN buffer."""
    return [k for k in sd if k.startswith(prefix) and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


def buffer_keys(sd: Dict[str, Tensor], prefix: str) -> List[str]:
    return [k for k in sd if k.startswith(prefix) and k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------
def _bn(sd, name: str, x: Tensor, train: bool) -> Tensor:
    """nn.BatchNorm2d forward; train mode uses batch stats (biased var) and updates the
    running buffers in place with momentum 0.1 and the UNBIASED variance."""
    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    if train:
        with torch.no_grad():
            sd[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], train, BN_MOMENTUM, BN_EPS)


def residual_block(sd, p: str, x: Tensor, train: bool) -> Tensor:
    """ResidualBlock.forward, train_soft_intro_vae.py:65-75."""
    if (p + ".conv_expand.weight") in sd:
        identity = F.conv2d(x, sd[p + ".conv_expand.weight"], None, 1, 0)
    else:
        identity = x
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1)
    out = F.leaky_relu(_bn(sd, p + ".bn1", out, train), LRELU_SLOPE)
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = _bn(sd, p + ".bn2", out, train)
    return F.leaky_relu(out + identity, LRELU_SLOPE)


def encoder_forward(sd, arch: Arch, x: Tensor, train: bool = True, prefix: str = "encoder",
                    o_cond: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Encoder.forward, :116-122 with `main` from :88-101.  o_cond: the condition rows a conditional model concatenates
    to the flattened features (:118-119; its fc is Linear(features + cond_dim, 2z), :106-107)."""
    m = prefix + ".main"
    y = F.conv2d(x, sd[m + ".0.weight"], None, 1, 2)
    y = F.leaky_relu(_bn(sd, m + ".1", y, train), LRELU_SLOPE)
    y = F.avg_pool2d(y, 2)
    blocks = arch.enc_blocks()
    for i, (name, _inc, _outc, _sz) in enumerate(blocks):
        y = residual_block(sd, m + "." + name, y, train)
        if i < len(blocks) - 1:
            y = F.avg_pool2d(y, 2)
    y = y.reshape(x.size(0), -1)
    if o_cond is not None:
        y = torch.cat([y, o_cond], dim=1)
    y = F.linear(y, sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"])
    mu, logvar = y.chunk(2, dim=1)
    return mu, logvar


def decoder_forward(sd, arch: Arch, z: Tensor, train: bool = True, prefix: str = "decoder",
                    y_cond: Optional[Tensor] = None) -> Tensor:
    """Decoder.forward, :161-169 with `fc`/`main` from :145-159.  y_cond: condition rows concatenated to z (:163-165)."""
    z = z.reshape(z.size(0), -1)
    if y_cond is not None:
        z = torch.cat([z, y_cond.reshape(y_cond.size(0), -1)], dim=1)
    y = F.relu(F.linear(z, sd[prefix + ".fc.0.weight"], sd[prefix + ".fc.0.bias"]))
    y = y.reshape(z.size(0), *arch.conv_output_size())
    m = prefix + ".main"
    blocks = arch.dec_blocks()
    for i, (name, _inc, _outc, _sz) in enumerate(blocks):
        y = residual_block(sd, m + "." + name, y, train)
        if i < len(blocks) - 1:
            y = F.interpolate(y, scale_factor=2, mode="nearest")
    return F.conv2d(y, sd[m + ".predict.weight"], sd[m + ".predict.bias"], 1, 2)


# --------------------------------------------------------------------------------------
# loss helpers
# --------------------------------------------------------------------------------------
def calc_kl(logvar: Tensor, mu: Tensor, reduce: str = "sum") -> Tensor:
    """calc_kl with mu_o = logvar_o = 0, :231-251."""
    kl = -0.5 * (1 + logvar - logvar.exp() - mu.pow(2)).sum(1)
    if reduce == "sum":
        kl = kl.sum()
    elif reduce == "mean":
        kl = kl.mean()
    return kl


def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """reparameterize, :254-265, with the N(0,1) draw supplied (teacher forcing)."""
    return mu + eps * torch.exp(0.5 * logvar)


def rec_loss(x: Tensor, recon_x: Tensor, reduction: str, loss_type: str = "mse") -> Tensor:
    """calc_reconstruction_loss, :268-294.  Arg order is (x, recon_x).  'mse' sums the squared error per sample BEFORE the
    batch reduction (:282-287); 'l1' / 'bce' hand `reduction` straight to F.l1_loss / F.binary_cross_entropy on the
    [B, D] views (:288-291), so 'mean' divides by B*D and 'none' is element-wise -- the callers of the 'none' form then
    sum the trailing dimensions away (`while len(shape) > 1: sum(-1)`, :574-578), which is folded in here."""
    b = recon_x.size(0)
    r, t = recon_x.reshape(b, -1), x.reshape(b, -1)
    if loss_type == "mse":
        err = (r - t).pow(2).sum(1)
        if reduction == "sum":
            err = err.sum()
        elif reduction == "mean":
            err = err.mean()
        return err
    if loss_type == "l1":
        err = F.l1_loss(r, t, reduction=reduction)
    elif loss_type == "bce":
        err = F.binary_cross_entropy(r, t, reduction=reduction)
    else:
        raise NotImplementedError
    while err.dim() > 1:
        err = err.sum(-1)
    return err


# --------------------------------------------------------------------------------------
# one introspective iteration
# --------------------------------------------------------------------------------------
@dataclass
class Hyper:
    beta_kl: float = 1.0
    beta_rec: float = 1.0
    beta_neg: float = 256.0
    gamma_r: float = 1e-8
    scale: float = 1.0 / (3 * 32 * 32)      # :456
    lr_e: float = 2e-4
    lr_d: float = 2e-4
    adam_b1: float = 0.9
    adam_b2: float = 0.999
    adam_eps: float = 1e-8
    loss_type: str = "mse"                  # recon_loss_type kwarg, :339


def _set_grad(sd, keys, flag: bool):
    for k in keys:
        sd[k].requires_grad_(flag)
        sd[k].grad = None


def e_step(sd, arch: Arch, real: Tensor, noise: Tensor, eps: Sequence[Tensor], hp: Hyper,
           bootstrap: bool = False):
    """Update-E half of the iteration, train_soft_intro_vae.py:551-589 (bootstrap :576-612:
    rec_rec / rec_fake come from `target_decoder`).  eps = (eps1, eps2, eps3) in draw order.
    Returns (scalars, grads of encoder params, z) -- z is reused by the D half (:598)."""
    ek, dk = param_keys(sd, "encoder."), param_keys(sd, "decoder.")
    _set_grad(sd, ek, True)
    _set_grad(sd, dk, False)
    tk = param_keys(sd, "target_decoder.") if bootstrap else []
    _set_grad(sd, tk, False)
    tgt = "target_decoder" if bootstrap else "decoder"

    fake = decoder_forward(sd, arch, noise)                                   # :557
    real_mu, real_logvar = encoder_forward(sd, arch, real)                    # :559
    z = reparameterize(real_mu, real_logvar, eps[0])                          # :560
    rec = decoder_forward(sd, arch, z)                                        # :561
    loss_rec = rec_loss(real, rec, "mean", hp.loss_type)                                    # :563
    lossE_real_kl = calc_kl(real_logvar, real_mu, "mean")                     # :565
    rec_mu, rec_logvar = encoder_forward(sd, arch, rec.detach())              # :567
    z_rec = reparameterize(rec_mu, rec_logvar, eps[1])
    rec_rec = decoder_forward(sd, arch, z_rec, prefix=tgt)
    fake_mu, fake_logvar = encoder_forward(sd, arch, fake.detach())           # :568
    z_fake = reparameterize(fake_mu, fake_logvar, eps[2])
    rec_fake = decoder_forward(sd, arch, z_fake, prefix=tgt)
    kl_rec = calc_kl(rec_logvar, rec_mu, "none")                              # :570
    kl_fake = calc_kl(fake_logvar, fake_mu, "none")                           # :571
    l_rr = rec_loss(rec, rec_rec, "none", hp.loss_type)          # :573 -- `rec` NOT detached (graph fidelity)
    l_rf = rec_loss(fake, rec_fake, "none", hp.loss_type)                                   # :576
    expelbo_rec = (-2 * hp.scale * (hp.beta_rec * l_rr + hp.beta_neg * kl_rec)).exp().mean()    # :580
    expelbo_fake = (-2 * hp.scale * (hp.beta_rec * l_rf + hp.beta_neg * kl_fake)).exp().mean()  # :581
    lossE_fake = 0.25 * (expelbo_rec + expelbo_fake)                          # :583
    lossE_real = hp.scale * (hp.beta_rec * loss_rec + hp.beta_kl * lossE_real_kl)  # :584
    lossE = lossE_real + lossE_fake
    lossE.backward()                                                          # :588
    grads = {k: sd[k].grad.detach().clone() for k in ek}
    scal = dict(loss_rec_e=loss_rec.item(), lossE_real_kl=lossE_real_kl.item(),
                expelbo_rec=expelbo_rec.item(), expelbo_fake=expelbo_fake.item(), lossE=lossE.item())
    tens = dict(fake=fake.detach(), rec=rec.detach(), rec_rec=rec_rec.detach(), rec_fake=rec_fake.detach(),
                real_mu=real_mu.detach(), real_logvar=real_logvar.detach(),
                rec_mu=rec_mu.detach(), rec_logvar=rec_logvar.detach(),
                fake_mu=fake_mu.detach(), fake_logvar=fake_logvar.detach())
    _set_grad(sd, ek, False)
    return scal, grads, z.detach(), tens


def d_step(sd, arch: Arch, real: Tensor, noise: Tensor, z: Tensor, eps: Sequence[Tensor], hp: Hyper,
           bootstrap: bool = False):
    """Update-D half, :591-624 (bootstrap :617-650: target decoder, nothing detached).
    eps = (eps4, eps5)."""
    ek, dk = param_keys(sd, "encoder."), param_keys(sd, "decoder.")
    _set_grad(sd, ek, False)
    _set_grad(sd, dk, True)
    tk = param_keys(sd, "target_decoder.") if bootstrap else []
    _set_grad(sd, tk, False)

    fake = decoder_forward(sd, arch, noise)                                   # :597
    rec = decoder_forward(sd, arch, z.detach())                               # :598
    loss_rec = rec_loss(real, rec, "mean", hp.loss_type)                                    # :599
    rec_mu, rec_logvar = encoder_forward(sd, arch, rec)                       # :601
    z_rec = reparameterize(rec_mu, rec_logvar, eps[0])
    fake_mu, fake_logvar = encoder_forward(sd, arch, fake)                    # :604
    z_fake = reparameterize(fake_mu, fake_logvar, eps[1])
    if bootstrap:
        rec_rec = decoder_forward(sd, arch, z_rec, prefix="target_decoder")   # bootstrap :635
        rec_fake = decoder_forward(sd, arch, z_fake, prefix="target_decoder")
        loss_rec_rec = rec_loss(rec, rec_rec, "mean", hp.loss_type)                         # bootstrap :638
        loss_fake_rec = rec_loss(fake, rec_fake, "mean", hp.loss_type)
    else:
        rec_rec = decoder_forward(sd, arch, z_rec.detach())                   # :607
        rec_fake = decoder_forward(sd, arch, z_fake.detach())                 # :608
        loss_rec_rec = rec_loss(rec.detach(), rec_rec, "mean", hp.loss_type)                # :610
        loss_fake_rec = rec_loss(fake.detach(), rec_fake, "mean", hp.loss_type)             # :612
    lossD_rec_kl = calc_kl(rec_logvar, rec_mu, "mean")                        # :615
    lossD_fake_kl = calc_kl(fake_logvar, fake_mu, "mean")                     # :616
    lossD = hp.scale * (loss_rec * hp.beta_rec + (lossD_rec_kl + lossD_fake_kl) * 0.5 * hp.beta_kl
                        + hp.gamma_r * 0.5 * hp.beta_rec * (loss_rec_rec + loss_fake_rec))   # :618-620
    lossD.backward()                                                          # :623
    grads = {k: sd[k].grad.detach().clone() for k in dk}
    scal = dict(loss_rec=loss_rec.item(), lossD_rec_kl=lossD_rec_kl.item(), lossD_fake_kl=lossD_fake_kl.item(),
                loss_rec_rec=loss_rec_rec.item(), loss_fake_rec=loss_fake_rec.item(), lossD=lossD.item())
    tens = dict(fake=fake.detach(), rec=rec.detach(), rec_rec=rec_rec.detach(), rec_fake=rec_fake.detach(),
                rec_mu=rec_mu.detach(), rec_logvar=rec_logvar.detach(),
                fake_mu=fake_mu.detach(), fake_logvar=fake_logvar.detach())
    _set_grad(sd, dk, False)
    return scal, grads, tens


def vae_step(sd, arch: Arch, real: Tensor, eps: Tensor, hp: Hyper, bootstrap: bool = False):
    """Vanilla-VAE warm-up step, :512-540 (loss NOT multiplied by `scale`).  Bootstrap trainer (:540-564 there): the same
    code, but `model(real_batch)` decodes with the frozen target decoder (forward(..., target=True) is the default,
    bootstrap :196-217), so the trainable decoder gets no gradient (gd comes back empty) and optimizer_d.step() is a no-op."""
    ek, dk = param_keys(sd, "encoder."), param_keys(sd, "decoder.")
    _set_grad(sd, ek, True)
    _set_grad(sd, dk, not bootstrap)
    mu, logvar = encoder_forward(sd, arch, real)
    z = reparameterize(mu, logvar, eps)
    rec = decoder_forward(sd, arch, z, prefix="target_decoder" if bootstrap else "decoder")
    loss_rec = rec_loss(real, rec, "mean", hp.loss_type)
    loss_kl = calc_kl(logvar, mu, "mean")
    loss = hp.beta_rec * loss_rec + hp.beta_kl * loss_kl
    loss.backward()
    ge = {k: sd[k].grad.detach().clone() for k in ek}
    gd = {} if bootstrap else {k: sd[k].grad.detach().clone() for k in dk}
    _set_grad(sd, ek, False)
    _set_grad(sd, dk, False)
    return dict(loss_rec=loss_rec.item(), loss_kl=loss_kl.item(), loss=loss.item()), ge, gd


# --------------------------------------------------------------------------------------
# optimiser
# --------------------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, Tensor] = field(default_factory=dict)
    v: Dict[str, Tensor] = field(default_factory=dict)


def adam_update(sd, grads: Dict[str, Tensor], st: AdamState, lr: float, hp: Hyper):
    """torch.optim.Adam single-tensor step (betas .9/.999, eps 1e-8, no weight decay, no amsgrad),
    as constructed at :450-451 and stepped at :589/:624."""
    st.step += 1
    b1, b2 = hp.adam_b1, hp.adam_b2
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    with torch.no_grad():
        for k, g in grads.items():
            if k not in st.m:
                st.m[k] = torch.zeros_like(sd[k])
                st.v[k] = torch.zeros_like(sd[k])
            st.m[k].lerp_(g, 1 - b1)
            st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (st.v[k].sqrt() / math.sqrt(bc2)).add_(hp.adam_eps)
            sd[k].addcdiv_(st.m[k], denom, value=-(lr / bc1))


def clone_sd(sd: Dict[str, Tensor], dtype: Optional[torch.dtype] = None) -> Dict[str, Tensor]:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        out[k] = t
    return out


def full_iteration(sd, arch: Arch, real, noise, eps5, hp: Hyper, st_e: AdamState, st_d: AdamState,
                   bootstrap: bool = False, world_grads_e=None, world_grads_d=None):
    """E half + Adam(encoder) + D half + Adam(decoder), exactly the order of :551-624."""
    se, ge, z, te = e_step(sd, arch, real, noise, eps5[:3], hp, bootstrap)
    adam_update(sd, ge if world_grads_e is None else world_grads_e, st_e, hp.lr_e, hp)
    sdd, gd, td = d_step(sd, arch, real, noise, z, eps5[3:], hp, bootstrap)
    adam_update(sd, gd if world_grads_d is None else world_grads_d, st_d, hp.lr_d, hp)
    scal = dict(se)
    scal.update(sdd)
    return scal, ge, gd, te, td


# --------------------------------------------------------------------------------------
# random-init state_dict of the reference architecture without importing the reference
# --------------------------------------------------------------------------------------
def make_state_dict(arch: Arch, seed: int = 0, bootstrap: bool = False) -> Dict[str, Tensor]:
    """Builds the tensors of SURVEY App. B with torch.nn default initialisers consumed in the
    reference's construction order (encoder modules in definition order :88-109, then decoder
    :140-159 [, then target decoder]), so that for the same seed the values are bit-identical to
    `SoftIntroVAE(...)` of the reference.  Reproduces the constructor side effect of the
    train-mode dummy forward (:102,111-114): encoder BN running_var = 0.9, nbt = 1."""
    import torch.nn as nn
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cin, cout, k, bias):
        c = nn.Conv2d(cin, cout, k, 1, k // 2, bias=bias)
        sd[name + ".weight"] = c.weight.detach().clone()
        if bias:
            sd[name + ".bias"] = c.bias.detach().clone()

    def bn(name, c, enc):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.full((c,), 0.9 if enc else 1.0)
        sd[name + ".num_batches_tracked"] = torch.tensor(1 if enc else 0, dtype=torch.long)

    def block(p, inc, outc, enc):
        if inc != outc:
            conv(p + ".conv_expand", inc, outc, 1, False)
        conv(p + ".conv1", inc, outc, 3, False)
        bn(p + ".bn1", outc, enc)
        conv(p + ".conv2", outc, outc, 3, False)
        bn(p + ".bn2", outc, enc)

    def linear(name, fin, fout):
        l = nn.Linear(fin, fout)
        sd[name + ".weight"] = l.weight.detach().clone()
        sd[name + ".bias"] = l.bias.detach().clone()

    conv("encoder.main.0", arch.cdim, arch.channels[0], 5, False)
    bn("encoder.main.1", arch.channels[0], True)
    for name, inc, outc, _ in arch.enc_blocks():
        block("encoder.main." + name, inc, outc, True)
    C, h, w = arch.conv_output_size()
    linear("encoder.fc", C * h * w, 2 * arch.zdim)
    for pref in (["decoder", "target_decoder"] if bootstrap else ["decoder"]):
        linear(pref + ".fc.0", arch.zdim, C * h * w)
        for name, inc, outc, _ in arch.dec_blocks():
            block(pref + ".main." + name, inc, outc, False)
        conv(pref + ".main.predict", arch.channels[0], arch.cdim, 5, True)
    torch.random.set_rng_state(g)
    return sd
