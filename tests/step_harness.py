"""Shared harness: run one teacher-forced introspective iteration (E half, Adam, D half, Adam) through the engine's
C ABI and through the oracle from identical state and inputs, and compare.  Used by the -m gpu parity tests and by
__graft_entry__.smoke().  (Test infrastructure: the only place besides bench.py's baseline legs that touches oracle/.)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "soft-intro-vae-pytorch_b200"


def mods(bootstrap=False):
    L = importlib.import_module(PKG + ".lib")
    name = ".train_soft_intro_vae_bootstrap" if bootstrap else ".train_soft_intro_vae"
    return L, importlib.import_module(PKG + name)


def make_inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    real = torch.rand(batch, cfg["cdim"], cfg["image_size"], cfg["image_size"], generator=g)
    noise = torch.randn(batch, cfg["zdim"], generator=g)
    eps = torch.randn(5, batch, cfg["zdim"], generator=g)
    return real, noise, eps


DEFAULT_HP = dict(beta_kl=1.0, beta_rec=1.0, beta_neg=256.0, gamma_r=1e-8, lr_e=2e-4, lr_d=2e-4, loss_type="mse")


def run_engine_iteration(cfg, batch, seed, backend=0, bootstrap=False, init_sd=None, inputs=None, hp=None, device="cuda:0",
                         teacher_enc=None):
    L, M = mods(bootstrap)
    E = importlib.import_module(PKG + ".engine")
    hp = dict(DEFAULT_HP, **(hp or {}))
    if "scale" not in hp:
        hp["scale"] = 1.0 / (cfg["cdim"] * cfg["image_size"] ** 2)
    torch.manual_seed(seed)
    model = M.SoftIntroVAE(cdim=cfg["cdim"], zdim=cfg["zdim"], channels=cfg["channels"], image_size=cfg["image_size"])
    if init_sd is not None:
        model.load_state_dict(init_sd)
    model._conv_backend = backend
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(device)
    real, noise, eps = inputs if inputs is not None else make_inputs(cfg, batch, seed)
    real, noise, eps = real.to(device), noise.to(device), eps.to(device).contiguous()
    eng = model.reserve(batch)
    eng.recon_loss = hp["loss_type"]
    h = E.make_hyper(hp["beta_kl"], hp["beta_rec"], hp["beta_neg"], hp["gamma_r"], hp["scale"])
    eng.e_step(real, noise, eps[:3].contiguous(), h)
    torch.cuda.synchronize()
    grads_e = {"encoder." + n: p.grad.detach().clone().cpu().contiguous() for n, p in model.encoder.named_parameters()}
    imgs_e = {k: eng.last_image(i).cpu() for i, k in enumerate(["fake", "rec", "rec_rec", "rec_fake"])}
    eng.adam(L.NET_ENCODER, hp["lr_e"])
    torch.cuda.synchronize()
    enc_after_e = {"encoder." + k: v.detach().clone().cpu().contiguous() for k, v in model.encoder.state_dict().items()}
    if teacher_enc is not None:
        # teacher forcing of the D half: Adam's first step moves every weight by ~lr*sign(g), which turns round-off
        # in tiny gradients into O(lr) weight differences; loading the comparison run's post-E-step encoder weights
        # makes the D half a function of identical state (the free-running variant is tested separately).
        with torch.no_grad():
            for n, p in model.encoder.named_parameters():
                p.copy_(teacher_enc["encoder." + n].to(device=device, dtype=torch.float32))
        model.params_changed()
    eng.d_step(eps[3:].contiguous(), h)
    torch.cuda.synchronize()
    grads_d = {"decoder." + n: p.grad.detach().clone().cpu().contiguous() for n, p in model.decoder.named_parameters()}
    eng.adam(L.NET_DECODER, hp["lr_d"])
    torch.cuda.synchronize()
    st = eng.stats.cpu()
    scal = dict(loss_rec_e=st[0].item(), lossE_real_kl=st[1].item(), expelbo_rec=st[2].item(), expelbo_fake=st[3].item(),
                lossE=st[4].item(), loss_rec=st[5].item(), lossD_rec_kl=st[6].item(), lossD_fake_kl=st[7].item(),
                loss_rec_rec=st[8].item(), loss_fake_rec=st[9].item(), lossD=st[10].item(), nan=st[15].item(),
                bce_domain=st[14].item())
    post = {k: v.detach().clone().cpu().contiguous() for k, v in model.state_dict().items()}
    return dict(scalars=scal, grads_e=grads_e, grads_d=grads_d, post=post, init=init, images_e=imgs_e, model=model,
                enc_after_e=enc_after_e)


def run_oracle_iteration(cfg, batch, seed, bootstrap=False, init_sd=None, inputs=None, hp=None, dtype=torch.float64):
    from oracle import sivae_oracle as O
    hp = dict(DEFAULT_HP, **(hp or {}))
    if "scale" not in hp:
        hp["scale"] = 1.0 / (cfg["cdim"] * cfg["image_size"] ** 2)
    arch = O.Arch(cdim=cfg["cdim"], zdim=cfg["zdim"], channels=cfg["channels"], image_size=cfg["image_size"])
    sd = O.clone_sd(init_sd if init_sd is not None else O.make_state_dict(arch, seed=seed, bootstrap=bootstrap), dtype)
    real, noise, eps = inputs if inputs is not None else make_inputs(cfg, batch, seed)
    real, noise, eps = real.to(dtype), noise.to(dtype), [e.to(dtype) for e in eps]
    ohp = O.Hyper(beta_kl=hp["beta_kl"], beta_rec=hp["beta_rec"], beta_neg=hp["beta_neg"], gamma_r=hp["gamma_r"],
                  scale=hp["scale"], lr_e=hp["lr_e"], lr_d=hp["lr_d"], loss_type=hp["loss_type"])
    scal, ge, gd, te, td = O.full_iteration(sd, arch, real, noise, eps, ohp, O.AdamState(), O.AdamState(), bootstrap)
    return dict(scalars=scal, grads_e=ge, grads_d=gd, post=sd, images_e=te)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def compare(eng, ora, tol, label="", lr=2e-4, verbose=True, tensor_tol=None, noise=None):
    """tol: relative tolerance for scalars; tensors (gradients, BN statistics) get 10*tol as relative-L2 / max-rel.
    All deviations are measured first (and appended to gpurun_out/parity_report.jsonl), then asserted."""
    import json
    ttol = 10 * tol if tensor_tol is None else tensor_tol
    # `noise`: the same oracle run in fp32.  Some seeds are ill-conditioned (the fp32 reference itself is 2e-3 away from
    # fp64 on the CIFAR-shaped case); a gradient tensor may deviate from the fp64 truth by 3x the reference's own fp32
    # round-off before it counts as a failure.
    def floor(name, k):
        if noise is None or k not in noise.get(name, {}):
            return 0.0
        return 3.0 * rel_l2(noise[name][k], ora[name][k])
    dev, fails = {}, []
    es, os_ = eng["scalars"], ora["scalars"]
    if es.get("nan", 0.0) != 0.0:
        fails.append("NaN flag set")
    for k in ("loss_rec_e", "lossE_real_kl", "expelbo_rec", "expelbo_fake", "lossE", "loss_rec", "lossD_rec_kl",
              "lossD_fake_kl", "loss_rec_rec", "loss_fake_rec", "lossD"):
        if k not in os_:
            continue
        r = abs(es[k] - os_[k]) / (abs(os_[k]) + 1e-30)
        dev["scalar:" + k] = r
        if not r < tol:
            fails.append("scalar %s engine %.8g oracle %.8g rel %.3g > %.3g" % (k, es[k], os_[k], r, tol))
    for name in ("grads_e", "grads_d"):
        if set(eng[name]) != set(ora[name]):
            fails.append(name + ": key sets differ")
            continue
        for k in ora[name]:
            r = rel_l2(eng[name][k], ora[name][k])
            dev[name + ":" + k] = r
            if not r < max(ttol, floor(name, k)):
                fails.append("%s[%s] rel-L2 %.3g > %.3g" % (name, k, r, max(ttol, floor(name, k))))
    for k, v in ora["post"].items():
        e = eng["post"][k]
        if k.endswith("num_batches_tracked"):
            if int(e) != int(v):                                   # integer work: exact
                fails.append("%s %d != %d" % (k, int(e), int(v)))
        elif k.endswith(("running_mean", "running_var")):
            r = float((e.double() - v.double()).abs().max() / (v.double().abs().max() + 1e-12))
            dev["bn:" + k] = r
            if not r < ttol:
                fails.append("%s max-rel %.3g" % (k, r))
        else:
            # against the ORACLE only a bound holds: one Adam step from zero moments moves each weight by ~lr*sign(g), so
            # round-off in a tiny gradient flips whole steps
            d = float((e.double() - v.double()).abs().max())
            if not d <= 2.05 * lr + 1e-7:
                fails.append("post-step %s differs by %.3g" % (k, d))
    # the optimiser step itself, teacher-forced with the ENGINE's own gradient: after the first Adam step (m = (1-b1) g,
    # v = (1-b2) g^2, both bias-corrected back to g and g^2) every parameter must sit at p0 - lr * g / (|g| + eps)
    if "init" in eng:
        for name, after in (("grads_e", eng.get("enc_after_e")), ("grads_d", eng["post"])):
            if after is None:
                continue
            for k, g in eng[name].items():
                g64 = g.double()
                want = eng["init"][k].double() - lr * g64 / (g64.abs() + 1e-8)
                d = float((after[k].double() - want).abs().max())
                dev["adam:" + k] = d / lr
                if not d <= 2e-7 + 1e-3 * lr:
                    fails.append("Adam step of %s: |p - (p0 - lr g/(|g|+eps))| = %.3g" % (k, d))
    top = sorted(dev.items(), key=lambda kv: -kv[1])
    if verbose:
        sc = [kv for kv in top if kv[0].startswith("scalar:")]
        te = [kv for kv in top if not kv[0].startswith(("scalar:", "adam:"))]
        print("[%s] worst scalars: %s | worst tensors: %s" % (label, ", ".join("%s=%.2e" % (k[7:], v) for k, v in sc[:3]),
                                                               ", ".join("%s=%.2e" % kv for kv in te[:4])))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(label=label, tol=tol, fails=fails, dev=dict(top))) + "\n")
    except OSError:
        pass
    assert not fails, "%s: %d deviations over tolerance; first: %s" % (label, len(fails), "; ".join(fails[:4]))
    return dev
