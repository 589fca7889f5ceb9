"""
ORACLE TOOLING -- TEST INFRASTRUCTURE ONLY (build container only; needs /root/reference).

Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
`train_soft_intro_vae()` (soft_intro_vae/train_soft_intro_vae.py:337-702, and the bootstrap twin)
for exactly one introspective iteration on CPU, and recording everything at its boundary:

  * the model's state_dict right after construction (SoftIntroVAE ctor, :442)
  * the real batch, the `noise_batch` draw (:547) and the five reparameterisation draws
    (:560,567,568,602,605) -- captured by wrapping `torch.randn` / `ref.reparameterize`
  * encoder grads at `optimizer_e.step()` (:589) and decoder grads at `optimizer_d.step()` (:624)
    -- captured by a recording subclass of optim.Adam
  * the logged scalars (tqdm postfix :629-631 and the pickle :695-697)
  * the state_dict right after optimizer_d.step() (snapshotted inside the tqdm `set_postfix` call, :629)

No reference source is copied: the reference module is imported from /root/reference and only its
module globals are patched (dataset class, matplotlib stub, model hyper-parameters for the tiny
config).  Usage:  python oracle/make_golden.py   (writes every tests/golden/*.pt, ~25 MB total), or one group:
  --losses   tiny_l1.pt, tiny_bce.pt       the same iteration with recon_loss_type = 'l1' / 'bce' (:288-291)
  --cond     tiny_cond.pt                  SoftIntroVAE(conditional=True) train / eval forward with a one-hot condition
  --vae      tiny_vae_std.pt, tiny_vae_bootstrap.pt   the VAE warm-up iteration (epoch < num_vae, :512-540) of both trainers
  --helpers  helpers.pt                    calc_kl / calc_reconstruction_loss / reparameterize called directly
  --toy      toy2d.pt                      the 2-D trainer's printed log (BASELINE config 1)
(the image-loader fixtures come from oracle/make_image_golden.py)
"""
import hashlib
import os
import pickle
import sys
import tempfile
from unittest.mock import MagicMock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference(subdir: str, modname: str):
    """Import the reference trainer unmodified (container lacks matplotlib; no network)."""
    mpl, plt = MagicMock(), MagicMock()
    plt.subplots = lambda *a, **k: (MagicMock(), MagicMock())
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    mpl.pyplot = plt
    for m in ("dataset", "metrics", "metrics.fid_score", "metrics.inception", modname):
        sys.modules.pop(m, None)
    sys.path.insert(0, os.path.join(REF, subdir))
    try:
        mod = __import__(modname)
    finally:
        sys.path.pop(0)
    return mod


def run_reference_iteration(subdir, modname, fn_name, *, tiny, batch, z_dim, seed, beta_neg, extra_kwargs=None,
                            data_seed=1234, image_size=32, init_edit=None):
    """init_edit(model): optional in-place edit of the freshly constructed reference model (recorded as part of `init`);
    used for recon_loss_type='bce', whose F.binary_cross_entropy needs decoder outputs inside [0, 1] (:291)."""
    ref = _import_reference(subdir, modname)
    rec = {"eps": [], "noise": [], "grads": [], "postfix": None}

    class FakeCIFAR(torch.utils.data.Dataset):       # ref.CIFAR10 is looked up as a module global (:379)
        def __init__(self, root=None, train=True, download=False, transform=None):
            self.x = torch.rand(batch, 3, image_size, image_size, generator=torch.Generator().manual_seed(data_seed))

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], 0

    ref.CIFAR10 = FakeCIFAR

    holder = {}
    orig_init = ref.SoftIntroVAE.__init__           # the class itself stays the reference's

    def patched_init(self, cdim=3, zdim=512, channels=(64, 128, 256, 512, 512, 512), image_size=256, **kw):
        if tiny is not None:                         # only the architecture hyper-parameters are substituted
            channels, image_size = tiny["channels"], tiny["image_size"]
        orig_init(self, cdim=cdim, zdim=zdim, channels=channels, image_size=image_size, **kw)
        holder["model"] = self
        if init_edit is not None:
            with torch.no_grad():
                init_edit(self)
        holder["init"] = {k: v.detach().clone() for k, v in self.state_dict().items()}
        holder["arch"] = dict(cdim=cdim, zdim=zdim, channels=list(channels), image_size=image_size)

    ref.SoftIntroVAE.__init__ = patched_init

    def reparameterize(mu, logvar):                 # same maths as :254-265, draw recorded
        eps = torch.randn_like(logvar)
        rec["eps"].append(eps.detach().clone())
        return mu + eps * torch.exp(0.5 * logvar)

    ref.reparameterize = reparameterize

    orig_randn = torch.randn

    def randn(*a, **k):
        t = orig_randn(*a, **k)
        if k.get("size", None) is not None and tuple(k["size"]) == (batch, z_dim):
            rec["noise"].append(t.detach().clone())
        return t

    class RecAdam(torch.optim.Adam):
        def step(self, closure=None):
            g = {}
            for grp in self.param_groups:
                for p in grp["params"]:
                    g[id(p)] = None if p.grad is None else p.grad.detach().clone()
            rec["grads"].append(g)
            return super().step(closure)

    ref.optim.Adam = RecAdam

    class Bar:                                       # stands in for tqdm(iterable=loader), :506
        def __init__(self, iterable=None, **k):
            self.it = iterable

        def __iter__(self):
            for b in self.it:                        # the loader shuffles (:458): record the batch actually seen
                rec["real"] = b[0].detach().clone()
                yield b

        def set_description_str(self, *a, **k):
            pass

        def set_postfix(self, **k):
            # called at :629-631, i.e. after optimizer_d.step() (:624) and BEFORE the iteration-0 sample figure
            # (:641-646, a train-mode forward that also moves the BN running stats) -- snapshot here.
            rec["postfix"] = dict(k)
            holder["post"] = {n: v.detach().clone() for n, v in holder["model"].state_dict().items()}

        def close(self):
            pass

    ref.tqdm = Bar

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="sivae_golden_")
    os.chdir(tmp)
    torch.set_num_threads(8)
    try:
        torch.randn = randn
        kwargs = dict(dataset="cifar10", z_dim=z_dim, batch_size=batch, num_workers=0, num_epochs=1, num_vae=0,
                      beta_kl=1.0, beta_neg=beta_neg, beta_rec=1.0, device=torch.device("cpu"), seed=seed,
                      test_iter=10 ** 9, save_interval=50, start_epoch=0, lr_e=2e-4, lr_d=2e-4)
        kwargs.update(extra_kwargs or {})
        getattr(ref, fn_name)(**kwargs)
    finally:
        torch.randn = orig_randn
        ref.optim.Adam = torch.optim.Adam
        os.chdir(cwd)
    with open(os.path.join(tmp, "soft_intro_train_graphs_data.pickle"), "rb") as fp:
        graphs = pickle.load(fp)
    model = holder["model"]
    names = {id(p): n for n, p in model.named_parameters()}
    assert len(rec["grads"]) == 2, len(rec["grads"])
    grads_e = {names[i]: g for i, g in rec["grads"][0].items() if g is not None}
    grads_d = {names[i]: g for i, g in rec["grads"][1].items() if g is not None}
    post = holder["post"]
    # the first (batch, z) torch.randn call is noise_batch (:547); a second one happens at the end-of-run
    # sample grid (:679) and is not part of the iteration.
    out = dict(
        arch=holder["arch"], batch=batch, seed=seed, data_seed=data_seed,
        hyper=dict(beta_kl=1.0, beta_rec=1.0, beta_neg=float(beta_neg), gamma_r=float(kwargs.get("gamma_r", 1e-8 if "bootstrap" not in modname else 1.0)),
                   scale=1.0 / (3 * 32 ** 2), lr_e=2e-4, lr_d=2e-4, loss_type=kwargs.get("recon_loss_type", "mse")),
        real=rec["real"], noise=rec["noise"][0], eps=rec["eps"][:5],
        init=holder["init"], post=post, grads_e=grads_e, grads_d=grads_d,
        scalars=dict(r_loss=rec["postfix"]["r_loss"], kl=rec["postfix"]["kl"], diff_kl=rec["postfix"]["diff_kl"],
                     expelbo_f=rec["postfix"]["expelbo_f"],
                     kl_real=float(graphs["kl_real"][0]), kl_fake=float(graphs["kl_fake"][0]),
                     kl_rec=float(graphs["kl_rec"][0]), rec_err=float(graphs["rec_err"][0])),
        torch_version=torch.__version__, threads=torch.get_num_threads(),
        reference_commit="b6dbf16",
    )
    assert len(rec["eps"]) >= 5
    return out


def summarise(full):
    """Large-config fixture: keep inputs that cannot be regenerated from a seed (noise, eps), the scalars, and
    per-tensor fingerprints of init / grads / post-step state instead of the tensors themselves."""
    def fp(d):
        return {k: (float(v.double().sum()), float(v.double().abs().sum()), float(v.double().norm()))
                for k, v in d.items() if v.is_floating_point()}
    out = {k: full[k] for k in ("arch", "batch", "seed", "data_seed", "hyper", "real", "noise", "eps", "scalars",
                                "torch_version", "threads", "reference_commit")}
    out["init_fp"] = fp(full["init"])
    out["post_fp"] = fp(full["post"])
    out["grads_e_fp"] = fp(full["grads_e"])
    out["grads_d_fp"] = fp(full["grads_d"])
    out["nbt_post"] = {k: int(v) for k, v in full["post"].items() if k.endswith("num_batches_tracked")}
    # a few full tensors that are small: fc biases and BN stats of the last encoder block
    out["post_small"] = {k: v for k, v in full["post"].items() if v.numel() <= 512}
    out["grads_small"] = {k: v for k, v in list(full["grads_e"].items()) + list(full["grads_d"].items()) if v.numel() <= 512}
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    tiny = dict(channels=[32, 64], image_size=16)
    g = run_reference_iteration("soft_intro_vae", "train_soft_intro_vae", "train_soft_intro_vae",
                                tiny=tiny, batch=8, z_dim=16, seed=0, beta_neg=256, image_size=16)
    torch.save(g, os.path.join(OUT, "tiny_std.pt"))
    print("tiny_std", g["scalars"])
    g = run_reference_iteration("soft_intro_vae_bootstrap", "train_soft_intro_vae_bootstrap", "train_soft_intro_vae",
                                tiny=tiny, batch=8, z_dim=16, seed=0, beta_neg=256, image_size=16,
                                extra_kwargs=dict(gamma_r=1.0, copy_to_target_freq=1))
    torch.save(g, os.path.join(OUT, "tiny_bootstrap.pt"))
    print("tiny_bootstrap", g["scalars"])
    g = run_reference_iteration("soft_intro_vae", "train_soft_intro_vae", "train_soft_intro_vae",
                                tiny=None, batch=8, z_dim=128, seed=0, beta_neg=256, image_size=32)
    torch.save(summarise(g), os.path.join(OUT, "cifar_std_summary.pt"))
    print("cifar_std", g["scalars"])
    extra_losses()
    conditional_forward()
    vae_warmups()
    helper_functions()
    for f in sorted(os.listdir(OUT)):
        p = os.path.join(OUT, f)
        print(f, os.path.getsize(p), hashlib.sha256(open(p, "rb").read()).hexdigest()[:16])


def _unit_range_decoder(model):
    """decoder outputs inside (0, 1): small `predict` filters around a bias of 0.5 (for the bce golden)"""
    model.decoder.main.predict.weight.mul_(0.05)
    model.decoder.main.predict.bias.fill_(0.5)


def extra_losses():
    """recon_loss_type = 'l1' / 'bce' (:288-291) at the tiny architecture, same seeds as tiny_std"""
    tiny = dict(channels=[32, 64], image_size=16)
    for lt, edit in (("l1", None), ("bce", _unit_range_decoder)):
        g = run_reference_iteration("soft_intro_vae", "train_soft_intro_vae", "train_soft_intro_vae",
                                    tiny=tiny, batch=8, z_dim=16, seed=0, beta_neg=256, image_size=16,
                                    extra_kwargs=dict(recon_loss_type=lt), init_edit=edit)
        torch.save(g, os.path.join(OUT, "tiny_%s.pt" % lt))
        print("tiny_" + lt, g["scalars"])


def conditional_forward(seed=0, batch=6, cond_dim=10):
    """SoftIntroVAE(conditional=True) (:106-109, :139-143, :186-193): the UNMODIFIED reference model's train-mode and eval-mode
    forward with a one-hot condition at the tiny architecture -- the only use the reference has for the branch (its training
    step passes no condition)."""
    ref = _import_reference("soft_intro_vae", "train_soft_intro_vae")
    torch.set_num_threads(8)
    torch.manual_seed(seed)
    arch = dict(cdim=3, zdim=16, channels=[32, 64], image_size=16)
    model = ref.SoftIntroVAE(conditional=True, cond_dim=cond_dim, **arch)
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    x = torch.rand(batch, 3, 16, 16, generator=g)
    z = torch.randn(batch, 16, generator=g)
    cond = torch.nn.functional.one_hot(torch.randint(0, cond_dim, (batch,), generator=g), cond_dim).float()
    out = dict(arch=arch, cond_dim=cond_dim, seed=seed, init=init, x=x, z=z, cond=cond, torch_version=torch.__version__,
               reference_commit="b6dbf16")
    with torch.no_grad():
        model.train()
        mu, lv, _, y = model(x, o_cond=cond, deterministic=True)
        out["train"] = dict(mu=mu.clone(), logvar=lv.clone(), y=y.clone(), sample=model.sample(z, y_cond=cond).clone())
        out["post_train"] = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
        model.eval()
        mu, lv, _, y = model(x, o_cond=cond, deterministic=True)
        out["eval"] = dict(mu=mu.clone(), logvar=lv.clone(), y=y.clone(), sample=model.sample(z, y_cond=cond).clone())
        try:
            model(x)                      # conditional model without a condition: fc shape error (:120)
            out["uncond_error"] = None
        except RuntimeError as ex:
            out["uncond_error"] = str(ex)[:120]
    torch.save(out, os.path.join(OUT, "tiny_cond.pt"))
    print("tiny_cond", float(out["train"]["mu"].abs().sum()), float(out["eval"]["y"].abs().sum()), out["uncond_error"])


def vae_warmup(subdir, modname, fn_name, out_name, extra_kwargs=None, batch=8, z_dim=16, seed=0):
    """VAE warm-up iteration (`epoch < num_vae`, :512-540; bootstrap twin: decodes through the frozen target decoder) of the
    UNMODIFIED reference trainer at the tiny architecture: num_vae = 1, num_epochs = 2 (with num_epochs = 1 the reference dies of a
    NameError after the epoch -- `b_size` is only assigned in the introspective branch, :547 vs :678), everything recorded at the
    FIRST tqdm `set_postfix` call, i.e. right after optimizer_e.step() / optimizer_d.step() of the warm-up iteration (:533-537)."""
    ref = _import_reference(subdir, modname)
    tiny = dict(channels=[32, 64], image_size=16)
    rec = {"eps": [], "grads": [], "postfix": None}
    holder = {}

    class FakeCIFAR(torch.utils.data.Dataset):
        def __init__(self, root=None, train=True, download=False, transform=None):
            self.x = torch.rand(batch, 3, 16, 16, generator=torch.Generator().manual_seed(1234))

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], 0

    ref.CIFAR10 = FakeCIFAR
    orig_init = ref.SoftIntroVAE.__init__

    def patched_init(self, cdim=3, zdim=512, channels=(64, 128, 256, 512, 512, 512), image_size=256, **kw):
        orig_init(self, cdim=cdim, zdim=zdim, channels=tiny["channels"], image_size=tiny["image_size"], **kw)
        holder["model"] = self
        holder["init"] = {k: v.detach().clone() for k, v in self.state_dict().items()}
        holder["arch"] = dict(cdim=cdim, zdim=zdim, channels=list(tiny["channels"]), image_size=tiny["image_size"])

    ref.SoftIntroVAE.__init__ = patched_init

    def reparameterize(mu, logvar):
        eps = torch.randn_like(logvar)
        rec["eps"].append(eps.detach().clone())
        return mu + eps * torch.exp(0.5 * logvar)

    ref.reparameterize = reparameterize

    class RecAdam(torch.optim.Adam):
        def step(self, closure=None):
            g = {}
            for grp in self.param_groups:
                for p in grp["params"]:
                    g[id(p)] = None if p.grad is None else p.grad.detach().clone()
            rec["grads"].append(g)
            return super().step(closure)

    ref.optim.Adam = RecAdam

    class Bar:
        def __init__(self, iterable=None, **k):
            self.it = iterable

        def __iter__(self):
            for b in self.it:
                if "real" not in rec:
                    rec["real"] = b[0].detach().clone()
                yield b

        def set_description_str(self, *a, **k):
            pass

        def set_postfix(self, **k):
            if rec["postfix"] is None:
                rec["postfix"] = dict(k)
                holder["post"] = {n: v.detach().clone() for n, v in holder["model"].state_dict().items()}

        def close(self):
            pass

    ref.tqdm = Bar
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="sivae_golden_vae_")
    os.chdir(tmp)
    torch.set_num_threads(8)
    try:
        kwargs = dict(dataset="cifar10", z_dim=z_dim, batch_size=batch, num_workers=0, num_epochs=2, num_vae=1, beta_kl=0.7,
                      beta_neg=256, beta_rec=1.3, device=torch.device("cpu"), seed=seed, test_iter=10 ** 9, save_interval=50,
                      start_epoch=0, lr_e=2e-4, lr_d=2e-4)
        kwargs.update(extra_kwargs or {})
        getattr(ref, fn_name)(**kwargs)
    finally:
        ref.optim.Adam = torch.optim.Adam
        os.chdir(cwd)
    model = holder["model"]
    names = {id(p): n for n, p in model.named_parameters()}
    assert set(rec["postfix"]) == {"r_loss", "kl"}, rec["postfix"]          # the warm-up branch's postfix (:536)
    grads_e = {names[i]: g for i, g in rec["grads"][0].items() if g is not None}      # optimizer_e.step() (:533)
    grads_d = {names[i]: g for i, g in rec["grads"][1].items() if g is not None}      # optimizer_d.step() (:534)
    out = dict(arch=holder["arch"], batch=batch, seed=seed, hyper=dict(beta_kl=0.7, beta_rec=1.3), real=rec["real"], eps=rec["eps"][0],
               init=holder["init"], post=holder["post"], grads_e=grads_e, grads_d=grads_d,
               scalars=dict(r_loss=rec["postfix"]["r_loss"], kl=rec["postfix"]["kl"]), torch_version=torch.__version__,
               threads=torch.get_num_threads(), reference_commit="b6dbf16")
    torch.save(out, os.path.join(OUT, out_name))
    print(out_name, out["scalars"], len(grads_e), "encoder /", len(grads_d), "decoder gradient tensors")


def vae_warmups():
    vae_warmup("soft_intro_vae", "train_soft_intro_vae", "train_soft_intro_vae", "tiny_vae_std.pt")
    vae_warmup("soft_intro_vae_bootstrap", "train_soft_intro_vae_bootstrap", "train_soft_intro_vae", "tiny_vae_bootstrap.pt",
               extra_kwargs=dict(gamma_r=1.0, copy_to_target_freq=1))


def helper_functions():
    """the tensor-level helpers of the reference module called directly with non-default arguments: calc_kl with an outlier
    prior (mu_o, logvar_o; :231-251), calc_reconstruction_loss for every loss type x reduction (:268-294), reparameterize under
    a fixed seed (:254-265) -> tests/golden/helpers.pt (a few kB)"""
    ref = _import_reference("soft_intro_vae", "train_soft_intro_vae")
    g = torch.Generator().manual_seed(99)
    mu, lv = torch.randn(5, 7, generator=g), torch.randn(5, 7, generator=g) * 0.3
    x, y = torch.rand(4, 3, 6, 6, generator=g), torch.rand(4, 3, 6, 6, generator=g)
    out = dict(mu=mu, logvar=lv, x=x, recon=y, kl={}, rec={}, torch_version=torch.__version__, reference_commit="b6dbf16")
    for red in ("sum", "mean", "none"):
        out["kl"][("default", red)] = ref.calc_kl(lv, mu, reduce=red)
        out["kl"][("outlier", red)] = ref.calc_kl(lv, mu, mu_o=0.3, logvar_o=-0.2, reduce=red)
        out["kl"][("tensor_prior", red)] = ref.calc_kl(lv, mu, mu_o=torch.full((7,), 0.1), logvar_o=torch.full((7,), 0.4), reduce=red)
        for lt in ("mse", "l1", "bce"):
            out["rec"][(lt, red)] = ref.calc_reconstruction_loss(x, y, loss_type=lt, reduction=red)
    torch.manual_seed(5)
    out["reparam_seed5"] = ref.reparameterize(mu, lv)
    torch.save(out, os.path.join(OUT, "helpers.pt"))
    print("helpers.pt", os.path.getsize(os.path.join(OUT, "helpers.pt")), "bytes")


if __name__ == "__main__" and "--helpers" in sys.argv:
    helper_functions()
elif __name__ == "__main__" and "--vae" in sys.argv:
    vae_warmups()
elif __name__ == "__main__" and "--cond" in sys.argv:
    conditional_forward()
elif __name__ == "__main__" and "--losses" in sys.argv:
    extra_losses()
elif __name__ == "__main__" and "--toy" not in sys.argv:
    main()


def run_reference_toy(n_iter=12, num_vae=4, batch=64, seed=92):
    """2-D toy (BASELINE config 1): the UNMODIFIED reference `train_soft_intro_vae_toy` for a few iterations; records the
    lines it prints (test_iter=1) and fingerprints of the final weights."""
    import contextlib
    import io
    ref = _import_reference("soft_intro_vae_2d", "train_soft_intro_vae_2d")
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="sivae_golden2d_")
    os.chdir(tmp)
    buf = io.StringIO()
    try:
        torch.set_num_threads(8)
        with contextlib.redirect_stdout(buf):
            model = ref.train_soft_intro_vae_toy(z_dim=2, lr_e=2e-4, lr_d=2e-4, batch_size=batch, n_iter=n_iter, num_vae=num_vae,
                                                 save_interval=5000, recon_loss_type="mse", beta_kl=0.3, beta_rec=0.2,
                                                 beta_neg=0.9, test_iter=1, seed=seed, scale=1, device=torch.device("cpu"),
                                                 dataset="8Gaussians")
        res_line = open("results_log_soft_intro_vae.txt").read().strip()
    finally:
        os.chdir(cwd)
    lines = [l.strip() for l in buf.getvalue().splitlines() if l.startswith("Iter:")]
    import re
    lines = [re.sub(r"time:\s*[\d.]+:\s*", "", l) for l in lines]                      # drop the wall-clock field
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return dict(n_iter=n_iter, num_vae=num_vae, batch=batch, seed=seed, lines=lines, state=sd, results_line=res_line,
                torch_version=torch.__version__)


if __name__ == "__main__" and "--toy" in sys.argv:
    g = run_reference_toy()
    torch.save(g, os.path.join(OUT, "toy2d.pt"))
    print("\n".join(g["lines"]))
    print(g["results_line"])
