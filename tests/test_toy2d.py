"""BASELINE config 1 (8Gaussians 2-D Soft-IntroVAE on CPU, plumbing / correctness): the 2-D twin module reproduces the
log of the UNMODIFIED reference trainer (tests/golden/toy2d.pt from oracle/make_golden.py --toy) for a seeded run."""
import contextlib
import importlib
import inspect
import io
import os
import re

import torch

from tests.step_harness import PKG

T = importlib.import_module(PKG + ".train_soft_intro_vae_2d")


def test_toy_signatures():
    sig = list(inspect.signature(T.train_soft_intro_vae_toy).parameters)
    assert sig == ["z_dim", "lr_e", "lr_d", "batch_size", "n_iter", "num_vae", "save_interval", "recon_loss_type", "beta_kl",
                   "beta_rec", "beta_neg", "test_iter", "seed", "pretrained", "scale", "device", "dataset", "gamma_r"]
    assert list(inspect.signature(T.calc_kl).parameters) == ["logvar", "mu", "mu_o", "is_outlier", "reduce"]
    m = T.SoftIntroVAESimple(x_dim=2, zdim=2, n_layers=3, num_hidden=256)
    assert sum(p.numel() for p in m.parameters()) == 397831       # incl. the unused loggamma
    assert "decoder.loggamma" in m.state_dict()


def test_toy_run_reproduces_reference_log(golden_dir, tmp_path):
    g = torch.load(os.path.join(golden_dir, "toy2d.pt"), weights_only=False)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    buf = io.StringIO()
    try:
        torch.set_num_threads(8)
        with contextlib.redirect_stdout(buf):
            model = T.train_soft_intro_vae_toy(z_dim=2, lr_e=2e-4, lr_d=2e-4, batch_size=g["batch"], n_iter=g["n_iter"],
                                               num_vae=g["num_vae"], save_interval=5000, recon_loss_type="mse", beta_kl=0.3,
                                               beta_rec=0.2, beta_neg=0.9, test_iter=1, seed=g["seed"], scale=1,
                                               device=torch.device("cpu"), dataset="8Gaussians")
        res = open("results_log_soft_intro_vae.txt").read().strip()
    finally:
        os.chdir(cwd)
    lines = [re.sub(r"time:\s*[\d.]+:\s*", "", l.strip()) for l in buf.getvalue().splitlines() if l.startswith("Iter:")]
    assert len(lines) == len(g["lines"]) == g["n_iter"]

    def nums(s):
        return [float(x) for x in re.findall(r"-?\d+\.\d+", s)]
    for mine, ref in zip(lines, g["lines"]):
        assert re.sub(r"-?\d+\.\d+", "#", mine) == re.sub(r"-?\d+\.\d+", "#", ref)          # same fields, same order
        for a, b in zip(nums(mine), nums(ref)):
            assert abs(a - b) <= 2e-4, (mine, ref)                                          # 4-decimal log lines
    for k, v in g["state"].items():
        assert torch.allclose(model.state_dict()[k], v, rtol=1e-4, atol=2e-6), k
    a, b = res.split("_gnelbo_"), g["results_line"].split("_gnelbo_")
    assert a[0] == b[0]
    for x, y in zip(re.findall(r"[-\d.e]+", a[1])[:3], re.findall(r"[-\d.e]+", b[1])[:3]):
        assert abs(float(x) - float(y)) <= 1e-3 * abs(float(y)) + 1e-9
