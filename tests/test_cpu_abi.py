"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/sivae.h declares, the model
description queried through it matches the reference's state_dict schema, the Python boundary mirrors the
reference module's names/signatures, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes as C
import importlib
import inspect
import os
import re

import pytest
import torch

from tests.step_harness import PKG, ROOT

L = importlib.import_module(PKG + ".lib")
M = importlib.import_module(PKG + ".train_soft_intro_vae")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sivae.h")).read()
    declared = set(re.findall(r"\b(sivae_[a-z0-9_]+)\s*\(", header))
    lib = L.load()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libsivae_b200.so does not export " + name
    assert declared == set(L.EXPORTS), (declared ^ set(L.EXPORTS))
    assert lib.sivae_version() >= 100


def _engine_schema(cfg, variant=0):
    lib = L.load()
    c = L.Config()
    c.cdim, c.zdim, c.image_size, c.n_channels = cfg["cdim"], cfg["zdim"], cfg["image_size"], len(cfg["channels"])
    for i, ch in enumerate(cfg["channels"]):
        c.channels[i] = ch
    c.max_batch, c.variant, c.conv_backend = 4, variant, 0
    c.cond_dim = cfg.get("cond_dim", 0)
    h = C.c_void_p()
    L.check(lib.sivae_create(C.byref(c), C.byref(h)), "create")
    out = {}
    for net, prefix in ((0, "encoder."), (1, "decoder.")) + (((2, "target_decoder."),) if variant else ()):
        for i in range(lib.sivae_num_tensors(h, net)):
            ti = L.TensorInfo()
            L.check(lib.sivae_tensor(h, net, i, C.byref(ti)), "tensor")
            out[prefix + ti.name.decode()] = tuple(ti.shape[:ti.ndim])
    ws = lib.sivae_workspace_bytes(h)
    lib.sivae_destroy(h)
    return out, ws


@pytest.mark.parametrize("cfg", [dict(cdim=3, zdim=128, channels=[64, 128, 256], image_size=32),
                                 dict(cdim=3, zdim=512, channels=[64, 128, 256, 512, 512, 512], image_size=256),
                                 dict(cdim=1, zdim=32, channels=[64, 128], image_size=28 + 4),
                                 # celeb1024 (:405-417): 8 stages from 16 channels
                                 dict(cdim=3, zdim=512, channels=[16, 32, 64, 128, 256, 512, 512, 512], image_size=1024)])
def test_engine_schema_equals_reference_state_dict(cfg):
    """index / shape work must be bit-exact: names, order and shapes of every parameter"""
    from oracle import sivae_oracle as O
    sd = O.make_state_dict(O.Arch(**cfg), seed=0)
    want = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    got, ws = _engine_schema(cfg)
    assert list(got.keys()) == list(want.keys())
    assert got == want
    assert ws > 0


def test_conditional_engine_schema():
    """cond_dim widens the two fc layers only (:106-109, :139-143)"""
    cfg = dict(cdim=3, zdim=16, channels=[32, 64], image_size=16)
    plain, _ = _engine_schema(cfg)
    cond, _ = _engine_schema(dict(cfg, cond_dim=10))
    assert list(plain.keys()) == list(cond.keys())
    for k in plain:
        if k == "encoder.fc.weight":
            assert cond[k] == (32, 64 * 4 * 4 + 10) and plain[k] == (32, 1024)
        elif k == "decoder.fc.0.weight":
            assert cond[k] == (1024, 16 + 10) and plain[k] == (1024, 16)
        else:
            assert cond[k] == plain[k], k


def test_recon_loss_and_conditional_entry_points_validate_their_arguments():
    """host-side argument checks of the round-2 entry points (no GPU work is enqueued before they fail)"""
    lib = L.load()
    c = L.Config()
    c.cdim, c.zdim, c.image_size, c.n_channels, c.max_batch = 3, 16, 16, 2, 2
    c.channels[0], c.channels[1] = 32, 64
    h = C.c_void_p()
    L.check(lib.sivae_create(C.byref(c), C.byref(h)), "create")
    assert lib.sivae_get_recon_loss(h) == L.LOSS_MSE
    assert lib.sivae_set_recon_loss(h, 7) != 0 and b"recon loss" in lib.sivae_last_error()
    assert lib.sivae_set_recon_loss(h, L.LOSS_L1) == 0 and lib.sivae_get_recon_loss(h) == L.LOSS_L1
    assert lib.sivae_set_recon_loss(h, L.LOSS_BCE) == 0 and lib.sivae_get_recon_loss(h) == L.LOSS_BCE
    assert lib.sivae_set_recon_loss(None, L.LOSS_MSE) != 0 and lib.sivae_get_recon_loss(None) == -1
    # the conditional entry points need an engine created with cond_dim > 0
    assert lib.sivae_encode_cond(h, None, None, 1, None, None, 0, None) != 0 and b"cond_dim" in lib.sivae_last_error()
    assert lib.sivae_decode_cond(h, 1, None, None, 1, None, 0, None) != 0 and b"cond_dim" in lib.sivae_last_error()
    lib.sivae_destroy(h)
    c.cond_dim = -1
    assert lib.sivae_create(C.byref(c), C.byref(h)) != 0 and b"cond_dim" in lib.sivae_last_error()
    # loader entry points: exactly one output, window inside the source
    assert lib.sivae_image_batch_u8_ex(None, None, None, 1, 8, 8, 3, 8, 8, 4, 4, None, None, None, None) != 0
    assert lib.sivae_jpeg_decode_batch(None, None, 0, 8, 8, None, None) != 0
    assert L.LOSS_TYPES == {"mse": 0, "l1": 1, "bce": 2}


def test_create_rejects_bad_configs():
    lib = L.load()
    c = L.Config()
    c.cdim, c.zdim, c.image_size, c.n_channels, c.max_batch = 3, 8, 30, 2, 1      # 30 not divisible by 4
    c.channels[0], c.channels[1] = 32, 64
    h = C.c_void_p()
    assert lib.sivae_create(C.byref(c), C.byref(h)) != 0
    assert b"image_size" in lib.sivae_last_error()


def test_boundary_mirrors_reference_signatures():
    want = {
        "train_soft_intro_vae": ["dataset", "z_dim", "lr_e", "lr_d", "batch_size", "num_workers", "start_epoch",
                                 "exit_on_negative_diff", "num_epochs", "num_vae", "save_interval", "recon_loss_type",
                                 "beta_kl", "beta_rec", "beta_neg", "test_iter", "seed", "pretrained", "device", "num_row",
                                 "gamma_r", "with_fid"],
        "calc_kl": ["logvar", "mu", "mu_o", "logvar_o", "reduce"],
        "reparameterize": ["mu", "logvar"],
        "calc_reconstruction_loss": ["x", "recon_x", "loss_type", "reduction"],
        "load_model": ["model", "pretrained", "device"],
        "save_checkpoint": ["model", "epoch", "iteration", "prefix"],
    }
    for fn, args in want.items():
        assert list(inspect.signature(getattr(M, fn)).parameters) == args, fn
    d = inspect.signature(M.train_soft_intro_vae).parameters
    assert d["gamma_r"].default == 1e-8 and d["batch_size"].default == 128 and d["num_epochs"].default == 250
    assert list(inspect.signature(M.SoftIntroVAE.__init__).parameters)[1:] == ["cdim", "zdim", "channels", "image_size", "conditional", "cond_dim"]
    assert list(inspect.signature(M.Decoder.__init__).parameters)[1:] == ["cdim", "zdim", "channels", "image_size", "conditional", "conv_input_size", "cond_dim"]
    assert list(inspect.signature(M.ResidualBlock.__init__).parameters)[1:] == ["inc", "outc", "groups", "scale"]


def test_model_init_is_bit_identical_to_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_std.pt"), weights_only=False)
    torch.manual_seed(g["seed"])
    model = M.SoftIntroVAE(**g["arch"])
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["init"].keys())
    for k, v in g["init"].items():
        assert torch.equal(sd[k], v), k


def test_conditional_model_init_is_bit_identical_to_reference(golden_dir):
    """conditional=True widens both fc layers by cond_dim (:106-109, :139-143): same constructors in the same order"""
    g = torch.load(os.path.join(golden_dir, "tiny_cond.pt"), weights_only=False)
    torch.manual_seed(g["seed"])
    model = M.SoftIntroVAE(conditional=True, cond_dim=g["cond_dim"], **g["arch"])
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["init"].keys())
    for k, v in g["init"].items():
        assert sd[k].shape == v.shape and torch.equal(sd[k], v), k
    assert model.conditional and model.cond_dim == g["cond_dim"] and model._arch["cond_dim"] == g["cond_dim"]
    assert M.SoftIntroVAE(cdim=3, zdim=8, channels=[32, 32], image_size=8, conditional=False, cond_dim=10)._arch["cond_dim"] == 0


def test_no_cpu_fallback():
    model = M.SoftIntroVAE(cdim=3, zdim=8, channels=[32, 32], image_size=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="cuda"):
        M.train_soft_intro_vae(dataset="synthetic32", device=torch.device("cpu"), num_epochs=1)


def test_helpers_match_oracle():
    from oracle import sivae_oracle as O
    mu, lv = torch.randn(5, 7), torch.randn(5, 7) * 0.3
    for red in ("sum", "mean", "none"):
        assert torch.allclose(M.calc_kl(lv, mu, reduce=red), O.calc_kl(lv, mu, red), rtol=1e-6)
    x, y = torch.rand(4, 3, 8, 8), torch.rand(4, 3, 8, 8)
    for red in ("sum", "mean", "none"):
        assert torch.allclose(M.calc_reconstruction_loss(x, y, "mse", red), O.rec_loss(x, y, red), rtol=1e-6)
    with pytest.raises(NotImplementedError):
        M.calc_reconstruction_loss(x, y, "mse", "bogus")
    with pytest.raises(NotImplementedError):
        M.calc_reconstruction_loss(x, y, "huber", "sum")
    assert M.str_to_list("1,2,3") == [1, 2, 3] and M.is_image_file("a.png") and not M.is_image_file("a.txt")


def test_learning_rate_schedule_equals_multisteplr():
    """a11: the reference builds MultiStepLR(milestones=(350,), gamma=0.1) on each optimiser (:453-454) and steps it once per epoch
    (:649-650), counting from the run's first epoch whatever start_epoch is; the drop-in computes the same rate in closed form"""
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=2e-4)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=(350,), gamma=0.1)
    for steps in range(0, 720):
        assert M._milestone_lr(2e-4, steps) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12), steps
        opt.step()
        sch.step()
    assert M._milestone_lr(1e-3, 349) == 1e-3 and M._milestone_lr(1e-3, 350) == pytest.approx(1e-4)


def test_helpers_match_reference_golden(golden_dir):
    """the drop-in's tensor-level helpers against the UNMODIFIED reference functions (tests/golden/helpers.pt,
    oracle/make_golden.py --helpers): calc_kl with scalar and tensor outlier priors, every loss type x reduction of
    calc_reconstruction_loss (l1 / bce hand `reduction` to F.*_loss on the [B, D] views: 'none' is element-wise), reparameterize
    under the same seed"""
    g = torch.load(os.path.join(golden_dir, "helpers.pt"), weights_only=False)
    mu, lv, x, y = g["mu"], g["logvar"], g["x"], g["recon"]
    for (kind, red), want in g["kl"].items():
        kw = {"default": {}, "outlier": dict(mu_o=0.3, logvar_o=-0.2),
              "tensor_prior": dict(mu_o=torch.full((7,), 0.1), logvar_o=torch.full((7,), 0.4))}[kind]
        got = M.calc_kl(lv, mu, reduce=red, **kw)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-6, atol=1e-7), (kind, red)
    for (lt, red), want in g["rec"].items():
        got = M.calc_reconstruction_loss(x, y, loss_type=lt, reduction=red)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-6, atol=1e-7), (lt, red)
    assert g["rec"][("l1", "none")].shape == (4, 108) and g["rec"][("mse", "none")].shape == (4,)
    torch.manual_seed(5)
    assert torch.allclose(M.reparameterize(mu, lv), g["reparam_seed5"], rtol=1e-6, atol=1e-7)


def test_bootstrap_model_init_is_bit_identical_to_reference(golden_dir):
    MB = importlib.import_module(PKG + ".train_soft_intro_vae_bootstrap")
    g = torch.load(os.path.join(golden_dir, "tiny_bootstrap.pt"), weights_only=False)
    torch.manual_seed(g["seed"])
    model = MB.SoftIntroVAE(**g["arch"])
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["init"].keys())
    for k, v in g["init"].items():
        assert torch.equal(sd[k], v), k
    sig = list(inspect.signature(MB.train_soft_intro_vae).parameters)
    assert "copy_to_target_freq" in sig and inspect.signature(MB.train_soft_intro_vae).parameters["gamma_r"].default == 1.0
    got, _ = _engine_schema(g["arch"], variant=1)
    want = {k: tuple(v.shape) for k, v in g["init"].items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    assert got == want
