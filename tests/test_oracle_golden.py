"""Pins oracle/sivae_oracle.py (the CPU restatement) against outputs of the UNMODIFIED reference
(tests/golden/*.pt, produced by oracle/make_golden.py from /root/reference in the build container)."""
import os

import pytest
import torch

from oracle import sivae_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _run(g, bootstrap):
    torch.set_num_threads(g["threads"])
    arch = O.Arch(**g["arch"])
    sd = O.clone_sd(g["init"])
    hp = O.Hyper(**g["hyper"])
    st_e, st_d = O.AdamState(), O.AdamState()
    scal, ge, gd, te, td = O.full_iteration(sd, arch, g["real"], g["noise"], g["eps"], hp, st_e, st_d, bootstrap)
    return sd, scal, ge, gd


def _relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("name,bootstrap", [("tiny_std.pt", False), ("tiny_bootstrap.pt", True),
                                            ("tiny_l1.pt", False), ("tiny_bce.pt", False)])   # recon_loss_type l1 / bce (:288-291)
def test_oracle_matches_reference_full(golden_dir, name, bootstrap):
    g = _load(golden_dir, name)
    assert g["hyper"].get("loss_type", "mse") == {"tiny_l1.pt": "l1", "tiny_bce.pt": "bce"}.get(name, "mse")
    sd, scal, ge, gd = _run(g, bootstrap)
    s = g["scalars"]
    # same ops, same thread count => agreement to fp32 round-off
    assert scal["loss_rec"] == pytest.approx(s["rec_err"], rel=1e-6)
    assert scal["lossE_real_kl"] == pytest.approx(s["kl_real"], rel=1e-6)
    assert scal["lossD_fake_kl"] == pytest.approx(s["kl_fake"], rel=1e-6)
    assert scal["lossD_rec_kl"] == pytest.approx(s["kl_rec"], rel=1e-6)
    assert scal["expelbo_fake"] == pytest.approx(s["expelbo_f"], rel=1e-5)
    assert set(ge) == set(g["grads_e"]) and set(gd) == set(g["grads_d"])
    for k in ge:
        assert _relerr(ge[k], g["grads_e"][k]) < 1e-5, k
    for k in gd:
        assert _relerr(gd[k], g["grads_d"][k]) < 1e-5, k
    for k, v in g["post"].items():
        if k.startswith("target_decoder") and not bootstrap:
            continue
        if v.is_floating_point():
            assert torch.allclose(sd[k], v, rtol=1e-5, atol=1e-7), k
        else:
            assert int(sd[k]) == int(v), k      # num_batches_tracked: exact


def test_make_state_dict_is_reference_init(golden_dir):
    """index/shape work + RNG order: the oracle's own constructor reproduces the reference init bit-exactly."""
    for name, boot in (("tiny_std.pt", False), ("tiny_bootstrap.pt", True)):
        g = _load(golden_dir, name)
        sd = O.make_state_dict(O.Arch(**g["arch"]), seed=g["seed"], bootstrap=boot)
        assert list(sd.keys()) == list(g["init"].keys())
        for k, v in g["init"].items():
            assert sd[k].shape == v.shape and torch.equal(sd[k], v), k


def test_oracle_matches_reference_cifar_summary(golden_dir):
    g = _load(golden_dir, "cifar_std_summary.pt")
    arch = O.Arch(**g["arch"])
    sd = O.make_state_dict(arch, seed=g["seed"])
    for k, (s, a, n) in g["init_fp"].items():
        assert float(sd[k].double().sum()) == pytest.approx(s, rel=1e-12, abs=1e-12), k
    torch.set_num_threads(g["threads"])
    real = g["real"]
    hp = O.Hyper(**g["hyper"])
    scal, ge, gd, _, _ = O.full_iteration(sd, arch, real, g["noise"], g["eps"], hp, O.AdamState(), O.AdamState())
    s = g["scalars"]
    assert scal["loss_rec"] == pytest.approx(s["rec_err"], rel=1e-6)
    assert scal["lossE_real_kl"] == pytest.approx(s["kl_real"], rel=1e-6)
    assert scal["lossD_fake_kl"] == pytest.approx(s["kl_fake"], rel=1e-6)
    assert scal["lossD_rec_kl"] == pytest.approx(s["kl_rec"], rel=1e-6)
    assert scal["expelbo_fake"] == pytest.approx(s["expelbo_f"], rel=1e-5)
    for k, (ssum, sabs, nrm) in g["grads_e_fp"].items():
        assert float(ge[k].double().norm()) == pytest.approx(nrm, rel=1e-4), k
    for k, (ssum, sabs, nrm) in g["grads_d_fp"].items():
        assert float(gd[k].double().norm()) == pytest.approx(nrm, rel=1e-4), k
    for k, v in g["nbt_post"].items():
        assert int(sd[k]) == v, k
    for k, v in g["post_small"].items():
        if v.is_floating_point():
            assert torch.allclose(sd[k], v, rtol=1e-4, atol=1e-6), k


def test_oracle_conditional_forward_matches_reference(golden_dir):
    """SoftIntroVAE(conditional=True): fc over [features | cond] (:106-109, :118-119) and [z | cond] (:139-143, :163-165) --
    the oracle's forward against the unmodified reference model in train and eval mode (tests/golden/tiny_cond.pt)"""
    g = _load(golden_dir, "tiny_cond.pt")
    torch.set_num_threads(8)
    arch = O.Arch(**g["arch"])
    sd = O.clone_sd(g["init"])
    assert sd["encoder.fc.weight"].shape[1] == 64 * 4 * 4 + g["cond_dim"] and sd["decoder.fc.0.weight"].shape[1] == 16 + g["cond_dim"]
    for mode in ("train", "eval"):
        t = mode == "train"
        with torch.no_grad():
            mu, lv = O.encoder_forward(sd, arch, g["x"], train=t, o_cond=g["cond"])
            y = O.decoder_forward(sd, arch, mu, train=t, y_cond=g["cond"])
            smp = O.decoder_forward(sd, arch, g["z"], train=t, y_cond=g["cond"])
        for name, v in (("mu", mu), ("logvar", lv), ("y", y), ("sample", smp)):
            assert torch.allclose(v, g[mode][name], rtol=1e-5, atol=1e-6), (mode, name)
        if t:
            for k, v in g["post_train"].items():
                assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-7), k
    assert "shapes cannot be multiplied" in g["uncond_error"]       # what the reference does without a condition


@pytest.mark.parametrize("name,bootstrap", [("tiny_vae_std.pt", False), ("tiny_vae_bootstrap.pt", True)])
def test_oracle_vae_warmup_step_matches_reference(golden_dir, name, bootstrap):
    """VAE warm-up iteration (`epoch < num_vae`, :512-540): the oracle's vae_step + both Adam steps against the UNMODIFIED
    reference trainer run with num_vae = 1 (oracle/make_golden.py --vae).  Bootstrap: `model(real_batch)` decodes through the
    frozen target decoder (bootstrap trainer :196-217), so the trainable decoder receives NO gradient and optimizer_d.step()
    moves nothing -- recorded as an empty decoder-gradient set."""
    g = _load(golden_dir, name)
    torch.set_num_threads(g["threads"])
    arch = O.Arch(**g["arch"])
    sd = O.clone_sd(g["init"])
    hp = O.Hyper(beta_kl=g["hyper"]["beta_kl"], beta_rec=g["hyper"]["beta_rec"])
    scal, ge, gd = O.vae_step(sd, arch, g["real"], g["eps"], hp, bootstrap=bootstrap)
    assert scal["loss_rec"] == pytest.approx(g["scalars"]["r_loss"], rel=1e-6)
    assert scal["loss_kl"] == pytest.approx(g["scalars"]["kl"], rel=1e-6)
    assert set(ge) == set(g["grads_e"]) and set(gd) == set(g["grads_d"])
    assert (len(gd) == 0) == bootstrap
    for k in ge:
        assert _relerr(ge[k], g["grads_e"][k]) < 1e-5, k
    for k in gd:
        assert _relerr(gd[k], g["grads_d"][k]) < 1e-5, k
    O.adam_update(sd, ge, O.AdamState(), 2e-4, hp)          # optimizer_e.step(), :533
    O.adam_update(sd, gd, O.AdamState(), 2e-4, hp)          # optimizer_d.step(), :534 (bootstrap: nothing to move)
    for k, v in g["post"].items():
        if v.is_floating_point():
            assert torch.allclose(sd[k], v, rtol=1e-5, atol=1e-7), k
        else:
            assert int(sd[k]) == int(v), k
    if bootstrap:
        for k, v in g["init"].items():
            if k.startswith("decoder.") and v.is_floating_point() and "running" not in k:
                assert torch.equal(g["post"][k], v), k      # the reference's own decoder did not move in the warm-up step
