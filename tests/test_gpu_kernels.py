"""-m gpu: single-kernel parity through the C ABI against plain torch fp32/fp64 references of the same op."""
import ctypes as C
import os
import importlib

import pytest
import torch
import torch.nn.functional as F

from tests.step_harness import PKG

pytestmark = pytest.mark.gpu
L = importlib.import_module(PKG + ".lib")
DEV = "cuda:0"


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _nhwc(x):           # NCHW tensor -> contiguous NHWC storage
    return x.permute(0, 2, 3, 1).contiguous()


def _krsc(w):           # [Cout,Cin,kh,kw] -> [Cout,kh,kw,Cin]
    return w.permute(0, 2, 3, 1).contiguous()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _round_tf32(x):
    # round-to-nearest (ties away, like cvt.rna) to 10 mantissa bits
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)


CONV_SHAPES = [
    # N, H, W, Cin, Cout, k
    (2, 8, 8, 3, 32, 5),       # stem-like
    (2, 8, 8, 32, 3, 5),       # predict-like
    (3, 16, 16, 32, 64, 3),
    (2, 4, 4, 64, 64, 3),
    (5, 8, 8, 32, 64, 1),      # conv_expand
    (1, 6, 10, 8, 12, 3),      # ragged / non-power-of-two
    (3, 20, 12, 3, 64, 5),     # stem-like, partial 16x16 tiles
    (2, 16, 16, 64, 3, 5),     # predict-like at the narrow-correlation kernel's channel count
    (1, 8, 8, 1, 32, 5),       # cdim = 1 (mnist-like)
    (2, 32, 16, 3, 64, 5),     # row-separable tensor-core stem (two stacked 16-row tiles)
    (3, 16, 24, 3, 32, 5),     # row-separable stem, single tile row, 32 output channels
    (1, 16, 16, 3, 128, 5),    # row-separable stem with two output-channel tiles
    (2, 32, 16, 64, 3, 5),     # row-separable predict conv
    (1, 16, 8, 32, 3, 5),      # row-separable predict conv, one input chunk
    (2, 16, 16, 1, 32, 5),     # row-separable, cdim = 1
    (2, 16, 16, 32, 1, 5),
    (2, 16, 16, 64, 128, 3),   # Cout >= 128: eligible for the CTA-pair (cta_group::2) kernel when SIVAE_TC_2CTA=1
    (2, 32, 16, 32, 256, 3),
    (2, 16, 16, 32, 320, 3),   # second 256-wide output tile only partly filled
]


def _conv_case(N, H, W, Cin, Cout, k, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    dy = torch.randn(N, Cout, H, W, generator=g)
    return x, w, dy


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("backend", [L.CONV_SIMT, L.CONV_TCGEN05, L.CONV_AUTO, L.CONV_TF32])
def test_conv_fwd_dgrad_wgrad(shape, backend):
    lib = L.load()
    N, H, W, Cin, Cout, k = shape
    x, w, dy = _conv_case(*shape)
    tc = backend == L.CONV_TCGEN05
    if backend == L.CONV_AUTO:
        # the default mode mixes split32 forward, bf16 / tf32 backward kernels: bf16 values are exact operands of all of them
        x, w, dy = x.bfloat16().float(), w.bfloat16().float(), dy.bfloat16().float()
    elif backend != L.CONV_SIMT:
        x, w, dy = _round_tf32(x), _round_tf32(w), _round_tf32(dy)      # operands exactly representable in tf32
    xd, wd, dyd = x.double(), w.double(), dy.double()
    xd.requires_grad_(True)
    wd.requires_grad_(True)
    bias = torch.randn(Cout)
    addend = torch.randn(N, Cout, H, W)
    y_ref = F.conv2d(xd, wd, bias.double(), 1, k // 2) + addend.double()
    y_ref.backward(dyd)
    tol = 2e-5
    xg, wg, dyg = _nhwc(x).to(DEV), _krsc(w).to(DEV), _nhwc(dy).to(DEV)
    bg, ag = bias.to(DEV), _nhwc(addend).to(DEV)
    y = torch.empty(N, H, W, Cout, device=DEV)
    rc = lib.sivae_conv2d_fwd(L.ptr(xg), L.ptr(wg), L.ptr(bg), L.ptr(ag), L.ptr(y), N, H, W, Cin, Cout, k, backend, _s())
    if tc and rc == -7:
        pytest.skip("shape not eligible for the tcgen05 kernel (SIMT handles it)")
    L.check(rc, "conv fwd")
    torch.cuda.synchronize()
    assert _rel(y.cpu(), _nhwc(y_ref.detach())) < tol
    # dgrad
    ws = torch.empty(max(1 << 25, Cout * Cin * k * k * 4 * 64), dtype=torch.uint8, device=DEV)
    dx = torch.empty(N, H, W, Cin, device=DEV)
    rc = lib.sivae_conv2d_dgrad(L.ptr(dyg), L.ptr(wg), None, L.ptr(dx), N, H, W, Cin, Cout, k, backend, L.ptr(ws), ws.numel(), _s())
    if not (tc and rc == -7):
        L.check(rc, "conv dgrad")
        torch.cuda.synchronize()
        assert _rel(dx.cpu(), _nhwc(xd.grad)) < tol
    # wgrad (accumulate on top of a known value)
    dw = torch.full((Cout, k, k, Cin), 0.5, device=DEV)
    rc = lib.sivae_conv2d_wgrad(L.ptr(xg), L.ptr(dyg), L.ptr(dw), N, H, W, Cin, Cout, k, 1, backend, L.ptr(ws), ws.numel(), _s())
    if not (tc and rc == -7):
        L.check(rc, "conv wgrad")
        torch.cuda.synchronize()
        assert _rel(dw.cpu() - 0.5, _krsc(wd.grad)) < tol


@pytest.mark.parametrize("shape", [sh for sh in CONV_SHAPES if sh[3] % 32 == 0] + [(2, 32, 32, 64, 64, 3), (2, 8, 8, 128, 128, 3)])
def test_conv_tc3x_unrounded_operands(shape):
    """compensated tensor-core conv (SIVAE_CONV_TC3X): full-precision fp32 operands, three tf32 MMAs per product on the
    hi/lo split -- fp32-class agreement with fp64 (plain tf32 on the same operands is ~5e-4)"""
    lib = L.load()
    N, H, W, Cin, Cout, k = shape
    x, w, dy = _conv_case(*shape, seed=3)
    xd, wd, dyd = x.double().requires_grad_(True), w.double().requires_grad_(True), dy.double()
    bias = torch.randn(Cout)
    addend = torch.randn(N, Cout, H, W)
    y_ref = F.conv2d(xd, wd, bias.double(), 1, k // 2) + addend.double()
    y_ref.backward(dyd)
    tol = 2e-5
    xg, wg, dyg = _nhwc(x).to(DEV), _krsc(w).to(DEV), _nhwc(dy).to(DEV)
    bg, ag = bias.to(DEV), _nhwc(addend).to(DEV)
    y = torch.empty(N, H, W, Cout, device=DEV)
    rc = lib.sivae_conv2d_fwd(L.ptr(xg), L.ptr(wg), L.ptr(bg), L.ptr(ag), L.ptr(y), N, H, W, Cin, Cout, k, L.CONV_TC3X, _s())
    if rc == -7:
        pytest.skip("shape not eligible for the tcgen05 kernel")
    L.check(rc, "conv fwd 3x")
    torch.cuda.synchronize()
    assert _rel(y.cpu(), _nhwc(y_ref.detach())) < tol
    ws = torch.empty(max(1 << 25, Cout * Cin * k * k * 4 * 64), dtype=torch.uint8, device=DEV)
    if Cout % 32 == 0:          # the dgrad is a forward conv with Cout input channels
        dx = torch.empty(N, H, W, Cin, device=DEV)
        L.check(lib.sivae_conv2d_dgrad(L.ptr(dyg), L.ptr(wg), None, L.ptr(dx), N, H, W, Cin, Cout, k, L.CONV_TC3X, L.ptr(ws),
                                       ws.numel(), _s()), "conv dgrad 3x")
        torch.cuda.synchronize()
        assert _rel(dx.cpu(), _nhwc(xd.grad)) < tol
    for acc in (1, 0):
        dw = torch.full((Cout, k, k, Cin), 0.5, device=DEV)
        rc = lib.sivae_conv2d_wgrad(L.ptr(xg), L.ptr(dyg), L.ptr(dw), N, H, W, Cin, Cout, k, acc, L.CONV_TC3X, L.ptr(ws), ws.numel(), _s())
        if rc == -7:
            break
        L.check(rc, "conv wgrad 3x")
        torch.cuda.synchronize()
        assert _rel(dw.cpu() - 0.5 * acc, _krsc(wd.grad)) < tol


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("with_id,use_mask", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("dims", [(3, 8, 8, 32), (2, 6, 10, 24), (5, 32, 32, 8)])      # small maps: fused one-block-per-channel-group backward; last: rows > 4096 -> reduce / finalize / apply
def test_bn_act_fwd_bwd(mode, with_id, use_mask, dims):
    lib = L.load()
    N, H, W, Cc = dims
    g = torch.Generator().manual_seed(5)
    t = torch.randn(N, Cc, H, W, generator=g) * 2 + 0.3
    idn = torch.randn(N, Cc, H, W, generator=g) if with_id else None
    gamma, beta = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.1
    rm, rv = torch.randn(Cc, generator=g) * 0.1, torch.rand(Cc, generator=g) + 0.5
    td = t.double().requires_grad_(True)
    idd = idn.double().requires_grad_(True) if with_id else None
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rmd, rvd = rm.double().clone(), rv.double().clone()
    y = F.batch_norm(td, rmd, rvd, gd, bd, True, 0.1, 1e-5)
    if with_id:
        y = y + idd
    y = F.leaky_relu(y, 0.2)
    if mode == 1:
        y = F.avg_pool2d(y, 2)
    elif mode == 2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    dout = torch.randn(y.shape, generator=g)
    y.backward(dout.double())
    tg = _nhwc(t).to(DEV)
    ig = _nhwc(idn).to(DEV) if with_id else None
    gg, bg, rmg, rvg = gamma.to(DEV), beta.to(DEV), rm.to(DEV), rv.to(DEV)
    nbt = torch.zeros(1, dtype=torch.int64, device=DEV)
    mi = torch.empty(2 * Cc, device=DEV)
    out = torch.empty(_nhwc(y.detach()).shape, device=DEV)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    mask = torch.full((N * H * W * Cc // 4,), 0xF0, dtype=torch.uint8, device=DEV) if use_mask else None
    if use_mask:
        L.check(lib.sivae_bn_act_fwd_m(L.ptr(tg), L.ptr(ig), L.ptr(gg), L.ptr(bg), L.ptr(rmg), L.ptr(rvg), L.ptr(nbt), L.ptr(mi),
                                       L.ptr(out), N, H, W, Cc, mode, 1, L.ptr(ws), ws.numel(), L.ptr(mask), _s()), "bn fwd (mask)")
    else:
        L.check(lib.sivae_bn_act_fwd(L.ptr(tg), L.ptr(ig), L.ptr(gg), L.ptr(bg), L.ptr(rmg), L.ptr(rvg), L.ptr(nbt), L.ptr(mi),
                                     L.ptr(out), N, H, W, Cc, mode, 1, L.ptr(ws), ws.numel(), _s()), "bn fwd")
    torch.cuda.synchronize()
    assert _rel(out.cpu(), _nhwc(y.detach())) < 1e-5
    assert torch.allclose(rmg.cpu().double(), rmd, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rvg.cpu().double(), rvd, rtol=1e-5, atol=1e-6)
    assert int(nbt) == 1
    dt = torch.empty_like(tg)
    gi = torch.empty_like(tg)
    dgam, dbet = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    doutg = _nhwc(dout).to(DEV)
    if use_mask:
        # sign bytes = the pre-activation's signs (bit j <-> channel 4*c4 + j), exactly (integer work)
        yb = F.batch_norm(t.double(), None, None, gamma.double(), beta.double(), True, 0.1, 1e-5) + idn.double()
        bits = (_nhwc(yb) > 0).reshape(N * H * W, Cc // 4, 4).to(torch.int64)
        want = (bits * torch.tensor([1, 2, 4, 8])).sum(-1).reshape(-1)
        near0 = (_nhwc(yb).abs() < 1e-5).reshape(N * H * W, Cc // 4, 4).any(-1).reshape(-1)      # fp32 vs fp64 sign flips at |y| ~ 0
        assert torch.equal(mask.cpu().to(torch.int64)[~near0], want[~near0])
        bogus = torch.full_like(ig, float("nan"))      # the masked backward must not read the identity tensor
        L.check(lib.sivae_bn_act_bwd_m(L.ptr(doutg), L.ptr(tg), L.ptr(bogus), L.ptr(gg), L.ptr(bg), L.ptr(mi), L.ptr(dt),
                                       L.ptr(gi), L.ptr(dgam), L.ptr(dbet), 0, N, H, W, Cc, mode, L.ptr(ws), ws.numel(), L.ptr(mask),
                                       _s()), "bn bwd (mask)")
    else:
        L.check(lib.sivae_bn_act_bwd(L.ptr(doutg), L.ptr(tg), L.ptr(ig), L.ptr(gg), L.ptr(bg), L.ptr(mi), L.ptr(dt),
                                     L.ptr(gi), L.ptr(dgam), L.ptr(dbet), 0, N, H, W, Cc, mode, L.ptr(ws), ws.numel(), _s()), "bn bwd")
    torch.cuda.synchronize()
    assert _rel(dt.cpu(), _nhwc(td.grad)) < 2e-5
    assert _rel(dgam.cpu(), gd.grad) < 2e-5 and _rel(dbet.cpu(), bd.grad) < 2e-5
    if with_id:
        assert _rel(gi.cpu(), _nhwc(idd.grad)) < 2e-5


def test_bn_eval_mode():
    lib = L.load()
    N, H, W, Cc = 2, 4, 4, 8
    t = torch.randn(N, Cc, H, W)
    gamma, beta, rm, rv = torch.rand(Cc) + 0.5, torch.randn(Cc), torch.randn(Cc), torch.rand(Cc) + 0.5
    y = F.leaky_relu(F.batch_norm(t, rm.clone(), rv.clone(), gamma, beta, False, 0.1, 1e-5), 0.2)
    tg = _nhwc(t).to(DEV)
    rmg, rvg = rm.to(DEV), rv.to(DEV)
    mi = torch.empty(2 * Cc, device=DEV)
    out = torch.empty_like(tg)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=DEV)
    gg, bg = gamma.to(DEV), beta.to(DEV)       # keep the device tensors alive while the kernel runs
    L.check(lib.sivae_bn_act_fwd(L.ptr(tg), None, L.ptr(gg), L.ptr(bg), L.ptr(rmg), L.ptr(rvg), None,
                                 L.ptr(mi), L.ptr(out), N, H, W, Cc, 0, 0, L.ptr(ws), ws.numel(), _s()), "bn eval")
    torch.cuda.synchronize()
    assert _rel(out.cpu(), _nhwc(y)) < 1e-5
    assert torch.equal(rmg.cpu(), rm) and torch.equal(rvg.cpu(), rv)       # eval mode must not touch the buffers


@pytest.mark.parametrize("B,per", [(4, 3 * 32 * 32), (3, 3 * 16 * 16), (2, 3 * 128 * 128), (1, 4)])
def test_mse3_fused_loss(B, per):
    lib = L.load()
    g = torch.Generator().manual_seed(B)
    t = [torch.rand(B, per, generator=g) for _ in range(5)]
    real, rec, rec_rec, fake, rec_fake = t
    ref = torch.stack([(rec.double() - real.double()).pow(2).sum(1), (rec_rec.double() - rec.double()).pow(2).sum(1),
                       (rec_fake.double() - fake.double()).pow(2).sum(1)], 1)
    d = [x.to(DEV) for x in t]
    out = torch.empty(B, 3, device=DEV)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    L.check(lib.sivae_mse3(*[L.ptr(x) for x in d], L.ptr(out), B, per, L.ptr(ws), ws.numel(), _s()), "mse3")
    torch.cuda.synchronize()
    assert torch.allclose(out.cpu().double(), ref, rtol=2e-6)


def test_kl_reparam():
    lib = L.load()
    B, z = 7, 48
    ml = torch.randn(B, 2 * z) * 0.5
    eps = torch.randn(B, z)
    mu, lv = ml[:, :z].double(), ml[:, z:].double()
    kl_ref = -0.5 * (1 + lv - mu.pow(2) - lv.exp()).sum(1)
    z_ref = mu + eps.double() * torch.exp(0.5 * lv)
    zz, kl = torch.empty(B, z, device=DEV), torch.empty(B, device=DEV)
    mlg, epg = ml.to(DEV), eps.to(DEV)
    L.check(lib.sivae_kl_reparam(L.ptr(mlg), L.ptr(epg), L.ptr(zz), L.ptr(kl), B, z, _s()), "kl")
    torch.cuda.synchronize()
    assert torch.allclose(kl.cpu().double(), kl_ref, rtol=1e-5)
    assert torch.allclose(zz.cpu().double(), z_ref, rtol=1e-5, atol=1e-6)


def test_adam_matches_torch_optim():
    lib = L.load()
    n = 10007
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=2e-4)
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * 10.0 ** float(torch.randint(-6, 2, (1,), generator=g))
        p_ref.grad = grad.clone()
        opt.step()
        gdev = (grad * 4).to(DEV)
        L.check(lib.sivae_adam_flat(L.ptr(p), L.ptr(gdev), L.ptr(m), L.ptr(v), n, 2e-4, 0.25, step, _s()), "adam")
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    assert torch.allclose(p.cpu(), p_ref.detach(), rtol=0, atol=2e-7)


@pytest.mark.parametrize("dims", [(32, 8192, 1024), (32, 512, 8192), (128, 4096, 256), (128, 128, 4096), (5, 36, 10), (3, 7, 9),
                                  (33, 260, 17)])
def test_linear_fwd_dgrad(dims):
    """fc kernels (vectorised paths for F % 4 == 0 at the config shapes, scalar fallbacks otherwise) against fp64 matmul"""
    lib = L.load()
    B, F, O = dims
    g = torch.Generator().manual_seed(B + F + O)
    x, w, b, dy = torch.randn(B, F, generator=g), torch.randn(O, F, generator=g) / F ** 0.5, torch.randn(O, generator=g), torch.randn(B, O, generator=g)
    xd, wd, bd, dyd = x.cuda(), w.cuda(), b.cuda(), dy.cuda()
    for relu in (0, 1):
        y = torch.empty(B, O, device="cuda")
        L.check(lib.sivae_linear_fwd(L.ptr(xd), L.ptr(wd), L.ptr(bd), L.ptr(y), B, F, O, relu, None), "sivae_linear_fwd")
        ref = x.double() @ w.double().t() + b.double()
        ref = ref.clamp_min(0) if relu else ref
        assert (y.cpu().double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    nws = lib.sivae_linear_dgrad_workspace_bytes(B, F, O)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    dx = torch.empty(B, F, device="cuda")
    L.check(lib.sivae_linear_dgrad(L.ptr(dyd), L.ptr(wd), L.ptr(dx), B, F, O, L.ptr(ws), nws, None), "sivae_linear_dgrad")
    ref = dy.double() @ w.double()
    assert (dx.cpu().double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    dx2 = torch.empty_like(dx)
    L.check(lib.sivae_linear_dgrad(L.ptr(dyd), L.ptr(wd), L.ptr(dx2), B, F, O, L.ptr(ws), nws, None), "sivae_linear_dgrad")
    assert torch.equal(dx, dx2)                     # deterministic (fixed-order split reduction)


SPLIT_SHAPES = [sh for sh in CONV_SHAPES if sh[3] % 32 == 0 or (sh[3] <= 3 and sh[1] % 16 == 0 and sh[2] % 8 == 0)] + [
    (2, 32, 32, 64, 64, 3), (2, 8, 8, 128, 128, 3), (4, 32, 32, 64, 64, 3), (2, 64, 64, 128, 128, 3), (2, 32, 32, 256, 256, 3),
    (4, 16, 16, 128, 512, 3), (3, 4, 4, 512, 512, 3), (2, 32, 32, 64, 128, 1), (32, 4, 4, 256, 256, 3)]


@pytest.mark.parametrize("shape", SPLIT_SHAPES)
def test_conv_fwd_split32_unrounded_operands(shape):
    """the engine's default FORWARD conv (SIVAE_CONV_AUTO): unrounded fp32 operands stored as bf16 hi + lo pairs (split32), three
    kind::f16 MMAs per product (lo*hi + hi*lo + hi*hi, the 2^-18 lo*lo term dropped), fp32 accumulation -- fp32-class agreement
    with the fp64 conv of the same operands (plain tf32 on them is ~5e-4); every tensor-core forward kernel family: CTA-pair
    halo, single-CTA halo, one-box-per-tap (1x1, small maps, split-K), row-separable stem / predict"""
    lib = L.load()
    N, H, W, Cin, Cout, k = shape
    x, w, _ = _conv_case(*shape, seed=5)
    bias = torch.randn(Cout)
    addend = torch.randn(N, Cout, H, W)
    ref = F.conv2d(x.double(), w.double(), bias.double(), 1, k // 2) + addend.double()
    xg, wg, bg, ag = _nhwc(x).to(DEV), _krsc(w).to(DEV), bias.to(DEV), _nhwc(addend).to(DEV)
    y = torch.empty(N, H, W, Cout, device=DEV)
    L.check(lib.sivae_conv2d_fwd(L.ptr(xg), L.ptr(wg), L.ptr(bg), L.ptr(ag), L.ptr(y), N, H, W, Cin, Cout, k, L.CONV_AUTO, _s()), "conv fwd split32")
    torch.cuda.synchronize()
    # subtract the exactly known bias + addend so that the relative error is that of the convolution itself
    got = y.cpu().double() - _nhwc(addend).double() - bias.double()
    want = _nhwc(F.conv2d(x.double(), w.double(), None, 1, k // 2))
    assert _rel(got, want) < 1e-5
    assert _rel(y.cpu(), _nhwc(ref)) < 1e-5


BWD16_SHAPES = [(2, 16, 16, 64, 64, 3), (2, 32, 32, 64, 64, 3), (2, 64, 64, 128, 128, 3), (4, 16, 16, 128, 512, 3), (2, 32, 32, 256, 256, 3),
                (3, 4, 4, 512, 512, 3), (2, 8, 8, 128, 128, 3), (2, 32, 32, 64, 128, 1), (32, 4, 4, 256, 256, 3), (5, 16, 24, 64, 192, 3),
                (2, 128, 128, 64, 64, 3), (3, 8, 8, 128, 64, 1)]


@pytest.mark.parametrize("shape", BWD16_SHAPES)
def test_conv_bwd_bf16_operands(shape):
    """dgrad / wgrad of the residual blocks in the default mode: plain bf16 operands (the BatchNorm backward writes bf16
    gradients, the forward keeps bf16 copies of its activations), kind::f16 MMAs, fp32 accumulation.  With bf16-exact inputs the
    kernels must agree with fp64 to fp32 round-off: CTA-pair / single-CTA halo and one-box-per-tap dgrad (with and without the
    in-place addend), halo (3 taps per instruction pair) and generic MN-major wgrad, accumulate on / off, several splits."""
    lib = L.load()
    N, H, W, Cin, Cout, k = shape
    x, w, dy = _conv_case(*shape, seed=9)
    x, w, dy = x.bfloat16().float(), w.bfloat16().float(), dy.bfloat16().float()
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(xd, wd, None, 1, k // 2).backward(dy.double())
    xg, wg, dyg = _nhwc(x).to(DEV), _krsc(w).to(DEV), _nhwc(dy).to(DEV)
    ws = torch.empty(max(1 << 26, Cout * Cin * k * k * 4 * 160), dtype=torch.uint8, device=DEV)
    addend = torch.randn(N, Cin, H, W)
    for add in (None, addend):
        dx = torch.empty(N, H, W, Cin, device=DEV)
        ag = _nhwc(add).to(DEV) if add is not None else None
        L.check(lib.sivae_conv2d_dgrad(L.ptr(dyg), L.ptr(wg), L.ptr(ag), L.ptr(dx), N, H, W, Cin, Cout, k, L.CONV_AUTO, L.ptr(ws),
                                       ws.numel(), _s()), "bf16 dgrad")
        torch.cuda.synchronize()
        got = dx.cpu().double() - (_nhwc(add).double() if add is not None else 0)
        assert _rel(got, _nhwc(xd.grad)) < 1e-5
    for acc in (1, 0):
        dw = torch.full((Cout, k, k, Cin), 0.5, device=DEV)
        L.check(lib.sivae_conv2d_wgrad(L.ptr(xg), L.ptr(dyg), L.ptr(dw), N, H, W, Cin, Cout, k, acc, L.CONV_AUTO, L.ptr(ws), ws.numel(),
                                       _s()), "bf16 wgrad")
        torch.cuda.synchronize()
        assert _rel(dw.cpu() - 0.5 * acc, _krsc(wd.grad)) < 1e-5
