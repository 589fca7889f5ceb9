"""CPU study for round 2: which operand format lets the tensor-core path hold the north-star 1e-4?

Same method as tests/study_operand_formats.py (oracle iteration in fp64, only the convolution operands rounded, exact
accumulation), with SPLIT formats and per-direction / per-operand switches:

  tf32             both operands rounded to 11 significant bits                          (round-1 tensor path)
  tf32 act-only    activations / gradients rounded, filters exact                        (which operand carries the error?)
  tf32 w-only      filters rounded, activations / gradients exact
  tf32 fwd-only    forward convs rounded, dgrad / wgrad exact                            (which direction?)
  tf32 bwd-only
  bf16x2           x = hi + lo, hi = bf16(x), lo = bf16(x - hi); product = hi*hi + lo*hi + hi*lo  (3 MMAs at kind::f16 rate,
                   the lo*lo term dropped)                                                (round-2 tensor path)
  bf16x2 fwd / tf32 bwd

Test infrastructure (imports oracle/); usage:  python tests/study_split_formats.py > profiles/r02a_split_formats.md
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sivae_oracle as O          # noqa: E402
from tests.step_harness import run_oracle_iteration, rel_l2   # noqa: E402


def r_tf32(x):
    i = x.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


def r_none(x):
    return x


def split_bf16(x):
    x32 = x.float()
    hi = x32.bfloat16().float()
    lo = (x32 - hi).bfloat16().float()
    return hi.double(), lo.double()


def r_bf16(x):
    return x.float().bfloat16().double()


class Fmt:
    """product(a, b, op): op(a', b') summed over the terms the format issues.  split: True = both operands hi+lo (3 products),
    'a' = first operand hi+lo and second bf16 (2 products), 'b' = the reverse"""
    def __init__(self, ra=r_none, rw=r_none, split=False):
        self.ra, self.rw, self.split = ra, rw, split

    def product(self, a, b, op):
        if self.split == "a":
            ah, al = split_bf16(a)
            bh = r_bf16(b)
            return op(ah, bh) + op(al, bh)
        if self.split == "b":
            bh, bl = split_bf16(b)
            ah = r_bf16(a)
            return op(ah, bh) + op(ah, bl)
        if self.split:
            ah, al = split_bf16(a)
            bh, bl = split_bf16(b)
            return op(ah, bh) + op(al, bh) + op(ah, bl)
        return op(self.ra(a), self.rw(b))


def make_conv(fwd, bwd):
    class RConv(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b, pad):
            ctx.save_for_backward(x, w)
            ctx.pad, ctx.has_b = pad, b is not None
            y = fwd.product(x, w, lambda a, c: F.conv2d(a, c, None, 1, pad))
            return y if b is None else y + b.view(1, -1, 1, 1)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            dx = dw = db = None
            if ctx.needs_input_grad[0]:
                dx = bwd.product(dy, w, lambda a, c: torch.nn.grad.conv2d_input(x.shape, c, a, 1, ctx.pad))
            if ctx.needs_input_grad[1]:
                # both wgrad operands are activation-like
                f = Fmt(bwd.ra, bwd.ra, {"b": "a", "a": "b"}.get(bwd.split, bwd.split))   # wgrad(x, dy): dy is the second operand
                dw = f.product(x, dy, lambda a, c: torch.nn.grad.conv2d_weight(a, w.shape, c, 1, ctx.pad))
            if ctx.has_b and ctx.needs_input_grad[2]:
                db = dy.sum((0, 2, 3))
            return dx, dw, db, None

    def conv2d(x, w, b=None, stride=1, padding=0):
        assert stride == 1
        return RConv.apply(x, w, b, padding)
    return conv2d


class PatchedF:
    def __init__(self, conv):
        self._conv = conv

    def __getattr__(self, k):
        return self._conv if k == "conv2d" else getattr(F, k)


def run(cfg, batch, seed, fmts, hp=None, teacher=None):
    """teacher = the unrounded run: its E-half gradients drive Adam(encoder), so the D half of every format starts from the
    same encoder (what the -m gpu parity tests do with `teacher_enc`)."""
    from tests.step_harness import DEFAULT_HP, make_inputs
    saved = O.F
    if fmts is not None:
        O.F = PatchedF(make_conv(*fmts))
    try:
        h = dict(DEFAULT_HP, **(hp or {}))
        h.setdefault("scale", 1.0 / (cfg["cdim"] * cfg["image_size"] ** 2))
        arch = O.Arch(**cfg)
        sd = O.clone_sd(O.make_state_dict(arch, seed=seed), torch.float64)
        real, noise, eps = make_inputs(cfg, batch, seed)
        ohp = O.Hyper(beta_kl=h["beta_kl"], beta_rec=h["beta_rec"], beta_neg=h["beta_neg"], gamma_r=h["gamma_r"],
                      scale=h["scale"], lr_e=h["lr_e"], lr_d=h["lr_d"])
        scal, ge, gd, te, td = O.full_iteration(sd, arch, real.double(), noise.double(), [e.double() for e in eps], ohp,
                                                O.AdamState(), O.AdamState(), False,
                                                world_grads_e=None if teacher is None else teacher["grads_e"])
        return dict(scalars=scal, grads_e=ge, grads_d=gd)
    finally:
        O.F = saved


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cases = [("tiny 16x16 [32,64] z16 B8", dict(cdim=3, zdim=16, channels=[32, 64], image_size=16), 8, 0, None),
             ("config C 32x32 [64,128,256] z128 B8", dict(cdim=3, zdim=128, channels=[64, 128, 256], image_size=32), 8, 11, None)]
    if os.environ.get("STUDY_BIG"):
        cases.append(("64x64 [64,128,256,512] z256 B4 beta_neg 256", dict(cdim=3, zdim=256, channels=[64, 128, 256, 512], image_size=64), 4, 3, None))
    T, X, S = Fmt(r_tf32, r_tf32), Fmt(), Fmt(split=True)
    fmts = [("tf32", (T, T)), ("tf32 act-only", (Fmt(r_tf32, r_none), Fmt(r_tf32, r_none))),
            ("tf32 w-only", (Fmt(r_none, r_tf32), Fmt(r_none, r_tf32))),
            ("tf32 fwd-only", (T, X)), ("tf32 bwd-only", (X, T)),
            ("bf16x2 (3 products)", (S, S)), ("bf16x2 fwd, tf32 bwd", (S, T)),
            ("bf16x2 fwd, bf16 bwd (1 product)", (S, Fmt(r_bf16, r_bf16))),
            ("bf16x2 fwd, bwd: dy bf16, w / x hi+lo (2 products)", (S, Fmt(split="b"))),
            ("bf16x2 fwd, bwd: dy hi+lo, w / x bf16 (2 products)", (S, Fmt(split="a")))]
    print("# Split operand formats vs. step parity (fp64 accumulation, oracle on CPU)\n")
    print("Deviation from the unrounded fp64 step: worst logged scalar (relative, which) / median and worst gradient tensor (relative L2).\n")
    for name, cfg, batch, seed, hp in cases:
        ref = run(cfg, batch, seed, None, hp)
        print("## %s\n\n| operand format | worst scalar | which | median grad | worst grad |\n|---|---:|---|---:|---:|" % name)
        for fname, f in fmts:
            out = run(cfg, batch, seed, f, hp, teacher=ref)
            sc = max((abs(out["scalars"][k] - v) / (abs(v) + 1e-30), k) for k, v in ref["scalars"].items())
            gr = sorted(rel_l2(out[n][k], ref[n][k]) for n in ("grads_e", "grads_d") for k in ref[n])
            print("| %s | %.2e | %s | %.2e | %.2e |" % (fname, sc[0], sc[1], gr[len(gr) // 2], gr[-1]))
            sys.stdout.flush()
        print()


if __name__ == "__main__":
    main()
