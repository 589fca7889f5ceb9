"""CPU study for DESIGN.md section 10, item 1: may the three gradient tensors a residual block still moves in fp32 -- the dgrad
results `DA1` (conv2) and `dx` (conv1 + identity path) and the identity-path gradient `G2` -- be STORED in bf16?

Method of tests/study_split_formats.py (the oracle iteration in fp64 with exact accumulation and only the modelled roundings
applied, teacher-forced like the -m gpu parity tests).  Baseline = the benchmarked backend: forward convs on bf16 hi + lo operands
(3 products), dgrad / wgrad on plain bf16 operands.  Variants add a round-to-bf16 of what would be stored:

  + DA1, dx partials   every conv's input gradient is rounded when it leaves the conv (conv2's is DA1 itself)
  + dx                 the gradient arriving at a block input (conv1 dgrad + identity path, accumulated in fp32) is rounded
  + G2                 the identity-path gradient of a block without conv_expand is rounded before it is added to dx
                       (with conv_expand it already is a bf16 conv operand today)

Printed: worst logged scalar (unchanged by construction: the forward is the same), median / worst relative-L2 deviation of the
gradient tensors from the unrounded fp64 step -- against the 4e-2 bound tests/test_gpu_step.py holds the default backend to.

Test infrastructure (imports oracle/); usage:  python tests/study_grad_storage.py > profiles/r02q_grad_storage_study.md
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sivae_oracle as O                                      # noqa: E402
from tests.step_harness import rel_l2                                      # noqa: E402
from tests.study_split_formats import Fmt, PatchedF, r_bf16, run as run_fmt  # noqa: E402


def make_conv(fwd, bwd, store_dx):
    """the RConv of study_split_formats with one addition: the input gradient is rounded to bf16 when it leaves the conv"""
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
                if store_dx and w.shape[2] == 3:             # the residual blocks' 3x3 convs; stem / predict stay fp32 (tf32 path)
                    dx = r_bf16(dx)
            if ctx.needs_input_grad[1]:
                f = Fmt(bwd.ra, bwd.ra, {"b": "a", "a": "b"}.get(bwd.split, bwd.split))
                dw = f.product(x, dy, lambda a, c: torch.nn.grad.conv2d_weight(a, w.shape, c, 1, ctx.pad))
            if ctx.has_b and ctx.needs_input_grad[2]:
                db = dy.sum((0, 2, 3))
            return dx, dw, db, None

    def conv2d(x, w, b=None, stride=1, padding=0):
        assert stride == 1
        return RConv.apply(x, w, b, padding)
    return conv2d


def make_block(store_in, store_g2):
    """O.residual_block (reference :65-75) with gradient-rounding hooks at the block input and on the identity path"""
    def residual_block(sd, p, x, train):
        if store_in and x.requires_grad:
            x = x * 1.0                                       # own node: its gradient = conv1 dgrad + identity path = the stored dx
            x.register_hook(lambda g: r_bf16(g))
        if (p + ".conv_expand.weight") in sd:
            identity = O.F.conv2d(x, sd[p + ".conv_expand.weight"], None, 1, 0)
        else:
            identity = x
            if store_g2 and x.requires_grad:
                identity = x * 1.0
                identity.register_hook(lambda g: r_bf16(g))   # G2 as it would be stored
        out = O.F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1)
        out = O.F.leaky_relu(O._bn(sd, p + ".bn1", out, train), O.LRELU_SLOPE)
        out = O.F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
        out = O._bn(sd, p + ".bn2", out, train)
        return O.F.leaky_relu(out + identity, O.LRELU_SLOPE)
    return residual_block


def run(cfg, batch, seed, variant, hp=None, teacher=None):
    S, B16 = Fmt(split=True), Fmt(r_bf16, r_bf16)
    store_dx, store_in, store_g2 = variant
    saved_block = O.residual_block
    O.residual_block = make_block(store_in, store_g2)
    import tests.study_split_formats as SF
    saved_make = SF.make_conv
    SF.make_conv = lambda fwd, bwd: make_conv(fwd, bwd, store_dx)
    try:
        return run_fmt(cfg, batch, seed, (S, B16), hp, teacher=teacher)
    finally:
        SF.make_conv = saved_make
        O.residual_block = saved_block


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cases = [("tiny 16x16 [32,64] z16 B8", dict(cdim=3, zdim=16, channels=[32, 64], image_size=16), 8, 0, None),
             ("config C 32x32 [64,128,256] z128 B8", dict(cdim=3, zdim=128, channels=[64, 128, 256], image_size=32), 8, 11, None),
             ("64x64 [64,128,256,512] z256 B4 beta_neg 256", dict(cdim=3, zdim=256, channels=[64, 128, 256, 512], image_size=64), 4, 3, None)]
    if os.environ.get("STUDY_BIG"):
        cases.append(("config M 128x128 [64,128,256,512,512] z256 B2", dict(cdim=3, zdim=256, channels=[64, 128, 256, 512, 512], image_size=128), 2, 7, None))
    variants = [("baseline: split32 forward, bf16 dgrad / wgrad operands (today)", (False, False, False)),
                ("+ conv input gradients (DA1, conv1 partial) stored bf16", (True, False, False)),
                ("+ block-input gradient dx stored bf16", (True, True, False)),
                ("+ identity-path gradient G2 stored bf16 (all three)", (True, True, True))]
    print("# bf16 STORAGE of the residual blocks' fp32 gradient tensors vs. gradient parity (fp64 accumulation, oracle on CPU)\n")
    print("Deviation from the unrounded fp64 step; bound on the default backend in tests/test_gpu_step.py: 4e-2 (relative L2 per tensor).\n")
    for name, cfg, batch, seed, hp in cases:
        ref = run_fmt(cfg, batch, seed, None, hp)
        print("## %s\n\n| stored in bf16 | worst scalar | median grad | 90th pct | worst grad | worst tensor |\n|---|---:|---:|---:|---:|---|" % name)
        for vname, v in variants:
            out = run(cfg, batch, seed, v, hp, teacher=ref)
            sc = max(abs(out["scalars"][k] - val) / (abs(val) + 1e-30) for k, val in ref["scalars"].items())
            gr = sorted((rel_l2(out[n][k], ref[n][k]), k) for n in ("grads_e", "grads_d") for k in ref[n])
            print("| %s | %.2e | %.2e | %.2e | %.2e | %s |" % (vname, sc, gr[len(gr) // 2][0], gr[int(len(gr) * 0.9)][0], gr[-1][0], gr[-1][1]))
            sys.stdout.flush()
        print()


if __name__ == "__main__":
    main()
