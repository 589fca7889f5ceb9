"""CPU study behind DESIGN.md section 9 item 1: what does the precision of the tensor-core OPERANDS cost the step?

Runs the oracle's introspective iteration in fp64 with every convolution's operands (forward: x, w; dgrad: dy, w; wgrad: x, dy)
rounded to a candidate storage format, accumulation exact, and measures the deviation of the logged scalars and of the
gradients from the unrounded fp64 run -- i.e. the pure operand-rounding error, the quantity that separates the engine's
tensor-core path from its exact path.

  tf32        10-bit mantissa, fp32 exponent (cvt.rna)                         -- today's tensor-core path
  fp16        10-bit mantissa, 5-bit exponent; activations / filters as they are, every gradient tensor multiplied by a
              per-tensor power of two (exact) that puts its max into [2^14, 2^15) before rounding  -- the proposed path
  fp16-noscale  the same without the gradient scale (shows why the scale is needed)
  bf16        7-bit mantissa

Test infrastructure (imports oracle/, hence under tests/; not collected by pytest); usage:  python tests/study_operand_formats.py > profiles/r01v_operand_formats.md
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sivae_oracle as O          # noqa: E402
from tests.step_harness import run_oracle_iteration, rel_l2   # noqa: E402


def r_tf32(x, grad=False):
    i = x.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


def r_fp16(x, grad=False, scale=True):
    if grad and scale:
        m = float(x.abs().max())
        if m == 0.0:
            return x
        k = 14 - int(torch.floor(torch.log2(torch.tensor(m))))          # max * 2^k in [2^14, 2^15)
        s = 2.0 ** k
        return (x * s).half().double() / s
    return x.half().double()


def r_fp16_noscale(x, grad=False):
    return r_fp16(x, grad, scale=False)


def r_bf16(x, grad=False):
    return x.bfloat16().double()


def make_conv(rnd):
    class RConv(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b, pad):
            ctx.save_for_backward(x, w)
            ctx.pad, ctx.has_b = pad, b is not None
            return F.conv2d(rnd(x), rnd(w), b, 1, pad)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            dx = torch.nn.grad.conv2d_input(x.shape, rnd(w), rnd(dy, True), 1, ctx.pad) if ctx.needs_input_grad[0] else None
            dw = torch.nn.grad.conv2d_weight(rnd(x), w.shape, rnd(dy, True), 1, ctx.pad) if ctx.needs_input_grad[1] else None
            db = dy.sum((0, 2, 3)) if ctx.has_b and ctx.needs_input_grad[2] else None
            return dx, dw, db, None

    def conv2d(x, w, b=None, stride=1, padding=0):
        assert stride == 1
        return RConv.apply(x, w, b, padding)
    return conv2d


class PatchedF:
    """oracle.F with conv2d replaced"""
    def __init__(self, conv):
        self._conv = conv

    def __getattr__(self, k):
        return self._conv if k == "conv2d" else getattr(F, k)


def run(cfg, batch, seed, rnd):
    saved = O.F
    if rnd is not None:
        O.F = PatchedF(make_conv(rnd))
    try:
        return run_oracle_iteration(cfg, batch, seed)
    finally:
        O.F = saved


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cases = [("tiny 16x16 [32,64] z16 B8", dict(cdim=3, zdim=16, channels=[32, 64], image_size=16), 8, 0),
             ("config C 32x32 [64,128,256] z128 B8", dict(cdim=3, zdim=128, channels=[64, 128, 256], image_size=32), 8, 11)]
    fmts = [("tf32", r_tf32), ("fp16 (+pow2 grad scale)", r_fp16), ("fp16-noscale", r_fp16_noscale), ("bf16", r_bf16)]
    print("# Operand storage format vs. step parity (fp64 accumulation, oracle on CPU)\n")
    print("Deviation from the unrounded fp64 step: worst logged scalar (relative) / median and worst gradient tensor (relative L2).\n")
    for name, cfg, batch, seed in cases:
        ref = run(cfg, batch, seed, None)
        print("## %s\n\n| operand format | worst scalar | median grad | worst grad |\n|---|---:|---:|---:|" % name)
        for fname, rnd in fmts:
            out = run(cfg, batch, seed, rnd)
            sc = max(abs(out["scalars"][k] - v) / (abs(v) + 1e-30) for k, v in ref["scalars"].items())
            gr = sorted(rel_l2(out[n][k], ref[n][k]) for n in ("grads_e", "grads_d") for k in ref[n])
            print("| %s | %.2e | %.2e | %.2e |" % (fname, sc, gr[len(gr) // 2], gr[-1]))
        print()


if __name__ == "__main__":
    main()
