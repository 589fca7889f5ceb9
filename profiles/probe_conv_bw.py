"""Times single tcgen05 conv launches through the C ABI, hot (same buffers every launch) and cold (rotating over
buffer sets far larger than L2), to separate kernel time from the in-situ memory-system state.
usage: python profiles/probe_conv_bw.py  (GPU box; prints one line per case)"""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("soft-intro-vae-pytorch_b200.lib")
DEV = "cuda:0"

CASES_ALL = [
    # N, H, W, Cin, Cout, k
    (32, 128, 128, 64, 128, 1),
    (32, 128, 128, 128, 64, 1),
    (32, 256, 256, 64, 64, 3),
    (32, 128, 128, 128, 128, 3),
    (32, 64, 64, 128, 256, 1),
    (32, 8, 8, 512, 512, 3),
    (32, 4, 4, 512, 512, 3),
    (32, 16, 16, 512, 512, 3),
    (32, 64, 64, 256, 256, 3),
    (32, 32, 32, 512, 512, 3),
]


CASES = CASES_ALL[:4] if os.environ.get("PROBE_SHORT") else CASES_ALL


def timed(fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    lib = L.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (N, H, W, Cin, Cout, k) in CASES:
        in_b = N * H * W * Cin * 4
        out_b = N * H * W * Cout * 4
        sets = max(2, int((1 << 31) // (in_b + out_b)))        # ~2 GB of distinct buffers
        sets = min(sets, 16)
        xs = [torch.randn(N, H, W, Cin, device=DEV) for _ in range(sets)]
        ys = [torch.empty(N, H, W, Cout, device=DEV) for _ in range(sets)]
        w = torch.randn(Cout, k, k, Cin, device=DEV) * 0.05

        def run(i, rot, add=0):
            j = (i % sets) if rot else 0
            a = None if add == 0 else (L.ptr(ys[j]) if add == 1 else L.ptr(ys[(j + 1) % sets]))
            rc = lib.sivae_conv2d_fwd(L.ptr(xs[j]), L.ptr(w), None, a, L.ptr(ys[j]), N, H, W, Cin, Cout, k, L.CONV_TCGEN05, st)
            assert rc == 0, rc

        for _ in range(3):
            run(0, False)
        hot = timed(lambda i: run(i, False), 20)
        for i in range(sets):
            run(i, True)
        cold = timed(lambda i: run(i, True), 2 * sets)
        for y in ys:
            y.zero_()
        addt = timed(lambda i: run(i, True, 1), 2 * sets)          # in-place addend (dx += conv), as the block backward does
        for y in ys:
            y.zero_()
        adds = timed(lambda i: run(i, True, 2), 2 * sets)          # addend from a different buffer
        gb = (in_b + out_b) / 1e9
        fl = 2.0 * N * H * W * Cin * Cout * k * k / 1e12
        print("N%d %dx%d %d->%d k%d  sets=%d  hot %.4f ms (%.2f TB/s, %.0f TF)   cold %.4f ms (%.2f TB/s, %.0f TF)   +addend in-place %.4f ms, separate %.4f ms" % (
            N, H, W, Cin, Cout, k, sets, hot, gb / hot, fl / hot * 1e3, cold, gb / cold, fl / cold * 1e3, addt, adds), flush=True)
        del xs, ys


if __name__ == "__main__":
    main()
