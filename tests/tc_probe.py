"""Stand-alone probe of the tcgen05 conv kernels (run under `timeout` on the GPU box; prints one flushed line per
case so that a trap / hang is attributable).  Not collected by pytest.  usage: python tests/tc_probe.py [fwd|wgrad|all]"""
import ctypes as C
import importlib
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("soft-intro-vae-pytorch_b200.lib")
DEV = "cuda:0"


def rt(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def say(*a):
    print(*a, flush=True)


def case(N, H, W, Cin, Cout, k, what):
    lib = L.load()
    g = torch.Generator().manual_seed(N * 1000 + H + Cin)
    x = rt(torch.randn(N, Cin, H, W, generator=g))
    w = rt(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    dy = rt(torch.randn(N, Cout, H, W, generator=g))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xg, wg, dyg = nhwc(x).to(DEV), nhwc(w).to(DEV), nhwc(dy).to(DEV)
    tag = "N%d H%d W%d Cin%d Cout%d k%d" % (N, H, W, Cin, Cout, k)
    if what in ("fwd", "all"):
        y = torch.zeros(N, H, W, Cout, device=DEV)
        t0 = time.time()
        rc = lib.sivae_conv2d_fwd(L.ptr(xg), L.ptr(wg), None, None, L.ptr(y), N, H, W, Cin, Cout, k, L.CONV_TCGEN05, st)
        say("fwd   %-40s launch rc=%d" % (tag, rc))
        if rc == 0:
            torch.cuda.synchronize()
            ref = nhwc(F.conv2d(x.double(), w.double(), None, 1, k // 2))
            say("fwd   %-40s rel=%.3e  (%.1f ms incl. sync)" % (tag, rel(y.cpu(), ref), (time.time() - t0) * 1e3))
    if what in ("wgrad", "all"):
        ws = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
        dw = torch.zeros(Cout, k, k, Cin, device=DEV)
        rc = lib.sivae_conv2d_wgrad(L.ptr(xg), L.ptr(dyg), L.ptr(dw), N, H, W, Cin, Cout, k, 0, L.CONV_TCGEN05, L.ptr(ws), ws.numel(), st)
        say("wgrad %-40s launch rc=%d" % (tag, rc))
        if rc == 0:
            torch.cuda.synchronize()
            wd = w.double().requires_grad_(True)
            F.conv2d(x.double(), wd, None, 1, k // 2).backward(dy.double())
            say("wgrad %-40s rel=%.3e" % (tag, rel(dw.cpu(), wd.grad.permute(0, 2, 3, 1))))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    say("device:", torch.cuda.get_device_name(0))
    for shp in [(1, 16, 16, 32, 32, 1), (1, 16, 16, 32, 64, 3), (2, 8, 8, 64, 64, 3), (8, 4, 4, 64, 128, 3),
                (3, 16, 16, 32, 64, 3), (5, 8, 8, 32, 64, 1), (2, 32, 32, 64, 64, 3), (2, 16, 16, 128, 256, 3),
                (4, 32, 32, 64, 3 * 0 + 64, 5), (32, 4, 4, 512, 512, 3), (2, 32, 32, 64, 128, 3), (1, 64, 64, 64, 64, 3),
                (3, 32, 16, 32, 96, 3), (2, 64, 32, 128, 512, 3)]:
        case(*shp, what=what)
    say("probe done")
