"""cold-L2 timing of the fc kernels at the config-H shapes (CUDA events, 256 MB L2 flush between launches)"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("soft-intro-vae-pytorch_b200.lib")
lib = L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, B, F, O in (("enc fc 8192->1024", 32, 8192, 1024), ("dec fc 512->8192", 32, 512, 8192)):
    x, w, b, dy = torch.randn(B, F, device="cuda"), torch.randn(O, F, device="cuda"), torch.randn(O, device="cuda"), torch.randn(B, O, device="cuda")
    y, dx = torch.empty(B, O, device="cuda"), torch.empty(B, F, device="cuda")
    nws = lib.sivae_linear_dgrad_workspace_bytes(B, F, O)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    for what, fn in (("fwd", lambda: lib.sivae_linear_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), B, F, O, 0, None)),
                     ("dgrad", lambda: lib.sivae_linear_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, F, O, L.ptr(ws), nws, None))):
        ts = []
        for _ in range(6):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) * 1e3)
        print("%s %s: %.1f us (min of 6, cold L2); weights %.1f MB -> %.0f GB/s" % (name, what, min(ts[1:]), O * F * 4 / 1e6, O * F * 4 / min(ts[1:]) / 1e3))
