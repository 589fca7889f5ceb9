"""Launches one convolution shape a few times through the C ABI (for ncu captures of a specific kernel instance).
usage: python profiles/probe_one_conv.py N H W Cin Cout k [backend: 0 auto | 2 tcgen05] [addend 0|1]"""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("soft-intro-vae-pytorch_b200.lib")
N, H, W, Cin, Cout, k = [int(a) for a in sys.argv[1:7]]
backend = int(sys.argv[7]) if len(sys.argv) > 7 else 0
add = int(sys.argv[8]) if len(sys.argv) > 8 else 0
lib = L.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(N, H, W, Cin, device="cuda:0")
w = torch.randn(Cout, k, k, Cin, device="cuda:0") * 0.05
y = torch.zeros(N, H, W, Cout, device="cuda:0")
for _ in range(3):
    rc = lib.sivae_conv2d_fwd(L.ptr(x), L.ptr(w), None, L.ptr(y) if add else None, L.ptr(y), N, H, W, Cin, Cout, k, backend, st)
    assert rc == 0, rc
torch.cuda.synchronize()
print("ok")
