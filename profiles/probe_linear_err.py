"""precision of the fc kernels vs fp64: vectorised paths and (through deliberately misaligned operands) the scalar fallbacks"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("soft-intro-vae-pytorch_b200.lib")
lib = L.load()
def mis(t):      # same values at a 4-byte-misaligned address
    s = torch.empty(t.numel() + 1, device="cuda"); s[1:].copy_(t.reshape(-1)); return s[1:].view(t.shape)
for B, F, O in ((8, 4096, 64), (8, 32, 4096), (32, 8192, 1024), (32, 512, 8192), (128, 4096, 256)):
    g = torch.Generator().manual_seed(B + F + O)
    x, w, b, dy = torch.randn(B, F, generator=g), torch.randn(O, F, generator=g) / F ** 0.5, torch.randn(O, generator=g), torch.randn(B, O, generator=g)
    ref_y = x.double() @ w.double().t() + b.double(); ref_dx = dy.double() @ w.double()
    for tag, xd, wd in (("vec", x.cuda(), w.cuda()), ("scalar", mis(x.cuda()), mis(w.cuda()))):
        y = torch.empty(B, O, device="cuda"); dx = torch.empty(B, F, device="cuda")
        L.check(lib.sivae_linear_fwd(L.ptr(xd), L.ptr(wd), L.ptr(b.cuda()), L.ptr(y), B, F, O, 0, None), "fwd")
        nws = lib.sivae_linear_dgrad_workspace_bytes(B, F, O); ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
        L.check(lib.sivae_linear_dgrad(L.ptr(dy.cuda()), L.ptr(wd), L.ptr(dx), B, F, O, L.ptr(ws), nws, None), "dgrad")
        torch.cuda.synchronize()
        ey = (y.cpu().double() - ref_y); edx = (dx.cpu().double() - ref_dx)
        print("B%d F%d O%d %-6s fwd relL2 %.2e max %.2e | dgrad relL2 %.2e max %.2e" % (B, F, O, tag, ey.norm() / ref_y.norm(), ey.abs().max(), edx.norm() / ref_dx.norm(), edx.abs().max()))
