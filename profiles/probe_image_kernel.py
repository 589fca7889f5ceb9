"""three launches of the image batch-assembly kernel at the bench geometry (for ncu: -k regex:k_image_batch -s 2 -c 1)"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = importlib.import_module("soft-intro-vae-pytorch_b200.gpu_dataset")
g = torch.Generator().manual_seed(1)
src = [torch.randint(0, 256, (32, 1024, 1024, 3), dtype=torch.uint8, generator=g).cuda() for _ in range(3)]
flags = (torch.arange(32) % 2).to(torch.uint8).cuda()
bt = G.ImageBatcher(256, 256, "cuda:0")
for s in src:
    out = bt(s, flags)
torch.cuda.synchronize()
print("ok", float(out.mean()))
