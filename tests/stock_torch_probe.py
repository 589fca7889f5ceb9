"""GPU comparator ("recommended"): the same introspective E+D step executed by stock PyTorch on the same
B200 -- the oracle's functional restatement of the reference step (torch.nn.functional conv2d / batch_norm / ... with
autograd, i.e. cuDNN + cuBLAS + ATen kernels), fp32 and with TF32 allowed.  This is what the unmodified reference would
dispatch to on this box; it is the honest "kernel to beat", next to the CPU baseline that bench.py reports.

Test infrastructure (it executes oracle/): run by hand on a GPU box,
    python tests/stock_torch_probe.py --config H --steps 5 --warmup 2 > gpurun_out/stock_torch_H.json
not part of the product path and not imported by it.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sivae_oracle as O   # noqa: E402

CONFIGS = {"C": (32, 128, [64, 128, 256], 128, 256.0), "M": (128, 256, [64, 128, 256, 512, 512], 64, 256.0),
           "H": (256, 512, [64, 128, 256, 512, 512, 512], 32, 1024.0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="H", choices=list(CONFIGS))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    size, zdim, channels, batch, beta_neg = CONFIGS[args.config]
    batch = args.batch or batch
    dev = torch.device("cuda:0")
    arch = O.Arch(cdim=3, zdim=zdim, channels=channels, image_size=size)
    hp = O.Hyper(beta_neg=beta_neg, gamma_r=1e-8, scale=1.0 / (3 * size * size))
    g = torch.Generator().manual_seed(1234)
    real = torch.rand(batch, 3, size, size, generator=g).to(dev)
    noise = torch.randn(batch, zdim, generator=g).to(dev)
    eps = [e.to(dev) for e in torch.randn(5, batch, zdim, generator=g)]
    out = dict(config=args.config, batch=batch, steps=args.steps, warmup=args.warmup, torch=torch.__version__,
               cudnn=torch.backends.cudnn.version(), gpu=torch.cuda.get_device_name(0), runs=[])
    for name, tf32, cl in (("fp32 (allow_tf32=False, the reference's setting)", False, False),
                           ("tf32 (cudnn.allow_tf32 + matmul.allow_tf32)", True, False),
                           ("tf32 + cudnn.benchmark", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = cl
        sd = {k: v.to(dev) for k, v in O.make_state_dict(arch, seed=0).items()}
        se, sdd = O.AdamState(), O.AdamState()
        for _ in range(args.warmup):
            O.full_iteration(sd, arch, real, noise, eps, hp, se, sdd, False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            scal = O.full_iteration(sd, arch, real, noise, eps, hp, se, sdd, False)[0]
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        out["runs"].append(dict(mode=name, ms_per_step=round(ms, 2), images_per_s=round(batch / ms * 1e3, 2),
                                peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                                lossE=float(scal.get("lossE", float("nan"))), lossD=float(scal.get("lossD", float("nan")))))
        del sd, se, sdd
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
