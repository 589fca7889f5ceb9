"""torchrun worker of tests/test_gpu_dist.py (one process per GPU, NCCL): the data-parallel introspective iteration.

Checks, on every rank:
  1. after introspective_iteration() the flat encoder / decoder gradient buffers hold the SUM over ranks of the gradients
     each rank computes alone from its own shard (the engine's kernels are deterministic, so the local gradients of a second
     model with the same init and inputs are reproduced bit for bit and all-gathered for the comparison);
  2. parameters, Adam moments and statistics are bit-identical on all ranks after the step (one all-reduce result feeds the
     same Adam kernel everywhere) while BatchNorm running statistics stay rank-local;
  3. the CUDA-graph replay of the step (three graph segments, the two NCCL all-reduces launched eagerly between them)
     leaves the same state as eager execution over several iterations.
Prints "DIST_OK" from rank 0 on success."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "soft-intro-vae-pytorch_b200"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = importlib.import_module(PKG + ".lib")
    M = importlib.import_module(PKG + ".train_soft_intro_vae")
    E = importlib.import_module(PKG + ".engine")
    cfg = dict(cdim=3, zdim=32, channels=[32, 64], image_size=32)
    B, iters = 8, 4
    hp = E.make_hyper(1.0, 1.0, 256.0, 1e-8, 1.0 / (3 * 32 * 32))
    g = torch.Generator().manual_seed(100 + rank)                 # rank-local shard, noise and eps
    reals = [torch.rand(B, 3, 32, 32, generator=g).to(dev) for _ in range(iters)]
    noises = [torch.randn(B, 32, generator=g).to(dev) for _ in range(iters)]
    epss = [torch.randn(5, B, 32, generator=g).to(dev) for _ in range(iters)]

    def fresh():
        torch.manual_seed(4)                                      # same init on every rank
        stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
        try:
            return M.SoftIntroVAE(**cfg).to(dev)
        finally:
            sys.stdout = stdout

    # ---- 1. local gradients (no collective) vs the all-reduced buffers ---------------------------------------
    a = fresh()
    ea = a.reserve(B)
    ea.e_step(reals[0], noises[0], epss[0][:3].contiguous(), hp)
    torch.cuda.synchronize()
    local_ge = ea.mem[L.NET_ENCODER].grads.clone()
    b = fresh()
    M.introspective_iteration(b, reals[0], noises[0], epss[0], hp, 2e-4, 2e-4, use_graph=False)
    torch.cuda.synchronize()
    eb = b._engine
    gathered = [torch.empty_like(local_ge) for _ in range(world)]
    dist.all_gather(gathered, local_ge)
    want = torch.stack(gathered).double().sum(0)
    got = eb.mem[L.NET_ENCODER].grads.double()
    err = float((got - want).norm() / (want.norm() + 1e-30))
    assert err < 1e-6, "all-reduced encoder gradient differs from the sum of the local ones: rel %.3g" % err
    # the encoder step must have used the MEAN gradient: redo Adam on model a with the gathered mean
    ea.mem[L.NET_ENCODER].grads.copy_(eb.mem[L.NET_ENCODER].grads)
    ea.adam(L.NET_ENCODER, 2e-4, 1.0 / world)
    torch.cuda.synchronize()
    assert torch.equal(ea.mem[L.NET_ENCODER].params, eb.mem[L.NET_ENCODER].params), "encoder Adam step is not grad/world"

    # ---- 2. replicas stay identical, BN buffers rank-local ----------------------------------------------------
    for net in (L.NET_ENCODER, L.NET_DECODER):
        for name in ("params", "m", "v"):
            t = getattr(eb.mem[net], name)
            ref = t.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(t, ref), "rank %d: %s of net %d differs from rank 0" % (rank, name, net)
    bn = eb.mem[L.NET_ENCODER].bn
    ref = bn.clone()
    dist.broadcast(ref, 0)
    if rank != 0:
        assert not torch.equal(bn, ref), "BatchNorm running statistics should be rank-local (different shards)"

    # ---- 3. segmented graph replay == eager --------------------------------------------------------------------
    outs = []
    for use_graph in (False, True):
        m = fresh()
        stats = []
        for i in range(iters):
            st = M.introspective_iteration(m, reals[i], noises[i], epss[i], hp, 2e-4, 2e-4, use_graph=use_graph)
            stats.append(st.clone())
        torch.cuda.synchronize()
        if use_graph:
            assert len(m._engine._graphs) == 3, "the graph path did not capture its three segments"
        outs.append(({k: v.detach().clone() for k, v in m.state_dict().items()}, stats))
        dist.barrier()
    (sd_a, st_a), (sd_b, st_b) = outs
    for x, y in zip(st_a, st_b):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-6), "graph vs eager statistics differ: %s vs %s" % (x.tolist(), y.tolist())
    worst = 0.0
    for k in sd_a:
        if sd_a[k].is_floating_point():
            worst = max(worst, float((sd_a[k].double() - sd_b[k].double()).abs().max()))
        else:
            assert torch.equal(sd_a[k], sd_b[k]), k
    assert worst <= 2.05 * 2e-4 * iters, "graph vs eager parameters differ by %.3g" % worst
    bit_identical = all(torch.equal(sd_a[k], sd_b[k]) for k in sd_a)
    dist.barrier()
    if rank == 0:
        print("DIST_OK world=%d graph_vs_eager_bit_identical=%s worst_param_diff=%.3g allreduce_rel_err=%.3g"
              % (world, bit_identical, worst, err))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
