"""torchrun worker of tests/test_gpu_dist.py (one process per GPU, NCCL): the data-parallel introspective iteration.

Checks, on every rank:
  1. after introspective_iteration() the flat encoder / decoder gradient buffers hold the SUM over ranks of the gradients
     each rank computes alone from its own shard (the engine's kernels are deterministic, so the local gradients of a second
     model with the same init and inputs are reproduced bit for bit and all-gathered for the comparison);
  2. parameters, Adam moments and statistics are bit-identical on all ranks after the step (one all-reduce result feeds the
     same Adam kernel everywhere) while BatchNorm running statistics stay rank-local;
  3. the CUDA-graph replay of the step leaves the same state as eager execution over several iterations, in both
     collective modes: "lib" (default: ONE graph per iteration, the two all-reduces are raw ncclAllReduce calls of the
     library's own communicator captured inside it) and "torch" (round-1 path: three graph segments, torch.distributed
     all-reduces launched eagerly between them);
  4. SURVEY 8(e)'s N-GPU oracle: every rank runs the ORACLE's E half on its own shard from the common state_dict, the
     gradients are averaged over ranks, one oracle Adam step, then the same for the D half -- the engine's all-reduced
     gradients (x 1/world) must match the averaged oracle gradients, its post-step parameters the oracle's (within the
     first-Adam-step bound), and its rank-local BatchNorm buffers the oracle's rank-local ones;
  5. Engine.broadcast_state makes replicas with different initial weights identical to rank 0.
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

    # ---- 3. graph replay == eager, both collective modes --------------------------------------------------------
    report = []
    for comm_mode, n_graphs in (("lib", 1), ("torch", 3)):
        os.environ["SIVAE_DP_COMM"] = comm_mode
        outs = []
        for use_graph in (False, True):
            m = fresh()
            stats = []
            for i in range(iters):
                st = M.introspective_iteration(m, reals[i], noises[i], epss[i], hp, 2e-4, 2e-4, use_graph=use_graph)
                stats.append(st.clone())
            torch.cuda.synchronize()
            if use_graph:
                assert len(m._engine._graphs) == n_graphs, "mode %s: expected %d captured graph(s), got %d (failed: %s)" % (
                    comm_mode, n_graphs, len(m._engine._graphs), m._engine._graph_failed)
            if comm_mode == "lib":
                assert m._engine.comm_world == world, "the library-owned communicator was not created"
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
        report.append("%s:bit_identical=%s,worst=%.3g" % (comm_mode, all(torch.equal(sd_a[k], sd_b[k]) for k in sd_a), worst))
        dist.barrier()
    os.environ["SIVAE_DP_COMM"] = "lib"

    # ---- 4. the N-GPU oracle of SURVEY 8(e) ----------------------------------------------------------------------
    from oracle import sivae_oracle as O
    from tests.step_harness import rel_l2
    arch = O.Arch(**cfg)
    m = fresh()
    init = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    sd = O.clone_sd(init, torch.float64)
    ohp = O.Hyper(beta_kl=1.0, beta_rec=1.0, beta_neg=256.0, gamma_r=1e-8, scale=1.0 / (3 * 32 * 32), lr_e=2e-4, lr_d=2e-4)
    real64, noise64 = reals[0].cpu().double(), noises[0].cpu().double()
    eps64 = [e.cpu().double() for e in epss[0]]

    def rank_mean(grads):
        out = {}
        for k, g_ in grads.items():
            t = g_.to(dev)
            dist.all_reduce(t)
            out[k] = (t / world).cpu()
        return out
    se, ge, z, _ = O.e_step(sd, arch, real64, noise64, eps64[:3], ohp)
    ge_mean = rank_mean(ge)
    st_e, st_d = O.AdamState(), O.AdamState()
    O.adam_update(sd, ge_mean, st_e, ohp.lr_e, ohp)
    M.introspective_iteration(m, reals[0], noises[0], epss[0], hp, 2e-4, 2e-4, use_graph=False)
    torch.cuda.synchronize()
    eng = m._engine
    got_ge = {"encoder." + n: p.grad.detach().cpu() / world for n, p in m.encoder.named_parameters()}
    worst_ge = max(rel_l2(got_ge[k], ge_mean[k]) for k in ge_mean)
    assert worst_ge < 3e-2, "engine mean encoder gradient vs averaged oracle gradients: rel-L2 %.3g" % worst_ge
    # D half of the oracle from the ENGINE's post-E encoder (teacher forcing, like the single-GPU parity tests)
    for n, p in m.encoder.named_parameters():
        d_ = float((p.detach().cpu().double() - sd["encoder." + n]).abs().max())
        assert d_ <= 2.05 * 2e-4 + 1e-7, "post-all-reduce encoder parameter %s differs from the oracle DP step by %.3g" % (n, d_)
        sd["encoder." + n] = p.detach().cpu().double().clone()
    sdd, gd, _ = O.d_step(sd, arch, real64, noise64, z, eps64[3:], ohp)
    gd_mean = rank_mean(gd)
    O.adam_update(sd, gd_mean, st_d, ohp.lr_d, ohp)
    got_gd = {"decoder." + n: p.grad.detach().cpu() / world for n, p in m.decoder.named_parameters()}
    worst_gd = max(rel_l2(got_gd[k], gd_mean[k]) for k in gd_mean)
    assert worst_gd < 3e-2, "engine mean decoder gradient vs averaged oracle gradients: rel-L2 %.3g" % worst_gd
    post = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            assert int(post[k]) == int(v), k
        elif k.endswith(("running_mean", "running_var")):      # rank-local statistics: each rank against ITS oracle
            r_ = float((post[k].double() - v).abs().max() / (v.abs().max() + 1e-12))
            assert r_ < 3e-2, "%s: rank-local BN buffer differs from this rank's oracle by %.3g" % (k, r_)
        elif k.startswith("decoder."):
            d_ = float((post[k].double() - v).abs().max())
            assert d_ <= 2.05 * 2e-4 + 1e-7, "post-all-reduce decoder parameter %s differs from the oracle DP step by %.3g" % (k, d_)
    st = eng.stats.cpu()
    for idx, key, ref in ((0, "loss_rec_e", se["loss_rec_e"]), (1, "lossE_real_kl", se["lossE_real_kl"]), (4, "lossE", se["lossE"]),
                          (5, "loss_rec", sdd["loss_rec"]), (10, "lossD", sdd["lossD"])):
        assert abs(float(st[idx]) - ref) <= 1e-4 * abs(ref), "rank %d scalar %s: engine %.8g oracle %.8g" % (rank, key, float(st[idx]), ref)

    # ---- 5. broadcast_state: replicas built from different seeds become rank 0's ------------------------------------
    torch.manual_seed(1000 + rank)
    stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
    try:
        c = M.SoftIntroVAE(**cfg).to(dev)
    finally:
        sys.stdout = stdout
    ec = c.reserve(B)
    ec.broadcast_state(dist, src=0)
    for net in (L.NET_ENCODER, L.NET_DECODER):
        for name in ("params", "m", "v", "bn", "nbt"):
            t = getattr(ec.mem[net], name)
            ref = t.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(t, ref), "broadcast_state: %s of net %d differs from rank 0 on rank %d" % (name, net, rank)
    dist.barrier()
    torch.cuda.synchronize()
    E.Engine.comm_finalize()
    if rank == 0:
        print("DIST_OK world=%d graph_vs_eager[%s] allreduce_rel_err=%.3g dp_oracle_grad_rel_l2=(%.3g, %.3g)"
              % (world, " ".join(report), err, worst_ge, worst_gd))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
