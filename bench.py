#!/usr/bin/env python
"""bench.py -- images/sec of the Soft-IntroVAE introspective E+D step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--config H|M|C|Bs] [--impl reference]

A "step" = one full introspective iteration: E half (4 decoder + 3 encoder forwards, loss, backward), NCCL all-reduce
of the flat encoder gradient, Adam(encoder), D half (4 decoder + 2 encoder forwards, loss, backward), all-reduce of the
decoder gradient, Adam(decoder) -- reference soft_intro_vae/train_soft_intro_vae.py:542-624.
Default workload: config H = CelebA-HQ-shaped synthetic 256x256x3, z=512, channels [64,128,256,512,512,512],
batch 32 per GPU (weak scaling), random-init weights, uniform [0,1) images.

Output: ONE JSON line on rank 0 (contract in the task description): value = whole-job images/s with inputs resident
in HBM (CUDA-event timed, max over ranks); e2e = same metric through the public Python API with host (pinned) inputs
and the statistics read back every step; roofline of the dominant kernel class (tcgen05 conv) from CUDA events
recorded around every convolution launch inside the timed region; cpu_baseline = the oracle (CPU restatement of the
reference step) timed on the host cores on a bounded sample.
`--impl reference` times the reference's CPU implementation of the path (the oracle port; the unmodified reference
needs /root/reference which does not exist on the GPU box) on the host cores with the same JSON shape.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "soft-intro-vae-pytorch_b200"

CONFIGS = {
    # name: (image_size, zdim, channels, per-GPU batch, beta_neg, bootstrap, GFLOP per image per iteration [SURVEY 8d])
    "C": (32, 128, [64, 128, 256], 128, 256.0, False, 9.80),
    "M": (128, 256, [64, 128, 256, 512, 512], 64, 256.0, False, 204.0),
    "H": (256, 512, [64, 128, 256, 512, 512, 512], 32, 1024.0, False, 820.6),
    "Bs": (256, 512, [64, 128, 256, 512, 512, 512], 16, 1024.0, True, 770.3),
}
WORKLOAD_NAMES = {"C": "CIFAR-10-shaped synthetic 32x32x3 z128 [64,128,256]",
                  "M": "CelebA-shaped synthetic 128x128x3 z256 [64,128,256,512,512]",
                  "H": "CelebA-HQ-shaped synthetic 256x256x3 z512 [64,128,256,512,512,512]",
                  "Bs": "FFHQ-shaped synthetic 256x256x3 bootstrap (target decoder)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_burst=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(self.rows))


def build_model(cfg_name, device, batch):
    import torch
    size, zdim, channels, _, beta_neg, boot, _ = CONFIGS[cfg_name]
    mod = importlib.import_module(PKG + (".train_soft_intro_vae_bootstrap" if boot else ".train_soft_intro_vae"))
    torch.manual_seed(0)
    stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
    try:
        model = mod.SoftIntroVAE(cdim=3, zdim=zdim, channels=channels, image_size=size).to(device)
    finally:
        sys.stdout = stdout
    model.train()
    model.reserve(batch)
    return mod, model


DTYPE_NAMES = {
    0: "bf16x2-compensated forward convs (operands = bf16 hi+lo pairs, 3 kind::f16 MMAs per product, fp32 accumulate; every logged "
       "scalar within 1e-4 of the fp64 oracle) + bf16 dgrad/wgrad of the residual blocks (tf32 for the image-facing convs); "
       "fp32 conv outputs / BN / loss / Adam / master weights",
    4: "tf32 operands (fp32 storage, fp32 accumulate, fp32 BN/loss/Adam)",
    3: "3xTF32 compensated (three tf32 convs per product)", 1: "fp32 CUDA cores (exact path)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    L = importlib.import_module(PKG + ".lib")
    E = importlib.import_module(PKG + ".engine")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    size, zdim, channels, batch, beta_neg, boot, gflop_img = CONFIGS[args.config]
    batch = args.batch or batch
    mod, model = build_model(args.config, device, batch)
    eng = model._engine
    lib = L.load()
    hp = E.make_hyper(1.0, 1.0, beta_neg, 1.0 if boot else 1e-8, 1.0 / (3 * size * size))
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4        # distinct host batches cycled through (inputs larger than L2 at config H: 4 x 25 MB)
    host_real = [torch.rand(batch, 3, size, size, generator=g).pin_memory() for _ in range(n_host)]
    dev_real = [h.to(device) for h in host_real]
    noise_d = torch.randn(batch, zdim, device=device)
    eps_d = torch.randn(5, batch, zdim, device=device)
    stats_host = torch.empty(16).pin_memory()
    noise_host = [torch.empty(batch, zdim).pin_memory() for _ in range(2)]
    stats_host2 = [torch.empty(16).pin_memory() for _ in range(2)]

    def step_resident(i, use_graph=None):
        mod.introspective_iteration(model, dev_real[i % n_host], noise_d, eps_d, hp, 2e-4, 2e-4, use_graph=use_graph)

    class _HostBatches:
        """the loader of the e2e leg: pinned host batches, cycled"""
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield host_real[i % n_host]

    e2e_prefetch = os.environ.get("SIVAE_PREFETCH", "1") != "0"          # the trainer's defaults
    e2e_async = os.environ.get("SIVAE_ASYNC_STATS", "0") == "1"

    def run_e2e(steps, prefetch=None, defer=None):
        """what the drop-in trainer does per iteration (train_soft_intro_vae.py _run_training): the batch moves host->device
        (prefetch: batch i+1 on a side stream while step i runs, DevicePrefetcher), noise from the CPU generator (:547)
        through a pinned buffer, the five eps draws on the device, the step, and the logged statistics copied back every
        step (:628) -- read right away, or (defer) like the trainer once the next step has been queued"""
        prefetch = e2e_prefetch if prefetch is None else prefetch
        defer = e2e_async if defer is None else defer
        src = mod.DevicePrefetcher(_HostBatches(steps), device) if prefetch else _HostBatches(steps)
        pend = None
        for i, real in enumerate(src):
            if not prefetch:
                real = real.to(device, non_blocking=True)
            nh = noise_host[i % 2]
            torch.randn((batch, zdim), out=nh)
            noise = nh.to(device, non_blocking=True)
            for k in range(5):
                torch.randn((batch, zdim), out=eps_d[k])
            st = mod.introspective_iteration(model, real, noise, eps_d, hp, 2e-4, 2e-4)
            stats_host2[i % 2].copy_(st, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pend is not None:
                pend[1].synchronize()
                assert float(pend[0][15]) == 0.0, "NaN flag"
            pend = (stats_host2[i % 2], ev)
            if not defer:
                ev.synchronize()
                pend = None
        if pend is not None:
            pend[1].synchronize()
        stats_host.copy_(stats_host2[(steps - 1) % 2])
        return stats_host

    def time_e2e(steps, **kw):
        run_e2e(5, **kw)                       # warm-up: side-stream allocator pool, pinned staging, graph replay path
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        run_e2e(steps, **kw)
        eb.record()
        barrier()
        t = torch.tensor([ea.elapsed_time(eb)], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # the public call replays the step from a CUDA graph from its third invocation on (first eager, second captures):
    # count the kernels of one eager step, then warm up until the replay path is the one being timed
    launches_a = lib.sivae_launch_count()
    step_resident(0, use_graph=False)
    launches_per_step = lib.sivae_launch_count() - launches_a
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    graph_ms = None
    if args.graph:
        # optional: the same step replayed from a CUDA graph (removes ~1100 kernel-launch latencies per step at the
        # small configs).  Captured on a side stream; the Adam step counters live on the device, so replays advance.
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            step_resident(0, use_graph=False)
        for _ in range(2):
            g_.replay()
        graph_ms = timed(lambda i: g_.replay(), args.steps) / args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)                     # headline region: K steps, no per-launch events
    launches = launches_per_step * args.steps                        # kernels executed in the region (graph replays)
    # second pass over the same K steps with a CUDA-event pair around every conv / BN / loss launch (the events cost
    # ~3 % of the step, so they stay out of the headline region); roofline numbers come from this pass
    lib.sivae_profile_enable(1)
    ms_prof = timed(lambda i: step_resident(i + args.steps, use_graph=False), args.steps)
    buf = C.create_string_buffer(1 << 16)
    lib.sivae_profile_dump(buf, len(buf))
    rows = [l.split() for l in buf.value.decode().splitlines()]
    # BatchNorm+activation passes (classes 8 / 9 of the dump; their "gflop" column carries algorithmic GB): HBM-bound
    bn_cls = {}
    for r in rows:
        if int(r[0]) in (8, 9):
            a = bn_cls.setdefault(int(r[0]), [0.0, 0.0, 0])
            a[0] += float(r[8]); a[1] += float(r[9]); a[2] += int(r[7])
    if args.layers and rank == 0:
        names = {0: "tc tf32 (dgrad)", 1: "tc_wgrad", 2: "cuda-core fwd/dgrad", 3: "cuda-core wgrad", 4: "fused loss pass (GB, TB/s)",
                 5: "tc split32 fwd (3 MMAs/product)", 6: "tc bf16 dgrad", 7: "tc bf16 wgrad",
                 8: "bn+act fwd (k=4+mode: residual; TB/s)", 9: "bn+act bwd (k=4+mode: residual; TB/s)"}
        rows.sort(key=lambda r: -float(r[8]))
        with open(args.layers, "w") as f:
            f.write("| class | N | H | W | Cin | Cout | k | launches/step | ms/step | ms/launch | TFLOP/s |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                n, ms, gf = int(r[7]), float(r[8]), float(r[9])
                f.write("| %s | %s | %s | %s | %s | %s | %s | %.1f | %.3f | %.4f | %.1f |\n" % (
                    names[int(r[0])], r[1], r[2], r[3], r[4], r[5], r[6], n / args.steps, ms / args.steps, ms / n, gf / ms if ms > 0 else 0))
    prof = (C.c_double * 24)()
    lib.sivae_profile_read(prof)
    lib.sivae_profile_enable(0)
    ms_e2e = time_e2e(args.steps)
    e2e_sweep = None
    if args.e2e_sweep:
        e2e_sweep = {"prefetch=%d,defer=%d" % (p_, d_): round(time_e2e(args.steps, prefetch=bool(p_), defer=bool(d_)) / args.steps, 3)
                     for p_ in (0, 1) for d_ in (0, 1)}
    # secondary number, NOT the headline: the opt-in pass re-use (the D half takes fake / rec from the E half's identical
    # decoder passes instead of recomputing them; bit-identical results, tests/test_gpu_step.py).  Headline and e2e above
    # execute all 13 forward passes of the reference step.
    reuse_ms = None
    if not args.no_reuse_leg:
        def step_reuse(i):
            mod.introspective_iteration(model, dev_real[i % n_host], noise_d, eps_d, hp, 2e-4, 2e-4, reuse_decoder_passes=True)
        for i in range(max(args.warmup, 3)):
            step_reuse(i)
        reuse_ms = timed(step_reuse, args.steps) / args.steps
        model._engine.reuse_decoder_passes = False
    sampler.stop_flag = True
    st = stats_host.clone()
    ms_step = ms_total / args.steps
    value = world * batch / (ms_step / 1e3)
    e2e_value = world * batch / (ms_e2e / args.steps / 1e3)
    peaks = measured_peaks()
    # kernel classes (CUDA events per launch): 0 = tcgen05 kind::tf32 conv (dgrad; also the forward of the TF32 backend),
    # 1 = tcgen05 wgrad, 5 = tcgen05 forward conv on split32 operands (three kind::f16 MMAs per product).  The dominant class
    # is the one with the most time in the step.
    tc_ms, tc_flops, tc_n = prof[0], prof[1], prof[2]
    wg_ms, wg_flops, wg_n = prof[3], prof[4], prof[5]
    f3_ms, f3_flops, f3_n = prof[15], prof[16], prof[17]
    d16_ms, d16_flops, d16_n = prof[18], prof[19], prof[20]
    w16_ms, w16_flops, w16_n = prof[21], prof[22], prof[23]
    simt_ms = prof[6] + prof[9]
    roof = None

    def tcls(ms, fl, n, mma_per_product=1):
        if not (n > 0 and ms > 0):
            return None
        a = fl / (ms * 1e-3) / 1e12
        return dict(achieved=round(a, 2), frac=round(a / peaks["bf16_sustained"], 4), ms_per_step=round(ms / args.steps, 3),
                    launches_per_step=n / args.steps, share_of_step=round(ms / ms_prof, 4),
                    tensor_pipe_issued_tflops=round(a * mma_per_product, 2),
                    tensor_pipe_issued_frac=round(a * mma_per_product / peaks["bf16_sustained"], 4))
    if (tc_n > 0 and tc_ms > 0) or (f3_n > 0 and f3_ms > 0):
        split_mode = f3_n > 0 and f3_ms >= tc_ms
        dom = tcls(f3_ms, f3_flops, f3_n, 3) if split_mode else tcls(tc_ms, tc_flops, tc_n)
        roof = dict(bound="tensor",
                    kernel=("k_conv_halo2<.., FMT_SPLIT> family (tcgen05 kind::f16 implicit-GEMM FORWARD conv on bf16 hi+lo operands, "
                            "3 MMAs per product)" if split_mode else "k_conv_halo2 family (tcgen05 kind::tf32 implicit-GEMM conv, fwd+dgrad)"),
                    achieved=dom["achieved"], peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=dom["frac"],
                    peak_note="MEASURED_PEAKS bf16_tflops_sustained (%s).  achieved = ALGORITHMIC conv FLOPs / time; the compensated "
                              "forward issues 3 bf16 MMAs per product (frac <= 0.333 by construction, tensor_pipe_issued_* = 3x), "
                              "kind::tf32 classes issue at half the bf16 rate (frac <= 0.5)" % peaks["source"],
                    launches_per_step=dom["launches_per_step"], ms_per_step=dom["ms_per_step"], share_of_step=dom["share_of_step"],
                    tensor_pipe_issued_tflops=dom["tensor_pipe_issued_tflops"], tensor_pipe_issued_frac=dom["tensor_pipe_issued_frac"],
                    ms_per_step_with_events=round(ms_prof / args.steps, 3),
                    traffic=_traffic("fwd"), traffic_detail=_traffic("fwd_detail"),
                    fwd_split32=tcls(f3_ms, f3_flops, f3_n, 3), tf32_conv=tcls(tc_ms, tc_flops, tc_n),
                    dgrad_bf16=tcls(d16_ms, d16_flops, d16_n), wgrad_bf16=tcls(w16_ms, w16_flops, w16_n),
                    wgrad_tf32=tcls(wg_ms, wg_flops, wg_n),
                    conv_stack=dict(achieved=round((tc_flops + wg_flops + f3_flops + d16_flops + w16_flops) /
                                                   ((tc_ms + wg_ms + f3_ms + d16_ms + w16_ms) * 1e-3) / 1e12, 2),
                                    ms_per_step=round((tc_ms + wg_ms + f3_ms + d16_ms + w16_ms) / args.steps, 3)),
                    simt_conv_ms_per_step=round(simt_ms / args.steps, 3))
        for cls, key, kern in ((8, "bn_fwd", "k_bn_act_fwd (BN apply + residual + LeakyReLU + pool / upsample)"),
                               (9, "bn_bwd", "k_bn_bwd_reduce + k_bn_bwd_apply (train-mode BN + LeakyReLU backward)")):
            if cls in bn_cls and bn_cls[cls][0] > 0:
                ms_c, gb_c, n_c = bn_cls[cls]
                gbs = gb_c / (ms_c * 1e-3)
                roof[key] = dict(bound="hbm", kernel=kern, achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                                 frac=round(gbs / peaks["hbm_gbs"], 4), ms_per_step=round(ms_c / args.steps, 3),
                                 launches_per_step=n_c / args.steps, share_of_step=round(ms_c / ms_prof, 4),
                                 note="algorithmic bytes = full-tensor passes each launch must make (DESIGN.md section 5)")
        if prof[14] > 0 and prof[12] > 0:
            gbs = prof[13] / (prof[12] * 1e-3) / 1e9
            roof["loss_pass"] = dict(bound="hbm", kernel="k_mse3_partial (fused 3x per-sample MSE over five images)",
                                     achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s", frac=round(gbs / peaks["hbm_gbs"], 4),
                                     ms_per_launch=round(prof[12] / prof[14], 4), bytes_per_launch=int(prof[13] / prof[14]))
    line = dict(metric="images/sec per introspective E+D step", value=round(value, 2), unit="images/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(ms_step, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype=DTYPE_NAMES.get(int(model._conv_backend), "?"), data="synthetic",
                config=dict(workload=WORKLOAD_NAMES[args.config], image_size=size, z_dim=zdim, channels=channels,
                            batch_per_gpu=batch, global_batch=batch * world, parallelism="dp%d" % world,
                            l2_policy="inputs cycle over %d distinct batches; activations per step (>20 GB) far exceed the 126 MB L2" % n_host,
                            algorithmic_gflop_per_image=gflop_img,
                            step_tflops=round(gflop_img * batch * world / ms_step, 2)),
                e2e=dict(value=round(e2e_value, 2), unit="images/s", ms_per_step=round(ms_e2e / args.steps, 3),
                         h2d_bytes_per_step=batch * 3 * size * size * 4 + batch * zdim * 4, d2h_bytes_per_step=64),
                gpu_launches=int(launches),
                launch_mode=(("cuda graph replay" if world == 1 else ("replay of ONE cuda graph per iteration, the two ncclAllReduce calls of the library-owned communicator captured inside"
                                                                    if eng.comm_world == world else "replay of 3 cuda graph segments, torch.distributed all-reduces eager between them"))
                             if importlib.import_module(PKG + ".train_soft_intro_vae")._graph_mode() >= 1 else "eager") +
                            " of %d kernels per step (counted on an eager step)" % launches_per_step,
                roofline=roof, clocks=sampler.summary() if rank == 0 else None,
                last_stats=dict(loss_rec=float(st[5]), kl_real=float(st[1]), lossE=float(st[4]), lossD=float(st[10])))
    line["e2e"]["mode"] = "prefetch=%d,defer_stats=%d" % (int(e2e_prefetch), int(e2e_async))
    if e2e_sweep is not None:
        line["e2e"]["sweep_ms_per_step"] = e2e_sweep
    if reuse_ms is not None:
        line["decoder_pass_reuse"] = dict(in_headline=False, ms_per_step=round(reuse_ms, 3),
                                          value=round(world * batch / (reuse_ms / 1e3), 2), unit="images/s",
                                          note="opt-in (SIVAE_REUSE_DEC=1 / reuse_decoder_passes=True): D half re-uses the E half's "
                                               "fake/rec decoder passes, BN running-stat updates replayed; bit-identical results")
    if graph_ms is not None:
        line["cuda_graph"] = dict(ms_per_step=round(graph_ms, 3), value=round(world * batch / (graph_ms / 1e3), 2))
    if rank == 0 and world == 1 and not args.no_loader_leg and size >= 128:
        try:
            line["image_loader"] = image_loader_leg(device, size, batch)
        except Exception as ex:                       # a secondary leg must never take the headline down with it
            line["image_loader"] = dict(error=repr(ex)[:300])
    if rank == 0 and world == 1 and not args.no_stock_leg:
        try:
            torch.cuda.empty_cache()
            line["stock_pytorch_same_gpu"] = stock_torch_leg(args.config, device)
            line["stock_pytorch_same_gpu"]["speedup_of_this_engine"] = round(line["stock_pytorch_same_gpu"]["ms_per_step"] / ms_step, 2)
        except Exception as ex:                       # a secondary leg must never take the headline down with it
            line["stock_pytorch_same_gpu"] = dict(error=repr(ex)[:300])
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.config, budget_s=args.cpu_budget)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        barrier()
        # the library's own communicator is deliberately NOT destroyed here: the process is about to exit and the OS reclaims it
        # (an explicit teardown exists -- Engine.comm_finalize -- and is exercised by tests/dist_worker.py)
        dist.destroy_process_group()


def image_loader_leg(device, out_size, batch, steps=20, warmup=3, src_size=1024):
    """Secondary kernel: batch assembly of decoded 8-bit images -- mirror + Pillow-exact bicubic resize
    + ToTensor -- for the workload's loader geometry (CelebA-HQ / FFHQ sources are 1024x1024).  HBM-bound byte work:
    algorithmic bytes = source pixels once + float output once."""
    import numpy as np
    import torch
    G = importlib.import_module(PKG + ".gpu_dataset")
    peaks = measured_peaks()
    n_host = 4                                   # 4 x 100 MB distinct source batches: every launch reads cold data
    g = torch.Generator().manual_seed(99)
    host = [torch.randint(0, 256, (batch, src_size, src_size, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_host)]
    dev = [h.to(device) for h in host]
    flags = (torch.arange(batch) % 2).to(torch.uint8)
    flags_d = flags.to(device)
    bt = G.ImageBatcher(out_size, out_size, device)
    out = torch.empty(batch, 3, out_size, out_size, dtype=torch.float32, device=device)

    def timed(fn):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps
    ms = timed(lambda i: bt(dev[i % n_host], flags_d, out=out))
    ms_e2e = timed(lambda i: bt(host[i % n_host], flags, out=out))          # pinned host pixels -> device inside the region
    nbytes = batch * src_size * src_size * 3 + batch * 3 * out_size * out_size * 4
    gbs = nbytes / (ms * 1e-3) / 1e9
    # the reference's own implementation of the same work on one host core (main.py:47 passes num_workers=0): Pillow + ToTensor
    from PIL import Image
    a = host[0][0].numpy()
    t0, n_cpu = time.time(), 0
    while time.time() - t0 < 3.0:
        im = Image.fromarray(a, "RGB")
        if n_cpu % 2:
            im = im.transpose(Image.FLIP_LEFT_RIGHT)
        r = np.asarray(im.resize((out_size, out_size), Image.BICUBIC))
        torch.from_numpy(r.copy()).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
        n_cpu += 1
    cpu_ips = n_cpu / (time.time() - t0)
    return dict(kernel="k_image_batch<3> (mirror + Pillow-exact bicubic %dx%d -> %dx%d + ToTensor, batch %d)" % (src_size, src_size, out_size, out_size, batch),
                value=round(batch / (ms * 1e-3), 1), unit="images/s", ms_per_batch=round(ms, 4),
                e2e=dict(value=round(batch / (ms_e2e * 1e-3), 1), unit="images/s", ms_per_batch=round(ms_e2e, 4),
                         h2d_bytes_per_step=batch * src_size * src_size * 3 + batch, d2h_bytes_per_step=0),
                roofline=dict(bound="hbm", achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s", frac=round(gbs / peaks["hbm_gbs"], 4),
                              bytes_per_launch=nbytes, traffic=_traffic("image"), traffic_detail=_traffic("image_detail"),
                              note="instruction-bound, not HBM-bound: 20 M warp-level integer MACs per batch (17-tap bicubic at 4x "
                                   "down-scaling) each need a byte load and an IMAD; see DESIGN.md section 5"),
                cpu_baseline=dict(value=round(cpu_ips, 1), unit="images/s", cores=1, kind="reference",
                                  sample="Pillow %s resize + ToTensor of one %dx%d image repeated for 3 s on one core "
                                         "(the reference's loader runs in the training process, num_workers=0)" % (__import__("PIL").__version__, src_size, src_size)))


def _traffic(kind):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kind)
        except Exception:
            return None
    return None


def cpu_baseline(cfg_name, budget_s=25.0, steps=1, warmup=1):
    """The reference's CPU implementation of the step on the host cores, on a bounded sample (reduced batch) of the workload.
    kind "reference": the UNMODIFIED reference trainer from oracle/_ref (oracle/build_ref.py; its own train_soft_intro_vae()
    loop body timed between two batch requests, oracle/ref_arm.py); kind "port" (only when oracle/_ref is absent): the oracle's
    functional restatement.  The batch is sized adaptively: a WARM probe (one untimed + one timed iteration at the smallest
    legal batch, so a cold oneDNN / thread pool cannot shrink the sample) measures seconds per image on this host, then
    `warmup` + `steps` iterations run at the largest batch (<= the workload's) whose iteration fits the budget (~10-30 s)."""
    import torch
    from oracle import ref_arm
    size, zdim, channels, batch, beta_neg, boot, gflop_img = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    use_ref = ref_arm.reference_root() is not None and os.environ.get("SIVAE_CPU_BASELINE", "reference") != "port"
    if use_ref:
        def run(b, n_warm, n_it):
            return ref_arm.time_reference_iterations(size, zdim, b, beta_neg, boot, warmup=n_warm, steps=n_it, threads=cores)[0]
    else:
        from oracle import sivae_oracle as O
        arch = O.Arch(cdim=3, zdim=zdim, channels=channels, image_size=size)
        sd = O.make_state_dict(arch, seed=0, bootstrap=boot)
        hp = O.Hyper(beta_neg=beta_neg, gamma_r=1.0 if boot else 1e-8, scale=1.0 / (3 * size * size))

        def run(b, n_warm, n_it):
            g = torch.Generator().manual_seed(1234)
            real = torch.rand(b, 3, size, size, generator=g)
            noise = torch.randn(b, zdim, generator=g)
            eps = list(torch.randn(5, b, zdim, generator=g))
            se, sdd = O.AdamState(), O.AdamState()
            for _ in range(n_warm):
                O.full_iteration(sd, arch, real, noise, eps, hp, se, sdd, boot)
            t0 = time.time()
            for _ in range(n_it):
                O.full_iteration(sd, arch, real, noise, eps, hp, se, sdd, boot)
            return (time.time() - t0) / n_it

    b0 = 1 if size > 64 else 2                      # train-mode BN at 4x4 needs more than one value per channel
    probe = run(b0, 1, 1)                           # warm probe: seconds per iteration at the smallest batch
    b = int(max(b0, min(batch, budget_s / (probe / b0))))
    if b > 4 * b0:                                  # small batches under-use the cores: re-estimate at a quarter of the target
        bm = max(b0, b // 4)
        b = int(max(b0, min(batch, budget_s / (run(bm, 0, 1) / bm))))
    dt = run(b, warmup, steps)
    return dict(value=round(b / dt, 4), unit="images/s", cores=cores, kind="reference" if use_ref else "port",
                sample="%d timed (+%d warm-up) full E+D iteration(s) of %s at batch %d (of %d) on torch CPU fp32, %d threads; %.1f s/iter "
                       "(batch grown from a warm %.1f s probe at batch %d towards a %.0f s budget)"
                       % (steps, warmup, "the unmodified reference train_soft_intro_vae() (oracle/_ref)" if use_ref else "the oracle port",
                          b, batch, cores, dt, probe, b0, budget_s))


def stock_torch_leg(cfg_name, device, steps=3, warmup=2):
    """The honest same-GPU comparator: the reference's arithmetic through STOCK PyTorch on this B200 -- cuDNN
    convolutions / batch norm, cuBLAS, ATen elementwise, autograd -- at torch's default flags (cudnn.allow_tf32 = True, i.e. TF32
    convolutions: what the unmodified reference does on this GPU under torch 2.11).  Runs the oracle's functional restatement
    (test infrastructure, timed here only as the thing being compared AGAINST) with the workload's shapes on the device."""
    import torch
    from oracle import ref_arm
    from oracle import sivae_oracle as O
    size, zdim, channels, batch, beta_neg, boot, gflop_img = CONFIGS[cfg_name]
    flags = "torch.backends.cudnn.allow_tf32 = %s, matmul.allow_tf32 = %s: torch defaults = what the unmodified reference runs" % (
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    if ref_arm.reference_root() is not None:
        # the UNMODIFIED reference trainer itself (oracle/_ref), device = this GPU
        dt, _ = ref_arm.time_reference_iterations(size, zdim, batch, beta_neg, boot, warmup=warmup, steps=steps, device=str(device))
        ms = dt * 1e3
        return dict(what="the UNMODIFIED reference train_soft_intro_vae() (oracle/_ref) on this GPU through stock PyTorch %s / cuDNN %s "
                         "(%s); full batch of %d, its own loop incl. the per-iteration .item() reads" % (
                             torch.__version__, torch.backends.cudnn.version(), flags, batch),
                    ms_per_step=round(ms, 2), value=round(batch / (ms * 1e-3), 2), unit="images/s", steps=steps, warmup=warmup)
    arch = O.Arch(cdim=3, zdim=zdim, channels=channels, image_size=size)
    sd = {k: v.to(device) for k, v in O.make_state_dict(arch, seed=0, bootstrap=boot).items()}
    hp = O.Hyper(beta_neg=beta_neg, gamma_r=1.0 if boot else 1e-8, scale=1.0 / (3 * size * size))
    g = torch.Generator().manual_seed(1234)
    real = torch.rand(batch, 3, size, size, generator=g).to(device)
    noise = torch.randn(batch, zdim, generator=g).to(device)
    eps = [e.to(device) for e in torch.randn(5, batch, zdim, generator=g)]
    se, sdd = O.AdamState(), O.AdamState()
    for _ in range(warmup):
        O.full_iteration(sd, arch, real, noise, eps, hp, se, sdd, boot)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        O.full_iteration(sd, arch, real, noise, eps, hp, se, sdd, boot)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    del sd
    torch.cuda.empty_cache()
    return dict(what="the same introspective iteration through stock PyTorch %s / cuDNN %s on this GPU (autograd, NCHW fp32 tensors, "
                     "torch.backends.cudnn.allow_tf32 = %s, matmul.allow_tf32 = %s: torch defaults = what the unmodified reference runs)"
                     % (torch.__version__, torch.backends.cudnn.version(), torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32),
                ms_per_step=round(ms, 2), value=round(batch / (ms * 1e-3), 2), unit="images/s", steps=steps, warmup=warmup)


def run_reference(args):
    """reference arm: the reference's own CPU implementation of the path (the unmodified trainer from oracle/_ref), all host
    threads, on a bounded sample of the workload; rank 0 only"""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    size, zdim, channels, batch, beta_neg, boot, gflop_img = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", 1))
    t0 = time.time()
    # bounded: every step is one iteration on the sample batch (~cpu_budget seconds each): cap the counts to stay within minutes
    steps = max(1, min(args.steps, 3 if size <= 64 else 2))
    warmup = 1 if args.warmup > 0 else 0
    cb = cpu_baseline(args.config, budget_s=args.cpu_budget, steps=steps, warmup=warmup)
    line = dict(impl="reference", metric="images/sec per introspective E+D step", value=cb["value"], unit="images/s",
                n_gpus=world, steps=steps, warmup=warmup, ms_per_step=None, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="fp32", data="synthetic",
                config=dict(workload=WORKLOAD_NAMES[args.config], image_size=size, z_dim=zdim, channels=channels,
                            batch_per_gpu=batch, note="CPU path: bounded sample, see cpu_baseline.sample"),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=round(time.time() - t0, 1))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="H", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock-leg", action="store_true", help="skip the stock-PyTorch-on-the-same-GPU comparator")
    ap.add_argument("--e2e-sweep", action="store_true", help="also time the e2e leg with / without batch prefetch and deferred statistics reads")
    ap.add_argument("--no-loader-leg", action="store_true", help="skip the image batch-assembly kernel timing")
    ap.add_argument("--no-reuse-leg", action="store_true", help="skip the secondary timing with decoder-pass re-use")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed from a CUDA graph")
    ap.add_argument("--layers", default="", help="write the per-conv-shape timing table (CUDA events) to this file")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    run_ours(args)


if __name__ == "__main__":
    main()
