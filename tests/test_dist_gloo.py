"""world_size-2 gloo test (CPU) of the data-parallel host logic of the trainer: the flat encoder / decoder gradient
buffers are sum-all-reduced once per half step and the Adam kernel receives grad_scale = 1/world, i.e. every rank ends
up with the mean of the per-rank gradients.  The native engine is replaced by a recording stub (no GPU
here); the real NCCL path is exercised by bench.py --gpus N on the GPU box."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.step_harness import PKG


class _Mem:
    def __init__(self, n):
        self.grads = torch.zeros(n)


class _StubEngine:
    """records calls; e_step / d_step deposit rank-dependent gradients like the engine's backward would"""

    def __init__(self, rank):
        L = importlib.import_module(PKG + ".lib")
        self.rank = rank
        self.mem = {L.NET_ENCODER: _Mem(7), L.NET_DECODER: _Mem(5)}
        self.calls = []
        self.stats = torch.zeros(16)
        self.L = L

    def e_step(self, real, noise, eps, hp):
        self.mem[self.L.NET_ENCODER].grads.copy_(torch.arange(7.0) * (self.rank + 1))
        self.calls.append("e_step")

    def d_step(self, eps, hp):
        self.mem[self.L.NET_DECODER].grads.copy_(torch.arange(5.0) + 10 * self.rank)
        self.calls.append("d_step")

    def vae_step(self, real, eps, hp):
        self.e_step(real, None, eps, hp)
        self.d_step(eps, hp)

    def adam(self, net, lr, grad_scale=1.0):
        self.calls.append(("adam", net, lr, grad_scale, self.mem[net].grads.clone()))

    def graphed(self, key, inputs, fn):
        """Engine.graphed without a GPU: run the segment, remember its key"""
        self.calls.append(("segment", key[0]))
        return fn(*inputs)


class _StubModel:
    def __init__(self, rank):
        self.eng = _StubEngine(rank)

    def _ensure_engine(self, batch):
        return self.eng


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = importlib.import_module(PKG + ".train_soft_intro_vae")
        model = _StubModel(rank)
        real = torch.zeros(2, 3, 8, 8)
        M.introspective_iteration(model, real, torch.zeros(2, 4), torch.zeros(5, 2, 4), None, 1e-3, 2e-3)
        calls = model.eng.calls
        assert [c if isinstance(c, str) else c[0] for c in calls] == ["e_step", "adam", "d_step", "adam"]
        _, net_e, lr_e, gs_e, g_e = calls[1]
        _, net_d, lr_d, gs_d, g_d = calls[3]
        assert (net_e, net_d) == (0, 1) and (lr_e, lr_d) == (1e-3, 2e-3)
        assert gs_e == gs_d == 1.0 / world
        # sum over ranks of arange*(rank+1) = arange*3 ; mean (after grad_scale) = arange*1.5
        assert torch.allclose(g_e, torch.arange(7.0) * 3) and torch.allclose(g_e * gs_e, torch.arange(7.0) * 1.5)
        assert torch.allclose(g_d, 2 * torch.arange(5.0) + 10.0)
        # the graph-segmented form used on the GPU under torch.distributed: same order, all-reduces between the segments
        model3 = _StubModel(rank)
        M._segmented_step(model3.eng, dist, real, torch.zeros(2, 4), torch.zeros(5, 2, 4), None, 1e-3, 2e-3, 1.0 / world, ("k",))
        names = [c if isinstance(c, str) else (c[1] if c[0] == "segment" else c[0]) for c in model3.eng.calls]
        assert names == ["introspective/E", "e_step", "introspective/D", "adam", "d_step", "introspective/A", "adam"], names
        seg_adams = [c for c in model3.eng.calls if not isinstance(c, str) and c[0] == "adam"]
        assert torch.allclose(seg_adams[0][4], torch.arange(7.0) * 3) and seg_adams[0][3] == 0.5
        assert torch.allclose(seg_adams[1][4], 2 * torch.arange(5.0) + 10.0) and seg_adams[1][1] == 1
        model2 = _StubModel(rank)
        M.vae_iteration(model2, real, torch.zeros(2, 4), None, 1e-3, 1e-3)
        adams = [c for c in model2.eng.calls if not isinstance(c, str)]
        assert len(adams) == 2 and all(a[3] == 0.5 for a in adams)
        assert torch.allclose(adams[0][4], torch.arange(7.0) * 3)
        out.put((rank, "ok"))
    except Exception as e:            # surface the failure to the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
