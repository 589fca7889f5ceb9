"""Torch-side owner of the memory the native engine borrows.

torch is plumbing here: it allocates the flat fp32 parameter / gradient / Adam / BN buffers and the workspace
arena, exposes them as ``nn.Parameter`` views with the reference's logical shapes (conv filters are
``[Cout,Cin,kh,kw]`` views with channels_last strides over the engine's ``[Cout][kh][kw][Cin]`` storage), supplies
the CUDA stream, and (multi-GPU) all-reduces the flat gradient buffers over NCCL.  All compute is in
libsivae_b200.so.
"""
import ctypes as C

import torch

from . import lib as L


def make_hyper(beta_kl, beta_rec, beta_neg, gamma_r, scale):
    return L.Hyper(float(beta_kl), float(beta_rec), float(beta_neg), float(gamma_r), float(scale))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class NetMemory:
    """Flat buffers of one net (encoder / decoder / target decoder)."""

    def __init__(self, handle, net, device, trainable=True):
        lib = L.load()
        self.net = net
        n = lib.sivae_param_count(handle, net)
        self.params = torch.zeros(n, dtype=torch.float32, device=device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=device) if trainable else None
        self.m = torch.zeros(n, dtype=torch.float32, device=device) if trainable else None
        self.v = torch.zeros(n, dtype=torch.float32, device=device) if trainable else None
        self.bn = torch.zeros(lib.sivae_bn_floats(handle, net), dtype=torch.float32, device=device)
        self.nbt = torch.zeros(max(1, lib.sivae_num_bn(handle, net)), dtype=torch.int64, device=device)
        self.tensors = []
        for i in range(lib.sivae_num_tensors(handle, net)):
            ti = L.TensorInfo()
            L.check(lib.sivae_tensor(handle, net, i, C.byref(ti)), "sivae_tensor")
            self.tensors.append((ti.name.decode(), ti.kind, ti.offset, ti.numel, tuple(ti.shape[:ti.ndim])))
        self.bns = []
        for i in range(lib.sivae_num_bn(handle, net)):
            bi = L.BnInfo()
            L.check(lib.sivae_bn(handle, net, i, C.byref(bi)), "sivae_bn")
            self.bns.append((bi.name.decode(), bi.channels, bi.bn_offset, bi.index))

    @staticmethod
    def view(flat, kind, off, numel, shape):
        v = flat[off:off + numel]
        if kind == L.T_CONV:
            co, ci, kh, kw = shape
            return v.view(co, kh, kw, ci).permute(0, 3, 1, 2)      # logical [Cout,Cin,kh,kw], channels_last strides
        return v.view(*shape)


class Engine:
    """One native engine instance bound to the modules of a SoftIntroVAE."""

    def __init__(self, cdim, zdim, channels, image_size, max_batch, device, bootstrap=False, conv_backend=L.CONV_AUTO,
                 cond_dim=0):
        lib = L.load()
        if torch.device(device).type != "cuda":
            raise RuntimeError("the B200 engine runs on CUDA devices only (no CPU fallback); got %s" % (device,))
        self.device = torch.device(device)
        self.cfg = L.Config()
        self.cfg.cdim, self.cfg.zdim, self.cfg.image_size = int(cdim), int(zdim), int(image_size)
        self.cfg.n_channels = len(channels)
        for i, c in enumerate(channels):
            self.cfg.channels[i] = int(c)
        self.cfg.max_batch = int(max_batch)
        self.cfg.variant = 1 if bootstrap else 0
        self.cfg.conv_backend = int(conv_backend)
        self.cfg.cond_dim = int(cond_dim)          # 0 = unconditional; SoftIntroVAE(conditional=True, cond_dim) otherwise
        self.bootstrap = bootstrap
        self.handle = C.c_void_p()
        L.check(lib.sivae_create(C.byref(self.cfg), C.byref(self.handle)), "sivae_create")
        with torch.cuda.device(self.device):
            self.mem = {L.NET_ENCODER: NetMemory(self.handle, L.NET_ENCODER, self.device),
                        L.NET_DECODER: NetMemory(self.handle, L.NET_DECODER, self.device)}
            if bootstrap:
                self.mem[L.NET_TARGET] = NetMemory(self.handle, L.NET_TARGET, self.device, trainable=False)
            self.stats = torch.zeros(16, dtype=torch.float32, device=self.device)
            self._bind()
        self._graphs = {}          # key -> (CUDAGraph, static input tensors); see graphed()
        self._graph_seen = set()
        self._graph_failed = set() # keys whose capture failed once: they stay eager
        self._comm_world = 1

    def _bind(self):
        lib = L.load()
        for net, m in self.mem.items():
            L.check(lib.sivae_bind_net(self.handle, net, L.ptr(m.params), L.ptr(m.grads), L.ptr(m.m), L.ptr(m.v),
                                       L.ptr(m.bn), L.ptr(m.nbt)), "sivae_bind_net")
        nbytes = lib.sivae_workspace_bytes(self.handle)
        self.workspace = torch.empty(nbytes + 512, dtype=torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        self._ws_ptr = (base + 255) // 256 * 256
        L.check(lib.sivae_bind_workspace(self.handle, C.c_void_p(self._ws_ptr), nbytes), "sivae_bind_workspace")
        self.workspace_bytes = nbytes

    @property
    def max_batch(self):
        return self.cfg.max_batch

    @property
    def reuse_decoder_passes(self):
        """D half takes fake = D(noise) and rec = D(z) (reference :597-598) from the E half's identical passes (:557,
        :561) instead of recomputing them; the BatchNorm running-statistics updates of the skipped passes are replayed
        from the saved batch statistics, so every result stays bit-identical.  Off unless SIVAE_REUSE_DEC=1 or set here."""
        return bool(L.load().sivae_get_reuse_decoder_passes(self.handle))

    @reuse_decoder_passes.setter
    def reuse_decoder_passes(self, on):
        L.check(L.load().sivae_set_reuse_decoder_passes(self.handle, 1 if on else 0), "sivae_set_reuse_decoder_passes")

    @property
    def recon_loss(self):
        """recon_loss_type of the step ('mse' | 'l1' | 'bce'; reference calc_reconstruction_loss :268-294)"""
        code = L.load().sivae_get_recon_loss(self.handle)
        return {v: k for k, v in L.LOSS_TYPES.items()}[code]

    @recon_loss.setter
    def recon_loss(self, loss_type):
        if loss_type not in L.LOSS_TYPES:
            raise NotImplementedError("recon_loss_type must be one of %s" % (sorted(L.LOSS_TYPES),))     # :292-293
        if loss_type != self.recon_loss:
            self.drop_graphs()                    # the loss kernels are baked into the captured step
            L.check(L.load().sivae_set_recon_loss(self.handle, L.LOSS_TYPES[loss_type]), "sivae_set_recon_loss")

    def check_loss_domain(self, stats_host):
        """stats[14]: recon_loss_type='bce' saw a reconstruction outside [0, 1] -- F.binary_cross_entropy raises there"""
        if bool(stats_host[14] != 0):
            raise RuntimeError("all elements of input should be between 0 and 1")

    def drop_graphs(self):
        """Forget every captured step graph (pointers or derived-filter state they baked in are no longer valid)."""
        self._graphs = {}
        self._graph_seen = set()
        self._graph_failed = set()

    def graphed(self, key, inputs, fn):
        """Run fn(*static_inputs) -- a sequence of engine calls on the current stream -- from a CUDA graph.
        `inputs` are copied into per-key static tensors first.  The first call with a given key runs eagerly (it also
        performs the one-time cudaFuncSetAttribute calls), the second captures, later ones replay.  Everything that
        changes between iterations lives on the device (parameters, Adam moments and step counters, BN statistics),
        so a replay is one more training step.  Scalars baked into the graph (learning rates, betas) are part of `key`."""
        if key not in self._graph_seen:
            self._graph_seen.add(key)
            return fn(*inputs)
        if key in self._graph_failed:
            return fn(*inputs)
        ent = self._graphs.get(key)
        if ent is None:
            statics = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(statics, inputs):
                s_.copy_(t, non_blocking=True)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            try:
                # thread_local: a cudaHostAlloc / cudaMalloc from ANOTHER thread (the DataLoader's pin-memory thread is still
                # warming up when the second iteration captures) must not invalidate this thread's capture
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    fn(*statics)
            except RuntimeError as ex:
                # a failed capture leaves no partial work behind (nothing executes during capture): run this step eagerly
                # and keep the key eager from now on
                import warnings
                warnings.warn("CUDA graph capture of %r failed (%s); this step shape stays eager" % (key[0], str(ex)[:200]))
                self._graph_failed.add(key)
                torch.cuda.synchronize(self.device)
                return fn(*inputs)
            ent = (g, statics)
            self._graphs[key] = ent
        g, statics = ent
        for s_, t in zip(statics, inputs):
            s_.copy_(t, non_blocking=True)
        g.replay()
        return None

    def grow(self, max_batch):
        """Re-create the native handle with a larger workspace; parameter / optimiser memory is kept."""
        self.drop_graphs()
        lib = L.load()
        steps = {net: lib.sivae_adam_get_step(self.handle, net) for net in self.mem}
        reuse = self.reuse_decoder_passes
        loss = self.recon_loss
        lib.sivae_destroy(self.handle)
        self.cfg.max_batch = int(max_batch)
        self.handle = C.c_void_p()
        L.check(lib.sivae_create(C.byref(self.cfg), C.byref(self.handle)), "sivae_create")
        self.workspace = None
        with torch.cuda.device(self.device):
            self._bind()
        for net, s in steps.items():
            lib.sivae_adam_set_step(self.handle, net, s)
        self.reuse_decoder_passes = reuse
        self.recon_loss = loss
        self._comm_world = 1        # the new handle is not attached to the process-global communicator yet (re-attached on next use)

    # ---- data parallel: the library-owned NCCL communicator ------------------------------------------------------
    @property
    def comm_world(self):
        return self._comm_world

    def comm_init(self, dist):
        """Create the engine's own NCCL communicator over the ranks of the initialised torch.distributed default group:
        rank 0 draws the ncclUniqueId, torch.distributed ships its 128 bytes (plumbing), every rank calls ncclCommInitRank
        inside the library.  From then on the gradient all-reduces are raw ncclAllReduce calls on the step's stream."""
        lib = L.load()
        world, rank = dist.get_world_size(), dist.get_rank()
        if lib.sivae_comm_global_world() == world:
            # the communicator is process-global: a second engine of this process attaches to the existing one
            L.check(lib.sivae_comm_init(self.handle, None, world, rank), "sivae_comm_init (attach)")
        else:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (C.c_ubyte * 128)()
                L.check(lib.sivae_comm_unique_id(buf), "sivae_comm_unique_id")
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            on_dev = dist.get_backend() == "nccl"
            t = uid.to(self.device) if on_dev else uid
            dist.broadcast(t, src=0)
            raw = bytes(t.cpu().tolist())
            with torch.cuda.device(self.device):
                L.check(lib.sivae_comm_init(self.handle, (C.c_ubyte * 128).from_buffer_copy(raw), world, rank), "sivae_comm_init")
        self._comm_world = world
        self.drop_graphs()

    @staticmethod
    def comm_finalize():
        """destroy the process-global communicator (all ranks, after the last step, before dist.destroy_process_group)"""
        L.check(L.load().sivae_comm_finalize(), "sivae_comm_finalize")

    def allreduce_grads(self, net):
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_allreduce_grads(self.handle, net, _stream()), "sivae_allreduce_grads")

    def iteration(self, real, noise, eps5, hp, lr_e, lr_d):
        """E half, all-reduce, Adam(encoder), D half, all-reduce, Adam(decoder) as one library call (sivae_iteration)"""
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_iteration(self.handle, L.ptr(real), L.ptr(noise), L.ptr(eps5), real.shape[0], C.byref(hp),
                                             float(lr_e), float(lr_d), L.ptr(self.stats), _stream()), "sivae_iteration")

    def broadcast_state(self, dist, src=0):
        """make every rank's replica identical to rank `src`: parameters, Adam moments and step counters, BatchNorm running
        statistics and num_batches_tracked (the reference's DP precedent, DistributedDataParallel, does this at construction)"""
        lib = L.load()
        for net, m in self.mem.items():
            for t in (m.params, m.m, m.v, m.bn, m.nbt):
                if t is not None:
                    dist.broadcast(t, src=src)
            if m.m is not None:
                st = torch.tensor([lib.sivae_adam_get_step(self.handle, net)], dtype=torch.int64, device=self.device)
                dist.broadcast(st, src=src)
                lib.sivae_adam_set_step(self.handle, net, int(st.item()))
        self.params_changed()

    def close(self):
        if self.handle:
            L.load().sivae_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- steps --------------------------------------------------------------------------------------------
    def params_changed(self, net=None):
        self.drop_graphs()
        lib = L.load()
        for n in (self.mem if net is None else [net]):
            L.check(lib.sivae_params_changed(self.handle, n), "sivae_params_changed")

    def e_step(self, real, noise, eps3, hp):
        B = real.shape[0]
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_e_step(self.handle, L.ptr(real), L.ptr(noise), L.ptr(eps3), B, C.byref(hp),
                                          L.ptr(self.stats), _stream()), "sivae_e_step")

    def d_step(self, eps2, hp):
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_d_step(self.handle, L.ptr(eps2), C.byref(hp), L.ptr(self.stats), _stream()), "sivae_d_step")

    def vae_step(self, real, eps, hp):
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_vae_step(self.handle, L.ptr(real), L.ptr(eps), real.shape[0], C.byref(hp),
                                            L.ptr(self.stats), _stream()), "sivae_vae_step")

    def adam(self, net, lr, grad_scale=1.0):
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_adam_step(self.handle, net, float(lr), float(grad_scale), _stream()), "sivae_adam_step")

    def _cond(self, cond, B):
        """[B, cond_dim] float32 rows on the engine's device (the o_cond / y_cond argument of the reference's forward calls)"""
        cond = cond.reshape(cond.size(0), -1).to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(cond.shape) != (B, self.cfg.cond_dim):
            raise RuntimeError("condition of shape %s does not match [batch %d, cond_dim %d]" % (tuple(cond.shape), B, self.cfg.cond_dim))
        return cond

    def encode(self, x, train, cond=None):
        B = x.shape[0]
        mu = torch.empty(B, self.cfg.zdim, dtype=torch.float32, device=self.device)
        lv = torch.empty_like(mu)
        with torch.cuda.device(self.device):
            if cond is not None:
                cond = self._cond(cond, B)
                L.check(L.load().sivae_encode_cond(self.handle, L.ptr(x), L.ptr(cond), B, L.ptr(mu), L.ptr(lv), 1 if train else 0,
                                                   _stream()), "sivae_encode_cond")
            else:
                L.check(L.load().sivae_encode(self.handle, L.ptr(x), B, L.ptr(mu), L.ptr(lv), 1 if train else 0, _stream()), "sivae_encode")
        return mu, lv

    def decode(self, z, train, net=L.NET_DECODER, cond=None):
        B = z.shape[0]
        out = torch.empty(B, self.cfg.cdim, self.cfg.image_size, self.cfg.image_size, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if cond is not None:
                cond = self._cond(cond, B)
                L.check(L.load().sivae_decode_cond(self.handle, net, L.ptr(z), L.ptr(cond), B, L.ptr(out), 1 if train else 0,
                                                   _stream()), "sivae_decode_cond")
            else:
                L.check(L.load().sivae_decode(self.handle, net, L.ptr(z), B, L.ptr(out), 1 if train else 0, _stream()), "sivae_decode")
        return out

    def note_batch(self, B):
        """graph replays do not run the library's host code: tell it the batch size of the step just replayed (last_image)"""
        L.load().sivae_set_last_batch(self.handle, int(B))

    def last_image(self, slot):
        """decoder output of the last half step as NCHW: 0 = fake, 1 = rec, 2 = rec_rec, 3 = rec_fake"""
        B = L.load().sivae_last_batch(self.handle)
        out = torch.empty(B, self.cfg.cdim, self.cfg.image_size, self.cfg.image_size, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().sivae_last_image(self.handle, int(slot), L.ptr(out), _stream()), "sivae_last_image")
        return out
