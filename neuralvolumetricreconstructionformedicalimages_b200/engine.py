"""NAFEngine -- the opt-in fused training / inference driver ("an opt-in
fused train_step entry that bypasses autograd").

One optimisation step is four kernel launches, captured once in a CUDA graph (nothing changes between replays: step
count, learning rate and the sampler's RNG seed live in a device-resident state the kernels read and advance):

    density_forward (detector pixels or rays -> ray generation -> sampling -> gather -> MLP -> sum sigma*delta -> acc[N];
                     leaves the encodings in the stash)
    mse_loss        (masked chunk-wise MSE + d loss / d acc; clears acc)
    density_backward(stash -> MLP forward/backward on tensor cores -> aggregated scatter into the flat gradient -> grid barrier ->
                     reduction of the per-CTA MLP gradients)
    adam_step_dev   (fused dense Adam over ONE flat vector [table | MLP], zeroing the gradient) -- or, on several GPUs, the
                    fused exchange kernel (reduce-scatter + Adam + all-gather over NVLink peer memory)

It takes the place of ``Trainer.train_step`` (reference src/trainer.py:134-142) +
``BasicTrainer.compute_loss`` (train.py:48-135) for the shipped configurations.  Parameters stay
ordinary ``nn.Parameter`` views of the flat vector, so ``state_dict()`` / ``ckpt.tar`` round-trip
with the reference layout (encoder.embeddings, layers.i.weight/bias).

Multi-GPU: one process per GPU, rays sharded across ranks, parameters replicated.  The step's one exchange
(sum of the flat gradient over ranks -> Adam -> identical parameters everywhere) runs in one of these ways:

  exchange="push" (default when it can be set up): ONE kernel over NVLink peer memory (csrc/exchange.cu): every rank owns a
      slice of the flat vector; ranks push their gradient slices into the owners' staging areas, the owner adds them in rank
      order, applies Adam (optimizer state for the slice only) and pushes the new parameters into all replicas -- stores
      only over NVLink.  Buffers come from nafb_peer_alloc and are mapped into the peers with CUDA IPC.
  exchange="peer": the pull edition (owners LOAD the peers' gradient slices; gradients double buffered by step parity);
  exchange="nvls": as "peer" on torch symmetric memory, the sum done by the NVSwitch (multimem.ld_reduce / multimem.st).
  exchange="nccl": dist.all_reduce of the flat gradient, then the dense Adam kernel with grad_scale = 1/world.

The voxel query shards by slabs of the outermost index and needs no collective.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, parallel
from .fused import density_backward, density_forward, stash_bytes
from .network.network import DensityNetwork


def _round_up(n, m):
    return (n + m - 1) // m * m


class _NoTimer:
    def __call__(self, name):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class EventTimer:
    """CUDA-event stopwatch around individual launches on the current stream (profiling runs only)."""

    def __init__(self):
        self.spans = []
        self._name = None

    def __call__(self, name):
        self._name = name
        return self

    def __enter__(self):
        self._e0 = torch.cuda.Event(enable_timing=True)
        self._e0.record()
        return self

    def __exit__(self, *a):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.spans.append((self._name, self._e0, e1))
        return False

    def summary(self):
        """name -> (mean ms, count); call after torch.cuda.synchronize()."""
        acc = {}
        for name, e0, e1 in self.spans:
            t, c = acc.get(name, (0.0, 0))
            acc[name] = (t + e0.elapsed_time(e1), c + 1)
        return {k: (t / c, c) for k, (t, c) in acc.items()}


class PeerSetupError(RuntimeError):
    """Peer-memory exchange could not be set up on every rank (raised on ALL ranks together)."""


class PeerExchange:
    """Peer-mapped flat buffers of all ranks + the descriptors of nafb_adam_exchange_step (one per gradient parity).

    Two ways to get memory every rank can address:
      "nvls": torch symmetric memory (CUDA VMM + multicast object): peers' pointers AND one multicast address, so the kernel
              can let the NVSwitch add the gradients (multimem.ld_reduce) and multicast the parameters (multimem.st);
      "ipc" : cudaMalloc + CUDA IPC handles (nafb_peer_alloc / nafb_peer_open): plain P2P loads and stores (pull edition);
      "push": same memory plus a staging area per rank: gradients are PUSHED to their owners and parameters pushed back --
              stores only over NVLink, one gradient buffer per rank (the default: fastest measured).
    """

    FLAG_BYTES = 256

    def __init__(self, n, device, group, rank, world, backend="ipc"):
        """Set-up is COLLECTIVE and staged so that a failure on one rank never leaves the others inside a collective that
        rank skipped: every stage that can fail locally is followed by an all-reduce(MIN) of its outcome; handles are exchanged
        and peers are opened only when every rank got that far.  On any failure every rank releases what it holds and raises
        PeerSetupError (the engine then falls back to NCCL on all ranks together)."""
        self.n, self.rank, self.world, self.device, self.backend = int(n), rank, world, device, backend
        self.bufs, self.local = [], None
        fb, seg = self.FLAG_BYTES, self.n * 4
        n_grad = 1 if backend == "push" else 2
        slot = ((self.n // 4 + world - 1) // world) * 4 if backend == "push" else 0       # floats per staging slot
        stage_off = fb + (1 + n_grad) * seg
        nbytes = stage_off + world * slot * 4
        mc_base = 0

        def agree(ok, what, err=None):
            """all ranks succeeded?  (one all-reduce; every rank calls it at the same point whatever happened locally)"""
            if world > 1:
                t = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
                ok_all = bool(int(t.item()))
            else:
                ok_all = ok
            if not ok_all:
                self.close()
                raise PeerSetupError(f"{backend}: {what} failed on {'this rank: ' + str(err) if not ok else 'a peer'}")

        if backend == "nvls":
            err = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._symm = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)
                self._symm.zero_()
            except Exception as e:   # noqa: BLE001
                err = e
            agree(err is None, "symmetric allocation", err)
            try:
                self._hdl = symm_mem.rendezvous(self._symm, group if group is not None else dist.group.WORLD)
                mc_base = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
                if mc_base == 0:
                    raise RuntimeError("symmetric memory has no multicast (NVLS) address on this system")
                ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            except Exception as e:   # noqa: BLE001
                err = e
            agree(err is None, "rendezvous / multicast", err)
            flat = self._symm
            self.flags = flat[: fb // 4].view(torch.int32)[: _lib.XFLAG_WORDS]
            self.param = flat[fb // 4 : fb // 4 + self.n]
            self.grad = [flat[fb // 4 + (1 + k) * self.n : fb // 4 + (2 + k) * self.n] for k in (0, 1)]
            local_ptr = ptrs[rank]
        else:
            err = None
            try:                                           # stage 1: local allocation
                self.local = _lib.peer_alloc(nbytes)
            except Exception as e:   # noqa: BLE001
                err = e
            agree(err is None, "peer_alloc", err)
            handles = [None] * world                       # stage 2: handle exchange (every rank is here)
            if world > 1:
                dist.all_gather_object(handles, self.local.handle, group=group)
            try:                                           # stage 3: map the peers
                for w in range(world):
                    if w == rank:
                        self.bufs.append(self.local)
                    else:
                        if os.environ.get("NAFB_TEST_FAIL_PEER_OPEN") == str(rank):
                            raise RuntimeError("injected peer_open failure (NAFB_TEST_FAIL_PEER_OPEN)")
                        self.bufs.append(_lib.peer_open(handles[w], nbytes))
            except Exception as e:   # noqa: BLE001
                err = e
            agree(err is None, "peer_open", err)
            ptrs = [b.ptr for b in self.bufs]
            self.flags = self.local.tensor(0, _lib.XFLAG_WORDS, torch.int32, device)
            self.param = self.local.tensor(fb, self.n, torch.float32, device)
            self.grad = [self.local.tensor(fb + (1 + k) * seg, self.n, torch.float32, device) for k in range(n_grad)]
            local_ptr = self.local.ptr
        i0, i1 = _lib.exchange_slice(self.n, rank, world)
        self.slice = (i0, i1)
        self.exp_avg = torch.zeros(i1 - i0, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(i1 - i0, device=device, dtype=torch.float32)
        self.desc = []
        for par in range(n_grad):
            x = _lib.Exchange()
            x.world, x.rank, x.n = world, rank, self.n
            for w, base in enumerate(ptrs):
                x.flags[w] = base
                x.param[w] = base + fb
                x.grad[w] = base + fb + (1 + par) * seg
                if backend == "push":
                    x.stage[w] = base + stage_off
            if backend == "push":
                x.stage_slot = slot
            else:
                x.grad_zero = local_ptr + fb + (1 + (1 - par)) * seg
            x.exp_avg, x.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
            if mc_base:
                x.mc_param = mc_base + fb
                x.mc_grad = mc_base + fb + (1 + par) * seg
            self.desc.append(x)
        if world > 1:
            torch.cuda.synchronize(device)
            dist.barrier(group=group)   # every rank has mapped every buffer before anybody launches

    def reset_flags(self, group=None):
        """Zero the epoch flags of every rank (collective).  Needed whenever the optimizer step -- the synchronisation epoch --
        moves BACKWARDS (a checkpoint rollback): stale flags of later epochs would satisfy the kernel's waits at once."""
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)
        self.flags.zero_()
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)

    def step(self, par, lr, betas, eps, step, stream):
        # grad_scale 1: the step's loss is the SUM of the ranks' chunked losses -- the reference's own rule (train.py:69-127 adds the
        # chunk means), so W ranks equal one GPU working on the concatenated batch with the same chunk boundaries
        _lib.check(_lib.lib().nafb_adam_exchange_step(ctypes.byref(self.desc[par]), lr, betas[0], betas[1], eps, step, 1.0, stream))

    def error_word(self) -> int:
        """0, or 1 + index of the peer flag a bounded spin gave up on (synchronises the stream)."""
        return int(self.flags[_lib.XFLAG_ERROR].item())

    def close(self):
        """Release every mapping and the local allocation (idempotent)."""
        for b in self.bufs:
            if b is not self.local:
                b.release()
        self.bufs = []
        if self.local is not None:
            self.local.release()
            self.local = None


def _stage(dst: torch.Tensor, src: torch.Tensor, dtype):
    """src -> dst (a view of a pinned staging buffer).  dtype None: any 1-byte type (uint8 / bool masks)."""
    if (src.device.type == "cpu" and src.is_contiguous() and src.numel() == dst.numel()
            and (src.dtype == dtype or (dtype is None and src.element_size() == 1))):
        ctypes.memmove(dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size())
    else:
        dst.copy_(src.reshape(dst.shape))


class PendingLoss:
    """The loss of a step enqueued by NAFEngine.train_step_host: `result()` returns the float as soon as the step's FORWARD launch
    has written it into pinned host memory (it polls the completion word the kernel stores after the loss; the backward pass and
    the optimizer of the step keep running, stream-ordered before anything enqueued later)."""
    __slots__ = ("_event", "_buf", "_flag", "_expect", "_value")

    def __init__(self, event, buf, flag=None, expect=0):
        self._event, self._buf, self._flag, self._expect, self._value = event, buf, flag, int(expect) & 0xFFFFFFFF, None

    def done(self) -> bool:
        if self._value is not None:
            return True
        if self._flag is not None:
            return int(self._flag[0]) == self._expect
        return self._event.query()

    def result(self) -> float:
        if self._value is None:
            if self._flag is not None:
                flag, expect = self._flag, self._expect
                spins = 0
                while int(flag[0]) != expect:
                    spins += 1
                    if spins > 2000000:           # seconds: something is wrong (a launch failed?) -- fall back to the event, which raises
                        self._event.synchronize()
                        if int(flag[0]) != expect:
                            raise RuntimeError("train_step_host: the step finished without writing its loss")
            else:
                self._event.synchronize()
            self._value = float(self._buf[0])
        return self._value


class NAFEngine:
    def __init__(self, net: DensityNetwork, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, n_samples=192, perturb=True, loss_chunk=None,
                 use_cuda_graph=True, process_group=None, use_stash=True, exchange="auto", seed=None):
        meta = net.fused_meta()
        if meta is None:
            raise RuntimeError("NAFEngine needs a DensityNetwork in a fused-capable configuration "
                               "(hash grid with L*C == 32, hidden_dim == 32, out_dim == 1)")
        dev = net.encoder.embeddings.device
        if dev.type != "cuda":
            raise RuntimeError("NAFEngine needs the network on a CUDA device (there is no CPU path)")
        _lib.lib()
        self.net, self.meta, self.device = net, meta, dev
        self.state = None
        self._lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.n_samples, self.perturb, self.loss_chunk = int(n_samples), bool(perturb), loss_chunk
        self.step_count = 0
        self.use_cuda_graph = use_cuda_graph
        self.use_stash = bool(use_stash)  # forward leaves the encodings (128 B/point) for backward instead of a second gather
        self.pg = process_group
        self.rank, self.world_size = parallel.world_info(process_group)
        if exchange not in ("auto", "push", "nvls", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'push', 'peer', 'nvls' or 'nccl'")
        self.px = None
        self._flatten(exchange)
        if self.world_size > 1:   # replicas start identical
            parallel.broadcast_(self.flat_param, 0, process_group)
        self._init_state(seed)
        self._graphs = {}
        self._seen_batches = {}
        self._prefetch = None        # (sampler, n_rays, buffers, sampler version) of the batch train_step_sampled has drawn ahead
        self._side_stream = None
        self._eager_runs = {}
        self._host_seq = 0
        self._static = {}
        # backward workspace of THIS engine (per-CTA partial MLP gradients + the words of the kernel's grid barrier): never shared
        # with another engine or stream, zero-filled once
        from .fused import new_workspace
        self._bwd_ws = new_workspace(self.meta.mlp(self.mlp_params), self.device)

    # ------------------------------------------------------------------ flat parameter vector
    def _flatten(self, exchange="auto"):
        params = [self.net.encoder.embeddings] + self.net.flat_params()
        # Lead-in of 2 floats: the table starts 8 bytes past a 16-byte boundary.  Every hashed level of the shipped grids
        # starts at an ODD entry offset (4913 + 35937 + 274625 + k * 2^19), so its entry pairs (2k, 2k+1) -- the x-neighbour
        # pairs of the xor hash -- land on 16-byte boundaries and the kernels serve them with one 128-bit access
        # (common.cuh load_entry_pair / red_add_entry_pair; they test the alignment at run time, so any base is correct).
        offs, n = [], 2
        for p in params:
            offs.append(n)
            n += _round_up(p.numel(), 4)
        n = _round_up(n, 4)
        # exchange over peer memory: NVLS (switch-side reduction + multicast) when the system offers it, else plain P2P over
        # CUDA IPC, else NCCL.  Every rank must take the same decision, hence the all-reduce of the outcome.
        candidates = {"auto": ["push"] if self.world_size > 1 else [], "push": ["push"], "nvls": ["nvls"], "peer": ["ipc"], "nccl": []}[exchange]
        for backend in candidates:
            try:
                if self.world_size > _lib.NAFB_MAX_RANKS:
                    raise PeerSetupError(f"peer exchange supports up to {_lib.NAFB_MAX_RANKS} ranks")      # same decision on every rank
                if backend == "nvls" and self.world_size == 1:
                    raise PeerSetupError("NVLS needs more than one rank")
                with torch.cuda.device(self.device):
                    self.px = PeerExchange(n, self.device, self.pg, self.rank, self.world_size, backend)
            except PeerSetupError as e:   # raised on every rank together (PeerExchange agrees on each stage's outcome)
                if self.world_size == 1:
                    raise
                self._peer_error, self.px = str(e), None
            if self.px is not None:
                break
        if self.px is None and exchange in ("push", "nvls", "peer"):
            raise RuntimeError(f"exchange='{exchange}' could not be set up on every rank: " + getattr(self, "_peer_error", "failed on a peer"))
        if self.px is not None:
            self.flat_param = self.px.param
            self.flat_grads = self.px.grad                     # double buffered by step parity
            self.exp_avg, self.exp_avg_sq = self.px.exp_avg, self.px.exp_avg_sq   # slice-local
        else:
            self.flat_param = torch.zeros(n, device=self.device, dtype=torch.float32)
            self.flat_grads = [torch.zeros_like(self.flat_param)]
            self.exp_avg = torch.zeros_like(self.flat_param)
            self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.flat_grad = self.flat_grads[0]
        self._views = []
        self._grad_views = [[] for _ in self.flat_grads]
        for p, o in zip(params, offs):
            v = self.flat_param[o:o + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v  # the module's parameters now alias the flat vector
            self._views.append(v)
            for k, fg in enumerate(self.flat_grads):
                self._grad_views[k].append(fg[o:o + p.numel()].view_as(p))
        self.n_params = n
        self.table = self._views[0]
        self.mlp_params = self._views[1:]
        self.grad_table = self._grad_views[0][0]
        self.grad_mlp = self._grad_views[0][1:]

    @property
    def exchange_mode(self):
        if self.px is not None:
            return {"push": "push", "nvls": "nvls", "ipc": "peer"}[self.px.backend]
        return "nccl" if self.world_size > 1 else "local"

    def _parity(self):
        """Which gradient buffer the coming step accumulates into (peer exchange: alternates; otherwise always 0)."""
        return (self.step_count & 1) if (self.px is not None and len(self.flat_grads) == 2) else 0

    # ------------------------------------------------------------------ device-resident step state
    def _init_state(self, seed):
        """nafb_step_state: [step, seed_lo, seed_hi, lr bits, ticket, ...] -- what changes from step to step lives on the
        device, so one captured graph replays the whole iteration (sampler uniforms and Adam's bias corrections included)."""
        seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        seed ^= (0x9E3779B97F4A7C15 * (self.rank + 1)) & 0xFFFFFFFFFFFFFFFF      # every rank draws its own jitter
        st = np.zeros(_lib.STATE_WORDS, dtype=np.uint32)
        st[_lib.STATE_SEED_LO], st[_lib.STATE_SEED_HI] = seed & 0xFFFFFFFF, seed >> 32
        st[_lib.STATE_LR : _lib.STATE_LR + 2] = np.array([self._lr], dtype=np.float64).view(np.uint32)
        self.state = torch.from_numpy(st.view(np.int32)).to(self.device)
        if self.px is not None:
            for x in self.px.desc:
                x.state = self.state.data_ptr()

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        """Learning rate (follow a scheduler by assignment); mirrored into the device state the kernels read."""
        self._lr = float(value)
        if getattr(self, "state", None) is not None:
            self.state[_lib.STATE_LR : _lib.STATE_LR + 2] = torch.tensor(_lib.lr_words(self._lr), dtype=torch.int32)

    def _set_step(self, step):
        self.step_count = int(step)
        self.state[_lib.STATE_STEP : _lib.STATE_STEP + 1] = torch.tensor([self.step_count], dtype=torch.int32)

    # ------------------------------------------------------------------ one training step
    def set_geometry(self, angles, geo):
        """Scanner geometry for the pixel source: the kernels generate the rays of (projection, row, col) themselves
        (reference src/dataset/tigre.py:402-528) instead of reading a [N,8] rays tensor."""
        from .dataset import geometry as G
        self.poses = G.pose_table(angles, geo, self.device)
        self.det = G.detector_fields(geo)

    def _ray_sampler(self, rays, pixels, t_rand):
        """nafb_sampler of a ray batch: explicit rays [N,8], or detector pixels [N,3] (needs set_geometry)."""
        kw = dict(t_rand=t_rand.data_ptr() if t_rand is not None else None,
                  rng_state=self.state.data_ptr() if (self.perturb and t_rand is None) else None,
                  n_samples=self.n_samples, perturb=int(self.perturb))
        if pixels is not None:
            if getattr(self, "poses", None) is None:
                raise RuntimeError("pixel batches need NAFEngine.set_geometry(angles, geo) first")
            return self.meta.sampler(pixels=pixels.data_ptr(), poses=self.poses.data_ptr(), n_rays=pixels.shape[0], **self.det, **kw)
        return self.meta.sampler(rays=rays.data_ptr(), n_rays=rays.shape[0], **kw)

    def _step_kernels(self, rays, projs, mask, t_rand, loss_out, dacc, acc, timer=None, stash=None, par=0, pixels=None, done_flag=None):
        """density_fwd -> mse_loss -> density_bwd (+ reduce).  t_rand None: the sampler draws its uniforms in-kernel."""
        L_ = _lib.lib()
        tm = timer or _NoTimer()
        N = pixels.shape[0] if pixels is not None else rays.shape[0]
        # acc is zero on entry: allocated zeroed, and mse_loss clears it after reading (zero_pred)
        grid = self.meta.grid(self.table)
        mlp = self.meta.mlp(self.mlp_params)
        explicit = self.perturb and t_rand is not None
        smp = self._ray_sampler(rays, pixels, t_rand if explicit else None)
        st = _lib.stream_ptr()
        chunk = int(self.loss_chunk or 0)
        if stash is not None:
            # tensor-core configuration: ONE launch = ray generation + sampling + gather + MLP + ray integral, and its last CTA
            # evaluates the masked chunk-wise MSE and d loss / d acc (nafb_density_forward_loss)
            tail = _lib.LossTail(target=projs.data_ptr(), mask=mask.data_ptr() if mask is not None else None, chunk=chunk, gscale=1.0,
                                 loss_out=loss_out.data_ptr(), dacc=dacc.data_ptr(), zero_pred=1,
                                 ticket=self.state.data_ptr() + 4 * _lib.STATE_TICKET_FWD,
                                 done_flag=done_flag, step_state=self.state.data_ptr() if done_flag else None)
            with tm("density_fwd_loss"):
                _lib.check(L_.nafb_density_forward_loss(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), _lib.ptr(acc), None, _lib.ptr(stash),
                                                        ctypes.byref(tail), st))
        else:
            with tm("density_fwd"):
                _lib.check(L_.nafb_density_forward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), _lib.SRC_RAYS, None, _lib.ptr(acc),
                                                   None, None, None, None, st))
            with tm("mse_loss"):
                _lib.check(L_.nafb_mse_loss(_lib.ptr(acc), _lib.ptr(projs), _lib.ptr(mask), N, chunk, 1.0, _lib.ptr(loss_out), _lib.ptr(dacc), 1, st))
        with tm("density_bwd"):
            gv = self._grad_views[par]
            density_backward(self.meta, self.table, self.mlp_params, dacc, gv[0], gv[1:], sampler=smp, n_points=N * self.n_samples, stash=stash,
                             workspace=self._bwd_ws)

    def _optimizer_kernel(self, par, timer=None):
        """Peer mode: the fused exchange kernel.  Otherwise the dense Adam kernel.  Both take step / lr from the device state
        and increment the step when their last block retires (graph-capturable)."""
        L_ = _lib.lib()
        tm = timer or _NoTimer()
        if self.px is not None:
            with tm("adam_exchange"):
                self.px.step(par, self._lr, self.betas, self.eps, 0, _lib.stream_ptr())
            return
        with tm("adam"):
            _lib.check(L_.nafb_adam_step_dev(_lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                             _lib.ptr(self.exp_avg_sq), self.n_params, self.betas[0], self.betas[1], self.eps,
                                             1.0, 1, _lib.ptr(self.state), _lib.stream_ptr()))

    def _whole_step(self, s, par, timer=None, t_rand=None, with_optimizer=True, use_pixels=False, loss_out=None, done_flag=None):
        self._step_kernels(None if use_pixels else s["rays"], s["projs"], s["mask"], t_rand, s["loss"] if loss_out is None else loss_out,
                           s["dacc"], s["acc"], timer,
                           stash=s["stash"], par=par, pixels=s["pixels"] if use_pixels else None, done_flag=done_flag)
        if with_optimizer:
            self._finish_step(par, timer)

    def _finish_step(self, par, timer=None):
        if self.world_size > 1 and self.px is None:
            with (timer or _NoTimer())("all_reduce"):
                parallel.allreduce_sum_(self.flat_grad, self.pg)
        self._optimizer_kernel(par, timer)

    @property
    def launches_per_step(self):
        """Kernels of this library launched by one train_step: density_fwd_loss, density_bwd, adam / adam_exchange in the
        tensor-core configuration; the fp32 SIMT arithmetic runs density_fwd, mse_loss, density_bwd, reduce_partials, adam."""
        return 3 if self.use_stash and stash_bytes(self.meta, self.table, self.mlp_params, 128) else 5

    def _load_inputs(self, s, rays, projs, mask, t_rand, pixels=None):
        """Inputs -> the static buffers the graph reads (device tensors, or pinned host tensors: one H2D copy each)."""
        N = s["rays"].shape[0]
        if pixels is not None:
            s["pixels"].copy_(pixels.reshape(N, 3), non_blocking=True)
        else:
            s["rays"].copy_(rays.reshape(N, 8), non_blocking=True)
        s["projs"].copy_(projs.reshape(N), non_blocking=True)
        if mask is not None:
            s["mask"].copy_(mask.reshape(N), non_blocking=True)
        if self.perturb and t_rand is not None:
            if s["t_rand"] is None:
                s["t_rand"] = torch.zeros(N, self.n_samples, device=self.device)
            s["t_rand"].copy_(t_rand, non_blocking=True)
            return s["t_rand"]
        return None

    def profiled_step(self, rays, projs, mask, t_rand, timer, pixels=None):
        """Same work as train_step, launched eagerly (no graph) with a CUDA-event pair around every kernel."""
        N = pixels.shape[0] if pixels is not None else rays.shape[0]
        s = self._get_static(N, mask is not None)
        with torch.cuda.device(self.device):
            tr = self._load_inputs(s, rays, projs, mask, t_rand, pixels)
            par = self._parity()
            self._whole_step(s, par, timer, tr, use_pixels=pixels is not None)
            self.step_count += 1
        return s["loss"][0]

    def _get_static(self, N, with_mask):
        key = (N, with_mask)
        s = self._static.get(key)
        if s is None:
            d = self.device
            nb = stash_bytes(self.meta, self.table, self.mlp_params, N * self.n_samples) if self.use_stash else 0
            # the per-step inputs live in ONE device buffer [rays | pixels | projs | mask] (16-byte aligned regions): a step fed
            # from the host is one H2D copy of the regions it uses (train_step_host)
            r16 = lambda n: (n + 15) // 16 * 16
            o_rays, o_pix = 0, r16(32 * N)
            o_projs = o_pix + r16(12 * N)
            o_mask = o_projs + r16(4 * N)
            packed = torch.zeros(o_mask + r16(N), device=d, dtype=torch.uint8)
            s = dict(packed=packed, offsets=(o_rays, o_pix, o_projs, o_mask),
                     rays=packed[o_rays:o_rays + 32 * N].view(torch.float32).view(N, 8),
                     pixels=packed[o_pix:o_pix + 12 * N].view(torch.int32).view(N, 3),
                     projs=packed[o_projs:o_projs + 4 * N].view(torch.float32),
                     stash=torch.empty(nb, dtype=torch.uint8, device=d) if nb else None,
                     mask=packed[o_mask:o_mask + N] if with_mask else None, t_rand=None,
                     loss=torch.zeros(2, device=d), dacc=torch.zeros(N, device=d), acc=torch.zeros(N, device=d))
            self._static[key] = s
        return s

    def train_step(self, rays, projs, mask=None, t_rand=None, pixels=None):
        """One optimisation step.  rays [N,8] -- or rays=None and pixels [N,3] int32 (projection, row, col; the kernels then
        generate the rays, see set_geometry) --, projs [N] (+ optional uint8/bool mask [N]) on the device or in pinned host
        memory; t_rand [N,S]: explicit uniforms of render.py:99 (parity runs) -- by default the kernels draw them themselves.
        Returns the loss as a 0-dim device tensor (no host synchronisation)."""
        use_pixels = pixels is not None
        N = pixels.shape[0] if use_pixels else rays.shape[0]
        s = self._get_static(N, mask is not None)
        with torch.cuda.device(self.device):
            # Device-resident batches are read IN PLACE: the captured graph is keyed by the addresses of the caller's tensors
            # (a training loop that cycles through resident batches replays one graph per batch, no staging copies).  Host tensors
            # -- or device tensors that are not in the layout the kernels read -- are copied into the engine's own buffers.
            src = pixels if use_pixels else rays

            def resident(x, dtype, shape):
                return (x is not None and x.is_cuda and x.device == self.device and x.is_contiguous() and x.dtype in dtype
                        and tuple(x.shape) == shape)

            in_place = (self.use_cuda_graph and resident(src, (torch.int32,) if use_pixels else (torch.float32,), (N, 3) if use_pixels else (N, 8))
                        and resident(projs, (torch.float32,), (N,)) and (mask is None or resident(mask, (torch.uint8, torch.bool), (N,)))
                        and (t_rand is None or not self.perturb or resident(t_rand, (torch.float32,), (N, self.n_samples))))
            if in_place:
                # ... but only for batches that come back: the first time an address tuple is seen the batch takes the staging
                # path (three small copies); a capture (~1 ms of host time) pays off from its second visit on
                probe = (src.data_ptr(), projs.data_ptr(), mask.data_ptr() if mask is not None else 0, N)
                seen = self._seen_batches.get(probe, 0)
                if seen < 1:
                    if len(self._seen_batches) > 4 * self.MAX_GRAPHS:
                        self._seen_batches.clear()
                    self._seen_batches[probe] = seen + 1
                    in_place = False
            if in_place:
                sv = dict(s)
                sv["pixels" if use_pixels else "rays"] = src
                sv["projs"] = projs
                sv["mask"] = mask.view(torch.uint8) if mask is not None else None
                tr = t_rand if (self.perturb and t_rand is not None) else None
                addr = (src.data_ptr(), projs.data_ptr(), mask.data_ptr() if mask is not None else 0, tr.data_ptr() if tr is not None else 0)
                s_run = sv
            else:
                tr = self._load_inputs(s, rays, projs, mask, t_rand, pixels)
                addr = ()
                s_run = s
            par = self._parity()
            shape_key = (N, mask is not None, par, tr is not None, use_pixels)
            key = shape_key + addr
            # one graph per (shape, parity[, batch addresses]) holds the whole iteration; only an NCCL all-reduce (+ the Adam after it)
            # stays outside
            in_graph = not (self.world_size > 1 and self.px is None)
            if not self.use_cuda_graph:
                self._whole_step(s_run, par, None, tr, use_pixels=use_pixels)
            else:
                g = self._graphs.get(key)
                if g is None and self._eager_runs.get(shape_key, 0) < 1:
                    # first use of a shape: run eagerly (kernel attributes / lazy module loading must not happen under capture)
                    self._eager_runs[shape_key] = self._eager_runs.get(shape_key, 0) + 1
                    self._whole_step(s_run, par, None, tr, use_pixels=use_pixels)
                else:
                    if g is None:
                        if len(self._graphs) >= self.MAX_GRAPHS:      # bound the cache: drop the oldest entries
                            for k in list(self._graphs)[: self.MAX_GRAPHS // 4]:
                                del self._graphs[k]
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._whole_step(s_run, par, None, tr, with_optimizer=in_graph, use_pixels=use_pixels)
                        self._graphs[key] = g   # capture does not execute: the replay below performs this step
                    g.replay()
                    if not in_graph:
                        self._finish_step(par)
            self.step_count += 1
        return s["loss"][0]

    MAX_GRAPHS = 2048   # captured iterations kept (one per batch shape, gradient parity and resident batch)

    def train_step_sampled(self, sampler, n_rays: int):
        """One optimisation step whose batch is DRAWN ON THE DEVICE (dataset.mask.PixelSampler; reference
        src/dataset/tigre.py:354-382 + train.py:59-60,93-95): the draw kernel writes the pixels, projection values and mask bits of
        the next projection straight into the buffers the fused step reads, and the whole iteration -- draw, forward + loss,
        backward, optimizer -- is ONE CUDA graph with no per-step host input at all (needs set_geometry).  Returns the loss as a
        0-dim device tensor.

        The draw of step k+1 is made by step k, on a second stream beside the optimizer kernel (the draw is one latency-bound
        CTA): the sampler is therefore ONE draw ahead of the steps run (`sampler.draws_done() == steps + 1`); the sequence of
        batches is exactly that of `draw_into` + `train_step(pixels=...)`.  If anything else moves the sampler between two steps
        (set_draw, draw, draw_into) the prefetched batch is discarded and the step draws afresh."""
        if getattr(self, "poses", None) is None:
            raise RuntimeError("train_step_sampled needs NAFEngine.set_geometry(angles, geo) first")
        N, with_mask = int(n_rays), sampler.mask is not None
        s = self._get_static(N, with_mask)
        with torch.cuda.device(self.device):
            par = self._parity()
            in_graph = not (self.world_size > 1 and self.px is None)
            # the step's batch: normally already there -- the previous step drew it while its optimizer ran (below).  Otherwise
            # (first step, another n_rays, or somebody else moved the sampler: set_draw / draw / draw_into) draw it now.
            pf = self._prefetch
            if not (pf is not None and pf[0] is sampler and pf[1] == N and pf[2] is s and pf[3] == sampler.version):
                sampler.draw_into(N, s["pixels"], s["projs"], s["mask"])
            if self._side_stream is None:
                # high priority: when the backward pass retires, the draw's single CTA must get its SM before the optimizer's
                # persistent blocks fill every thread slot (it would otherwise run after them)
                self._side_stream = torch.cuda.Stream(device=self.device, priority=-1)
            side = self._side_stream

            def body(with_optimizer):
                self._whole_step(s, par, None, None, with_optimizer=False, use_pixels=True)
                # fork: the NEXT step's draw (one CTA, ~40 us of latency-bound work that needs nothing but the sampler's device
                # state) runs beside the optimizer; forward and backward have consumed this step's batch
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    sampler.draw_into(N, s["pixels"], s["projs"], s["mask"])
                if with_optimizer:
                    self._finish_step(par)
                torch.cuda.current_stream(self.device).wait_stream(side)

            key = (N, with_mask, par, "sampled", id(sampler))
            g = self._graphs.get(key)
            if not self.use_cuda_graph or self._eager_runs.get(key, 0) < 1:
                self._eager_runs[key] = self._eager_runs.get(key, 0) + 1
                body(True)
            else:
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        body(in_graph)
                    self._graphs[key] = g
                g.replay()
                if not in_graph:
                    self._finish_step(par)
            self._prefetch = (sampler, N, s, sampler.version)     # the objects themselves: an id() could be recycled
            self.step_count += 1
        return s["loss"][0]

    HOST_SLOTS = 3   # pinned staging slots of train_step_host: a step can be enqueued while the previous one is still running

    def train_step_host(self, projs, mask=None, pixels=None, rays=None, wait=True):
        """One optimisation step fed from HOST memory, result read back to the host: returns the loss as a python float.

        The end-to-end form of train_step: the inputs are copied into a pinned staging slot (a few KB of CPU memcpy), and
        ONE graph launch performs the H2D copy and the whole iteration.  The forward launch writes the loss straight into the
        slot's pinned host buffer (zero-copy D2H) followed by a completion word; the call polls that word and returns the loss
        as soon as the forward pass has produced it -- the backward pass and the optimizer of the step are still running then,
        stream-ordered before whatever is enqueued next, so the host stages and launches step k+1 while step k finishes and
        the GPU never drains.  (torch.cuda.synchronize() when the parameters themselves are needed on the host.)

        wait=False returns the PendingLoss instead of polling: `.result()` gives the float.  The staging slots rotate
        (HOST_SLOTS), so several steps can be in flight."""
        use_pixels = pixels is not None
        src = pixels if use_pixels else rays
        N = src.shape[0]
        s = self._get_static(N, mask is not None)
        o_rays, o_pix, o_projs, o_mask = s["offsets"]
        slot = self._host_seq % self.HOST_SLOTS
        self._host_seq += 1
        hk = ("host_pix" if use_pixels else "host_rays", slot)
        if s.get(hk) is None:
            hp = torch.zeros(s["packed"].numel(), dtype=torch.uint8).pin_memory()      # host mirror of the packed input buffer
            dp = torch.zeros_like(s["packed"])                                          # ... and the slot's own device copy of it
            loss = torch.zeros(4, dtype=torch.float32).pin_memory()      # loss, number of valid rays, completion word, pad
            sv = dict(s)                                                  # the step's buffers with this slot's input block
            sv.update(packed=dp, rays=dp[o_rays:o_rays + 32 * N].view(torch.float32).view(N, 8),
                      pixels=dp[o_pix:o_pix + 12 * N].view(torch.int32).view(N, 3), projs=dp[o_projs:o_projs + 4 * N].view(torch.float32),
                      mask=dp[o_mask:o_mask + N] if mask is not None else None)
            s[hk] = dict(packed=hp, dev=sv, inp=(hp[o_pix:o_pix + 12 * N].view(torch.int32).view(N, 3) if use_pixels
                                                 else hp[o_rays:o_rays + 32 * N].view(torch.float32).view(N, 8)),
                         projs=hp[o_projs:o_projs + 4 * N].view(torch.float32), mask=hp[o_mask:o_mask + N],
                         loss=loss, loss_np=loss.numpy(), flag_np=loss.numpy()[2:3].view(np.uint32), event=torch.cuda.Event(),
                         copied=torch.cuda.Event(), pending=None)
        h = s[hk]
        sv = h["dev"]
        if h["pending"] is not None:
            h["pending"].result()      # the launch that last used this slot has read its inputs and written its loss
        # staging: a few KB; plain memmove when the caller's tensors already have the staged layout (a torch copy_ call costs
        # more than the copy itself at this size, and in the waited form every host microsecond is on the critical path)
        _stage(h["inp"], src, torch.int32 if use_pixels else torch.float32)
        _stage(h["projs"], projs, torch.float32)
        if mask is not None:
            _stage(h["mask"], mask, None)
        end = o_mask + N if mask is not None else o_projs + 4 * N
        with torch.cuda.device(self.device):
            par = self._parity()
            in_graph = not (self.world_size > 1 and self.px is None)
            # H2D on a copy stream into the slot's own device block: the transfer of step k+1 runs while step k is still
            # computing (the slot -- and with it this block -- was last used HOST_SLOTS steps ago, and that step has retired:
            # its successor's loss has been read).  The step's graph waits for the copy's event.
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(self._copy_stream):
                if use_pixels:     # [pixels | projs | mask] are contiguous: one H2D copy
                    sv["packed"][o_pix:end].copy_(h["packed"][o_pix:end], non_blocking=True)
                else:
                    sv["packed"][o_rays:o_rays + 32 * N].copy_(h["packed"][o_rays:o_rays + 32 * N], non_blocking=True)
                    sv["packed"][o_projs:end].copy_(h["packed"][o_projs:end], non_blocking=True)
                h["copied"].record(self._copy_stream)
            torch.cuda.current_stream().wait_event(h["copied"])

            def body(with_optimizer):
                # the loss kernel stores its two floats straight into the slot's pinned (device-mapped) host buffer: no D2H copy
                # node at the end of the graph, and the PCIe write is long done when the backward pass and the optimizer retire
                # (the fused forward + loss launch also raises the slot's completion word; without it -- fp32 SIMT arithmetic -- the
                #  event below is what result() waits for)
                self._whole_step(sv, par, None, None, with_optimizer=with_optimizer, use_pixels=use_pixels, loss_out=h["loss"],
                                 done_flag=(h["loss"].data_ptr() + 8) if s["stash"] is not None else None)

            shape_key = (N, mask is not None, par, "host", use_pixels)
            key = shape_key + (slot,)
            g = self._graphs.get(key)
            if not self.use_cuda_graph or self._eager_runs.get(shape_key, 0) < 1:
                self._eager_runs[shape_key] = self._eager_runs.get(shape_key, 0) + 1
                body(True)
            else:
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        body(in_graph)
                    self._graphs[key] = g
                g.replay()
                if not in_graph:
                    self._finish_step(par)
            expect = self.step_count + 1          # what the kernel stores: device step count (completed steps) + 1
            self.step_count += 1
            h["event"].record()
        h["pending"] = PendingLoss(h["event"], h["loss_np"], h["flag_np"] if s["stash"] is not None else None, expect)
        return h["pending"].result() if wait else h["pending"]

    def check_health(self):
        """Raise if a kernel of a past step reported a failure it could not signal otherwise: the backward kernel's grid barrier
        timed out (a CTA never became resident: the MLP gradients of that step are missing), or -- several GPUs -- a bounded
        spin of the exchange kernel gave up on a peer (replicas have diverged).  Synchronises the stream; call it wherever the
        host already waits (evaluation, checkpoints)."""
        bad = int(self._bwd_ws[-64:].view(torch.int32)[2].item())
        if bad:
            raise RuntimeError("density_backward: the grid-wide barrier timed out (workspace shared between concurrent launches, or the "
                               "grid was not resident); MLP gradients of at least one step were lost")
        if self.px is not None:
            w = self.px.error_word()
            if w:
                raise RuntimeError(f"gradient exchange: rank {self.rank} gave up waiting for flag word {w - 1} of a peer; replicas have diverged")

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def render_projection(self, rays=None, t_rand=None, perturb=None, pixels=None):
        """acc [N] for rays [N,8] -- or for detector pixels [N,3] (projection, row, col), the rays being generated in the
        kernel (set_geometry) -- forward only; the reference's eval path renders a whole view this way (train.py:235-240)."""
        perturb = self.perturb if perturb is None else perturb
        if pixels is not None:
            pixels = pixels.reshape(-1, 3).to(torch.int32).contiguous()
            N = pixels.shape[0]
        else:
            rays = rays.reshape(-1, 8).contiguous()
            N = rays.shape[0]
        smp = self._ray_sampler(rays, pixels, t_rand if perturb else None)
        smp.perturb = int(bool(perturb))
        if not perturb:
            smp.rng_state = None
        acc = torch.zeros(N, device=self.device, dtype=torch.float32)
        grid, mlp = self.meta.grid(self.table), self.meta.mlp(self.mlp_params)
        _lib.check(_lib.lib().nafb_density_forward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), _lib.SRC_RAYS, None, _lib.ptr(acc),
                                                   None, None, None, None, _lib.stream_ptr()))
        return acc

    @torch.no_grad()
    def render_view(self, view: int, H: int, W: int, perturb=None, shard: bool = True):
        """Projection `view` rendered for every detector pixel (reference train.py:235-240: the rays of the chosen view in chunks of
        n_rays through render()), rays generated in-kernel.  With several ranks and shard=True the H*W pixels are SPLIT over the
        ranks (contiguous ranges, parallel.shard_range) and the line integrals all-gathered: every rank returns the whole [H,W]
        image having rendered 1/W of it.  Collective when sharded."""
        n = H * W
        i0, i1 = parallel.shard_range(n, self.rank, self.world_size) if (shard and self.world_size > 1) else (0, n)
        flat = torch.arange(i0, i1, device=self.device, dtype=torch.int64)
        pixels = torch.stack([torch.full_like(flat, int(view)), flat // W, flat % W], dim=-1).to(torch.int32)
        local = self.render_projection(pixels=pixels, perturb=perturb) if i1 > i0 else torch.zeros(0, device=self.device)
        if shard and self.world_size > 1:
            local = parallel.gather_shards(local, n, self.pg)
        return local.reshape(H, W)

    @torch.no_grad()
    def eval_step(self, view: int, proj_gt: torch.Tensor, image_gt: torch.Tensor, n_voxel, s_half, perturb=None, shard: bool = True):
        """The reference's eval_step (train.py:220-286) without its host round trips: render projection `view` (every
        detector pixel, rays generated in-kernel; split over the ranks and all-gathered when there are several), query the
        whole volume, score both on the device.
        proj_gt [H,W], image_gt [n1,n2,n3].  Returns {"proj_mse", "proj_psnr", "psnr_3d", "ssim_3d", "projs_pred", "image_pred"}."""
        from .utils import get_mse, get_psnr, get_psnr_3d, get_ssim_3d
        self.check_health()
        H, W = proj_gt.shape
        projs_pred = self.render_view(view, H, W, perturb=perturb, shard=shard)
        image_pred = self.voxel_query(n_voxel, s_half)
        out = {"proj_mse": float(get_mse(projs_pred, proj_gt)), "proj_psnr": float(get_psnr(projs_pred, proj_gt)),
               "psnr_3d": get_psnr_3d(image_pred, image_gt), "projs_pred": projs_pred, "image_pred": image_pred}
        out["ssim_3d"] = get_ssim_3d(image_pred, image_gt) if min(image_gt.shape) >= 7 else float("nan")
        return out

    @torch.no_grad()
    def voxel_query(self, n_voxel, s_half, slab=None, out=None):
        """Full-volume density query (reference train.py:246-250 + tigre.py:388-400) without ever
        materialising the [n1,n2,n3,3] coordinate tensor: positions are generated in-kernel from
        the float64 linspace end points ``s_half = sVoxel/2 - dVoxel/2``.

        slab=(i0, i1) restricts the outermost index (multi-GPU: z-slab sharding, no collective).
        Returns a tensor [(i1-i0), n2, n3]."""
        n1, n2, n3 = [int(v) for v in n_voxel]
        i0, i1 = (0, n1) if slab is None else (int(slab[0]), int(slab[1]))
        res = density_forward(self.meta, self.table, self.mlp_params, sigma_out=out,
                              voxels=(n1, n2, n3, i0, i1, float(s_half[0]), float(s_half[1]), float(s_half[2])))
        return res["sigma"].view(i1 - i0, n2, n3)

    def rank_slab(self, n1):
        """[i0, i1) of the outermost voxel index owned by this rank."""
        return parallel.shard_range(n1, self.rank, self.world_size)

    # ------------------------------------------------------------------ optimizer state (ckpt compatibility)
    def _full_state(self, t):
        """exp_avg / exp_avg_sq as a full-length vector (peer exchange keeps only the owned slice per rank)."""
        if self.px is None:
            return t.clone()
        full = torch.zeros(self.n_params, device=self.device, dtype=torch.float32)
        i0, i1 = self.px.slice
        full[i0:i1] = t
        if self.world_size > 1:
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.pg)   # slices are disjoint
        return full

    def optimizer_moments(self):
        """(step, exp_avg, exp_avg_sq) with the moments as FLAT full-length vectors in the layout of flat_param (diagnostics and
        tests; collective with a peer-memory exchange, like optimizer_state_dict)."""
        return self.step_count, self._full_state(self.exp_avg), self._full_state(self.exp_avg_sq)

    def _param_slices(self):
        """(parameter, offset, numel) of the module's parameters inside the flat vector, in net.parameters() order."""
        params = [self.net.encoder.embeddings] + self.net.flat_params()
        out, n = [], 2
        for p in params:
            out.append((p, n, p.numel()))
            n += _round_up(p.numel(), 4)
        return out

    def optimizer_state_dict(self):
        """The optimizer state in the layout of ``torch.optim.Adam(net.parameters(), lr, betas).state_dict()`` -- what the reference
        trainer stores under ckpt["optimizer"] (src/trainer.py:118-126) and restores with optimizer.load_state_dict (:60-70): per
        parameter {"step", "exp_avg", "exp_avg_sq"} in net.parameters() order + "param_groups".  A ckpt.tar written by the
        reference trainer resumes here and vice versa.
        COLLECTIVE with a peer-memory exchange (each rank holds the moments of its slice only: they are all-gathered): call it on
        EVERY rank, then let rank 0 save."""
        self.check_health()
        m, v = self._full_state(self.exp_avg), self._full_state(self.exp_avg_sq)
        opt = torch.optim.Adam(list(self.net.parameters()), lr=self._lr, betas=self.betas, eps=self.eps)
        if self.step_count > 0:
            for p, o, n in self._param_slices():
                opt.state[p] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m[o:o + n].view_as(p).clone(),
                                "exp_avg_sq": v[o:o + n].view_as(p).clone()}
        return opt.state_dict()

    def load_optimizer_state_dict(self, sd):
        """Accepts a torch.optim.Adam state_dict (the reference's ckpt["optimizer"]; int or tensor `step`) -- or the flat dict
        {"step", "exp_avg", "exp_avg_sq", "lr"} of the first round.  Collective with a peer-memory exchange."""
        if "param_groups" in sd:
            slices = self._param_slices()
            group = sd["param_groups"][0]
            if len(group["params"]) != len(slices):
                raise ValueError(f"optimizer state has {len(group['params'])} parameters, the network {len(slices)}")
            m = torch.zeros(self.n_params, device=self.device, dtype=torch.float32)
            v = torch.zeros_like(m)
            step = 0
            for pid, (p, o, n) in zip(group["params"], slices):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state of parameter {pid} has shape {tuple(st['exp_avg'].shape)}, expected {tuple(p.shape)}")
                m[o:o + n] = st["exp_avg"].reshape(-1).to(self.device, torch.float32)
                v[o:o + n] = st["exp_avg_sq"].reshape(-1).to(self.device, torch.float32)
                step = max(step, int(float(st["step"])))
            sd = {"step": step, "exp_avg": m, "exp_avg_sq": v, "lr": float(group["lr"])}
            self.betas = (float(group["betas"][0]), float(group["betas"][1]))
            self.eps = float(group["eps"])
        self._set_step(sd["step"])
        i0, i1 = self.px.slice if self.px is not None else (0, self.n_params)
        self.exp_avg.copy_(sd["exp_avg"][i0:i1])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"][i0:i1])
        self.lr = float(sd.get("lr", self._lr))
        self._graphs.clear()        # betas / eps are kernel arguments of the captured launches
        self._eager_runs.clear()
        if self.px is not None:
            for g in self.flat_grads:   # the parity of step_count selects the gradient buffer: both must be clean
                g.zero_()
            # the optimizer step doubles as the synchronisation epoch of the exchange kernel: after a rollback the flags of later
            # epochs would satisfy its waits immediately -> clear them on every rank, between barriers
            self.px.reset_flags(self.pg)
