"""NAFEngine -- the opt-in fused training / inference driver ("an opt-in
fused train_step entry that bypasses autograd").

One optimisation step is four kernel launches, captured once in a CUDA graph:

    density_forward (rays -> sampling -> gather -> MLP -> sum sigma*delta -> acc[N])
    mse_loss        (masked chunk-wise MSE + d loss / d acc)
    density_backward(recompute -> MLP backward -> scatter into the flat gradient) + reduce
    adam_step       (fused dense Adam over ONE flat vector [table | MLP], zeroing the gradient)

It takes the place of ``Trainer.train_step`` (reference src/trainer.py:134-142) +
``BasicTrainer.compute_loss`` (train.py:48-135) for the shipped configurations.  Parameters stay
ordinary ``nn.Parameter`` views of the flat vector, so ``state_dict()`` / ``ckpt.tar`` round-trip
with the reference layout (encoder.embeddings, layers.i.weight/bias).

Multi-GPU: one process per GPU, rays sharded across ranks, parameters/optimizer state replicated;
the flat gradient is all-reduced (sum) with NCCL and Adam divides by world size (== DDP's mean).
The voxel query shards by slabs of the outermost index and needs no collective.
"""
from __future__ import annotations

import ctypes

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


class NAFEngine:
    def __init__(self, net: DensityNetwork, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, n_samples=192, perturb=True, loss_chunk=None,
                 use_cuda_graph=True, process_group=None, use_stash=True):
        meta = net.fused_meta()
        if meta is None:
            raise RuntimeError("NAFEngine needs a DensityNetwork in a fused-capable configuration "
                               "(hash grid with L*C == 32, hidden_dim == 32, out_dim == 1)")
        dev = net.encoder.embeddings.device
        if dev.type != "cuda":
            raise RuntimeError("NAFEngine needs the network on a CUDA device (there is no CPU path)")
        _lib.lib()
        self.net, self.meta, self.device = net, meta, dev
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.n_samples, self.perturb, self.loss_chunk = int(n_samples), bool(perturb), loss_chunk
        self.step_count = 0
        self.use_cuda_graph = use_cuda_graph
        self.use_stash = bool(use_stash)  # forward leaves the encodings (128 B/point) for backward instead of a second gather
        self.pg = process_group
        self.rank, self.world_size = parallel.world_info(process_group)
        self._flatten()
        if self.world_size > 1:   # replicas start identical
            parallel.broadcast_(self.flat_param, 0, process_group)
        self._graphs = {}
        self._static = {}

    # ------------------------------------------------------------------ flat parameter vector
    def _flatten(self):
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
        self.flat_param = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self._views, self._grad_views = [], []
        for p, o in zip(params, offs):
            v = self.flat_param[o:o + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v  # the module's parameters now alias the flat vector
            self._views.append(v)
            self._grad_views.append(self.flat_grad[o:o + p.numel()].view_as(p))
        self.n_params = n
        self.table = self._views[0]
        self.mlp_params = self._views[1:]
        self.grad_table = self._grad_views[0]
        self.grad_mlp = self._grad_views[1:]

    # ------------------------------------------------------------------ one training step
    def _step_kernels(self, rays, projs, mask, t_rand, loss_out, dacc, acc, timer=None, stash=None):
        L_ = _lib.lib()
        tm = timer or _NoTimer()
        N = rays.shape[0]
        # acc is zero on entry: allocated zeroed, and mse_loss clears it after reading (zero_pred)
        grid = self.meta.grid(self.table)
        mlp = self.meta.mlp(self.mlp_params)
        smp = self.meta.sampler(rays=rays.data_ptr(), t_rand=t_rand.data_ptr() if self.perturb else None, n_rays=N,
                                n_samples=self.n_samples, perturb=int(self.perturb))
        st = _lib.stream_ptr()
        with tm("density_fwd"):
            _lib.check(L_.nafb_density_forward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), _lib.SRC_RAYS, None, _lib.ptr(acc),
                                               None, None, None, _lib.ptr(stash), st))
        chunk = int(self.loss_chunk or 0)
        with tm("mse_loss"):
            _lib.check(L_.nafb_mse_loss(_lib.ptr(acc), _lib.ptr(projs), _lib.ptr(mask), N, chunk, 1.0, _lib.ptr(loss_out), _lib.ptr(dacc), 1, st))
        with tm("density_bwd"):
            density_backward(self.meta, self.table, self.mlp_params, dacc, self.grad_table, self.grad_mlp, rays=rays, t_rand=t_rand,
                             n_samples=self.n_samples, perturb=self.perturb, stash=stash)

    def _adam(self, timer=None):
        L_ = _lib.lib()
        self.step_count += 1
        with (timer or _NoTimer())("adam"):
            _lib.check(L_.nafb_adam_step(_lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                         _lib.ptr(self.exp_avg_sq), self.n_params, self.lr, self.betas[0], self.betas[1], self.eps,
                                         self.step_count, 1.0 / self.world_size, 1, _lib.stream_ptr()))

    # kernels of this library launched by one train_step (density_fwd, mse_loss, density_bwd, reduce_partials, adam)
    LAUNCHES_PER_STEP = 5

    def profiled_step(self, rays, projs, mask, t_rand, timer):
        """Same work as train_step, launched eagerly (no graph) with a CUDA-event pair around every kernel."""
        N = rays.shape[0]
        s = self._get_static(N, mask is not None)
        with torch.cuda.device(self.device):
            s["rays"].copy_(rays.reshape(N, 8))
            s["projs"].copy_(projs.reshape(N))
            if mask is not None:
                s["mask"].copy_(mask.reshape(N))
            if self.perturb:
                with timer("t_rand"):
                    if t_rand is None:
                        s["t_rand"].uniform_(0.0, 1.0)
                    else:
                        s["t_rand"].copy_(t_rand)
            self._step_kernels(s["rays"], s["projs"], s["mask"], s["t_rand"], s["loss"], s["dacc"], s["acc"], timer, stash=s["stash"])
            if self.world_size > 1:
                with timer("all_reduce"):
                    parallel.allreduce_sum_(self.flat_grad, self.pg)
            self._adam(timer)
        return s["loss"][0]

    def _get_static(self, N, with_mask):
        key = (N, with_mask)
        s = self._static.get(key)
        if s is None:
            d = self.device
            nb = stash_bytes(self.meta, self.table, self.mlp_params, N * self.n_samples) if self.use_stash else 0
            s = dict(rays=torch.zeros(N, 8, device=d), projs=torch.zeros(N, device=d),
                     stash=torch.empty(nb, dtype=torch.uint8, device=d) if nb else None,
                     mask=torch.ones(N, device=d, dtype=torch.uint8) if with_mask else None,
                     t_rand=torch.zeros(N, self.n_samples, device=d) if self.perturb else None,
                     loss=torch.zeros(2, device=d), dacc=torch.zeros(N, device=d), acc=torch.zeros(N, device=d))
            self._static[key] = s
        return s

    def train_step(self, rays, projs, mask=None, t_rand=None):
        """rays [N,8], projs [N] (+ optional uint8/bool mask [N], uniforms t_rand [N,S]) on the device.
        Returns the loss as a 0-dim device tensor (no host synchronisation)."""
        N = rays.shape[0]
        s = self._get_static(N, mask is not None)
        with torch.cuda.device(self.device):
            s["rays"].copy_(rays.reshape(N, 8), non_blocking=True)
            s["projs"].copy_(projs.reshape(N), non_blocking=True)
            if mask is not None:
                s["mask"].copy_(mask.reshape(N), non_blocking=True)
            if self.perturb:
                if t_rand is None:
                    s["t_rand"].uniform_(0.0, 1.0)  # on-device Philox, same distribution as torch.rand (render.py:99)
                else:
                    s["t_rand"].copy_(t_rand, non_blocking=True)
            self._run_fwd_bwd(s, (N, mask is not None))
            parallel.allreduce_sum_(self.flat_grad, self.pg)
            self._adam()
        return s["loss"][0]

    def _run_fwd_bwd(self, s, key):
        if not self.use_cuda_graph:
            self._step_kernels(s["rays"], s["projs"], s["mask"], s["t_rand"], s["loss"], s["dacc"], s["acc"], stash=s["stash"])
            return
        g = self._graphs.get(key)
        if g is None:
            # warm up on a side stream (lazy kernel attribute setup must not happen under capture)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._step_kernels(s["rays"], s["projs"], s["mask"], s["t_rand"], s["loss"], s["dacc"], s["acc"], stash=s["stash"])
                self.flat_grad.zero_()  # the gradient is zero between steps (Adam clears it)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_kernels(s["rays"], s["projs"], s["mask"], s["t_rand"], s["loss"], s["dacc"], s["acc"], stash=s["stash"])
            self._graphs[key] = g
            # capture does not execute: replay below performs the first real step
        g.replay()

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def render_projection(self, rays, t_rand=None, perturb=None):
        """acc [N] for rays [N,8] (forward only; reference eval path train.py:235-240)."""
        perturb = self.perturb if perturb is None else perturb
        rays = rays.reshape(-1, 8).contiguous()
        if perturb and t_rand is None:
            t_rand = torch.rand(rays.shape[0], self.n_samples, device=self.device)
        out = density_forward(self.meta, self.table, self.mlp_params, rays=rays, t_rand=t_rand, n_samples=self.n_samples, perturb=perturb,
                              want_acc=True, want_sigma=False)
        return out["acc"]

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
    def optimizer_state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "lr": self.lr}

    def load_optimizer_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr = float(sd.get("lr", self.lr))
