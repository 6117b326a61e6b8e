#!/usr/bin/env python
"""bench.py -- headline benchmark of the NAF hot path (BASELINE.json: "train samples/sec").

    python bench.py --gpus N --steps K --warmup W            # this repository's sm_100a engine
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A "step" is one complete training iteration of config/chest_50.yaml's shape on synthetic data:
1024 rays x 192 samples -> stratified sampling -> 16x2 hash-grid encode (2^19 tables) -> 4x32 MLP
-> Beer-Lambert integral -> masked MSE -> backward -> dense Adam over 14.27 M parameters.
Throughput = n_gpus * n_rays * n_samples * K / (time of K steps, max over ranks).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: `roofline` (dominant kernel),
`kernels` (per-kernel CUDA-event means from an instrumented eager pass), `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_RAYS, N_SAMPLES = 1024, 192            # config/chest_50.yaml: train.n_rays, render.n_samples
N_PROJ, DET, N_VOXEL = 50, 256, 128      # 50 cone-beam projections of 256x256, 128^3 volume
LOSS_CHUNK = 200                         # train.py:56
WORKLOAD = ("chest_50: 1024 rays x 192 samples/step, hash grid 16x2 (2^19, base 16) + 4x32 MLP (skips [2], sigmoid), "
            "perturbed sampling, masked chunk(200) MSE, dense Adam over 14 266 663 fp32 params; 50 cone-beam projections 256x256")
FLOP_FWD_PER_POINT = 8256                # 2*(32*32 + 32*32 + 64*32 + 32*1), SURVEY.md 8d
FLOP_FWDBWD_PER_POINT = 24768
TABLE_BYTES_PER_POINT = 1024             # 16 levels x 8 corners x 2 ch x 4 B gathered (and the same reduced in backward)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback")


def ncu_summary(kernel):
    """Metrics of the newest committed `ncu --set full` summary of a kernel under profiles/ (scripts/ncu_summary.py output):
    {"file": ..., metric name: value in base units} or None."""
    import glob
    import re
    pat = {"density_bwd": "*density_bwd*_full.txt", "density_fwd": "*density_fwd*_full.txt", "density_fwd_loss": "*density_fwd*_full.txt",
           "adam": "*adam*_full.txt", "voxel_query": "*voxel*_full.txt"}.get(kernel)
    if not pat:
        return None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pat)), key=lambda f: (os.path.basename(f).split("_")[0], f))
    if not files:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
    out = {"file": os.path.relpath(files[-1], ROOT)}
    for line in open(files[-1]):
        m = re.match(r"([a-z0-9_.]+)\s+([0-9.eE+-]+)\s*([A-Za-z/%]*)", line)
        if m and m.group(1) not in out:
            out[m.group(1)] = float(m.group(2)) * mult.get(m.group(3), 1.0)
    return out


def ncu_dram_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel from its committed ncu summary (None when absent)."""
    d = ncu_summary(kernel)
    if not d or "dram__bytes_read.sum" not in d or "dram__bytes_write.sum" not in d:
        return None
    return {"bytes_per_launch": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"], "source": d["file"]}


# measured L2-resident random-access rates of this pool's B200 (scripts/microbench.py -> profiles/r1_microbench_l2_rates.json):
# what bounds the encoder's gather (address-divergent 8-byte loads) and scatter (8 / 16-byte float reductions), in G operations / s
L2_GATHER_GOPS, L2_RED_GOPS = 263.5, 184.0


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # pragma: no cover
            self.ok = False

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.01)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------- synthetic data
def synthetic_batches(n_batches, device, seed, rays_all_cpu=None):
    """chest_50-like cone-beam scan (the package's own geometry code) with U(0, 0.05) projection values; every batch =
    1024 distinct pixels of one projection (np.random.choice(replace=False), tigre.py:358) + a pixel mask.
    Returns pixel batches [B,N,3] (projection, row, col), the same rays gathered from the host-generated ray table
    [B,N,8], projections [B,N], masks [B,N] and (data, geo)."""
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    data = G.chest50_like(N_VOXEL, DET, N_PROJ)
    geo = G.ConeGeometry(data)
    rays_all = G.rays_with_near_far(data["angles"], geo, "cpu").to(device)   # [50,256,256,8] fp32, 105 MB on the device
    gen = torch.Generator(device=device).manual_seed(seed)
    projs_all = torch.rand(N_PROJ, DET, DET, device=device, generator=gen) * 0.05
    mask_all = torch.rand(N_PROJ, DET, DET, device=device, generator=gen) > 0.02  # ~2 % of the pixels masked out
    pix_b, rays_b, projs_b, mask_b = [], [], [], []
    for b in range(n_batches):
        p = b % N_PROJ
        pix = torch.randperm(DET * DET, device=device, generator=gen)[:N_RAYS]
        pix_b.append(torch.stack([torch.full_like(pix, p), pix // DET, pix % DET], 1).to(torch.int32))
        rays_b.append(rays_all[p].reshape(-1, 8)[pix])
        projs_b.append(projs_all[p].reshape(-1)[pix])
        mask_b.append(mask_all[p].reshape(-1)[pix])
    del rays_all
    return torch.stack(pix_b), torch.stack(rays_b), torch.stack(projs_b), torch.stack(mask_b).to(torch.uint8), (data, geo)


def build_engine(device):
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    torch.manual_seed(0)
    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(device)
    return NAFEngine(net, lr=1e-3, betas=(0.9, 0.999), n_samples=N_SAMPLES, perturb=True, loss_chunk=LOSS_CHUNK, use_cuda_graph=True,
                     exchange=os.environ.get("NAFB_EXCHANGE", "auto"))


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(steps, warmup, n_rays=N_RAYS):
    """The reference's training step on the host cores.

    kind "reference": the reference's OWN Python (render, DensityNetwork, HashEncoder module, calc_mse_loss; staged under
    baseline/_ref by baseline/stage_ref.sh) + torch.optim.Adam, with its one native op served by the reference's own
    kernel text compiled for the host (oracle/_ref, OpenMP) -- one render() call per step (the best case for the
    reference; train.py's 200-ray chunk loop is slower).
    kind "port": when that staging is absent, the oracle's restatement of the same code (oracle/).
    This is the one place bench.py executes oracle/ -- as the baseline being reported."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_rays
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    kind = "port"
    try:
        from baseline import ref_loader
        if ref_loader.available("host"):
            kind = "reference"
    except Exception:
        kind = "port"
    from oracle import hashgrid as oh
    if kind == "reference":
        get_encoder, get_network, render, calc_mse_loss = ref_loader.import_reference("host")
        enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
        net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))

        def one_step(rays, projs, mask):
            opt.zero_grad()
            loss = {"loss": 0.0}
            ret = render(rays, net, None, N_SAMPLES, 0, True, 409600, 0.0)
            calc_mse_loss(loss, projs[mask], ret["acc"][mask])
            loss["loss"].backward()
            opt.step()
    else:
        from oracle import naf
        enc = oh.OracleHashEncoder(3, 16, 2, 16, 19, use_ref=False, normalise="div")
        net = naf.OracleDensityNetwork(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))

        def one_step(rays, projs, mask):
            naf.train_step(net, opt, rays, projs, N_SAMPLES, True, mask=mask, chunk=LOSS_CHUNK)
    times = []
    for it in range(warmup + steps):
        rays = torch.from_numpy(make_rays(n_rays, rng))
        projs = torch.from_numpy(rng.uniform(0, 0.05, n_rays).astype(np.float32))
        mask = torch.from_numpy(rng.uniform(0, 1, n_rays) > 0.02)
        t0 = time.perf_counter()
        one_step(rays, projs, mask)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    return dict(value=n_rays * N_SAMPLES * steps / total, ms_per_step=1e3 * total / steps, cores=cores, threads=oh.num_threads(),
                n_rays=n_rays, kind=kind)


def cpu_extra_baselines():
    """Bounded CPU samples (host cores of this box) of the two other single-GPU workloads, through the same reference code as
    cpu_reference_run: one training step of 4 096 rays x 384 samples (large batch is 16x that), and the voxel query on 2^18 voxels of
    the 512^3 lattice (the query is 512x that)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    out = {}
    try:
        global N_SAMPLES
        keep = N_SAMPLES
        N_SAMPLES = 384
        try:
            r = cpu_reference_run(steps=2, warmup=1, n_rays=4096)
        finally:
            N_SAMPLES = keep
        out["large_batch"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                              "sample": "2 timed steps of 4096 rays x 384 samples after 1 warm-up (1/16 of the large-batch step)"}
    except Exception as e:   # noqa: BLE001
        out["large_batch"] = {"unavailable": str(e)[:200]}
    try:
        from oracle import hashgrid as oh
        from oracle import naf
        torch.manual_seed(0)
        enc = oh.OracleHashEncoder(3, 16, 2, 16, 19, use_ref=False, normalise="mul_recip")
        net = naf.OracleDensityNetwork(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
        n, m = 512, 64
        s = (n * 0.001) / 2 - 0.001 / 2
        lin = np.linspace(-s, s, n)
        xyz = np.stack(np.meshgrid(lin[:m], lin[:m], lin[:m], indexing="ij"), -1).astype(np.float32).reshape(-1, 3)
        with torch.no_grad():
            naf.run_network(torch.from_numpy(xyz[:65536]), net, 409600)
            t0 = time.perf_counter()
            naf.run_network(torch.from_numpy(xyz), net, 409600)
            dt = time.perf_counter() - t0
        out["voxel_query_512"] = {"value": xyz.shape[0] / dt, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "port",
                                  "ms_for_512_cubed": 1e3 * dt * (n ** 3) / xyz.shape[0],
                                  "sample": "64^3 voxels of the 512^3 lattice (1/512 of the query) through the oracle (C hash grid with OpenMP + torch-CPU MLP)"}
    except Exception as e:   # noqa: BLE001
        out["voxel_query_512"] = {"unavailable": str(e)[:200]}
    return out


def reference_cuda_numbers():
    """The reference's own CUDA build timed on this GPU (baseline/ref_cuda_bench.py in a subprocess, ~2 s), when it is staged
    under baseline/_ref: the denominator of BASELINE.json's ">= 20x the reference's CUDA build" target, reported next to our
    number in the same run.  None when the staging is absent."""
    import subprocess
    script = os.path.join(ROOT, "baseline", "ref_cuda_bench.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "build")):
        return None
    try:
        r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=300, cwd=ROOT,
                           env=dict(os.environ, REF_STEPS="20", REF_WARMUP="5", REF_EXTRA="1"))
        out = {}
        for l in r.stdout.splitlines():
            if l.startswith("{"):
                d = json.loads(l)
                out[d["variant"]] = {k: d[k] for k in ("value", "unit", "ms_per_step", "ms") if k in d}
        if not out:
            return None
        out["note"] = ("the reference's render / DensityNetwork / HashEncoder (its CUDA extension, 2-line compile fix) / calc_mse_loss + "
                       "torch.optim.Adam on the same workload: 'chunked' = train.py's 200-ray chunk loop, 'one_call' = one render() per step; "
                       "'large_batch' = one 65536 x 384 step, 'voxel_query_512' = run_network over the 512^3 lattice")
        return out
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bound the run: a full 1024x192 step costs ~0.3-0.6 s on 8-16 cores
    n_rays = N_RAYS if args.steps + args.warmup <= 60 else max(64, (N_RAYS * 60) // (args.steps + args.warmup) // 8 * 8)
    r = cpu_reference_run(args.steps, args.warmup, n_rays)
    what = ("the reference's own Python + its hash-grid kernel text compiled for the host (OpenMP)" if r["kind"] == "reference"
            else "oracle port of the reference")
    sample = (f"{args.steps} timed steps of {n_rays} rays x {N_SAMPLES} samples after {args.warmup} warm-up, full 16x2/2^19 table, "
              f"dense Adam; {what}")
    line = {
        "impl": "reference", "metric": "train samples/sec (rays x samples / s)", "value": r["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": what + " on the host cores (the reference's CUDA op has no CPU path of its own)"},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# --------------------------------------------------------------------------------------- the other BASELINE.json configs
def lamino_like(n_proj=187, det_w=356, det_h=256):
    """config/lamino_chip.yaml-like geometry: parallel beam, 29 degree laminography tilt, 187 angles
    0.72..179.28 degrees (the grid of the reference's data/angles_real.npy), 256x356 detector, 356x356x70 volume."""
    return dict(DSD=1500.0, DSO=1000.0, nDetector=[det_w, det_h], dDetector=[1.0, 1.0], nVoxel=[356, 356, 70], dVoxel=[1.0, 1.0, 1.0],
                offOrigin=[0, 0, 0], offDetector=[0, 0], accuracy=0.5, mode="parallel", filter=None, tilt_angle=29,
                angles=np.deg2rad(0.72 + 0.96 * np.arange(n_proj)))


def _time_steps(eng, rays_b, projs_b, mask_b, K, W, world):
    import torch.distributed as dist
    nb = rays_b.shape[0]
    W = max(W, 2 * nb + 1) // nb * nb      # every resident batch at least twice (graph capture on the second visit), whole cycles
    for i in range(W):
        eng.train_step(rays_b[i % nb], projs_b[i % nb], mask_b[i % nb])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = eng.train_step(rays_b[(W + i) % nb], projs_b[(W + i) % nb], mask_b[(W + i) % nb])
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=rays_b.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / K, float(loss.item())


def extra_workloads(device, world, rank):
    """The remaining configurations of BASELINE.json, measured briefly after the headline run (same engine, same
    kernels): tilted-laminography training, large-batch training (65536 x 384, 256^3 volume) and the 512^3 voxel query
    (z-slab sharded, no collective).  All ranks take part; times are the max over ranks."""
    import torch.distributed as dist
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    out = {}
    gen = torch.Generator(device=device).manual_seed(99 + rank)

    def make_engine(n_samples, chunk):
        torch.manual_seed(0)
        enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
        net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(device)
        return NAFEngine(net, lr=1e-3, n_samples=n_samples, perturb=True, loss_chunk=chunk, use_cuda_graph=True)

    def batches(rays_all, n_rays, n_b):
        flat = rays_all.reshape(-1, 8)
        rb, pb, mb = [], [], []
        for _ in range(n_b):
            pix = torch.randint(0, flat.shape[0], (n_rays,), device=device, generator=gen)
            rb.append(flat[pix])
            pb.append(torch.rand(n_rays, device=device, generator=gen) * 0.05)
            mb.append((torch.rand(n_rays, device=device, generator=gen) > 0.02).to(torch.uint8))
        return torch.stack(rb), torch.stack(pb), torch.stack(mb)

    # ---- lamino_chip-like: parallel beam + tilt, masked loss, 1024 x 192
    data = lamino_like()
    geo = G.ConeGeometry(data)
    rays_all = G.rays_with_near_far(data["angles"][rank::max(world, 1)][:24], geo, device)    # 24 projections per rank are plenty
    eng = make_engine(192, LOSS_CHUNK)
    rb, pb, mb = batches(rays_all, 1024, 16)
    ms, loss = _time_steps(eng, rb, pb, mb, 40, 5, world)
    out["lamino_chip"] = {"workload": "tilted laminography (parallel beam, tilt 29 deg, 256x356 detector, 187 angles), 1024 rays x 192 samples, "
                          "masked chunk(200) MSE", "ms_per_step": ms, "value": world * 1024 * 192 / (ms * 1e-3), "unit": "samples/s", "loss": loss}
    del eng, rays_all, rb, pb, mb
    # ---- large batch: 65536 rays x 384 samples per step per GPU, 256^3 volume
    data = G.chest50_like(256, 256, 50)
    geo = G.ConeGeometry(data)
    rays_all = G.rays_with_near_far(data["angles"][:16], geo, device)
    eng = make_engine(384, None)
    rb, pb, mb = batches(rays_all, 65536, 3)
    ms, loss = _time_steps(eng, rb, pb, mb, 4, 4, world)    # 4 warm-up steps: eager + graph capture for both gradient parities
    pts = 65536 * 384
    pts = 65536 * 384
    out["large_batch"] = {"workload": "65536 rays x 384 samples per step per GPU, 256^3 volume, cone beam, one MSE chunk", "ms_per_step": ms,
                          "value": world * pts / (ms * 1e-3), "unit": "samples/s", "loss": loss,
                          "stash_gb": (eng._static[(65536, True)]["stash"].numel() / 1e9) if eng._static[(65536, True)]["stash"] is not None else 0.0}
    # per-kernel times of the same step (instrumented eager pass) and the roofline of its dominant kernel
    from neuralvolumetricreconstructionformedicalimages_b200.engine import EventTimer
    tm = EventTimer()
    eng.profiled_step(rb[0], pb[0], mb[0], None, EventTimer())
    for i in range(2):
        eng.profiled_step(rb[i % 3], pb[i % 3], mb[i % 3], None, tm)
    torch.cuda.synchronize()
    prof = {k: v[0] for k, v in tm.summary().items()}
    out["large_batch"]["kernels_ms"] = prof
    dom = max(prof, key=prof.get)
    out["large_batch"]["roofline"] = {
        "kernel": dom, "bound": "hbm", "achieved": TABLE_BYTES_PER_POINT * pts / (prof[dom] * 1e-3) / 1e9, "peak": measured_peaks()["hbm"], "unit": "GB/s",
        "frac": TABLE_BYTES_PER_POINT * pts / (prof[dom] * 1e-3) / 1e9 / measured_peaks()["hbm"], "traffic": None,
        "note": "algorithmic bytes = 1024 B/point of table entries reduced (backward) or gathered (forward); the table is L2-resident, the "
                "limiter is the L2 operation rate: see `limiter`",
        "limiter": {"bound": "l2_atomic" if "bwd" in dom else "l1tex_gather", "ops_per_point_upper_bound": 128,
                    "achieved_gops_upper_bound": 128 * pts / (prof[dom] * 1e-3) / 1e9, "peak_gops": L2_RED_GOPS if "bwd" in dom else L2_GATHER_GOPS}}
    del rays_all, rb, pb, mb
    # ---- 512^3 voxel query, slabs of the outermost index per rank (no collective)
    n = 512
    s_half = (n * 0.001) / 2 - 0.001 / 2            # sVoxel/2 - dVoxel/2 with 1 mm voxels: 0.2555 m < bound 0.3
    i0, i1 = eng.rank_slab(n)
    vol = torch.empty(i1 - i0, n, n, device=device)
    eng.voxel_query((n, n, n), (s_half,) * 3, slab=(i0, i1), out=vol)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.voxel_query((n, n, n), (s_half,) * 3, slab=(i0, i1), out=vol)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vq_ms = float(t[0])
    vox_per_rank = (i1 - i0) * n * n
    gbs = TABLE_BYTES_PER_POINT * vox_per_rank / (vq_ms * 1e-3) / 1e9
    nc = ncu_summary("voxel_query")
    out["voxel_query_512"] = {"workload": f"forward-only fused encode+MLP over the 512^3 lattice, outermost-index slabs over {world} GPU(s), fp32 output",
                              "ms": vq_ms, "value": n ** 3 / (vq_ms * 1e-3), "unit": "voxels/s", "mean_sigma": float(vol.mean().item()),
                              "roofline": {"kernel": "k_density_fwd_tc<VOXELS>", "bound": "hbm", "achieved": gbs, "peak": measured_peaks()["hbm"],
                                           "unit": "GB/s", "frac": gbs / measured_peaks()["hbm"],
                                           "traffic": (nc["dram__bytes_read.sum"] + nc["dram__bytes_write.sum"]) if nc and "dram__bytes_read.sum" in nc else None,
                                           "traffic_source": (nc["file"] + " (128^3 query: scale by 64)") if nc else None,
                                           "algorithmic_bytes_per_launch": TABLE_BYTES_PER_POINT * vox_per_rank + 4 * vox_per_rank,
                                           "note": "1024 B of table entries gathered per voxel (+ 4 B written); the 57 MB table is L2-resident, so the "
                                                   "applicable limit is the address-divergent load rate of the L1TEX / L2 path: see `limiter`",
                                           "limiter": {"bound": "l1tex_gather", "achieved_gops_upper_bound": 128 * vox_per_rank / (vq_ms * 1e-3) / 1e9,
                                                       "peak_gops": L2_GATHER_GOPS,
                                                       "note": "128 corner loads per voxel before x-neighbour pair merging and L1 hits (an 8x4x4 block of "
                                                               "the lattice per tile shares sectors); the random-access micro-benchmark is not a ceiling "
                                                               "for spatially coherent points"}}}
    del eng, vol
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=20)
    ap.add_argument("--source", default="pixels", choices=["pixels", "rays"],
                    help="pixels: batches are (projection,row,col) and the kernels generate the rays; rays: [N,8] ray tensors")
    ap.add_argument("--no-extra", action="store_true", help="skip the lamino / large-batch / voxel-query side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this package has no CPU fallback (use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created; stdout must carry exactly one
        # JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    K, W = args.steps, args.warmup

    eng = build_engine(device)
    n_b = min(K + W, 256)
    pix_b, rays_b, projs_b, mask_b, (data, geo) = synthetic_batches(n_b, device, seed=1234 + rank)
    eng.set_geometry(data["angles"], geo)
    use_pixels = args.source == "pixels"     # the kernels generate the rays of (projection, row, col) themselves
    in_b = pix_b if use_pixels else rays_b
    # host copies in pinned memory for the end-to-end leg
    in_h, projs_h, mask_h = in_b.cpu().pin_memory(), projs_b.cpu().pin_memory(), mask_b.cpu().pin_memory()

    def step(inp, projs, mask):
        return eng.train_step(None, projs, mask, pixels=inp) if use_pixels else eng.train_step(inp, projs, mask)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()

    # ---------------- device-resident throughput ("value")
    # warm-up: at least W steps, and every resident batch twice (the engine captures one CUDA graph per resident batch on its
    # second visit and replays it from then on: the timed steps below launch nothing but graphs)
    for i in range(max(W, 2 * n_b)):
        step(in_b[i % n_b], projs_b[i % n_b], mask_b[i % n_b])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        j = (W + i) % n_b
        loss = step(in_b[j], projs_b[j], mask_b[j])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    final_loss = float(loss.item())
    clocks = sampler.finish()

    # ---------------- end to end: host inputs -> H2D -> step -> D2H of the loss, every step, through the engine's host entry
    # (NAFEngine.train_step_host: staging in pinned memory, one graph launch = H2D copies + iteration + D2H of the loss,
    #  stream synchronisation, python float back)
    def step_host(j):
        return (eng.train_step_host(projs_h[j], mask_h[j], pixels=in_h[j]) if use_pixels
                else eng.train_step_host(projs_h[j], mask_h[j], rays=in_h[j]))

    for i in range(12):    # first use of a gradient parity runs eagerly, then one graph capture per (parity, staging slot)
        step_host(i % n_b)
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(K):
        host_loss = step_host((W + i) % n_b)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    h2d = in_h[0].numel() * 4 + projs_h[0].numel() * 4 + mask_h[0].numel()
    d2h = 12     # loss, number of valid rays, completion word: written by the forward launch straight into pinned host memory

    # the same host entry used the way a training loop would: step k+1 is staged and enqueued before the loss of step k is read
    # (train_step_host(wait=False) -> PendingLoss; every step still copies its inputs H2D and its loss D2H inside the timed region)
    def step_host_nowait(j):
        return (eng.train_step_host(projs_h[j], mask_h[j], pixels=in_h[j], wait=False) if use_pixels
                else eng.train_step_host(projs_h[j], mask_h[j], rays=in_h[j], wait=False))

    barrier()
    t0 = time.perf_counter()
    pending = None
    for i in range(K):
        nxt = step_host_nowait((W + i) % n_b)
        if pending is not None:
            piped_loss = pending.result()
        pending = nxt
    piped_loss = pending.result()
    barrier()
    piped_wall_ms = 1e3 * (time.perf_counter() - t0)

    # ---------------- the same step with the batch DRAWN ON THE DEVICE inside the graph (NAFEngine.train_step_sampled: the reference's
    # per-iteration dataset work, tigre.py:354-382, as a kernel of the step): no per-step host input at all
    from neuralvolumetricreconstructionformedicalimages_b200.dataset.mask import PixelSampler
    gen_s = torch.Generator(device=device).manual_seed(4321 + rank)
    ps = PixelSampler(torch.rand(N_PROJ, DET, DET, device=device, generator=gen_s) * 0.05 + 1e-4,
                      torch.polar(torch.where(torch.rand(N_PROJ, DET, DET, device=device, generator=gen_s) < 0.02, 0.003, 1.0),
                                  torch.zeros(N_PROJ, DET, DET, device=device)), 0.007, seed=99 + rank)
    for i in range(5):
        eng.train_step_sampled(ps, N_RAYS)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(K):
        sampled_loss = eng.train_step_sampled(ps, N_RAYS)
    g1.record()
    barrier()
    sampled_ms = g0.elapsed_time(g1)
    ps.check()

    # ---------------- per-kernel CUDA-event timing (instrumented eager pass, same workload)
    from neuralvolumetricreconstructionformedicalimages_b200.engine import EventTimer
    timer = EventTimer()
    for i in range(3):
        eng.profiled_step(None if use_pixels else in_b[i], projs_b[i], mask_b[i], None, EventTimer(), pixels=in_b[i] if use_pixels else None)
    torch.cuda.synchronize()
    for i in range(args.profile_steps):
        j = i % n_b
        eng.profiled_step(None if use_pixels else in_b[j], projs_b[j], mask_b[j], None, timer, pixels=in_b[j] if use_pixels else None)
    torch.cuda.synchronize()
    prof = timer.summary()

    extra = None if args.no_extra else extra_workloads(device, world, rank)

    # ---------------- N > 1: are the replicas still identical, did a bounded spin of the exchange kernel give up?  (collective)
    from neuralvolumetricreconstructionformedicalimages_b200 import parallel
    torch.cuda.synchronize()
    replica_div = parallel.replica_divergence(eng.flat_param, eng.pg)
    err_word = torch.tensor([eng.px.error_word() if eng.px is not None else 0], device=device, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(err_word, op=dist.ReduceOp.MAX)
    eng.check_health()
    loss_all = float(parallel.combined_loss(loss, eng.pg).item())

    # ---------------- reduce over ranks (max time)
    t = torch.tensor([ms, e2e_wall_ms, piped_wall_ms, sampled_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_wall_ms, piped_wall_ms, sampled_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    pts_step = N_RAYS * N_SAMPLES
    value = world * pts_step * K / (ms * 1e-3)
    e2e_value = world * pts_step * K / (e2e_wall_ms * 1e-3)

    if rank == 0:
        peaks = measured_peaks()
        step_sum = sum(v[0] for v in prof.values())
        kernels = {}
        # algorithmic work per launch: table bytes gathered / reduced, MLP flops, optimizer bytes
        algo = {
            "density_fwd": (FLOP_FWD_PER_POINT * pts_step, TABLE_BYTES_PER_POINT * pts_step),
            "density_fwd_loss": (FLOP_FWD_PER_POINT * pts_step, TABLE_BYTES_PER_POINT * pts_step),
            "density_bwd": (FLOP_FWDBWD_PER_POINT * pts_step, TABLE_BYTES_PER_POINT * pts_step),
            "adam": (0, 32 * eng.n_params),
            # peer exchange: per rank 4 B/param zeroed + its 1/W slice: W gradient reads, p/m/v read+write, W parameter writes
            "adam_exchange": (0, 4 * eng.n_params + (eng.n_params // world) * (4 * world + 24 + 4 * world)),
        }
        for name, (mean_ms, cnt) in prof.items():
            k = {"ms": mean_ms, "share": mean_ms / step_sum}
            if name in algo:
                flops, nbytes = algo[name]
                k.update(tflops=flops / (mean_ms * 1e-3) / 1e12, algorithmic_gbs=nbytes / (mean_ms * 1e-3) / 1e9)
            kernels[name] = k
        # What actually bounds the encoder on this machine: the table is L2-resident, so neither kernel is near an
        # HBM byte roofline; the gather is bound by address-divergent L1TEX wavefronts / L2 sectors, the scatter by the rate at
        # which L2 performs float reductions.  Operation counts come from the committed ncu capture of the same kernel
        # (lts__t_sectors_srcunit_tex_op_{read,red}: sectors that reached L2 AFTER pair merging, warp aggregation and L1 hits).
        def limiter(name, metric, peak, label):
            nc = ncu_summary(name)
            if name not in kernels or not nc or metric not in nc:
                return None
            ops = nc[metric]
            g = ops / (kernels[name]["ms"] * 1e-3) / 1e9
            return {"bound": label, "ops_per_launch": ops, "ops_source": nc["file"] + ":" + metric, "achieved_gops": g, "peak_gops": peak,
                    "frac": g / peak, "peak_source": "profiles/r1_microbench_l2_rates.json (random 8-byte accesses over an L2-resident 57 MB buffer)"}
        for fk in ("density_fwd", "density_fwd_loss"):
            lim = limiter(fk, "lts__t_sectors_srcunit_tex_op_read.sum", L2_GATHER_GOPS, "l1tex_gather")
            if lim:
                kernels[fk]["limiter"] = lim
        lim = limiter("density_bwd", "lts__t_sectors_srcunit_tex_op_red.sum", L2_RED_GOPS, "l2_atomic")
        if lim:
            kernels["density_bwd"]["limiter"] = lim
        dom = max((n for n in kernels if n in algo), key=lambda n: kernels[n]["ms"])
        flops, nbytes = algo[dom]
        achieved = kernels[dom]["algorithmic_gbs"]
        traffic = ncu_dram_traffic(dom)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                    "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
                    "algorithmic_bytes_per_launch": nbytes, "peak_source": peaks["source"], "kernel_ms": kernels[dom]["ms"],
                    "kernel_share_of_step": kernels[dom]["share"],
                    "tensor_tflops": kernels[dom]["tflops"], "tensor_frac": kernels[dom]["tflops"] / peaks["tensor"],
                    "limiter": kernels[dom].get("limiter"),
                    "note": "algorithmic bytes = 1024 B/point of table entries reduced (gathered in forward) (SURVEY.md 8d); the 57 MB table "
                            "and the 28 MB stash are L2-resident, so DRAM traffic is far below the algorithmic bytes and `frac` against the "
                            "HBM peak is not the distance to the applicable roofline: that is `limiter` (L2 reduction / gather operation "
                            "rate against its measured peak)"}
        line = {
            "metric": "train samples/sec (rays x samples / s)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tables, sampling, loss, Adam: fp32; MLP products: bf16x3 split operands on tcgen05, fp32 accumulation in TMEM)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"dp{world} (rays sharded, parameters replicated; exchange: " + {
                           "push": "one fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (gradients pushed to their "
                                   "owners, parameters pushed back: stores only)",
                           "nvls": "one fused reduce-scatter + Adam + all-gather kernel over NVLink, gradients added in the NVSwitch "
                                   "(multimem.ld_reduce) and parameters multicast (multimem.st)",
                           "peer": "one fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (P2P loads / stores)",
                           "nccl": "NCCL all-reduce of the flat gradient + dense Adam", "local": "none (single GPU), dense Adam"}[eng.exchange_mode] + ")",
                       "per_gpu_batch": f"{N_RAYS}x{N_SAMPLES}", "cuda_graph": True,
                       "ray_source": "detector pixels, rays generated in-kernel" if use_pixels else "rays tensor [N,8]",
                       "sampler_uniforms": "in-kernel counter-based generator (no [N,S] buffer)",
                       "l2": "no explicit flush: every step streams 456 MB (dense Adam over 4 x 57 MB vectors) through the 126 MB L2; "
                             "each step uses a different ray batch"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_wall_ms / K, "device_event_ms_per_step": e2e_ms / K,
                    "note": "value = synchronous form: every step returns its own loss to the host before the next one is staged; the loss is "
                            "written by the step's forward launch and polled by the host, so the backward pass and the optimizer of step k "
                            "overlap the staging, H2D copy (copy stream) and launch of step k+1; the timed region ends with a full device "
                            "synchronisation",
                    "pipelined": {"value": world * pts_step * K / (piped_wall_ms * 1e-3), "unit": "samples/s", "ms_per_step": piped_wall_ms / K,
                                  "note": "train_step_host(wait=False): the loss of step k is read after step k+1 has been enqueued; "
                                          "same H2D / D2H bytes every step"}},
            "sampled": {"value": world * pts_step * K / (sampled_ms * 1e-3), "unit": "samples/s", "ms_per_step": sampled_ms / K,
                        "note": "NAFEngine.train_step_sampled: the pixel draw of the reference's dataset (non-zero pixels, without replacement, "
                                "mask lookup) is a kernel inside the step's CUDA graph; projections resident, 4 launches per step"},
            "gpu_launches": eng.launches_per_step * K,
            "roofline": roofline,
            "kernels": kernels,
            "final_loss": final_loss, "final_loss_all_ranks": loss_all, "host_loss": host_loss, "host_loss_pipelined": piped_loss,
            # N > 1 correctness, checked after the timed region: max |param - param on rank 0| over all ranks (0.0 = bit-identical
            # replicas) and the exchange kernel's error word (0 = no bounded spin ever gave up), max over ranks
            "replica_divergence": replica_div, "exchange_error_word": int(err_word.item()),
            "dp_loss_rule": "sum over ranks of the reference's chunked loss (= one GPU on the concatenated batch with the same chunk boundaries)",
        }
        if extra is not None:
            line["workloads"] = extra
        if world == 1 and not args.no_cpu_baseline:
            rc = reference_cuda_numbers()
            if rc:
                line["reference_cuda"] = rc
                for wk in ("large_batch", "voxel_query_512"):
                    if extra is not None and wk in rc and wk in extra:
                        extra[wk]["reference_cuda"] = rc[wk]
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=8, warmup=2)
            line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": "8 timed steps of 1024 rays x 192 samples after 2 warm-up ("
                                              + ("the reference's own Python with its hash-grid kernels compiled for the host, OpenMP"
                                                 if r["kind"] == "reference" else "oracle port: C hash grid with OpenMP + torch-CPU MLP/render/Adam") + ")"}
            if extra is not None:
                for wk, v in cpu_extra_baselines().items():
                    if wk in extra:
                        extra[wk]["cpu_baseline"] = v
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
