#!/usr/bin/env python
"""Time the REFERENCE's own CUDA build of the NAF training step on this GPU (the denominator of BASELINE.json's
">= 20x the reference's CUDA build" target).  Needs baseline/_ref/ (see baseline/stage_ref.sh; git-ignored, built in the
container that mounts /root/reference, shipped to the GPU box with the repository snapshot).

Runs the reference's render / DensityNetwork / HashEncoder / calc_mse_loss + torch.optim.Adam, unmodified except for the
2-line `.scalar_type()` compile fix, on the same synthetic chest_50 workload as bench.py, two ways:
  chunked   : train.py:69-127 as intended (200-ray chunks, masked MSE per chunk, coords slice corrected)
  one_call  : one render() call per step (best case for the reference)
Prints one JSON line per variant.  None of this repository's kernels are on that path.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def main():
    steps = int(os.environ.get("REF_STEPS", "30"))
    warmup = int(os.environ.get("REF_WARMUP", "5"))
    sys.path.insert(0, ROOT)
    from baseline.ref_loader import import_reference
    get_encoder, get_network, render, calc_mse_loss = import_reference("cuda")
    import bench
    dev = torch.device("cuda", 0)
    _, rays_b, projs_b, mask_b, _ = bench.synthetic_batches(steps + warmup, dev, seed=1234)
    mask_b = mask_b.bool()
    for variant in ("chunked", "one_call"):
        torch.manual_seed(0)
        enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
        net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))
        kw = dict(n_samples=bench.N_SAMPLES, n_fine=0, perturb=True, netchunk=409600, raw_noise_std=0.0)

        def step(i):
            rays, projs, mask = rays_b[i], projs_b[i], mask_b[i]
            opt.zero_grad()
            loss = {"loss": 0.0}
            if variant == "chunked":
                for c0 in range(0, rays.shape[0], bench.LOSS_CHUNK):            # train.py:69
                    sl = slice(c0, c0 + bench.LOSS_CHUNK)
                    ret = render(rays[sl], net, None, chunk_size=bench.LOSS_CHUNK, **kw)
                    m = mask[sl]
                    calc_mse_loss(loss, projs[sl][m], ret["acc"][m])             # train.py:127
            else:
                ret = render(rays, net, None, **kw)
                calc_mse_loss(loss, projs[mask], ret["acc"][mask])
            loss["loss"].backward()
            opt.step()
            return loss["loss"]

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            l = step(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"impl": "reference_cuda", "variant": variant, "ms_per_step": ms,
                          "value": bench.N_RAYS * bench.N_SAMPLES / (ms * 1e-3), "unit": "samples/s", "steps": steps, "warmup": warmup,
                          "final_loss": float(l.item()), "gpu": torch.cuda.get_device_name(0),
                          "workload": bench.WORKLOAD}), flush=True)


def extra():
    """REF_EXTRA=1: the reference's CUDA build on the two other single-GPU configurations of BASELINE.json: one large-batch
    training step (65 536 rays x 384 samples in one render() call) and the 512^3 voxel query (run_network over the materialised
    coordinate tensor in chunks of netchunk = 409 600, train.py:246-250)."""
    sys.path.insert(0, ROOT)
    from baseline.ref_loader import import_reference
    get_encoder, get_network, render, calc_mse_loss = import_reference("cuda")
    import bench
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    sys.path.insert(0, REF)
    from src.render import run_network
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))
    # ---- large batch
    data = G.chest50_like(256, 256, 50)
    geo = G.ConeGeometry(data)
    rays_all = G.rays_with_near_far(data["angles"][:4], geo, dev).reshape(-1, 8)
    N, S = 65536, 384
    gen = torch.Generator(device=dev).manual_seed(1)
    times = []
    for it in range(3):
        rays = rays_all[torch.randint(0, rays_all.shape[0], (N,), device=dev, generator=gen)]
        projs = torch.rand(N, device=dev, generator=gen) * 0.05
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad()
        loss = {"loss": 0.0}
        ret = render(rays, net, None, n_samples=S, n_fine=0, perturb=True, netchunk=409600, raw_noise_std=0.0)
        calc_mse_loss(loss, projs, ret["acc"])
        loss["loss"].backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        del ret, loss
    ms = float(np.median(times[1:]))
    print(json.dumps({"impl": "reference_cuda", "variant": "large_batch", "ms_per_step": ms, "value": N * S / (ms * 1e-3), "unit": "samples/s",
                      "workload": "65536 rays x 384 samples, one render() call per step"}), flush=True)
    opt.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    # ---- 512^3 voxel query
    n = 512
    s = (n * 0.001) / 2 - 0.001 / 2
    lin = torch.from_numpy(np.linspace(-s, s, n)).to(dev)
    vox = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).to(torch.float32)          # [n,n,n,3], 1.6 GB (tigre.py:388-400)
    times = []
    with torch.no_grad():
        for it in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            img = run_network(vox, net, 409600)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            del img
    print(json.dumps({"impl": "reference_cuda", "variant": "voxel_query_512", "ms": float(times[-1]), "value": n ** 3 / (times[-1] * 1e-3),
                      "unit": "voxels/s", "workload": "run_network over the 512^3 lattice, netchunk 409600"}), flush=True)


if __name__ == "__main__":
    main()
    if os.environ.get("REF_EXTRA") == "1":
        extra()
