#!/usr/bin/env python
"""End-to-end demonstration: reconstruct an analytic phantom from its cone-beam projections with the fused engine.

Follows config/chest_50.yaml of the reference (hash grid 16x2 / 2^19, 4x32 MLP with skip at 2, sigmoid head, bound 0.3,
1024 rays x 192 samples per iteration, Adam lr 1e-3, StepLR(1500 epochs, 0.1), 50 projections = 50 iterations per epoch)
on a synthetic stand-in for the missing pickle: 50 exact projections (256 x 256) of a sum of ellipsoids in a 128^3 volume.
Everything per-iteration stays on the GPU: pixel selection, ray generation, sampling, encode, MLP, integral, loss,
backward, Adam (one CUDA graph per iteration).  Evaluation: the 128^3 voxel query + PSNR-3D / SSIM-3D on the device.

    python examples/train_phantom.py --epochs 100
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_phantom.py --epochs 100
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G            # noqa: E402
from neuralvolumetricreconstructionformedicalimages_b200.dataset import phantom as PH            # noqa: E402
from neuralvolumetricreconstructionformedicalimages_b200.dataset.mask import PixelSampler       # noqa: E402
from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder              # noqa: E402
from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine                 # noqa: E402
from neuralvolumetricreconstructionformedicalimages_b200.network import get_network              # noqa: E402
from neuralvolumetricreconstructionformedicalimages_b200.utils import get_psnr_3d, get_ssim_3d   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100, help="passes over the 50 projections (the reference trains 1501)")
    ap.add_argument("--n-rays", type=int, default=1024)
    ap.add_argument("--n-samples", type=int, default=192)
    ap.add_argument("--lrate-step", type=int, default=1500)
    ap.add_argument("--eval-every", type=int, default=0)
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)

    data = G.chest50_like(n_voxel=128, n_detector=256, n_proj=50)
    geo = G.ConeGeometry(data)
    ells = PH.default_ellipsoids(float(geo.sVoxel[0]) / 2 * 0.9)
    vol_gt = torch.from_numpy(PH.phantom_volume(geo, ells)).to(dev)
    rays_all = G.rays_with_near_far(data["angles"], geo, "cpu")
    projs = PH.phantom_projections(rays_all, ells).to(dev)                      # [50, 256, 256] exact line integrals
    del rays_all
    sampler = PixelSampler(projs, seed=1234 + rank)

    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(dev)
    eng = NAFEngine(net, lr=1e-3, n_samples=args.n_samples, perturb=True, loss_chunk=200)
    eng.set_geometry(data["angles"], geo)

    def evaluate():
        vol = eng.voxel_query([int(v) for v in geo.nVoxel], G.voxel_half_extent(geo))
        return get_psnr_3d(vol, vol_gt), get_ssim_3d(vol, vol_gt)

    n_proj = projs.shape[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it = 0
    for epoch in range(args.epochs):
        eng.lr = 1e-3 * (0.1 ** (epoch // args.lrate_step))                    # StepLR(step_size in epochs, gamma 0.1), trainer.py:57
        # one pass over the projections in the order of the reference's un-shuffled DataLoader: every iteration is ONE graph launch --
        # pixel draw on the device (csrc/select.cu), forward + loss, backward, Adam -- with no per-step host input at all
        for k in range(n_proj):
            loss = eng.train_step_sampled(sampler, args.n_rays)
            it += 1
        if args.eval_every and (epoch + 1) % args.eval_every == 0 and rank == 0:
            psnr, ssim = evaluate()
            print(f"epoch {epoch + 1:5d}  loss {loss.item():.3e}  psnr_3d {psnr:.2f} dB  ssim_3d {ssim:.4f}", flush=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    psnr, ssim = evaluate()
    if rank == 0:
        n = it * args.n_rays * args.n_samples * world
        print(f"{args.epochs} epochs = {it} iterations on {world} GPU(s) in {dt:.2f} s ({n / dt / 1e6:.0f} M samples/s incl. on-device pixel "
              f"selection); final loss {loss.item():.3e}; PSNR-3D {psnr:.2f} dB, SSIM-3D {ssim:.4f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
