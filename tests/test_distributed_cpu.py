"""Multi-process host logic of the ray-sharded data-parallel path, world_size 2 on the gloo backend (CPU).

What is checked:
  * shard_range: disjoint, ordered, exact cover (rays across ranks, voxel slabs across ranks);
  * the exchange step: per-rank gradients of the per-rank losses, summed with allreduce_sum_ (what NAFEngine hands to Adam,
    unscaled), equal the single-process gradient of the reference's chunked loss (train.py:69-127) on the CONCATENATED batch --
    computed with the CPU oracle network on real NAF rays, masks included;
  * broadcast_ / replica_divergence / combined_loss / gather_shards (the sharded evaluation render, train.py:235-240).
The GPU kernels are not involved (tests/test_gpu_parity.py covers them); this is the N > 1 plumbing.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_rays
from neuralvolumetricreconstructionformedicalimages_b200 import parallel
from oracle import hashgrid as oh
from oracle import naf

WORLD = 2


def test_shard_range_properties():
    for n in [0, 1, 2, 7, 128, 1024, 1025, 70]:
        for world in [1, 2, 3, 4, 8]:
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)
    assert parallel.world_info() == (0, 1)
    t = torch.ones(3)
    assert parallel.allreduce_sum_(t) is t and parallel.replica_divergence(t) == 0.0   # single process: no-ops


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_net(seed=0):
    torch.manual_seed(seed)
    enc = oh.OracleHashEncoder(3, 4, 2, 4, 10, use_ref=False, normalise="mul_recip")      # tiny grid: 4 levels x 2, 2^10
    net = naf.OracleDensityNetwork(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    with torch.no_grad():
        enc.embeddings.uniform_(-0.3, 0.3)
    return net


def _flat(ts):
    return torch.cat([t.reshape(-1) for t in ts])


def _rank_batch(rank, n_rays=48, S=16):
    rng = np.random.default_rng(100 + rank)
    rays = torch.from_numpy(make_rays(n_rays, rng))
    projs = torch.from_numpy(rng.uniform(0, 0.05, n_rays).astype(np.float32))
    mask = torch.from_numpy(rng.uniform(0, 1, n_rays) > 0.1)
    t_rand = torch.from_numpy(rng.uniform(0, 1, (n_rays, S)).astype(np.float32))
    return rays, projs, mask, t_rand, S


CHUNK = 16            # divides the per-rank batch (48 rays), so the chunks of the concatenated batch are the ranks' chunks


def _loss(net, batch, chunk=CHUNK):
    rays, projs, mask, t_rand, S = batch
    ret = naf.render(rays, net, S, True, t_rand=t_rand)
    return naf.chunked_masked_mse(ret["acc"], projs, mask, chunk)


def _worker(rank, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(1)
        assert parallel.world_info() == (rank, WORLD)
        # replicas start from rank 0's parameters
        net = _small_net(seed=rank)                      # deliberately different initialisation per rank
        flat_p = _flat([p.detach() for p in net.parameters()]).clone()
        assert parallel.replica_divergence(flat_p) > 0.0
        parallel.broadcast_(flat_p, 0)
        assert parallel.replica_divergence(flat_p) == 0.0
        off = 0
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(flat_p[off:off + p.numel()].view_as(p))
                off += p.numel()
        # local gradient of the local loss, then the exchange step
        loss = _loss(net, _rank_batch(rank))
        loss.backward()
        flat_g = _flat([p.grad for p in net.parameters()]).clone()
        parallel.allreduce_sum_(flat_g)                  # handed to Adam unscaled (grad_scale 1)
        mean_loss = parallel.combined_loss(loss)
        # sharded evaluation render: every rank holds its shard of a 37-element "acc", all ranks get the whole
        j0, j1 = parallel.shard_range(37, rank, WORLD)
        whole = parallel.gather_shards(torch.arange(j0, j1, dtype=torch.float32) * 2.0, 37)
        assert torch.equal(whole, torch.arange(37, dtype=torch.float32) * 2.0)
        whole2 = parallel.gather_shards(torch.arange(j0, j1)[:, None].repeat(1, 3), 37)
        assert whole2.shape == (37, 3) and torch.equal(whole2[:, 1], torch.arange(37))
        # voxel slabs: every rank contributes its slab, the union is the whole lattice
        n1 = 7
        i0, i1 = parallel.shard_range(n1, rank, WORLD)
        owned = torch.zeros(n1)
        owned[i0:i1] = 1
        dist.all_reduce(owned)
        assert torch.equal(owned, torch.ones(n1))
        if rank == 0:
            torch.save({"grad": flat_g, "loss": mean_loss, "param": flat_p}, os.path.join(out_dir, "r0.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_exchange_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    got = torch.load(os.path.join(str(tmp_path), "r0.pt"))
    # single process: the reference's chunked loss on the concatenated batch, same (rank-0) parameters
    net = _small_net(seed=0)
    np.testing.assert_array_equal(got["param"].numpy(), _flat([p.detach() for p in net.parameters()]).numpy())
    batches = [_rank_batch(r) for r in range(WORLD)]
    concat = tuple(torch.cat([b[k] for b in batches], 0) for k in range(4)) + (batches[0][4],)
    total = _loss(net, concat)
    np.testing.assert_allclose(total.item(), sum(_loss(net, b).item() for b in batches), rtol=1e-6)
    total.backward()
    ref = _flat([p.grad for p in net.parameters()])
    np.testing.assert_allclose(got["loss"].item(), total.item(), rtol=1e-6)
    np.testing.assert_allclose(got["grad"].numpy(), ref.numpy(), rtol=1e-5, atol=1e-7 * float(ref.abs().max()))
