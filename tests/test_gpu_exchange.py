"""The fused exchange kernel (reduce-scatter + Adam + all-gather over NVLink peer memory, csrc/exchange.cu).

  * one GPU (world 1, flags and "peers" are the local buffers): bit-exact against the dense Adam kernel, which is itself
    pinned against torch.optim.Adam (test_gpu_parity.py::test_adam_matches_torch); gradient double buffering; flags;
  * NAFEngine(exchange="peer") on one GPU trains like the default engine;
  * two GPUs (skipped when the box has one): two ranks over NCCL + CUDA IPC; the peer path keeps the replicas
    bit-identical and agrees with the NCCL all-reduce path; a checkpoint rollback re-arms the epoch flags; two ranks equal ONE
    GPU working on the concatenated batch (the data-parallel loss rule); an injected peer_open failure on one rank makes
    every rank fall back to NCCL together.
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

from helpers import make_rays

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from neuralvolumetricreconstructionformedicalimages_b200 import _lib
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network

DEV = "cuda"


def test_exchange_kernel_world1_bit_exact_vs_adam():
    L_ = _lib.lib()
    n = 4 * 25_003          # not a multiple of the block size
    g = torch.Generator(device=DEV).manual_seed(5)
    buf = _lib.peer_alloc(256 + 3 * n * 4)
    flags = buf.tensor(0, _lib.XFLAG_WORDS, torch.int32, DEV)
    param = buf.tensor(256, n, torch.float32, DEV)
    grad0 = buf.tensor(256 + n * 4, n, torch.float32, DEV)
    grad1 = buf.tensor(256 + 2 * n * 4, n, torch.float32, DEV)
    param.copy_(torch.randn(n, device=DEV, generator=g))
    grad0.copy_(torch.randn(n, device=DEV, generator=g) * 1e-2)
    grad1.fill_(7.0)                                        # stale other-parity buffer: must come back zeroed
    m = torch.randn(n, device=DEV, generator=g) * 1e-3
    v = torch.rand(n, device=DEV, generator=g) * 1e-4
    ref = [param.clone(), grad0.clone(), m.clone(), v.clone()]
    x = _lib.Exchange()
    x.world, x.rank, x.n = 1, 0, n
    x.flags[0], x.param[0], x.grad[0] = buf.ptr, buf.ptr + 256, buf.ptr + 256 + n * 4
    x.grad_zero = buf.ptr + 256 + 2 * n * 4
    x.exp_avg, x.exp_avg_sq = m.data_ptr(), v.data_ptr()
    for step in (1, 2, 9):
        _lib.check(L_.nafb_adam_exchange_step(ctypes.byref(x), 1e-3, 0.9, 0.999, 1e-8, step, 1.0, _lib.stream_ptr()))
        _lib.check(L_.nafb_adam_step(_lib.ptr(ref[0]), _lib.ptr(ref[1]), _lib.ptr(ref[2]), _lib.ptr(ref[3]), n, 1e-3, 0.9, 0.999, 1e-8, step,
                                     1.0, 0, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(param, ref[0]) and torch.equal(m, ref[2]) and torch.equal(v, ref[3])
        assert torch.equal(grad0, ref[1])                   # this step's gradient is left alone (its owner clears it next step)
        assert int(grad1.abs().max().item()) == 0
        f = flags.cpu().numpy()
        assert f[_lib.XFLAG_ARRIVE] == step and f[_lib.XFLAG_DONE] == step and f[_lib.XFLAG_ERROR] == 0 and f[_lib.XFLAG_TICKET] == 0
    # ---- push edition, world 1: no peers to push to, the owner phase alone must reproduce the dense kernel
    stage = torch.zeros(n, device=DEV)
    x.stage[0], x.stage_slot, x.grad_zero = stage.data_ptr(), n, None
    for step in (10, 11):
        grad0.copy_(torch.randn(n, device=DEV, generator=g) * 1e-2)
        ref[1].copy_(grad0)
        _lib.check(L_.nafb_adam_exchange_step(ctypes.byref(x), 1e-3, 0.9, 0.999, 1e-8, step, 1.0, _lib.stream_ptr()))
        _lib.check(L_.nafb_adam_step(_lib.ptr(ref[0]), _lib.ptr(ref[1]), _lib.ptr(ref[2]), _lib.ptr(ref[3]), n, 1e-3, 0.9, 0.999, 1e-8, step,
                                     1.0, 1, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(param, ref[0]) and torch.equal(m, ref[2]) and torch.equal(v, ref[3])
        assert int(grad0.abs().max().item()) == 0           # the push edition's single gradient buffer leaves the kernel zeroed
        f = flags.cpu().numpy()
        assert f[_lib.XFLAG_DONE] == step and f[_lib.XFLAG_ERROR] == 0 and f[_lib.XFLAG_TICKET] == 0 and f[_lib.XFLAG_TICKET2] == 0
    x.stage_slot = n // 2 // 4 * 4
    with pytest.raises(RuntimeError):                       # staging slot smaller than the slice
        _lib.check(L_.nafb_adam_exchange_step(ctypes.byref(x), 1e-3, 0.9, 0.999, 1e-8, 12, 1.0, _lib.stream_ptr()))
    i0, i1 = _lib.exchange_slice(n, 0, 1)
    assert (i0, i1) == (0, n)
    spans = [_lib.exchange_slice(n, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] and a[0] % 4 == 0 for a, b in zip(spans, spans[1:]))
    with pytest.raises(RuntimeError):
        _lib.check(L_.nafb_adam_exchange_step(ctypes.byref(x), 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, _lib.stream_ptr()))
    del flags, param, grad0, grad1
    buf.release()


def _net(dev, seed=0):
    torch.manual_seed(seed)
    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(dev)
    with torch.no_grad():
        enc.embeddings.uniform_(-0.05, 0.05)
    return net


def _batches(rank, k, N=256, S=64):
    rng = np.random.default_rng(10 + rank)
    out = []
    for _ in range(k):
        rays = torch.from_numpy(make_rays(N, rng))
        projs = torch.from_numpy(rng.uniform(0, 0.05, N).astype(np.float32))
        t_rand = torch.from_numpy(rng.uniform(0, 1, (N, S)).astype(np.float32))
        out.append((rays, projs, t_rand))
    return out, S


def test_engine_peer_mode_single_gpu_matches_default():
    dev = torch.device("cuda", 0)
    batches, S = _batches(0, 4)
    params = []
    for mode in ("peer", "push", "auto"):
        eng = NAFEngine(_net(dev), lr=1e-3, n_samples=S, perturb=True, loss_chunk=100, use_cuda_graph=(mode != "auto"), exchange=mode)
        assert eng.exchange_mode == (mode if mode != "auto" else "local")
        for rays, projs, t_rand in batches:
            loss = eng.train_step(rays.to(dev), projs.to(dev), None, t_rand.to(dev))
        torch.cuda.synchronize()
        if eng.px is not None:
            assert eng.px.error_word() == 0
        step, m_flat, v_flat = eng.optimizer_moments()
        assert m_flat.numel() == eng.n_params and step == 4
        sd = eng.optimizer_state_dict()                      # torch.optim.Adam layout (what the reference's ckpt.tar holds)
        assert float(sd["state"][0]["step"]) == 4.0 and sd["state"][0]["exp_avg_sq"].shape == (7131219, 2)
        eng.check_health()
        params.append((eng.flat_param.clone(), v_flat, float(loss.item())))
    pb, vb, lb = params[-1]
    for pa, va, la in params[:-1]:
        assert abs(la - lb) <= 1e-5 * abs(lb)
        np.testing.assert_allclose(pa.cpu().numpy(), pb.cpu().numpy(), rtol=0, atol=5e-5)     # 4 steps of lr 1e-3; float atomics order differs (see _assert_same_parameters in test_gpu_parity.py)
        np.testing.assert_allclose(va.cpu().numpy(), vb.cpu().numpy(), rtol=1e-3, atol=1e-12)


# ----------------------------------------------------------------------------- two ranks
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from neuralvolumetricreconstructionformedicalimages_b200 import parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        batches, S = _batches(rank, 5)
        res = {}
        modes = ["push", "peer", "nccl"]
        try:   # NVLS needs NVSwitch multicast; exercised when the box has it
            probe = NAFEngine(_net(dev, seed=rank), lr=1e-3, n_samples=S, exchange="nvls")
            del probe
            modes.insert(0, "nvls")
        except RuntimeError:
            pass
        for mode in modes:
            eng = NAFEngine(_net(dev, seed=rank), lr=1e-3, n_samples=S, perturb=True, loss_chunk=128, use_cuda_graph=True, exchange=mode)
            assert eng.exchange_mode == mode, eng.exchange_mode
            for rays, projs, t_rand in batches:
                eng.train_step(rays.to(dev), projs.to(dev), None, t_rand.to(dev))
            torch.cuda.synchronize()
            div = parallel.replica_divergence(eng.flat_param)
            err = eng.px.error_word() if eng.px is not None else 0
            eng.check_health()
            _, _, v_flat = eng.optimizer_moments()
            res[mode] = dict(param=eng.flat_param.cpu(), v=v_flat.cpu(), div=div, err=err)
            # a rollback (checkpoint restore to an EARLIER step) must re-arm the epoch flags: two more steps stay in lock-step
            sd = eng.optimizer_state_dict()
            for _ in range(2):
                rays, projs, t_rand = batches[0]
                eng.train_step(rays.to(dev), projs.to(dev), None, t_rand.to(dev))
            eng.load_optimizer_state_dict(sd)
            for rays, projs, t_rand in batches[:2]:
                eng.train_step(rays.to(dev), projs.to(dev), None, t_rand.to(dev))
            torch.cuda.synchronize()
            res[mode]["div_after_rollback"] = parallel.replica_divergence(eng.flat_param)
            res[mode]["err_after_rollback"] = eng.px.error_word() if eng.px is not None else 0
            del eng
        # ---- the data-parallel rule: W ranks == ONE GPU on the concatenated batch with the same chunk boundaries (the loss is the
        # reference's sum of chunk means, train.py:69-127; the rank-summed gradient goes to Adam unscaled)
        solo_group = [dist.new_group([r]) for r in range(world)][rank]
        solo = NAFEngine(_net(dev, seed=0), lr=1e-3, n_samples=S, perturb=True, loss_chunk=128, use_cuda_graph=False, process_group=solo_group)
        assert solo.world_size == 1
        all_batches = [_batches(r, 5)[0] for r in range(world)]
        for k in range(5):
            cat = [torch.cat([all_batches[r][k][j] for r in range(world)], 0).to(dev) for j in range(3)]
            solo.train_step(cat[0], cat[1], None, cat[2])
        torch.cuda.synchronize()
        res["solo"] = dict(param=solo.flat_param.cpu())
        # ---- a peer_open failure on ONE rank: every rank must leave the set-up together and agree on the NCCL fallback
        os.environ["NAFB_TEST_FAIL_PEER_OPEN"] = "1"
        eng = NAFEngine(_net(dev, seed=rank), lr=1e-3, n_samples=S, perturb=True, loss_chunk=128, exchange="auto")
        del os.environ["NAFB_TEST_FAIL_PEER_OPEN"]
        assert eng.exchange_mode == "nccl", eng.exchange_mode
        rays, projs, t_rand = batches[0]
        eng.train_step(rays.to(dev), projs.to(dev), None, t_rand.to(dev))
        torch.cuda.synchronize()
        res["fallback_div"] = parallel.replica_divergence(eng.flat_param)
        if rank == 0:
            torch.save(res, os.path.join(out_dir, "res.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_peer_exchange_matches_nccl(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(str(tmp_path), "res.pt"), weights_only=False)
    print("exchange modes exercised:", sorted(res))
    modes = [m for m in res if isinstance(res[m], dict) and "div" in res[m]]
    for mode in modes:
        assert res[mode]["err"] == 0 and res[mode]["div"] == 0.0, mode    # no timed-out flag, replicas bit-identical
        assert res[mode]["err_after_rollback"] == 0 and res[mode]["div_after_rollback"] == 0.0, mode
        if mode != "nccl":
            np.testing.assert_allclose(res[mode]["param"].numpy(), res["nccl"]["param"].numpy(), rtol=0, atol=5e-5)
            np.testing.assert_allclose(res[mode]["v"].numpy(), res["nccl"]["v"].numpy(), rtol=2e-3, atol=1e-12)
    assert res["fallback_div"] == 0.0
    # two ranks == one GPU on the concatenated batch (replicas start from rank 0's parameters in both runs)
    a, b = res["nccl"]["param"].numpy(), res["solo"]["param"].numpy()
    assert (np.abs(a - b) > 5e-5).mean() < 1e-4 and np.abs(a - b).max() < 2.1e-3 * 5   # 5 steps of lr 1e-3; float-atomic order differs


def test_backward_grid_barrier_survives_a_busy_gpu():
    """The tensor-core backward kernel reduces the per-CTA MLP gradients behind a grid-wide barrier of its own (cooperative launch
    is not available to kernels that allocate tensor memory: the occupancy API answers 1 CTA per SM for them).  Its CTAs must all
    become resident, so a long kernel on ANOTHER stream may delay part of the grid while the rest spins: the launch has to finish
    (no trap, no time-out flag) with the same result as on a quiet GPU, and two engines stepping on two streams at once -- each
    with its own workspace -- must not disturb each other."""
    dev = torch.device("cuda", 0)
    batches, S = _batches(3, 3, N=1024, S=192)

    def run(eng, busy):
        side = torch.cuda.Stream(device=dev)
        a = torch.randn(8192, 8192, device=dev)
        out = []
        for rays, projs, t_rand in batches:
            if busy:
                with torch.cuda.stream(side):          # ~10 ms of work that fills every SM while the step is enqueued
                    for _ in range(12):
                        a = (a @ a).clamp_(-1, 1)
            out.append(float(eng.train_step(rays.to(dev), projs.to(dev), None, t_rand.to(dev))))
        torch.cuda.synchronize()
        eng.check_health()
        return out, eng.flat_param.clone()

    quiet, p_quiet = run(NAFEngine(_net(dev), lr=1e-3, n_samples=S, perturb=True, loss_chunk=200, use_cuda_graph=False), busy=False)
    loud, p_loud = run(NAFEngine(_net(dev), lr=1e-3, n_samples=S, perturb=True, loss_chunk=200, use_cuda_graph=False), busy=True)
    np.testing.assert_allclose(loud, quiet, rtol=1e-5)
    assert float((p_loud - p_quiet).abs().max()) < 5e-5      # float-atomic order differs from run to run, nothing else
    # two engines, two streams, at the same time
    e1 = NAFEngine(_net(dev), lr=1e-3, n_samples=S, perturb=True, loss_chunk=200, use_cuda_graph=False)
    e2 = NAFEngine(_net(dev), lr=1e-3, n_samples=S, perturb=True, loss_chunk=200, use_cuda_graph=False)
    assert e1._bwd_ws.data_ptr() != e2._bwd_ws.data_ptr()
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dev_batches = [(r.to(dev), p.to(dev), t.to(dev)) for r, p, t in batches]
    torch.cuda.synchronize()
    l1, l2 = [], []
    for rays, projs, t_rand in dev_batches:
        with torch.cuda.stream(s1):
            l1.append(e1.train_step(rays, projs, None, t_rand).clone())       # (the returned scalar is a view of a buffer the next step overwrites)
        with torch.cuda.stream(s2):
            l2.append(e2.train_step(rays, projs, None, t_rand).clone())
    torch.cuda.synchronize()
    e1.check_health()
    e2.check_health()
    np.testing.assert_allclose([float(v) for v in l1], quiet, rtol=1e-5)
    np.testing.assert_allclose([float(v) for v in l2], quiet, rtol=1e-5)
