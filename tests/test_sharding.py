"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group.  The per-rank compute is injected
(the oracle stands in for the CUDA path, which needs a GPU): what is tested is the shard arithmetic,
the spectrum broadcast and the gather — the N > 1 plumbing bench.py --gpus N uses."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bank_partitions(fc):
    from fftconv_b200.sharding import shard_bank
    for K in (0, 1, 3, 10, 1000, 20000):
        for world in (1, 2, 4, 8):
            parts = shard_bank([256.0] * K, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == K
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    rng = np.random.default_rng(0)
    costs = (rng.integers(6, 17, 500) * rng.integers(6, 17, 500)).astype(float)
    parts = shard_bank(list(costs), 8)
    loads = [costs[b:e].sum() for b, e in parts]
    assert max(loads) / (costs.sum() / 8) < 1.06


def _shard_bank_loop(costs, world):
    """the greedy walk the vectorised shard_bank restates: rank r ends at the first template whose cost midpoint lies
    beyond r / world of the total"""
    K = len(costs)
    total = float(sum(costs))
    bounds, acc, k = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while k < K and acc + costs[k] / 2.0 <= target:
            acc += costs[k]
            k += 1
        bounds.append(k)
    bounds.append(K)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def test_shard_bank_vectorised_equals_the_greedy_walk(fc):
    """shard_bank runs inside the multi-GPU step (cumulative sum + searchsorted since round 2e: the Python walk over 20 000
    templates cost 1 ms of the 6.3 ms config-5 step on 8 GPUs); same ranges as the walk, uniform shortcut included."""
    import time
    from fftconv_b200.sharding import shard_bank
    rng = np.random.default_rng(7)
    for trial in range(200):
        K, world = int(rng.integers(0, 80)), int(rng.integers(1, 10))
        costs = [float(a * b) for a, b in zip(rng.integers(1, 33, K), rng.integers(1, 33, K))] if trial % 3 else [1.0] * K
        assert shard_bank(costs, world) == _shard_bank_loop(costs, world), (costs, world)
    for K in (1000, 20000, 20001):
        for world in (1, 2, 3, 4, 8):
            assert shard_bank(None, world, K) == shard_bank([1.0] * K, world) == _shard_bank_loop([1.0] * K, world)
    t0 = time.perf_counter()
    for _ in range(20):
        shard_bank(None, 8, 20000)
    assert (time.perf_counter() - t0) / 20 < 2e-3                   # generous: ~0.1 ms here, the walk took 2.2 ms
    with pytest.raises(ValueError):
        shard_bank([1.0], 0)


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    import oracle
    from fftconv_b200.sharding import sharded_convolution
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)                      # same inputs on every rank
        data = rng.random((40, 30, 3), dtype=np.float32)
        kernels = [rng.standard_normal((int(rng.integers(3, 9)), int(rng.integers(3, 9)), 3)).astype(np.float32)
                   for _ in range(7)]
        shape = oracle.fft_data(data, 8, 8).shape

        def fft_fn(d, kh, kw):
            return torch.from_numpy(oracle.fft_data(d, kh, kw))

        def conv_fn(spec, shard):
            return oracle.conv_fft_data(spec.numpy(), shard)

        b, e, planes = sharded_convolution(data if rank == 0 else None, 8, 8, kernels, fft_fn, conv_fn,
                                           lambda: torch.zeros(shape, dtype=torch.complex64))
        full = sharded_convolution(data if rank == 0 else None, 8, 8, kernels, fft_fn, conv_fn,
                                   lambda: torch.zeros(shape, dtype=torch.complex64), gather=True)
        ref = oracle.convolution_fft(data, 8, 8, kernels)
        ok = len(full) == len(ref) and all(np.array_equal(a, r) for a, r in zip(full, ref))
        ok = ok and all(np.array_equal(p, ref[b + i]) for i, p in enumerate(planes))
        # the side-stream broadcast helper degrades to an in-order broadcast on CPU tensors and returns no event
        from fftconv_b200.sharding import broadcast_spectrum_async, bind_host_to_gpu
        sp = fft_fn(data, 8, 8) if rank == 0 else torch.zeros(shape, dtype=torch.complex64)
        ok = ok and broadcast_spectrum_async(sp, 0) is None
        ok = ok and np.array_equal(sp.numpy(), oracle.fft_data(data, 8, 8))
        ok = ok and bind_host_to_gpu(0) is None            # no GPU / NVML here: best effort, must not raise
        q.put((rank, b, e, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_convolution_gloo_world2(fc):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 7
    assert all(r[3] for r in res)


def _pyr_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    import oracle
    from fftconv_b200.pyramid import pyramid_convolution, pyramid_sides
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)                      # same inputs on every rank
        sides = pyramid_sides(40, 4, 2)
        assert sides == [40, 28, 20, 14]
        levels = [rng.random((s, s, 3), dtype=np.float32) for s in sides]
        kernels = [rng.standard_normal((int(rng.integers(3, 7)), int(rng.integers(3, 7)), 3)).astype(np.float32)
                   for _ in range(5)]
        shapes = [(s, s, 3) for s in sides]

        def fft_fn(lv, kh, kw):
            return torch.from_numpy(oracle.fft_data(lv, kh, kw))

        def alloc(H, W, F):
            return torch.zeros(oracle.fft_data(np.zeros((H, W, F), np.float32), 6, 6).shape, dtype=torch.complex64)

        def conv_fn(l, spec, b, e):
            return oracle.conv_fft_data(spec.numpy(), kernels[b:e])

        b, e, res = pyramid_convolution(levels if rank == 0 else None, 6, 6, len(kernels),
                                        [float(k.shape[0] * k.shape[1]) for k in kernels], fft_fn, alloc, conv_fn, shapes)
        ok = True
        for l, lv in enumerate(levels):
            ref = oracle.convolution_fft(lv, 6, 6, kernels)
            ok = ok and len(res[l]) == e - b and all(np.array_equal(p, ref[b + i]) for i, p in enumerate(res[l]))
        q.put((rank, b, e, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_pyramid_schedule_gloo_world2(fc):
    """config 5 plumbing: every level spectrum broadcast from rank 0, bank sharded, outputs stay sharded."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pyr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 5
    assert all(r[3] for r in res)


def test_pyramid_sides_config5(fc):
    from fftconv_b200.pyramid import pyramid_sides, level_plane
    sides = pyramid_sides()
    assert sides == [256, 223, 194, 169, 147, 128, 111, 97, 84, 74]
    assert [level_plane(s, s, 16, 16)[0] for s in sides] == [272, 240, 224, 192, 176, 144, 128, 112, 112, 96]


def test_channel_slices():
    from fftconv_b200.sharding import channel_slices
    for F in (1, 5, 31, 32):
        for world in (1, 2, 3, 8):
            b = channel_slices(F, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == F and all(b[i] <= b[i + 1] for i in range(world))
            sizes = [b[i + 1] - b[i] for i in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _ag_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    import oracle
    from fftconv_b200.sharding import PeerAllGatherSpectrum
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(4)
        data = rng.random((40, 30, 3), dtype=np.float32)              # replicated image
        full = oracle.fft_data(data, 8, 8)                            # [F][FW][CH]
        ag = PeerAllGatherSpectrum(full.shape)                        # no CUDA here: per-slice broadcast fallback
        d_fwh = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0)))

        def fft_fn(d_slice, nch, spec_slice):
            part = oracle.fft_data(np.ascontiguousarray(d_slice.numpy().transpose(2, 1, 0)), 8, 8)
            spec_slice.copy_(torch.from_numpy(part))

        ok = not ag.enabled
        for _ in range(2):
            ag.begin_fill()
            ag.fill(d_fwh, 40, 30, 8, 8, fft_fn=fft_fn)
            spec = ag.gather()
            ok = ok and np.array_equal(spec.numpy(), full)
        ag.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_allgather_spectrum_fallback_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ag_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def _bc_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    from fftconv_b200.sharding import PeerBroadcastRaw
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(6)
        bc = PeerBroadcastRaw(4 * 3 * 30 * 40)                        # no CUDA here: broadcast fallback through the process group
        ok = not bc.enabled
        for step in range(3):
            img = torch.from_numpy(rng.random((3, 30, 40), dtype=np.float32) + step)     # same stream of draws on both ranks
            bc.begin()
            if rank == 0:
                bc.publish(img)                                       # only rank 0 contributes its copy
            got = bc.fetch().view(torch.float32).view(3, 30, 40)
            ok = ok and bool(torch.equal(got, img))
        bc.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_broadcast_raw_fallback_world2():
    """sharding.PeerBroadcastRaw without CUDA IPC: the image of rank 0 reaches every rank through the process group."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
