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
