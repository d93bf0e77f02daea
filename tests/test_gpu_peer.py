"""Peer spectrum (fftconv_peer_*, fftconv_b200.sharding.PeerSpectrum): two processes share one GPU here (CUDA IPC works
between processes on the same device as well), a gloo group carries the handle exchange, and the device-side flag
protocol (signal / wait / pull / acknowledge) is exercised for several steps, including the owner's wait for the
acknowledgements before it refills the buffer."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    import fftconv_b200 as fc
    from fftconv_b200.sharding import PeerSpectrum
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        rng = np.random.default_rng(7)
        H, W, F, kh, kw = 40, 30, 3, 8, 8
        ps = PeerSpectrum((F, 48, 25))
        ok, used_ipc = True, ps.enabled
        for step in range(4):
            data = rng.random((H, W, F), dtype=np.float32) + step           # same stream of inputs on both ranks
            d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()
            ps.begin_fill()
            if rank == 0:
                fc.fft_data_device(d_t, H, W, F, kh, kw, spec_t=ps.spec)
            spec = ps.publish_and_fetch()
            want = fc.fft_data_device(d_t, H, W, F, kh, kw)                  # every rank can check locally
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(torch.view_as_real(spec), torch.view_as_real(want)))
        ok = ok and ps.status() == 0
        ps.close()
        q.put((rank, bool(ok), bool(used_ipc)))
    finally:
        dist.destroy_process_group()


def test_peer_spectrum_two_processes(fc):
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
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]                 # both ranks agree on IPC vs fallback


def _ag_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    import fftconv_b200 as fc
    from fftconv_b200.sharding import PeerAllGatherSpectrum
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        rng = np.random.default_rng(9)
        H, W, F, kh, kw = 40, 30, 5, 8, 8                                     # 5 channels over 3 ranks: slices 1, 2, 2
        ag = PeerAllGatherSpectrum((F, 48, 25))
        ok, used_ipc = True, ag.enabled
        for step in range(4):
            data = rng.random((H, W, F), dtype=np.float32) + step           # the image is replicated: same stream of inputs
            d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()
            ag.begin_fill()
            ag.fill(d_t, H, W, kh, kw)
            spec = ag.gather()
            want = fc.fft_data_device(d_t, H, W, F, kh, kw)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(torch.view_as_real(spec), torch.view_as_real(want)))
        ok = ok and ag.status() == 0
        ag.close()
        q.put((rank, bool(ok), bool(used_ipc)))
    finally:
        dist.destroy_process_group()


def test_peer_allgather_three_processes(fc):
    """fftconv_peer_allgather: every rank transforms its channel slice, one kernel per rank pulls the others' slices."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ag_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert len({r[2] for r in res}) == 1


def _bc_worker(rank, world, port, q, scheme):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
    import torch
    import torch.distributed as dist
    import fftconv_b200 as fc
    from fftconv_b200.sharding import PeerBroadcastRaw
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        rng = np.random.default_rng(11)
        H, W, F, kh, kw, K = 40, 30, 3, 8, 8, 70
        bank = (rng.standard_normal((K, F, kw, kh)) * 0.1).astype(np.float32)
        b_t = torch.from_numpy(bank).cuda()
        bc = PeerBroadcastRaw(4 * F * W * H, scheme=scheme)
        ok, used_ipc = True, bc.enabled
        for step in range(4):
            data = rng.random((F, W, H), dtype=np.float32) + step            # every rank draws the same stream; only rank 0 uses it
            d_t = torch.from_numpy(data).cuda()
            bc.begin()
            if rank == 0:
                bc.publish(d_t)
            img = bc.fetch().view(torch.float32).view(F, W, H)
            out = fc.convolution_fft_device(img, b_t)                       # one-shot call on the delivered image
            want = fc.conv_bank(fc.fft_data_device(d_t, H, W, F, kh, kw), b_t, kh, kw)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(img, d_t))
            ok = ok and float((out - want).norm() / want.norm()) < 1e-5
        ok = ok and bc.status() == 0
        bc.close()
        q.put((rank, bool(ok), bool(used_ipc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scheme", ["pull", "scatter"])
def test_peer_broadcast_raw_three_processes(fc, scheme):
    """sharding.PeerBroadcastRaw: the raw image of rank 0 reaches every rank by scatter + all-gather (device flags), and the
    one-shot device-resident entry point convolves it."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bc_worker, args=(r, 3, port, q, scheme)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert len({r[2] for r in res}) == 1
