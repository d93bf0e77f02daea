"""GPU parity at the EXACT BASELINE.json shapes (configs[1..4]), on the pipelines that bench.py times.

Each test runs the full-size workload through the C ABI and checks sampled output planes against the float64
oracle (direct convolution in C/OpenMP, or the float64 FFT convolution for the 512x512 templates of config 3
with a direct-convolution spot check).  The sampled planes are chosen so that every template block of the
per-bin GEMM (128 rows), every host-output chunk and both halves of the double-buffered output staging are hit.
Semantics: whole FH x FW plane, no flip (src/cudaConvFFTData.cu:191-282).  Tolerance rel-L2 <= 1e-5."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5

C2_PLANES = (0, 127, 128, 383, 384, 639, 640, 895, 896, 999)


def _c2_inputs():
    rng = np.random.default_rng(2)
    data = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
    bank = (rng.standard_normal((1000, 16, 16, 31)) * 0.05).astype(np.float32)
    return data, bank


def _profiled_kernels(fc, fn):
    fc.profile(True)
    try:
        fn()
        import torch
        torch.cuda.synchronize()
        return fc.profile_read()
    finally:
        fc.profile(False)


def test_c2_full_bank_device_outputs(fc, oracle):
    """config 2 as bench.py's `value` leg runs it: cudaFFTData spectrum -> device bank of 1000 templates -> 1000 device
    planes; overlap-save / tcgen05 pipeline, 8 template blocks (the last one holds 104 rows)."""
    import torch
    data, bank = _c2_inputs()
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    spec = fc.fft_data_device(d_t, 256, 256, 31, 16, 16)
    out = torch.full((1000, 272, 272), float("nan"), device="cuda")
    prof = _profiled_kernels(fc, lambda: fc.conv_bank(spec, b_t, 16, 16, out))
    assert {"os_data_fft(tiles)", "os_kern_fft(templates)", "os_gemm", "os_inverse"} <= set(prof), prof
    assert not bool(torch.isnan(out).any())
    for k in C2_PLANES:
        ref = oracle.direct_conv64_c(data, bank[k], 272, 272)
        assert oracle.rel_l2(out[k].cpu().numpy().T, ref) < TOL, k


def test_c2_ragged_bank_device_cells(fc, oracle):
    """config 2, secondary run of SURVEY 8(d): 1000 templates whose sizes differ per cell, kh, kw ~ U{6..16} iid
    (src/cudaConvFFTData.cu:204-231 reads the size of every cell), declared maximum 16 x 16, everything on the device,
    one-shot entry point on the overlap-save / tcgen05 pipeline."""
    import torch
    rng = np.random.default_rng(22)
    data = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
    khs, kws = rng.integers(6, 17, 1000), rng.integers(6, 17, 1000)
    khs[0], kws[0], khs[999], kws[999] = 6, 16, 16, 6
    cells = [(rng.standard_normal((int(a), int(b), 31)) * 0.05).astype(np.float32) for a, b in zip(khs, kws)]
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()
    c_t = [torch.from_numpy(np.ascontiguousarray(k.transpose(2, 1, 0))).cuda() for k in cells]
    dc = fc.DeviceCells(c_t, 31)
    assert (dc.max_kh, dc.max_kw) == (16, 16)
    out = torch.full((1000, 272, 272), float("nan"), device="cuda")
    prof = _profiled_kernels(fc, lambda: fc.convolution_fft_device(d_t, dc, out, max_kh=16, max_kw=16))
    assert {"os_data_fft(tiles)", "os_kern_fft(templates)", "os_gemm", "os_inverse"} <= set(prof), prof
    assert not bool(torch.isnan(out).any())
    for k in C2_PLANES:
        ref = oracle.direct_conv64_c(data, cells[k], 272, 272)
        assert oracle.rel_l2(out[k].cpu().numpy().T, ref) < TOL, (k, cells[k].shape)
    # the same cell through the reference-facing host entry point gives the same planes (bit for bit: same pipeline)
    outs = fc.cudaConvolutionFFT(data, 16, 16, [cells[k] for k in (0, 500, 999)] * 22)      # 66 templates: overlap-save path
    for i, k in enumerate((0, 500, 999)):
        assert oracle.rel_l2(outs[i], oracle.direct_conv64_c(data, cells[k], 272, 272)) < TOL


@pytest.mark.parametrize("entry", ["cudaConvolutionFFT", "cudaConvFFTData", "cudaConvFFTDataStreams"])
def test_c2_full_bank_host_outputs(fc, oracle, entry):
    """config 2 as bench.py's `e2e` leg runs it: host kernels in, 1000 pageable host planes out (the MEX contract,
    src/cudaConvFFTData.cu:275-279): >= 3 output chunks, both staging halves reused."""
    data, bank = _c2_inputs()
    cells = [bank[k] for k in range(1000)]
    if entry == "cudaConvolutionFFT":
        outs = fc.cudaConvolutionFFT(data, 16, 16, cells)
    else:
        spec = fc.cudaFFTData(data, 16, 16)
        outs = getattr(fc, entry)(spec, cells)
    assert len(outs) == 1000
    for k in C2_PLANES:
        ref = oracle.direct_conv64_c(data, bank[k], 272, 272)
        assert outs[k].shape == (272, 272)
        assert oracle.rel_l2(outs[k], ref) < TOL, (entry, k)


def test_c2_separate_pageable_planes_like_a_mex_cell(fc, oracle):
    """The MEX shim hands over K SEPARATE pageable mxArrays (mex/mex_common.h alloc_out_cell), not one slab:
    the library must stage them through its pinned bounce ring.  Raw C-ABI call, non-contiguous planes."""
    data, bank = _c2_inputs()
    K = 300
    d = np.ascontiguousarray(data.transpose(2, 1, 0))
    b = np.ascontiguousarray(bank[:K].transpose(0, 3, 2, 1))
    planes = [np.full((272, 272), np.nan, np.float32) for _ in range(K)]
    planes = planes[::-1]                                             # make adjacency unlikely
    kp = (ctypes.c_void_p * K)(*[b.ctypes.data + 4 * k * 31 * 256 for k in range(K)])
    op = (ctypes.c_void_p * K)(*[p.ctypes.data for p in planes])
    khs = (ctypes.c_int * K)(*([16] * K))
    rc = fc.lib().fftconv_convolution_fft(d.ctypes.data, 0, 256, 256, 31, 16, 16, K, kp, khs, khs, None, None, op, 0,
                                          None, 0, None, 0, None)
    assert rc == 0, fc.last_error()
    for k in (0, 127, 128, 255, 256, 299):
        ref = oracle.direct_conv64_c(data, bank[k], 272, 272)
        assert oracle.rel_l2(planes[k].T, ref) < TOL, k


def test_c3_full_size_large_plane(fc, oracle):
    """config 3 at full size: 4096 x 4096 single-channel image, 512 x 512 templates, 4608 x 4608 plane (radix plan
    [9, 32, 16]), large-plane pipeline.  Ground truth: float64 FFT convolution + a float64 direct
    convolution spot check on a window."""
    import scipy.fft
    import torch
    rng = np.random.default_rng(3)
    H = W = 4096; kh = kw = 512; K = 2
    data = rng.random((H, W), dtype=np.float32)
    ks = (rng.standard_normal((K, kh, kw)) / 512).astype(np.float32)
    rad_h = (ctypes.c_int * 8)()
    path = fc.lib().fftconv_query_path(H, W, 1, kh, kw, 64, None, rad_h, None)
    assert path == 4 and [r for r in rad_h if r] == [9, 32, 16]
    d_t = torch.from_numpy(np.ascontiguousarray(data.T))[None].cuda()          # [1][W][H]
    b_t = torch.from_numpy(np.ascontiguousarray(ks.transpose(0, 2, 1)))[:, None].cuda()   # [K][1][kw][kh]
    spec = fc.fft_data_device(d_t, H, W, 1, kh, kw)
    assert tuple(spec.shape) == (1, 4608, 2305)
    prof = _profiled_kernels(fc, lambda: fc.conv_bank(spec, b_t, kh, kw))
    assert "bp_conv_w" in prof and "bp_inv_h" in prof, prof
    out = fc.conv_bank(spec, b_t, kh, kw)
    torch.cuda.synchronize()
    D = scipy.fft.rfft2(data.astype(np.float64), s=(4608, 4608), workers=-1)
    for k in range(K):
        ref = scipy.fft.irfft2(D * scipy.fft.rfft2(ks[k].astype(np.float64), s=(4608, 4608), workers=-1), s=(4608, 4608), workers=-1)
        got = out[k].cpu().numpy().T
        assert oracle.rel_l2(got, ref) < TOL, k
        # direct float64 convolution on a 24 x 24 window in the interior
        y0, x0 = 2000 + 37 * k, 1500
        win = np.zeros((24, 24))
        kk = ks[k].astype(np.float64)[::-1, ::-1]
        for yy in range(24):
            for xx in range(24):
                y, x = y0 + yy, x0 + xx
                win[yy, xx] = (data[y - kh + 1:y + 1, x - kw + 1:x + 1].astype(np.float64) * kk).sum()
        assert oracle.rel_l2(got[y0:y0 + 24, x0:x0 + 24], win) < TOL, k


def test_c4_full_size_batch(fc, oracle):
    """config 4 at full size: 64 images 512 x 512 x 32 against 256 kernels 32 x 32 x 32 through fftconv_conv_batch
    (per-bin complex GEMM on tcgen05 over the tiles of all images of a group); 16 384 planes of 544 x 544 stay on the
    device (19.4 GB); sampled (image, kernel) pairs against the float64 direct convolution."""
    import torch
    N, H, W, F, kh, kw, K = 64, 512, 512, 32, 32, 32, 256
    g = torch.Generator(device="cuda").manual_seed(4)
    d_t = torch.rand((N, F, W, H), device="cuda", generator=g)
    b_t = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.03
    out = torch.empty((N, K, 544, 544), device="cuda")
    prof = _profiled_kernels(fc, lambda: fc.conv_batch(d_t, b_t, out))
    assert "os_gemm" in prof and "os_inverse" in prof, prof
    for n, k in ((0, 0), (3, 127), (4, 128), (31, 200), (63, 255)):
        data = d_t[n].cpu().numpy().transpose(2, 1, 0)                   # (H, W, F)
        ker = b_t[k].cpu().numpy().transpose(2, 1, 0)
        ref = oracle.direct_conv64_c(np.ascontiguousarray(data), np.ascontiguousarray(ker), 544, 544)
        assert oracle.rel_l2(out[n, k].cpu().numpy().T, ref) < TOL, (n, k)
    del out, d_t, b_t
    torch.cuda.empty_cache()


C5_SIDES = (256, 223, 194, 169, 147, 128, 111, 97, 84, 74)
C5_PLANES = (272, 240, 224, 192, 176, 144, 128, 112, 112, 96)


@pytest.mark.parametrize("prepared", [False, True])
def test_c5_all_pyramid_levels(fc, oracle, prepared):
    """config 5: the ten pyramid level sizes at F = 31 (plane sides 272 ... 96) against a bank large enough for the
    overlap-save / tcgen05 pipeline (160 templates: two GEMM blocks), one-shot and prepared bank."""
    rng = np.random.default_rng(5)
    K, F = 160, 31
    ks = [(rng.standard_normal((16, 16, F)) * 0.05).astype(np.float32) for _ in range(K)]
    bank = fc.Bank(ks) if prepared else None
    for side, plane in zip(C5_SIDES, C5_PLANES):
        assert fc.computeFFTsize16(side + 15) == plane
        data = (rng.random((side, side, F), dtype=np.float32) * 0.2).astype(np.float32)
        if prepared:
            outs = bank.conv(data)
        else:
            assert fc.lib().fftconv_query_path(side, side, F, 16, 16, K, None, None, None) == 3
            outs = fc.cudaConvolutionFFT(data, 16, 16, ks)
        for k in (0, 127, 128, K - 1):
            ref = oracle.direct_conv64_c(data, ks[k], plane, plane)
            assert oracle.rel_l2(outs[k], ref) < TOL, (side, k)
    if bank is not None:
        bank.close()


@pytest.mark.parametrize("ntblk,ahead", [(2, 0), (2, 1), (3, 2)])
def test_c2_chunked_schedules_in_a_fresh_process(ntblk, ahead):
    """Several chunks of device-resident templates, with and without the template transforms of chunk i + 1 running ahead
    on a side stream (OsAhead in csrc/fftconv.cu; FFTCONV_OS_NTBLK / FFTCONV_OS_AHEAD are read once per process, hence the
    subprocess).  scripts/oneshot_time.py checks planes at every chunk boundary against the float64 FFT convolution."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FFTCONV_OS_NTBLK=str(ntblk), FFTCONV_OS_AHEAD=str(ahead))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "oneshot_time.py"), "1000", "2"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"max rel-L2 over \d+ planes ([0-9.e+-]+) nan=(\w+)", r.stdout)
    assert m, r.stdout[-2000:]
    assert float(m.group(1)) < TOL and m.group(2) == "False", r.stdout[-500:]
    nchunks = -(-1000 // (128 * ntblk))
    assert re.search(rf"os_gemm=[0-9.]+x{nchunks}\b", r.stdout), r.stdout[-500:]      # one per-bin GEMM per chunk
