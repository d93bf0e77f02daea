"""Device-resident timing of the BASELINE configs beyond the bench workload (C3 large kernels, C4 batched,
C5 pyramid) on one GPU, with a parity spot check each.  python scripts/config_time.py [c3] [c4] [c5] [scale]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc
from fftconv_b200.pyramid import pyramid_convolution_cuda, pyramid_convolution_prepared, pyramid_sides, level_plane

which = [a for a in sys.argv[1:] if a.startswith("c")] or ["c4", "c5"]
scale = next((float(a) for a in sys.argv[1:] if not a.startswith("c")), 1.0)
g = torch.Generator(device="cuda").manual_seed(4)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def ref_fft(data, bank, FW, FH):
    d64 = data.double(); k64 = bank.double()
    return torch.fft.irfft2(torch.fft.rfft2(d64, s=(FW, FH)).unsqueeze(0) * torch.fft.rfft2(k64, s=(FW, FH)), s=(FW, FH)).sum(1)


if "c4" in which:
    N, H, W, F, kh, kw, K = max(2, int(64 * scale)), 512, 512, 32, 32, 32, 256
    data = torch.rand((N, F, W, H), device="cuda", generator=g)
    bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.03
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    out = torch.empty((N, K, FW, FH), device="cuda")
    ms = timeit(lambda: fc.conv_batch(data, bank, out), 2)
    ref = ref_fft(data[N - 1], bank[:2], FW, FH)
    err = float((out[N - 1, :2].double() - ref).norm() / ref.norm())
    print(f"[c4] {N} images 512x512x32 x {K} kernels 32x32: {ms:.2f} ms -> {N*K*FH*FW/ms/1e6:.2f} G outputs/s, rel-L2 {err:.2e}", flush=True)
    fc.profile(True); fc.profile_read(True); fc.conv_batch(data, bank, out); torch.cuda.synchronize()
    for name, (t, n) in fc.profile_read(True).items():
        print(f"     {name:28s} {t:9.3f} ms ({n} launches)")
    fc.profile(False)
    del data, out
    fc.lib().fftconv_release()

if "c5" in which:
    F, kh, kw, K = 31, 16, 16, max(256, int(20000 * scale))
    sides = pyramid_sides()
    levels = [torch.rand((F, s, s), device="cuda", generator=g) * 0.2 for s in sides]
    bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
    shapes = [(s, s, F) for s in sides]
    outs = [torch.empty((K,) + level_plane(s, s, kh, kw)[::-1], device="cuda") for s in sides]
    ms = timeit(lambda: pyramid_convolution_cuda(levels, shapes, bank, kh, kw, outs), 2)
    nout = sum(K * o.shape[1] * o.shape[2] for o in outs)
    FH, FW = level_plane(sides[3], sides[3], kh, kw)
    ref = ref_fft(levels[3], bank[:2], FW, FH)
    err = float((outs[3][:2].double() - ref).norm() / ref.norm())
    print(f"[c5] 10-level pyramid x {K} templates 16x16x31: {ms:.2f} ms -> {nout/ms/1e6:.2f} G outputs/s, rel-L2 {err:.2e}", flush=True)
    t0 = time.time()
    import ctypes        # device-resident templates handed to fftconv_bank_create without a host round trip
    h = ctypes.c_void_p(0)
    kp = (ctypes.c_void_p * K)(*[bank.data_ptr() + 4 * k * F * kw * kh for k in range(K)])
    khs = (ctypes.c_int * K)(*([kh] * K)); kws = (ctypes.c_int * K)(*([kw] * K)); ond = (ctypes.c_ubyte * K)(*([1] * K))
    rc = fc.lib().fftconv_bank_create(K, kp, khs, kws, None, ond, F, 0, torch.cuda.current_stream().cuda_stream, ctypes.byref(h))
    assert rc == 0, fc.last_error()
    pb = fc.Bank.__new__(fc.Bank); pb._h = h; pb.K, pb.F, pb.maxKH, pb.maxKW, pb.device = K, F, kh, kw, 0
    torch.cuda.synchronize(); t_prep = (time.time() - t0) * 1e3
    ms = timeit(lambda: pyramid_convolution_prepared(levels, shapes, pb, outs), 2)
    err = float((outs[3][:2].double() - ref).norm() / ref.norm())
    print(f"[c5] same, PREPARED bank (one-off transform {t_prep:.1f} ms): {ms:.2f} ms -> {nout/ms/1e6:.2f} G outputs/s, rel-L2 {err:.2e}", flush=True)
    pb.close()
    del outs
    fc.lib().fftconv_release()

if "c3" in which:
    H = W = 4096; F = 1; kh = kw = 512; K = max(2, int(64 * scale))
    data = torch.rand((F, W, H), device="cuda", generator=g)
    bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) / 512
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    spec = fc.fft_data_device(data, H, W, F, kh, kw)
    out = torch.empty((K, FW, FH), device="cuda")
    ms = timeit(lambda: fc.conv_bank(spec, bank, kh, kw, out), 2)
    ref = ref_fft(data, bank[:1], FW, FH)
    err = float((out[:1].double() - ref).norm() / ref.norm())
    print(f"[c3] 4096x4096 x {K} kernels 512x512: {ms:.2f} ms -> {K*FH*FW/ms/1e6:.2f} G outputs/s, rel-L2 {err:.2e}", flush=True)
    fc.profile(True); fc.profile_read(True); fc.conv_bank(spec, bank, kh, kw, out); torch.cuda.synchronize()
    for name, (t, n) in fc.profile_read(True).items():
        print(f"     {name:28s} {t:9.3f} ms ({n} launches)")
    fc.profile(False)
