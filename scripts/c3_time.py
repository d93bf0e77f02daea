"""Config 3 (4096^2 x K kernels 512^2) on the large-plane path: device-resident time and per-kernel breakdown.
python scripts/c3_time.py [K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import numpy as np, torch
import fftconv_b200 as fc
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = W = 4096; F = 1; kh = kw = 512
g = torch.Generator(device="cuda").manual_seed(3)
data = torch.rand((F, W, H), device="cuda", generator=g)
bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) / 512
FH = FW = 4608
spec = fc.fft_data_device(data, H, W, F, kh, kw)
out = torch.empty((K, FW, FH), device="cuda")
ref = torch.fft.irfft2(torch.fft.rfft2(data.double(), s=(FW, FH)) * torch.fft.rfft2(bank[:1].double(), s=(FW, FH)), s=(FW, FH)).sum(1)
variants = [dict()]
for v in variants:
    os.environ.update(v)
    fc.lib().fftconv_release()
    for _ in range(2):
        fc.conv_bank(spec, bank, kh, kw, out)
    torch.cuda.synchronize()
    err = float((out[:1].double() - ref).norm() / ref.norm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); fc.conv_bank(spec, bank, kh, kw, out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"[c3 {v}] K={K}: {ms:.2f} ms ({ms / K * 1e3:.1f} us/kernel) -> {K*FH*FW/ms/1e6:.2f} G outputs/s, rel-L2 {err:.2e}", flush=True)
    fc.profile(True); fc.profile_read(True); fc.conv_bank(spec, bank, kh, kw, out); torch.cuda.synchronize()
    for name, (t, n) in fc.profile_read(True).items():
        print(f"     {name:28s} {t:9.3f} ms ({n} launches)")
    fc.profile(False)
