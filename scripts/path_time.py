"""Device-resident timing of one workload on each pipeline with the per-kernel breakdown.
   python scripts/path_time.py [c2|c4img|c3s] [paths, e.g. 2,3] [K]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc

SHAPES = {
    "c2": (256, 256, 31, 16, 16, 1000),
    "c2s": (256, 256, 31, 16, 16, 256),
    "c4img": (512, 512, 32, 32, 32, 256),
    "c5l9": (74, 74, 31, 16, 16, 2000),
}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    paths = [int(p) for p in (sys.argv[2] if len(sys.argv) > 2 else "2,3").split(",")]
    H, W, F, kh, kw, K = SHAPES[wl]
    if len(sys.argv) > 3:
        K = int(sys.argv[3])
    g = torch.Generator(device="cuda").manual_seed(2)
    data = torch.rand((F, W, H), device="cuda", generator=g) * 0.2
    bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
    spec = fc.fft_data_device(data, H, W, F, kh, kw)
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    out = torch.empty((K, FW, FH), device="cuda")
    d64 = data.double(); k64 = bank[:4].double()
    ref = torch.fft.irfft2(torch.fft.rfft2(d64, s=(FW, FH)).unsqueeze(0) * torch.fft.rfft2(k64, s=(FW, FH)), s=(FW, FH)).sum(1)
    for path in paths:
        opt = fc.Options(path=path)
        out.zero_()
        for _ in range(3):
            fc.conv_bank(spec, bank, kh, kw, out, options=opt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(10):
            e0.record(); fc.conv_bank(spec, bank, kh, kw, out, options=opt); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        err = float((out[:4].double() - ref).norm() / ref.norm())
        print(f"[{wl}] path={path} K={K} plane {FH}x{FW}: median {ms:.3f} ms min {min(ts):.3f} -> "
              f"{K*FH*FW/ms/1e6:.2f} G outputs/s  rel-L2 {err:.2e}", flush=True)
        fc.profile(True)
        fc.profile_read(True)
        for _ in range(3):
            fc.conv_bank(spec, bank, kh, kw, out, options=opt)
        torch.cuda.synchronize()
        pr = fc.profile_read(True)
        fc.profile(False)
        for name, (t, n) in pr.items():
            print(f"     {name:28s} {t/3:8.3f} ms/call  ({n//3} launches/call)")


if __name__ == "__main__":
    main()
