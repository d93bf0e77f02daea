"""Quick device-resident timing of the C2 (HOG-DPM) workload; prints ms per call and outputs/s."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc

def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    generic = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    H = W = 256; F = 31; kh = kw = 16
    g = torch.Generator(device="cuda").manual_seed(2)
    data = torch.rand((F, W, H), device="cuda", generator=g) * 0.2
    bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
    spec = fc.fft_data_device(data, H, W, F, kh, kw)
    FH = FW = 272
    out = torch.empty((K, FW, FH), device="cuda")
    opt = fc.Options(force_generic=generic)
    for _ in range(3):
        fc.conv_bank(spec, bank, kh, kw, out, options=opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record(); fc.conv_bank(spec, bank, kh, kw, out, options=opt); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"K={K} generic={generic}: median {ms:.3f} ms  min {min(ts):.3f}  -> {K*FH*FW/ms/1e6:.2f} G outputs/s; launches so far {fc.launch_count()}")
    # correctness spot check against fp64 torch FFT conv
    d64 = data.double(); k64 = bank[:4].double()
    ref = torch.fft.irfft2(torch.fft.rfft2(d64, s=(FW, FH)).unsqueeze(0) * torch.fft.rfft2(k64, s=(FW, FH)), s=(FW, FH)).sum(1)
    err = (out[:4].double() - ref).norm() / ref.norm()
    print("rel-L2 vs fp64 FFT conv:", float(err))

if __name__ == "__main__":
    main()
