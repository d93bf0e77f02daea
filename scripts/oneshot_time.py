"""Device-resident timing of the one-shot entry point (fftconv_convolution_fft) on the C2 shapes, L2 flushed between calls;
   checks planes of every chunk against the float64 FFT convolution.   python scripts/oneshot_time.py [K] [reps]
   Used for the chunking / run-ahead experiments (FFTCONV_OS_NTBLK, FFTCONV_OS_AHEAD are read once at load)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    H = W = 256; F = 31; kh = kw = 16; FH = FW = 272
    g = torch.Generator(device="cuda").manual_seed(2)
    data = torch.rand((F, W, H), device="cuda", generator=g) * 0.2
    bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
    out = torch.full((K, FW, FH), float("nan"), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fc.convolution_fft_device(data, bank, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fc.convolution_fft_device(data, bank, out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    ks = sorted({0, 127, 128, 255, 256, 383, 511, 512, 639, 640, 767, 768, 895, 896, K - 1} & set(range(K)))
    idx = torch.tensor(ks, device="cuda")
    ref = torch.fft.irfft2(torch.fft.rfft2(data.double(), s=(FW, FH)).unsqueeze(0) *
                           torch.fft.rfft2(bank[idx].double(), s=(FW, FH)), s=(FW, FH)).sum(1)
    err = float(((out[idx].double() - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max())
    nan = bool(torch.isnan(out).any())
    fc.profile(True); fc.profile_read(True)
    for _ in range(3):
        fc.convolution_fft_device(data, bank, out)
    torch.cuda.synchronize()
    pr = fc.profile_read(True); fc.profile(False)
    env = {k: v for k, v in os.environ.items() if k.startswith("FFTCONV_")}
    print(f"{env} K={K}: median {ms:.4f} ms min {min(ts):.4f}  max rel-L2 over {len(ks)} planes {err:.2e} nan={nan}  "
          + " ".join(f"{n.split('(')[0]}={t/3:.3f}x{c//3}" for n, (t, c) in pr.items()), flush=True)


if __name__ == "__main__":
    main()
