"""Short run of the C2 workload on one pipeline for ncu captures: python scripts/ncu_os.py [path] [K] [calls]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc
path = int(sys.argv[1]) if len(sys.argv) > 1 else 3
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H = W = 256; F = 31; kh = kw = 16
g = torch.Generator(device="cuda").manual_seed(2)
data = torch.rand((F, W, H), device="cuda", generator=g) * 0.2
bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
spec = fc.fft_data_device(data, H, W, F, kh, kw)
out = torch.empty((K, 272, 272), device="cuda")
for _ in range(calls):
    fc.conv_bank(spec, bank, kh, kw, out, options=fc.Options(path=path))
torch.cuda.synchronize()
print("done", fc.launch_count())
