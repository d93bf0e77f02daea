"""Short run of config 3 (large-plane path) for ncu captures: python scripts/ncu_bp.py [K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = W = 4096; F = 1; kh = kw = 512
g = torch.Generator(device="cuda").manual_seed(3)
data = torch.rand((F, W, H), device="cuda", generator=g)
bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) / 512
spec = fc.fft_data_device(data, H, W, F, kh, kw)
out = torch.empty((K, 4608, 4608), device="cuda")
for _ in range(2):
    fc.conv_bank(spec, bank, kh, kw, out)
torch.cuda.synchronize()
print("done", fc.launch_count())
