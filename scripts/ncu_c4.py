"""Short run of config 4 (batched images, per-bin complex GEMM on tcgen05) for ncu captures: python scripts/ncu_c4.py [N images]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import fftconv_b200 as fc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator(device="cuda").manual_seed(4)
data = torch.rand((N, 32, 512, 512), device="cuda", generator=g)
bank = torch.randn((256, 32, 32, 32), device="cuda", generator=g) * 0.03
out = torch.empty((N, 256, 544, 544), device="cuda")
for _ in range(2):
    fc.conv_batch(data, bank, out)
torch.cuda.synchronize()
print("done", fc.launch_count())
