"""Config 5 on one GPU: 10-level pyramid x K templates, one call (fftconv_conv_pyramid) against the per-level loop.
python scripts/c5_time.py [K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import numpy as np, torch
import fftconv_b200 as fc
from fftconv_b200.pyramid import pyramid_convolution_cuda, pyramid_sides, level_plane
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
F, kh, kw = 31, 16, 16
sides = pyramid_sides()
g = torch.Generator(device="cuda").manual_seed(5)
levels = [torch.rand((F, s, s), device="cuda", generator=g) * 0.2 for s in sides]
bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
shapes = [(s, s, F) for s in sides]
outs = [torch.empty((K,) + level_plane(s, s, kh, kw)[::-1], device="cuda") for s in sides]
nout = sum(K * fh * fw for fh, fw in (level_plane(s, s, kh, kw) for s in sides))
res = {}
for one in (False, True):
    step = lambda: pyramid_convolution_cuda(levels, shapes, bank, kh, kw, outs, one_call=one)
    for _ in range(2): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    res[one] = [o[:3].clone() for o in outs]
    print(f"[c5 one_call={one}] K={K}: {ms:.2f} ms -> {nout / ms / 1e6:.1f} G outputs/s", flush=True)
    fc.profile(True); fc.profile_read(True); step(); torch.cuda.synchronize()
    for name, (t, n) in fc.profile_read(True).items():
        print(f"     {name:28s} {t:9.3f} ms ({n} launches)")
    fc.profile(False)
print("max rel diff one_call vs per level:", max(float((a - b).norm() / b.norm()) for a, b in zip(res[True], res[False])))
