"""Small invocations of every CUDA pipeline for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`
(out-of-bounds / misaligned accesses in shared and global memory, including the TMA destinations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import numpy as np, torch
import fftconv_b200 as fc
import oracle
rng = np.random.default_rng(0)

def check(name, outs, data, ks, FH, FW):
    err = max(oracle.rel_l2(o, oracle.direct_conv64_c(data, k, FH, FW)) for o, k in zip(outs[:2], ks[:2]))
    print(f"{name}: rel-L2 {err:.2e}", flush=True)
    assert err < 1e-5

# path 3 (overlap-save, tcgen05 GEMM, TMA-box inverse): 130 templates = two MMA blocks, ragged sizes
data = rng.random((70, 50, 5), dtype=np.float32)
ks = [rng.standard_normal((int(rng.integers(3, 12)), int(rng.integers(3, 10)), 5)).astype(np.float32) for _ in range(130)]
check("path 3", fc.cudaConvolutionFFT(data, 11, 9, ks, options=fc.Options(path=3)), data, ks, 80, 64)
check("path 3 correlate", [np.roll(o, (k.shape[0] - 1, k.shape[1] - 1), (0, 1)) for o, k in
      zip(fc.cudaConvolutionFFT(data, 11, 9, ks[:70], options=fc.Options(path=3, correlate=1)), ks)],
      data, [np.ascontiguousarray(k[::-1, ::-1]) for k in ks], 80, 64)
# path 2 and 1
check("path 2", fc.cudaConvolutionFFT(data, 11, 9, ks[:5], options=fc.Options(path=2)), data, ks, 80, 64)
check("path 1", fc.cudaConvolutionFFT(data, 11, 9, ks[:3], options=fc.Options(path=1)), data, ks, 80, 64)
# path 4 (in-place large-plane pipeline): odd radices, multi-channel and single-channel, pruned and unpruned stages
d4 = rng.random((250, 130, 2), dtype=np.float32)
k4 = [rng.standard_normal((23, 15, 2)).astype(np.float32), rng.standard_normal((9, 70, 2)).astype(np.float32)]
check("path 4 multi", fc.cudaConvolutionFFT(d4, 23, 15, k4, options=fc.Options(path=4)), d4, k4, 272, 144)
d5 = rng.random((100, 300, 1), dtype=np.float32)
k5 = [rng.standard_normal((29, 21, 1)).astype(np.float32)]
check("path 4 single", fc.cudaConvolutionFFT(d5, 29, 21, k5, options=fc.Options(path=4)), d5, k5, 128, 320)
# batched + prepared bank + fused maximum
dt = torch.from_numpy(np.ascontiguousarray(np.stack([data, data[::-1].copy()]).transpose(0, 3, 2, 1))).cuda()
bt = torch.from_numpy(np.stack([np.ascontiguousarray(np.pad(k, ((0, 11 - k.shape[0]), (0, 9 - k.shape[1]), (0, 0))).transpose(2, 1, 0)) for k in ks[:70]])).cuda()
out = fc.conv_batch(dt, bt); torch.cuda.synchronize()
print("batch:", tuple(out.shape), flush=True)
bank = fc.Bank(ks[:70]); pk = bank.conv_max(data); bank.close()
print("peaks:", float(pk[0][0]), int(pk[1][0]), int(pk[2][0]), flush=True)
print("sanitize_small done", fc.launch_count())
# round 2: size-specialised large-plane kernels (1152-point lines), pyramid batch (levels of different sizes in one call)
d6 = rng.random((60, 1152 - 20 + 1 - 2, 1), dtype=np.float32)
k6 = [(rng.standard_normal((5, 20, 1)) / 8).astype(np.float32)]
check("path 4 size-specialised w pass", fc.cudaConvolutionFFT(d6, 5, 20, k6, options=fc.Options(path=4)), d6, k6, 64, 1152)
d7 = rng.random((1152 - 9 + 1 - 3, 30, 2), dtype=np.float32)
k7 = [(rng.standard_normal((9, 7, 2)) / 8).astype(np.float32)]
check("path 4 size-specialised h pass", fc.cudaConvolutionFFT(d7, 9, 7, k7, options=fc.Options(path=4)), d7, k7, 1152, 48)
lv = [torch.from_numpy(np.ascontiguousarray(rng.random((s, s + 3, 5), dtype=np.float32).transpose(2, 1, 0))).cuda() for s in (70, 41, 23)]
po = fc.conv_pyramid(lv, bt, 11, 9); torch.cuda.synchronize()
print("pyramid:", [tuple(o.shape) for o in po], flush=True)
print("sanitize_small round-2 done", fc.launch_count())
