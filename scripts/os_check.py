"""Bring-up check of the overlap-save / tcgen05 path (run on the GPU box):
   python scripts/os_check.py [simt|tc|tcswap] ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import numpy as np
import fftconv_b200 as fc
import oracle

def run(tag, H, W, F, kh, kw, K, vary=False, spectrum=False):
    rng = np.random.default_rng(H + 3 * K)
    data = (rng.random((H, W, F), dtype=np.float32) * 0.2).astype(np.float32)
    ks = []
    for k in range(K):
        a, b = (kh, kw) if not vary or k % 3 == 0 else (int(rng.integers(1, kh + 1)), int(rng.integers(1, kw + 1)))
        ks.append((rng.standard_normal((a, b, F)) * 0.05).astype(np.float32))
    opt = fc.Options(path=3)
    t0 = time.time()
    if spectrum:
        spec = fc.cudaFFTData(data, kh, kw)
        outs = fc.cudaConvFFTData(spec, ks, options=opt)
    else:
        outs = fc.cudaConvolutionFFT(data, kh, kw, ks, options=opt)
    dt = time.time() - t0
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    worst = 0.0
    idx = sorted(set([0, 1, K // 2, K - 1]) & set(range(K)))
    for i in idx:
        ref = oracle.direct_conv64_c(data, ks[i], FH, FW)
        worst = max(worst, oracle.rel_l2(outs[i], ref))
    print(f"[{tag}] {H}x{W}x{F} k{kh}x{kw} K={K} spectrum={spectrum}: worst rel-L2 {worst:.3e}  ({dt*1e3:.1f} ms)", flush=True)
    return worst

for mode in (sys.argv[1:] or ["simt", "tc"]):
    os.environ["FFTCONV_OS_GEMM"] = "simt" if mode == "simt" else "tc"
    os.environ["FFTCONV_OS_LBO_SWAP"] = "1" if mode == "tcswap" else "0"
    run(mode, 64, 64, 5, 7, 7, 3)
    run(mode, 100, 37, 3, 16, 2, 5, vary=True)
    run(mode, 64, 8, 5, 10, 4, 10)
    run(mode, 40, 40, 2, 20, 31, 3, vary=True)
    run(mode, 256, 256, 31, 16, 16, 130, vary=True)
    run(mode, 256, 256, 31, 16, 16, 130, vary=True, spectrum=True)
    run(mode, 300, 200, 8, 9, 12, 300)
print("launches", fc.launch_count())
