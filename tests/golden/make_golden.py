"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Runs oracle/_ref/libref_replay.so — the reference's own device kernels (padData,
elementwiseProductAndNormalize, sumAlongFeatures, src/cudaConvFFTData.cuh:11-92) and its host
sequence (src/cudaConvolutionFFT.cu:109-310) compiled from /root/reference and linked against
cuFFT 11.4 — on seeded inputs on a B200, and stores inputs' seeds + outputs as .npz.

    gpurun -- python tests/golden/make_golden.py          (needs a GPU; run once, commit the .npz)

The fixtures pin BOTH the CPU oracle (tests/test_golden.py, no GPU) and the CUDA path (-m gpu).
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_replay.so")
    L = ctypes.CDLL(so)
    L.ref_convolution_fft.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 6
    return L


def ref_run(L, data_hwf, max_kh, max_kw, kernels, threads=None):
    d = oracle.to_fwh(data_hwf)
    F, W, H = d.shape
    FH, FW = oracle.compute_fft_size16(H + max_kh - 1), oracle.compute_fft_size16(W + max_kw - 1)
    K = len(kernels)
    ks = [oracle.to_fwh(k) for k in kernels]
    kp = (ctypes.c_void_p * K)(*[k.ctypes.data for k in ks])
    kh = (ctypes.c_int * K)(*[k.shape[2] for k in ks])
    kw = (ctypes.c_int * K)(*[k.shape[1] for k in ks])
    outs = np.zeros((K, FW, FH), dtype=np.float32)
    op = (ctypes.c_void_p * K)(*[outs.ctypes.data + 4 * i * FW * FH for i in range(K)])
    th = (ctypes.c_int * 4)(*threads) if threads else None
    ms = ctypes.c_float(0)
    rc = L.ref_convolution_fft(d.ctypes.data, H, W, F, max_kh, max_kw, K, kp, kh, kw, op, th, ctypes.byref(ms))
    assert rc == 0, rc
    return outs, float(ms.value)


def cases():
    """name -> (data, max_kh, max_kw, kernels, threads).  All inputs are rebuilt from seeds by the tests."""
    out = {}
    data, cells, cn, cm = oracle.demo_workload(seed=1, n_kernels=3)
    out["c1_demo"] = (data, cn, cm, cells, [8, 8, 8, 16])                 # demoCudaConvolutionFFT.m:115-129
    rng = np.random.default_rng(42)
    d = rng.random((40, 27, 3), dtype=np.float32)
    ks = [rng.standard_normal((7, 5, 3)).astype(np.float32), rng.standard_normal((4, 5, 3)).astype(np.float32)]
    out["small_48x32"] = (d, 7, 5, ks, None)
    rng = np.random.default_rng(2)
    d = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
    ks = [(rng.standard_normal((16, 16, 31)) * 0.05).astype(np.float32),
          (rng.standard_normal((9, 13, 31)) * 0.05).astype(np.float32)]
    out["c2_hog_272"] = (d, 16, 16, ks, None)
    return out


def main():
    L = load_ref()
    meta = {}
    for name, (data, mkh, mkw, ks, th) in cases().items():
        outs, ms = ref_run(L, data, mkh, mkw, ks, th)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), outs=outs)
        errs = [oracle.rel_l2(oracle.from_wh(outs[i]), oracle.direct_conv64_c(data, ks[i], outs.shape[2], outs.shape[1]))
                for i in range(len(ks))]
        meta[name] = {"planes": list(outs.shape), "rel_l2_vs_fp64_direct": errs}
        print(name, outs.shape, "rel-L2 vs float64 direct conv:", errs)
    # timing of the reference's own GPU path on the C2 workload (sample of templates, includes its
    # per-template cudaMalloc / blocking copies / device syncs exactly as the MEX loop does)
    rng = np.random.default_rng(2)
    d = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
    ks = [(rng.standard_normal((16, 16, 31)) * 0.05).astype(np.float32) for _ in range(100)]
    ref_run(L, d, 16, 16, ks[:10])
    t0 = time.perf_counter()
    _, ms = ref_run(L, d, 16, 16, ks)
    wall = time.perf_counter() - t0
    meta["ref_replay_c2_timing"] = {"templates": 100, "kernel_loop_ms": ms, "wall_s": wall,
                                    "outputs_per_s": 100 * 272 * 272 / (ms * 1e-3)}
    print("reference replay (cuFFT 11.4 + reference kernels) on C2: %.3f ms per template -> %.3e outputs/s"
          % (ms / 100, 100 * 272 * 272 / (ms * 1e-3)))
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
