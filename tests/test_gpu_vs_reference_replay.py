"""LIVE parity against the reference itself on the GPU box: the reference's own device kernels + host sequence
(oracle/ref_replay.cu #includes src/cudaConvFFTData.h/.cuh from /root/reference at build time and replays
src/cudaConvolutionFFT.cu:109-310 against cuFFT) run next to the CUDA product path on the same fresh random inputs.
The committed fixtures (tests/golden) pin three cases; this sweeps more shapes, every pipeline of the product, and the
BASELINE config-2 bank.  Skipped when oracle/_ref/libref_replay.so was not built (no reference sources at build time).
Tolerance: rel-L2 <= 1e-5 on the whole FFT_H x FFT_W plane (BASELINE.json:north_star)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_replay.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libref_replay.so not built")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    L = make_golden.load_ref()
    return lambda data, kh, kw, ks, th=None: make_golden.ref_run(L, data, kh, kw, ks, th)[0]


def _planes(ref_outs):
    return [o.T for o in ref_outs]          # [FW][FH] memory == (FH, FW) column-major


@pytest.mark.parametrize("path", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("H,W,F,kh,kw,K", [(64, 8, 5, 10, 4, 3), (97, 130, 3, 11, 16, 5), (200, 180, 2, 16, 9, 70)])
def test_every_pipeline_matches_the_reference_replay(fc, oracle, ref, path, H, W, F, kh, kw, K):
    rng = np.random.default_rng(1000 * path + H + K)
    data = rng.random((H, W, F), dtype=np.float32)
    ks = [rng.standard_normal((kh, kw, F)).astype(np.float32) for _ in range(K - 1)]
    ks.append(rng.standard_normal((max(1, kh - 3), max(1, kw - 2), F)).astype(np.float32))
    want = _planes(ref(data, kh, kw, ks))
    got = fc.cudaConvolutionFFT(data, kh, kw, ks, options=fc.Options(path=path))
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert oracle.rel_l2(g, w) < TOL


def test_c2_bank_sample_matches_the_reference_replay(fc, oracle, ref):
    """BASELINE config 2 shapes, 96 templates (enough for the overlap-save / tcgen05 path), reference thread shape of the demo."""
    rng = np.random.default_rng(2)
    data = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
    ks = [(rng.standard_normal((16, 16, 31)) * 0.05).astype(np.float32) for _ in range(64)]
    ks += [(rng.standard_normal((int(rng.integers(6, 17)), int(rng.integers(6, 17)), 31)) * 0.05).astype(np.float32)
           for _ in range(32)]
    want = _planes(ref(data, 16, 16, ks, [8, 8, 8, 16]))
    spec = fc.cudaFFTData(data, 16, 16)
    got = fc.cudaConvFFTData(spec, ks)
    worst = max(oracle.rel_l2(g, w) for g, w in zip(got, want))
    assert worst < TOL, worst


def test_large_plane_matches_the_reference_replay(fc, oracle, ref):
    """config 3 structure at 1/4 scale (plane 1152 x 1152, large-plane path) against cuFFT through the reference's loop."""
    rng = np.random.default_rng(3)
    data = rng.random((1024, 1024, 1), dtype=np.float32)
    ks = [(rng.standard_normal((128, 128, 1)) / 128).astype(np.float32) for _ in range(2)]
    want = _planes(ref(data, 128, 128, ks))
    got = fc.cudaConvolutionFFT(data, 128, 128, ks)
    for g, w in zip(got, want):
        assert oracle.rel_l2(g, w) < TOL
