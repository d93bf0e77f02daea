"""CPU tests of the oracle itself (no GPU): the numpy restatement of the reference pipeline
must agree with the float64 direct convolution, the C oracle with scipy, and the helper
functions with their known answers."""
import numpy as np
import pytest

TOL = 1e-5   # relative-L2 tolerance stated by BASELINE.json:north_star


def test_fft_size16_known_answers(oracle):
    # src/cudaConvFFTData.h:96-102; demo: 64+9 -> 80, 8+3 -> 16 (demoCudaConvolutionFFT.m:78-79)
    kat = {1: 16, 15: 16, 16: 16, 17: 32, 73: 80, 11: 16, 271: 272, 272: 272, 273: 288, 4607: 4608, 543: 544}
    for n, want in kat.items():
        assert oracle.compute_fft_size16(n) == want
    assert oracle.compute_fft_size_pow2(73) == 128 and oracle.compute_fft_size_pow2(64) == 64
    assert oracle.i_div_up(7, 2) == 4 and oracle.i_align_up(17, 16) == 32


def test_pad_data_zero(oracle):
    rng = np.random.default_rng(0)
    src = rng.random((3, 5, 7), dtype=np.float32)
    dst = oracle.pad_data(src, 16, 32)
    assert dst.shape == (3, 16, 32)
    assert np.array_equal(dst[:, :5, :7], src) and dst[:, 5:, :].sum() == 0 and dst[:, :, 7:].sum() == 0


def test_clamp_pad(oracle):
    src = np.arange(12, dtype=np.float32).reshape(3, 4)     # [W=3][H=4]
    p = oracle.clamp_pad_data(src, 8, 8, kernel_x=2, kernel_y=1)
    # x: 0,1,2 | 2,2 (replicate) | 0,0,0 (wrap) ; y: 0..3 | 3 | 0,0,0
    assert list(p[:, 0]) == [0, 4, 8, 8, 8, 0, 0, 0]
    assert list(p[0, :]) == [0, 1, 2, 3, 3, 0, 0, 0]


def test_demo_workload_matches_direct_conv(oracle):
    data, cells, cn, cm = oracle.demo_workload(seed=1, n_kernels=3)
    outs = oracle.convolution_fft(data, cn, cm, cells)
    assert outs[0].shape == (80, 16) and outs[0].dtype == np.float32
    for k, out in zip(cells, outs):
        ref = oracle.direct_conv64(data, k, 80, 16)
        assert oracle.rel_l2(out, ref) < TOL
        assert oracle.rel_l2(out[:73, :11], ref[:73, :11]) < TOL
    # cell{1} == cell{3} determinism, cell{2} differs (kernel2(1) = 100)
    assert np.array_equal(outs[0], outs[2]) and not np.array_equal(outs[0], outs[1])
    # the flipped planted pattern makes the response peak where the pattern was planted
    # (data(5:14,2:5,1) = kernel(:,:,1), demoCudaConvolutionFFT.m:58): correlation peak at
    # 0-based peaks: (4+cn-1, 1+cm-1) = (13, 4), (20+cn-1, cm-1) = (29, 3), (cn-1, m-1) = (9, 7)
    blk = outs[0][:73, :11]
    assert np.unravel_index(np.argmax(blk), blk.shape) in [(13, 4), (29, 3), (9, 7)]


@pytest.mark.parametrize("H,W,F,kh,kw", [(20, 9, 1, 3, 3), (33, 17, 4, 7, 5), (64, 64, 31, 16, 16), (50, 3, 2, 1, 1)])
def test_reference_pipeline_vs_direct(oracle, H, W, F, kh, kw):
    rng = np.random.default_rng(H * 131 + W)
    d = rng.random((H, W, F), dtype=np.float32)
    ks = [rng.standard_normal((kh, kw, F)).astype(np.float32), rng.standard_normal((max(kh - 1, 1), kw, F)).astype(np.float32)]
    outs = oracle.convolution_fft(d, kh, kw, ks)
    FH, FW = oracle.compute_fft_size16(H + kh - 1), oracle.compute_fft_size16(W + kw - 1)
    for k, o in zip(ks, outs):
        assert o.shape == (FH, FW)
        assert oracle.rel_l2(o, oracle.direct_conv64(d, k, FH, FW)) < TOL


def test_c_oracle_matches_scipy(oracle):
    rng = np.random.default_rng(5)
    d = rng.random((40, 23, 3), dtype=np.float32)
    k = rng.standard_normal((6, 9, 3)).astype(np.float32)
    a = oracle.direct_conv64(d, k, 48, 32)
    b = oracle.direct_conv64_c(d, k, 48, 32)
    assert oracle.rel_l2(b, a) < 1e-12
    c = oracle.direct_conv64_c(d, k, 48, 32, f32=True)
    assert oracle.rel_l2(c, a) < 1e-5


def test_oversize_kernel_is_circular(oracle):
    # a kernel larger than the declared maximum aliases (circular convolution)
    rng = np.random.default_rng(7)
    d = rng.random((30, 30, 2), dtype=np.float32)
    k = rng.standard_normal((8, 8, 2)).astype(np.float32)
    spec = oracle.fft_data(d, 3, 3)          # plane 32 x 32, but 30 + 8 - 1 = 37 > 32
    out = oracle.conv_fft_data(spec, [k])[0]
    assert oracle.rel_l2(out, oracle.direct_conv64(d, k, 32, 32)) < TOL


def test_kernel_shape_error(oracle):
    rng = np.random.default_rng(8)
    spec = oracle.fft_data(rng.random((10, 10, 2), dtype=np.float32), 3, 3)
    with pytest.raises(ValueError):
        oracle.conv_fft_data(spec, [np.zeros((3, 3, 3), np.float32)])
    with pytest.raises(ValueError):
        oracle.conv_fft_data(spec, [np.zeros((17, 3, 2), np.float32)])


def test_modulate_and_normalize(oracle):
    rng = np.random.default_rng(9)
    a = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    b = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    assert np.allclose(oracle.modulate_and_normalize(a, b, 64), a * b / 64, rtol=1e-6, atol=1e-7)


def test_cpu_fft_path_matches(oracle):
    rng = np.random.default_rng(10)
    d = rng.random((64, 48, 5), dtype=np.float32)
    ks = [rng.standard_normal((9, 6, 5)).astype(np.float32)]
    outs, _ = oracle.fft_conv_cpu(d, 9, 6, ks)
    assert oracle.rel_l2(outs[0], oracle.direct_conv64(d, ks[0], 80, 64)) < TOL
