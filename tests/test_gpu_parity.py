"""GPU parity tests: the CUDA path, called through the reference-shaped interface over the
C ABI, against the CPU oracle (numpy restatement of the reference pipeline and float64
direct convolution).  Tolerance: relative L2 <= 1e-5 (BASELINE.json:north_star), on the full
FFT_H x FFT_W plane and on the cropped (H+kh-1) x (W+kw-1) block."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _check(oracle, outs, data, kernels, FH, FW, tol=TOL):
    H, W, _ = data.shape
    for k, o in zip(kernels, outs):
        kh, kw, _ = k.shape
        assert o.shape == (FH, FW) and o.dtype == np.float32
        ref = oracle.direct_conv64_c(data, k, FH, FW)
        assert oracle.rel_l2(o, ref) < tol
        ch, cw = min(H + kh - 1, FH), min(W + kw - 1, FW)
        assert oracle.rel_l2(o[:ch, :cw], ref[:ch, :cw]) < tol


def test_demo_workload_c1(fc, oracle):
    """BASELINE config 1: demoCudaConvolutionFFT.m workload, called exactly like the demo (:124-129)."""
    data, cells, cn, cm = oracle.demo_workload(seed=1, n_kernels=10)
    outs = fc.cudaConvolutionFFT(data, cn, cm, cells, [8, 8, 8, 16], 0)
    _check(oracle, outs, data, cells, 80, 16)
    assert np.array_equal(outs[0], outs[2]) and not np.array_equal(outs[0], outs[1])
    ref = oracle.convolution_fft(data, cn, cm, cells)          # restated reference pipeline (fp32)
    for o, r in zip(outs, ref):
        assert oracle.rel_l2(o, r) < TOL
    blk = outs[0][:73, :11]
    assert np.unravel_index(np.argmax(blk), blk.shape) in [(13, 4), (29, 3), (9, 7)]


def test_two_call_path_equals_one_shot(fc, oracle):
    data, cells, cn, cm = oracle.demo_workload(seed=3, n_kernels=4)
    spec = fc.cudaFFTData(data, cn, cm)
    assert spec.shape == (41, 16, 5)
    outs2 = fc.cudaConvFFTData(spec, cells)
    outs1 = fc.cudaConvolutionFFT(data, cn, cm, cells)
    outs3 = fc.cudaConvFFTDataStreams(spec, cells, [16, 8, 8, 32])
    for a, b, c in zip(outs1, outs2, outs3):
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_spectrum_matches_reference_layout(fc, oracle):
    """cudaFFTData output = cuFFT R2C layout [(FH/2+1), FW, F] (src/cudaFFTData.cu:90-103,128-146)."""
    rng = np.random.default_rng(11)
    data = rng.random((50, 20, 3), dtype=np.float32)
    spec = fc.cudaFFTData(data, 7, 5)
    got = spec.tensor.cpu().numpy()                      # [F][FW][CH]
    ref = oracle.fft_data(data, 7, 5)
    assert got.shape == ref.shape == (3, 32, 33)
    assert oracle.rel_l2(np.stack([got.real, got.imag]), np.stack([ref.real, ref.imag])) < TOL
    assert spec.gather().shape == (33, 32, 3)


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("H,W,F,kh,kw", [
    (16, 16, 1, 1, 1), (20, 9, 2, 3, 3), (33, 17, 4, 7, 5), (64, 64, 31, 16, 16), (50, 3, 2, 1, 4),
    (100, 37, 3, 16, 2), (71, 130, 2, 11, 13), (40, 40, 2, 20, 33), (241, 97, 2, 16, 16), (31, 289, 1, 5, 16),
])
def test_shape_sweep(fc, oracle, H, W, F, kh, kw, generic):
    rng = np.random.default_rng(H * 1000 + W * 7 + F)
    data = rng.random((H, W, F), dtype=np.float32)
    kernels = [rng.standard_normal((kh, kw, F)).astype(np.float32),
               rng.standard_normal((max(1, kh - 2), max(1, kw - 1), F)).astype(np.float32),
               rng.standard_normal((1, 1, F)).astype(np.float32)]
    opt = fc.Options(force_generic=generic)
    outs = fc.cudaConvolutionFFT(data, kh, kw, kernels, options=opt)
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    _check(oracle, outs, data, kernels, FH, FW)


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("m", list(range(1, 41)))
def test_every_multiple_of_16(fc, oracle, m, generic):
    """size sweep over every 16*m plane side, m = 1..40 (odd radices 3..37)."""
    n = 16 * m
    kh = kw = 5
    H, W = n - kh + 1, max(1, (n // 2) - kw + 1 - (m % 3))
    rng = np.random.default_rng(m)
    data = rng.random((H, W, 2), dtype=np.float32)
    kernels = [rng.standard_normal((kh, kw, 2)).astype(np.float32)]
    outs = fc.cudaConvolutionFFT(data, kh, kw, kernels, options=fc.Options(force_generic=generic))
    _check(oracle, outs, data, kernels, n, fc.computeFFTsize16(W + kw - 1))


def test_hog_sized_c2_sample(fc, oracle):
    """BASELINE config 2 shapes (256x256x31 feature map, 16x16x31 templates), a sample of templates."""
    rng = np.random.default_rng(2)
    data = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
    kernels = [(rng.standard_normal((16, 16, 31)) * 0.05).astype(np.float32) for _ in range(5)]
    kernels += [(rng.standard_normal((int(rng.integers(6, 17)), int(rng.integers(6, 17)), 31)) * 0.05).astype(np.float32)
                for _ in range(5)]
    spec = fc.cudaFFTData(data, 16, 16)
    assert spec.shape == (137, 272, 31)
    outs = fc.cudaConvFFTData(spec, kernels)
    _check(oracle, outs, data, kernels, 272, 272)
    outs_g = fc.cudaConvFFTData(spec, kernels, options=fc.Options(force_generic=1))
    for a, b in zip(outs, outs_g):
        assert oracle.rel_l2(a, b) < TOL


def test_properties_linearity_delta_shift(fc, oracle):
    rng = np.random.default_rng(21)
    H, W, F, kh, kw = 90, 75, 3, 9, 12
    data = rng.random((H, W, F), dtype=np.float32)
    spec = fc.cudaFFTData(data, kh, kw)
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    a = rng.standard_normal((kh, kw, F)).astype(np.float32)
    b = rng.standard_normal((kh, kw, F)).astype(np.float32)
    oa, ob, oab = fc.cudaConvFFTData(spec, [a, b, (2 * a - 3 * b).astype(np.float32)])
    assert oracle.rel_l2(oab, 2 * oa.astype(np.float64) - 3 * ob) < TOL          # linearity
    delta = np.zeros((kh, kw, F), np.float32)
    delta[0, 0, :] = 1
    od, = fc.cudaConvFFTData(spec, [delta])                                        # delta => sum_f data
    assert oracle.rel_l2(od[:H, :W], data.sum(axis=2, dtype=np.float64)) < TOL
    assert np.abs(od[H:, :]).max() < 1e-4 and np.abs(od[:, W:]).max() < 1e-4
    sh = np.zeros((kh, kw, F), np.float32)
    sh[3, 5, :] = 1                                                                # shifted delta => shift
    os_, = fc.cudaConvFFTData(spec, [sh])
    assert oracle.rel_l2(os_[3:3 + H, 5:5 + W], data.sum(axis=2, dtype=np.float64)) < TOL


def test_oversize_kernel_wraps(fc, oracle):
    rng = np.random.default_rng(7)
    d = rng.random((30, 30, 2), dtype=np.float32)
    k = rng.standard_normal((8, 8, 2)).astype(np.float32)
    spec = fc.cudaFFTData(d, 3, 3)
    out, = fc.cudaConvFFTData(spec, [k])
    assert oracle.rel_l2(out, oracle.direct_conv64(d, k, 32, 32)) < TOL
    with pytest.raises(fc.FFTConvError) as e:
        fc.cudaConvFFTData(spec, [np.zeros((33, 3, 2), np.float32)])
    assert "same number of features" in str(e.value)
    with pytest.raises(fc.FFTConvError):
        fc.cudaConvFFTData(spec, [np.zeros((3, 3, 3), np.float32)])


def test_gpuarray_kernels_and_empty_cell(fc, oracle):
    data, cells, cn, cm = oracle.demo_workload(seed=5, n_kernels=4)
    spec = fc.cudaFFTData(data, cn, cm)
    mixed = [cells[0], fc.gpuArray(cells[1]), cells[2], fc.gpuArray(cells[3])]
    outs = fc.cudaConvFFTData(spec, mixed)
    ref = fc.cudaConvFFTData(spec, cells)
    for a, b in zip(outs, ref):
        assert np.array_equal(a, b)
    assert fc.cudaConvFFTData(spec, []) == []
    with pytest.raises(fc.FFTConvError):
        fc.cudaConvFFTDataStreams(spec, mixed)        # host kernels only (src/cudaConvFFTDataStreams.cu:352-374)


def test_extensions_correlate_crop_clamp(fc, oracle):
    rng = np.random.default_rng(31)
    H, W, F, kh, kw = 60, 44, 3, 8, 6
    data = rng.random((H, W, F), dtype=np.float32)
    k = rng.standard_normal((kh, kw, F)).astype(np.float32)
    spec = fc.cudaFFTData(data, kh, kw)
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    full, = fc.cudaConvFFTData(spec, [k])
    # correlation mode == convolution with the flipped kernel, circularly shifted by (kh-1, kw-1)
    corr, = fc.cudaConvFFTData(spec, [k], options=fc.Options(correlate=1))
    flip, = fc.cudaConvFFTData(spec, [np.ascontiguousarray(k[::-1, ::-1, :])])
    assert oracle.rel_l2(np.roll(corr, (kh - 1, kw - 1), axis=(0, 1)), flip) < TOL
    # fused crop (device-resident bank API)
    import torch
    bank = torch.from_numpy(np.ascontiguousarray(k.transpose(2, 1, 0))[None]).cuda()
    ch, cw = H + kh - 1, W + kw - 1
    out = fc.conv_bank(spec.tensor, bank, kh, kw, options=fc.Options(crop_h=ch, crop_w=cw))
    torch.cuda.synchronize()
    assert out.shape == (1, cw, ch)
    assert oracle.rel_l2(out[0].cpu().numpy().T, full[:ch, :cw]) < 1e-6
    # clamp pad (src/convolutionFFTkernel.cu:46-76) vs its restatement
    cs = fc.cudaFFTDataClamp(data, kh, kw, kh // 2, kw // 2)
    padded = np.stack([oracle.clamp_pad_data(np.ascontiguousarray(data[:, :, f].T), FW, FH, kw // 2, kh // 2) for f in range(F)])
    ref = np.fft.rfft2(padded.astype(np.float64), axes=(-2, -1))
    got = cs.tensor.cpu().numpy()
    assert oracle.rel_l2(np.stack([got.real, got.imag]), np.stack([ref.real, ref.imag])) < TOL
    # modulateAndNormalize
    a = fc.gpuArray(got.transpose(2, 1, 0).copy())
    b = fc.gpuArray(spec.tensor.cpu().numpy().transpose(2, 1, 0).copy())
    fc.modulateAndNormalize(a, b)
    want = got * spec.tensor.cpu().numpy() / got.size
    res = a.tensor.cpu().numpy()
    assert oracle.rel_l2(np.stack([res.real, res.imag]), np.stack([want.real, want.imag])) < TOL


def test_large_plane_c3_scaled(fc, oracle):
    """BASELINE config 3 structure (single channel, kernel = 1/8 of the image side) at a size the CPU
    oracle finishes in seconds; checked against float64 FFT convolution + direct spot check."""
    rng = np.random.default_rng(3)
    H = W = 1024
    kh = kw = 128
    data = rng.random((H, W, 1), dtype=np.float32)
    ks = [(rng.standard_normal((kh, kw, 1)) / kh).astype(np.float32) for _ in range(2)]
    outs = fc.cudaConvolutionFFT(data, kh, kw, ks)
    FH = fc.computeFFTsize16(H + kh - 1)
    import scipy.fft as sfft
    for k, o in zip(ks, outs):
        ref = sfft.irfft2(sfft.rfft2(data[:, :, 0].astype(np.float64), s=(FH, FH)) *
                          sfft.rfft2(k[:, :, 0].astype(np.float64), s=(FH, FH)), s=(FH, FH))
        assert oracle.rel_l2(o, ref) < TOL


# ------------------------------------------------------------------ path 4: large-plane in-place pipeline
# (kernels_bigplane.cuh).  Forced with Options(path=4) at small sizes so every radix of the in-place plan
# (odd 3..17, 8 / 16 / 32) runs; sizes whose odd part has a prime factor above 17 must fall back to the generic path.
@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 16, 17, 18, 19, 20, 26, 27, 32, 34, 36, 40, 45, 64, 68])
def test_bigplane_every_radix(fc, oracle, m):
    n = 16 * m
    kh, kw = 5, 7
    H, W = n - kh + 1, max(1, 48 - kw + 1 - (m % 3))
    rng = np.random.default_rng(100 + m)
    data = rng.random((H, W, 1), dtype=np.float32)
    kernels = [rng.standard_normal((kh, kw, 1)).astype(np.float32), rng.standard_normal((3, 2, 1)).astype(np.float32)]
    outs = fc.cudaConvolutionFFT(data, kh, kw, kernels, options=fc.Options(path=4))
    _check(oracle, outs, data, kernels, n, fc.computeFFTsize16(W + kw - 1))
    # and with the long axis along w (the strided pass)
    data_t = np.ascontiguousarray(data.transpose(1, 0, 2))
    kernels_t = [np.ascontiguousarray(k.transpose(1, 0, 2)) for k in kernels]
    outs_t = fc.cudaConvolutionFFT(data_t, kw, kh, kernels_t, options=fc.Options(path=4))
    _check(oracle, outs_t, data_t, kernels_t, fc.computeFFTsize16(W + kw - 1), n)


@pytest.mark.parametrize("H,W,F,kh,kw", [(200, 120, 3, 40, 33), (130, 250, 2, 70, 9), (90, 90, 5, 90, 90), (64, 300, 4, 1, 1)])
def test_bigplane_multichannel_unpruned(fc, oracle, H, W, F, kh, kw):
    """Channel sum in the frequency domain (2 lines + 2 accumulator lines per CTA); templates longer than the
    first-stage stride take the unpruned forward stage."""
    rng = np.random.default_rng(H + W + F)
    data = rng.random((H, W, F), dtype=np.float32)
    kernels = [rng.standard_normal((kh, kw, F)).astype(np.float32),
               rng.standard_normal((max(1, kh // 2), max(1, kw - 3), F)).astype(np.float32)]
    outs = fc.cudaConvolutionFFT(data, kh, kw, kernels, options=fc.Options(path=4))
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    _check(oracle, outs, data, kernels, FH, FW)
    outs_g = fc.cudaConvolutionFFT(data, kh, kw, kernels, options=fc.Options(path=1))
    for a, b in zip(outs, outs_g):
        assert oracle.rel_l2(a, b) < TOL


def test_bigplane_options_match_generic(fc, oracle):
    """correlate / crop / oversize (circular wrap) behave exactly as on the generic path."""
    rng = np.random.default_rng(77)
    data = rng.random((100, 70, 2), dtype=np.float32)
    ks = [rng.standard_normal((9, 6, 2)).astype(np.float32), rng.standard_normal((30, 20, 2)).astype(np.float32)]
    for kw_ in (dict(correlate=1), dict(crop_h=60, crop_w=33), dict(crop_h=50, crop_w=40, out_ld=64), dict()):
        a = fc.cudaConvolutionFFT(data, 9, 6, ks, options=fc.Options(path=4, **kw_))      # 2nd kernel exceeds maxK: wraps
        b = fc.cudaConvolutionFFT(data, 9, 6, ks, options=fc.Options(path=1, **kw_))
        for x, y in zip(a, b):
            assert x.shape == y.shape
            if "crop_h" in kw_:          # cropped planes are stored packed ([crop_w][out_ld]) at the start of the buffer
                ch, cw = kw_["crop_h"], kw_["crop_w"]
                ld = kw_.get("out_ld", ch)
                x = x.T.reshape(-1)[: cw * ld].reshape(cw, ld)[:, :ch]
                y = y.T.reshape(-1)[: cw * ld].reshape(cw, ld)[:, :ch]
            assert oracle.rel_l2(x, y) < TOL


def test_bigplane_c3_scaled_auto_and_device_outputs(fc, oracle):
    """BASELINE config 3 structure at 1/4 scale: plane 1152 = 9 * 128 picks path 4 automatically; checked against
    the float64 FFT convolution and against the generic path."""
    import scipy.fft as sfft
    rng = np.random.default_rng(33)
    H = W = 1024
    kh = kw = 128
    data = rng.random((H, W, 1), dtype=np.float32)
    ks = [(rng.standard_normal((kh, kw, 1)) / kh).astype(np.float32) for _ in range(3)]
    ks.append((rng.standard_normal((100, 37, 1)) / 64).astype(np.float32))
    before = fc.launch_count()
    outs = fc.cudaConvolutionFFT(data, kh, kw, ks)
    assert fc.launch_count() > before
    outs_g = fc.cudaConvolutionFFT(data, kh, kw, ks, options=fc.Options(path=1))
    FH = fc.computeFFTsize16(H + kh - 1)
    assert FH == 1152
    for k, o, g in zip(ks, outs, outs_g):
        ref = sfft.irfft2(sfft.rfft2(data[:, :, 0].astype(np.float64), s=(FH, FH)) *
                          sfft.rfft2(k[:, :, 0].astype(np.float64), s=(FH, FH)), s=(FH, FH))
        assert oracle.rel_l2(o, ref) < TOL
        assert oracle.rel_l2(o, g) < TOL
        assert not np.array_equal(o, g)          # really a different pipeline


@pytest.mark.parametrize("FH,FW,F,kh,kw,opts", [
    (64, 1152, 1, 5, 20, {}),                          # size-specialised w pass (pruned first stage), run-time h passes
    (64, 1152, 3, 5, 200, dict(correlate=1)),          # ... channel sum, template wider than the first-stage stride
    (1152, 48, 2, 30, 7, dict(crop_h=1000, crop_w=40)),  # size-specialised C2R along h, cropped store
    (1152, 1152, 1, 128, 128, dict(correlate=1)),
    (2048, 2048, 1, 49, 30, {}),
    (4096, 2048, 2, 300, 17, {}),
])
def test_bigplane_size_specialised_kernels(fc, oracle, FH, FW, F, kh, kw, opts):
    """kernels_bigplane_ct.cuh: the line lengths with a compile-time plan (1152, 2048, 4096, 4608 -- the last one is covered
    at full size by test_c3_full_size_large_plane) against the generic pipeline and the float64 FFT convolution."""
    import scipy.fft as sfft
    rng = np.random.default_rng(FH + FW + F)
    H, W = FH - kh + 1 - 3, FW - kw + 1 - 2
    assert fc.computeFFTsize16(H + kh - 1) == FH and fc.computeFFTsize16(W + kw - 1) == FW
    data = rng.random((H, W, F), dtype=np.float32)
    ks = [(rng.standard_normal((kh, kw, F)) / kh).astype(np.float32),
          (rng.standard_normal((max(1, kh - 2), max(1, kw // 2), F)) / kh).astype(np.float32)]
    a = fc.cudaConvolutionFFT(data, kh, kw, ks, options=fc.Options(path=4, **opts))
    b = fc.cudaConvolutionFFT(data, kh, kw, ks, options=fc.Options(path=1, **opts))
    for x, y in zip(a, b):
        assert x.shape == y.shape
        if "crop_h" in opts:
            ch, cw = opts["crop_h"], opts["crop_w"]
            x = x.T.reshape(-1)[: cw * ch].reshape(cw, ch)
            y = y.T.reshape(-1)[: cw * ch].reshape(cw, ch)
        assert oracle.rel_l2(x, y) < TOL
        assert not np.array_equal(x, y)
    if not opts:
        D = sfft.rfft2(data.astype(np.float64).transpose(2, 0, 1), s=(FH, FW), workers=-1)
        for k, o in zip(ks, a):
            ref = sfft.irfft2(D * sfft.rfft2(k.astype(np.float64).transpose(2, 0, 1), s=(FH, FW), workers=-1), s=(FH, FW), workers=-1).sum(0)
            assert oracle.rel_l2(o, ref) < TOL


@pytest.mark.parametrize("shape", ["c1", "os", "big_template"])
def test_graph_plan_replays_the_whole_schedule(fc, oracle, shape):
    """fftconv_plan_*: data transform + bank convolution captured into one CUDA graph (the persistent schedule that
    replaces the per-stream ConvPlans of src/cudaConvFFTDataStreams.cu:292-328,338-469); buffer contents change between
    executions, another call grows the cached scratch in between (transparent re-capture)."""
    import torch
    rng = np.random.default_rng(91)
    H, W, F, kh, kw, K = {"c1": (64, 8, 5, 10, 4, 10), "os": (120, 100, 6, 12, 9, 130), "big_template": (70, 100, 2, 40, 70, 3)}[shape]
    bank = rng.standard_normal((K, kh, kw, F)).astype(np.float32)
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    d_t = torch.zeros((F, W, H), device="cuda")
    plan = fc.Plan(d_t, b_t, kh, kw)
    assert plan.graph_nodes >= 3 and plan.path == {"c1": 2, "os": 3, "big_template": 1}[shape]
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    for rep in range(3):
        data = rng.random((H, W, F), dtype=np.float32)
        d_t.copy_(torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))))
        before = fc.launch_count()
        out = plan.execute()
        if rep != 1:                                                # (rep 1 re-captures: the scratch grew behind the plan)
            assert fc.launch_count() == before + 1                 # one graph launch
        torch.cuda.synchronize()
        for k in (0, K - 1):
            assert oracle.rel_l2(out[k].cpu().numpy().T, oracle.direct_conv64_c(data, bank[k], FH, FW)) < TOL, (rep, k)
        ref_spec = oracle.fft_data(data, kh, kw)
        got = plan.spec.cpu().numpy()
        assert oracle.rel_l2(np.stack([got.real, got.imag]), np.stack([ref_spec.real, ref_spec.imag])) < TOL
        if rep == 0:                                                # grow the cached scratch behind the plan's back
            big = rng.random((300, 260, F), dtype=np.float32)
            fc.cudaConvolutionFFT(big, kh, kw, [bank[k] for k in range(K)])
    plan.close()
