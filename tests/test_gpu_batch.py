"""GPU parity of the batched entry point (BASELINE config "batched": N images x one bank, the channel
reduction done as a per-frequency-bin complex GEMM on the tensor cores): fftconv_conv_batch against the
float64 direct convolution and against the single-image entry point.  Tolerance rel-L2 <= 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _case(fc, oracle, N, H, W, F, kh, kw, K, seed):
    import torch
    rng = np.random.default_rng(seed)
    data = rng.random((N, H, W, F), dtype=np.float32)                      # MATLAB order per image
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.05).astype(np.float32)
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(0, 3, 2, 1))).cuda()     # [N][F][W][H]
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()     # [K][F][kw][kh]
    before = fc.launch_count()
    out = fc.conv_batch(d_t, b_t)
    torch.cuda.synchronize()
    assert fc.launch_count() > before
    assert tuple(out.shape) == (N, K, FW, FH)
    got = out.cpu().numpy().transpose(0, 1, 3, 2)                           # [N][K][FH][FW]
    for n in range(N):
        for k in sorted({0, K // 2, K - 1}):
            ref = oracle.direct_conv64_c(data[n], bank[k], FH, FW)
            assert oracle.rel_l2(got[n, k], ref) < TOL, (n, k)
    return data, bank, got


def test_batch_small_kernels(fc, oracle):
    _case(fc, oracle, N=3, H=100, W=90, F=6, kh=9, kw=12, K=37, seed=41)


def test_batch_32x32_kernels_c4_shape_scaled(fc, oracle):
    """config 4 scaled down: 5 images 96x80x32 x 130 kernels 32x32x32 (two 128-row MMA blocks, NF = 2)."""
    _case(fc, oracle, N=5, H=96, W=80, F=32, kh=32, kw=32, K=130, seed=42)


def test_batch_matches_single_image_calls(fc, oracle):
    data, bank, got = _case(fc, oracle, N=2, H=64, W=48, F=5, kh=7, kw=5, K=70, seed=43)
    for n in range(2):
        outs = fc.cudaConvolutionFFT(data[n], 7, 5, [bank[k] for k in range(70)])
        for k in (0, 33, 69):
            assert oracle.rel_l2(got[n, k], outs[k]) < 2e-6


def test_batch_many_tile_blocks(fc, oracle):
    """more than 40 tiles per group -> several B operand blocks (NNB > 1), groups that straddle images."""
    _case(fc, oracle, N=4, H=200, W=150, F=3, kh=5, kw=6, K=9, seed=44)


def test_batch_several_groups_share_one_bank_transform(fc, oracle):
    """More images than one group holds (1280 tiles): the bank is transformed once for the whole call and every
    group runs against the resident template spectra."""
    import torch
    N, H, W, F, kh, kw, K = 11, 300, 300, 2, 32, 32, 130
    rng = np.random.default_rng(46)
    data = rng.random((N, H, W, F), dtype=np.float32)
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.05).astype(np.float32)
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(0, 3, 2, 1))).cuda()
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    out = fc.conv_batch(d_t, b_t)
    torch.cuda.synchronize()
    for n in (0, 9, 10):                       # first group, last image of it, second group
        for k in (0, 127, 129):
            got = out[n, k].cpu().numpy().T
            assert oracle.rel_l2(got, oracle.direct_conv64_c(data[n], bank[k], FH, FW)) < TOL, (n, k)


def test_batch_falls_back_for_large_kernels(fc, oracle):
    """kernels beyond 32x32 are outside the overlap-save path: the call loops over the images."""
    _case(fc, oracle, N=2, H=80, W=70, F=2, kh=40, kw=33, K=3, seed=45)


@pytest.mark.parametrize("feed", ["raw", "spectra", "mixed"])
def test_conv_pyramid_one_call_equals_level_by_level(fc, oracle, feed):
    """fftconv_conv_pyramid: the ten level sizes of BASELINE config 5 (F = 31, 16 x 16 templates) against one bank of 150
    templates in ONE call (tiles of all levels share the per-bin GEMM; 158 tiles = 4 tile blocks that straddle level
    boundaries) -- against the per-level two-call sequence cudaFFTData -> cudaConvFFTData, and against float64."""
    import torch
    from fftconv_b200.pyramid import pyramid_sides, level_plane
    rng = np.random.default_rng(57)
    F, kh, kw, K = 31, 16, 16, 150
    sides = pyramid_sides()
    levels = [(rng.random((s, s + (l % 3), F), dtype=np.float32) * 0.2).astype(np.float32) for l, s in enumerate(sides)]   # H = s, W = s + l%3
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.05).astype(np.float32)
    lt = [torch.from_numpy(np.ascontiguousarray(lv.transpose(2, 1, 0))).cuda() for lv in levels]
    bt = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    shapes = [(lv.shape[0], lv.shape[1]) for lv in levels]
    specs = [fc.fft_data_device(t, H, W, F, kh, kw).clone() for t, (H, W) in zip(lt, shapes)]
    ref_outs = [fc.conv_bank(sp, bt, kh, kw).clone() for sp in specs]
    fc.profile(True); fc.profile_read(True)
    if feed == "raw":
        outs = fc.conv_pyramid(lt, bt, kh, kw)
    elif feed == "spectra":
        outs = fc.conv_pyramid(None, bt, kh, kw, specs=specs, shapes=shapes)
    else:
        outs = fc.conv_pyramid([t if l % 2 else None for l, t in enumerate(lt)], bt, kh, kw, specs=specs, shapes=shapes)
    torch.cuda.synchronize()
    prof = fc.profile_read(True)
    fc.profile(False)
    assert prof["os_gemm"][1] == 1 and prof["os_inverse"][1] == 1 and prof["os_kern_fft(templates)"][1] == 1, prof
    for l, (H, W) in enumerate(shapes):
        FH, FW = level_plane(H, W, kh, kw)
        got, want = outs[l].cpu().numpy(), ref_outs[l].cpu().numpy()
        assert got.shape == (K, FW, FH)
        assert oracle.rel_l2(got, want) < TOL, l
        for k in (0, 127, 128, K - 1):
            ref = oracle.direct_conv64_c(levels[l], bank[k], FH, FW)
            assert oracle.rel_l2(got[k].T, ref) < TOL, (l, k)


@pytest.mark.parametrize("L,K,F,kh,kw", [(1, 1, 1, 1, 1), (2, 3, 1, 5, 32), (3, 129, 2, 32, 3)])
def test_conv_pyramid_edge_shapes(fc, oracle, L, K, F, kh, kw):
    """one level / one template / one channel, 1 x 1 and 32-wide templates, a bank one past a block of 128, ragged level sizes."""
    import torch
    rng = np.random.default_rng(60 + K)
    shapes = [(33 + 17 * l, 70 - 9 * l) for l in range(L)]
    levels = [rng.random((H, W, F), dtype=np.float32) for (H, W) in shapes]
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.1).astype(np.float32)
    lt = [torch.from_numpy(np.ascontiguousarray(lv.transpose(2, 1, 0))).cuda() for lv in levels]
    bt = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    outs = fc.conv_pyramid(lt, bt, kh, kw)
    torch.cuda.synchronize()
    for l, (H, W) in enumerate(shapes):
        FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
        got = outs[l].cpu().numpy()
        assert got.shape == (K, FW, FH)
        for k in sorted({0, K // 2, K - 1}):
            ref = oracle.direct_conv64_c(levels[l], bank[k], FH, FW)
            assert oracle.rel_l2(got[k].T, ref) < TOL, (l, k)


@pytest.mark.parametrize("fallback", [False, True])
def test_conv_pyramid_waits_for_levels_still_in_flight(fc, oracle, fallback):
    """data_ready (fftconv_spectrum_ready_event): the levels are filled on ANOTHER stream behind a long-running kernel -- the
    NCCL broadcast of the packed pyramid in the multi-GPU schedule -- and only the data side of the call waits for them.
    The level buffers hold NaN until that stream writes them, so a call that did not wait cannot pass.
    fallback: 40-wide templates, which take the level-by-level route (there the whole call waits)."""
    import torch
    rng = np.random.default_rng(61)
    F, K = 5, 130
    kh, kw = (40, 9) if fallback else (16, 12)
    shapes = [(90, 75), (64, 64), (47, 53)]
    levels = [rng.random((H, W, F), dtype=np.float32) for (H, W) in shapes]
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.1).astype(np.float32)
    src = [torch.from_numpy(np.ascontiguousarray(lv.transpose(2, 1, 0))).cuda() for lv in levels]
    bt = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    lt = [torch.full_like(t, float("nan")) for t in src]
    side = torch.cuda.Stream()
    busy = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(40):                                   # some milliseconds of work in front of the fill
            busy.mul_(1.0001)
        for d, t in zip(lt, src):
            d.copy_(t, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(side)
    outs = fc.conv_pyramid(lt, bt, kh, kw, data_ready=ready)
    torch.cuda.synchronize()
    for l, (H, W) in enumerate(shapes):
        FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
        got = outs[l].cpu().numpy()
        assert not np.isnan(got).any(), l
        for k in (0, K - 1):
            assert oracle.rel_l2(got[k].T, oracle.direct_conv64_c(levels[l], bank[k], FH, FW)) < TOL, (l, k)
    # the event was one-shot: the next call does not wait for (or touch) it
    outs2 = fc.conv_pyramid(src, bt, kh, kw)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(outs, outs2))


def test_conv_pyramid_too_many_tiles_goes_level_by_level(fc, oracle):
    """more than 1280 overlap-save tiles (the scratch bound of one GEMM problem): the levels are convolved one by one."""
    import torch
    rng = np.random.default_rng(59)
    F, K, kh, kw = 2, 64, 16, 16
    sides = (800, 790, 780, 770, 760)                          # 17 x 17 tiles each
    lt = [torch.from_numpy(rng.random((F, s, s), dtype=np.float32)).cuda() for s in sides]
    bt = torch.from_numpy((rng.standard_normal((K, F, kw, kh)) * 0.05).astype(np.float32)).cuda()
    fc.profile(True); fc.profile_read(True)
    outs = fc.conv_pyramid(lt, bt, kh, kw)
    torch.cuda.synchronize()
    prof = fc.profile_read(True)
    fc.profile(False)
    assert prof["os_gemm"][1] == len(sides), prof              # one GEMM launch per level
    for l in (0, 4):
        want = fc.convolution_fft_device(lt[l], bt)
        assert oracle.rel_l2(outs[l].cpu().numpy(), want.cpu().numpy()) < TOL


def test_conv_pyramid_falls_back_level_by_level(fc, oracle):
    """templates above 32 x 32 (no overlap-save tiles) and correlation mode: the call is L single-image calls."""
    import torch
    rng = np.random.default_rng(58)
    F, K = 2, 3
    for kh, kw, opts in ((40, 9, None), (7, 7, fc.Options(correlate=1))):
        levels = [rng.random((s, s + 5, F), dtype=np.float32) for s in (70, 50)]
        bank = rng.standard_normal((K, kh, kw, F)).astype(np.float32)
        lt = [torch.from_numpy(np.ascontiguousarray(lv.transpose(2, 1, 0))).cuda() for lv in levels]
        bt = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
        outs = fc.conv_pyramid(lt, bt, kh, kw, options=opts)
        torch.cuda.synchronize()
        for l, lv in enumerate(levels):
            H, W = lv.shape[:2]
            spec = fc.fft_data_device(lt[l], H, W, F, kh, kw)
            want = fc.conv_bank(spec, bt, kh, kw, options=opts)
            assert oracle.rel_l2(outs[l].cpu().numpy(), want.cpu().numpy()) < TOL


def test_pyramid_schedule_single_rank(fc, oracle):
    """config 5 scaled down: 4-level pyramid x 80 templates through the sharded schedule (world = 1)."""
    import torch
    from fftconv_b200.pyramid import pyramid_convolution_cuda, pyramid_sides, level_plane
    rng = np.random.default_rng(51)
    F, kh, kw, K = 7, 6, 6, 80
    sides = pyramid_sides(60, 4, 2)
    levels = [(rng.random((s, s, F), dtype=np.float32) * 0.2).astype(np.float32) for s in sides]
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.05).astype(np.float32)
    lt = [torch.from_numpy(np.ascontiguousarray(lv.transpose(2, 1, 0))).cuda() for lv in levels]
    bt = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    b, e, outs = pyramid_convolution_cuda(lt, [(s, s, F) for s in sides], bt, kh, kw)
    torch.cuda.synchronize()
    assert (b, e) == (0, K)
    for l, s in enumerate(sides):
        FH, FW = level_plane(s, s, kh, kw)
        got = outs[l].cpu().numpy()
        assert got.shape == (K, FW, FH)
        for k in (0, 41, K - 1):
            ref = oracle.direct_conv64_c(levels[l], bank[k], FH, FW)
            assert oracle.rel_l2(got[k].T, ref) < TOL


def test_prepared_bank_matches_one_shot_and_oracle(fc, oracle):
    """fftconv_bank_*: the bank is transformed once, then serves images of different sizes (pyramid levels)."""
    rng = np.random.default_rng(61)
    F, K = 6, 150
    ks = []
    for k in range(K):
        a, b = (13, 16) if k % 4 == 0 else (int(rng.integers(3, 14)), int(rng.integers(2, 17)))
        ks.append((rng.standard_normal((a, b, F)) * 0.05).astype(np.float32))
    bank = fc.Bank(ks)
    assert (bank.K, bank.F, bank.maxKH, bank.maxKW) == (K, F, 13, 16) and bank.bytes > 0
    for (H, W) in ((90, 70), (41, 133)):
        data = (rng.random((H, W, F), dtype=np.float32) * 0.2).astype(np.float32)
        FH, FW = bank.plane(H, W)
        before = fc.launch_count()
        outs = bank.conv(data)
        assert fc.launch_count() - before >= 3 and len(outs) == K
        one = fc.cudaConvolutionFFT(data, 13, 16, ks, options=fc.Options(path=3))
        for k in (0, 1, 77, K - 1):
            assert outs[k].shape == (FH, FW)
            assert np.array_equal(outs[k], one[k])                      # same kernels, same arithmetic
            assert oracle.rel_l2(outs[k], oracle.direct_conv64_c(data, ks[k], FH, FW)) < TOL
    bank.close()


def test_prepared_bank_device_path_and_errors(fc, oracle):
    import torch
    rng = np.random.default_rng(62)
    F, K, kh, kw = 4, 40, 8, 8
    ks = [(rng.standard_normal((kh, kw, F)) * 0.1).astype(np.float32) for _ in range(K)]
    bank = fc.Bank(ks)
    data = rng.random((70, 50, F), dtype=np.float32)
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()
    out = bank.conv_device(d_t)
    torch.cuda.synchronize()
    FH, FW = bank.plane(70, 50)
    got = out.cpu().numpy()
    for k in (0, 19, 39):
        assert oracle.rel_l2(got[k].T, oracle.direct_conv64_c(data, ks[k], FH, FW)) < TOL
    with pytest.raises(fc.FFTConvError):
        bank.conv(rng.random((70, 50, F + 1), dtype=np.float32))        # feature mismatch (src/cudaConvFFTData.cu:229-230)
    with pytest.raises(fc.FFTConvError):
        fc.Bank([np.zeros((40, 8, F), np.float32)])                     # beyond 32 x 32: not a prepared-bank shape
    bank.close()


def test_fused_peak_reduction_matches_planes(fc, oracle):
    """fftconv_bank_conv_max: maximum + position of every template's full convolution, fused into the inverse store."""
    rng = np.random.default_rng(71)
    F, K, H, W = 5, 140, 83, 61
    ks = []
    for k in range(K):
        a, b = (int(rng.integers(2, 17)), int(rng.integers(2, 13)))
        ks.append((rng.standard_normal((a, b, F))).astype(np.float32))
    data = rng.standard_normal((H, W, F)).astype(np.float32)
    bank = fc.Bank(ks)
    planes = bank.conv(data)
    before = fc.launch_count()
    val, ys, xs = bank.conv_max(data)
    assert fc.launch_count() > before
    for k in range(K):
        kh, kw, _ = ks[k].shape
        blk = planes[k][:H + kh - 1, :W + kw - 1]
        x, y = np.unravel_index(np.argmax(blk.T), blk.T.shape)             # first maximum, smallest x then y
        assert val[k] == blk.max(), k                                       # same arithmetic as the plane: bit-equal
        assert (ys[k], xs[k]) == (y, x), k
    # and against the float64 direct convolution
    for k in (0, 70, K - 1):
        kh, kw, _ = ks[k].shape
        ref = oracle.direct_conv64_c(data, ks[k], *bank.plane(H, W))[:H + kh - 1, :W + kw - 1]
        assert abs(val[k] - ref.max()) <= 1e-5 * np.abs(ref).max() * 10
    bank.close()


@pytest.mark.parametrize("path", [2, 3])
def test_correlation_mode_per_template_sizes(fc, oracle, path):
    """correlate = 1 (complexConjMulAndScale, src/cudaConvFFTData.cuh:42-45,63): circular cross-correlation on the
    plane = convolution with the flipped template, shifted by (kh-1, kw-1) per template."""
    rng = np.random.default_rng(81)
    H, W, F, K = 75, 58, 4, 70
    data = rng.random((H, W, F), dtype=np.float32)
    ks = [rng.standard_normal((int(rng.integers(1, 15)), int(rng.integers(1, 13)), F)).astype(np.float32) for _ in range(K)]
    ks[0] = rng.standard_normal((14, 12, F)).astype(np.float32)
    corr = fc.cudaConvolutionFFT(data, 14, 12, ks, options=fc.Options(correlate=1, path=path))
    FH, FW = fc.computeFFTsize16(H + 13), fc.computeFFTsize16(W + 11)
    for k in (0, 1, 35, K - 1):
        kh, kw, _ = ks[k].shape
        flip = oracle.direct_conv64_c(data, np.ascontiguousarray(ks[k][::-1, ::-1, :]), FH, FW)
        want = np.roll(flip, (-(kh - 1), -(kw - 1)), axis=(0, 1))
        assert oracle.rel_l2(corr[k], want) < TOL, (path, k)


def test_spectrum_ready_event_orders_only_the_data_side(fc, oracle):
    """fftconv_spectrum_ready_event: the spectrum is filled on ANOTHER stream (a copy standing in for the NCCL broadcast);
    the call stream never waits for it explicitly — only the event handed to the library orders the data-side work."""
    import torch
    rng = np.random.default_rng(48)
    H = W = 128; F = 4; kh = kw = 9; K = 70
    data = rng.random((H, W, F), dtype=np.float32)
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.1).astype(np.float32)
    spec_src = fc.cudaFFTData(data, kh, kw).tensor.clone()
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for path in (0, 1, 3):
        spec = torch.zeros_like(spec_src)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            torch.cuda._sleep(20_000_000)                  # the "collective" is late
            spec.copy_(spec_src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        out = fc.conv_bank(spec, b_t, kh, kw, options=fc.Options(path=path), spectrum_ready=ev)
        torch.cuda.synchronize()
        FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
        for k in (0, K - 1):
            assert oracle.rel_l2(out[k].cpu().numpy().T, oracle.direct_conv64_c(data, bank[k], FH, FW)) < TOL, (path, k)


def test_torch_ops_match_the_mirror(fc, oracle):
    import torch
    import fftconv_b200.torch_ops  # noqa: F401
    rng = np.random.default_rng(49)
    H, W, F, kh, kw, K = 70, 50, 3, 9, 7, 5
    data = rng.random((H, W, F), dtype=np.float32)
    bank = rng.standard_normal((K, kh, kw, F)).astype(np.float32)
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    spec = torch.ops.fftconv.fft_data(d_t, kh, kw)
    out = torch.ops.fftconv.conv_fft_data(spec, b_t)
    out2 = torch.ops.fftconv.convolution_fft(d_t[None], b_t)
    torch.cuda.synchronize()
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    for k in range(K):
        ref = oracle.direct_conv64_c(data, bank[k], FH, FW)
        assert oracle.rel_l2(out[k].cpu().numpy().T, ref) < TOL
        assert oracle.rel_l2(out2[0, k].cpu().numpy().T, ref) < TOL


def test_batch_tail_group_of_one_image_small_bank_declared_max_larger(fc, oracle):
    """A tail group that holds ONE image, a bank below the automatic overlap-save threshold (K < 64) and a declared
    maximum larger than the kernels (so no per-call bank is built): the group still has no compat spectrum, so it must be
    served by the overlap-save path (used to dereference a null spectrum)."""
    import ctypes
    import torch
    N, H, W, F, kh, kw, K = 11, 600, 600, 2, 5, 5, 9              # 121 tiles per image -> 10 images per group, tail of 1
    rng = np.random.default_rng(47)
    data = rng.random((N, H, W, F), dtype=np.float32)
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.1).astype(np.float32)
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(0, 3, 2, 1))).cuda()
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    mk = 8                                                          # declared maximum 8 x 8 > 5 x 5
    FH, FW = fc.computeFFTsize16(H + mk - 1), fc.computeFFTsize16(W + mk - 1)
    out = torch.empty((N, K, FW, FH), device="cuda")
    kp = (ctypes.c_void_p * K)(*[b_t.data_ptr() + 4 * k * F * kw * kh for k in range(K)])
    op = (ctypes.c_void_p * (N * K))(*[out.data_ptr() + 4 * FW * FH * i for i in range(N * K)])
    khs = (ctypes.c_int * K)(*([kh] * K))
    ond = (ctypes.c_ubyte * K)(*([1] * K))
    rc = fc.lib().fftconv_conv_batch(d_t.data_ptr(), 1, N, H, W, F, mk, mk, K, kp, khs, khs, None, ond, op, 1, None, 0,
                                     torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fc.last_error()
    torch.cuda.synchronize()
    for n, k in ((0, 0), (9, 4), (10, 8)):
        ref = oracle.direct_conv64_c(data[n], bank[k], FH, FW)
        assert oracle.rel_l2(out[n, k].cpu().numpy().T, ref) < TOL, (n, k)


def test_calls_on_two_streams_and_two_host_threads_do_not_race(fc, oracle):
    """The cached scratch is shared by all streams of a device: device-output calls issued back to back on two streams,
    and host-output calls from two host threads, must each see their own image."""
    import threading
    import torch
    rng = np.random.default_rng(52)
    H = W = 120; F = 4; kh = kw = 9; K = 70
    datas = [rng.random((H, W, F), dtype=np.float32) for _ in range(2)]
    bank = (rng.standard_normal((K, kh, kw, F)) * 0.1).astype(np.float32)
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()
    FH, FW = fc.computeFFTsize16(H + kh - 1), fc.computeFFTsize16(W + kw - 1)
    refs = [oracle.direct_conv64_c(d, bank[K - 1], FH, FW) for d in datas]
    # two streams, device outputs, no host synchronisation in between
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    d_ts = [torch.from_numpy(np.ascontiguousarray(d.transpose(2, 1, 0))).cuda() for d in datas]
    outs = [torch.empty((K, FW, FH), device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                spec = fc.fft_data_device(d_ts[i], H, W, F, kh, kw, stream=streams[i])
                fc.conv_bank(spec, b_t, kh, kw, outs[i], stream=streams[i])
    torch.cuda.synchronize()
    for i in range(2):
        assert oracle.rel_l2(outs[i][K - 1].cpu().numpy().T, refs[i]) < TOL, i
    # two host threads through the one-shot entry point (host buffers)
    res = [None, None]

    def work(i):
        for _ in range(4):
            res[i] = fc.cudaConvolutionFFT(datas[i], kh, kw, [bank[k] for k in range(K)])

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(2):
        assert oracle.rel_l2(res[i][K - 1], refs[i]) < TOL, i


def _responses(planes, ks, H, W, bias):
    out = []
    for k, p in enumerate(planes):
        kh, kw, _ = ks[k].shape
        out.append((p[:H + kh - 1, :W + kw - 1] + np.float32(bias[k])).astype(np.float32))
    return out


def test_fused_threshold_detection_with_bias_matches_planes(fc, oracle):
    """fftconv_bank_conv_detect: responses conv + bias >= threshold, fused into the inverse store; compared
    bit-for-bit with thresholding the planes of the same bank on the host."""
    rng = np.random.default_rng(72)
    F, K, H, W = 5, 140, 83, 61
    ks = [rng.standard_normal((int(rng.integers(2, 17)), int(rng.integers(2, 13)), F)).astype(np.float32) for _ in range(K)]
    data = rng.standard_normal((H, W, F)).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32) * 3
    bank = fc.Bank(ks)
    resp = _responses(bank.conv(data), ks, H, W, bias)
    thr = float(np.quantile(np.concatenate([r.ravel() for r in resp]), 0.9995))
    M = 32
    counts, val, ys, xs = bank.detect(data, thr, bias=bias, max_per_template=M)
    assert counts.sum() > K // 2                                            # the threshold is meaningful
    for k in range(K):
        r = resp[k]
        yy, xx = np.nonzero(r >= np.float32(thr))
        assert counts[k] == yy.size, k
        order = sorted(zip(-r[yy, xx].astype(np.float64), xx, yy))          # value descending, then smallest x, then y
        n = min(M, len(order))
        for i in range(n):
            assert val[k, i] == np.float32(-order[i][0]) and (xs[k, i], ys[k, i]) == (order[i][1], order[i][2]), (k, i)
        assert np.all(np.isneginf(val[k, n:])) and np.all(ys[k, n:] == -1)
    bank.close()


@pytest.mark.parametrize("k", [1, 5, 16])
def test_fused_topk_matches_planes(fc, oracle, k):
    """fftconv_bank_conv_topk: exact top-k of every template (candidate pass + threshold pass over the same spectra)."""
    rng = np.random.default_rng(73)
    F, K, H, W = 4, 96, 150, 97
    ks = [rng.standard_normal((int(rng.integers(3, 17)), int(rng.integers(3, 17)), F)).astype(np.float32) for _ in range(K)]
    data = rng.standard_normal((H, W, F)).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32)
    bank = fc.Bank(ks)
    resp = _responses(bank.conv(data), ks, H, W, bias)
    val, ys, xs = bank.topk(data, k, bias=bias)
    for t in range(K):
        r = resp[t]
        flat = sorted(zip(-r.ravel().astype(np.float64), np.tile(np.arange(r.shape[1]), r.shape[0]), np.repeat(np.arange(r.shape[0]), r.shape[1])))[:k]
        for i in range(k):
            assert val[t, i] == np.float32(-flat[i][0]), (t, i)
            assert (xs[t, i], ys[t, i]) == (flat[i][1], flat[i][2]), (t, i)
    v1, y1, x1 = bank.conv_max(data)
    if k == 1:                                                             # bias-free maximum agrees with conv_max
        v0, y0, x0 = bank.topk(data, 1)
        assert np.array_equal(v0[:, 0], v1) and np.array_equal(y0[:, 0], y1) and np.array_equal(x0[:, 0], x1)
    bank.close()
