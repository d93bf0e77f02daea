"""The overlap-save path works from raw data; a two-call caller hands back the spectrum of cudaFFTData.  The library keeps
the raw data of the last fftconv_fft_data call and a hash of the spectrum it wrote, and re-hashes what the convolution
receives ON THE DEVICE (csrc/fftconv.cu, Ctx::SpecCache).  These tests make sure a spectrum that no longer matches that
provenance (modified in place, replaced at the same address, produced for other data) is never served from the kept raw
data, and that the eager tile transform of steady-state loops stays correct.  Reference: cudaFFTData -> cudaConvFFTData
(src/cudaFFTData.cu:128-147, src/cudaConvFFTData.cu:191-282); tolerance rel-L2 <= 1e-5 against float64 direct convolution."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5
H, W, F, KH, KW, K = 96, 80, 6, 9, 12, 70          # K >= 64: overlap-save / tcgen05 path


def _mk(seed):
    import torch
    rng = np.random.default_rng(seed)
    data = rng.random((H, W, F), dtype=np.float32)
    bank = (rng.standard_normal((K, KH, KW, F)) * 0.05).astype(np.float32)
    d_t = torch.from_numpy(np.ascontiguousarray(data.transpose(2, 1, 0))).cuda()        # [F][W][H]
    b_t = torch.from_numpy(np.ascontiguousarray(bank.transpose(0, 3, 2, 1))).cuda()     # [K][F][kw][kh]
    return data, bank, d_t, b_t


def _check(fc, oracle, out_t, data, bank, scale=1.0):
    FH, FW = fc.computeFFTsize16(H + KH - 1), fc.computeFFTsize16(W + KW - 1)
    got = out_t.cpu().numpy().transpose(0, 2, 1)
    for k in (0, K // 2, K - 1):
        ref = scale * oracle.direct_conv64_c(data, bank[k], FH, FW)
        assert oracle.rel_l2(got[k], ref) < TOL, k


def _ran(fc):
    return set(fc.profile_read())


def test_cached_raw_data_serves_the_unmodified_spectrum(fc, oracle):
    import torch
    data, bank, d_t, b_t = _mk(1)
    spec = fc.fft_data_device(d_t, H, W, F, KH, KW)
    fc.profile(True); fc.profile_read()
    out = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    assert "os_gemm" in _ran(fc)
    fc.profile(False)
    _check(fc, oracle, out, data, bank)


def test_spectrum_modified_in_place_is_not_served_from_the_cache(fc, oracle):
    import torch
    data, bank, d_t, b_t = _mk(2)
    spec = fc.fft_data_device(d_t, H, W, F, KH, KW)
    spec.mul_(2.0)                                     # same address, different contents
    out = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    _check(fc, oracle, out, data, bank, scale=2.0)


def test_other_spectrum_at_the_same_address(fc, oracle):
    import torch
    data1, bank, d1, b_t = _mk(3)
    data2, _, d2, _ = _mk(4)
    spec_a = fc.fft_data_device(d2, H, W, F, KH, KW).clone()     # spectrum of data2, kept aside
    spec = fc.fft_data_device(d1, H, W, F, KH, KW)               # provenance now says: data1 at `spec`
    spec.copy_(spec_a)                                           # ... but the buffer holds data2's spectrum
    out = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    _check(fc, oracle, out, data2, bank)


def test_steady_state_loop_with_changing_data(fc, oracle):
    """fft_data / conv in a loop (what bench.py times): from the second iteration on the tile spectra are computed next to
    the forward transform; every iteration must see ITS data."""
    import torch
    _, bank, _, b_t = _mk(5)
    spec = None
    for it in range(4):
        data, _, d_t, _ = _mk(10 + it)
        spec = fc.fft_data_device(d_t, H, W, F, KH, KW, spec_t=spec)
        out = fc.conv_bank(spec, b_t, KH, KW)
        torch.cuda.synchronize()
        _check(fc, oracle, out, data, bank)


def test_eager_tiles_survive_an_unrelated_call_in_between(fc, oracle):
    import torch
    data, bank, d_t, b_t = _mk(6)
    other, _, o_t, _ = _mk(7)
    spec = fc.fft_data_device(d_t, H, W, F, KH, KW)
    fc.conv_bank(spec, b_t, KH, KW)                              # arms the eager transform for this geometry
    spec = fc.fft_data_device(d_t, H, W, F, KH, KW, spec_t=spec)     # tiles of `data` now sit in the shared scratch
    o_np = np.ascontiguousarray(other)
    ks = [np.ascontiguousarray(bank[k]) for k in range(K)]
    outs_other = fc.cudaConvolutionFFT(o_np, KH, KW, ks)         # one-shot call on other data overwrites that scratch
    out = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    _check(fc, oracle, out, data, bank)
    FH, FW = fc.computeFFTsize16(H + KH - 1), fc.computeFFTsize16(W + KW - 1)
    assert oracle.rel_l2(outs_other[3], oracle.direct_conv64_c(other, bank[3], FH, FW)) < TOL


def test_second_convolution_with_the_same_spectrum(fc, oracle):
    import torch
    data, bank, d_t, b_t = _mk(8)
    spec = fc.fft_data_device(d_t, H, W, F, KH, KW)
    o1 = fc.conv_bank(spec, b_t, KH, KW).clone()
    o2 = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    _check(fc, oracle, o1, data, bank)
    _check(fc, oracle, o2, data, bank)


def _bind(fc, spec, d_t):
    import torch
    rc = fc.lib().fftconv_spectrum_bind_raw(spec.data_ptr(), d_t.data_ptr(), H, W, F, KH, KW, 0,
                                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fc.last_error()


def test_bind_raw_spectrum_assembled_from_channel_slices(fc, oracle):
    """fftconv_spectrum_bind_raw: a spectrum assembled from channel slices (what the multi-GPU all-gather delivers) is
    declared as the transform of the replicated raw data; the convolution then tiles the raw data."""
    import torch
    data, bank, d_t, b_t = _mk(20)
    FH, FW = fc.computeFFTsize16(H + KH - 1), fc.computeFFTsize16(W + KW - 1)
    spec = torch.empty((F, FW, FH // 2 + 1), dtype=torch.complex64, device="cuda")
    for f0, f1 in ((0, 2), (2, 3), (3, F)):
        fc.fft_data_device(d_t[f0:f1], H, W, f1 - f0, KH, KW, spec_t=spec[f0:f1])
    _bind(fc, spec, d_t)
    d_t.zero_()                                        # the library kept its own copy of the raw data
    out = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    _check(fc, oracle, out, data, bank)


def test_bind_raw_then_modified_spectrum_is_inverted_not_trusted(fc, oracle):
    import torch
    data, bank, d_t, b_t = _mk(21)
    spec = fc.fft_data_device(d_t, H, W, F, KH, KW).clone()
    _bind(fc, spec, d_t)
    spec.mul_(-3.0)                                    # changed after the declaration: the device-side hash no longer matches
    out = fc.conv_bank(spec, b_t, KH, KW)
    torch.cuda.synchronize()
    _check(fc, oracle, out, data, bank, scale=-3.0)


def test_bind_raw_rejects_bad_arguments(fc):
    assert fc.lib().fftconv_spectrum_bind_raw(None, None, H, W, F, KH, KW, 0, None) != 0
