"""CPU restatement of the reference hot path (TEST INFRASTRUCTURE, see __init__).

Array convention (everywhere in this repo's Python): MATLAB arrays ``A(h, w, f)`` are
numpy arrays of shape ``(H, W, F)``; the reference's memory order (column-major, h
contiguous, ``src/cudaConvFFTData.cuh:26-27``) is the C-order array ``[F][W][H]`` =
``np.ascontiguousarray(A.transpose(2, 1, 0))``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

__all__ = [
    "compute_fft_size16", "compute_fft_size_pow2", "i_div_up", "i_align_up",
    "pad_data", "clamp_pad_data", "elementwise_product_and_normalize",
    "modulate_and_normalize", "sum_along_features", "fft_data",
    "conv_fft_data", "convolution_fft", "direct_conv64", "direct_conv64_c",
    "fft_conv_cpu", "rel_l2", "demo_workload", "to_fwh", "from_wh", "build_c_oracle",
]


# --------------------------------------------------------------------------- helpers
def i_div_up(a: int, b: int) -> int:
    """``iDivUp`` — src/cudaConvFFTData.h:36-38."""
    return a // b + 1 if a % b != 0 else a // b


def i_align_up(a: int, b: int) -> int:
    """``iAlignUp`` — src/cudaConvFFTData.h:41-43."""
    return a - a % b + b if a % b != 0 else a


def compute_fft_size16(n: int) -> int:
    """``computeFFTsize16`` — src/cudaConvFFTData.h:96-102 (next multiple of 16)."""
    mod, rem = divmod(int(n), 16)
    return mod * 16 + (16 if rem > 0 else 0)


def compute_fft_size_pow2(n: int) -> int:
    """``computeFFTsize`` (unused by the reference) — src/cudaConvFFTData.h:67-94."""
    n = i_align_up(int(n), 16)
    hi = n.bit_length() - 1
    low = 1 << hi
    return n if low == n else 1 << (hi + 1)


def to_fwh(a: np.ndarray) -> np.ndarray:
    """(H, W, F) MATLAB-shaped array -> C-order [F][W][H] float32 (reference memory)."""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 2:
        a = a[:, :, None]
    return np.ascontiguousarray(a.transpose(2, 1, 0))


def from_wh(p: np.ndarray) -> np.ndarray:
    """[FW][FH] plane (reference memory) -> (FH, FW) MATLAB-shaped view."""
    return p.T


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return float(d / n) if n > 0 else float(d)


# ------------------------------------------------------------- device-kernel restatements
def pad_data(src_fwh: np.ndarray, fft_w: int, fft_h: int) -> np.ndarray:
    """``padData`` (zero pad) — src/cudaConvFFTData.cuh:11-31.

    dst[f][w][h] = src[f][w][h] if w < dataW and h < dataH else 0.  When the source is
    LARGER than the plane the reference simply never reads the excess (bounds test is on
    the destination index), i.e. it truncates — kept here.
    """
    f, w, h = src_fwh.shape
    dst = np.zeros((f, fft_w, fft_h), dtype=np.float32)
    dst[:, : min(w, fft_w), : min(h, fft_h)] = src_fwh[:, :fft_w, :fft_h]
    return dst


def clamp_pad_data(src_wh: np.ndarray, fft_w: int, fft_h: int, kernel_x: int, kernel_y: int) -> np.ndarray:
    """Clamp/wrap ``padData`` of the orphan SDK file — src/convolutionFFTkernel.cu:46-76.

    Index rule per axis (``:63-68``): i < data -> i; data <= i < data+kernelOfs -> data-1
    (replicate far edge); i >= data+kernelOfs -> 0 (wrap to the near edge).  NOTE the SDK
    kernel writes a ROW-major plane ``dst[y*fftW + x]`` (``:70``); we return ``[W][H]``
    (x slow, y fast) to stay in this repo's layout.
    """
    w, h = src_wh.shape

    def idx(n, data, ofs):
        i = np.arange(n)
        return np.where(i < data, i, np.where(i < data + ofs, data - 1, 0))

    ix = idx(fft_w, w, kernel_x)
    iy = idx(fft_h, h, kernel_y)
    return np.ascontiguousarray(src_wh[np.ix_(ix, iy)].astype(np.float32))


def elementwise_product_and_normalize(d: np.ndarray, k: np.ndarray, scale: float) -> np.ndarray:
    """``elementwiseProductAndNormalize`` — src/cudaConvFFTData.cuh:47-67 (no conjugate)."""
    d = d.astype(np.complex64)
    k = k.astype(np.complex64)
    re = np.float32(scale) * (d.real * k.real - d.imag * k.imag)
    im = np.float32(scale) * (d.imag * k.real + d.real * k.imag)
    return (re + 1j * im).astype(np.complex64)


def modulate_and_normalize(a: np.ndarray, b: np.ndarray, data_n: int) -> np.ndarray:
    """``modulateAndNormalize`` — src/convolutionFFTkernel.cu:84-100 (a = a*b/dataN)."""
    return elementwise_product_and_normalize(a, b, 1.0 / float(data_n))


def sum_along_features(per_feature: np.ndarray) -> np.ndarray:
    """``sumAlongFeatures`` — src/cudaConvFFTData.cuh:70-92 (sequential fp32, f ascending)."""
    acc = per_feature[0].astype(np.float32).copy()
    for z in range(1, per_feature.shape[0]):
        acc += per_feature[z].astype(np.float32)
    return acc


# --------------------------------------------------------------- MEX-level restatements
def _rfft2(x: np.ndarray) -> np.ndarray:
    """cufftPlanMany rank-2 R2C, n={FFT_W, FFT_H}, batch F — src/cudaFFTData.cu:137-146.

    Unnormalised forward transform over the last two axes, last axis (h) halved.
    """
    import scipy.fft as sfft
    return sfft.rfft2(x.astype(np.float32), axes=(-2, -1)).astype(np.complex64)


def _irfft2_unnorm(x: np.ndarray, fft_w: int, fft_h: int) -> np.ndarray:
    """cufftExecC2R (unnormalised inverse) — src/cudaConvFFTData.cu:178-184,262."""
    import scipy.fft as sfft
    y = sfft.irfft2(x.astype(np.complex64), s=(fft_w, fft_h), axes=(-2, -1))
    return (y * np.float32(fft_w * fft_h)).astype(np.float32)


def fft_data(data_hwf: np.ndarray, kernel_h: int, kernel_w: int) -> np.ndarray:
    """``cudaFFTData`` — src/cudaFFTData.cu:18-160.

    Returns the spectrum in reference memory order, C-order ``[F][FFT_W][FFT_H/2+1]``
    complex64 (= MATLAB ``[(FFT_H/2+1), FFT_W, F]``, ``:92-94``).
    """
    d = to_fwh(data_hwf)
    _, w, h = d.shape
    fft_h = compute_fft_size16(h + kernel_h - 1)     # :78
    fft_w = compute_fft_size16(w + kernel_w - 1)     # :79
    return _rfft2(pad_data(d, fft_w, fft_h))


def conv_fft_data(spec: np.ndarray, kernels_hwf: Sequence[np.ndarray]) -> List[np.ndarray]:
    """``cudaConvFFTData`` — src/cudaConvFFTData.cu:24-306, hot loop :191-282.

    ``spec`` is ``[F][FFT_W][CFFT_H]`` complex64.  Returns K arrays of MATLAB shape
    ``(FFT_H, FFT_W)`` float32 (the whole padded plane, no crop, no flip).
    """
    f, fft_w, cfft_h = spec.shape
    fft_h = (cfft_h - 1) * 2                                    # :95
    scale = np.float32(1.0) / np.float32(fft_w * fft_h)          # :259
    outs = []
    for ker in kernels_hwf:
        k = to_fwh(ker)
        if k.shape[0] != f or k.shape[1] > fft_w or k.shape[2] > fft_h:      # :229
            raise ValueError("Kernel and Data must have the same number of features and "
                             "kernel size should be smaller than data size")
        kspec = _rfft2(pad_data(k, fft_w, fft_h))                # :233-244
        prod = elementwise_product_and_normalize(spec, kspec, scale)   # :252-260
        per_feature = _irfft2_unnorm(prod, fft_w, fft_h)         # :262 (F inverse FFTs)
        outs.append(from_wh(sum_along_features(per_feature)))    # :265-271
    return outs


def convolution_fft(data_hwf: np.ndarray, max_kh: int, max_kw: int,
                    kernels_hwf: Sequence[np.ndarray]) -> List[np.ndarray]:
    """``cudaConvolutionFFT`` — src/cudaConvolutionFFT.cu:27-311 (= the two calls fused)."""
    return conv_fft_data(fft_data(data_hwf, max_kh, max_kw), kernels_hwf)


# ----------------------------------------------------------------------- ground truth
def direct_conv64(data_hwf: np.ndarray, ker_hwf: np.ndarray, fft_h: int, fft_w: int) -> np.ndarray:
    """float64 ``sum_f conv2(D_f, k_f)`` (demoCudaConvolutionFFT.m:91-96) embedded top-left
    in a zero ``(FFT_H, FFT_W)`` plane; anything beyond the plane wraps (circular), which is
    what an FFT of that size computes."""
    from scipy.signal import convolve2d
    d = np.asarray(data_hwf, dtype=np.float64)
    k = np.asarray(ker_hwf, dtype=np.float64)
    if d.ndim == 2:
        d = d[:, :, None]
    if k.ndim == 2:
        k = k[:, :, None]
    full = np.zeros((d.shape[0] + k.shape[0] - 1, d.shape[1] + k.shape[1] - 1))
    for f in range(d.shape[2]):
        full += convolve2d(d[:, :, f], k[:, :, f], mode="full")
    out = np.zeros((fft_h, fft_w))
    hh, ww = full.shape
    for y0 in range(0, hh, fft_h):
        for x0 in range(0, ww, fft_w):
            blk = full[y0:y0 + fft_h, x0:x0 + fft_w]
            out[: blk.shape[0], : blk.shape[1]] += blk
    return out


_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/direct_conv.c -> oracle/_build/liboracle.so (gcc -O3 -fopenmp)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "direct_conv.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC",
                               "-o", so, src, "-lm"])
    return so


def _clib():
    global _CLIB
    if _CLIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.oracle_direct_conv.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] + \
            [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int]
        lib.oracle_direct_conv.restype = None
        lib.oracle_direct_conv_f32.argtypes = lib.oracle_direct_conv.argtypes
        lib.oracle_direct_conv_f32.restype = None
        _CLIB = lib
    return _CLIB


def direct_conv64_c(data_hwf: np.ndarray, ker_hwf: np.ndarray, fft_h: int, fft_w: int,
                    threads: int = 0, f32: bool = False) -> np.ndarray:
    """Same contract as :func:`direct_conv64`, in C + OpenMP (oracle/direct_conv.c)."""
    d = to_fwh(data_hwf)
    k = to_fwh(ker_hwf)
    f, w, h = d.shape
    _, kw, kh = k.shape
    assert k.shape[0] == f
    if f32:
        out = np.zeros((fft_w, fft_h), dtype=np.float32)
        _clib().oracle_direct_conv_f32(d.ctypes.data, h, w, f, k.ctypes.data, kh, kw, fft_h, fft_w,
                                       out.ctypes.data, threads)
    else:
        out = np.zeros((fft_w, fft_h), dtype=np.float64)
        _clib().oracle_direct_conv(d.ctypes.data, h, w, f, k.ctypes.data, kh, kw, fft_h, fft_w,
                                   out.ctypes.data, threads)
    return from_wh(out)


def fft_conv_cpu(data_hwf: np.ndarray, max_kh: int, max_kw: int, kernels_hwf: Sequence[np.ndarray],
                 workers: int = -1, spec=None):
    """The demo's CPU ``fft2 .* fft2 -> ifft2 -> sum`` path (demoCudaConvolutionFFT.m:78-102)
    at the reference plane size, float32, multi-threaded pocketfft.  Returns (outs, spec)."""
    import scipy.fft as sfft
    d = to_fwh(data_hwf)
    f, w, h = d.shape
    fft_h = compute_fft_size16(h + max_kh - 1)
    fft_w = compute_fft_size16(w + max_kw - 1)
    if spec is None:
        spec = sfft.rfft2(d, s=(fft_w, fft_h), axes=(-2, -1), workers=workers)
    outs = []
    for ker in kernels_hwf:
        k = to_fwh(ker)
        ks = sfft.rfft2(k, s=(fft_w, fft_h), axes=(-2, -1), workers=workers)
        per = sfft.irfft2(spec * ks, s=(fft_w, fft_h), axes=(-2, -1), workers=workers)
        outs.append(from_wh(per.sum(axis=0, dtype=np.float32)))
    return outs, spec


# --------------------------------------------------------------------------- workloads
def demo_workload(seed: int = 1, n_kernels: int = 3):
    """demoCudaConvolutionFFT.m:37-69,108-113 with a fixed seed: 64x8x5 data, 10x4x5 kernels
    with the planted ``reshape(1:40,10,4)`` block, manual flip, kernel2(1)=100, cell{3}==cell{1}."""
    rng = np.random.default_rng(seed)
    n, m, k, cn, cm = 64, 8, 5, 10, 4
    data = rng.random((n, m, k), dtype=np.float32)
    kernel = np.zeros((cn, cm, k), dtype=np.float32)
    kernel[:, :, 0] = np.arange(1, cn * cm + 1, dtype=np.float32).reshape(cm, cn).T   # :52
    for i in range(1, k):
        kernel[:, :, i] = rng.random((cn, cm), dtype=np.float32)
    data[4:4 + cn, 1:1 + cm, 0] = kernel[:, :, 0]          # :58
    data[20:20 + cn, 0:cm, 1] = kernel[:, :, 0]            # :59
    data[0:cn, m - cm:m, k - 1] = kernel[:, :, 0]          # :60
    kernel[:, :, k - 1] = kernel[:, :, 0]                  # :61
    kernel = kernel[::-1, ::-1, :].copy()                  # :67-69 flip
    k2 = kernel.copy()
    k2[0, 0, 0] = 100.0                                    # :110-111 kernel2(1) = 100
    cells = [kernel, k2, kernel]
    while len(cells) < n_kernels:
        cells.append(rng.random((cn, cm, k), dtype=np.float32))
    return data, cells[:n_kernels], cn, cm
