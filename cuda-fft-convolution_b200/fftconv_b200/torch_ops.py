"""torch.ops.fftconv.* — the device-resident entry points as PyTorch custom operators (SURVEY section 8f, rank 4:
"a Python/torch operator so results stay on device for downstream code").  Importing this module registers

    torch.ops.fftconv.fft_data(data[F,W,H], kh, kw)            -> complex64 [F][FW][CH]     (cudaFFTData, src/cudaFFTData.cu:18-160)
    torch.ops.fftconv.conv_fft_data(spec, bank[K,F,kw,kh])      -> float32  [K][FW][FH]      (cudaConvFFTData, src/cudaConvFFTData.cu:24-306)
    torch.ops.fftconv.convolution_fft(data[N,F,W,H], bank)      -> float32  [N][K][FW][FH]   (cudaConvolutionFFT, src/cudaConvolutionFFT.cu:27-311,
                                                                                             batched: per-bin complex GEMM on tcgen05)

All tensors use the reference's memory order (h contiguous, src/cudaConvFFTData.cuh:26-27), i.e. MATLAB arrays seen
from C.  The operators run on the current CUDA stream through the C ABI (libfftconv.so); there is no CPU
implementation — calling them with CPU tensors raises.  Fake (meta) kernels give shapes to torch.compile / export."""
from __future__ import annotations

import torch

from . import computeFFTsize16, conv_bank, conv_batch, fft_data_device

_LIB = torch.library.Library("fftconv", "DEF")
_LIB.define("fft_data(Tensor data, int kh, int kw) -> Tensor")
_LIB.define("conv_fft_data(Tensor spec, Tensor bank) -> Tensor")
_LIB.define("convolution_fft(Tensor data, Tensor bank) -> Tensor")


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("fftconv operators are CUDA (sm_100a) only: there is no CPU fallback")


def _fft_data(data, kh, kw):
    _need_cuda(data)
    F, W, H = (int(x) for x in data.shape)
    return fft_data_device(data.contiguous().float(), H, W, F, int(kh), int(kw))


def _conv_fft_data(spec, bank):
    _need_cuda(spec, bank)
    K, F, kw, kh = (int(x) for x in bank.shape)
    return conv_bank(spec.contiguous(), bank.contiguous().float(), kh, kw)


def _convolution_fft(data, bank):
    _need_cuda(data, bank)
    return conv_batch(data.contiguous().float(), bank.contiguous().float())


def _plane(H, W, kh, kw):
    return computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)


def _fft_data_fake(data, kh, kw):
    F, W, H = data.shape
    FH, FW = _plane(int(H), int(W), int(kh), int(kw))
    return data.new_empty((F, FW, FH // 2 + 1), dtype=torch.complex64)


def _conv_fft_data_fake(spec, bank):
    F, FW, CH = spec.shape
    return bank.new_empty((bank.shape[0], FW, (int(CH) - 1) * 2), dtype=torch.float32)


def _convolution_fft_fake(data, bank):
    N, F, W, H = data.shape
    K, _, kw, kh = bank.shape
    FH, FW = _plane(int(H), int(W), int(kh), int(kw))
    return data.new_empty((N, K, FW, FH), dtype=torch.float32)


_LIB.impl("fft_data", _fft_data, "CUDA")
_LIB.impl("conv_fft_data", _conv_fft_data, "CUDA")
_LIB.impl("convolution_fft", _convolution_fft, "CUDA")
_LIB.impl("fft_data", _fft_data_fake, "Meta")
_LIB.impl("conv_fft_data", _conv_fft_data_fake, "Meta")
_LIB.impl("convolution_fft", _convolution_fft_fake, "Meta")
