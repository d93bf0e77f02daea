"""Host-side mirror of the reference's MEX interface over the C ABI of libfftconv.so.

Same function names, argument order, argument meaning and error behaviour as the MEX entry
points of chrischoy/CUDA-FFT-Convolution:

    cudaFFTData(data, kernelH, kernelW)                       src/cudaFFTData.cu:18-160
    cudaConvFFTData(fftData, kernelCell[, threads])           src/cudaConvFFTData.cu:24-306
    cudaConvolutionFFT(data, maxKH, maxKW, kernelCell[, threads[, gpuId]])
                                                              src/cudaConvolutionFFT.cu:27-311
    cudaConvFFTDataStreams(fftData, kernelCell[, threads])    src/cudaConvFFTDataStreams.cu:121-522

MATLAB arrays ``A(h, w, f)`` are numpy arrays of shape ``(H, W, F)`` (any memory order; they are
marshalled to the reference's column-major memory = C-order ``[F][W][H]``).  A MATLAB ``gpuArray``
is a :class:`GpuArray` (device memory held by a torch tensor — torch is only the allocator here).
A cell array is a Python list.

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked),
but every compute call needs the CUDA library and a B200 and raises otherwise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Union

import numpy as np

__all__ = [
    "FFTConvError", "GpuArray", "gpuArray", "gather", "computeFFTsize16", "computeFFTsize",
    "cudaFFTData", "cudaConvFFTData", "cudaConvolutionFFT", "cudaConvFFTDataStreams",
    "cudaFFTDataClamp", "modulateAndNormalize", "Options", "conv_bank", "fft_data_device",
    "conv_batch", "Bank", "Plan", "lib", "LIB_PATH", "launch_count", "last_error", "EXPORTED_SYMBOLS", "profile", "profile_read",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFTCONV_LIB") or os.path.join(_HERE, "libfftconv.so")   # override: A/B builds only

EXPORTED_SYMBOLS = [
    "fftconv_fft_size16", "fftconv_fft_size_pow2", "fftconv_fft_data", "fftconv_fft_data_clamp",
    "fftconv_conv_fft_data", "fftconv_conv_fft_data_streams", "fftconv_convolution_fft",
    "fftconv_conv_bank", "fftconv_conv_batch", "fftconv_conv_pyramid", "fftconv_bank_create", "fftconv_bank_info", "fftconv_bank_conv", "fftconv_bank_conv_max",
    "fftconv_bank_conv_detect", "fftconv_bank_conv_topk",
    "fftconv_plan_create", "fftconv_plan_execute", "fftconv_plan_info", "fftconv_plan_destroy",
    "fftconv_bank_destroy", "fftconv_modulate_and_normalize", "fftconv_launch_count",
    "fftconv_workspace_bytes", "fftconv_release", "fftconv_last_error", "fftconv_version",
    "fftconv_profile_enable", "fftconv_profile_kinds", "fftconv_profile_name", "fftconv_profile_read",
    "fftconv_spectrum_ready_event", "fftconv_spectrum_bind_raw", "fftconv_query_path",
    "fftconv_peer_alloc", "fftconv_peer_open", "fftconv_peer_close", "fftconv_peer_free", "fftconv_peer_signal",
    "fftconv_peer_wait", "fftconv_peer_wait_all", "fftconv_peer_pull", "fftconv_peer_status", "fftconv_peer_allgather",
]

# error ids / messages of the reference
ERRID_FFTDATA = "parallel:gpu:mexGPUExample:InvalidInput"      # src/cudaFFTData.cu:28
ERRID_CONV = "cudaConvFFTData:InvalidInput"                     # src/cudaConvFFTData.cu:47
MSG_INVALID = "Invalid input to MEX file."                      # src/cudaFFTData.cu:29
MSG_NOT_GPU = "The data must be FFT-ed real array in GPU"       # src/cudaConvFFTData.cu:69
MSG_NOT_CELL = "Kernel must be a cell array"                    # src/cudaConvFFTData.cu:107
MSG_KERNEL_TYPE = "Kernels must be of type float and have features larger than 1"   # :198
MSG_KERNEL_SHAPE = ("Kernel and Data must have the same number of features and kernel size should be smaller "
                    "than data size")                            # src/cudaConvFFTData.cu:230
MSG_WRONG_NARGS = "Wrong number of inputs"                      # src/cudaConvolutionFFT.cu:46
MSG_INVALID_DATA = "Invalid data input"                         # src/cudaConvolutionFFT.cu:54


class FFTConvError(RuntimeError):
    """Raised where the MEX would call mexErrMsgIdAndTxt / mexErrMsgTxt (and for CUDA errors,
    where the reference prints and calls exit(), src/cudaConvFFTData.h:6-29)."""

    def __init__(self, identifier: str, message: str, code: int = -1):
        super().__init__(f"{identifier}: {message}" if identifier else message)
        self.identifier = identifier
        self.message = message
        self.code = code


class Options(ctypes.Structure):
    """fftconv_options (include/fftconv.h) — all zero = exact reference behaviour."""
    _fields_ = [("correlate", ctypes.c_int), ("crop_h", ctypes.c_int), ("crop_w", ctypes.c_int),
                ("out_ld", ctypes.c_int), ("force_generic", ctypes.c_int), ("path", ctypes.c_int),
                ("reserved", ctypes.c_int * 2)]


_lib = None


def lib() -> ctypes.CDLL:
    """Load libfftconv.so (in-tree build).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FFTConvError("fftconv:LibraryMissing",
                               f"{LIB_PATH} not found — build it with `make -C cuda-fft-convolution_b200/csrc` "
                               "(or __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        c_int, c_vp, c_ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong
        L.fftconv_fft_size16.argtypes = [c_int]
        L.fftconv_spectrum_ready_event.argtypes = [c_int, c_vp]
        L.fftconv_query_path.argtypes = [c_int] * 6 + [c_vp, c_vp, c_vp]
        L.fftconv_peer_alloc.argtypes = [ctypes.c_size_t, c_int, ctypes.POINTER(c_vp), c_vp]
        L.fftconv_peer_open.argtypes = [c_vp, c_int, ctypes.POINTER(c_vp)]
        L.fftconv_peer_close.argtypes = [c_vp, c_int]
        L.fftconv_peer_free.argtypes = [c_vp, c_int]
        L.fftconv_peer_signal.argtypes = [c_vp, ctypes.c_ulonglong, c_int, c_vp]
        L.fftconv_peer_wait.argtypes = [c_vp, ctypes.c_ulonglong, c_int, c_vp]
        L.fftconv_peer_wait_all.argtypes = [c_vp, c_int, ctypes.c_ulonglong, c_int, c_vp]
        L.fftconv_peer_pull.argtypes = [c_vp, c_vp, ctypes.c_size_t, c_int, c_vp]
        L.fftconv_peer_status.argtypes = [c_int]
        L.fftconv_peer_allgather.argtypes = [c_vp, c_int, c_int, c_vp, ctypes.c_ulonglong, ctypes.c_ulonglong, c_int, c_vp]
        L.fftconv_fft_size_pow2.argtypes = [c_int]
        L.fftconv_fft_data.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]
        L.fftconv_fft_data_clamp.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]
        L.fftconv_conv_fft_data.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                            c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp]
        L.fftconv_conv_fft_data_streams.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                                    c_vp, c_vp, c_int, c_vp, c_int]
        L.fftconv_convolution_fft.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                              c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp]
        L.fftconv_conv_batch.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                         c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp]
        L.fftconv_spectrum_bind_raw.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]
        L.fftconv_conv_pyramid.argtypes = [c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                           c_vp, c_vp, c_vp, c_vp, c_int, c_vp]
        L.fftconv_bank_create.argtypes = [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp]
        L.fftconv_bank_info.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.fftconv_bank_conv.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp]
        L.fftconv_bank_conv_max.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp]
        L.fftconv_bank_conv_detect.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_vp, ctypes.c_float, c_int, c_vp, c_vp, c_int, c_vp]
        L.fftconv_bank_conv_topk.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp]
        L.fftconv_plan_create.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp]
        L.fftconv_plan_execute.argtypes = [c_vp, c_vp]
        L.fftconv_plan_info.argtypes = [c_vp, c_vp, c_vp]
        L.fftconv_plan_destroy.argtypes = [c_vp]
        L.fftconv_plan_destroy.restype = None
        L.fftconv_bank_destroy.argtypes = [c_vp]
        L.fftconv_bank_destroy.restype = None
        L.fftconv_conv_bank.argtypes = [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp]
        L.fftconv_modulate_and_normalize.argtypes = [c_vp, c_vp, c_ll, c_int, c_vp]
        L.fftconv_launch_count.restype = c_ll
        L.fftconv_workspace_bytes.argtypes = [c_int]
        L.fftconv_workspace_bytes.restype = c_ll
        L.fftconv_release.restype = None
        L.fftconv_last_error.restype = ctypes.c_char_p
        L.fftconv_version.restype = ctypes.c_char_p
        L.fftconv_profile_enable.argtypes = [c_int]
        L.fftconv_profile_enable.restype = None
        L.fftconv_profile_name.argtypes = [c_int]
        L.fftconv_profile_name.restype = ctypes.c_char_p
        L.fftconv_profile_read.argtypes = [c_int, c_vp, c_vp, c_int]
        _lib = L
    return _lib


def last_error() -> str:
    return lib().fftconv_last_error().decode()


def launch_count() -> int:
    return int(lib().fftconv_launch_count())


def profile(enable: bool) -> None:
    """Bracket every kernel launch with CUDA events (roofline leg of bench.py)."""
    lib().fftconv_profile_enable(1 if enable else 0)


def profile_read(reset: bool = True):
    """-> {kernel name: (total ms, launches)} for every kernel kind that ran while profiling."""
    L = lib()
    out = {}
    for kind in range(L.fftconv_profile_kinds()):
        ms, n = ctypes.c_double(0), ctypes.c_longlong(0)
        if L.fftconv_profile_read(kind, ctypes.byref(ms), ctypes.byref(n), 1 if reset else 0) != 0:
            raise FFTConvError("fftconv:CudaError", last_error())
        if n.value:
            out[L.fftconv_profile_name(kind).decode()] = (ms.value, n.value)
    return out


def _check(rc: int, errid: str):
    if rc != 0:
        raise FFTConvError(errid if rc > -9 else "fftconv:CudaError", last_error(), rc)


def computeFFTsize16(n: int) -> int:
    """computeFFTsize16 — src/cudaConvFFTData.h:96-102 (through the C ABI)."""
    return int(lib().fftconv_fft_size16(int(n)))


def computeFFTsize(n: int) -> int:
    """computeFFTsize — src/cudaConvFFTData.h:67-94."""
    return int(lib().fftconv_fft_size_pow2(int(n)))


# ----------------------------------------------------------------------------- gpuArray
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise FFTConvError("fftconv:NoGpu", "no CUDA device available — this engine has no CPU fallback")
    return torch


class GpuArray:
    """Stand-in for a MATLAB gpuArray: `shape` is the MATLAB size, `tensor` the device memory in
    column-major order (a torch tensor of the reversed shape)."""

    def __init__(self, tensor, shape, is_complex: bool):
        self.tensor = tensor
        self.shape = tuple(int(s) for s in shape)
        self.is_complex = is_complex

    @property
    def device_index(self) -> int:
        return int(self.tensor.device.index or 0)

    def data_ptr(self) -> int:
        return int(self.tensor.data_ptr())

    def gather(self) -> np.ndarray:
        a = self.tensor.cpu().numpy()
        return np.asfortranarray(a.transpose(*range(a.ndim - 1, -1, -1)))


def gpuArray(a: np.ndarray, device: int = 0) -> GpuArray:
    """MATLAB gpuArray(a): upload a host array, keeping MATLAB (column-major) semantics."""
    torch = _torch()
    a = np.asarray(a)
    mem = np.ascontiguousarray(a.transpose(*range(a.ndim - 1, -1, -1)))
    t = torch.from_numpy(mem).to(f"cuda:{device}")
    return GpuArray(t, a.shape, np.iscomplexobj(a))


def gather(g: GpuArray) -> np.ndarray:
    return g.gather()


def _as_single_3d(a, errid, msg) -> np.ndarray:
    """MEX type/shape check (mxSINGLE_CLASS, 3 dims, not a gpuArray) then marshal to [F][W][H]."""
    if isinstance(a, GpuArray) or not isinstance(a, np.ndarray) or a.ndim != 3 or a.dtype != np.float32:
        raise FFTConvError(errid, msg)
    return np.ascontiguousarray(a.transpose(2, 1, 0))


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    return int(getattr(stream, "cuda_stream", stream))


# --------------------------------------------------------------------------- cudaFFTData
def cudaFFTData(*args) -> GpuArray:
    """fftData = cudaFFTData(data, kernelH, kernelW)   — src/cudaFFTData.cu:18-160.

    data: host single H x W x F (3-D required, gpuArray rejected, :49-54).
    Returns a complex-single gpuArray of MATLAB size [(FFT_H/2+1), FFT_W, F] (:90-103)."""
    if len(args) != 3:
        raise FFTConvError(ERRID_FFTDATA, MSG_INVALID)
    data, kh, kw = args
    d = _as_single_3d(data, ERRID_FFTDATA, MSG_INVALID)
    return _fft_data(d, int(kh), int(kw), 0)


def cudaFFTDataClamp(data, kernelH, kernelW, kernelY, kernelX, device: int = 0) -> GpuArray:
    """cudaFFTData with the clamp/wrap pad of the SDK padData (src/convolutionFFTkernel.cu:46-76)."""
    d = _as_single_3d(data, ERRID_FFTDATA, MSG_INVALID)
    return _fft_data(d, int(kernelH), int(kernelW), device, clamp=(int(kernelY), int(kernelX)))


def _fft_data(d_fwh: np.ndarray, kh: int, kw: int, device: int, clamp=None) -> GpuArray:
    torch = _torch()
    L = lib()
    F, W, H = d_fwh.shape
    FH, FW = computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)
    CH = FH // 2 + 1
    with torch.cuda.device(device):
        spec = torch.empty((F, FW, CH), dtype=torch.complex64, device=f"cuda:{device}")
        st = torch.cuda.current_stream().cuda_stream
        if clamp is None:
            rc = L.fftconv_fft_data(d_fwh.ctypes.data, 0, H, W, F, kh, kw, spec.data_ptr(), device, st)
        else:
            rc = L.fftconv_fft_data_clamp(d_fwh.ctypes.data, 0, H, W, F, kh, kw, clamp[0], clamp[1],
                                          spec.data_ptr(), device, st)
    _check(rc, ERRID_FFTDATA)
    return GpuArray(spec, (CH, FW, F), True)


def fft_data_device(data_t, H: int, W: int, F: int, kh: int, kw: int, spec_t=None, stream=None):
    """Extension: data already on the device (torch float32 tensor, memory [F][W][H]); stream-ordered."""
    torch = _torch()
    FH, FW = computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)
    CH = FH // 2 + 1
    dev = int(data_t.device.index or 0)
    if spec_t is None:
        spec_t = torch.empty((F, FW, CH), dtype=torch.complex64, device=data_t.device)
    st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    rc = lib().fftconv_fft_data(data_t.data_ptr(), 1, H, W, F, kh, kw, spec_t.data_ptr(), dev, st)
    _check(rc, ERRID_FFTDATA)
    return spec_t


# ----------------------------------------------------------------------- kernel marshalling
class _Cell:
    """Marshals a kernel cell (list) into the pointer/size arrays of the C ABI."""

    def __init__(self, cell, F_expected: Optional[int], allow_gpu: bool = True):
        if not isinstance(cell, (list, tuple)):
            raise FFTConvError(ERRID_CONV, MSG_NOT_CELL)
        K = len(cell)
        self.K = K
        self.keep = []
        self.ptrs = (ctypes.c_void_p * max(K, 1))()
        self.kh = (ctypes.c_int * max(K, 1))()
        self.kw = (ctypes.c_int * max(K, 1))()
        self.kf = (ctypes.c_int * max(K, 1))()
        self.on_dev = (ctypes.c_ubyte * max(K, 1))()
        host = []
        for k, ker in enumerate(cell):
            if isinstance(ker, GpuArray):
                if not allow_gpu or ker.is_complex or len(ker.shape) != 3 or str(ker.tensor.dtype) != "torch.float32":
                    raise FFTConvError(ERRID_CONV, MSG_KERNEL_TYPE)
                self.keep.append(ker.tensor)
                self.ptrs[k] = ker.data_ptr()
                self.kh[k], self.kw[k], self.kf[k] = ker.shape
                self.on_dev[k] = 1
            else:
                if not isinstance(ker, np.ndarray) or ker.dtype != np.float32 or ker.ndim != 3:
                    raise FFTConvError(ERRID_CONV, MSG_KERNEL_TYPE)      # src/cudaConvFFTData.cu:197-198
                self.kh[k], self.kw[k], self.kf[k] = ker.shape
                host.append((k, ker))
        # host kernels are packed back to back so the library uploads them with one copy
        if host:
            total = sum(int(ker.size) for _, ker in host)
            pack = np.empty(total, dtype=np.float32)
            off = 0
            for k, ker in host:
                n = int(ker.size)
                pack[off:off + n] = ker.transpose(2, 1, 0).ravel()
                self.ptrs[k] = pack.ctypes.data + 4 * off
                off += n
            self.keep.append(pack)


def _threads_arg(threads):
    if threads is None:
        return None, 0
    t = np.ascontiguousarray(np.asarray(threads, dtype=np.float64).ravel())
    return t, int(t.size)


def _alloc_outs(K: int, FH: int, FW: int):
    outs = np.empty((K, FW, FH), dtype=np.float32)
    ptrs = (ctypes.c_void_p * max(K, 1))()
    for k in range(K):
        ptrs[k] = outs.ctypes.data + 4 * k * FW * FH
    return outs, ptrs


def _wrap_outs(outs: np.ndarray) -> List[np.ndarray]:
    # [FW][FH] memory == MATLAB (FH, FW) column-major
    return [outs[k].T for k in range(outs.shape[0])]


# ------------------------------------------------------------------------ cudaConvFFTData
def cudaConvFFTData(*args, options: Optional[Options] = None) -> List[np.ndarray]:
    """cvcell = cudaConvFFTData(fftData, kernelCell[, threads])  — src/cudaConvFFTData.cu:24-306.

    fftData must be a gpuArray (:68); kernelCell a cell of single 3-D arrays kh x kw x F, host or
    gpuArray, sizes may differ per cell (:194-231); optional 4-vector of thread-block sizes (:71-81,
    validated and ignored).  Returns a 1 x K list of host single (FFT_H, FFT_W) arrays (:111,275-279)."""
    return _conv_fft_data(args, options, streams=False)


def cudaConvFFTDataStreams(*args, options: Optional[Options] = None) -> List[np.ndarray]:
    """Same contract as cudaConvFFTData, host kernels only — src/cudaConvFFTDataStreams.cu:121-522."""
    return _conv_fft_data(args, options, streams=True)


def _conv_fft_data(args, options, streams: bool):
    if len(args) < 2 or len(args) > 3 or not isinstance(args[0], GpuArray):
        raise FFTConvError(ERRID_CONV if not streams else ERRID_FFTDATA, MSG_NOT_GPU)
    spec, cell = args[0], args[1]
    threads, nthreads = _threads_arg(args[2] if len(args) == 3 else None)
    if len(args) == 3 and nthreads != 4:
        _check(lib().fftconv_conv_fft_data(spec.data_ptr(), 2, 16, 1, 0, None, None, None, None, None, None, 0,
                                           threads.ctypes.data, nthreads, None, 0, None), ERRID_CONV)
    CH, FW, F = spec.shape                         # :92-98
    FH = (CH - 1) * 2
    c = _Cell(cell, F, allow_gpu=not streams)
    outs, optrs = _alloc_outs(c.K, FH, FW)
    L = lib()
    torch = _torch()
    dev = spec.device_index
    tp = threads.ctypes.data if threads is not None else None
    op = ctypes.byref(options) if options is not None else None
    with torch.cuda.device(dev):
        if streams:
            rc = L.fftconv_conv_fft_data_streams(spec.data_ptr(), CH, FW, F, c.K, c.ptrs, c.kh, c.kw, c.kf,
                                                 optrs, tp, nthreads, op, dev)
        else:
            st = torch.cuda.current_stream().cuda_stream
            rc = L.fftconv_conv_fft_data(spec.data_ptr(), CH, FW, F, c.K, c.ptrs, c.kh, c.kw, c.kf, c.on_dev,
                                         optrs, 0, tp, nthreads, op, dev, st)
    _check(rc, ERRID_CONV)
    return _wrap_outs(outs)


# ---------------------------------------------------------------------- cudaConvolutionFFT
def cudaConvolutionFFT(*args, options: Optional[Options] = None) -> List[np.ndarray]:
    """cvcell = cudaConvolutionFFT(data, maxKH, maxKW, kernelCell[, threads[, gpuId]])
    — src/cudaConvolutionFFT.cu:27-311 (gpuId is 0-based, :84-89)."""
    if len(args) < 4 or len(args) > 6:
        raise FFTConvError(ERRID_CONV, MSG_WRONG_NARGS)          # :45-46
    data, max_kh, max_kw, cell = args[:4]
    d = _as_single_3d(data, "", MSG_INVALID_DATA)                # :50-54 (mexErrMsgTxt: no id)
    if not isinstance(cell, (list, tuple)):
        raise FFTConvError(ERRID_CONV, MSG_NOT_CELL)             # :64-65
    threads, nthreads = _threads_arg(args[4] if len(args) > 4 else None)
    gpu = int(args[5]) if len(args) > 5 else 0
    F, W, H = d.shape
    FH, FW = computeFFTsize16(H + int(max_kh) - 1), computeFFTsize16(W + int(max_kw) - 1)
    L = lib()
    tp = threads.ctypes.data if threads is not None else None
    if threads is not None and nthreads != 4:
        _check(L.fftconv_convolution_fft(d.ctypes.data, 0, H, W, F, int(max_kh), int(max_kw), 0, None, None, None,
                                         None, None, None, 0, tp, nthreads, None, gpu, None), ERRID_CONV)
    c = _Cell(cell, F)
    outs, optrs = _alloc_outs(c.K, FH, FW)
    torch = _torch()
    op = ctypes.byref(options) if options is not None else None
    with torch.cuda.device(gpu):
        st = torch.cuda.current_stream().cuda_stream
        rc = L.fftconv_convolution_fft(d.ctypes.data, 0, H, W, F, int(max_kh), int(max_kw), c.K, c.ptrs, c.kh, c.kw,
                                       c.kf, c.on_dev, optrs, 0, tp, nthreads, op, gpu, st)
    _check(rc, ERRID_CONV)
    return _wrap_outs(outs)


# ------------------------------------------------------------------------------ extensions
def conv_bank(spec_t, bank_t, kh: int, kw: int, out_t=None, options: Optional[Options] = None, stream=None,
              spectrum_ready=None):
    """Device-resident bank: spec_t complex64 [F][FW][CH], bank_t float32 [K][F][kw][kh] (torch, cuda).
    Writes K planes [FW][FH] (or the crop) into out_t; stream-ordered, no host sync.
    spectrum_ready: torch.cuda.Event recorded behind a collective that is still filling spec_t on another stream
    (fftconv_spectrum_ready_event): only the data-side work waits for it, the template transforms start at once."""
    torch = _torch()
    F, FW, CH = spec_t.shape
    FH = (CH - 1) * 2
    K = int(bank_t.shape[0])
    dev = int(spec_t.device.index or 0)
    if out_t is None:
        ch = options.crop_h if options is not None and options.crop_h > 0 else FH
        cw = options.crop_w if options is not None and options.crop_w > 0 else FW
        ld = options.out_ld if options is not None and options.out_ld > 0 else ch
        out_t = torch.empty((K, cw, ld), dtype=torch.float32, device=spec_t.device)
    st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    op = ctypes.byref(options) if options is not None else None
    if spectrum_ready is not None:
        _check(lib().fftconv_spectrum_ready_event(dev, ctypes.c_void_p(spectrum_ready.cuda_event)), ERRID_CONV)
    rc = lib().fftconv_conv_bank(spec_t.data_ptr(), CH, FW, F, K, bank_t.data_ptr(), kh, kw, out_t.data_ptr(), op, dev, st)
    _check(rc, ERRID_CONV)
    return out_t


class DeviceCells:
    """A kernel cell resident on the device, marshalled once: K float32 tensors [F][kw_k][kh_k] whose sizes may differ per
    template (the gpuArray cells of src/cudaConvFFTData.cu:204-231).  Holds the pointer / size arrays of the C ABI so that a
    repeated call costs no per-template Python work."""

    def __init__(self, cells, F: int):
        for t in cells:
            if t.dim() != 3 or int(t.shape[0]) != F or not t.is_contiguous() or str(t.dtype) != "torch.float32" or not t.is_cuda:
                raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        self.K = len(cells)
        self.F = F
        self.keep = list(cells)
        self.kws = np.array([int(t.shape[1]) for t in cells], dtype=np.int32).reshape(-1)
        self.khs = np.array([int(t.shape[2]) for t in cells], dtype=np.int32).reshape(-1)
        self.ptrs = np.array([t.data_ptr() for t in cells], dtype=np.uint64).reshape(-1)
        self.max_kh = int(self.khs.max(initial=1))
        self.max_kw = int(self.kws.max(initial=1))


def convolution_fft_device(data_t, bank_t, out_t=None, options: Optional[Options] = None, stream=None, data_ready=None,
                           max_kh: Optional[int] = None, max_kw: Optional[int] = None):
    """cudaConvolutionFFT with everything resident on the device (fftconv_convolution_fft, src/cudaConvolutionFFT.cu:27-311
    is its host-buffer original): data_t float32 [F][W][H], bank_t float32 [K][F][kw][kh] (kh x kw is also the declared
    maximum template size) -> out_t [K][FW][FH].  One call: on the overlap-save path the raw data is tiled directly and
    no full-plane spectrum is ever formed.  Stream-ordered, no host sync.
    bank_t may also be a DeviceCells (or a list of K device tensors [F][kw_k][kh_k]): template sizes that differ per cell,
    as the reference's kernel cell allows; max_kh x max_kw is then the declared maximum (default: the largest template).
    data_ready: torch.cuda.Event recorded behind whatever is still filling data_t on ANOTHER stream (a peer delivery over
    NVLink, fftconv_spectrum_ready_event): only the data-side work waits for it, the template transforms start at once."""
    torch = _torch()
    F, W, H = (int(x) for x in data_t.shape)
    if isinstance(bank_t, (list, tuple)):
        bank_t = DeviceCells(bank_t, F)
    if isinstance(bank_t, DeviceCells):
        if bank_t.F != F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        K, kp, khs, kws = bank_t.K, bank_t.ptrs, bank_t.khs, bank_t.kws
        kh = int(max_kh) if max_kh is not None else bank_t.max_kh
        kw = int(max_kw) if max_kw is not None else bank_t.max_kw
    else:
        K, Fk, kw, kh = (int(x) for x in bank_t.shape)
        if Fk != F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        kp = np.uint64(bank_t.data_ptr()) + np.uint64(4 * F * kw * kh) * np.arange(K, dtype=np.uint64)
        khs = np.full(K, kh, dtype=np.int32)
        kws = np.full(K, kw, dtype=np.int32)
        kh = int(max_kh) if max_kh is not None else kh
        kw = int(max_kw) if max_kw is not None else kw
    FH, FW = computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)
    dev = int(data_t.device.index or 0)
    if out_t is None:
        out_t = torch.empty((K, FW, FH), dtype=torch.float32, device=data_t.device)
    op = np.uint64(out_t.data_ptr()) + np.uint64(4 * FW * FH) * np.arange(K, dtype=np.uint64)
    ond = np.ones(max(K, 1), dtype=np.uint8)
    st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    o = ctypes.byref(options) if options is not None else None
    if data_ready is not None:
        _check(lib().fftconv_spectrum_ready_event(dev, ctypes.c_void_p(data_ready.cuda_event)), ERRID_CONV)
    rc = lib().fftconv_convolution_fft(data_t.data_ptr(), 1, H, W, F, kh, kw, K, kp.ctypes.data, khs.ctypes.data, kws.ctypes.data,
                                       None, ond.ctypes.data, op.ctypes.data, 1, None, 0, o, dev, st)
    _check(rc, ERRID_CONV)
    return out_t


def conv_batch(data_t, bank_t, out_t=None, options: Optional[Options] = None, stream=None):
    """Batched images against one device-resident bank (fftconv_conv_batch): data_t float32 [N][F][W][H],
    bank_t float32 [K][F][kw][kh] (torch, cuda) -> out_t [N][K][FW][FH].  Stream-ordered, no host sync."""
    torch = _torch()
    N, F, W, H = (int(x) for x in data_t.shape)
    K, Fk, kw, kh = (int(x) for x in bank_t.shape)
    if Fk != F:
        raise FFTConvError(ERRID_CONV, "Kernel and Data must have the same number of features and kernel size "
                                       "should be smaller than data size", -5)
    FH, FW = computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)
    dev = int(data_t.device.index or 0)
    if out_t is None:
        out_t = torch.empty((N, K, FW, FH), dtype=torch.float32, device=data_t.device)
    plane = FW * FH * 4
    kp = np.uint64(bank_t.data_ptr()) + np.uint64(4 * F * kw * kh) * np.arange(K, dtype=np.uint64)     # (no Python loop over
    op = np.uint64(out_t.data_ptr()) + np.uint64(plane) * np.arange(N * K, dtype=np.uint64)            # N*K plane pointers)
    khs = np.full(K, kh, dtype=np.int32)
    kws = np.full(K, kw, dtype=np.int32)
    ond = np.ones(K, dtype=np.uint8)
    st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    o = ctypes.byref(options) if options is not None else None
    rc = lib().fftconv_conv_batch(data_t.data_ptr(), 1, N, H, W, F, kh, kw, K, kp.ctypes.data, khs.ctypes.data, kws.ctypes.data,
                                  None, ond.ctypes.data, op.ctypes.data, 1, o, dev, st)
    _check(rc, ERRID_CONV)
    return out_t


def conv_pyramid(levels, bank_t, kh: int, kw: int, outs=None, specs=None, shapes=None,
                 options: Optional[Options] = None, stream=None, data_ready=None):
    """Feature pyramid x one device-resident bank in ONE call (fftconv_conv_pyramid, BASELINE config 5).

    levels   list of float32 torch tensors [F][W_l][H_l] on the device -- or None when every level arrives as a spectrum
    specs    optional list of complex64 [F][FW_l][CH_l] spectra from fft_data_device (used where levels is None / levels[l] is None);
             `shapes` = [(H_l, W_l)] is then required
    bank_t   float32 [K][F][kw][kh] on the device; kh x kw is also the declared maximum template size
    data_ready: torch.cuda.Event recorded behind whatever is still filling the levels on ANOTHER stream (the NCCL broadcast
             of the packed pyramid, fftconv_spectrum_ready_event): only the data-side work waits for it, the template
             transforms of the first chunk start at once.
    Returns [out_l float32 [K][FW_l][FH_l]] (allocated here unless `outs` is given).  Stream-ordered, no host sync."""
    torch = _torch()
    K, F, kw_, kh_ = (int(x) for x in bank_t.shape)
    if (kh_, kw_) != (kh, kw):
        raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
    L = len(levels) if levels is not None else len(specs)
    if shapes is None:
        shapes = [(int(t.shape[2]), int(t.shape[1])) for t in levels]
    dev_t = bank_t.device
    dev = int(dev_t.index or 0)
    planes = [(computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)) for (H, W) in shapes]
    if outs is None:
        outs = [torch.empty((K, FW, FH), dtype=torch.float32, device=dev_t) for (FH, FW) in planes]
    dp = (ctypes.c_void_p * L)(*[(levels[l].data_ptr() if levels is not None and levels[l] is not None else None) for l in range(L)])
    sp = None
    if specs is not None:
        sp = (ctypes.c_void_p * L)(*[(specs[l].data_ptr() if specs[l] is not None else None) for l in range(L)])
    Hs = (ctypes.c_int * L)(*[h for h, _ in shapes])
    Ws = (ctypes.c_int * L)(*[w for _, w in shapes])
    # pointer tables as numpy arrays (a bank of 20 000 templates x 10 levels is 200 000 plane pointers: a Python loop over
    # them costs more than the convolution)
    ks = np.arange(K, dtype=np.uint64)
    kp = np.uint64(bank_t.data_ptr()) + np.uint64(4 * F * kw * kh) * ks
    op = np.concatenate([np.uint64(outs[l].data_ptr()) + np.uint64(4 * FW * FH) * ks for l, (FH, FW) in enumerate(planes)])
    khs = np.full(K, kh, dtype=np.int32)
    kws = np.full(K, kw, dtype=np.int32)
    ond = np.ones(K, dtype=np.uint8)
    st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    o = ctypes.byref(options) if options is not None else None
    if data_ready is not None:
        _check(lib().fftconv_spectrum_ready_event(dev, ctypes.c_void_p(data_ready.cuda_event)), ERRID_CONV)
    rc = lib().fftconv_conv_pyramid(L, dp if levels is not None else None, sp, Hs, Ws, F, kh, kw, K, kp.ctypes.data,
                                    khs.ctypes.data, kws.ctypes.data, None, ond.ctypes.data, op.ctypes.data, o, dev, st)
    _check(rc, ERRID_CONV)
    return outs


class Bank:
    """Prepared template bank (fftconv_bank_*): the kernel-side counterpart of cudaFFTData.  The template spectra
    are computed once and stay on the device; conv() then serves any image size.

        bank = fc.Bank(cell_of_kernels)              # kh x kw x F float32 arrays (host) or GpuArrays
        planes = bank.conv(data)                     # list of (FH, FW) float32 arrays, like cudaConvolutionFFT
        out_t = bank.conv_device(data_t)             # torch [F][W][H] on the device -> [K][FW][FH] on the device
    """

    def __init__(self, cell, device: int = 0):
        torch = _torch()
        c = _Cell(cell, None)
        if c.K == 0:
            raise FFTConvError(ERRID_CONV, MSG_INVALID)
        c.F = int(c.kf[0])
        self._keep = c
        self.device = device
        h = ctypes.c_void_p(0)
        st = torch.cuda.current_stream(device).cuda_stream
        rc = lib().fftconv_bank_create(c.K, c.ptrs, c.kh, c.kw, c.kf, c.on_dev, c.F, device, st, ctypes.byref(h))
        _check(rc, ERRID_CONV)
        self._h = h
        K, F, mh, mw, nb = (ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_longlong(0))
        lib().fftconv_bank_info(h, ctypes.byref(K), ctypes.byref(F), ctypes.byref(mh), ctypes.byref(mw), ctypes.byref(nb))
        self.K, self.F, self.maxKH, self.maxKW, self.bytes = K.value, F.value, mh.value, mw.value, nb.value
        self._keep = None                            # the kernels are no longer needed once transformed

    def plane(self, H: int, W: int):
        return computeFFTsize16(H + self.maxKH - 1), computeFFTsize16(W + self.maxKW - 1)

    def conv(self, data, options: Optional[Options] = None) -> List[np.ndarray]:
        d_fwh = _as_single_3d(data, ERRID_CONV, MSG_INVALID)        # marshalled to [F][W][H]
        F, W, H = d_fwh.shape
        if F != self.F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        FH, FW = self.plane(H, W)
        outs, ptrs = _alloc_outs(self.K, FH, FW)
        torch = _torch()
        op = ctypes.byref(options) if options is not None else None
        rc = lib().fftconv_bank_conv(self._h, d_fwh.ctypes.data, 0, H, W, ptrs, 0, op,
                                     torch.cuda.current_stream(self.device).cuda_stream)
        _check(rc, ERRID_CONV)
        return _wrap_outs(outs)

    def conv_device(self, data_t, out_t=None, options: Optional[Options] = None, stream=None):
        torch = _torch()
        F, W, H = (int(x) for x in data_t.shape)
        if F != self.F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        FH, FW = self.plane(H, W)
        if out_t is None:
            out_t = torch.empty((self.K, FW, FH), dtype=torch.float32, device=data_t.device)
        plane = FW * FH * 4
        ptrs = (ctypes.c_void_p * self.K)(*[out_t.data_ptr() + plane * k for k in range(self.K)])
        st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        op = ctypes.byref(options) if options is not None else None
        rc = lib().fftconv_bank_conv(self._h, data_t.data_ptr(), 1, H, W, ptrs, 1, op, st)
        _check(rc, ERRID_CONV)
        return out_t

    def conv_max(self, data):
        """Fused detection reduction (fftconv_bank_conv_max): -> (value[K] float32, y[K], x[K] int32), the maximum of
        every template's full linear convolution and its 0-based position; no plane is written or copied."""
        d_fwh = _as_single_3d(data, ERRID_CONV, MSG_INVALID)
        F, W, H = d_fwh.shape
        if F != self.F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        peaks = np.zeros((self.K, 4), dtype=np.int32)
        torch = _torch()
        rc = lib().fftconv_bank_conv_max(self._h, d_fwh.ctypes.data, 0, H, W, peaks.ctypes.data, 0,
                                         torch.cuda.current_stream(self.device).cuda_stream)
        _check(rc, ERRID_CONV)
        return peaks[:, 0].copy().view(np.float32), peaks[:, 1].copy(), peaks[:, 2].copy()

    def _bias_arg(self, bias):
        if bias is None:
            return None, None
        b = np.ascontiguousarray(np.asarray(bias, dtype=np.float32).ravel())
        if b.size != self.K:
            raise FFTConvError(ERRID_CONV, MSG_INVALID)
        return b, b.ctypes.data

    def detect(self, data, threshold: float, bias=None, max_per_template: int = 16):
        """Fused detection (fftconv_bank_conv_detect): every response conv + bias[k] >= threshold of every template's
        full linear convolution.  -> (counts[K] int32, value[K, M] float32, y[K, M], x[K, M] int32), the M =
        max_per_template largest per template in descending order (unused slots: -inf, -1, -1); counts may exceed M."""
        d_fwh = _as_single_3d(data, ERRID_CONV, MSG_INVALID)
        F, W, H = d_fwh.shape
        if F != self.F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        M = int(max_per_template)
        dets = np.zeros((self.K, M, 4), dtype=np.int32)
        counts = np.zeros(self.K, dtype=np.int32)
        keep, bp = self._bias_arg(bias)
        torch = _torch()
        rc = lib().fftconv_bank_conv_detect(self._h, d_fwh.ctypes.data, 0, H, W, bp, float(threshold), M, dets.ctypes.data,
                                            counts.ctypes.data, 0, torch.cuda.current_stream(self.device).cuda_stream)
        _check(rc, ERRID_CONV)
        return counts, dets[:, :, 0].copy().view(np.float32), dets[:, :, 1].copy(), dets[:, :, 2].copy()

    def topk(self, data, k: int, bias=None):
        """Fused top-k (fftconv_bank_conv_topk): the k largest responses conv + bias[t] of every template, exact,
        descending.  -> (value[K, k] float32, y[K, k], x[K, k] int32)."""
        d_fwh = _as_single_3d(data, ERRID_CONV, MSG_INVALID)
        F, W, H = d_fwh.shape
        if F != self.F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        dets = np.zeros((self.K, int(k), 4), dtype=np.int32)
        keep, bp = self._bias_arg(bias)
        torch = _torch()
        rc = lib().fftconv_bank_conv_topk(self._h, d_fwh.ctypes.data, 0, H, W, bp, int(k), dets.ctypes.data, 0,
                                          torch.cuda.current_stream(self.device).cuda_stream)
        _check(rc, ERRID_CONV)
        return dets[:, :, 0].copy().view(np.float32), dets[:, :, 1].copy(), dets[:, :, 2].copy()

    def close(self):
        if getattr(self, "_h", None):
            lib().fftconv_bank_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """Graph plan (fftconv_plan_*): the launch sequence of cudaFFTData + cudaConvFFTData over fixed device buffers,
    captured once into a CUDA graph; execute() is one graph launch.  The tensors' CONTENTS may change between executions.

        plan = fc.Plan(data_t, bank_t, kh, kw)       # data_t [F][W][H], bank_t [K][F][kw][kh] float32 on the device
        data_t.copy_(next_frame); plan.execute()      # -> plan.out [K][FW][FH] (and plan.spec, the cudaFFTData spectrum)
    With data_t=None and spec_t given the plan starts from the spectrum (cudaConvFFTData only)."""

    def __init__(self, data_t, bank_t, kh: int, kw: int, spec_t=None, out_t=None, shape=None, options: Optional[Options] = None,
                 stream=None):
        torch = _torch()
        K, F, bkw, bkh = (int(x) for x in bank_t.shape)
        if data_t is not None:
            Fd, W, H = (int(x) for x in data_t.shape)
        else:
            H, W, Fd = shape
        if Fd != F:
            raise FFTConvError(ERRID_CONV, MSG_KERNEL_SHAPE, -5)
        FH, FW = computeFFTsize16(H + kh - 1), computeFFTsize16(W + kw - 1)
        dev = bank_t.device
        self.device = int(dev.index or 0)
        self.spec = spec_t if spec_t is not None else torch.empty((F, FW, FH // 2 + 1), dtype=torch.complex64, device=dev)
        ch = options.crop_h if options is not None and options.crop_h > 0 else FH
        cw = options.crop_w if options is not None and options.crop_w > 0 else FW
        ld = options.out_ld if options is not None and options.out_ld > 0 else ch
        self.out = out_t if out_t is not None else torch.empty((K, cw, ld), dtype=torch.float32, device=dev)
        self._keep = (data_t, bank_t, options)
        st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        h = ctypes.c_void_p(0)
        rc = lib().fftconv_plan_create(data_t.data_ptr() if data_t is not None else None, H, W, F, kh, kw, self.spec.data_ptr(),
                                       K, bank_t.data_ptr(), bkh, bkw, self.out.data_ptr(),
                                       ctypes.byref(options) if options is not None else None, self.device, st, ctypes.byref(h))
        _check(rc, ERRID_CONV)
        self._h = h
        n, pth = ctypes.c_int(0), ctypes.c_int(0)
        lib().fftconv_plan_info(h, ctypes.byref(n), ctypes.byref(pth))
        self.graph_nodes, self.path = n.value, pth.value

    def execute(self, stream=None):
        torch = _torch()
        st = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        _check(lib().fftconv_plan_execute(self._h, st), ERRID_CONV)
        return self.out

    def close(self):
        if getattr(self, "_h", None):
            lib().fftconv_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def modulateAndNormalize(a: GpuArray, b: GpuArray) -> None:
    """In place a = a .* b / numel(a)  — modulateAndNormalize, src/convolutionFFTkernel.cu:84-100."""
    torch = _torch()
    n = int(a.tensor.numel())
    dev = a.device_index
    rc = lib().fftconv_modulate_and_normalize(a.data_ptr(), b.data_ptr(), n, dev, torch.cuda.current_stream(dev).cuda_stream)
    _check(rc, ERRID_CONV)
