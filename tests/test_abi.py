"""CPU tests of the boundary: libfftconv.so loads, exports every symbol include/fftconv.h
declares, and the argument checks that do not need a device behave like the reference."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fftconv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fftconv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(fc):
    L = fc.lib()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/fftconv.h but not exported"
    assert sorted(fc.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in L.fftconv_version()


def test_fft_size16_through_abi(fc, oracle):
    for n in list(range(1, 600)) + [4095, 4096, 4607, 4608, 4609]:
        assert fc.computeFFTsize16(n) == oracle.compute_fft_size16(n)
    for n in (1, 16, 17, 73, 100, 512, 513, 1000):
        assert fc.computeFFTsize(n) == oracle.compute_fft_size_pow2(n)


def test_argument_errors_without_device(fc):
    L = fc.lib()
    # null spectrum -> "The data must be FFT-ed real array in GPU" (src/cudaConvFFTData.cu:68-69)
    rc = L.fftconv_conv_fft_data(None, 9, 16, 1, 0, None, None, None, None, None, None, 0, None, 0, None, 0, None)
    assert rc == -6 and "FFT-ed real array in GPU" in fc.last_error()
    # thread vector with != 4 entries (src/cudaConvFFTData.cu:71-72)
    t = (ctypes.c_double * 3)(8, 8, 8)
    dummy = ctypes.c_void_p(16)
    rc = L.fftconv_conv_fft_data(dummy, 9, 16, 1, 0, None, None, None, None, None, None, 0, t, 3, None, 0, None)
    assert rc == -3 and "CUDA Thread Size must be 4 integers" in fc.last_error()
    # invalid data
    rc = L.fftconv_fft_data(None, 0, 4, 4, 1, 3, 3, None, 0, None)
    assert rc == -1 and fc.last_error() == "Invalid input to MEX file."
    # kernel with the wrong feature count / larger than the plane (src/cudaConvFFTData.cu:229-230)
    kp = (ctypes.c_void_p * 1)(64)
    op = (ctypes.c_void_p * 1)(64)
    one = (ctypes.c_int * 1)
    rc = L.fftconv_conv_fft_data(dummy, 9, 16, 2, 1, kp, one(3), one(3), one(5), None, op, 0, None, 0, None, 0, None)
    assert rc == -5 and "same number of features" in fc.last_error()
    rc = L.fftconv_conv_fft_data(dummy, 9, 16, 2, 1, kp, one(17), one(3), one(2), None, op, 0, None, 0, None, 0, None)
    assert rc == -5


def test_python_mirror_argument_errors(fc):
    with pytest.raises(fc.FFTConvError) as e:
        fc.cudaFFTData(np.zeros((4, 4), np.float32), 3, 3)           # not 3-D (src/cudaFFTData.cu:51)
    assert e.value.identifier == "parallel:gpu:mexGPUExample:InvalidInput"
    with pytest.raises(fc.FFTConvError):
        fc.cudaFFTData(np.zeros((4, 4, 2), np.float64), 3, 3)        # not single
    with pytest.raises(fc.FFTConvError):
        fc.cudaFFTData(np.zeros((4, 4, 2), np.float32), 3)           # nrhs != 3
    with pytest.raises(fc.FFTConvError) as e:
        fc.cudaConvolutionFFT(np.zeros((4, 4, 2), np.float32), 3, 3)  # too few inputs
    assert "Wrong number of inputs" in str(e.value)
    with pytest.raises(fc.FFTConvError) as e:
        fc.cudaConvolutionFFT(np.zeros((4, 4, 2), np.float32), 3, 3, np.zeros((3, 3, 2), np.float32))
    assert "Kernel must be a cell array" in str(e.value)
    with pytest.raises(fc.FFTConvError) as e:
        fc.cudaConvFFTData(np.zeros((4, 4, 2), np.float32), [])       # spectrum must be a gpuArray
    assert "FFT-ed real array in GPU" in str(e.value)
    with pytest.raises(fc.FFTConvError) as e:
        fc.cudaConvolutionFFT(np.zeros((4, 4, 2), np.float32), 3, 3, [], [8, 8, 8])
    assert "CUDA Thread Size must be 4 integers" in str(e.value)


def test_device_cells_refuse_host_tensors_and_wrong_shapes(fc):
    """fc.DeviceCells marshals a DEVICE-resident kernel cell (the gpuArray kernels of src/cudaConvFFTData.cu:204-231) for the
    one-shot entry point: host tensors, a channel count that differs from the data's (:229), non-contiguous or non-float32
    cells are refused with the reference's message before anything reaches the library -- there is no CPU fallback."""
    import torch
    ok = torch.zeros((3, 4, 5))
    for bad in ([ok],                                                       # on the host
                [torch.zeros((2, 4, 5))],                                   # F differs
                [torch.zeros((3, 4, 5), dtype=torch.float64)],              # not single
                [torch.zeros((3, 5, 4)).transpose(1, 2)],                   # not contiguous
                [torch.zeros((3, 4))]):                                     # not 3-D
        with pytest.raises(fc.FFTConvError) as e:
            fc.DeviceCells(bad, 3)
        assert "same number of features" in str(e.value)


def test_no_product_import_of_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "cuda-fft-convolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, fn


def test_torch_ops_register_and_infer_shapes():
    """torch.ops.fftconv.*: registered, shape inference on the meta device, and no CPU fallback."""
    import torch
    import fftconv_b200.torch_ops  # noqa: F401
    d = torch.empty((31, 256, 256), device="meta")
    spec = torch.ops.fftconv.fft_data(d, 16, 16)
    assert tuple(spec.shape) == (31, 272, 137) and spec.dtype == torch.complex64
    bank = torch.empty((1000, 31, 16, 16), device="meta")
    assert tuple(torch.ops.fftconv.conv_fft_data(spec, bank).shape) == (1000, 272, 272)
    imgs = torch.empty((64, 32, 512, 512), device="meta")
    assert tuple(torch.ops.fftconv.convolution_fft(imgs, torch.empty((256, 32, 32, 32), device="meta")).shape) == (64, 256, 544, 544)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.fftconv.fft_data(torch.zeros((2, 8, 8)), 3, 3)          # CPU tensors: no fallback


def _query(fc, H, W, F, kh, kw, K, **opt):
    L = fc.lib()
    rh, rw = (ctypes.c_int * 8)(), (ctypes.c_int * 8)()
    o = fc.Options(**opt) if opt else None
    path = L.fftconv_query_path(H, W, F, kh, kw, K, ctypes.byref(o) if o is not None else None, rh, rw)
    return path, [r for r in rh if r], [r for r in rw if r]


def test_pipeline_selection_host_logic(fc):
    """Which pipeline serves which BASELINE config (pure host logic, no device): the overlap-save / tcgen05 path needs a
    bank that fills an MMA block, small banks take the 16-point-tiled path, planes of 1024 and more with large templates
    the in-place large-plane path, sizes with a prime factor above 17 the generic one."""
    assert _query(fc, 256, 256, 31, 16, 16, 1000)[0] == 3          # C2
    assert _query(fc, 256, 256, 31, 16, 16, 10)[0] == 2            # same shapes, 10 templates
    assert _query(fc, 64, 8, 5, 10, 4, 10)[0] == 2                 # C1
    assert _query(fc, 512, 512, 32, 32, 32, 256)[0] == 3           # C4 (per image)
    path, rh, rw = _query(fc, 4096, 4096, 1, 512, 512, 64)         # C3: plane 4608 = 9 * 512
    assert path == 4 and rh == [9, 32, 16] and rw == [9, 32, 16]
    path, rh, rw = _query(fc, 1024, 1024, 3, 128, 100, 4)          # plane 1152 x 1136 = (9 * 128) x (16 * 71): 71 is prime
    assert path == 1 and rh == [] and rw == []
    path, rh, rw = _query(fc, 1920, 1080, 3, 129, 65, 8)           # 2048 x 1152
    assert path == 4 and int(np.prod(rh)) == 2048 and int(np.prod(rw)) == 1152 and rw[0] == 9
    # forced paths fall back when the shape is outside the path's range
    assert _query(fc, 256, 256, 31, 16, 16, 1000, path=1)[0] == 1
    assert _query(fc, 256, 256, 31, 16, 16, 1000, force_generic=1)[0] == 1
    assert _query(fc, 300, 300, 2, 40, 40, 100, path=3)[0] in (1, 2)       # templates above 32 x 32
    assert _query(fc, 100, 100, 1, 8, 8, 1, path=4)[0] == 4                # any 16 m plane with factors <= 17
    assert _query(fc, 290, 100, 1, 15, 8, 1, path=4)[0] == 1               # 304 = 16 * 19
    assert fc.lib().fftconv_query_path(0, 8, 1, 3, 3, 1, None, None, None) < 0


def test_in_place_plans_cover_every_supported_size(fc):
    """Every plane side 16 m up to 8192 whose odd part factors over {3, 5, 7, 9, 11, 13, 17} gets a plan whose radices
    multiply back to the side, odd radices first, at most 8 stages; nothing else is claimed."""
    def smooth17(m):
        for p in (2, 3, 5, 7, 11, 13, 17):
            while m % p == 0:
                m //= p
        return m == 1
    for m in range(1, 513):
        n = 16 * m
        path, rh, _ = _query(fc, n, 64, 1, 1, 1, 1, path=4)
        if smooth17(m):
            assert path == 4 and int(np.prod(rh)) == n, (n, rh)
            odd_done = False
            for r in rh:
                if r % 2 == 0:
                    odd_done = True
                else:
                    assert not odd_done, (n, rh)
                assert r in (2, 3, 4, 5, 7, 8, 9, 11, 13, 16, 17, 32), (n, rh)
        else:
            assert path != 4, n


def test_reference_arm_restarts_with_all_host_threads():
    """`bench.py --impl reference` under torchrun inherits OMP_NUM_THREADS=1; rank 0 restarts itself with the core count so the
    host arm uses the threads it reports (`cpu_baseline.cores`), other ranks print nothing and exit 0."""
    import json
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    env.pop("FFTCONV_BENCH_REEXEC", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                         "--warmup", "1"], env=dict(env, RANK="1"), capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
