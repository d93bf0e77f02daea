"""The MEX shims (cuda-fft-convolution_b200/mex/*.cpp) compiled against a stand-in mex.h / gpu/mxGPUArray.h
(tests/mex_stub/) and driven with fake mxArrays: marshalling (marshal_cell, alloc_out_cell), the reference's error
ids and messages (src/cudaFFTData.cu:28-29,49-54; src/cudaConvFFTData.cu:47,69,72,107,198,230;
src/cudaConvolutionFFT.cu:45-54), gpuArray handle hygiene, and on a GPU the demo workload through all four entry points."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
STUB = os.path.join(HERE, "mex_stub")
DRIVER = os.path.join(STUB, "mex_driver")

ERR_FFT = "parallel:gpu:mexGPUExample:InvalidInput"
ERR_CONV = "cudaConvFFTData:InvalidInput"
MSG_INVALID = "Invalid input to MEX file."
MSG_NOT_GPU = "The data must be FFT-ed real array in GPU"
MSG_KTYPE = "Kernels must be of type float and have features larger than 1"
MSG_KSHAPE = "Kernel and Data must have the same number of features and kernel size should be smaller than data size"
MSG_THREADS = "CUDA Thread Size must be 4 integers"


def _driver():
    if not os.path.exists(DRIVER) or os.path.getmtime(DRIVER) < os.path.getmtime(os.path.join(STUB, "mex_driver.cpp")):
        subprocess.check_call(["make", "-C", STUB, "all"])
    return DRIVER


def _run(scenario):
    r = subprocess.run([_driver(), scenario], capture_output=True, text=True, timeout=300)
    return r.returncode, [l.split("|") for l in r.stdout.strip().splitlines()], r.stderr


def test_shims_compile_and_raise_the_reference_errors():
    rc, rows, err = _run("errors")
    assert rc == 0, err
    got = {r[0]: (r[1], r[2], r[3]) for r in rows}
    want = {
        "fftdata_wrong_nargs": (ERR_FFT, MSG_INVALID), "fftdata_2d_data": (ERR_FFT, MSG_INVALID),
        "fftdata_double_data": (ERR_FFT, MSG_INVALID), "fftdata_gpu_data": (ERR_FFT, MSG_INVALID),
        "conv_host_spectrum": (ERR_CONV, MSG_NOT_GPU), "conv_wrong_nargs": (ERR_CONV, MSG_NOT_GPU),
        "conv_not_a_cell": (ERR_CONV, "Kernel must be a cell array"),
        "conv_double_kernel": (ERR_CONV, MSG_KTYPE), "conv_2d_kernel": (ERR_CONV, MSG_KTYPE),
        "conv_thread_vector_3": (ERR_CONV, MSG_THREADS), "conv_feature_mismatch": (ERR_CONV, MSG_KSHAPE),
        "conv_kernel_larger_than_plane": (ERR_CONV, MSG_KSHAPE),
        "oneshot_wrong_nargs": (ERR_CONV, "Wrong number of inputs"), "oneshot_gpu_data": ("", "Invalid data input"),
        "oneshot_not_a_cell": (ERR_CONV, "Kernel must be a cell array"), "oneshot_thread_vector_5": (ERR_CONV, MSG_THREADS),
        "streams_host_spectrum": (ERR_FFT, MSG_NOT_GPU), "streams_gpu_kernel": (ERR_CONV, MSG_KTYPE),
    }
    assert set(got) == set(want)
    for name, (eid, msg) in want.items():
        assert got[name][0] == eid, name
        assert got[name][1].startswith(msg), (name, got[name][1])
        assert got[name][2] == "leaked_handles=0", (name, got[name][2])     # every mxGPUCreate* matched by a destroy


def test_marshal_cell_and_output_cell():
    rc, rows, err = _run("marshal")
    assert rc == 0 and rows[-1] == ["marshal", "ok"], (rows, err)


@pytest.mark.gpu
def test_demo_workload_through_the_shims_on_the_gpu():
    rc, rows, err = _run("demo")
    assert rc == 0, (rows, err)
    last = rows[-1]
    assert last[0] == "demo" and last[1] == "ok", rows
    assert float(last[2].split("=")[1]) < 1e-5
