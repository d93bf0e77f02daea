// Shared marshalling for the MEX shims.  The shims only translate mxArray / mxGPUArray to the plain-pointer C ABI of
// include/fftconv.h.  Build on a MATLAB box:
//   mex -largeArrayDims cudaConvolutionFFT.cpp -I../../include -L../fftconv_b200 -lfftconv -lmwgpu -lcudart
// No MATLAB in this image: tests/mex_stub/ compiles the same sources against a stand-in mex.h / mxGPUArray.h and drives
// marshal_cell / alloc_out_cell and the reference's error ids with fake mxArrays (tests/test_mex_shims.py).
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include "mex.h"
#include "gpu/mxGPUArray.h"
#include "../../include/fftconv.h"

static const char* kErrConv = "cudaConvFFTData:InvalidInput";               // src/cudaConvFFTData.cu:47
static const char* kErrFft = "parallel:gpu:mexGPUExample:InvalidInput";     // src/cudaFFTData.cu:28

struct KernelCell {
    std::vector<const float*> ptr;
    std::vector<int> kh, kw, kf;
    std::vector<unsigned char> on_dev;
    std::vector<const mxGPUArray*> handles;      // destroyed by release() (the reference leaks all but the
    void release() {                              // last one, src/cudaConvFFTData.cu:296)
        for (auto h : handles) mxGPUDestroyGPUArray(h);
        handles.clear();
    }
};

// kernel cell -> pointer arrays; raises the reference's errors (src/cudaConvFFTData.cu:106-107,194-225)
static void marshal_cell(const mxArray* cell, bool allow_gpu, KernelCell& c) {
    if (mxGetClassID(cell) != mxCELL_CLASS) {
        c.release();                                  // the caller may already hold the spectrum handle in c
        mexErrMsgIdAndTxt(kErrConv, "Kernel must be a cell array");
    }
    const mwSize K = mxGetNumberOfElements(cell);
    for (mwSize k = 0; k < K; ++k) {
        const mxArray* a = mxGetCell(cell, k);
        const mwSize* dims;
        if (!mxIsGPUArray(a)) {
            if (mxGetClassID(a) != mxSINGLE_CLASS || mxGetNumberOfDimensions(a) != 3) {
                c.release();
                mexErrMsgIdAndTxt(kErrConv, "Kernels must be of type float and have features larger than 1");
            }
            dims = mxGetDimensions(a);
            c.ptr.push_back((const float*)mxGetData(a));
            c.on_dev.push_back(0);
        } else {
            const mxGPUArray* g = allow_gpu ? mxGPUCreateFromMxArray(a) : nullptr;
            if (!g || mxGPUGetClassID(g) != mxSINGLE_CLASS || mxGPUGetNumberOfDimensions(g) != 3) {
                if (g) mxGPUDestroyGPUArray(g);
                c.release();
                mexErrMsgIdAndTxt(kErrConv, "Kernels must be of type float and have features larger than 1");
            }
            c.handles.push_back(g);
            dims = mxGPUGetDimensions(g);
            c.ptr.push_back((const float*)mxGPUGetDataReadOnly(g));
            c.on_dev.push_back(1);
        }
        c.kh.push_back((int)dims[0]); c.kw.push_back((int)dims[1]); c.kf.push_back((int)dims[2]);
    }
}

// 1 x K cell of host single FFT_H x FFT_W planes (src/cudaConvFFTData.cu:111,186-188,275-279)
static mxArray* alloc_out_cell(int K, int FH, int FW, std::vector<float*>& outs) {
    mxArray* cell = mxCreateCellMatrix(1, K);
    mwSize d[2] = {(mwSize)FH, (mwSize)FW};
    for (int k = 0; k < K; ++k) {
        mxArray* p = mxCreateUninitNumericArray(2, d, mxSINGLE_CLASS, mxREAL);
        outs.push_back((float*)mxGetData(p));
        mxSetCell(cell, k, p);
    }
    return cell;
}

// (CH, FW, F) of a spectrum gpuArray (src/cudaConvFFTData.cu:92-98).  MATLAB drops trailing singleton dimensions, so a
// single-channel spectrum has two dimensions: F defaults to 1 instead of reading past the dimension vector.
static void spectrum_dims(const mxGPUArray* spec, int& CH, int& FW, int& F) {
    const mwSize nd = mxGPUGetNumberOfDimensions(spec);
    const mwSize* sd = mxGPUGetDimensions(spec);
    CH = (int)sd[0]; FW = nd > 1 ? (int)sd[1] : 1; F = nd > 2 ? (int)sd[2] : 1;
}

static const double* thread_arg(int nrhs, const mxArray* prhs[], int idx, int& n) {
    n = 0;
    if (nrhs <= idx) return nullptr;
    n = (int)mxGetNumberOfElements(prhs[idx]);
    return (const double*)mxGetData(prhs[idx]);        // length validated by the library (:71-72)
}

static void raise_if(int rc, const char* id, KernelCell* c = nullptr) {
    if (rc == 0) return;
    if (c) c->release();
    mexErrMsgIdAndTxt(id, "%s", fftconv_last_error()); // never exit() (reference: src/cudaConvFFTData.h:6-29)
}
