/* TEST INFRASTRUCTURE: a minimal stand-in for MATLAB's mex.h / matrix.h, just large enough to COMPILE AND DRIVE the MEX
 * shims of cuda-fft-convolution_b200/mex/ without MATLAB (the image has none).  Semantics kept from MATLAB where the
 * shims depend on them: column-major data, trailing singleton dimensions are dropped (so an H x W x 1 array reports two
 * dimensions, src/cudaConvFFTData.cu:197-198), mexErrMsgIdAndTxt does not return (here: throws MexError). */
#ifndef FFTCONV_TEST_MEX_H_
#define FFTCONV_TEST_MEX_H_
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxDOUBLE_CLASS, mxSINGLE_CLASS } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;

struct mxArray {
    mxClassID cls = mxUNKNOWN_CLASS;
    mxComplexity cplx = mxREAL;
    std::vector<mwSize> dims;
    void* data = nullptr;            /* host data (numeric) or, for a gpuArray, the device (or stand-in) pointer */
    bool is_gpu = false;
    bool owns = true;
    std::vector<mxArray*> cells;
    size_t numel() const { size_t n = 1; for (mwSize d : dims) n *= d; return dims.empty() ? 0 : n; }
    size_t elsize() const { return (cls == mxDOUBLE_CLASS ? 8 : 4) * (cplx == mxCOMPLEX ? 2 : 1); }
};

struct MexError { std::string id, msg; };

inline void mx_normalise_dims(std::vector<mwSize>& d) { while (d.size() > 2 && d.back() == 1) d.pop_back(); }

inline mxClassID mxGetClassID(const mxArray* a) { return a->cls; }
inline size_t mxGetNumberOfElements(const mxArray* a) { return a->cls == mxCELL_CLASS ? a->cells.size() : a->numel(); }
inline mxArray* mxGetCell(const mxArray* a, mwSize i) { return a->cells[i]; }
inline mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
inline const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
inline void* mxGetData(const mxArray* a) { return a->data; }
inline double mxGetScalar(const mxArray* a) {
    return a->cls == mxDOUBLE_CLASS ? *(const double*)a->data : (double)*(const float*)a->data;
}
inline mxArray* mxCreateCellMatrix(mwSize m, mwSize n) {
    mxArray* a = new mxArray; a->cls = mxCELL_CLASS; a->dims = {m, n}; a->cells.assign(m * n, nullptr); return a;
}
inline mxArray* mxCreateUninitNumericArray(mwSize nd, const mwSize* d, mxClassID cls, mxComplexity c) {
    mxArray* a = new mxArray; a->cls = cls; a->cplx = c; a->dims.assign(d, d + nd); mx_normalise_dims(a->dims);
    a->data = malloc(a->numel() * a->elsize() + 16);
    memset(a->data, 0xAB, a->numel() * a->elsize());       /* "uninitialised" */
    return a;
}
inline void mxSetCell(mxArray* c, mwSize i, mxArray* v) { c->cells[i] = v; }
inline void mxDestroyArray(mxArray* a) {
    if (!a) return;
    for (mxArray* c : a->cells) mxDestroyArray(c);
    if (a->owns && a->data && !a->is_gpu) free(a->data);
    delete a;
}
[[noreturn]] inline void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[2048]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw MexError{id, buf};
}
[[noreturn]] inline void mexErrMsgTxt(const char* msg) { throw MexError{"", msg}; }
#endif
