// MEX shim: cvcell = cudaConvolutionFFT(data, maxKH, maxKW, kernelCell[, threads[, gpuId]])
// replaces src/cudaConvolutionFFT.cu:27-311
#include "mex_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (mxInitGPU() != MX_GPU_SUCCESS) mexErrMsgTxt("mxInitGPU fail");
    if (nrhs < 4 || nrhs > 6) mexErrMsgIdAndTxt(kErrConv, "Wrong number of inputs");           // :45-46
    if (mxIsGPUArray(prhs[0]) || mxGetNumberOfDimensions(prhs[0]) != 3 || mxGetClassID(prhs[0]) != mxSINGLE_CLASS)
        mexErrMsgTxt("Invalid data input");                                                     // :50-54
    const int maxKH = (int)mxGetScalar(prhs[1]), maxKW = (int)mxGetScalar(prhs[2]);
    int nthreads = 0;
    const double* threads = thread_arg(nrhs, prhs, 4, nthreads);
    int dev = 0;
    cudaGetDevice(&dev);
    if (nrhs > 5) dev = (int)mxGetScalar(prhs[5]);                                              // 0-based, :84-89
    const mwSize* d = mxGetDimensions(prhs[0]);
    const int H = (int)d[0], W = (int)d[1], F = (int)d[2];
    const int FH = fftconv_fft_size16(H + maxKH - 1), FW = fftconv_fft_size16(W + maxKW - 1);
    KernelCell c;
    marshal_cell(prhs[3], true, c);
    std::vector<float*> outs;
    plhs[0] = alloc_out_cell((int)c.ptr.size(), FH, FW, outs);
    const int rc = fftconv_convolution_fft((const float*)mxGetData(prhs[0]), 0, H, W, F, maxKH, maxKW,
                                           (int)c.ptr.size(), c.ptr.data(), c.kh.data(), c.kw.data(), c.kf.data(),
                                           c.on_dev.data(), outs.data(), 0, threads, nthreads, nullptr, dev, nullptr);
    raise_if(rc, kErrConv, &c);
    c.release();
}
