// MEX shim: fftData = cudaFFTData(data, kernelH, kernelW)      replaces src/cudaFFTData.cu:18-160
#include "mex_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (mxInitGPU() != MX_GPU_SUCCESS) mexErrMsgTxt("mxInitGPU fail");
    if (nrhs != 3 || mxIsGPUArray(prhs[0]) || mxGetNumberOfDimensions(prhs[0]) != 3 ||
        mxGetClassID(prhs[0]) != mxSINGLE_CLASS)
        mexErrMsgIdAndTxt(kErrFft, "Invalid input to MEX file.");                 // :49-54
    const mwSize* d = mxGetDimensions(prhs[0]);
    const int H = (int)d[0], W = (int)d[1], F = (int)d[2];
    const int KH = (int)mxGetScalar(prhs[1]), KW = (int)mxGetScalar(prhs[2]);
    const int FH = fftconv_fft_size16(H + KH - 1), FW = fftconv_fft_size16(W + KW - 1);
    mwSize sd[3] = {(mwSize)(FH / 2 + 1), (mwSize)FW, (mwSize)F};                 // :90-94
    mxGPUArray* spec = mxGPUCreateGPUArray(3, sd, mxSINGLE_CLASS, mxCOMPLEX, MX_GPU_DO_NOT_INITIALIZE);
    int dev = 0;
    cudaGetDevice(&dev);
    const int rc = fftconv_fft_data((const float*)mxGetData(prhs[0]), 0, H, W, F, KH, KW,
                                    (fftconv_float2*)mxGPUGetData(spec), dev, nullptr);
    if (rc) { mxGPUDestroyGPUArray(spec); raise_if(rc, kErrFft); }
    plhs[0] = mxGPUCreateMxArrayOnGPU(spec);
    mxGPUDestroyGPUArray(spec);
}
