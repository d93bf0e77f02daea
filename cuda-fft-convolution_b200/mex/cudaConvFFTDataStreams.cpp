// MEX shim: cvcell = cudaConvFFTDataStreams(fftData, kernelCell[, threads])
// replaces src/cudaConvFFTDataStreams.cu:121-522 (host kernels only, :352-374)
#include "mex_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    mxInitGPU();
    if (nrhs < 2 || nrhs > 3 || !mxIsGPUArray(prhs[0]))
        mexErrMsgIdAndTxt(kErrFft, "The data must be FFT-ed real array in GPU");   // :159-160
    int nthreads = 0;
    const double* threads = thread_arg(nrhs, prhs, 2, nthreads);
    const mxGPUArray* spec = mxGPUCreateFromMxArray(prhs[0]);
    int CH, FW, F;
    spectrum_dims(spec, CH, FW, F);                                                 // :92-98
    const int FH = (CH - 1) * 2;
    KernelCell c;
    c.handles.push_back(spec);
    marshal_cell(prhs[1], false, c);
    std::vector<float*> outs;
    plhs[0] = alloc_out_cell((int)c.ptr.size(), FH, FW, outs);
    int dev = 0;
    cudaGetDevice(&dev);
    const int rc = fftconv_conv_fft_data_streams((const fftconv_float2*)mxGPUGetDataReadOnly(spec), CH, FW, F,
                                                 (int)c.ptr.size(), c.ptr.data(), c.kh.data(), c.kw.data(),
                                                 c.kf.data(), outs.data(), threads, nthreads, nullptr, dev);
    raise_if(rc, kErrFft, &c);
    c.release();
}
