// TEST INFRASTRUCTURE ONLY — never linked into the product.
//
// Replays the reference's GPU path without MATLAB: the reference's OWN device kernels
// (padData, elementwiseProductAndNormalize, sumAlongFeatures from src/cudaConvFFTData.cuh) and
// host helpers (computeFFTsize16, iDivUp from src/cudaConvFFTData.h) are #included from where
// they lie under /root/reference at build time (never copied into this repo), and driven in
// the order of the MEX hot loop (src/cudaConvolutionFFT.cu:109-310) with raw pointers instead
// of mxArrays, linked against the image's cuFFT 11.4.  Used (a) to generate the golden
// fixtures under tests/golden/ and (b) as the "reference's own cuFFT path on B200" comparator.
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cstdlib>
#include "cudaConvFFTData.h"      // -I/root/reference/src
#include "cudaConvFFTData.cuh"

#define RR_CUDA(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "ref_replay: %s at line %d\n", cudaGetErrorString(e), __LINE__); return -1; } } while (0)
#define RR_FFT(x)  do { cufftResult e = (x); if (e != CUFFT_SUCCESS) { fprintf(stderr, "ref_replay: cufft error %d at line %d\n", (int)e, __LINE__); return -2; } } while (0)

extern "C" int ref_fft_size16(int n) { return computeFFTsize16(n); }

// data: host [F][W][H]; kernels[k]: host [F][kw][kh]; outs[k]: host FH*FW floats.
// threads: {H, W, D, 2D} block shape (defaults 16,8,8,32, src/cudaConvolutionFFT.cu:32-35).
// ms_out (optional): GPU+host wall time of the kernel loop in milliseconds (CUDA events).
extern "C" int ref_convolution_fft(const float* data, int H, int W, int F, int maxKH, int maxKW,
                                   int K, const float* const* kernels, const int* kh, const int* kw,
                                   float* const* outs, const int* threads, float* ms_out)
{
    const int tH = threads ? threads[0] : 16, tW = threads ? threads[1] : 8;
    const int tD = threads ? threads[2] : 8, t2 = threads ? threads[3] : 32;
    const int FH = computeFFTsize16(H + maxKH - 1), FW = computeFFTsize16(W + maxKW - 1);
    const int CH = FH / 2 + 1;
    const size_t data_b = sizeof(float) * (size_t)W * H * F;
    const size_t fft_b = sizeof(float) * (size_t)FW * FH * F;
    const size_t cfft_b = sizeof(float2) * (size_t)FW * CH * F;
    const size_t conv_b = sizeof(float) * (size_t)FW * FH;

    int n[2] = {FW, FH}, cn[2] = {FW, CH};
    cufftHandle r2c, c2r;
    RR_FFT(cufftPlanMany(&r2c, 2, n, n, 1, FW * FH, cn, 1, FW * CH, CUFFT_R2C, F));
    RR_FFT(cufftPlanMany(&c2r, 2, n, cn, 1, FW * CH, n, 1, FW * FH, CUFFT_C2R, F));

    float *d_data, *d_padded, *d_ifft, *d_conv, *d_kernel = nullptr;
    cufftComplex *d_spec, *d_kspec, *d_prod;
    RR_CUDA(cudaMalloc(&d_data, data_b));
    RR_CUDA(cudaMalloc(&d_padded, fft_b));
    RR_CUDA(cudaMemcpy(d_data, data, data_b, cudaMemcpyHostToDevice));
    dim3 b3(tH, tW, tD), g3(iDivUp(FW, b3.x), iDivUp(FH, b3.y), iDivUp(F, b3.z));
    dim3 b2(t2, t2), g2(iDivUp(FW, b2.x), iDivUp(FH, b2.y));
    padData<<<g3, b3>>>(d_padded, d_data, FW, FH, W, H, F);
    RR_CUDA(cudaMalloc(&d_spec, cfft_b));
    RR_FFT(cufftExecR2C(r2c, d_padded, d_spec));
    RR_CUDA(cudaDeviceSynchronize());
    RR_CUDA(cudaFree(d_data));
    RR_CUDA(cudaMalloc(&d_ifft, fft_b));
    RR_CUDA(cudaMalloc(&d_conv, conv_b));
    RR_CUDA(cudaMalloc(&d_kspec, cfft_b));
    RR_CUDA(cudaMalloc(&d_prod, cfft_b));

    cudaEvent_t e0, e1;
    RR_CUDA(cudaEventCreate(&e0));
    RR_CUDA(cudaEventCreate(&e1));
    RR_CUDA(cudaEventRecord(e0));
    for (int k = 0; k < K; ++k) {
        const size_t kb = sizeof(float) * (size_t)kw[k] * kh[k] * F;
        RR_CUDA(cudaMalloc(&d_kernel, kb));
        RR_CUDA(cudaMemcpy(d_kernel, kernels[k], kb, cudaMemcpyHostToDevice));
        padData<<<g3, b3>>>(d_padded, d_kernel, FW, FH, kw[k], kh[k], F);
        RR_FFT(cufftExecR2C(r2c, d_padded, d_kspec));
        RR_CUDA(cudaDeviceSynchronize());
        elementwiseProductAndNormalize<<<g3, b3>>>(d_prod, d_spec, d_kspec, CH, FW, F, 1.0f / (FW * FH));
        RR_FFT(cufftExecC2R(c2r, d_prod, d_ifft));
        RR_CUDA(cudaDeviceSynchronize());
        sumAlongFeatures<<<g2, b2>>>(d_conv, d_ifft, FH, FW, F);
        RR_CUDA(cudaMemcpy(outs[k], d_conv, conv_b, cudaMemcpyDeviceToHost));
        RR_CUDA(cudaFree(d_kernel));
    }
    RR_CUDA(cudaEventRecord(e1));
    RR_CUDA(cudaEventSynchronize(e1));
    if (ms_out) RR_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cufftDestroy(r2c); cufftDestroy(c2r);
    cudaFree(d_spec); cudaFree(d_ifft); cudaFree(d_conv); cudaFree(d_kspec); cudaFree(d_prod); cudaFree(d_padded);
    return 0;
}
