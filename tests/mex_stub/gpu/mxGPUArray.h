/* TEST INFRASTRUCTURE: stand-in for MATLAB's gpu/mxGPUArray.h (see ../mex.h).  A gpuArray is an mxArray whose data
 * pointer is device memory (cudaMalloc) -- or plain host memory when the driver runs its argument-error scenarios on a
 * machine without a GPU (those never dereference it). */
#ifndef FFTCONV_TEST_MXGPUARRAY_H_
#define FFTCONV_TEST_MXGPUARRAY_H_
#include "../mex.h"
#include <cuda_runtime.h>

#define MX_GPU_SUCCESS 0
typedef enum { MX_GPU_DO_NOT_INITIALIZE = 0, MX_GPU_INITIALIZE_VALUES } mxGPUInitialize;
typedef mxArray mxGPUArray;            /* a handle onto the same object */
extern int g_mx_live_gpu_handles;      /* leak check: every mxGPUCreate* must be matched by a destroy */
extern bool g_mx_fake_gpu;             /* no device: "device" memory is host memory */

inline int mxInitGPU() { return MX_GPU_SUCCESS; }
inline bool mxIsGPUArray(const mxArray* a) { return a->is_gpu; }
inline const mxGPUArray* mxGPUCreateFromMxArray(const mxArray* a) { ++g_mx_live_gpu_handles; return a; }
inline mxClassID mxGPUGetClassID(const mxGPUArray* g) { return g->cls; }
inline mwSize mxGPUGetNumberOfDimensions(const mxGPUArray* g) { return g->dims.size(); }
inline const mwSize* mxGPUGetDimensions(const mxGPUArray* g) { return g->dims.data(); }
inline const void* mxGPUGetDataReadOnly(const mxGPUArray* g) { return g->data; }
inline void* mxGPUGetData(mxGPUArray* g) { return g->data; }
inline void mxGPUDestroyGPUArray(const mxGPUArray*) { --g_mx_live_gpu_handles; }
inline mxGPUArray* mxGPUCreateGPUArray(mwSize nd, const mwSize* d, mxClassID cls, mxComplexity c, mxGPUInitialize) {
    mxArray* a = new mxArray; a->cls = cls; a->cplx = c; a->dims.assign(d, d + nd); mx_normalise_dims(a->dims);
    a->is_gpu = true;
    const size_t bytes = a->numel() * a->elsize();
    if (g_mx_fake_gpu) a->data = malloc(bytes + 16);
    else if (cudaMalloc(&a->data, bytes) != cudaSuccess) throw MexError{"stub:cudaMalloc", "cudaMalloc failed"};
    ++g_mx_live_gpu_handles;
    return a;
}
inline mxArray* mxGPUCreateMxArrayOnGPU(const mxGPUArray* g) { return const_cast<mxArray*>(g); }
#endif
