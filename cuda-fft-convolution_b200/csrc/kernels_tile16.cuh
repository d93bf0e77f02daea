// 16-point-tiled fast path (placeholder until the tiled kernels land): reports "unsupported"
// so that every call takes the generic pipeline.
#pragma once
#include "line_fft.cuh"
#include "../../include/fftconv.h"

namespace fftconv {
struct DevBuf;
struct SrcDesc;
static inline int tile16_opt_in() { return 0; }
static inline bool tile16_supported(int, int, int, int) { return false; }
static inline size_t tile16_scratch_per_kernel(int, int, int, int, int) { return 0; }
static inline int tile16_round_chunk(int kc, int) { return kc; }
static inline int tile16_reserve(DevBuf&, DevBuf&, int, int, int, int, int, int) { return 0; }
static inline int tile16_prepare_spectrum(DevBuf&, const cpx*, int, int, int, cudaStream_t) { return 0; }
static inline int tile16_chunk(DevBuf&, DevBuf&, DevBuf&, int, int, int, int, int, const SrcDesc*, int,
                               float* const*, const fftconv_options&, int, cudaStream_t) { return 0; }
}  // namespace fftconv
