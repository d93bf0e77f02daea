/* TEST INFRASTRUCTURE ONLY — CPU oracle, never linked into the product.
 *
 * float64 (and float32) direct full convolution summed over feature channels:
 *     out = sum_f conv2(D(:,:,f), k(:,:,f))          demoCudaConvolutionFFT.m:91-96
 * embedded top-left in a zero FFT_H x FFT_W plane — the plane cudaConvFFTData returns
 * (src/cudaConvFFTData.cu:111,186-188,275-279).  Anything past the plane is folded back
 * (circular), which is what a transform of that size does when the kernel is larger than
 * the declared maximum (size check only at src/cudaConvFFTData.cu:229).
 *
 * Memory layouts are the reference's (column-major, h contiguous,
 * src/cudaConvFFTData.cuh:26-27): data [F][W][H], kernel [F][kw][kh], out [FFT_W][FFT_H].
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DEFINE_DIRECT(NAME, ACC)                                                              \
void NAME(const float* data, int H, int W, int F, const float* ker, int kh, int kw,          \
          int FH, int FW, ACC* out, int threads)                                              \
{                                                                                             \
    const int OH = H + kh - 1, OW = W + kw - 1;                                               \
    ACC* full = (ACC*)calloc((size_t)OH * OW, sizeof(ACC));                                   \
    if (threads <= 0) threads = 0;                                                            \
    _Pragma("omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : omp_get_max_threads())") \
    for (int ox = 0; ox < OW; ++ox) {                                                         \
        ACC* col = full + (size_t)ox * OH;                                                    \
        for (int f = 0; f < F; ++f) {                                                         \
            for (int kx = 0; kx < kw; ++kx) {                                                 \
                const int dx = ox - kx;                                                       \
                if (dx < 0 || dx >= W) continue;                                              \
                const float* dcol = data + ((size_t)f * W + dx) * H;                          \
                const float* kcol = ker + ((size_t)f * kw + kx) * kh;                         \
                for (int ky = 0; ky < kh; ++ky) {                                             \
                    const ACC kv = (ACC)kcol[ky];                                             \
                    ACC* o = col + ky;                                                        \
                    for (int y = 0; y < H; ++y) o[y] += kv * (ACC)dcol[y];                    \
                }                                                                             \
            }                                                                                 \
        }                                                                                     \
    }                                                                                         \
    memset(out, 0, (size_t)FH * FW * sizeof(ACC));                                            \
    for (int ox = 0; ox < OW; ++ox)                                                           \
        for (int oy = 0; oy < OH; ++oy)                                                       \
            out[(size_t)(ox % FW) * FH + (oy % FH)] += full[(size_t)ox * OH + oy];            \
    free(full);                                                                               \
}

DEFINE_DIRECT(oracle_direct_conv, double)
DEFINE_DIRECT(oracle_direct_conv_f32, float)
