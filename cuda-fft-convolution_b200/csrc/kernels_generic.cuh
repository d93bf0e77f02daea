// Generic (any plane size) pipeline kernels.
//
//   fwd_h_pass           real columns -> half spectrum along h (zero pad fused into the load,
//                        two real columns per complex line FFT)
//   fwd_w_pass           complex FFT along w of the half-transformed data -> compat spectrum
//   conv_w_pass_generic  per kernel & u-tile: for every channel forward w-FFT of the kernel's
//                        half-transform, multiply by the data spectrum, accumulate over
//                        channels IN THE FREQUENCY DOMAIN, then ONE inverse w-FFT
//   inv_h_pass           C2R along h (two columns per complex line), 1/(FH*FW) scale and
//                        optional crop fused into the coalesced store
//
// Replaces (reference, per kernel): padData (src/cudaConvFFTData.cuh:11-31), cufftExecR2C
// (src/cudaConvFFTData.cu:244), elementwiseProductAndNormalize (.cuh:47-67), F x cufftExecC2R
// (.cu:262) and sumAlongFeatures (.cuh:70-92).
//
// Layouts (h contiguous everywhere, as in the reference, src/cudaConvFFTData.cuh:26-27):
//   real source  [plane][col][row]            half transform T [plane][col][CH]
//   spectrum     [F][FW][CH]  (compat)        Z intermediate   [k][FW][CH]
//   output       plane k at out[k], [FW][FH] (or cropped [cw][ch])
#pragma once
#include "line_fft.cuh"

namespace fftconv {

struct SrcDesc {          // one multi-channel real source (the data, or one kernel of the bank)
    const float* ptr;     // [F][cols][rows]
    int rows;             // h extent (kh / H)
    int cols;             // w extent (kw / W)
};

enum PadMode { PAD_ZERO = 0, PAD_CLAMP = 1 };

// clamp/wrap index rule of the SDK padData (src/convolutionFFTkernel.cu:63-68)
__device__ __forceinline__ int clamp_index(int i, int data, int ofs) {
    return i < data ? i : (i < data + ofs ? data - 1 : 0);
}

// ------------------------------------------------------------------------------- fwd_h_pass
// grid.x = ceil(nsrc*F*npairs / NL), NL lines per CTA; dynamic smem = 2*NL*ld*sizeof(cpx)
template <int PAD>
__global__ void fwd_h_pass(const SrcDesc* __restrict__ srcs, int nsrc, int F, int maxcols,
                           int FH, int CH, LinePlan plan, const cpx* __restrict__ tw,
                           cpx* __restrict__ T, int NL, int ld,
                           int clamp_ofs_h, int clamp_ofs_w, int padded_cols)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* b0 = reinterpret_cast<cpx*>(smem_raw);
    cpx* b1 = b0 + (size_t)NL * ld;
    const int ncol_out = (PAD == PAD_CLAMP) ? padded_cols : maxcols;   // columns produced per plane
    const int npairs = (ncol_out + 1) / 2;
    const long long nlines = (long long)nsrc * F * npairs;
    const long long line0 = (long long)blockIdx.x * NL;

    // ---- load (pad fused)
    for (int idx = threadIdx.x; idx < NL * FH; idx += blockDim.x) {
        const int l = idx / FH, y = idx - l * FH;
        const long long line = line0 + l;
        cpx v = make_float2(0.f, 0.f);
        if (line < nlines) {
            const int cp = (int)(line % npairs);
            const long long pf = line / npairs;
            const int f = (int)(pf % F);
            const int s = (int)(pf / F);
            const SrcDesc d = srcs[s];
            const float* base = d.ptr + (size_t)f * d.rows * d.cols;
            const int xa = 2 * cp, xb = 2 * cp + 1;
            if (PAD == PAD_ZERO) {
                if (y < d.rows) {
                    if (xa < d.cols) v.x = base[(size_t)xa * d.rows + y];
                    if (xb < d.cols) v.y = base[(size_t)xb * d.rows + y];
                }
            } else {
                const int sy = clamp_index(y, d.rows, clamp_ofs_h);
                if (xa < ncol_out) v.x = base[(size_t)clamp_index(xa, d.cols, clamp_ofs_w) * d.rows + sy];
                if (xb < ncol_out) v.y = base[(size_t)clamp_index(xb, d.cols, clamp_ofs_w) * d.rows + sy];
            }
        }
        b0[(size_t)l * ld + y] = v;
    }
    __syncthreads();
    const cpx* res = fft_lines<false>(b0, b1, NL, ld, plan, tw);

    // ---- split the two packed real transforms and store u = 0..CH-1
    for (int idx = threadIdx.x; idx < NL * CH; idx += blockDim.x) {
        const int l = idx / CH, u = idx - l * CH;
        const long long line = line0 + l;
        if (line >= nlines) continue;
        const int cp = (int)(line % npairs);
        const long long pf = line / npairs;
        const int xa = 2 * cp, xb = 2 * cp + 1;
        const cpx zu = res[(size_t)l * ld + u];
        const cpx zn = cconj(res[(size_t)l * ld + (u == 0 ? 0 : FH - u)]);
        const cpx a = make_float2(0.5f * (zu.x + zn.x), 0.5f * (zu.y + zn.y));
        const cpx d = make_float2(0.5f * (zu.x - zn.x), 0.5f * (zu.y - zn.y));
        cpx* o = T + ((size_t)pf * ncol_out) * CH + u;
        o[(size_t)xa * CH] = a;
        if (xb < ncol_out) o[(size_t)xb * CH] = make_float2(d.y, -d.x);   // -i * d
    }
}

// ------------------------------------------------------------------------------- fwd_w_pass
// grid = (ceil(CH/TU), planes); smem = 2*TU*ld*sizeof(cpx).  T [plane][ncols][CH] -> S [plane][FW][CH]
__global__ void fwd_w_pass(const cpx* __restrict__ T, int ncols, int FW, int CH,
                           LinePlan plan, const cpx* __restrict__ tw,
                           cpx* __restrict__ S, int TU, int ld)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* b0 = reinterpret_cast<cpx*>(smem_raw);
    cpx* b1 = b0 + (size_t)TU * ld;
    const int u0 = blockIdx.x * TU;
    const size_t p = blockIdx.y;
    const cpx* Tp = T + p * (size_t)ncols * CH;
    for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
        const int x = idx / TU, u = idx - x * TU;
        cpx v = make_float2(0.f, 0.f);
        if (x < ncols && u0 + u < CH) v = Tp[(size_t)x * CH + u0 + u];
        b0[(size_t)u * ld + x] = v;
    }
    __syncthreads();
    const cpx* res = fft_lines<false>(b0, b1, TU, ld, plan, tw);
    cpx* Sp = S + p * (size_t)FW * CH;
    for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
        const int v = idx / TU, u = idx - v * TU;
        if (u0 + u < CH) Sp[(size_t)v * CH + u0 + u] = res[(size_t)u * ld + v];
    }
}

// ---------------------------------------------------------------------- conv_w_pass_generic
// grid = (ceil(CH/TU), nk); smem = 3*TU*ld*sizeof(cpx)
// T: kernel half transforms [k][F][maxcols][CH]; S: data spectrum [F][FW][CH]; Z: [k][FW][CH]
template <bool CONJ>
__global__ void conv_w_pass_generic(const cpx* __restrict__ T, const int* __restrict__ kcols, int maxcols,
                                    const cpx* __restrict__ S, int F, int FW, int CH,
                                    LinePlan plan, const cpx* __restrict__ tw,
                                    cpx* __restrict__ Z, int TU, int ld)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* b0 = reinterpret_cast<cpx*>(smem_raw);
    cpx* b1 = b0 + (size_t)TU * ld;
    cpx* acc = b1 + (size_t)TU * ld;
    const int u0 = blockIdx.x * TU;
    const size_t k = blockIdx.y;
    const int ncols = min(kcols[k], FW);          // pad semantics: never read past the plane
    for (int idx = threadIdx.x; idx < TU * ld; idx += blockDim.x) acc[idx] = make_float2(0.f, 0.f);

    for (int f = 0; f < F; ++f) {
        const cpx* Tp = T + (k * F + f) * (size_t)maxcols * CH;
        for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
            const int x = idx / TU, u = idx - x * TU;
            cpx v = make_float2(0.f, 0.f);
            if (x < ncols && u0 + u < CH) v = Tp[(size_t)x * CH + u0 + u];
            b0[(size_t)u * ld + x] = v;
        }
        __syncthreads();
        const cpx* res = fft_lines<false>(b0, b1, TU, ld, plan, tw);
        const cpx* Sf = S + (size_t)f * FW * CH;
        for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
            const int v = idx / TU, u = idx - v * TU;
            if (u0 + u < CH) {
                const cpx d = __ldg(&Sf[(size_t)v * CH + u0 + u]);
                const cpx kx = res[(size_t)u * ld + v];
                cpx a = acc[(size_t)u * ld + v];
                cfma(a, d, CONJ ? cconj(kx) : kx);
                acc[(size_t)u * ld + v] = a;
            }
        }
        __syncthreads();
    }
    const cpx* res = fft_lines<true>(acc, b0, TU, ld, plan, tw);
    cpx* Zk = Z + k * (size_t)FW * CH;
    for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
        const int x = idx / TU, u = idx - x * TU;
        if (u0 + u < CH) Zk[(size_t)x * CH + u0 + u] = res[(size_t)u * ld + x];
    }
}

// ------------------------------------------------------------------------------- inv_h_pass
// grid.x = ceil(nk*(FW/2)/NL); smem = 2*NL*ld*sizeof(cpx).  Z [k][FW][CH] -> out[k] (FW x FH, or crop)
// crop: only rows < crop_h and columns < crop_w are stored, with leading dimension out_ld.
__global__ void inv_h_pass(const cpx* __restrict__ Z, int nk, int FH, int FW, int CH,
                           LinePlan plan, const cpx* __restrict__ tw, float scale,
                           float* const* __restrict__ outs, int crop_h, int crop_w, int out_ld,
                           int NL, int ld, const unsigned long long* __restrict__ skip_if_equal)
{
    // spectrum -> plane leg of the overlap-save path: nothing to do when the spectrum is bit-identical to the one
    // fftconv_fft_data produced from raw data this library still holds (skip_if_equal[0] == [1], see SpecCache)
    if (skip_if_equal && skip_if_equal[0] == skip_if_equal[1]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* b0 = reinterpret_cast<cpx*>(smem_raw);
    cpx* b1 = b0 + (size_t)NL * ld;
    const int npairs = FW / 2;
    const long long nlines = (long long)nk * npairs;
    const long long line0 = (long long)blockIdx.x * NL;
    const int half = FH / 2;

    for (int idx = threadIdx.x; idx < NL * CH; idx += blockDim.x) {
        const int l = idx / CH, u = idx - l * CH;
        const long long line = line0 + l;
        cpx za = make_float2(0.f, 0.f), zb = za;
        if (line < nlines) {
            const int cp = (int)(line % npairs);
            const size_t k = (size_t)(line / npairs);
            const cpx* Zk = Z + k * (size_t)FW * CH;
            za = Zk[(size_t)(2 * cp) * CH + u];
            zb = Zk[(size_t)(2 * cp + 1) * CH + u];
        }
        cpx* L = b0 + (size_t)l * ld;
        if (u == 0 || u == half) {                 // C2R ignores Im of DC / Nyquist
            L[u] = make_float2(za.x, zb.x);
        } else {
            L[u] = make_float2(za.x - zb.y, za.y + zb.x);            // za + i*zb
            L[FH - u] = make_float2(za.x + zb.y, zb.x - za.y);       // conj(za) + i*conj(zb)
        }
    }
    __syncthreads();
    const cpx* res = fft_lines<true>(b0, b1, NL, ld, plan, tw);
    for (int idx = threadIdx.x; idx < NL * crop_h; idx += blockDim.x) {
        const int l = idx / crop_h, y = idx - l * crop_h;
        const long long line = line0 + l;
        if (line >= nlines) continue;
        const int cp = (int)(line % npairs);
        const size_t k = (size_t)(line / npairs);
        const cpx r = res[(size_t)l * ld + y];
        float* o = outs[k];
        const int xa = 2 * cp, xb = xa + 1;
        if (xa < crop_w) o[(size_t)xa * out_ld + y] = r.x * scale;
        if (xb < crop_w) o[(size_t)xb * out_ld + y] = r.y * scale;
    }
}

// in-place  a = a*b/dataN, grid-stride (modulateAndNormalize, src/convolutionFFTkernel.cu:84-100)
__global__ void modulate_and_normalize_kernel(cpx* __restrict__ a, const cpx* __restrict__ b, long long n, float q)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const cpx p = cmul(a[i], b[i]);
        a[i] = make_float2(q * p.x, q * p.y);
    }
}

}  // namespace fftconv
