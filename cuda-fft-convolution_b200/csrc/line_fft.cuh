// Generic shared-memory line FFT: Stockham autosort, mixed radix, any length
// n = product(radix[s]).  Radices {2,3,4,5,7,8,9,11,13,16,17} run as register butterflies;
// any other (odd prime) radix falls back to a direct O(R) sum per output.
//
// All threads of the CTA cooperate on `nl` lines laid out at  buf + line*ld  (float2).
// The transform ping-pongs between `src` and `dst`; the buffer holding the result is
// returned.  Twiddles come from a table  tw[j] = e^{-2*pi*i*j/n}, j < n  (computed in double
// on the host) so no sin/cos is evaluated on the device.
#pragma once
#include "cplx.cuh"

namespace fftconv {

#define FFTCONV_MAX_STAGES 12

struct LinePlan {
    int n;
    int nstages;
    int radix[FFTCONV_MAX_STAGES];
};

template <int R, bool INV>
__device__ __forceinline__ void stage_regs(const cpx* __restrict__ src, cpx* __restrict__ dst, int nl, int ld,
                                           int n, int Ns, const cpx* __restrict__ tw) {
    const int nb = n / R;                 // butterflies per line
    const int tscale = n / (Ns * R);      // twiddle index scale
    const int total = nl * nb;
    for (int it = threadIdx.x; it < total; it += blockDim.x) {
        const int l = it / nb;
        const int j = it - l * nb;
        const int k = j % Ns;
        const cpx* in = src + (size_t)l * ld;
        float re[R], im[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cpx v = in[j + r * nb];
            if (r > 0 && Ns > 1) v = cmul(v, twd<INV>(tw[r * k * tscale]));
            re[r] = v.x; im[r] = v.y;
        }
        dft_regs<R, INV>(re, im);
        cpx* out = dst + (size_t)l * ld + (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) out[q * Ns] = make_float2(re[q], im[q]);
    }
}

// direct-sum stage for an arbitrary radix R (runtime): one output per work item
template <bool INV>
__device__ __forceinline__ void stage_generic(const cpx* __restrict__ src, cpx* __restrict__ dst, int nl, int ld,
                                              int n, int Ns, int R, const cpx* __restrict__ tw) {
    const int nb = n / R;
    const int tscale = n / (Ns * R);
    const int total = nl * n;             // (line, butterfly j, output q)
    for (int it = threadIdx.x; it < total; it += blockDim.x) {
        const int l = it / n;
        const int rem = it - l * n;
        const int q = rem / nb;
        const int j = rem - q * nb;
        const int k = j % Ns;
        const cpx* in = src + (size_t)l * ld + j;
        // exponent step per r:  k*tscale (stage twiddle) + q*nb (DFT_R kernel), both mod n
        const int step = (k * tscale + q * nb) % n;
        int idx = 0;
        cpx acc = make_float2(0.f, 0.f);
        for (int r = 0; r < R; ++r) {
            cfma(acc, in[r * nb], twd<INV>(tw[idx]));
            idx += step; if (idx >= n) idx -= n;
        }
        dst[(size_t)l * ld + (j - k) * R + k + q * Ns] = acc;
    }
}

template <bool INV>
__device__ __forceinline__ cpx* fft_lines(cpx* src, cpx* dst, int nl, int ld, const LinePlan& P,
                                          const cpx* __restrict__ tw) {
    int Ns = 1;
    const int n = P.n;
    for (int s = 0; s < P.nstages; ++s) {
        const int R = P.radix[s];
        switch (R) {
            case 2:  stage_regs<2, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 3:  stage_regs<3, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 4:  stage_regs<4, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 5:  stage_regs<5, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 7:  stage_regs<7, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 8:  stage_regs<8, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 9:  stage_regs<9, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 11: stage_regs<11, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 13: stage_regs<13, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 16: stage_regs<16, INV>(src, dst, nl, ld, n, Ns, tw); break;
            case 17: stage_regs<17, INV>(src, dst, nl, ld, n, Ns, tw); break;
            default: stage_generic<INV>(src, dst, nl, ld, n, Ns, R, tw); break;
        }
        __syncthreads();
        cpx* t = src; src = dst; dst = t;
        Ns *= R;
    }
    return src;
}

}  // namespace fftconv
