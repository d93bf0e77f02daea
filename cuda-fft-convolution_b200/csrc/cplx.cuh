// Complex helpers and register-resident small DFT butterflies (sm_100a, fp32).
//
// All butterflies are written for the FORWARD transform  X[q] = sum_r x[r] e^{-2*pi*i*r*q/R}
// on split re/im register arrays.  The inverse transform is obtained for free by calling
// the same butterfly with the re/im arrays swapped:  IDFT(x) = swap(DFT(swap(x))).
#pragma once
#include <cuda_runtime.h>

namespace fftconv {

typedef float2 cpx;

__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
// a * conj(b)
__device__ __forceinline__ cpx cmulc(cpx a, cpx b) {
    return make_float2(fmaf(a.y, b.y, a.x * b.x), fmaf(a.y, b.x, -(a.x * b.y)));
}
__device__ __forceinline__ cpx cconj(cpx a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
// acc += a*b
__device__ __forceinline__ void cfma(cpx& acc, cpx a, cpx b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}
// twiddle fetch: table holds e^{-2*pi*i*j/n}; inverse uses the conjugate
template <bool INV>
__device__ __forceinline__ cpx twd(cpx t) { return INV ? make_float2(t.x, -t.y) : t; }

// ------------------------------------------------------------------ radix 2 / 4 / 8 / 16
__device__ __forceinline__ void dft2(float& ar, float& ai, float& br, float& bi) {
    float tr = ar - br, ti = ai - bi;
    ar += br; ai += bi; br = tr; bi = ti;
}

// forward 4-point DFT in place on (x0,x1,x2,x3); natural order out
__device__ __forceinline__ void dft4(float& r0, float& i0, float& r1, float& i1,
                                     float& r2, float& i2, float& r3, float& i3) {
    float ar = r0 + r2, ai = i0 + i2;
    float br = r0 - r2, bi = i0 - i2;
    float cr = r1 + r3, ci = i1 + i3;
    float dr = r1 - r3, di = i1 - i3;
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    // X1 = b - i*d ; X3 = b + i*d      (-i*(dr + i di) = di - i dr)
    r1 = br + di; i1 = bi - dr;
    r3 = br - di; i3 = bi + dr;
}

#define FFTCONV_SQRT1_2 0.70710678118654752440f
#define FFTCONV_COS_PI_8 0.92387953251128675613f
#define FFTCONV_SIN_PI_8 0.38268343236508977173f

template <int R> struct Dft;

template <> struct Dft<2> {
    __device__ __forceinline__ static void run(float* re, float* im) { dft2(re[0], im[0], re[1], im[1]); }
};
template <> struct Dft<4> {
    __device__ __forceinline__ static void run(float* re, float* im) {
        dft4(re[0], im[0], re[1], im[1], re[2], im[2], re[3], im[3]);
    }
};
template <> struct Dft<8> {
    __device__ __forceinline__ static void run(float* re, float* im) {
        // 2 x DFT4 over even / odd samples, then radix-2 combine with w8^k
        float er[4] = {re[0], re[2], re[4], re[6]}, ei[4] = {im[0], im[2], im[4], im[6]};
        float orr[4] = {re[1], re[3], re[5], re[7]}, oi[4] = {im[1], im[3], im[5], im[7]};
        dft4(er[0], ei[0], er[1], ei[1], er[2], ei[2], er[3], ei[3]);
        dft4(orr[0], oi[0], orr[1], oi[1], orr[2], oi[2], orr[3], oi[3]);
        // twiddles w8^1 = (1-i)/sqrt2, w8^2 = -i, w8^3 = (-1-i)/sqrt2
        float t1r = (orr[1] + oi[1]) * FFTCONV_SQRT1_2, t1i = (oi[1] - orr[1]) * FFTCONV_SQRT1_2;
        float t2r = oi[2], t2i = -orr[2];
        float t3r = (oi[3] - orr[3]) * FFTCONV_SQRT1_2, t3i = -(orr[3] + oi[3]) * FFTCONV_SQRT1_2;
        re[0] = er[0] + orr[0]; im[0] = ei[0] + oi[0];
        re[4] = er[0] - orr[0]; im[4] = ei[0] - oi[0];
        re[1] = er[1] + t1r; im[1] = ei[1] + t1i;
        re[5] = er[1] - t1r; im[5] = ei[1] - t1i;
        re[2] = er[2] + t2r; im[2] = ei[2] + t2i;
        re[6] = er[2] - t2r; im[6] = ei[2] - t2i;
        re[3] = er[3] + t3r; im[3] = ei[3] + t3i;
        re[7] = er[3] - t3r; im[7] = ei[3] - t3i;
    }
};

// multiply (r,i) by e^{-2*pi*i*K/16} for compile-time K
template <int K>
__device__ __forceinline__ void mul_w16(float& r, float& i) {
    constexpr int k = ((K % 16) + 16) % 16;
    if (k == 0) return;
    float nr, ni;
    if (k == 4) { nr = i; ni = -r; }
    else if (k == 8) { nr = -r; ni = -i; }
    else if (k == 12) { nr = -i; ni = r; }
    else if (k == 2) { nr = (r + i) * FFTCONV_SQRT1_2; ni = (i - r) * FFTCONV_SQRT1_2; }
    else if (k == 6) { nr = (i - r) * FFTCONV_SQRT1_2; ni = -(r + i) * FFTCONV_SQRT1_2; }
    else if (k == 10) { nr = -(r + i) * FFTCONV_SQRT1_2; ni = (r - i) * FFTCONV_SQRT1_2; }
    else if (k == 14) { nr = (r - i) * FFTCONV_SQRT1_2; ni = (r + i) * FFTCONV_SQRT1_2; }
    else {
        // generic: w = c - i*s with c = cos(pi*k/8), s = sin(pi*k/8)
        constexpr float c = (k == 1 || k == 15) ? FFTCONV_COS_PI_8 : (k == 3 || k == 13) ? FFTCONV_SIN_PI_8
                          : (k == 5 || k == 11) ? -FFTCONV_SIN_PI_8 : -FFTCONV_COS_PI_8;   // k = 7, 9
        constexpr float s = (k == 1 || k == 7) ? FFTCONV_SIN_PI_8 : (k == 3 || k == 5) ? FFTCONV_COS_PI_8
                          : (k == 9 || k == 15) ? -FFTCONV_SIN_PI_8 : -FFTCONV_COS_PI_8;   // k = 11, 13
        nr = fmaf(i, s, r * c);
        ni = fmaf(-r, s, i * c);
    }
    r = nr; i = ni;
}

template <> struct Dft<16> {
    // 4x4 Cooley-Tukey: x index = c + 4*a  (c = 0..3 column, a = 0..3), X index = q1 + 4*q2
    __device__ __forceinline__ static void run(float* re, float* im) {
        // stage 1: DFT4 over a for each column c -> y[c][q1] stored at [c + 4*q1]
#pragma unroll
        for (int c = 0; c < 4; ++c)
            dft4(re[c], im[c], re[c + 4], im[c + 4], re[c + 8], im[c + 8], re[c + 12], im[c + 12]);
        // twiddle w16^(c*q1)
        mul_w16<1>(re[1 + 4], im[1 + 4]);  mul_w16<2>(re[2 + 4], im[2 + 4]);  mul_w16<3>(re[3 + 4], im[3 + 4]);
        mul_w16<2>(re[1 + 8], im[1 + 8]);  mul_w16<4>(re[2 + 8], im[2 + 8]);  mul_w16<6>(re[3 + 8], im[3 + 8]);
        mul_w16<3>(re[1 + 12], im[1 + 12]); mul_w16<6>(re[2 + 12], im[2 + 12]); mul_w16<9>(re[3 + 12], im[3 + 12]);
        // stage 2: DFT4 over c for each q1 -> X[q1 + 4*q2] ; in-place result sits at [q2 + 4*q1]
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1)
            dft4(re[4 * q1], im[4 * q1], re[4 * q1 + 1], im[4 * q1 + 1],
                 re[4 * q1 + 2], im[4 * q1 + 2], re[4 * q1 + 3], im[4 * q1 + 3]);
        // transpose 4x4 so that output index q1 + 4*q2 is natural
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) {
                float t = re[4 * a + b]; re[4 * a + b] = re[4 * b + a]; re[4 * b + a] = t;
                t = im[4 * a + b]; im[4 * a + b] = im[4 * b + a]; im[4 * b + a] = t;
            }
    }
};

// ------------------------------------------------------------------ odd radices 3..31
// c_odd_tw[R][j] = (cos(2*pi*j/R), sin(2*pi*j/R)), filled by the host at library init.
__constant__ float2 c_odd_tw[32][32];

template <int R>
struct DftOdd {
    // Symmetric O(R^2/2) real-coefficient form:  s_j = x_j + x_{R-j},  d_j = x_j - x_{R-j}
    //   X[k]   = x0 + sum_j cos(jk) s_j  - i * sum_j sin(jk) d_j
    //   X[R-k] = x0 + sum_j cos(jk) s_j  + i * sum_j sin(jk) d_j
    __device__ __forceinline__ static void run(float* re, float* im) {
        constexpr int H = (R - 1) / 2;
        float sr[H + 1], si[H + 1], dr[H + 1], di[H + 1];
        float x0r = re[0], x0i = im[0];
        float sumr = x0r, sumi = x0i;
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            sr[j] = re[j] + re[R - j]; si[j] = im[j] + im[R - j];
            dr[j] = re[j] - re[R - j]; di[j] = im[j] - im[R - j];
            sumr += sr[j]; sumi += si[j];
        }
        re[0] = sumr; im[0] = sumi;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            float ar = x0r, ai = x0i, br = 0.f, bi = 0.f;
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const float2 t = c_odd_tw[R][(j * k) % R];
                ar = fmaf(t.x, sr[j], ar); ai = fmaf(t.x, si[j], ai);
                br = fmaf(t.y, di[j], br); bi = fmaf(t.y, dr[j], bi);
            }
            // -i*(dr + i di)*sin = (di - i dr)*sin
            re[k] = ar + br;     im[k] = ai - bi;
            re[R - k] = ar - br; im[R - k] = ai + bi;
        }
    }
};
template <> struct Dft<3> : DftOdd<3> {};
template <> struct Dft<5> : DftOdd<5> {};
template <> struct Dft<7> : DftOdd<7> {};
template <> struct Dft<9> : DftOdd<9> {};      // (DftOdd only needs R odd, not prime)
template <> struct Dft<11> : DftOdd<11> {};
template <> struct Dft<13> : DftOdd<13> {};
template <> struct Dft<17> : DftOdd<17> {};

// forward / inverse dispatch on split arrays
template <int R, bool INV>
__device__ __forceinline__ void dft_regs(float* re, float* im) {
    if (INV) Dft<R>::run(im, re); else Dft<R>::run(re, im);
}

}  // namespace fftconv
