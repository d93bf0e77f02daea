// Overlap-save tiling + per-frequency-bin complex GEMM on the 5th-gen tensor cores (tcgen05).
//
// A single image has no dense contraction in the frequency domain (one multiply per bin and channel,
// SURVEY 8d).  Cutting the FH x FW plane into NT overlapping 64 x 64 tiles (overlap-save, S = 65 - maxk
// valid outputs per tile side) turns the tiles into a batch: for every one of the 33*64 frequency bins
//
//      P[bin][template t][tile m] = sum_f  K^[t][f][bin] * D^[m][f][bin]          (complex, F channels)
//
// is a dense (templates x channels) x (channels x tiles) complex GEMM shared by all templates, which is
// what the tensor cores want.  fp32 parity (rel-L2 <= 1e-5) rules out plain TF32, so the real-valued
// form of the product runs as 3xTF32 (hi*hi + hi*lo + lo*hi) with fp32 accumulation in TMEM.
//
//   os_kern_fft                templates -> A operand images: zero pad fused into the load, pruned 2-D 64 x 64
//                              half spectrum in shared memory (run-time residue task: one copy of the transform
//                              code serves all four residues), operand-image store in plain fp32
//   os_data_fft                overlap-save window gather (zero fill, circular wrap) -> 2-D spectrum -> B images
//   os_gemm                    TMA bulk copies -> smem operand images -> tcgen05.mma (kind::tf32,
//                              M=128 templates, N=2*tiles, K=2*F) -> TMEM -> conflict-free padded staging tile -> one TMA
//                              tensor store of P per item (pad columns clipped by the tensor bounds)
//   os_inverse_z               per (template, 4 tiles): TMA tensor copies gather the product spectra into four independent
//                              column zones, 2-D C2R inverse, valid-region store (crop / peak, threshold, top-k detection /
//                              correlation shift / per-level geometry of a pyramid batch fused)
//   os_inverse_tma, os_inverse earlier forms of the same kernel (one mbarrier for all boxes; per-thread cp.async gather):
//                              fallbacks when the 5-D tensor maps / any tensor map cannot be built
//   inv_w_pass                 (spectrum -> plane, when the caller hands in a cudaFFTData spectrum)
//
// Replaces the same reference rows as kernels_tile16.cuh (padData, cufftExecR2C,
// elementwiseProductAndNormalize, F x cufftExecC2R, sumAlongFeatures: src/cudaConvFFTData.cuh:11-92,
// src/cudaConvFFTData.cu:233-271).
//
// Operand images (K-major, no swizzle; one 16-byte unit = 4 consecutive k = channels (2c,2c+1) x (re,im)):
//   Aimg [tblk][bin][ks][kc][128 templates][4]  plain fp32      (MMA A: rows = templates; hi/lo split on chip)
//   Bimg [nblk][bin][ks][kc][NMMA rows     ][4]  plain fp32      (MMA B: rows = (tile, re/im column); hi/lo split on chip)
//   P    [tblk][nblk][bin = u*64 + v][128 templates][RS]  fp32, (re,im) per tile: WORK-ITEM MAJOR -- the 128 x RS block
//        an os_gemm item produces is one contiguous run (one bulk store per item; with the earlier [u][template][v][RS]
//        order an item scattered 128 pieces of RS*4 bytes 64*RS*4 bytes apart and the kernel ran 0.28 instead of 0.21 ms
//        at config 2).  os_inverse_tma gathers boxes {8 floats, 1 template, 64 bins} through a 3-D tensor map
//        {RS, 128 templates, bins}: 32-byte pieces, but the CTAs of neighbouring tile groups / templates that are
//        resident at the same time read the neighbouring pieces, so DRAM still sees whole rows
// Core matrix = 8 rows x 16 B contiguous -> SBO = 128 B, LBO (next 16-byte k unit) = rows * 16 B.
#pragma once
#include <cstdint>
#include <type_traits>

#include "kernels_generic.cuh"
#include "kernels_tile16.cuh"   // mbarrier / bulk-copy helpers

namespace fftconv {

constexpr int OS_T = 64;              // tile side (FFT size)
constexpr int OS_CH = 33;             // half-spectrum rows of a tile
constexpr int OS_NBIN = OS_CH * OS_T; // 2112 frequency bins per tile
constexpr int OS_TM = 128;            // templates per GEMM block (MMA M)
constexpr int OS_ACC_COLS = 128;      // TMEM columns per accumulator buffer (2 buffers)
constexpr int OS_IG = 4;              // tiles per CTA of os_inverse (64 threads per tile): one 32-byte sector of P per bin
constexpr int OS_IGB = 2;             // log2(OS_IG)
constexpr int OS_IROW = 66;           // complex row stride of a 64-point line in shared memory (16-byte aligned rows)
constexpr int OS_ICOL = 33;           // os_inverse: complex stride of a spectrum column (odd: conflict-free 64-bit accesses)
constexpr int OS_ITILE = 64 * OS_ICOL + 8;   // os_inverse: tile stride (columns >= 32 sit 4 elements further; +32 B per tile):
                                             // the 16 lanes of a gather half-warp (2 tiles x 2 column halves x 4 columns) hit 32 distinct banks

// ------------------------------------------------------------------------------------------------
// compile-time twiddles  w64^e = cos(2 pi e/64) - i sin(2 pi e/64)  (double Taylor series, folded to
// immediates: the task transforms below are fully unrolled with compile-time indices)
constexpr double OS_PI = 3.14159265358979323846264338327950288;
__host__ __device__ constexpr double os_sin_taylor(double x) {
    double term = x, sum = x;
    for (int n = 1; n < 18; ++n) { term *= -x * x / ((2.0 * n) * (2.0 * n + 1.0)); sum += term; }
    return sum;
}
__host__ __device__ constexpr double os_snap(double v) { return (v < 1e-13 && v > -1e-13) ? 0.0 : v; }
__host__ __device__ constexpr double os_sin64d(int e) {
    double a = 2.0 * OS_PI * (double)e / 64.0;
    if (a > OS_PI) a -= 2.0 * OS_PI;
    return os_snap(os_sin_taylor(a));
}
__host__ __device__ constexpr double os_cos64d(int e) { return os_sin64d((e + 16) & 63); }
template <int E> struct OsW64 {
    static constexpr float c = (float)os_cos64d(E & 63);
    static constexpr float s = (float)os_sin64d(E & 63);
};

template <int B, int E, class Fn>
__device__ __forceinline__ void os_static_for(Fn&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        os_static_for<B + 1, E>(f);
    }
}

// acc (+)= (x + i y) * (-i)^K   (forward)   or   * (+i)^K   (inverse)
template <int K, bool INV, bool FIRST>
__device__ __forceinline__ void os_rot_acc(float x, float y, float& sr, float& si) {
    constexpr int k = INV ? ((4 - (K & 3)) & 3) : (K & 3);
    float rx, ry;
    if (k == 0) { rx = x; ry = y; } else if (k == 1) { rx = y; ry = -x; } else if (k == 2) { rx = -x; ry = -y; } else { rx = -y; ry = x; }
    if (FIRST) { sr = rx; si = ry; } else { sr += rx; si += ry; }
}
template <int E, bool INV>
__device__ __forceinline__ void os_twiddle(float sr, float si, float& orr, float& oi) {
    constexpr int e = E & 63;
    if (e == 0) { orr = sr; oi = si; return; }
    constexpr float c = OsW64<e>::c, s = OsW64<e>::s;
    if (!INV) { orr = fmaf(si, s, sr * c); oi = fmaf(-sr, s, si * c); }      // (sr + i si)(c - i s)
    else      { orr = fmaf(-si, s, sr * c); oi = fmaf(sr, s, si * c); }      // (sr + i si)(c + i s)
}

// One "task" of a 64-point transform of a sequence with 16*NF leading non-zero samples:
// the 16 outputs  X[4*j1 + R0], j1 = 0..15.
//     z[c] = w64^{R0*c} * sum_{q<NF} x[c + 16 q] * w4^{R0*q},      X[4*j1 + R0] = DFT16(z)[j1]
// ld(j) (j even) returns samples j and j+1 as (re, im, re, im).
template <int R0, int NF, bool INV, class Load>
__device__ __forceinline__ void os_fft64_task(Load&& ld, float* re, float* im) {
    os_static_for<0, 8>([&](auto c2c) {
        constexpr int c2 = decltype(c2c)::value;
        float sr0 = 0.f, si0 = 0.f, sr1 = 0.f, si1 = 0.f;
        os_static_for<0, NF>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const float4 v = ld(2 * c2 + 16 * q);
            os_rot_acc<R0 * q, INV, q == 0>(v.x, v.y, sr0, si0);
            os_rot_acc<R0 * q, INV, q == 0>(v.z, v.w, sr1, si1);
        });
        os_twiddle<R0 * (2 * c2), INV>(sr0, si0, re[2 * c2], im[2 * c2]);
        os_twiddle<R0 * (2 * c2 + 1), INV>(sr1, si1, re[2 * c2 + 1], im[2 * c2 + 1]);
    });
    dft_regs<16, INV>(re, im);
}
// The same task with the residue r0 as a RUN-TIME, warp-uniform value: the twiddles w64^{r0 c} come from a constant-memory
// table (one uniform constant load each) instead of being folded into the instruction stream.  os_kern_fft used to
// inline 12 compile-time specialisations of the task (10 128 instructions = 162 KB of SASS, more than the instruction
// cache holds: `no_instruction` was its top stall at 7 warps per issue cycle); with this form it inlines 3.
// c_os_w64[r0][c] = (cos, sin)(2 pi r0 c / 64), filled by the host from the same constexpr series (os_cos64d / os_sin64d).
__constant__ float2 c_os_w64[4][16];

template <int NF, class Load>
__device__ __forceinline__ void os_fft64_task_dyn(int r0, Load&& ld, float* re, float* im) {
    const float2* tw = c_os_w64[r0];
    os_static_for<0, 8>([&](auto c2c) {
        constexpr int c2 = decltype(c2c)::value;
        float sr0 = 0.f, si0 = 0.f, sr1 = 0.f, si1 = 0.f;
        os_static_for<0, NF>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const float4 v = ld(2 * c2 + 16 * q);
            if (q == 0) { sr0 = v.x; si0 = v.y; sr1 = v.z; si1 = v.w; }
            else {                                   // (x + i y) (-i)^(r0 q), q = 1: k = r0
                const int k = (r0 * q) & 3;
                const float ax = (k & 1) ? v.y : v.x, ay = (k & 1) ? -v.x : v.y;      // k odd: (y, -x)
                const float bx = (k & 1) ? v.w : v.z, by = (k & 1) ? -v.z : v.w;
                const float sg = (k & 2) ? -1.f : 1.f;                                    // k = 2: (-x, -y); k = 3: (-y, x) = -(y, -x)
                sr0 += sg * ax; si0 += sg * ay; sr1 += sg * bx; si1 += sg * by;
            }
        });
        const float2 w0 = tw[2 * c2], w1 = tw[2 * c2 + 1];                             // (sr + i si)(c - i s)
        re[2 * c2] = fmaf(si0, w0.y, sr0 * w0.x);     im[2 * c2] = fmaf(-sr0, w0.y, si0 * w0.x);
        re[2 * c2 + 1] = fmaf(si1, w1.y, sr1 * w1.x); im[2 * c2 + 1] = fmaf(-sr1, w1.y, si1 * w1.x);
    });
    dft_regs<16, false>(re, im);
}

// r0 is warp-uniform at every call site, so the switch does not diverge
template <int NF, bool INV, class Load>
__device__ __forceinline__ void os_fft64_task_rt(int r0, Load&& ld, float* re, float* im) {
    switch (r0) {
        case 0: os_fft64_task<0, NF, INV>(ld, re, im); break;
        case 1: os_fft64_task<1, NF, INV>(ld, re, im); break;
        case 2: os_fft64_task<2, NF, INV>(ld, re, im); break;
        default: os_fft64_task<3, NF, INV>(ld, re, im); break;
    }
}

// Two tasks of a full 64-sample sequence that share their loads and the first radix-4 butterfly:
// outputs X[4*j1 + P] -> (reA, imA) and X[4*j1 + P + 2] -> (reB, imB), P = 0 or 1.
//     P = 0:  a = x0 + x2, b = x1 + x3:  z_0 = a + b,            z_2 = (a - b) w64^{2c}
//     P = 1:  a = x0 - x2, b = x1 - x3:  z_1 = (a -+ i b) w64^c, z_3 = (a +- i b) w64^{3c}    (x_q = x[c + 16 q])
template <int P, bool INV, class Load, bool DFT = true>     // DFT = false: the first radix-4 stage only (the caller runs the 16-point transforms)
__device__ __forceinline__ void os_fft64_pair(Load&& ld, float* reA, float* imA, float* reB, float* imB) {
    os_static_for<0, 8>([&](auto c2c) {
        constexpr int c2 = decltype(c2c)::value;
        const float4 v0 = ld(2 * c2), v1 = ld(2 * c2 + 16), v2 = ld(2 * c2 + 32), v3 = ld(2 * c2 + 48);
        os_static_for<0, 2>([&](auto hc) {
            constexpr int h = decltype(hc)::value;
            constexpr int c = 2 * c2 + h;
            const float x0r = h ? v0.z : v0.x, x0i = h ? v0.w : v0.y;
            const float x1r = h ? v1.z : v1.x, x1i = h ? v1.w : v1.y;
            const float x2r = h ? v2.z : v2.x, x2i = h ? v2.w : v2.y;
            const float x3r = h ? v3.z : v3.x, x3i = h ? v3.w : v3.y;
            if (P == 0) {
                const float ar = x0r + x2r, ai = x0i + x2i, br = x1r + x3r, bi = x1i + x3i;
                reA[c] = ar + br; imA[c] = ai + bi;
                os_twiddle<2 * c, INV>(ar - br, ai - bi, reB[c], imB[c]);
            } else {
                const float ar = x0r - x2r, ai = x0i - x2i, br = x1r - x3r, bi = x1i - x3i;
                // forward: z1 = a - i b, z3 = a + i b ; inverse: the other way round
                const float mr = ar + bi, mi = ai - br;     // a - i b
                const float pr = ar - bi, pi = ai + br;     // a + i b
                os_twiddle<c, INV>(INV ? pr : mr, INV ? pi : mi, reA[c], imA[c]);
                os_twiddle<3 * c, INV>(INV ? mr : pr, INV ? mi : pi, reB[c], imB[c]);
            }
        });
    });
    if (DFT) {
        dft_regs<16, INV>(reA, imA);
        dft_regs<16, INV>(reB, imB);
    }
}

__device__ __forceinline__ int os_wrap(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

__device__ __forceinline__ float os_tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ------------------------------------------------------------------------------------------------
// os_kern_fft: templates -> A operand images in ONE kernel (pad fused into the load, 2-D 64 x 64 half
// spectrum of a 16*NF x 16*NF support, operand-image store in plain fp32 -- os_gemm splits hi/lo on chip).  Replaces the template
// legs of os_hpass + os_wpass and their [template][F][33][XC] intermediate.
// CTA = 16 templates x one channel pair, 256 threads.  Two phases (odd spectrum rows, then even ones) so
// that the h-transformed rows of the 32 planes fit 70 KB (NF=1) of shared memory:
//   h step: thread = (plane, column pair): two real columns ride as one complex sequence; the two tasks
//           that hold rows u and 64-u are computed together and unpacked in registers
//   w step: thread = (spectrum row, task r0, template); both channels of the pair, then 16 128-bit stores;
//           16 consecutive lanes write 256 contiguous bytes of the image
// grid = (ntblk*128/16, NKS*KC).
struct OsKArgs {
    const SrcDesc* descs;
    int nk, F;
    float* img;
    int NKS, KC;
    int flip;               // correlation mode: the template is read flipped in h and w (its own extent)
};
constexpr int OS_KSL = 16;            // templates per CTA

template <int NF>
__global__ void __launch_bounds__(256, 2) os_kern_fft(OsKArgs a)
{
    constexpr int XC = 16 * NF, NCP = XC / 2;
    constexpr int PS = 17 * XC + 2;                    // plane stride in cpx (+16 B: conflict-free LDS.128 across templates)
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx* Hs = reinterpret_cast<cpx*>(os_smem_raw);     // [2 channels][16 templates][17 rows][XC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fp = blockIdx.y;
    const int slot0 = blockIdx.x * OS_KSL;
    const int tblk = slot0 / OS_TM, sl0 = slot0 - tblk * OS_TM;
    const int ks = fp / a.KC, kc = fp - ks * a.KC;
    const size_t stage_floats = (size_t)a.KC * OS_TM * 4;           // one K-stage of one bin
    const size_t bin_stride = (size_t)a.NKS * stage_floats;

    // NF = 1: the 32 zero-padded (and, in correlation mode, flipped) 16 x 16 template planes of this CTA are staged in
    // shared memory once, by asynchronous 4-byte copies (all 32 of a thread in flight); both phases read them from there
    constexpr bool STAGED = NF == 1;
    constexpr int RCOL = 18, RPL = 16 * RCOL + 2;      // raw column / plane stride (floats): conflict-free 64-bit reads
    float* raw = reinterpret_cast<float*>(os_smem_raw + (size_t)32 * PS * sizeof(cpx));   // [32 planes][RPL]
    if (STAGED) {
        const int y = threadIdx.x & 15, x = threadIdx.x >> 4;
#pragma unroll 4
        for (int pl = 0; pl < 32; ++pl) {
            const int item = slot0 + (pl & 15), f = 2 * fp + (pl >> 4);
            float* dst = raw + pl * RPL + x * RCOL + y;
            bool issued = false;
            if (item < a.nk && f < a.F) {
                const SrcDesc d = a.descs[item];
                if (x < d.cols && y < d.rows) {
                    const int sx = a.flip ? d.cols - 1 - x : x, sy = a.flip ? d.rows - 1 - y : y;
                    const float* src = d.ptr + ((size_t)f * d.cols + sx) * d.rows + sy;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
                    issued = true;
                }
            }
            if (!issued) *dst = 0.f;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }

#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
        // ---- h step
        for (int line = threadIdx.x; line < 32 * NCP; line += 256) {
            const int cp = line % NCP, pl = line / NCP;
            const int ch = pl >> 4, slot = pl & 15;
            const int item = slot0 + slot, f = 2 * fp + ch;
            float4* hrow = reinterpret_cast<float4*>(Hs + (size_t)pl * PS) + cp;      // + ro * (XC/2)
            if (item < a.nk && f < a.F) {
                // direct loads (NF = 2): flipped read = element (rows-1-j, cols-1-x) of the template
                SrcDesc d{};
                if (!STAGED) d = a.descs[item];
                const int xa = 2 * cp, xb = xa + 1;
                const bool va = xa < d.cols, vb = xb < d.cols;
                const int rows = d.rows;
                const float* pa = STAGED ? nullptr : d.ptr + ((size_t)f * d.cols + (a.flip ? d.cols - 1 - xa : xa)) * d.rows + (a.flip ? rows - 1 : 0);
                const float* pb = STAGED ? nullptr : d.ptr + ((size_t)f * d.cols + (a.flip ? d.cols - 1 - xb : xb)) * d.rows + (a.flip ? rows - 1 : 0);
                const int sj = a.flip ? -1 : 1;
                const float* ca = raw + pl * RPL + xa * RCOL;          // staged (NF = 1): two 64-bit reads per sample pair
                auto ld = [&](int j) {
                    float4 v;
                    if (STAGED) {
                        const float2 p = *reinterpret_cast<const float2*>(ca + j), q = *reinterpret_cast<const float2*>(ca + RCOL + j);
                        v = make_float4(p.x, q.x, p.y, q.y);
                    } else {
                        v.x = (va && j < rows) ? __ldg(pa + sj * j) : 0.f;
                        v.y = (vb && j < rows) ? __ldg(pb + sj * j) : 0.f;
                        v.z = (va && j + 1 < rows) ? __ldg(pa + sj * (j + 1)) : 0.f;
                        v.w = (vb && j + 1 < rows) ? __ldg(pb + sj * (j + 1)) : 0.f;
                    }
                    return v;
                };
                float pr[16], pi[16], qr[16], qi[16];
                // A[u] = (Z[u] + conj Z[64-u]) / 2 (column xa),  B[u] = -i (Z[u] - conj Z[64-u]) / 2 (column xb)
                auto put = [&](int ro, float zr, float zi, float nr, float ni) {
                    hrow[ro * NCP] = make_float4(0.5f * (zr + nr), 0.5f * (zi - ni), 0.5f * (zi + ni), -0.5f * (zr - nr));
                };
                os_fft64_task_dyn<NF>(ph == 0 ? 1 : 0, ld, pr, pi);   // Z[4 j1 + 1]  |  Z[4 j1]
                os_fft64_task_dyn<NF>(ph == 0 ? 3 : 2, ld, qr, qi);   // Z[4 j1 + 3]  |  Z[4 j1 + 2]
                if (ph == 0) {
#pragma unroll
                    for (int j1 = 0; j1 < 8; ++j1) {
                        put(2 * j1, pr[j1], pi[j1], qr[15 - j1], qi[15 - j1]);          // u = 4 j1 + 1
                        put(2 * j1 + 1, qr[j1], qi[j1], pr[15 - j1], pi[15 - j1]);      // u = 4 j1 + 3
                    }
                } else {
#pragma unroll
                    for (int j1 = 0; j1 < 9; ++j1) put(2 * j1, pr[j1], pi[j1], pr[(16 - j1) & 15], pi[(16 - j1) & 15]);   // u = 4 j1
#pragma unroll
                    for (int j1 = 0; j1 < 8; ++j1) put(2 * j1 + 1, qr[j1], qi[j1], qr[15 - j1], qi[15 - j1]);              // u = 4 j1 + 2
                }
            } else {
                const int nr = ph == 0 ? 16 : 17;
                for (int ro = 0; ro < nr; ++ro) hrow[ro * NCP] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        // ---- w step
        const int NR = ph == 0 ? 16 : 17;
        const int njobs = ((NR + 1) >> 1) * 4;
        for (int c2 = warp; c2 < njobs; c2 += 8) {
            const int r0 = c2 & 3;
            const int ro = (c2 >> 2) * 2 + (lane >> 4);
            const int slot = lane & 15;
            if (ro >= NR) continue;
            const int u = ph == 0 ? 2 * ro + 1 : 2 * ro;
            float* base = a.img + ((size_t)tblk * OS_NBIN + (size_t)u * 64 + r0) * bin_stride + (size_t)ks * stage_floats +
                          (size_t)kc * OS_TM * 4 + (size_t)(sl0 + slot) * 4;
            float re0[16], im0[16], re1[16], im1[16];
            {
                const float4* row = reinterpret_cast<const float4*>(Hs + (size_t)slot * PS + ro * XC);
                auto ld = [&](int j) { return row[j >> 1]; };
                os_fft64_task_dyn<NF>(r0, ld, re0, im0);
            }
            {
                const float4* row = reinterpret_cast<const float4*>(Hs + (size_t)(16 + slot) * PS + ro * XC);
                auto ld = [&](int j) { return row[j >> 1]; };
                os_fft64_task_dyn<NF>(r0, ld, re1, im1);
            }
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1)
                *reinterpret_cast<float4*>(base + (size_t)(4 * j1) * bin_stride) = make_float4(re0[j1], im0[j1], re1[j1], im1[j1]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// os_data_fft: source plane -> B operand images in ONE kernel.  CTA = (overlap-save tile, channel pair):
// the two 64 x 64 windows are gathered with coalesced loads (zero fill beyond the source, circular wrap at
// FH x FW), transformed along h (two real columns per complex sequence, two threads per line) and along w
// in shared memory, split hi/lo and stored as the (re-row, im-row) pair of the tile in the B image.
// grid = NT * NKS*KC (channel pair fastest), 128 threads.  NT = images * tiles per image (batched calls: the images only add tiles).
// Per-level geometry of a PYRAMID batch (fftconv_conv_pyramid): the levels of a feature pyramid differ in size, so the
// "images" of the batch no longer share one tile grid.  Tiles are numbered level-major; level l owns tiles
// m0 .. m0 + nth*ntw - 1.  With a table the uniform fields (NTimg, nth, FH, FW, crop, out_ld, src) are ignored.
struct OsLevel {
    const float* src;       // [F][cols][rows] on the device (raw level, or the plane recovered from its spectrum)
    int rows, cols;
    int FH, FW;             // plane of the level (computeFFTsize16 of size + maxK - 1)
    int nth, m0;            // tile rows of the level's grid, first tile
    int crop_h, crop_w, out_ld;
};
constexpr int OS_MAX_LEVELS = 64;
__device__ __forceinline__ int os_level_of(const OsLevel* __restrict__ lv, int nlevels, int m) {
    int l = 0;
    while (l + 1 < nlevels && m >= lv[l + 1].m0) ++l;
    return l;
}

struct OsDArgs {
    const OsLevel* levels;  // pyramid batch: per-level geometry (device), else nullptr
    int nlevels;
    SrcDesc src;            // image 0: [F][cols][rows]; image n follows at n*F*cols*rows
    int F, nth, NTimg, Sh, Sw, oy0, ox0, FH, FW;
    float* img;
    int NKS, KC, NMMA, NTn;
    int correlate;
    int st256;              // one 256-bit store per (tile, channel pair, bin) instead of two 128-bit ones (FFTCONV_OS_DATA_ST256)
    // provenance of a caller-supplied spectrum (SpecCache in fftconv.cu): hsel[0] = hash of the spectrum when
    // fftconv_fft_data produced it, hsel[1] = hash of what the caller handed to the convolution.  Equal: the raw data
    // kept from that call (`alt`) is tiled directly -- or, when the tile spectra were already computed next to the
    // forward transform (alt_done), nothing is left to do.  Different: `src` (the plane recovered from the spectrum).
    const unsigned long long* hsel;
    SrcDesc alt;
    int alt_done;
};
// 64-bit position-sensitive hash of a buffer of 8-byte words (wrapping sum of mixed (index, word) pairs: the order of the
// additions does not matter, so blocks combine with one atomicAdd each).  *out must be zero before the launch.
__device__ __forceinline__ unsigned long long os_mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) os_hash64(const unsigned long long* __restrict__ p, size_t n, unsigned long long* out) {
    unsigned long long h = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        h += os_mix64(p[i] + 0x9E3779B97F4A7C15ull * (unsigned long long)(2 * i + 1));
#pragma unroll
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        atomicAdd(out, t);
    }
}
constexpr int OS_DRAW = 65;           // raw window column stride (floats)
constexpr size_t OS_DATA_SMEM = 2 * 64 * OS_DRAW * sizeof(float) + 2 * 33 * OS_IROW * sizeof(cpx);

__global__ void __launch_bounds__(128) os_data_fft(OsDArgs a)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    float* raw = reinterpret_cast<float*>(os_smem_raw);                              // [2][64][65]
    cpx* Hs = reinterpret_cast<cpx*>(os_smem_raw + 2 * 64 * OS_DRAW * sizeof(float));   // [2][33][66]
    SrcDesc src = a.src;
    if (a.hsel && a.hsel[0] == a.hsel[1]) {                             // CTA-uniform (see OsDArgs)
        if (a.alt_done) return;
        src = a.alt;
    }
    const int npair = a.NKS * a.KC;
    const int m = blockIdx.x / npair, fp = blockIdx.x - m * npair;      // channel pair fastest
    int img = m / a.NTimg, mt = m - img * a.NTimg;                // tiles of a batch are numbered image-major
    int nth = a.nth, FH = a.FH, FW = a.FW;
    if (a.levels) {                                               // pyramid batch: the level's own plane and tile grid
        const OsLevel lv = a.levels[os_level_of(a.levels, a.nlevels, m)];
        src.ptr = lv.src; src.rows = lv.rows; src.cols = lv.cols;
        FH = lv.FH; FW = lv.FW; nth = lv.nth; mt = m - lv.m0; img = 0;
    }
    const int tj = mt / nth, ti = mt - tj * nth;
    const int oy = ti * a.Sh - a.oy0, ox = tj * a.Sw - a.ox0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- gather: thread = (row y, column parity); lanes run along h (contiguous in the source).  Asynchronous 4-byte
    // copies straight into shared memory: all 64 of a thread are in flight at once and cost no registers (the register-
    // staged form ran 4 dependent rounds of 16 loads: the kernel's top stall was long_scoreboard at 4.1 warps per issue)
    {
        const int y = threadIdx.x & 63;
        const int gy = os_wrap(oy + y, FH);
        const bool vy = gy < src.rows;
        for (int ch = 0; ch < 2; ++ch) {
            const int f = 2 * fp + ch;
            const float* pl = src.ptr + ((size_t)img * a.F + f) * src.cols * src.rows + gy;
            float* dst = raw + (size_t)ch * 64 * OS_DRAW + y + (threadIdx.x >> 6) * OS_DRAW;
            int gx = os_wrap(ox + (threadIdx.x >> 6), FW);
            const bool vf = vy && f < a.F;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                if (vf && gx < src.cols)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(pl + (size_t)gx * src.rows) : "memory");
                else
                    *dst = 0.f;
                dst += 2 * OS_DRAW;
                gx += 2;
                if (gx >= FW) gx -= FW;                             // FW >= 16 > 2: one conditional subtract is enough
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    const int par = warp >> 1;                       // warp-uniform: tasks (par, par + 2)
    // ---- h step: line = (channel, column pair)
    {
        const int line = threadIdx.x & 63;
        const int ch = line >> 5, cp = line & 31;
        const float* ca = raw + ((size_t)ch * 64 + 2 * cp) * OS_DRAW;
        const float* cb = ca + OS_DRAW;
        auto ld = [&](int j) { return make_float4(ca[j], cb[j], ca[j + 1], cb[j + 1]); };
        float pr[16], pi[16], qr[16], qi[16];
        float4* hcol = reinterpret_cast<float4*>(Hs + (size_t)ch * 33 * OS_IROW) + cp;       // + u * (OS_IROW/2)
        auto put = [&](int u, float zr, float zi, float nr, float ni) {
            hcol[u * (OS_IROW / 2)] = make_float4(0.5f * (zr + nr), 0.5f * (zi - ni), 0.5f * (zi + ni), -0.5f * (zr - nr));
        };
        if (par == 0) {
            os_fft64_pair<0, false>(ld, pr, pi, qr, qi);              // Z[4 j1], Z[4 j1 + 2]
#pragma unroll
            for (int j1 = 0; j1 < 9; ++j1) put(4 * j1, pr[j1], pi[j1], pr[(16 - j1) & 15], pi[(16 - j1) & 15]);
#pragma unroll
            for (int j1 = 0; j1 < 8; ++j1) put(4 * j1 + 2, qr[j1], qi[j1], qr[15 - j1], qi[15 - j1]);
        } else {
            os_fft64_pair<1, false>(ld, pr, pi, qr, qi);              // Z[4 j1 + 1], Z[4 j1 + 3]
#pragma unroll
            for (int j1 = 0; j1 < 8; ++j1) {
                put(4 * j1 + 1, pr[j1], pi[j1], qr[15 - j1], qi[15 - j1]);
                put(4 * j1 + 3, qr[j1], qi[j1], pr[15 - j1], pi[15 - j1]);
            }
        }
    }
    __syncthreads();
    // ---- w step: thread = (spectrum row u, task pair); both channels, then the operand-image stores.
    // (Tried: warp = (channel, task pair), lane = row, rows 0 and 32 packed into one complex transform -- all 128 threads busy
    // instead of 66, but each thread then owns 8 of the 16 bytes of an operand unit: twice the store instructions, half-written
    // sectors meeting in L2.  Config 4 at quarter scale 3.27 -> 5.35 ms, C2 45 -> 62 us: the kernel is bound by its stores.)
    // (Tried, round 2e: rows 0 and 32 packed as H[0] + i H[32] in row 0, lane 0 leaves the transform in shared memory and warps
    // 1 and 3 -- which otherwise run this whole step for ONE lane each -- only unpack and store the two rows: a quarter fewer
    // warp instructions, 153 registers instead of 167, parity green, and no change in time (config 4: 13.3 vs 13.2 ms, C2 53 vs
    // 50 us): the kernel waits for its window gather and its stores, not for issue slots.  scripts/exp_datafft.sh.)
    {
        const int u = threadIdx.x & 63;
        if (u < OS_CH) {
            const int nblk = m / a.NTn, sl = m - nblk * a.NTn;
            const int ks = fp / a.KC, kc = fp - ks * a.KC;
            const size_t stage_stride = (size_t)a.KC * a.NMMA * 4;
            const size_t bin_stride = (size_t)a.NKS * stage_stride;
            float* base = a.img + ((size_t)nblk * OS_NBIN + (size_t)u * 64 + par) * bin_stride + (size_t)ks * stage_stride +
                          (size_t)kc * a.NMMA * 4 + (size_t)(2 * sl) * 4;
            float r0A[16], i0A[16], r0B[16], i0B[16], r1A[16], i1A[16], r1B[16], i1B[16];
            {
                const cpx* row = Hs + (size_t)u * OS_IROW;
                auto ld = [&](int j) { return *reinterpret_cast<const float4*>(row + j); };
                if (par == 0) os_fft64_pair<0, false>(ld, r0A, i0A, r0B, i0B);
                else          os_fft64_pair<1, false>(ld, r0A, i0A, r0B, i0B);
            }
            {
                const cpx* row = Hs + (size_t)(33 + u) * OS_IROW;
                auto ld = [&](int j) { return *reinterpret_cast<const float4*>(row + j); };
                if (par == 0) os_fft64_pair<0, false>(ld, r1A, i1A, r1B, i1B);
                else          os_fft64_pair<1, false>(ld, r1A, i1A, r1B, i1B);
            }
            // out_re = sum a*c - b*d, out_im = sum a*d + b*c  with K^ = a + i b (A rows), D^ = c + i d
            // correlate: K^ -> conj(K^):  out_re = sum a*c + b*d, out_im = sum a*d - b*c
            // the 1/(64*64) of the inverse transform rides here: a power of two, so the scaling is exact
            const float sc = 1.0f / 4096.0f, sg = a.correlate ? -sc : sc;
            // plain fp32 (re-row, im-row): os_gemm splits hi / lo on chip, as it does for the A images
            const bool st256 = a.st256 != 0;
            auto emit = [&](float* o, float c0x, float c0y, float c1x, float c1y) {
                if (st256) {                       // the 32 bytes of a (tile, channel pair, bin) unit as ONE whole-sector store
                    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(sc * c0x), "f"(-sg * c0y),
                                 "f"(sc * c1x), "f"(-sg * c1y), "f"(sc * c0y), "f"(sg * c0x), "f"(sc * c1y), "f"(sg * c1x) : "memory");
                    return;
                }
                float4* oh = reinterpret_cast<float4*>(o);
                oh[0] = make_float4(sc * c0x, -sg * c0y, sc * c1x, -sg * c1y);
                oh[1] = make_float4(sc * c0y, sg * c0x, sc * c1y, sg * c1x);
            };
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) {
                emit(base + (size_t)(4 * j1) * bin_stride, r0A[j1], i0A[j1], r1A[j1], i1A[j1]);
                emit(base + (size_t)(4 * j1 + 2) * bin_stride, r0B[j1], i0B[j1], r1B[j1], i1B[j1]);
            }
        }
    }
}

// os_data_fft_occ: the same transform laid out for occupancy (FFTCONV_OS_DATA=2, experiment of round 2e).  os_data_fft is
// latency-bound (ncu at config-4 shapes: long_scoreboard on the window gather, 31 % issue active) and holds three CTAs per
// SM: 167 registers (the w step keeps the spectra of BOTH channels of a row, 128 floats, in registers) and 68 KB.  Here
//   * the h-transformed rows overlay the raw windows: a barrier separates the first radix-4 stage of the h step (which
//     consumes every raw sample) from the register DFTs and their stores;
//   * the w step runs the two channels one after the other: the spectrum of channel 0 is parked IN PLACE in its own row
//     (a barrier lets both parities finish reading the row first) and read back while channel 1 is emitted.
// 35 KB of shared memory and about half the live registers.
constexpr size_t OS_DATA_SMEM_OCC = (2 * 64 * OS_DRAW * sizeof(float) > 2 * 33 * OS_IROW * sizeof(cpx))
                                        ? 2 * 64 * OS_DRAW * sizeof(float) : 2 * 33 * OS_IROW * sizeof(cpx);

__global__ void __launch_bounds__(128, 5) os_data_fft_occ(OsDArgs a)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    float* raw = reinterpret_cast<float*>(os_smem_raw);                              // [2][64][65]
    cpx* Hs = reinterpret_cast<cpx*>(os_smem_raw);                                   // [2][33][66], overlays raw
    SrcDesc src = a.src;
    if (a.hsel && a.hsel[0] == a.hsel[1]) {                             // CTA-uniform (see OsDArgs)
        if (a.alt_done) return;
        src = a.alt;
    }
    const int npair = a.NKS * a.KC;
    const int m = blockIdx.x / npair, fp = blockIdx.x - m * npair;      // channel pair fastest
    int img = m / a.NTimg, mt = m - img * a.NTimg;
    int nth = a.nth, FH = a.FH, FW = a.FW;
    if (a.levels) {
        const OsLevel lv = a.levels[os_level_of(a.levels, a.nlevels, m)];
        src.ptr = lv.src; src.rows = lv.rows; src.cols = lv.cols;
        FH = lv.FH; FW = lv.FW; nth = lv.nth; mt = m - lv.m0; img = 0;
    }
    const int tj = mt / nth, ti = mt - tj * nth;
    const int oy = ti * a.Sh - a.oy0, ox = tj * a.Sw - a.ox0;
    const int warp = threadIdx.x >> 5;
    {   // ---- gather (as in os_data_fft)
        const int y = threadIdx.x & 63;
        const int gy = os_wrap(oy + y, FH);
        const bool vy = gy < src.rows;
        for (int ch = 0; ch < 2; ++ch) {
            const int f = 2 * fp + ch;
            const float* pl = src.ptr + ((size_t)img * a.F + f) * src.cols * src.rows + gy;
            float* dst = raw + (size_t)ch * 64 * OS_DRAW + y + (threadIdx.x >> 6) * OS_DRAW;
            int gx = os_wrap(ox + (threadIdx.x >> 6), FW);
            const bool vf = vy && f < a.F;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                if (vf && gx < src.cols)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(pl + (size_t)gx * src.rows) : "memory");
                else
                    *dst = 0.f;
                dst += 2 * OS_DRAW;
                gx += 2;
                if (gx >= FW) gx -= FW;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    const int par = warp >> 1;                       // warp-uniform: tasks (par, par + 2)
    {   // ---- h step: line = (channel, column pair)
        const int line = threadIdx.x & 63;
        const int ch = line >> 5, cp = line & 31;
        const float* ca = raw + ((size_t)ch * 64 + 2 * cp) * OS_DRAW;
        const float* cb = ca + OS_DRAW;
        auto ld = [&](int j) { return make_float4(ca[j], cb[j], ca[j + 1], cb[j + 1]); };
        float pr[16], pi[16], qr[16], qi[16];
        if (par == 0) os_fft64_pair<0, false, decltype(ld)&, false>(ld, pr, pi, qr, qi);
        else          os_fft64_pair<1, false, decltype(ld)&, false>(ld, pr, pi, qr, qi);
        __syncthreads();                             // every raw sample has been consumed: the rows may overwrite the windows
        dft_regs<16, false>(pr, pi);
        dft_regs<16, false>(qr, qi);
        float4* hcol = reinterpret_cast<float4*>(Hs + (size_t)ch * 33 * OS_IROW) + cp;       // + u * (OS_IROW/2)
        auto put = [&](int u, float zr, float zi, float nr, float ni) {
            hcol[u * (OS_IROW / 2)] = make_float4(0.5f * (zr + nr), 0.5f * (zi - ni), 0.5f * (zi + ni), -0.5f * (zr - nr));
        };
        if (par == 0) {
#pragma unroll
            for (int j1 = 0; j1 < 9; ++j1) put(4 * j1, pr[j1], pi[j1], pr[(16 - j1) & 15], pi[(16 - j1) & 15]);
#pragma unroll
            for (int j1 = 0; j1 < 8; ++j1) put(4 * j1 + 2, qr[j1], qi[j1], qr[15 - j1], qi[15 - j1]);
        } else {
#pragma unroll
            for (int j1 = 0; j1 < 8; ++j1) {
                put(4 * j1 + 1, pr[j1], pi[j1], qr[15 - j1], qi[15 - j1]);
                put(4 * j1 + 3, qr[j1], qi[j1], pr[15 - j1], pi[15 - j1]);
            }
        }
    }
    __syncthreads();
    {   // ---- w step: thread = (spectrum row u, task pair); channel 0 first, parked in place, then channel 1 and the stores
        const int u = threadIdx.x & 63;
        const bool active = u < OS_CH;
        cpx* row0 = Hs + (size_t)(active ? u : 0) * OS_IROW;
        float rA[16], iA[16], rB[16], iB[16];
        if (active) {
            auto ld = [&](int j) { return *reinterpret_cast<const float4*>(row0 + j); };
            if (par == 0) os_fft64_pair<0, false, decltype(ld)&, false>(ld, rA, iA, rB, iB);
            else          os_fft64_pair<1, false, decltype(ld)&, false>(ld, rA, iA, rB, iB);
        }
        __syncthreads();                             // both parities have read row u of channel 0
        if (active) {
            dft_regs<16, false>(rA, iA);
            dft_regs<16, false>(rB, iB);
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) {
                row0[4 * j1 + par] = make_float2(rA[j1], iA[j1]);
                row0[4 * j1 + par + 2] = make_float2(rB[j1], iB[j1]);
            }
            {
                const cpx* row = Hs + (size_t)(33 + u) * OS_IROW;
                auto ld = [&](int j) { return *reinterpret_cast<const float4*>(row + j); };
                if (par == 0) os_fft64_pair<0, false>(ld, rA, iA, rB, iB);
                else          os_fft64_pair<1, false>(ld, rA, iA, rB, iB);
            }
            const int nblk = m / a.NTn, sl = m - nblk * a.NTn;
            const int ks = fp / a.KC, kc = fp - ks * a.KC;
            const size_t stage_stride = (size_t)a.KC * a.NMMA * 4;
            const size_t bin_stride = (size_t)a.NKS * stage_stride;
            float* base = a.img + ((size_t)nblk * OS_NBIN + (size_t)u * 64 + par) * bin_stride + (size_t)ks * stage_stride +
                          (size_t)kc * a.NMMA * 4 + (size_t)(2 * sl) * 4;
            const float sc = 1.0f / 4096.0f, sg = a.correlate ? -sc : sc;
            const bool st256 = a.st256 != 0;
            auto emit = [&](float* o, float c0x, float c0y, float c1x, float c1y) {
                if (st256) {
                    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(sc * c0x), "f"(-sg * c0y),
                                 "f"(sc * c1x), "f"(-sg * c1y), "f"(sc * c0y), "f"(sg * c0x), "f"(sc * c1y), "f"(sg * c1x) : "memory");
                    return;
                }
                float4* oh = reinterpret_cast<float4*>(o);
                oh[0] = make_float4(sc * c0x, -sg * c0y, sc * c1x, -sg * c1y);
                oh[1] = make_float4(sc * c0y, sg * c0x, sc * c1y, sg * c1x);
            };
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) {
                const cpx c0a = row0[4 * j1 + par], c0b = row0[4 * j1 + par + 2];
                emit(base + (size_t)(4 * j1) * bin_stride, c0a.x, c0a.y, rA[j1], iA[j1]);
                emit(base + (size_t)(4 * j1 + 2) * bin_stride, c0b.x, c0b.y, rB[j1], iB[j1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM PTX wrappers (sm_100a)
__device__ __forceinline__ void os_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void os_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void os_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void os_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32 (K = 8 per instruction), issued by ONE thread
__device__ __forceinline__ void os_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the descriptors given as (low word, high word): between the instructions of one K-stage only the
// 14-bit start-address field (low word) moves, so the issue loop works on 32-bit values
__device__ __forceinline__ void os_mma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool os_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void os_mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void os_tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void os_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void os_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void os_named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// TMA 1-D bulk copy shared -> global (SASS: UBLKCP), bulk-group completion
__device__ __forceinline__ void os_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
struct alignas(64) OsTensorMap { unsigned long long opaque[16]; };       // CUtensorMap (128 bytes), built on the host
// TMA tensor store shared -> global (SASS: UTMASTG), bulk-group completion
__device__ __forceinline__ void os_tma_store_3d(const void* tmap, const void* ssrc, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(ssrc)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// L2 prefetch of a contiguous run (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void os_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void os_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void os_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t os_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t os_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// os_gemm: persistent, warp-specialised.  Work item = (bin, tile block, template block); a CTA owns a
// contiguous range of items ordered so that consecutive items share the B operand (same bin & tile block).
//   warp 0 : TMA producer (cp.async.bulk + mbarrier expect_tx), A ring of `nsta` K-stages, 2 B buffers
//   warp 1 : TMEM allocation + single-thread tcgen05.mma issue; 3 passes (lo*hi, hi*lo, hi*hi) per K-stage
//   warps 2-5 : epilogue, tcgen05.ld -> registers -> smem staging -> one bulk store per template row
//   warps 6-9 : A splitter.  The A images travel through HBM as plain fp32 (half the bytes of a hi/lo pair);
//               once a K-stage has landed in the deep raw ring these warps rewrite it in place as hi = tf32(a) and
//               write a - hi into a 2-deep lo ring, then hand both to the MMA warp (fence.proxy.async + mbarrier)
struct OsGemmArgs {
    const float* Aimg;
    const float* Bimg;
    float* P;
    int NTBLK, NNB, NKS, KC, NMMA, RS;
    int RSP;          // row pitch of the epilogue staging tile in floats (RS + 4 when RS/4 is even, see os_config_tiles; P itself is dense)
    int use_pmap;     // epilogue: TMA tensor store of the padded staging tile, the pad columns clipped by the tensor bounds
    long long nitems;
    int nsta;
    int lbo_swap;     // debug: swap the LBO / SBO fields of the smem descriptors
    int use_tmap;     // epilogue: one bulk store per item (the whole 128 x RS staging tile) instead of one per template row
    int pf_items;     // L2 prefetch distance of the TMA producer, in work items (0 = off)
    int dbg;          // timing experiments only (FFTCONV_OS_DBG): 1 = no P store
    int hi_inplace;   // 1: the splitter rewrites the landed fp32 K-stage as hi = tf32(a) (debug); 0: the tensor core
                      // itself ignores the 13 low mantissa bits of a kind::tf32 operand, so the raw stage IS the hi operand
};

// Walks the work items (tile block, bin, template block) of one CTA without 64-bit divisions in the loop
// (the single-thread producer / MMA roles execute every instruction at full dependent latency).
struct OsItemIter {
    int tblk, bin, nblk, ntblk, nnb;
    long long key;
    // item = ((bin * NNB + nblk) * NTBLK + tblk): BIN-major.  With several tile blocks (batched images) a CTA stays on one
    // bin while it walks the tile blocks, so the A images of that bin (NTBLK x 32 KB) are re-read from L2 instead of
    // being streamed from HBM once per tile block (tile-block-major order: 4.4 GB of A reads per launch at config 4).
    __device__ __forceinline__ OsItemIter(long long it, int NTBLK, int NNB) : ntblk(NTBLK), nnb(NNB) {
        key = it / NTBLK;
        tblk = (int)(it - key * NTBLK);
        bin = (int)(key / NNB);
        nblk = (int)(key - (long long)bin * NNB);
    }
    __device__ __forceinline__ void next() {
        if (++tblk == ntblk) {
            tblk = 0; ++key;
            if (++nblk == nnb) { nblk = 0; ++bin; }
        }
    }
    __device__ __forceinline__ bool last_of_key() const { return tblk + 1 == ntblk; }
    __device__ __forceinline__ size_t b_block() const { return (size_t)nblk * OS_NBIN + bin; }     // index of the B image block
};

__global__ void __launch_bounds__(320, 1) os_gemm(OsGemmArgs g, const __grid_constant__ OsTensorMap pmap)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    const uint32_t a_half = (uint32_t)g.KC * OS_TM * 16u;       // one K-stage of A (fp32 as it travels; one hi or lo image)
    const uint32_t b_stage = (uint32_t)g.KC * g.NMMA * 16u;     // one K-stage of B (fp32 as it travels = hi operand)
    const uint32_t b_buf = b_stage * g.NKS;
    unsigned char* a_sm = os_smem_raw;                          // raw ring [nsta][a_half]: TMA target = hi operand
    unsigned char* lo_sm = a_sm + (size_t)g.nsta * a_half;      // lo ring [2][a_half]
    unsigned char* b_sm = lo_sm + 2 * (size_t)a_half;           // [2][b_buf] raw B of the current / next key
    unsigned char* blo_sm = b_sm + 2 * (size_t)b_buf;           // [2][b_buf] b - tf32(b)
    float* stage_sm = reinterpret_cast<float*>(blo_sm + 2 * (size_t)b_buf);   // [128][RSP] epilogue staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(stage_sm) + (size_t)OS_TM * g.RSP * 4u);
    uint64_t* a_full = bars;                 // [nsta]  TMA -> splitter
    uint64_t* a_empty = bars + 8;            // [nsta]  MMA -> TMA
    uint64_t* a_ready = bars + 16;           // [nsta]  splitter -> MMA
    uint64_t* lo_empty = bars + 24;          // [2]     MMA -> splitter
    uint64_t* b_full = bars + 26;            // [2]
    uint64_t* b_empty = bars + 28;           // [2]
    uint64_t* acc_full = bars + 30;          // [2]
    uint64_t* acc_empty = bars + 32;         // [2]
    uint64_t* b_ready = bars + 34;           // [2]     splitter -> MMA (lo image of B written)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long lo = g.nitems * (long long)blockIdx.x / gridDim.x;
    const long long hi = g.nitems * (long long)(blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < g.nsta; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], 128); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&lo_empty[i], 1);
            mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128);
            mbar_init(&b_ready[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) os_tmem_alloc(tmem_slot, 2 * OS_ACC_COLS);
    os_tc_fence_before();
    __syncthreads();
    os_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            long long curkey = -1;
            uint32_t nb = 0, na = 0;
            OsItemIter w(lo, g.NTBLK, g.NNB);
            // The A images are streamed exactly once, so every ring slot would otherwise pay the full HBM latency under
            // load (~2.3 us measured: with 4 slots of 16 KB that caps the CTA at ~30 GB/s).  A second iterator runs
            // `pf` items ahead and pulls the A block of that item (and the B block when a new key starts) into L2: the
            // bytes in flight towards HBM sit in L2 instead of shared memory and the ring only sees L2 latency.
            const int pf = g.pf_items;
            OsItemIter wp(lo, g.NTBLK, g.NNB);
            long long pit = lo, pkey = -1;
            auto prefetch_item = [&]() {
                if (pit >= hi) return;
                if (wp.key != pkey) {
                    pkey = wp.key;
                    if (pit != lo) os_prefetch_l2(reinterpret_cast<const unsigned char*>(g.Bimg) + wp.b_block() * b_buf, b_buf);
                }
                os_prefetch_l2(reinterpret_cast<const unsigned char*>(g.Aimg) + ((size_t)wp.tblk * OS_NBIN + wp.bin) * g.NKS * a_half,
                               a_half * g.NKS);
                ++pit; wp.next();
            };
            for (int i = 0; i < pf; ++i) prefetch_item();
            for (long long it = lo; it < hi; ++it, w.next()) {
                const long long key = w.key;
                const int tblk = w.tblk, bin = w.bin;
                if (pf) prefetch_item();
                if (key != curkey) {
                    const uint32_t bb = nb & 1;
                    if (nb >= 2) mbar_wait(&b_empty[bb], ((nb >> 1) - 1) & 1);
                    mbar_expect_tx(&b_full[bb], b_buf);
                    bulk_g2s(b_sm + (size_t)bb * b_buf, reinterpret_cast<const unsigned char*>(g.Bimg) + w.b_block() * b_buf,
                             b_buf, &b_full[bb]);
                    ++nb;
                    curkey = key;
                }
                const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.Aimg) +
                                            ((size_t)tblk * OS_NBIN + bin) * g.NKS * a_half;
                for (int ks = 0; ks < g.NKS; ++ks) {
                    const uint32_t st = na % g.nsta, fill = na / g.nsta;
                    if (fill >= 1) mbar_wait(&a_empty[st], (fill - 1) & 1);
                    mbar_expect_tx(&a_full[st], a_half);
                    bulk_g2s(a_sm + (size_t)st * a_half, asrc + (size_t)ks * a_half, a_half, &a_full[st]);
                    ++na;
                }
            }
        }
    } else if (warp == 1) {
        // MMA issue.  The WHOLE warp walks the item list with warp-uniform values and one elected lane issues the
        // instructions: with a single active thread (`if (lane == 0)`) the compiler cannot prove that the operands of the
        // uniform-datapath instructions (UTCHMMA / UTCBAR take uniform registers) are uniform and wraps every one of them
        // in a vote / elect / broadcast loop -- 19 dependent instructions per MMA, 4 600 cycles per item, which made this
        // single thread the bottleneck of the kernel (ncu r01i: the role never waits on a barrier).
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = os_idesc_tf32(OS_TM, g.NMMA);
        const uint32_t a_lbo = OS_TM * 16u, b_lbo = (uint32_t)g.NMMA * 16u, sbo = 128u;
        // descriptor words: only the 14-bit start-address field of the low word changes between instructions
        const uint64_t adesc0 = g.lbo_swap ? os_smem_desc(0, sbo, a_lbo) : os_smem_desc(0, a_lbo, sbo);
        const uint64_t bdesc0 = g.lbo_swap ? os_smem_desc(0, sbo, b_lbo) : os_smem_desc(0, b_lbo, sbo);
        const uint32_t a_w0 = (uint32_t)adesc0, a_w1 = (uint32_t)(adesc0 >> 32);
        const uint32_t b_w0 = (uint32_t)bdesc0, b_w1 = (uint32_t)(bdesc0 >> 32);
        const uint32_t a_step = 2u * a_lbo >> 4, b_step = 2u * b_lbo >> 4;     // K advances by 2 units of 16 bytes
        const uint32_t a_sm0 = smem_u32(a_sm), lo_sm0 = smem_u32(lo_sm), b_sm0 = smem_u32(b_sm), blo_sm0 = smem_u32(blo_sm);
        const int nj = g.KC >> 1;                                              // <= 4
        long long curkey = -1;
        uint32_t nb = 0, na = 0, nit = 0, bcur = 0;
        OsItemIter w(lo, g.NTBLK, g.NNB);
        for (long long it = lo; it < hi; ++it, ++nit, w.next()) {
            const long long key = w.key;
            if (key != curkey) {
                bcur = nb & 1;
                mbar_wait(&b_ready[bcur], (nb >> 1) & 1);
                ++nb;
                curkey = key;
            }
            const uint32_t acc = nit & 1;
            if (nit >= 2) mbar_wait(&acc_empty[acc], ((nit >> 1) - 1) & 1);
            const uint32_t tmem_d = tb + acc * OS_ACC_COLS;
            for (int ks = 0; ks < g.NKS; ++ks) {
                const uint32_t st = na % g.nsta, ls = na & 1;
                mbar_wait(&a_ready[st], (na / g.nsta) & 1);
                os_tc_fence_after();
                const uint32_t ad_hi = a_w0 | (((a_sm0 + st * a_half) >> 4) & 0x3FFFu);
                const uint32_t ad_lo = a_w0 | (((lo_sm0 + ls * a_half) >> 4) & 0x3FFFu);
                const uint32_t b_off = bcur * b_buf + (uint32_t)ks * b_stage;
                const uint32_t bd_hi = b_w0 | (((b_sm0 + b_off) >> 4) & 0x3FFFu);
                const uint32_t bd_lo = b_w0 | (((blo_sm0 + b_off) >> 4) & 0x3FFFu);
                if (os_elect_one()) {
                    // small terms first: (A lo, B hi), (A hi, B lo), then (A hi, B hi)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nj && !(g.dbg & 8)) os_mma_tf32_w(tmem_d, ad_lo + j * a_step, a_w1, bd_hi + j * b_step, b_w1, idesc, (ks | j) != 0 ? 1u : 0u);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nj && !(g.dbg & 8)) os_mma_tf32_w(tmem_d, ad_hi + j * a_step, a_w1, bd_lo + j * b_step, b_w1, idesc, 1u);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nj) os_mma_tf32_w(tmem_d, ad_hi + j * a_step, a_w1, bd_hi + j * b_step, b_w1, idesc, 1u);
                    os_mma_commit(&a_empty[st]);
                    os_mma_commit(&lo_empty[ls]);
                    if (ks + 1 == g.NKS) {
                        os_mma_commit(&acc_full[acc]);
                        if (it + 1 == hi || w.last_of_key()) os_mma_commit(&b_empty[bcur]);
                    }
                }
                __syncwarp();
                ++na;
            }
        }
    } else if (warp >= 6) {
        // A splitter: lo = a - tf32(a) into the lo ring (the raw stage itself serves as the hi operand: kind::tf32 reads
        // only the sign, the exponent and the 10 high mantissa bits of each 32-bit element)
        // The B image of a key gets the same treatment once per key (it serves all template blocks of the bin).
        const int sp = threadIdx.x - 192;             // 0..127
        const int kc = g.KC;                          // float4 per thread and stage (a_half / 16 / 128), <= 8
        const uint32_t nbv = b_buf / 16u;
        uint32_t na = 0, nb = 0;
        long long curkey = -1;
        OsItemIter w(lo, g.NTBLK, g.NNB);
        for (long long it = lo; it < hi; ++it, w.next()) {
            if (w.key != curkey) {
                curkey = w.key;
                const uint32_t bb = nb & 1;
                mbar_wait(&b_full[bb], (nb >> 1) & 1);
                const float4* bp = reinterpret_cast<const float4*>(b_sm + (size_t)bb * b_buf);
                float4* blp = reinterpret_cast<float4*>(blo_sm + (size_t)bb * b_buf);
                for (uint32_t i = sp; i < nbv; i += 128) {
                    const float4 v = bp[i];
                    blp[i] = make_float4(v.x - os_tf32_hi(v.x), v.y - os_tf32_hi(v.y), v.z - os_tf32_hi(v.z), v.w - os_tf32_hi(v.w));
                }
                fence_proxy_async();
                os_mbar_arrive(&b_ready[bb]);
                ++nb;
            }
            for (int ks = 0; ks < g.NKS; ++ks, ++na) {
                const uint32_t st = na % g.nsta, ls = na & 1;
                mbar_wait(&a_full[st], (na / g.nsta) & 1);
                float4* hp = reinterpret_cast<float4*>(a_sm + (size_t)st * a_half) + sp;
                float4* lp = reinterpret_cast<float4*>(lo_sm + (size_t)ls * a_half) + sp;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < kc) v[i] = hp[128 * i];
                if (na >= 2) mbar_wait(&lo_empty[ls], ((na >> 1) - 1) & 1);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < kc && !(g.dbg & 4)) {
                        const float4 h = make_float4(os_tf32_hi(v[i].x), os_tf32_hi(v[i].y), os_tf32_hi(v[i].z), os_tf32_hi(v[i].w));
                        if (g.hi_inplace) hp[128 * i] = h;
                        lp[128 * i] = make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
                    }
                fence_proxy_async();                  // generic-proxy writes -> visible to tcgen05.mma (async proxy)
                os_mbar_arrive(&a_ready[st]);
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                // template row of the block
        const bool leader = threadIdx.x == 64;        // issues the tensor store of the staging tile
        uint32_t nit = 0;
        OsItemIter w(lo, g.NTBLK, g.NNB);
        for (long long it = lo; it < hi; ++it, ++nit, w.next()) {
            const int tblk = w.tblk, bin = w.bin, nblk = w.nblk;
            const uint32_t acc = nit & 1;
            mbar_wait(&acc_full[acc], (nit >> 1) & 1);
            os_tc_fence_after();
            const uint32_t taddr = tmem_base + acc * OS_ACC_COLS + ((uint32_t)(q * 32) << 16);
            // row pitch RSP with RSP/4 odd: the 8 lanes of a 128-bit store wavefront (rows RSP words apart) hit 8 different
            // bank groups; with the dense pitch RS = 80 they hit 2 (ncu r02b: 48 % of the kernel's shared-memory wavefronts were
            // bank conflicts, and the kernel is bound by shared-memory bandwidth: TMA in, split, 3 operand reads per MMA,
            // staging, store)
            float* srow = stage_sm + (size_t)row * g.RSP;
            uint32_t r[80];                           // all TMEM loads of the row in flight, one wait (RS <= 80 columns)
#pragma unroll
            for (int i = 0; i < 10; ++i)
                if (8 * i < g.RS) os_tmem_ld8(taddr + 8 * i, r + 8 * i);
            os_tmem_ld_wait();
            os_tc_fence_before();
            os_mbar_arrive(&acc_empty[acc]);          // the accumulator is free again before the store is even staged
            if (g.use_tmap) {
                if (leader) os_bulk_wait_read0();     // the previous tensor store has finished reading the staging tile
                os_named_bar_sync(1, 128);
            } else {
                os_bulk_wait_read0();                 // this thread's previous bulk store has finished reading its staging row
            }
#pragma unroll
            for (int i = 0; i < 10; ++i)
                if (8 * i < g.RS) {
                    float4* o = reinterpret_cast<float4*>(srow + 8 * i);
                    o[0] = make_float4(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1]), __uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3]));
                    o[1] = make_float4(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5]), __uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7]));
                }
            fence_proxy_async();                      // this thread's staging writes -> visible to the bulk-copy engine
            // P[tblk][nblk][bin][template][RS]: the item's 128 x RS block is ONE contiguous run of global memory
            const size_t item_index = ((size_t)tblk * g.NNB + nblk) * OS_NBIN + bin;
            float* dst = g.P + item_index * ((size_t)OS_TM * g.RS);
            if (g.use_tmap) {
                os_named_bar_sync(1, 128);
                if (leader && !(g.dbg & 1)) {
                    // padded staging tile (pitch RSP): P seen as {RS, 128, items}, box {RSP, 128, 1} -- the RSP - RS pad
                    // columns lie outside the tensor and are clipped by the TMA engine, P stays dense
                    if (g.use_pmap) os_tma_store_3d(&pmap, stage_sm, 0, 0, (int)item_index);
                    else os_bulk_s2g(dst, stage_sm, (uint32_t)(OS_TM * g.RS * 4));
                }
            } else {
                os_bulk_s2g(dst + (size_t)row * g.RS, srow, (uint32_t)g.RS * 4u);     // one bulk copy per template row
            }
        }
        os_bulk_wait0();
    }
    os_tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        os_tc_fence_after();
        os_tmem_dealloc(tmem_base, 2 * OS_ACC_COLS);
    }
}

// Reference implementation of os_gemm on the SIMT pipe (exact fp32 FMA over hi+lo).  Debug/validation
// only (FFTCONV_OS_GEMM=simt); never selected by the product path.
__global__ void __launch_bounds__(128) os_gemm_simt(OsGemmArgs g)
{
    const long long it = blockIdx.x;
    const long long key = it / g.NTBLK;
    const int tblk = (int)(it - key * g.NTBLK);
    const int bin = (int)(key / g.NNB);
    const int nblk = (int)(key % g.NNB);
    const size_t a_stage = (size_t)g.KC * OS_TM * 4, b_stage = (size_t)g.KC * g.NMMA * 4;   // floats
    const float* A = g.Aimg + ((size_t)tblk * OS_NBIN + bin) * g.NKS * a_stage;
    const float* B = g.Bimg + ((size_t)nblk * OS_NBIN + bin) * g.NKS * b_stage;
    const int t = threadIdx.x;
    float* P = g.P + ((((size_t)tblk * g.NNB + nblk) * OS_NBIN + bin) * OS_TM + t) * (size_t)g.RS;
    for (int n = 0; n < g.RS; ++n) {
        float acc = 0.f;
        for (int ks = 0; ks < g.NKS; ++ks)
            for (int kc = 0; kc < g.KC; ++kc) {
                const float4 ah = *reinterpret_cast<const float4*>(A + ks * a_stage + ((size_t)kc * OS_TM + t) * 4);
                const float4 al = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 bh = *reinterpret_cast<const float4*>(B + ks * b_stage + ((size_t)kc * g.NMMA + n) * 4);
                const float4 bl = make_float4(0.f, 0.f, 0.f, 0.f);
                acc = fmaf(ah.x + al.x, bh.x + bl.x, acc);
                acc = fmaf(ah.y + al.y, bh.y + bl.y, acc);
                acc = fmaf(ah.z + al.z, bh.z + bl.z, acc);
                acc = fmaf(ah.w + al.w, bh.w + bl.w, acc);
            }
        P[n] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// os_inverse: CTA = (template, group of OS_IG tiles), 64 threads per tile, two threads per 64-point line.
//   gather   the 33 x 64 product spectrum of each tile from P into shared memory (cp.async), column (v) major;
//            columns 0 and 32 (spectra of real sequences along h) are then combined into one complex column
//   pass 1   inverse along h: line v (0..31) = column v for rows 0..32 and the conjugate of column 64-v above
//            (Hermitian symmetry of the 2-D spectrum of a real tile); result in place
//   pass 2   C2R along w: line y packs rows y and y+32 into one complex sequence; lanes run along h, so every
//            store instruction writes one contiguous run of a plane column -- the valid (65-maxkh) x (65-maxkw)
//            block goes straight from registers to the output plane (crop fused), no staging
// The 1/4096 of the inverse transform is folded into the B operand images (os_data_fft).
// grid = (ceil(NT/OS_IG), templates in chunk); smem = OS_IG * OS_ITILE * 8 B.
struct OsInvArgs {
    const OsLevel* levels;  // pyramid batch (os_inverse_z only): per-level geometry; plane of (level l, template t) =
    int nlevels;            // outs[l * out_img_stride + t], stored with the level's own crop / out_ld
    const float* P;
    float* const* outs;
    int nk, NNB, NTn, RS, NT, NTimg, nth, Sh, Sw, oy0, ox0;
    int FH, FW, crop_h, crop_w, out_ld;
    int out_img_stride;     // plane of (image n, template t) = outs[n * out_img_stride + t]
    // fused reduction (fftconv_bank_conv_max): when peak_keys != nullptr no plane is written; every template keeps the
    // maximum of its full linear convolution (H + kh - 1) x (W + kw - 1) as a packed (ordered value, position) key
    unsigned long long* peak_keys;
    const int2* khw;        // (kh, kw) per template of the chunk (peak, detection and correlation modes)
    // fused detection (fftconv_bank_conv_detect / _topk): no plane is written either.  Every response r = conv + bias[t]
    // of the template's full linear convolution is compared with thr[t]:
    //   det_mode 1   r >= thr[t]  ->  key appended to det_keys[t][slot], slot = atomicAdd(det_count[t]) (kept while < det_cap)
    //   det_mode 2   candidate pass of the top-k selection: every lane writes the key of its own best response to
    //                det_keys[t][tile * 32 + lane] (0 = none); the k-th largest candidate is a lower bound of the k-th
    //                largest response, so a det_mode 1 pass with that bound as thr[t] captures the exact top-k
    int det_mode, det_cap;
    unsigned long long* det_keys;
    unsigned int* det_count;
    const float* det_bias;  // per template of the chunk, or nullptr
    const float* det_thr;   // per template of the chunk (det_mode 1)
    int H, W;
    int dbg;                // timing experiments only (FFTCONV_OS_DBG): 16 = contiguous load instead of the gather, 32 = no plane store
    int corr;               // correlation mode: plane position (Y, X) of the flipped-template convolution is stored at
                            // ((Y - kh + 1) mod FH, (X - kw + 1) mod FW)
};

// order-preserving map float -> uint32 (larger float <=> larger uint)
__device__ __forceinline__ unsigned os_ordered(float f) {
    const unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
// key = ordered value << 32 | ~(x << 16 | y): ties resolve to the smallest position
__device__ __forceinline__ unsigned long long os_peak_key(float v, int y, int x) {
    return ((unsigned long long)os_ordered(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (((unsigned)x << 16) | (unsigned)y));
}

struct fftconv_peak_dev { float value; int y, x, pad; };
__global__ void os_peak_finalize(const unsigned long long* keys, int K, fftconv_peak_dev* out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const unsigned long long key = keys[k];
    const unsigned o = (unsigned)(key >> 32), pos = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu);
    const unsigned u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    out[k].value = __uint_as_float(u); out[k].y = (int)(pos & 0xFFFFu); out[k].x = (int)(pos >> 16); out[k].pad = 0;
}

// Tail of both inverse kernels: the 64 outputs a lane holds after pass 2 (rows ylo / yhi = lane - oy0 (+32), columns
// 4*j1 + par (+2) - ox0 of the valid block of its tile) go to the plane (crop fused), to the plane shifted by the template
// extent (correlation mode), or into the fused reductions (maximum, detections) -- in which case no plane is written.
struct OsTileOut { float* dst; int ny, nx, y0, x0, ld; };
__device__ __forceinline__ void os_inverse_emit(const OsInvArgs& a, int t, int lane, int par, int tile_index, const OsTileOut& T,
                                                const float (&reA)[16], const float (&imA)[16], const float (&reB)[16], const float (&imB)[16])
{
    const int ny = T.ny, nx = T.nx;
    const int y = lane;
    const int ylo = y - a.oy0, yhi = y + 32 - a.oy0;                   // rows of the valid block held by this lane
    const bool wlo = ylo >= 0 && ylo < ny, whi = yhi < ny;
    if (a.peak_keys || a.det_mode == 2) {
        const float bias = (a.det_mode && a.det_bias) ? a.det_bias[t] : 0.f;
        float best = -INFINITY; int by = 0, bx = 0;
        auto upd = [&](float v, int yy, int xx) { if (v > best) { best = v; by = yy; bx = xx; } };
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) {                             // ascending x: the first maximum wins
            const int xa = 4 * j1 + par - a.ox0, xb = xa + 2;
            if (xa >= 0 && xa < nx) { if (wlo) upd(reA[j1], ylo, xa); if (whi) upd(imA[j1], yhi, xa); }
            if (xb >= 0 && xb < nx) { if (wlo) upd(reB[j1], ylo, xb); if (whi) upd(imB[j1], yhi, xb); }
        }
        unsigned long long key = best > -INFINITY ? os_peak_key(best + bias, T.y0 + by, T.x0 + bx) : 0ull;
        if (a.det_mode == 2) {                                        // one candidate per (tile, parity, lane)
            a.det_keys[(size_t)t * a.det_cap + (size_t)(tile_index * 2 + par) * 32 + lane] = key;
            return;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0 && key) atomicMax(a.peak_keys + t, key);
        return;
    }
    if (a.det_mode == 1) {
        const float bias = a.det_bias ? a.det_bias[t] : 0.f, thr = a.det_thr[t];
        unsigned long long* keys = a.det_keys + (size_t)t * a.det_cap;
        auto hit = [&](float v, int yy, int xx) {
            const float r = v + bias;
            if (r >= thr) {
                const unsigned slot = atomicAdd(a.det_count + t, 1u);
                if (slot < (unsigned)a.det_cap) keys[slot] = os_peak_key(r, T.y0 + yy, T.x0 + xx);
            }
        };
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) {
            const int xa = 4 * j1 + par - a.ox0, xb = xa + 2;
            if (xa >= 0 && xa < nx) { if (wlo) hit(reA[j1], ylo, xa); if (whi) hit(imA[j1], yhi, xa); }
            if (xb >= 0 && xb < nx) { if (wlo) hit(reB[j1], ylo, xb); if (whi) hit(imB[j1], yhi, xb); }
        }
        return;
    }
    if (a.corr) {
        const int2 k = a.khw[t];
        float* base = T.dst;
        int ydl = T.y0 + ylo - (k.x - 1), ydh = T.y0 + yhi - (k.x - 1);
        if (ydl < 0) ydl += a.FH;
        if (ydh < 0) ydh += a.FH;
        const bool slo = wlo && ydl < a.crop_h, shi = whi && ydh < a.crop_h;
        const int xs = T.x0 - (k.y - 1);
        auto put = [&](int xr, float vlo, float vhi) {
            if (xr < 0 || xr >= nx) return;
            int xd = xs + xr;
            if (xd < 0) xd += a.FW;
            if (xd >= a.crop_w) return;
            float* d = base + (size_t)xd * a.out_ld;
            if (slo) d[ydl] = vlo;
            if (shi) d[ydh] = vhi;
        };
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) {
            put(4 * j1 + par - a.ox0, reA[j1], imA[j1]);
            put(4 * j1 + par + 2 - a.ox0, reB[j1], imB[j1]);
        }
        return;
    }
    if (a.dbg & 32) { if (reA[3] == 1.2345f) T.dst[0] = imB[5]; return; }
    // every store carries its own predicate and a 32-bit element offset (the branchy form with 64-bit products cost
    // 9 instructions per store: a quarter of all instructions of the inverse)
    float* dlo = T.dst + ylo;
    const int ld = T.ld;
    const unsigned unx = (unsigned)nx;
    int off = (par - a.ox0) * ld;
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) {
        const int xa = 4 * j1 + par - a.ox0;
        const bool pa = (unsigned)xa < unx, pb = (unsigned)(xa + 2) < unx;
        float* d = dlo + off;
        float* e = d + 2 * ld;
        if (pa && wlo) d[0] = reA[j1];
        if (pa && whi) d[32] = imA[j1];
        if (pb && wlo) e[0] = reB[j1];
        if (pb && whi) e[32] = imB[j1];
        off += 4 * ld;
    }
}

// Detection finalisation, one CTA per template: the `nsel` largest of the template's candidate keys, in descending order
// (repeated block-wide maximum below the previous pick; keys are unique because they carry the position).
//   n = min(count[t], cap) when count != nullptr, else cap (fixed candidate slots, empty = 0)
//   out_peaks != nullptr : peaks[t][i] = pick i (value -inf, y = x = -1 when fewer than nsel candidates exist)
//   out_thr   != nullptr : thr[t] = value of pick nsel-1, or -inf when fewer exist (lower bound for the top-k second pass)
//   out_counts!= nullptr : counts[t] = count[t] (every response found, including those beyond cap)
__global__ void __launch_bounds__(256) os_det_select(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ count,
                                                     int cap, int nsel, fftconv_peak_dev* out_peaks, float* out_thr, int* out_counts)
{
    __shared__ unsigned long long red[8];
    __shared__ unsigned long long last_s;
    const int t = blockIdx.x;
    const unsigned long long* k = keys + (size_t)t * cap;
    const int n = count ? (int)min(count[t], (unsigned)cap) : cap;
    if (out_counts && threadIdx.x == 0) out_counts[t] = count ? (int)min(count[t], 0x7fffffffu) : 0;
    unsigned long long last = ~0ull;
    for (int i = 0; i < nsel; ++i) {
        unsigned long long best = 0;
        for (int j = threadIdx.x; j < n; j += 256) {
            const unsigned long long v = k[j];
            if (v < last && v > best) best = v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) best = red[w] > best ? red[w] : best;
            last_s = best;
            if (out_peaks) {
                fftconv_peak_dev pk;
                if (best) {
                    const unsigned o = (unsigned)(best >> 32), pos = 0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFu);
                    const unsigned u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
                    pk.value = __uint_as_float(u); pk.y = (int)(pos & 0xFFFFu); pk.x = (int)(pos >> 16); pk.pad = 0;
                } else { pk.value = -INFINITY; pk.y = -1; pk.x = -1; pk.pad = 0; }
                out_peaks[(size_t)t * nsel + i] = pk;
            }
            if (out_thr && i == nsel - 1) {
                float thr = -INFINITY;
                if (best) {
                    const unsigned o = (unsigned)(best >> 32);
                    thr = __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
                }
                out_thr[t] = thr;
            }
        }
        __syncthreads();
        last = last_s;
        if (last == 0) last = 0;           // nothing left: the remaining picks stay empty (best stays 0)
    }
}

__global__ void os_fill_f32(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__device__ __forceinline__ int os_icol(int v) { return v * OS_ICOL + ((v >> 5) << 2); }

__global__ void __launch_bounds__(OS_IG * 64, 12 / OS_IG) os_inverse(OsInvArgs a)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(os_smem_raw);                        // [4][OS_ITILE]
    const int t = blockIdx.y;
    const int m0 = blockIdx.x * OS_IG;
    const int tblk = t / OS_TM, tl = t - tblk * OS_TM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-tile store parameters, fetched / computed now by one thread per tile so that the plane-pointer load hides
    // behind the gather: base of the valid block in the plane, rows / columns to store
    __shared__ float* tile_dst[OS_IG];
    __shared__ int tile_ny[OS_IG], tile_nx[OS_IG], tile_y0[OS_IG], tile_x0[OS_IG];
    if (threadIdx.x < OS_IG) {
        const int m = m0 + threadIdx.x;
        float* d = nullptr; int ny = 0, nx = 0;
        if (m < a.NT) {
            const int img = m / a.NTimg, mt = m - img * a.NTimg;
            const int tj = mt / a.nth, ti = mt - tj * a.nth;
            const int Y0 = ti * a.Sh, X0 = tj * a.Sw;
            tile_y0[threadIdx.x] = Y0; tile_x0[threadIdx.x] = X0;
            if (a.peak_keys || a.det_mode) {      // region of the full linear convolution of THIS template
                const int2 k = a.khw[t];
                ny = min(a.Sh, a.H + k.x - 1 - Y0); nx = min(a.Sw, a.W + k.y - 1 - X0);
            } else if (a.corr) {
                ny = min(a.Sh, a.FH - Y0); nx = min(a.Sw, a.FW - X0);
                d = a.outs[(size_t)img * a.out_img_stride + t];
            } else {
                ny = min(a.Sh, a.crop_h - Y0); nx = min(a.Sw, a.crop_w - X0);
                d = a.outs[(size_t)img * a.out_img_stride + t] + (size_t)X0 * a.out_ld + Y0;
            }
        }
        tile_dst[threadIdx.x] = d; tile_ny[threadIdx.x] = ny; tile_nx[threadIdx.x] = nx;
    }

    // ---- gather: thread = (tile gq, column v); the 4 tiles of a group share one 32-byte sector of P.
    //      33 asynchronous 8-byte copies per thread (cp.async, no register staging): all of them in flight at once
    {
        const int gq = threadIdx.x & (OS_IG - 1);
        const int v = threadIdx.x >> OS_IGB;
        const int m = m0 + gq;
        cpx* dst = buf + gq * OS_ITILE + os_icol(v);
        if (m < a.NT) {
            const int nblk = m / a.NTn, ml = m - nblk * a.NTn;
            const size_t ustride = (size_t)64 * OS_TM * (a.RS / 2);        // cpx units; P[tblk][nblk][u*64 + v][template][RS]
            const cpx* pp = reinterpret_cast<const cpx*>(
                a.P + ((((size_t)tblk * a.NNB + nblk) * OS_NBIN + v) * OS_TM + tl) * (size_t)a.RS + 2 * ml);
            const uint32_t d0 = smem_u32(dst);
#pragma unroll
            for (int u = 0; u <= 32; ++u) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + 8u * u), "l"(pp) : "memory");
                pp += ustride;
            }
        } else {
#pragma unroll
            for (int u = 0; u <= 32; ++u) dst[u] = make_float2(0.f, 0.f);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    // columns 0 and 32 are spectra of real sequences along h: combine them into one complex column pair
    //     col0[u] := Z[u][0] + i Z[u][32],   col32[u] := Z[u][0] - i Z[u][32]
    if (threadIdx.x < OS_IG * 33) {
        const int gq = threadIdx.x / 33, u = threadIdx.x - gq * 33;
        cpx* c0 = buf + gq * OS_ITILE + os_icol(0) + u;
        cpx* c32 = buf + gq * OS_ITILE + os_icol(32) + u;
        const cpx za = *c0, zb = *c32;
        *c0 = make_float2(za.x - zb.y, za.y + zb.x);
        *c32 = make_float2(za.x + zb.y, za.y - zb.x);
    }
    __syncthreads();
    const int par = warp >> OS_IGB;                                        // warp-uniform: tasks (par, par + 2)
    const int gq = warp & (OS_IG - 1);                                     // one warp = the 32 lines of one tile
    cpx* tile = buf + gq * OS_ITILE;
    // ---- pass 1: inverse along h, in place.  x[u] = col_v[u] (u <= 32), conj(col_mv[64-u]) (u > 32)
    {
        const int v = lane, mv = lane ? 64 - lane : 32;
        cpx* cv = tile + os_icol(v);
        cpx* cm = tile + os_icol(mv);
        float reA[16], imA[16], reB[16], imB[16];
        auto ld1 = [&](int u) -> cpx {
            if (u <= 32) return cv[u];
            const cpx q = cm[64 - u];
            return make_float2(q.x, -q.y);
        };
        auto ld = [&](int j) { const cpx z0 = ld1(j), z1 = ld1(j + 1); return make_float4(z0.x, z0.y, z1.x, z1.y); };
        if (par == 0) os_fft64_pair<0, true>(ld, reA, imA, reB, imB);
        else          os_fft64_pair<1, true>(ld, reA, imA, reB, imB);
        __syncthreads();
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) {                                  // Y[y]: y < 32 -> col v, y >= 32 -> col mv
            const int ya = 4 * j1 + par, yb = ya + 2;                      // (compile-time split: j1 < 8 <=> y < 32)
            if (j1 < 8) { cv[ya] = make_float2(reA[j1], imA[j1]); cv[yb] = make_float2(reB[j1], imB[j1]); }
            else        { cm[ya - 32] = make_float2(reA[j1], imA[j1]); cm[yb - 32] = make_float2(reB[j1], imB[j1]); }
        }
    }
    __syncthreads();
    // ---- pass 2: C2R along w.  W[v] = Y[y][v] + i Y[y+32][v]; outputs straight to the plane
    {
        const int y = lane;
        const int m = m0 + gq;
        if (m >= a.NT) return;                                             // (no barrier below)
        const cpx* ty = tile + y;
        float reA[16], imA[16], reB[16], imB[16];
        auto ld1 = [&](int v) -> cpx {
            if (v == 0) { const cpx p0 = ty[os_icol(0)], p32 = ty[os_icol(32)]; return make_float2(p0.x, p32.x); }
            if (v == 32) { const cpx p0 = ty[os_icol(0)], p32 = ty[os_icol(32)]; return make_float2(p0.y, p32.y); }
            if (v < 32) { const cpx p = ty[os_icol(v)], q = ty[os_icol(64 - v)]; return make_float2(p.x - q.y, p.y + q.x); }
            const cpx p = ty[os_icol(64 - v)], q = ty[os_icol(v)];
            return make_float2(p.x + q.y, q.x - p.y);
        };
        auto ld = [&](int j) { const cpx z0 = ld1(j), z1 = ld1(j + 1); return make_float4(z0.x, z0.y, z1.x, z1.y); };
        if (par == 0) os_fft64_pair<0, true>(ld, reA, imA, reB, imB);
        else          os_fft64_pair<1, true>(ld, reA, imA, reB, imB);
        const OsTileOut T{tile_dst[gq], tile_ny[gq], tile_nx[gq], tile_y0[gq], tile_x0[gq], a.out_ld};
        os_inverse_emit(a, t, lane, par, m0 + gq, T, reA, imA, reB, imB);
    }
}

// ------------------------------------------------------------------------------------------------
// os_inverse_tma: same transform as os_inverse; the gather is done by the TMA engine.
// P keeps the bin-major layout os_gemm streams out (P[tblk][nblk][u][template][v][RS]); seen as a 3-D tensor
// {RS floats, 64 columns v, rows (tblk, nblk, u, template)} the 4 tiles of a group are a box {8, 64, 1}: 64 pieces of 32
// bytes, 320 bytes apart.  One thread issues 33 tensor copies (cp.async.bulk.tensor.3d, one mbarrier) instead of the CTA
// issuing 8 448 eight-byte cp.async — that issue was os_inverse's top stall (mio_throttle) and on its own ran at 2.1 TB/s.
// The box lands dense, [u][v][tile] (2 KB per row, 128-byte aligned as the tensor copy requires):
//   pass 1   lane = (column within a block of 8, tile): every access of a warp is 256 contiguous bytes.  All inputs are in
//            registers when the block barrier falls, so the results are written in os_inverse's per-tile layout
//            (column stride 33) over the same buffer;
//   pass 2   and the store are os_inverse's: lanes along h, one warp per tile, every store instruction writes one
//            contiguous run of a plane column.
constexpr int OS_TROW = 256;          // complex pitch of a landed row: 64 columns x 4 tiles
// per-tile stride of the re-laid buffer: 64 columns x 33 + 4 (columns >= 32 sit 4 elements further).  2116 = 4 mod 16, so
// the 16 lanes of a half warp (4 tiles x 4 columns) of a pass-1 store hit 16 different 8-byte banks (the 2120 of
// os_inverse put tiles q and q + 2 on the same banks: 4 wavefronts per store instead of 2)
constexpr int OS_ITILE2 = 64 * OS_ICOL + 4;
constexpr int OS_ITMA_SMEM = (OS_IG * OS_ITILE2 > OS_CH * OS_TROW ? OS_IG * OS_ITILE2 : OS_CH * OS_TROW) * 8;

__device__ __forceinline__ void os_tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

template <int P, class Slot, bool DFT = true>
__device__ __forceinline__ void os_c2r64_pair(Slot&& S, float* reA, float* imA, float* reB, float* imB) {
    auto bfly = [&](auto cc, const cpx x0, const cpx x1, const cpx x2, const cpx x3) {
        constexpr int c = decltype(cc)::value;
        if (P == 0) {
            const float ar = x0.x + x2.x, ai = x0.y + x2.y, br = x1.x + x3.x, bi = x1.y + x3.y;
            reA[c] = ar + br; imA[c] = ai + bi;
            os_twiddle<2 * c, true>(ar - br, ai - bi, reB[c], imB[c]);
        } else {
            const float ar = x0.x - x2.x, ai = x0.y - x2.y, br = x1.x - x3.x, bi = x1.y - x3.y;
            os_twiddle<c, true>(ar - bi, ai + br, reA[c], imA[c]);          // a + i b
            os_twiddle<3 * c, true>(ar + bi, ai - br, reB[c], imB[c]);      // a - i b
        }
    };
    auto lo = [](const cpx p, const cpx z) { return make_float2(p.x - z.y, p.y + z.x); };    // W[v]
    auto hi = [](const cpx p, const cpx z) { return make_float2(p.x + z.y, z.x - p.y); };    // W[64 - v]
    {
        const cpx p0 = S(0), p32 = S(32), p16 = S(16), p48 = S(48);
        bfly(std::integral_constant<int, 0>{}, make_float2(p0.x, p32.x), lo(p16, p48), make_float2(p0.y, p32.y), hi(p16, p48));
        const cpx p8 = S(8), p56 = S(56), p24 = S(24), p40 = S(40);
        bfly(std::integral_constant<int, 8>{}, lo(p8, p56), lo(p24, p40), hi(p24, p40), hi(p8, p56));
    }
    os_static_for<1, 8>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const cpx pa = S(c), za = S(64 - c), pb = S(c + 16), zb = S(48 - c);
        const cpx pc = S(32 - c), zc = S(c + 32), pd = S(16 - c), zd = S(c + 48);
        bfly(cc, lo(pa, za), lo(pb, zb), hi(pc, zc), hi(pd, zd));
        bfly(std::integral_constant<int, 16 - c>{}, lo(pd, zd), lo(pc, zc), hi(pb, zb), hi(pa, za));
    });
    if (DFT) {
        dft_regs<16, true>(reA, imA);
        dft_regs<16, true>(reB, imB);
    }
}
// One-shot form (one CTA per item, 33 boxes on one mbarrier): FFTCONV_OS_INV_Z=0.  VAR & 1: one warp polls the mbarrier
// (measured slower: 0.267 vs 0.256 ms), VAR & 2: conflict-free tile stride, VAR & 4: paired pass-2 loads.
template <int VAR>
__global__ void __launch_bounds__(256, 3) os_inverse_tma(OsInvArgs a, const __grid_constant__ OsTensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(os_smem_raw);
    __shared__ __align__(8) uint64_t bar;
    __shared__ float* tile_dst[OS_IG];
    __shared__ int tile_ny[OS_IG], tile_nx[OS_IG], tile_y0[OS_IG], tile_x0[OS_IG], tile_ok[OS_IG];
    const int t = blockIdx.y;
    const int NG = a.RS >> 3;
    const int nblk = blockIdx.x / NG, g = blockIdx.x - nblk * NG;
    const int mbase = nblk * a.NTn + 4 * g;                                // first tile of the group
    const int tblk = t / OS_TM, tl = t - tblk * OS_TM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < OS_IG) {
        const int m = mbase + threadIdx.x;
        const bool ok = 4 * g + (int)threadIdx.x < a.NTn && m < a.NT;
        float* d = nullptr; int ny = 0, nx = 0;
        if (ok) {
            const int img = m / a.NTimg, mt = m - img * a.NTimg;
            const int tj = mt / a.nth, ti = mt - tj * a.nth;
            const int Y0 = ti * a.Sh, X0 = tj * a.Sw;
            tile_y0[threadIdx.x] = Y0; tile_x0[threadIdx.x] = X0;
            if (a.peak_keys || a.det_mode) {      // region of the full linear convolution of THIS template
                const int2 k = a.khw[t];
                ny = min(a.Sh, a.H + k.x - 1 - Y0); nx = min(a.Sw, a.W + k.y - 1 - X0);
            } else if (a.corr) {
                ny = min(a.Sh, a.FH - Y0); nx = min(a.Sw, a.FW - X0);
                d = a.outs[(size_t)img * a.out_img_stride + t];
            } else {
                ny = min(a.Sh, a.crop_h - Y0); nx = min(a.Sw, a.crop_w - X0);
                d = a.outs[(size_t)img * a.out_img_stride + t] + (size_t)X0 * a.out_ld + Y0;
            }
        }
        tile_dst[threadIdx.x] = d; tile_ny[threadIdx.x] = ny; tile_nx[threadIdx.x] = nx; tile_ok[threadIdx.x] = ok ? 1 : 0;
    }
    __syncthreads();
    // ---- gather: 33 boxes {4 tiles, 64 columns, 1 row}, one thread, all in flight at once
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, OS_CH * 2048u);
        const int bin0 = (tblk * a.NNB + nblk) * OS_NBIN;                  // tensor {RS, 128 templates, bins}: box {8, 1, 64}
        if (a.dbg & 16) {                                                  // timing experiment: contiguous bytes instead of the gather
            const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            bulk_g2s(buf, reinterpret_cast<const unsigned char*>(a.P) + cta * (OS_CH * 2048u), OS_CH * 2048u, &bar);
        } else
#pragma unroll 1
        for (int u = 0; u < OS_CH; ++u) os_tma_load_3d(buf + u * OS_TROW, &tmap, 8 * g, tl, bin0 + u * 64, &bar);
    }
    if (!(VAR & 1) || warp == 0) mbar_wait(&bar, 0);
    if (VAR & 1) __syncthreads();
    // columns 0 and 32 are spectra of real sequences along h: combine them into one complex column pair
    //     col0[u] := Z[u][0] + i Z[u][32],   col32[u] := Z[u][0] - i Z[u][32]
    if (threadIdx.x < OS_IG * 33) {
        const int q = threadIdx.x & 3, u = threadIdx.x >> 2;
        cpx* c0 = buf + u * OS_TROW + q;
        cpx* c32 = c0 + 4 * 32;
        const cpx za = *c0, zb = *c32;
        *c0 = make_float2(za.x - zb.y, za.y + zb.x);
        *c32 = make_float2(za.x + zb.y, za.y - zb.x);
    }
    __syncthreads();
    const int par = warp >> OS_IGB;                                        // warp-uniform: tasks (par, par + 2)
    // ---- pass 1: inverse along h.  x[u] = col_v[u] (u <= 32), conj(col_mv[64-u]) (u > 32); landed layout in,
    //      per-tile layout out (tile q at q * (VAR & 2 ? OS_ITILE2 : OS_ITILE), column v at os_icol(v))
    {
        const int q = lane & 3;
        const int v = 8 * (warp & 3) + (lane >> 2), mv = v ? 64 - v : 32;
        const cpx* cv = buf + q + 4 * v;                                   // element u: cv[u * OS_TROW]
        const cpx* cm = buf + q + 4 * mv;
        float reA[16], imA[16], reB[16], imB[16];
        auto ld1 = [&](int u) -> cpx {
            if (u <= 32) return cv[u * OS_TROW];
            const cpx z = cm[(64 - u) * OS_TROW];
            return make_float2(z.x, -z.y);
        };
        auto ld = [&](int j) { const cpx z0 = ld1(j), z1 = ld1(j + 1); return make_float4(z0.x, z0.y, z1.x, z1.y); };
        if (par == 0) os_fft64_pair<0, true>(ld, reA, imA, reB, imB);
        else          os_fft64_pair<1, true>(ld, reA, imA, reB, imB);
        __syncthreads();                                                   // every input of the CTA has been read
        cpx* ov = buf + q * (VAR & 2 ? OS_ITILE2 : OS_ITILE) + os_icol(v);
        cpx* om = buf + q * (VAR & 2 ? OS_ITILE2 : OS_ITILE) + os_icol(mv);
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) {                                  // Y[y]: y < 32 -> col v, y >= 32 -> col mv
            const int ya = 4 * j1 + par, yb = ya + 2;                      // (compile-time split: j1 < 8 <=> y < 32)
            if (j1 < 8) { ov[ya] = make_float2(reA[j1], imA[j1]); ov[yb] = make_float2(reB[j1], imB[j1]); }
            else        { om[ya - 32] = make_float2(reA[j1], imA[j1]); om[yb - 32] = make_float2(reB[j1], imB[j1]); }
        }
    }
    __syncthreads();
    // ---- pass 2: C2R along w.  W[v] = Y[y][v] + i Y[y+32][v]; outputs straight to the plane (as os_inverse)
    {
        const int gq = warp & (OS_IG - 1);                                 // one warp = the 32 lines of one tile
        const cpx* tile = buf + gq * (VAR & 2 ? OS_ITILE2 : OS_ITILE);
        const int y = lane;
        if (!tile_ok[gq]) return;                                          // warp-uniform (no barrier below)
        const cpx* ty = tile + y;
        float reA[16], imA[16], reB[16], imB[16];
        auto ld1 = [&](int v) -> cpx {
            if (v == 0) { const cpx p0 = ty[os_icol(0)], p32 = ty[os_icol(32)]; return make_float2(p0.x, p32.x); }
            if (v == 32) { const cpx p0 = ty[os_icol(0)], p32 = ty[os_icol(32)]; return make_float2(p0.y, p32.y); }
            if (v < 32) { const cpx p = ty[os_icol(v)], z = ty[os_icol(64 - v)]; return make_float2(p.x - z.y, p.y + z.x); }
            const cpx p = ty[os_icol(64 - v)], z = ty[os_icol(v)];
            return make_float2(p.x + z.y, z.x - p.y);
        };
        auto ld = [&](int j) { const cpx z0 = ld1(j), z1 = ld1(j + 1); return make_float4(z0.x, z0.y, z1.x, z1.y); };
        auto slot = [&](int v) -> cpx { return ty[os_icol(v)]; };
        if (VAR & 4) {
            if (par == 0) os_c2r64_pair<0>(slot, reA, imA, reB, imB);
            else          os_c2r64_pair<1>(slot, reA, imA, reB, imB);
        } else {
            if (par == 0) os_fft64_pair<0, true>(ld, reA, imA, reB, imB);
            else          os_fft64_pair<1, true>(ld, reA, imA, reB, imB);
        }
        const OsTileOut T{tile_dst[gq], tile_ny[gq], tile_nx[gq], tile_y0[gq], tile_x0[gq], a.out_ld};
        os_inverse_emit(a, t, lane, par, mbase + gq, T, reA, imA, reB, imB);
    }
}

// ------------------------------------------------------------------------------------------------
// os_inverse_z: the inverse with the product spectra landing in four independent ZONES, persistent CTAs.
// Zone z holds the 16 spectrum columns that pass 1 couples: v in [8z, 8z+8) ("direct") and their mirrors 64 - v.  P is
// seen as a 5-D tensor {RS floats, 128 templates, 64 v, 33 u, template block x tile block}; one tensor copy with box
// {8 floats = 4 tiles, 1 template, 8 columns, 33 rows} brings the direct columns of a zone, a second one the mirrors
// (zone 0: columns 56..63, of which 57..63 are mirrors; column 32, the partner of column 0, comes through a third map
// with a one-column box).  9 tensor copies per item instead of 33, one mbarrier per zone:
//   * the two warps of a zone (par 0 / 1) start pass 1 as soon as THEIR 17 KB have landed, and rewrite the zone in place
//     in the per-tile layout pass 2 reads (64 threads, one named barrier) -- no CTA-wide barrier until pass 2;
//   * the item loop lets a CTA run persistently (FFTCONV_OS_INV_PERSIST=1: 3 per SM, the boxes of the next item are
//     requested as soon as pass 2 has its inputs in registers); the default launches one CTA per item, which measured
//     faster (see os_chunk_inverse in fftconv.cu).
// Zone layout (complex units): direct box [u][8 v][4 tiles] at +0, mirror box at +1056; after pass 1
// [4 tiles][16 slots][33] with tile stride 532 (slot s < 8: Y[y][8z+s], slot 8+s: Y[y+32][8z+s], y < 32).
constexpr int OS_ZS = 4 * 532;                       // zone stride (2128 complex >= 2 * 1056)
constexpr int OS_ZC32 = 4 * OS_ZS;                   // column 32: [u][4 tiles]
constexpr int OS_IZ_SMEM = (OS_ZC32 + OS_CH * 4) * 8;

__device__ __forceinline__ void os_tma_load_5d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void os_zone_bar(int z) {                      // 64 threads: the two warps of zone z (ids 1..4)
    switch (z) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}
// complex offset of slot(v) of pass 2 inside a tile's share of the zones (compile-time v)
__host__ __device__ constexpr int os_zslot(int v) {
    return v < 32 ? (v >> 3) * OS_ZS + (v & 7) * 33
         : v == 32 ? 8 * 33
         : ((64 - v) >> 3) * OS_ZS + (8 + ((64 - v) & 7)) * 33;
}

__global__ void __launch_bounds__(256, 3) os_inverse_z(OsInvArgs a, const __grid_constant__ OsTensorMap tmap8,
                                                       const __grid_constant__ OsTensorMap tmap1, int nitems)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(os_smem_raw);
    __shared__ __align__(8) uint64_t full[4];
    __shared__ float* tile_dst[2][OS_IG];
    __shared__ int tile_ny[2][OS_IG], tile_nx[2][OS_IG], tile_y0[2][OS_IG], tile_x0[2][OS_IG], tile_ok[2][OS_IG], tile_ld[2][OS_IG];
    const int NG = a.RS >> 3;
    const int NX = a.NNB * NG;                                             // items per template
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int par = warp >> 2, zq = warp & 3;                              // pass 1: zone zq; pass 2: tile zq
    // (Tried, round 2e: parity = warp & 1, so that the warps of one scheduler partition (warp index mod 4) all run the same
    // parity's specialised first stages and its L0 instruction cache sees one copy of the pass code: 0.244 vs 0.240 ms at
    // config 2, 28.5 vs 27.8 ms at config 4 -- slower.  scripts/exp_invpar.sh.)
    auto request = [&](int item) {                                         // one thread: 9 tensor copies
        const int t = item / NX, bx = item - t * NX;
        const int nblk = bx / NG, g = bx - nblk * NG;
        const int tblk = t / OS_TM, tl = t - tblk * OS_TM;
        const int blk = tblk * a.NNB + nblk;
#pragma unroll
        for (int z = 0; z < 4; ++z) {
            mbar_expect_tx(&full[z], 2u * 8448u + (z == 0 ? 1056u : 0u));
            os_tma_load_5d(buf + z * OS_ZS, &tmap8, 8 * g, tl, 8 * z, 0, blk, &full[z]);
            os_tma_load_5d(buf + z * OS_ZS + 1056, &tmap8, 8 * g, tl, z ? 57 - 8 * z : 56, 0, blk, &full[z]);
            if (z == 0) os_tma_load_5d(buf + OS_ZC32, &tmap1, 8 * g, tl, 32, 0, blk, &full[0]);
        }
    };
    // grid-stride item order: the CTAs resident at any time work on neighbouring tile groups of the same templates, so the
    // 32-byte pieces they request are neighbours in DRAM at about the same time.  (A contiguous share of the item list per
    // CTA makes every CTA walk its own rows: 0.465 instead of 0.31 ms.)
    const int item0 = (int)blockIdx.x, item1 = nitems, istep = (int)gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int z = 0; z < 4; ++z) mbar_init(&full[z], 1);
        fence_barrier_init();
        if (item0 < item1) request(item0);
    }
    __syncthreads();                                                       // the barriers exist before anyone polls them
    int it = 0;
#pragma unroll 1
    for (int item = item0; item < item1; item += istep, ++it) {
        const int sl = it & 1;
        const int t = item / NX, bx = item - t * NX;
        const int nblk = bx / NG, g = bx - nblk * NG;
        const int mbase = nblk * a.NTn + 4 * g;                            // first tile of the group
        if (threadIdx.x >= 32 && threadIdx.x < 32 + OS_IG) {               // (warp 1: warp 0 issues the copies)
            const int i = threadIdx.x - 32;
            const int m = mbase + i;
            const bool ok = 4 * g + i < a.NTn && m < a.NT;
            float* d = nullptr; int ny = 0, nx = 0, ld = a.out_ld;
            if (ok && a.levels) {                     // pyramid batch: planes only (no peak / detection / correlation mode)
                const int l = os_level_of(a.levels, a.nlevels, m);
                const OsLevel lv = a.levels[l];
                const int mt = m - lv.m0;
                const int tj = mt / lv.nth, ti = mt - tj * lv.nth;
                const int Y0 = ti * a.Sh, X0 = tj * a.Sw;
                tile_y0[sl][i] = Y0; tile_x0[sl][i] = X0;
                ld = lv.out_ld;
                ny = min(a.Sh, lv.crop_h - Y0); nx = min(a.Sw, lv.crop_w - X0);
                d = a.outs[(size_t)l * a.out_img_stride + t] + (size_t)X0 * ld + Y0;
            } else if (ok) {
                const int img = m / a.NTimg, mt = m - img * a.NTimg;
                const int tj = mt / a.nth, ti = mt - tj * a.nth;
                const int Y0 = ti * a.Sh, X0 = tj * a.Sw;
                tile_y0[sl][i] = Y0; tile_x0[sl][i] = X0;
                if (a.peak_keys || a.det_mode) {      // region of the full linear convolution of THIS template
                    const int2 k = a.khw[t];
                    ny = min(a.Sh, a.H + k.x - 1 - Y0); nx = min(a.Sw, a.W + k.y - 1 - X0);
                } else if (a.corr) {
                    ny = min(a.Sh, a.FH - Y0); nx = min(a.Sw, a.FW - X0);
                    d = a.outs[(size_t)img * a.out_img_stride + t];
                } else {
                    ny = min(a.Sh, a.crop_h - Y0); nx = min(a.Sw, a.crop_w - X0);
                    d = a.outs[(size_t)img * a.out_img_stride + t] + (size_t)X0 * a.out_ld + Y0;
                }
            }
            tile_dst[sl][i] = d; tile_ny[sl][i] = ny; tile_nx[sl][i] = nx; tile_ok[sl][i] = ok ? 1 : 0; tile_ld[sl][i] = ld;
        }
        float reA[16], imA[16], reB[16], imB[16];
        // Both passes run through ONE copy of the 16-point register transforms (a two-trip loop that stays rolled); only
        // the first radix-4 stages are specialised by pass and parity.  With everything inlined per parity the kernel
        // carried 61 KB of live straight-line code against a 32 KB instruction cache: once the persistent loop had
        // removed the wait for the boxes, `no_instruction` became 37 % of all stall samples.
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0) {
                // ---- pass 1 (zone zq): inverse along h.  x[u] = col_v[u] (u <= 32), conj(col_mv[64-u]) (u > 32)
                cpx* Z = buf + zq * OS_ZS;
                const int q = lane & 3, vl = lane >> 2;
                mbar_wait(&full[zq], (uint32_t)sl);
                if (zq == 0) {
                    // columns 0 and 32 are spectra of real sequences along h: combine them into one complex column pair
                    //     col0[u] := Z[u][0] + i Z[u][32],   col32[u] := Z[u][0] - i Z[u][32]     (col32 takes the unused
                    //     slot of column 56 in the mirror box)
                    for (int i = par * 32 + lane; i < OS_CH * 4; i += 64) {
                        const int u = i >> 2, qq = i & 3;
                        cpx* c0 = Z + u * 32 + qq;
                        const cpx za = *c0, zb = buf[OS_ZC32 + i];
                        *c0 = make_float2(za.x - zb.y, za.y + zb.x);
                        c0[1056] = make_float2(za.x + zb.y, za.y - zb.x);
                    }
                    os_zone_bar(0);
                }
                const int ml = zq ? 7 - vl : ((8 - vl) & 7);
                const cpx* cv = Z + vl * 4 + q;                            // element u: cv[u * 32]
                const cpx* cm = Z + 1056 + ml * 4 + q;
                auto ld1 = [&](int u) -> cpx {
                    if (u <= 32) return cv[u * 32];
                    const cpx z = cm[(64 - u) * 32];
                    return make_float2(z.x, -z.y);
                };
                auto ld = [&](int j) { const cpx z0 = ld1(j), z1 = ld1(j + 1); return make_float4(z0.x, z0.y, z1.x, z1.y); };
                if (par == 0) os_fft64_pair<0, true, decltype(ld)&, false>(ld, reA, imA, reB, imB);
                else          os_fft64_pair<1, true, decltype(ld)&, false>(ld, reA, imA, reB, imB);
            } else {
                // ---- pass 2 (tile zq): C2R along w, rows y and y + 32 packed; one pair of loads gives W[v] and W[64 - v]
                const cpx* ty = buf + zq * 532 + lane;
                auto slot = [&](int v) -> cpx { return ty[os_zslot(v)]; };
                if (par == 0) os_c2r64_pair<0, decltype(slot)&, false>(slot, reA, imA, reB, imB);
                else          os_c2r64_pair<1, decltype(slot)&, false>(slot, reA, imA, reB, imB);
            }
            dft_regs<16, true>(reA, imA);
            dft_regs<16, true>(reB, imB);
            if (pass == 0) {
                os_zone_bar(zq);                                           // both warps of the zone have read their inputs
                cpx* ov = buf + zq * OS_ZS + (lane & 3) * 532 + (lane >> 2) * 33 + par;
#pragma unroll
                for (int j1 = 0; j1 < 16; ++j1) {                          // Y[y]: y < 32 -> slot vl, y >= 32 -> slot 8 + vl
                    constexpr int dummy = 0; (void)dummy;
                    const int o = j1 < 8 ? 4 * j1 : 8 * 33 + 4 * j1 - 32;  // y = 4 j1 + par (+ 2)
                    ov[o] = make_float2(reA[j1], imA[j1]); ov[o + 2] = make_float2(reB[j1], imB[j1]);
                }
                __syncthreads();
            }
        }
        __syncthreads();                                                   // the zones are free: request the next boxes
        if (threadIdx.x == 0 && item + istep < item1) {
            fence_proxy_async();
            request(item + istep);
        }
        if (tile_ok[sl][zq]) {                                             // warp-uniform
            const OsTileOut T{tile_dst[sl][zq], tile_ny[sl][zq], tile_nx[sl][zq], tile_y0[sl][zq], tile_x0[sl][zq], tile_ld[sl][zq]};
            os_inverse_emit(a, t, lane, par, mbase + zq, T, reA, imA, reB, imB);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// inv_w_pass: inverse complex FFT along w of the compat spectrum S [plane][FW][CH] -> Z [plane][FW][CH]
// (first half of spectrum -> plane; inv_h_pass finishes).  grid = (ceil(CH/TU), planes).
__global__ void inv_w_pass(const cpx* __restrict__ S, int FW, int CH, LinePlan plan, const cpx* __restrict__ tw,
                           cpx* __restrict__ Z, int TU, int ld, const unsigned long long* __restrict__ skip_if_equal)
{
    if (skip_if_equal && skip_if_equal[0] == skip_if_equal[1]) return;     // (see inv_h_pass)
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx* b0 = reinterpret_cast<cpx*>(os_smem_raw);
    cpx* b1 = b0 + (size_t)TU * ld;
    const int u0 = blockIdx.x * TU;
    const size_t p = blockIdx.y;
    const cpx* Sp = S + p * (size_t)FW * CH;
    for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
        const int x = idx / TU, u = idx - x * TU;
        cpx v = make_float2(0.f, 0.f);
        if (u0 + u < CH) v = Sp[(size_t)x * CH + u0 + u];
        b0[(size_t)u * ld + x] = v;
    }
    __syncthreads();
    const cpx* res = fft_lines<true>(b0, b1, TU, ld, plan, tw);
    cpx* Zp = Z + p * (size_t)FW * CH;
    for (int idx = threadIdx.x; idx < TU * FW; idx += blockDim.x) {
        const int x = idx / TU, u = idx - x * TU;
        if (u0 + u < CH) Zp[(size_t)x * CH + u0 + u] = res[(size_t)u * ld + x];
    }
}

}  // namespace fftconv
