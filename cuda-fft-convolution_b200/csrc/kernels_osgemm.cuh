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
//   os_hpass<TEMPLATES|DATA>   pad / window gather fused into the load, 64-point real FFT along h
//   os_wpass<A|B>              64-point FFT along w, hi/lo split, store as tcgen05 operand images
//   os_gemm                    TMA bulk copies -> smem operand images -> tcgen05.mma (kind::tf32,
//                              M=128 templates, N=2*tiles, K=2*F) -> TMEM -> bulk store of P
//   os_inverse                 per (template, tile): 2-D C2R inverse, scale, valid-region store (crop fused)
//   inv_w_pass                 (spectrum -> plane, when the caller hands in a cudaFFTData spectrum)
//
// Replaces the same reference rows as kernels_tile16.cuh (padData, cufftExecR2C,
// elementwiseProductAndNormalize, F x cufftExecC2R, sumAlongFeatures: src/cudaConvFFTData.cuh:11-92,
// src/cudaConvFFTData.cu:233-271).
//
// Operand images (K-major, no swizzle; one 16-byte unit = 4 consecutive k = channels (2c,2c+1) x (re,im)):
//   Aimg [tblk][bin][ks][term hi/lo][kc][128 templates][4]      (MMA A: rows = templates)
//   Bimg [nblk][bin][ks][term hi/lo][kc][NMMA rows     ][4]      (MMA B: rows = (tile, re/im column))
//   P    [tblk][nblk][bin][128 templates][RS]  fp32, (re,im) per tile
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

__device__ __forceinline__ int os_wrap(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// ------------------------------------------------------------------------------------------------
// os_hpass: 64-point real-to-half-complex transform along h of every column of every plane.
//   MODE 0 (templates): plane = (template, channel), XC = 16*NF columns, 16*NF rows, zero pad fused
//   MODE 1 (data tiles): plane = (tile, channel), 64 columns x 64 rows gathered from the source plane
//                        with zero fill beyond (srcH, srcW) and circular wrap at (FH, FW)
// Two real columns are packed into one complex line; a line is 4 tasks (r0 = 0..3) on 4 warps.
// out H: [plane][33][XC] complex.   grid.x = ceil(planes * XC/2 / 64), 256 threads.
struct OsHArgs {
    const SrcDesc* descs;   // MODE 0: one per template
    SrcDesc src;            // MODE 1: the source plane [F][cols][rows]
    int nitems, F, XC;
    cpx* H;
    int FH, FW, nth, Sh, Sw, oy0, ox0;    // MODE 1 only
};

template <int MODE, int NF>
__global__ void __launch_bounds__(256) os_hpass(OsHArgs a)
{
    __shared__ cpx Zs[64][65];
    const int ncp = a.XC >> 1;
    const long long nlines = (long long)a.nitems * a.F * ncp;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int r0 = wq & 3;
    const int ll = (wq >> 2) * 32 + lane;
    const long long line = (long long)blockIdx.x * 64 + ll;
    if (line < nlines) {
        const int cp = (int)(line % ncp);
        const long long plane = line / ncp;
        const int f = (int)(plane % a.F);
        const int item = (int)(plane / a.F);
        float re[16], im[16];
        if (MODE == 0) {
            const SrcDesc d = a.descs[item];
            const int xa = 2 * cp, xb = xa + 1;
            const float* pa = d.ptr + ((size_t)f * d.cols + xa) * d.rows;
            const float* pb = pa + d.rows;
            const bool va = xa < d.cols, vb = xb < d.cols;
            const int rows = d.rows;
            auto ld = [&](int j) {
                float4 v;
                v.x = (va && j < rows) ? __ldg(pa + j) : 0.f;
                v.y = (vb && j < rows) ? __ldg(pb + j) : 0.f;
                v.z = (va && j + 1 < rows) ? __ldg(pa + j + 1) : 0.f;
                v.w = (vb && j + 1 < rows) ? __ldg(pb + j + 1) : 0.f;
                return v;
            };
            os_fft64_task_rt<NF, false>(r0, ld, re, im);
        } else {
            const SrcDesc d = a.src;
            const int ti = item % a.nth, tj = item / a.nth;
            const int oy = ti * a.Sh - a.oy0, ox = tj * a.Sw - a.ox0;
            const int gxa = os_wrap(ox + 2 * cp, a.FW), gxb = os_wrap(ox + 2 * cp + 1, a.FW);
            const bool va = gxa < d.cols, vb = gxb < d.cols;
            const float* pa = d.ptr + ((size_t)f * d.cols + gxa) * d.rows;
            const float* pb = d.ptr + ((size_t)f * d.cols + gxb) * d.rows;
            const int rows = d.rows, FH = a.FH;
            auto ld = [&](int j) {
                const int g0 = os_wrap(oy + j, FH), g1 = os_wrap(oy + j + 1, FH);
                float4 v;
                v.x = (va && g0 < rows) ? __ldg(pa + g0) : 0.f;
                v.y = (vb && g0 < rows) ? __ldg(pb + g0) : 0.f;
                v.z = (va && g1 < rows) ? __ldg(pa + g1) : 0.f;
                v.w = (vb && g1 < rows) ? __ldg(pb + g1) : 0.f;
                return v;
            };
            os_fft64_task_rt<NF, false>(r0, ld, re, im);
        }
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) Zs[ll][4 * j1 + r0] = make_float2(re[j1], im[j1]);
    }
    __syncthreads();
    // split the packed pair:  A[u] = (Z[u] + conj Z[64-u])/2,  B[u] = -i (Z[u] - conj Z[64-u])/2
    const int ppc = 64 / ncp;                                 // whole planes per CTA
    const long long nplanes = (long long)a.nitems * a.F;
    const int total = ppc * OS_CH * a.XC;
    for (int idx = threadIdx.x; idx < total; idx += 256) {
        const int x = idx % a.XC;
        const int u = (idx / a.XC) % OS_CH;
        const int pl = idx / (a.XC * OS_CH);
        const long long plane = (long long)blockIdx.x * ppc + pl;
        if (plane >= nplanes) continue;
        const int l2 = pl * ncp + (x >> 1);
        const cpx zu = Zs[l2][u];
        const cpx zn = Zs[l2][(64 - u) & 63];
        cpx o;
        if (x & 1) o = make_float2(0.5f * (zu.y + zn.y), -0.5f * (zu.x - zn.x));
        else       o = make_float2(0.5f * (zu.x + zn.x), 0.5f * (zu.y - zn.y));
        a.H[((size_t)plane * OS_CH + u) * a.XC + x] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// os_wpass: 64-point transform along w of the rows of H, written as tcgen05 operand images.
//   MODE 0 -> Aimg (rows = templates);  MODE 1 -> Bimg (rows = (tile, re/im column))
// CTA = 32 slots (lane = slot) x one channel pair x R=2 spectrum rows; the 16 outputs of a task go through
// a shared staging buffer so that the image is written in 512-byte (A) / 1-KB (B) runs.
// grid = (ceil(nblk*slots_per_blk/32), NKS*KC channel pairs, ceil(33/2)); 256 threads; smem 64 KB.
struct OsWArgs {
    const cpx* H;           // [item][F][33][XC]
    int F, XC;
    float* img;
    int NKS, KC;
    int rows;               // rows of one operand block (A: 128, B: NMMA)
    int nblk;               // operand blocks (A: template blocks, B: tile blocks)
    int slots_per_blk;      // A: 128, B: NMMA/2
    int valid_per_blk;      // A: 128, B: tiles per block
    int nvalid;             // A: templates in this chunk, B: NT
    int correlate;
};
constexpr int OS_WR = 2;

__device__ __forceinline__ float os_tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int MODE, int NF>
__global__ void __launch_bounds__(256) os_wpass(OsWArgs a)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx (*stag)[2][32] = reinterpret_cast<cpx (*)[2][32]>(os_smem_raw);      // [R*64 bins][2 channels][32 slots]
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int slot0 = blockIdx.x * 32, fp = blockIdx.y, u0 = blockIdx.z * OS_WR;
    int item = -1;
    {
        const int slot = slot0 + lane;
        const int blk = slot / a.slots_per_blk, sl = slot - blk * a.slots_per_blk;
        const int it = blk * a.valid_per_blk + sl;
        if (blk < a.nblk && sl < a.valid_per_blk && it < a.nvalid) item = it;
    }
    for (int wt = wq; wt < 2 * OS_WR * 4; wt += 8) {
        const int r0 = wt & 3, fq = (wt >> 2) & 1, r = wt >> 3;
        const int u = u0 + r, f = 2 * fp + fq;
        float re[16], im[16];
        if (item >= 0 && f < a.F && u < OS_CH) {
            const float4* row = reinterpret_cast<const float4*>(a.H + (((size_t)item * a.F + f) * OS_CH + u) * a.XC);
            auto ld = [&](int j) { return __ldg(row + (j >> 1)); };
            os_fft64_task_rt<NF, false>(r0, ld, re, im);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { re[j] = 0.f; im[j] = 0.f; }
        }
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) stag[r * 64 + 4 * j1 + r0][fq][lane] = make_float2(re[j1], im[j1]);
    }
    __syncthreads();
    const int ks = fp / a.KC, kc = fp - ks * a.KC;
    const size_t term_stride = (size_t)a.KC * a.rows * 4;                 // floats between hi and lo images
    for (int idx = threadIdx.x; idx < OS_WR * 64 * 32; idx += 256) {
        const int sl_lane = idx & 31, bl = idx >> 5;
        const int u = u0 + (bl >> 6);
        if (u >= OS_CH) continue;
        const int slot = slot0 + sl_lane;
        const int blk = slot / a.slots_per_blk, sl = slot - blk * a.slots_per_blk;
        if (blk >= a.nblk) continue;
        const int bin = u * 64 + (bl & 63);
        const cpx c0 = stag[bl][0][sl_lane], c1 = stag[bl][1][sl_lane];
        float* base = a.img + ((((size_t)blk * OS_NBIN + bin) * a.NKS + ks) * 2) * term_stride + (size_t)kc * a.rows * 4;
        if (MODE == 0) {
            const float4 v = make_float4(c0.x, c0.y, c1.x, c1.y);
            const float4 hi = make_float4(os_tf32_hi(v.x), os_tf32_hi(v.y), os_tf32_hi(v.z), os_tf32_hi(v.w));
            const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
            float4* o = reinterpret_cast<float4*>(base + (size_t)sl * 4);
            o[0] = hi;
            *reinterpret_cast<float4*>(base + term_stride + (size_t)sl * 4) = lo;
        } else {
            // out_re = sum a*c - b*d, out_im = sum a*d + b*c  with K^ = a + i b (A rows), D^ = c + i d
            // correlate: K^ -> conj(K^):  out_re = sum a*c + b*d, out_im = sum a*d - b*c
            float4 vr, vi;
            if (!a.correlate) { vr = make_float4(c0.x, -c0.y, c1.x, -c1.y); vi = make_float4(c0.y, c0.x, c1.y, c1.x); }
            else              { vr = make_float4(c0.x, c0.y, c1.x, c1.y);   vi = make_float4(c0.y, -c0.x, c1.y, -c1.x); }
            const float4 hr = make_float4(os_tf32_hi(vr.x), os_tf32_hi(vr.y), os_tf32_hi(vr.z), os_tf32_hi(vr.w));
            const float4 hq = make_float4(os_tf32_hi(vi.x), os_tf32_hi(vi.y), os_tf32_hi(vi.z), os_tf32_hi(vi.w));
            float4* o = reinterpret_cast<float4*>(base + (size_t)(2 * sl) * 4);
            o[0] = hr; o[1] = hq;
            float4* ol = reinterpret_cast<float4*>(base + term_stride + (size_t)(2 * sl) * 4);
            ol[0] = make_float4(vr.x - hr.x, vr.y - hr.y, vr.z - hr.z, vr.w - hr.w);
            ol[1] = make_float4(vi.x - hq.x, vi.y - hq.y, vi.z - hq.z, vi.w - hq.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// os_kern_fft: templates -> A operand images in ONE kernel (pad fused into the load, 2-D 64 x 64 half
// spectrum of a 16*NF x 16*NF support, hi/lo TF32 split, operand-image store).  Replaces the template
// legs of os_hpass + os_wpass and their [template][F][33][XC] intermediate.
// CTA = 16 templates x one channel pair, 256 threads.  Two phases (odd spectrum rows, then even ones) so
// that the h-transformed rows of the 32 planes fit 70 KB (NF=1) of shared memory:
//   h step: thread = (plane, column pair): two real columns ride as one complex sequence; the two tasks
//           that hold rows u and 64-u are computed together and unpacked in registers
//   w step: thread = (spectrum row, task r0, template); both channels of the pair, then 16 x (hi, lo)
//           128-bit stores; 16 consecutive lanes write 256 contiguous bytes of the image
// grid = (ntblk*128/16, NKS*KC).
struct OsKArgs {
    const SrcDesc* descs;
    int nk, F;
    float* img;
    int NKS, KC;
};
constexpr int OS_KSL = 16;            // templates per CTA

template <int NF>
__global__ void __launch_bounds__(256) os_kern_fft(OsKArgs a)
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
    const size_t term_stride = (size_t)a.KC * OS_TM * 4;
    const size_t bin_stride = (size_t)a.NKS * 2 * term_stride;

#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
        // ---- h step
        for (int line = threadIdx.x; line < 32 * NCP; line += 256) {
            const int cp = line % NCP, pl = line / NCP;
            const int ch = pl >> 4, slot = pl & 15;
            const int item = slot0 + slot, f = 2 * fp + ch;
            float4* hrow = reinterpret_cast<float4*>(Hs + (size_t)pl * PS) + cp;      // + ro * (XC/2)
            if (item < a.nk && f < a.F) {
                const SrcDesc d = a.descs[item];
                const int xa = 2 * cp, xb = xa + 1;
                const float* pa = d.ptr + ((size_t)f * d.cols + xa) * d.rows;
                const float* pb = pa + d.rows;
                const bool va = xa < d.cols, vb = xb < d.cols;
                const int rows = d.rows;
                auto ld = [&](int j) {
                    float4 v;
                    v.x = (va && j < rows) ? __ldg(pa + j) : 0.f;
                    v.y = (vb && j < rows) ? __ldg(pb + j) : 0.f;
                    v.z = (va && j + 1 < rows) ? __ldg(pa + j + 1) : 0.f;
                    v.w = (vb && j + 1 < rows) ? __ldg(pb + j + 1) : 0.f;
                    return v;
                };
                float pr[16], pi[16], qr[16], qi[16];
                // A[u] = (Z[u] + conj Z[64-u]) / 2 (column xa),  B[u] = -i (Z[u] - conj Z[64-u]) / 2 (column xb)
                auto put = [&](int ro, float zr, float zi, float nr, float ni) {
                    hrow[ro * NCP] = make_float4(0.5f * (zr + nr), 0.5f * (zi - ni), 0.5f * (zi + ni), -0.5f * (zr - nr));
                };
                if (ph == 0) {
                    os_fft64_task<1, NF, false>(ld, pr, pi);          // Z[4 j1 + 1]
                    os_fft64_task<3, NF, false>(ld, qr, qi);          // Z[4 j1 + 3]
#pragma unroll
                    for (int j1 = 0; j1 < 8; ++j1) {
                        put(2 * j1, pr[j1], pi[j1], qr[15 - j1], qi[15 - j1]);          // u = 4 j1 + 1
                        put(2 * j1 + 1, qr[j1], qi[j1], pr[15 - j1], pi[15 - j1]);      // u = 4 j1 + 3
                    }
                } else {
                    os_fft64_task<0, NF, false>(ld, pr, pi);          // Z[4 j1]
                    os_fft64_task<2, NF, false>(ld, qr, qi);          // Z[4 j1 + 2]
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
            float re0[16], im0[16], re1[16], im1[16];
            {
                const float4* row = reinterpret_cast<const float4*>(Hs + (size_t)slot * PS + ro * XC);
                auto ld = [&](int j) { return row[j >> 1]; };
                os_fft64_task_rt<NF, false>(r0, ld, re0, im0);
            }
            {
                const float4* row = reinterpret_cast<const float4*>(Hs + (size_t)(16 + slot) * PS + ro * XC);
                auto ld = [&](int j) { return row[j >> 1]; };
                os_fft64_task_rt<NF, false>(r0, ld, re1, im1);
            }
            float* base = a.img + ((size_t)tblk * OS_NBIN + (size_t)u * 64 + r0) * bin_stride + (size_t)ks * 2 * term_stride +
                          (size_t)kc * OS_TM * 4 + (size_t)(sl0 + slot) * 4;
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) {
                const float4 v = make_float4(re0[j1], im0[j1], re1[j1], im1[j1]);
                const float4 hi = make_float4(os_tf32_hi(v.x), os_tf32_hi(v.y), os_tf32_hi(v.z), os_tf32_hi(v.w));
                const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                float* o = base + (size_t)(4 * j1) * bin_stride;
                *reinterpret_cast<float4*>(o) = hi;
                *reinterpret_cast<float4*>(o + term_stride) = lo;
            }
        }
        __syncthreads();
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
// os_gemm: persistent, warp-specialised.  Work item = (tile block, bin, template block); a CTA owns a
// contiguous range of items ordered so that consecutive items share the B operand (same tile block & bin).
//   warp 0 : TMA producer (cp.async.bulk + mbarrier expect_tx), A ring of `nsta` K-stages, 2 B buffers
//   warp 1 : TMEM allocation + single-thread tcgen05.mma issue; 3 passes (lo*hi, hi*lo, hi*hi) per K-stage
//   warps 2-5 : epilogue, tcgen05.ld -> registers -> smem staging -> one bulk store of the 128 x RS block
struct OsGemmArgs {
    const float* Aimg;
    const float* Bimg;
    float* P;
    int NTBLK, NNB, NKS, KC, NMMA, RS;
    long long nitems;
    int nsta;
    int lbo_swap;     // debug: swap the LBO / SBO fields of the smem descriptors
};

__global__ void __launch_bounds__(192, 1) os_gemm(OsGemmArgs g)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    const uint32_t a_stage = 2u * g.KC * OS_TM * 16u;
    const uint32_t b_stage = 2u * g.KC * g.NMMA * 16u;
    const uint32_t b_buf = b_stage * g.NKS;
    const uint32_t p_blk = (uint32_t)OS_TM * g.RS * 4u;
    unsigned char* a_sm = os_smem_raw;
    unsigned char* b_sm = a_sm + (size_t)g.nsta * a_stage;
    float* stage_sm = reinterpret_cast<float*>(b_sm + 2 * (size_t)b_buf);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(stage_sm) + p_blk);
    uint64_t* a_full = bars;                 // [nsta]
    uint64_t* a_empty = bars + 8;            // [nsta]
    uint64_t* b_full = bars + 16;            // [2]
    uint64_t* b_empty = bars + 18;           // [2]
    uint64_t* acc_full = bars + 20;          // [2]
    uint64_t* acc_empty = bars + 22;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long lo = g.nitems * (long long)blockIdx.x / gridDim.x;
    const long long hi = g.nitems * (long long)(blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < g.nsta; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128);
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
            for (long long it = lo; it < hi; ++it) {
                const long long key = it / g.NTBLK;
                const int tblk = (int)(it - key * g.NTBLK);
                const int bin = (int)(key % OS_NBIN);
                if (key != curkey) {
                    const uint32_t bb = nb & 1;
                    if (nb >= 2) mbar_wait(&b_empty[bb], ((nb >> 1) - 1) & 1);
                    mbar_expect_tx(&b_full[bb], b_buf);
                    bulk_g2s(b_sm + (size_t)bb * b_buf, reinterpret_cast<const unsigned char*>(g.Bimg) + (size_t)key * b_buf,
                             b_buf, &b_full[bb]);
                    ++nb;
                    curkey = key;
                }
                const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.Aimg) +
                                            ((size_t)tblk * OS_NBIN + bin) * g.NKS * a_stage;
                for (int ks = 0; ks < g.NKS; ++ks) {
                    const uint32_t st = na % g.nsta, fill = na / g.nsta;
                    if (fill >= 1) mbar_wait(&a_empty[st], (fill - 1) & 1);
                    mbar_expect_tx(&a_full[st], a_stage);
                    bulk_g2s(a_sm + (size_t)st * a_stage, asrc + (size_t)ks * a_stage, a_stage, &a_full[st]);
                    ++na;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = os_idesc_tf32(OS_TM, g.NMMA);
            const uint32_t a_lbo = OS_TM * 16u, b_lbo = (uint32_t)g.NMMA * 16u, sbo = 128u;
            long long curkey = -1;
            uint32_t nb = 0, na = 0, nit = 0, bcur = 0;
            for (long long it = lo; it < hi; ++it, ++nit) {
                const long long key = it / g.NTBLK;
                if (key != curkey) {
                    bcur = nb & 1;
                    mbar_wait(&b_full[bcur], (nb >> 1) & 1);
                    ++nb;
                    curkey = key;
                }
                const uint32_t acc = nit & 1;
                if (nit >= 2) mbar_wait(&acc_empty[acc], ((nit >> 1) - 1) & 1);
                os_tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * OS_ACC_COLS;
                for (int ks = 0; ks < g.NKS; ++ks) {
                    const uint32_t st = na % g.nsta;
                    mbar_wait(&a_full[st], (na / g.nsta) & 1);
                    os_tc_fence_after();
                    const uint32_t a_base = smem_u32(a_sm + (size_t)st * a_stage);
                    const uint32_t b_base = smem_u32(b_sm + (size_t)bcur * b_buf + (size_t)ks * b_stage);
                    const uint32_t a_term = (uint32_t)g.KC * OS_TM * 16u, b_term = (uint32_t)g.KC * g.NMMA * 16u;
#pragma unroll 1
                    for (int pass = 0; pass < 3; ++pass) {
                        // small terms first: (A lo, B hi), (A hi, B lo), then (A hi, B hi)
                        const uint32_t ta = pass == 0 ? 1u : 0u, tb = pass == 1 ? 1u : 0u;
                        for (int j = 0; j < g.KC / 2; ++j) {
                            const uint32_t aaddr = a_base + ta * a_term + (uint32_t)j * 2u * a_lbo;
                            const uint32_t baddr = b_base + tb * b_term + (uint32_t)j * 2u * b_lbo;
                            const uint64_t ad = g.lbo_swap ? os_smem_desc(aaddr, sbo, a_lbo) : os_smem_desc(aaddr, a_lbo, sbo);
                            const uint64_t bd = g.lbo_swap ? os_smem_desc(baddr, sbo, b_lbo) : os_smem_desc(baddr, b_lbo, sbo);
                            os_mma_tf32(tmem_d, ad, bd, idesc, (ks | pass | j) != 0 ? 1u : 0u);
                        }
                    }
                    os_mma_commit(&a_empty[st]);
                    ++na;
                }
                os_mma_commit(&acc_full[acc]);
                if (it + 1 == hi || (it + 1) / g.NTBLK != key) os_mma_commit(&b_empty[bcur]);
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;              // 0..127
        uint32_t nit = 0;
        for (long long it = lo; it < hi; ++it, ++nit) {
            const long long key = it / g.NTBLK;
            const int tblk = (int)(it - key * g.NTBLK);
            const int bin = (int)(key % OS_NBIN);
            const int nblk = (int)(key / OS_NBIN);
            const uint32_t acc = nit & 1;
            mbar_wait(&acc_full[acc], (nit >> 1) & 1);
            os_tc_fence_after();
            if (et == 0) os_bulk_wait_read0();        // previous bulk store has finished reading the staging
            os_named_bar_sync(1, 128);
            const uint32_t taddr = tmem_base + acc * OS_ACC_COLS + ((uint32_t)(q * 32) << 16);
            float* srow = stage_sm + (size_t)row * g.RS;
            for (int c0 = 0; c0 < g.RS; c0 += 16) {
                uint32_t r[16];
                os_tmem_ld8(taddr + c0, r);
                const bool two = c0 + 8 < g.RS;
                if (two) os_tmem_ld8(taddr + c0 + 8, r + 8);
                os_tmem_ld_wait();
                float4* o = reinterpret_cast<float4*>(srow + c0);
                o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                if (two) {
                    o[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
                    o[3] = make_float4(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
                }
            }
            os_tc_fence_before();
            os_mbar_arrive(&acc_empty[acc]);
            fence_proxy_async();
            os_named_bar_sync(2, 128);
            if (et == 0) {
                float* dst = g.P + ((size_t)((size_t)tblk * g.NNB + nblk) * OS_NBIN + bin) * OS_TM * g.RS;
                os_bulk_s2g(dst, stage_sm, p_blk);
            }
        }
        if (et == 0) os_bulk_wait0();
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
    const int bin = (int)(key % OS_NBIN);
    const int nblk = (int)(key / OS_NBIN);
    const size_t a_stage = (size_t)2 * g.KC * OS_TM * 4, b_stage = (size_t)2 * g.KC * g.NMMA * 4;   // floats
    const float* A = g.Aimg + ((size_t)tblk * OS_NBIN + bin) * g.NKS * a_stage;
    const float* B = g.Bimg + (size_t)key * g.NKS * b_stage;
    float* P = g.P + ((size_t)((size_t)tblk * g.NNB + nblk) * OS_NBIN + bin) * OS_TM * g.RS;
    const int t = threadIdx.x;
    for (int n = 0; n < g.RS; ++n) {
        float acc = 0.f;
        for (int ks = 0; ks < g.NKS; ++ks)
            for (int kc = 0; kc < g.KC; ++kc) {
                const float4 ah = *reinterpret_cast<const float4*>(A + ks * a_stage + ((size_t)kc * OS_TM + t) * 4);
                const float4 al = *reinterpret_cast<const float4*>(A + ks * a_stage + ((size_t)(g.KC + kc) * OS_TM + t) * 4);
                const float4 bh = *reinterpret_cast<const float4*>(B + ks * b_stage + ((size_t)kc * g.NMMA + n) * 4);
                const float4 bl = *reinterpret_cast<const float4*>(B + ks * b_stage + ((size_t)(g.KC + kc) * g.NMMA + n) * 4);
                acc = fmaf(ah.x + al.x, bh.x + bl.x, acc);
                acc = fmaf(ah.y + al.y, bh.y + bl.y, acc);
                acc = fmaf(ah.z + al.z, bh.z + bl.z, acc);
                acc = fmaf(ah.w + al.w, bh.w + bl.w, acc);
            }
        P[(size_t)t * g.RS + n] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// os_inverse: CTA = (template, group of 4 tiles).  Per tile: gather the 33 x 64 product spectrum from P,
// inverse 64-point transform along w (rows u = 1..31 complex; rows 0 and 32 are spectra of real sequences
// and ride together as one complex row), then C2R along h (two columns per complex line), scale by
// 1/4096, and store the valid (65-maxkh) x (65-maxkw) block of the tile into the output plane.
// grid = (ceil(NT/4), templates in chunk); 512 threads; smem = 4*32*66*8 + 4*Sw*(Sh|1)*4.
struct OsInvArgs {
    const float* P;
    float* const* outs;
    int nk, NNB, NTn, RS, NT, nth, Sh, Sw, oy0, ox0;
    int FH, FW, crop_h, crop_w, out_ld;
    float scale;
};
constexpr int OS_IG = 4;          // tiles per CTA
constexpr int OS_IROW = 66;       // complex row stride (16-byte aligned rows, conflict-free LDS.128)

__global__ void __launch_bounds__(512) os_inverse(OsInvArgs a)
{
    extern __shared__ __align__(128) unsigned char os_smem_raw[];
    cpx* buf = reinterpret_cast<cpx*>(os_smem_raw);                        // [4][32][66]
    float* ostage = reinterpret_cast<float*>(os_smem_raw);                 // [4][Sw][shp], reuses buf after the C2R reads
    __shared__ int tile_y0[OS_IG], tile_x0[OS_IG];
    const int shp = a.Sh | 1;
    const int t = blockIdx.y;
    const int m0 = blockIdx.x * OS_IG;
    const int tblk = t / OS_TM, tl = t - tblk * OS_TM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x < OS_IG) {
        const int m = m0 + threadIdx.x;
        const int tj = m / a.nth, ti = m - tj * a.nth;
        tile_y0[threadIdx.x] = m < a.NT ? ti * a.Sh : a.crop_h;            // out of range -> nothing stored
        tile_x0[threadIdx.x] = m < a.NT ? tj * a.Sw : a.crop_w;
    }
    // ---- gather: thread = (tile gq, column v, row parity); 4 consecutive threads read one 32-byte sector of P;
    //      all 17 loads of a thread are in flight together
    {
        const int gq = threadIdx.x & 3, v = (threadIdx.x >> 2) & 63, rh = threadIdx.x >> 8;
        const int m = m0 + gq;
        cpx z[16];
        if (m < a.NT) {
            const int nblk = m / a.NTn, ml = m - nblk * a.NTn;
            const size_t bstride = (size_t)OS_TM * a.RS;
            const float* prow = a.P + ((size_t)((size_t)tblk * a.NNB + nblk) * OS_NBIN * OS_TM + tl) * a.RS + 2 * ml +
                                (size_t)(rh * 64 + v) * bstride;
#pragma unroll
            for (int it = 0; it < 16; ++it) z[it] = __ldg(reinterpret_cast<const cpx*>(prow + (size_t)(it * 128) * bstride));
            if (rh == 0) {
                const cpx x32 = __ldg(reinterpret_cast<const cpx*>(prow + (size_t)(32 * 64) * bstride));
                z[0] = make_float2(z[0].x - x32.y, z[0].y + x32.x);        // X0 + i X32
            }
        } else {
#pragma unroll
            for (int it = 0; it < 16; ++it) z[it] = make_float2(0.f, 0.f);
        }
        cpx* dst = buf + (gq * 32 + rh) * OS_IROW + v;
#pragma unroll
        for (int it = 0; it < 16; ++it) dst[it * 2 * OS_IROW] = z[it];
    }
    __syncthreads();
    const int r0 = warp & 3;
    const int line = (warp >> 2) * 32 + lane;                              // 0..127
    // ---- inverse along w, in place: line = (tile, row), 4 tasks per line on 4 warps
    {
        cpx* rowp = buf + line * OS_IROW;
        float re[16], im[16];
        auto ld = [&](int j) { return *reinterpret_cast<const float4*>(rowp + j); };
        os_fft64_task_rt<4, true>(r0, ld, re, im);
        __syncthreads();
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) rowp[4 * j1 + r0] = make_float2(re[j1], im[j1]);
    }
    __syncthreads();
    // ---- C2R along h: line = (tile, column pair), only column pairs that hold valid outputs
    {
        const int gq = line >> 5, p = line & 31;
        const int xa = 2 * p;
        const bool active = xa + 1 >= a.ox0 && m0 + gq < a.NT;
        float re[16], im[16];
        if (active) {
            const cpx* tb = buf + gq * 32 * OS_IROW + xa;
            // Z[j] = Ya[j] + i Yb[j] for j <= 32, conj(Ya[64-j]) + i conj(Yb[64-j]) above; row 0 carries
            // (y0, y32) of both columns as (re, im)
            auto ld1 = [&](int j) -> cpx {
                if (j == 0) { const float4 q = *reinterpret_cast<const float4*>(tb); return make_float2(q.x, q.z); }
                if (j == 32) { const float4 q = *reinterpret_cast<const float4*>(tb); return make_float2(q.y, q.w); }
                if (j < 32) { const float4 q = *reinterpret_cast<const float4*>(tb + j * OS_IROW); return make_float2(q.x - q.w, q.y + q.z); }
                const float4 q = *reinterpret_cast<const float4*>(tb + (64 - j) * OS_IROW);
                return make_float2(q.x + q.w, q.z - q.y);
            };
            auto ld = [&](int j) { const cpx z0 = ld1(j), z1 = ld1(j + 1); return make_float4(z0.x, z0.y, z1.x, z1.y); };
            os_fft64_task_rt<4, true>(r0, ld, re, im);
        }
        __syncthreads();                                                   // every read of buf is done: reuse it as ostage
        if (active) {
            float* oa = ostage + ((size_t)gq * a.Sw + (xa - a.ox0)) * shp;
            float* ob = oa + shp;
            const bool wa = xa >= a.ox0;
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) {
                const int y = 4 * j1 + r0 - a.oy0;
                if (y >= 0) {
                    if (wa) oa[y] = re[j1] * a.scale;
                    ob[y] = im[j1] * a.scale;
                }
            }
        }
    }
    __syncthreads();
    // ---- store the valid block of every tile: one warp per column, lanes along h (contiguous in the plane)
    float* out = a.outs[t];
#pragma unroll 1
    for (int gq = 0; gq < OS_IG; ++gq) {
        const int Y0 = tile_y0[gq], X0 = tile_x0[gq];
        const int ny = min(a.Sh, a.crop_h - Y0);
        if (ny <= 0) continue;
        for (int x = warp; x < a.Sw && X0 + x < a.crop_w; x += 16) {
            const float* src = ostage + ((size_t)gq * a.Sw + x) * shp;
            float* dst = out + (size_t)(X0 + x) * a.out_ld + Y0;
            for (int y = lane; y < ny; y += 32) dst[y] = src[y];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// inv_w_pass: inverse complex FFT along w of the compat spectrum S [plane][FW][CH] -> Z [plane][FW][CH]
// (first half of spectrum -> plane; inv_h_pass finishes).  grid = (ceil(CH/TU), planes).
__global__ void inv_w_pass(const cpx* __restrict__ S, int FW, int CH, LinePlan plan, const cpx* __restrict__ tw,
                           cpx* __restrict__ Z, int TU, int ld)
{
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
