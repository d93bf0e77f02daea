// 16-point-tiled fast path for the kernel-bank convolution (planes up to 544 wide).
//
// Both plane sides are multiples of 16 by contract (computeFFTsize16,
// src/cudaConvFFTData.h:96-102):  FH = 16*mh, FW = 16*mw.  Every transform is split as
//     index = a + m*b   (frequency side)        index = 16*a' + b'   (space side)
// so that all 16-point transforms run in registers (Dft<16>) and only the m-point transforms
// go through shared memory.  Because a template has at most 16*nya x 16*nxa non-zero taps, its
// forward transform is PRUNED: the m-point stage degenerates to nya (nxa) terms.
//
//   tile16_relayout     compat spectrum [F][FW][CH] -> private tiles Dp[t][f][ub][va][18]
//                       (tile t = rows u = t + mh*ub, ub = 0..15; rows above FH/2 by symmetry)
//   tile16_kern_hpass   templates -> Ag[t][kg][f][kk][ub][XCP]   (h transform, rows of tile t)
//   tile16_conv         per (tile, group of KB templates): for each channel
//                         TMA bulk copy of the Dp / Ag slabs (3-stage mbarrier pipeline)
//                         twiddle + FFT16 along w, multiply by Dp, accumulate over channels
//                       then inverse: IFFT16 (regs) + m-point IDFT (smem) along w,
//                       IFFT16 across the 16 rows of the tile, twiddle -> Wg[k][x][t][16]
//   tile16_c2r          mh-point inverse along h (two columns per complex line), scale,
//                       crop, coalesced store.
//
// Replaces, per template: padData + cufftExecR2C + elementwiseProductAndNormalize +
// F x cufftExecC2R + sumAlongFeatures (src/cudaConvFFTData.cu:233-271).
#pragma once
#include <cstdint>
#include "line_fft.cuh"
#include "kernels_generic.cuh"

namespace fftconv {

#define T16_PAD 18            // padded row of 16 complex (144 B): conflict-free LDS.128 across lanes

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_shared(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------ relayout
__global__ void tile16_relayout(const cpx* __restrict__ S, cpx* __restrict__ Dp, int F, int FH, int FW, int CH,
                                int mh, int mw, int NT)
{
    const long long total = (long long)NT * F * 16 * mw * T16_PAD;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int vb = (int)(i % T16_PAD);
        long long r = i / T16_PAD;
        const int va = (int)(r % mw); r /= mw;
        const int ub = (int)(r % 16); r /= 16;
        const int f = (int)(r % F);
        const int t = (int)(r / F);
        cpx val = make_float2(0.f, 0.f);
        if (vb < 16) {
            const int u = t + mh * ub, v = va + mw * vb;
            if (u <= FH / 2) val = S[((size_t)f * FW + v) * CH + u];
            else val = cconj(S[((size_t)f * FW + (v == 0 ? 0 : FW - v)) * CH + (FH - u)]);
        }
        Dp[i] = val;
    }
}

// ------------------------------------------------------------------------------ kernel h pass
// one thread per (template k, channel f, tile t, column x): 16-point pruned transform along h
__global__ void tile16_kern_hpass(const SrcDesc* __restrict__ descs, int nk, int F, int FH, int mh, int NT, int nya,
                                  int XC, int KB, int NG, const cpx* __restrict__ twH, cpx* __restrict__ Ag)
{
    const int XCP = XC + 2;
    const long long total = (long long)nk * F * NT * XC;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % XC);
    long long r = i / XC;
    const int t = (int)(r % NT); r /= NT;
    const int f = (int)(r % F);
    const int k = (int)(r / F);
    const SrcDesc d = descs[k];
    const int rows = min(d.rows, FH);
    float re[16], im[16];
    if (x < d.cols) {
        const float* col = d.ptr + ((size_t)f * d.cols + x) * d.rows;
        if (nya == 1) {
#pragma unroll
            for (int yb = 0; yb < 16; ++yb) { re[yb] = yb < rows ? col[yb] : 0.f; im[yb] = 0.f; }
        } else {
#pragma unroll
            for (int yb = 0; yb < 16; ++yb) { re[yb] = 0.f; im[yb] = 0.f; }
            for (int ya = 0; ya < nya; ++ya) {
                const cpx w = twH[((t * ya) % mh) * 16];          // w_mh^(ua*ya)
#pragma unroll
                for (int yb = 0; yb < 16; ++yb) {
                    const int y = 16 * ya + yb;
                    const float v = y < rows ? col[y] : 0.f;
                    re[yb] = fmaf(w.x, v, re[yb]); im[yb] = fmaf(w.y, v, im[yb]);
                }
            }
        }
#pragma unroll
        for (int yb = 1; yb < 16; ++yb) {                         // w_FH^(ua*yb)
            const cpx w = twH[t * yb];
            const float a = re[yb], b = im[yb];
            re[yb] = fmaf(-b, w.y, a * w.x); im[yb] = fmaf(b, w.x, a * w.y);
        }
        Dft<16>::run(re, im);
    } else {
#pragma unroll
        for (int yb = 0; yb < 16; ++yb) { re[yb] = 0.f; im[yb] = 0.f; }
    }
    const int kg = k / KB, kk = k - kg * KB;
    cpx* o = Ag + ((((size_t)t * NG + kg) * F + f) * KB + kk) * 16 * XCP + x;
#pragma unroll
    for (int ub = 0; ub < 16; ++ub) o[(size_t)ub * XCP] = make_float2(re[ub], im[ub]);
}

// ------------------------------------------------------------------------------ main kernel
struct Tile16Params {
    const cpx* Dp;        // [NT][F][16][mw][18]
    const cpx* Ag;        // [NT][NG][F][KB][16][XCP]
    cpx* Wg;              // [nk][FW][NT][16]
    const cpx* twW;       // n = FW
    const cpx* twH;       // n = FH
    const cpx* twM;       // n = mw
    LinePlan planM;       // n = mw
    int F, FH, FW, mh, mw, NT, NG, KB, nk, nxa, nstage;
    int nmain;            // items (row, va) owned by a thread for the whole channel loop (<= 512)
    int nextra;           // warp-passes per channel for the remaining items, dealt round-robin to the warps
};

// Work decomposition.  An item is (row = kk*16 + ub, va): 16 accumulators (the bins va + mw*vb).
// There are N = KB*16*mw items per CTA.  The first `nmain` (<= 512 = 16 warps, 4 per SM
// sub-partition, 128 registers each) keep their accumulators in registers for all channels.
// When N > 512 (e.g. FW = 272: 2*16*17 = 544) the remaining items are processed as `nextra`
// extra warp-passes per channel, handed round-robin to the 16 warps with accumulators in shared
// memory, so that every sub-partition issues the same number of passes.

// K^[u][va + mw*vb], vb = 0..15, for one item: load the row of the template half-transform, pruned
// m-point stage, twiddle, 16-point FFT in registers.
template <bool TW_REGS>
__device__ __forceinline__ void t16_item_spectrum(const cpx* __restrict__ As_row, int nxa, const cpx* __restrict__ twM,
                                                  int va, int mw, const float* twr, const float* twi,
                                                  const cpx* __restrict__ twW, float* re, float* im)
{
    const float4* ap = reinterpret_cast<const float4*>(As_row);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 q = ap[j];
        re[2 * j] = q.x; im[2 * j] = q.y; re[2 * j + 1] = q.z; im[2 * j + 1] = q.w;
    }
    // pruned m-point stage: in[xb] = sum_xa w_mw^(va*xa) * A[16*xa + xb]
    for (int xa = 1; xa < nxa; ++xa) {
        const cpx w = twM[(va * xa) % mw];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = ap[8 * xa + j];
            re[2 * j] = fmaf(w.x, q.x, fmaf(-w.y, q.y, re[2 * j]));
            im[2 * j] = fmaf(w.x, q.y, fmaf(w.y, q.x, im[2 * j]));
            re[2 * j + 1] = fmaf(w.x, q.z, fmaf(-w.y, q.w, re[2 * j + 1]));
            im[2 * j + 1] = fmaf(w.x, q.w, fmaf(w.y, q.z, im[2 * j + 1]));
        }
    }
#pragma unroll
    for (int xb = 1; xb < 16; ++xb) {
        float wr, wi;
        if (TW_REGS) { wr = twr[xb]; wi = twi[xb]; }
        else { const cpx w = __ldg(&twW[va * xb]); wr = w.x; wi = w.y; }       // va*xb < FW
        const float a = re[xb], b = im[xb];
        re[xb] = fmaf(-b, wi, a * wr);
        im[xb] = fmaf(b, wr, a * wi);
    }
    Dft<16>::run(re, im);
}

#define T16_THREADS 512

template <bool CONJ>
__global__ void __launch_bounds__(T16_THREADS, 1) tile16_conv(const Tile16Params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mw = P.mw, KB = P.KB, F = P.F;
    const int XCP = 16 * P.nxa + 2;
    const uint32_t dp_bytes = (uint32_t)(16 * mw * T16_PAD * sizeof(cpx));
    const uint32_t a_bytes = (uint32_t)(KB * 16 * XCP * sizeof(cpx));
    const uint32_t stage_bytes = dp_bytes + a_bytes;            // both multiples of 16
    const int nstage = P.nstage;
    const int t = blockIdx.y, kg = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int N = KB * 16 * mw;                                 // == KB * FW
    const int NM = P.nmain, E = P.nextra;
    const bool active = tid < NM;
    const int item = active ? tid : 0;
    const int va = item % mw, row = item / mw;                  // row = kk*16 + ub
    const int ub = row & 15;

    // smem: [pipeline stages | (aliased later) inverse ping-pong] [mbarriers] [extra accumulators]
    const int mwp = mw | 1;
    const size_t y_bytes = 2 * (size_t)KB * 256 * mwp * sizeof(cpx);
    const size_t pipe_bytes = (size_t)nstage * stage_bytes;
    const size_t bar_off = ((pipe_bytes > y_bytes ? pipe_bytes : y_bytes) + 15) & ~(size_t)15;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + bar_off);          // [nstage] "slab landed"
    uint32_t* done = reinterpret_cast<uint32_t*>(smem_raw + bar_off + 32);     // [nstage] warps finished with the slab
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem_raw + bar_off + 48);    // [nstage] "every warp is done with the slab"
    uint32_t* seq = reinterpret_cast<uint32_t*>(smem_raw + bar_off + 80);      // [E][4] passes done per accumulator copy
    cpx* Racc = reinterpret_cast<cpx*>(smem_raw + bar_off + 192);              // [E][4][16][32]

    const cpx* dp_src = P.Dp + (size_t)t * F * (dp_bytes / sizeof(cpx));
    const cpx* a_src = P.Ag + ((size_t)t * P.NG + kg) * F * (a_bytes / sizeof(cpx));

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (int)blockDim.x); done[s] = 0; }
        for (int e = 0; e < 4 * E; ++e) seq[e] = 0;
        fence_barrier_init();
        fence_proxy_async();
    }
    for (int i = tid; i < E * 2048; i += blockDim.x) Racc[i] = make_float2(0.f, 0.f);
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < nstage && s < F; ++s) {
            unsigned char* st = smem_raw + (size_t)s * stage_bytes;
            mbar_expect_tx(&full[s], stage_bytes);
            bulk_g2s(st, dp_src + (size_t)s * (dp_bytes / sizeof(cpx)), dp_bytes, &full[s]);
            bulk_g2s(st + dp_bytes, a_src + (size_t)s * (a_bytes / sizeof(cpx)), a_bytes, &full[s]);
        }
    }

    // per-thread w twiddles  w_FW^(va*xb), xb = 0..15
    float twr[16], twi[16];
    twr[0] = 1.f; twi[0] = 0.f;
#pragma unroll
    for (int xb = 1; xb < 16; ++xb) {
        const cpx w = P.twW[va * xb];
        twr[xb] = w.x; twi[xb] = w.y;
    }
    float accr[16], acci[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { accr[j] = 0.f; acci[j] = 0.f; }

    int s = 0, xw = 0;                                          // xw: warp owning the next extra pass
    uint32_t phase = 0;
    for (int f = 0; f < F; ++f) {
        mbar_wait(&full[s], phase);
        const unsigned char* st = smem_raw + (size_t)s * stage_bytes;
        const cpx* Dps = reinterpret_cast<const cpx*>(st);
        const cpx* As = reinterpret_cast<const cpx*>(st + dp_bytes);
        {
            float re[16], im[16];
            t16_item_spectrum<true>(As + (size_t)row * XCP, P.nxa, P.twM, va, mw, twr, twi, P.twW, re, im);
#ifdef T16_FENCE_DP
            asm volatile("" ::: "memory");
#endif
            const float4* dp = reinterpret_cast<const float4*>(Dps + ((size_t)ub * mw + va) * T16_PAD);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 q = dp[j];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float dr = h ? q.z : q.x, di = h ? q.w : q.y;
                    const float kr = re[2 * j + h], ki = CONJ ? -im[2 * j + h] : im[2 * j + h];
                    accr[2 * j + h] = fmaf(kr, dr, fmaf(-ki, di, accr[2 * j + h]));
                    acci[2 * j + h] = fmaf(kr, di, fmaf(ki, dr, acci[2 * j + h]));
                }
            }
        }
        for (int e = 0; e < E; ++e) {
            if (warp == xw) {
                // Each slot has 4 accumulator copies used in turn (channel f -> copy f & 3), so that
                // consecutive passes of a slot, which run on different warps, do not serialise; the
                // passes of one copy still run in channel order (deterministic accumulation).
                const int cpy = f & 3;
                while (ld_acquire_shared(&seq[4 * e + cpy]) != (uint32_t)(f >> 2)) __nanosleep(64);
                const int id = NM + 32 * e + lane;
                if (id < N) {
                    const int erow = id / mw, eva = id - erow * mw;
                    float re[16], im[16];
                    t16_item_spectrum<false>(As + (size_t)erow * XCP, P.nxa, P.twM, eva, mw, nullptr, nullptr, P.twW, re, im);
                    const float4* dp = reinterpret_cast<const float4*>(Dps + ((size_t)(erow & 15) * mw + eva) * T16_PAD);
                    cpx* ra = Racc + ((size_t)e * 4 + cpy) * 512 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q = dp[j];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float dr = h ? q.z : q.x, di = h ? q.w : q.y;
                            const float kr = re[2 * j + h], ki = CONJ ? -im[2 * j + h] : im[2 * j + h];
                            cpx a = ra[(2 * j + h) * 32];
                            a.x = fmaf(kr, dr, fmaf(-ki, di, a.x));
                            a.y = fmaf(kr, di, fmaf(ki, dr, a.y));
                            ra[(2 * j + h) * 32] = a;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) st_release_shared(&seq[4 * e + cpy], (uint32_t)(f >> 2) + 1);
            }
            if (++xw == nwarps) xw = 0;
        }
        // No block barrier in the channel loop: the LAST warp to finish with slab s refills it, so the
        // warps drift apart and their shared-memory and FMA phases overlap.  Every thread arrives on the slab's "empty"
        // mbarrier before its warp counts itself out; the counter only elects the refilling warp, which then passes through the
        // (already complete) empty barrier: the reads of all warps are ordered before the bulk copy by the barrier itself,
        // the edge compute-sanitizer's racecheck models (round 1 reported the refill against the reads of the slab because
        // the acq_rel counter was the only ordering).
        mbar_arrive(&empty[s]);                     // every thread that read the slab arrives itself
        __syncwarp();
        if (lane == 0) {
            if (atom_add_acq_rel_shared(&done[s], 1u) == (uint32_t)nwarps - 1u) {
                done[s] = 0;
                mbar_wait(&empty[s], phase);
                if (f + nstage < F) {
                    unsigned char* dst = smem_raw + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], stage_bytes);
                    bulk_g2s(dst, dp_src + (size_t)(f + nstage) * (dp_bytes / sizeof(cpx)), dp_bytes, &full[s]);
                    bulk_g2s(dst + dp_bytes, a_src + (size_t)(f + nstage) * (a_bytes / sizeof(cpx)), a_bytes, &full[s]);
                }
            }
        }
        if (++s == nstage) { s = 0; phase ^= 1; }
    }
    __syncthreads();                                         // all slabs consumed before the buffers are reused

    // ---- inverse along w: IFFT16 over vb (registers), twiddle, m-point IDFT over va (smem)
    cpx* Y0 = reinterpret_cast<cpx*>(smem_raw);
    cpx* Y1 = Y0 + (size_t)KB * 256 * mwp;
    dft_regs<16, true>(accr, acci);
    if (active) {
        Y0[((size_t)row * 16) * mwp + va] = make_float2(accr[0], acci[0]);
#pragma unroll
        for (int xb = 1; xb < 16; ++xb) {
            // multiply by conj(w_FW^(va*xb))
            const float a = accr[xb], b = acci[xb];
            Y0[((size_t)row * 16 + xb) * mwp + va] =
                make_float2(fmaf(b, twi[xb], a * twr[xb]), fmaf(b, twr[xb], -(a * twi[xb])));
        }
    }
    for (int e = warp; e < E; e += nwarps) {
        const int id = NM + 32 * e + lane;
        if (id < N) {
            const int erow = id / mw, eva = id - erow * mw;
            const cpx* ra = Racc + (size_t)e * 2048 + lane;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const cpx a0 = ra[j * 32], a1 = ra[512 + j * 32], a2 = ra[1024 + j * 32], a3 = ra[1536 + j * 32];
                accr[j] = (a0.x + a1.x) + (a2.x + a3.x); acci[j] = (a0.y + a1.y) + (a2.y + a3.y);
            }
            dft_regs<16, true>(accr, acci);
#pragma unroll
            for (int xb = 0; xb < 16; ++xb) {
                const cpx w = __ldg(&P.twW[eva * xb]);
                const float a = accr[xb], b = acci[xb];
                Y0[((size_t)erow * 16 + xb) * mwp + eva] =
                    make_float2(fmaf(b, w.y, a * w.x), fmaf(b, w.x, -(a * w.y)));
            }
        }
    }
    __syncthreads();
    const cpx* Zs = fft_lines<true>(Y0, Y1, KB * 256, mwp, P.planM, P.twM);

    // ---- IFFT16 across the 16 rows of the tile for each column x, twiddle w_FH^(-ua*yb)
    for (int it = tid; it < N; it += blockDim.x) {
        const int kk = it / P.FW, x = it - kk * P.FW;
        const int xa = x >> 4, xb = x & 15;
        const int k = kg * KB + kk;
        if (k < P.nk) {
            float zr[16], zi[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const cpx z = Zs[((size_t)(kk * 16 + u) * 16 + xb) * mwp + xa];
                zr[u] = z.x; zi[u] = z.y;
            }
            dft_regs<16, true>(zr, zi);
            float4* o = reinterpret_cast<float4*>(P.Wg + (((size_t)k * P.FW + x) * P.NT + t) * 16);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 q;
                {
                    const cpx w = __ldg(&P.twH[t * (2 * j)]);
                    const float a = zr[2 * j], b = zi[2 * j];
                    q.x = fmaf(b, w.y, a * w.x); q.y = fmaf(b, w.x, -(a * w.y));
                }
                {
                    const cpx w = __ldg(&P.twH[t * (2 * j + 1)]);
                    const float a = zr[2 * j + 1], b = zi[2 * j + 1];
                    q.z = fmaf(b, w.y, a * w.x); q.w = fmaf(b, w.x, -(a * w.y));
                }
                o[j] = q;
            }
        }
    }
}

// ------------------------------------------------------------------------------ final C2R
// grid.x = ceil(nk*(FW/2)/NP); smem = 2*NP*16*mhp*sizeof(cpx)
__global__ void tile16_c2r(const cpx* __restrict__ Wg, int nk, int FH, int FW, int mh, int NT,
                           LinePlan planH, const cpx* __restrict__ twMh, float scale,
                           float* const* __restrict__ outs, int crop_h, int crop_w, int out_ld, int NP, int mhp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* b0 = reinterpret_cast<cpx*>(smem_raw);
    cpx* b1 = b0 + (size_t)NP * 16 * mhp;
    const int npairs = FW / 2;
    const long long nlines = (long long)nk * npairs;
    const long long line0 = (long long)blockIdx.x * NP;
    const int tlen = NT * 16;
    const int mirror_max = (mh + 1) / 2 - 1;                 // tiles 1..mirror_max also fill mh - t

    for (int idx = threadIdx.x; idx < NP * tlen; idx += blockDim.x) {
        const int p = idx / tlen, r = idx - p * tlen;
        const int t = r >> 4, yb = r & 15;
        const long long line = line0 + p;
        cpx wa = make_float2(0.f, 0.f), wb = wa;
        if (line < nlines) {
            const int cp = (int)(line % npairs);
            const size_t k = (size_t)(line / npairs);
            const cpx* base = Wg + ((k * FW + 2 * cp) * NT) * 16;
            wa = base[r];
            wb = base[(size_t)tlen + r];
        }
        cpx* L = b0 + ((size_t)p * 16 + yb) * mhp;
        L[t] = make_float2(wa.x - wb.y, wa.y + wb.x);                    // wa + i*wb
        if (t >= 1 && t <= mirror_max)
            L[mh - t] = make_float2(wa.x + wb.y, wb.x - wa.y);           // conj(wa) + i*conj(wb)
    }
    __syncthreads();
    const cpx* res = fft_lines<true>(b0, b1, NP * 16, mhp, planH, twMh);
    for (int idx = threadIdx.x; idx < NP * crop_h; idx += blockDim.x) {
        const int p = idx / crop_h, y = idx - p * crop_h;
        const long long line = line0 + p;
        if (line >= nlines) continue;
        const int cp = (int)(line % npairs);
        const size_t k = (size_t)(line / npairs);
        const cpx r = res[((size_t)p * 16 + (y & 15)) * mhp + (y >> 4)];
        float* o = outs[k];
        const int xa = 2 * cp, xb = xa + 1;
        if (xa < crop_w) o[(size_t)xa * out_ld + y] = r.x * scale;
        if (xb < crop_w) o[(size_t)xb * out_ld + y] = r.y * scale;
    }
}

}  // namespace fftconv
