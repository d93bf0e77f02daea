// Large-plane pipeline (path 4): planes whose lines are thousands of points long (BASELINE config 3:
// 4096 x 4096 image, 512 x 512 templates, plane 4608 x 4608).
//
// The w pass walks the STRIDED dimension of the reference layout (h contiguous, src/cudaConvFFTData.cuh:26-27):
// to touch HBM in whole 32-byte sectors a CTA has to own 4 neighbouring h-bins of EVERY column, i.e. 4 lines of FW
// complex points.  A Stockham ping-pong (kernels_generic.cuh) needs two buffers per line and drops to 1 line per CTA
// at FW = 4608 (8-byte accesses, a quarter of each sector used).  Here the line transforms run IN PLACE:
//
//   forward  = decimation in frequency, natural order in  -> mixed-radix digit-reversed order out
//   inverse  = the same stages backwards (decimation in time), digit-reversed in -> natural order out
//
// so the pointwise product with the data spectrum happens in digit-reversed order (one 16-bit table lookup per bin
// gives the natural bin) and no reordering pass ever runs.  One buffer per line: 4 lines of 4608 points = 153 KB.
//
//   bp_repad_spec   compat spectrum [F][FW][CH] -> private [F][FW][CHp] (CHp = CH rounded up to 4: aligned sectors),
//                   pre-multiplied by 1/(FH*FW)   (elementwiseProductAndNormalize's scale, src/cudaConvFFTData.cuh:47-67)
//   bp_kern_h       template columns -> half spectrum along h; zero pad fused into the load
//                   (padData src/cudaConvFFTData.cuh:11-31), first stage pruned to the kh non-zero rows
//   bp_conv_w       per template and 4 h-bins: pruned forward w transform, product with the data spectrum, channel sum
//                   in the frequency domain, ONE inverse w transform (replaces F x cufftExecC2R + sumAlongFeatures,
//                   src/cudaConvFFTData.cu:262-271)
//   bp_inv_h        C2R along h (two real columns per complex line), crop fused into the coalesced store
#pragma once
#include "cplx.cuh"
#include "kernels_generic.cuh"

namespace fftconv {

#define BP_MAX_STAGES 8
struct IpPlan {
    int n, ns;
    int R[BP_MAX_STAGES];    // radix of stage s
    int L[BP_MAX_STAGES];    // sub-transform length entering stage s (L[0] = n, L[s+1] = L[s] / R[s])
    unsigned magic;          // ceil(2^32 / m0), m0 = n / R[0]  (0: no third pad term)
    int tws;                 // 1: later stages read their twiddles from the shared-memory table
};

// Padded position of point i of a line.  One pad slot per 16 points keeps the short-stride stages off a single bank;
// one more per m0 = n/R[0] points does the same for the stride of the first stage, which is also the stride between
// CONSECUTIVE natural indices after digit reversal (the scatter of bp_inv_h, the gather of bp_kern_h).
__device__ __forceinline__ int bp_pidx(int i, unsigned magic) { return i + (i >> 4) + (int)__umulhi((unsigned)i, magic); }

// ------------------------------------------------------------------------------- radix 32
template <> struct Dft<32> {
    __device__ __forceinline__ static void run(float* re, float* im) {
        float er[16], ei[16], qr[16], qi[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { er[k] = re[2 * k]; ei[k] = im[2 * k]; qr[k] = re[2 * k + 1]; qi[k] = im[2 * k + 1]; }
        Dft<16>::run(er, ei);
        Dft<16>::run(qr, qi);
        // w32^k = cos(pi k / 16) - i sin(pi k / 16)
        constexpr float c[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                                 0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                                 0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                                 -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                                 -0.92387953251128675613f, -0.98078528040323044913f};
        constexpr float s[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                                 0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                                 0.98078528040323044913f, 1.f, 0.98078528040323044913f, 0.92387953251128675613f,
                                 0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                                 0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float tr = fmaf(qi[k], s[k], qr[k] * c[k]);
            const float ti = fmaf(-qr[k], s[k], qi[k] * c[k]);
            re[k] = er[k] + tr; im[k] = ei[k] + ti;
            re[k + 16] = er[k] - tr; im[k + 16] = ei[k] - ti;
        }
    }
};

// ------------------------------------------------------------------------------- stage twiddles
// Stage 0 reads the global table tw[t] = w_n^t (t = j*q, coalesced enough, L1-resident).  Every later stage has a
// sub-length L_s dividing L1 = n / R[0], so all of them share ONE small table  w_L1^t, t < L1  (4 KB at n = 4608) kept in
// shared memory: those stages gather twiddles with large strides, which from global memory cost one L1 line per lane
// and were the main long-scoreboard stall of the first version.
struct TwSrc {
    const cpx* g;        // global table w_n^t
    unsigned sh;         // shared-memory byte address of the w_L1^t table (0: none)
    int gstep;           // index scale for the global table
    int sstep;           // index scale for the shared table
};
__device__ __forceinline__ cpx tw_fetch(const TwSrc& t, int jq) {
    if (t.sh) {
        cpx v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(t.sh + (unsigned)(jq * t.sstep) * 8u));
        return v;
    }
    return __ldg(&t.g[jq * t.gstep]);
}
#define BP_TW_SMEM_MAX 1024
// w_L1^t = w_n^(t * R0) for t < L1, filled by the whole CTA (followed by the caller's __syncthreads)
__device__ __forceinline__ void bp_fill_tw(cpx* tab, const IpPlan& P, const cpx* __restrict__ tw) {
    if (P.ns < 2 || P.L[1] > BP_TW_SMEM_MAX || P.tws == 0) return;
    for (int t = threadIdx.x; t < P.L[1]; t += blockDim.x) tab[t] = __ldg(&tw[t * P.R[0]]);
}
__device__ __forceinline__ unsigned bp_tw_addr(const cpx* tab, const IpPlan& P) {
    return (P.ns < 2 || P.L[1] > BP_TW_SMEM_MAX || P.tws == 0) ? 0u : (unsigned)__cvta_generic_to_shared(tab);
}

// ------------------------------------------------------------------------------- in-place stages
// One work item = one radix-R butterfly of one line.  DIF (forward): DFT_R over the R samples  base + r*m, then the
// output q is rotated by w_L^(j q).  DIT (inverse): the exact inverse — rotate input q by conj(w_L^(j q)), then the
// inverse DFT_R.  Twiddles for R = 16 / 32 come from a two-level split q = 4a + b (R/4 + 2 table reads instead of R - 1).
// LIN: the padded positions of the R samples are  A + r*S (+ r>>4 for a 32-point last stage): m is a multiple of 16,
// or 1 — true for every stage of every plan except strides 2, 4, 8.
template <int R, bool INV, bool LIN>
__device__ __forceinline__ void ip_bfly(cpx* __restrict__ ln, int base, int m, int js, bool rot, bool first,
                                        unsigned magic, const TwSrc& tw) {
    constexpr int G = (R >= 16 && R % 4 == 0) ? 4 : 1;
    cpx w1[G], wg[R / G];
    if (rot) {
#pragma unroll
        for (int q = 1; q < G; ++q) w1[q] = twd<INV>(tw_fetch(tw, js * q));
#pragma unroll
        for (int a = 1; a < R / G; ++a) wg[a] = twd<INV>(tw_fetch(tw, js * G * a));
    }
    const int A = bp_pidx(base, magic);
    const int S = m == 1 ? 1 : m + (m >> 4) + ((first && magic) ? 1 : 0);
    float re[R], im[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int o = LIN ? A + r * S + ((R == 32) ? ((m == 1) ? (r >> 4) : 0) : 0) : bp_pidx(base + r * m, magic);
        cpx v = ln[o];
        if (INV && rot && r > 0) {
            const int a = r / G, q = r % G;
            v = (G == 1 || q == 0) ? cmul(v, wg[a]) : (a == 0 ? cmul(v, w1[q]) : cmul(v, cmul(wg[a], w1[q])));
        }
        re[r] = v.x; im[r] = v.y;
    }
    dft_regs<R, INV>(re, im);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int o = LIN ? A + r * S + ((R == 32) ? ((m == 1) ? (r >> 4) : 0) : 0) : bp_pidx(base + r * m, magic);
        cpx v = make_float2(re[r], im[r]);
        if (!INV && rot && r > 0) {
            const int a = r / G, q = r % G;
            v = (G == 1 || q == 0) ? cmul(v, wg[a]) : (a == 0 ? cmul(v, w1[q]) : cmul(v, cmul(wg[a], w1[q])));
        }
        ln[o] = v;
    }
}

// __noinline__: every radix gets its own register allocation (a 32-point butterfly needs ~100 registers, a 9-point one
// 40); inlined into one kernel body the allocator spilled kilobytes per thread.
template <int R, bool INV>
__device__ __noinline__ void ip_stage(cpx* __restrict__ lines, int nl, int ldl, int n, int L, unsigned magic,
                                      const cpx* __restrict__ twg, unsigned twsh, int L1) {
    const int m = L / R, nb = n / R;
    TwSrc tw;
    tw.g = twg; tw.gstep = n / L;
    tw.sh = L == n ? 0u : twsh;                       // stage 0 walks the global table
    tw.sstep = L == n ? 0 : L1 / L;
    const int items = nb * nl;
    const bool lin = (m & 15) == 0 || m == 1;
    const bool rot = m > 1;                           // last stage: j = 0, every rotation is 1
    const bool first = L == n;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int l = it / nb, b = it - l * nb;
        const int blk = b / m, j = b - blk * m;
        cpx* ln = lines + l * ldl;
        const int base = blk * L + j;
        // (a 32-point stage is always followed by 16- or 32-point stages: its stride is a multiple of 16, or 1)
        if (lin || R == 32) ip_bfly<R, INV, true>(ln, base, m, j, rot, first, magic, tw);
        else ip_bfly<R, INV, false>(ln, base, m, j, rot, first, magic, tw);
    }
}

template <bool INV>
__device__ __forceinline__ void ip_run_stage(int R, cpx* lines, int nl, int ldl, int n, int L, unsigned magic,
                                             const cpx* __restrict__ tw, unsigned twsh, int L1) {
    switch (R) {
        case 2:  ip_stage<2, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 3:  ip_stage<3, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 4:  ip_stage<4, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 5:  ip_stage<5, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 7:  ip_stage<7, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 8:  ip_stage<8, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 9:  ip_stage<9, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 11: ip_stage<11, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 13: ip_stage<13, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 16: ip_stage<16, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        case 17: ip_stage<17, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
        default: ip_stage<32, INV>(lines, nl, ldl, n, L, magic, tw, twsh, L1); break;
    }
}

// forward first stage when only the first nz <= n/R samples of a line are non-zero (a zero-padded template):
// the DFT_R collapses to a broadcast,  Y_q[j] = x[j] * w_n^(j q).
__device__ __noinline__ void ip_stage_pruned_fwd(cpx* __restrict__ lines, int nl, int ldl, int n, int R, unsigned magic,
                                                    const cpx* __restrict__ tw) {
    const int m = n / R;
    const int S = m + (m >> 4) + (magic ? 1 : 0);     // m is a multiple of 16 (the last radix alone is >= 4 ... see make_ip_plan)
    const bool lin = (m & 15) == 0;
    for (int l = 0; l < nl; ++l) {
        cpx* ln = lines + l * ldl;
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const int A = bp_pidx(j, magic);
            const cpx x = ln[A];
            for (int q = 1; q < R; ++q)
                ln[lin ? A + q * S : bp_pidx(j + q * m, magic)] = cmul(x, __ldg(&tw[j * q]));
        }
    }
}

// natural order in (first nz samples non-zero, the caller zero-filled up to bp_fill_to) -> digit-reversed out
__device__ __forceinline__ void ip_forward(cpx* lines, int nl, int ldl, const IpPlan& P, const cpx* __restrict__ tw,
                                           unsigned twsh, int nz) {
    int s = 0;
    if (nz <= P.L[0] / P.R[0]) {
        ip_stage_pruned_fwd(lines, nl, ldl, P.n, P.R[0], P.magic, tw);
        __syncthreads();
        s = 1;
    }
    for (; s < P.ns; ++s) {
        ip_run_stage<false>(P.R[s], lines, nl, ldl, P.n, P.L[s], P.magic, tw, twsh, P.L[1]);
        __syncthreads();
    }
}
__device__ __forceinline__ int bp_fill_to(const IpPlan& P, int nz) { return nz <= P.L[0] / P.R[0] ? P.L[0] / P.R[0] : P.n; }

// digit-reversed in -> natural order out (unnormalised)
__device__ __forceinline__ void ip_inverse(cpx* lines, int nl, int ldl, const IpPlan& P, const cpx* __restrict__ tw,
                                           unsigned twsh) {
    for (int s = P.ns - 1; s >= 0; --s) {
        ip_run_stage<true>(P.R[s], lines, nl, ldl, P.n, P.L[s], P.magic, tw, twsh, P.L[1]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------- bp_repad_spec
// Row p of the private copy holds the w-bin nat_of[p]: the product pass of bp_conv_w then walks it sequentially.
__global__ void bp_repad_spec(const cpx* __restrict__ S, int CH, int CHp, int FW, long long rows, float scale,
                              const unsigned short* __restrict__ nat_of, cpx* __restrict__ Sp) {
    const long long total = rows * CHp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / CHp;
        const int u = (int)(i - r * CHp);
        const long long f = r / FW;
        const int p = (int)(r - f * FW);
        cpx v = make_float2(0.f, 0.f);
        if (u < CH) { v = S[(f * FW + nat_of[p]) * CH + u]; v.x *= scale; v.y *= scale; }
        Sp[i] = v;
    }
}

// ------------------------------------------------------------------------------- bp_kern_h
// grid (maxcols4 / 4, nk * F) x 256; 4 template columns per CTA as 2 packed complex lines.
// T: [nk * F][maxcols4][CHp]
template <int MINB>
__global__ void __launch_bounds__(256, MINB) bp_kern_h(const SrcDesc* __restrict__ srcs, int F, int maxcols4, int FH, int CH, int CHp,
                                                       const __grid_constant__ IpPlan plan, const cpx* __restrict__ tw,
                                                       const unsigned short* __restrict__ pos_of, cpx* __restrict__ T, int ldl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* lines = reinterpret_cast<cpx*>(smem_raw);
    const unsigned magic = plan.magic;
    cpx* twtab = lines + 2 * ldl;
    bp_fill_tw(twtab, plan, tw);
    const unsigned twsh = bp_tw_addr(twtab, plan);
    const int pf = blockIdx.y, s = pf / F, f = pf - s * F;
    const SrcDesc d = srcs[s];
    const int x0 = blockIdx.x * 4;
    if (x0 >= d.cols) return;
    const int rows = min(d.rows, FH);
    const float* base = d.ptr + (size_t)f * d.rows * d.cols;
    const int fill = bp_fill_to(plan, rows);
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int xa = x0 + 2 * l, xb = xa + 1;
        const float* ca = base + (size_t)xa * d.rows;
        const float* cb = base + (size_t)xb * d.rows;
        for (int y = threadIdx.x; y < fill; y += 256) {
            cpx v = make_float2(0.f, 0.f);
            if (y < rows) {
                if (xa < d.cols) v.x = ca[y];
                if (xb < d.cols) v.y = cb[y];
            }
            lines[l * ldl + bp_pidx(y, magic)] = v;
        }
    }
    __syncthreads();
    ip_forward(lines, 2, ldl, plan, tw, twsh, rows);
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const cpx* ln = lines + l * ldl;
        cpx* o = T + ((size_t)pf * maxcols4 + x0 + 2 * l) * CHp;
        for (int u = threadIdx.x; u < CHp; u += 256) {
            cpx a = make_float2(0.f, 0.f), b = a;
            if (u < CH) {
                const cpx zu = ln[bp_pidx(pos_of[u], magic)];
                const cpx zn = cconj(ln[bp_pidx(pos_of[u == 0 ? 0 : FH - u], magic)]);
                a = make_float2(0.5f * (zu.x + zn.x), 0.5f * (zu.y + zn.y));
                const cpx dd = make_float2(0.5f * (zu.x - zn.x), 0.5f * (zu.y - zn.y));
                b = make_float2(dd.y, -dd.x);                                   // -i * dd
            }
            o[u] = a;
            o[CHp + u] = b;
        }
    }
}

// ------------------------------------------------------------------------------- bp_conv_w
// Template index fastest in the grid: the CTAs resident at any moment share a handful of h-bin tiles, so the data
// spectrum is read from HBM once per call, not once per template.
// TU = 4: grid (nk, CHp / 4), single channel: one CTA per SM owns whole 32-byte sectors (4 h-bins of every column).
// TU = 2: grid (2 nk, CHp / 4), MULTI: 2 lines + 2 accumulator lines (channel sum in the frequency domain); the two
//         halves of a sector belong to CTAs that are neighbours in launch order, so they meet in L2.
template <bool CONJ, bool MULTI, int NT, int TU, int MINB>
__global__ void __launch_bounds__(NT, MINB) bp_conv_w(const cpx* __restrict__ T, const int* __restrict__ kcols, int maxcols4,
                                                    const cpx* __restrict__ Sp, int F, int FW, int CHp,
                                                    const __grid_constant__ IpPlan plan, const cpx* __restrict__ tw,
                                                    cpx* __restrict__ Z, int ldl)
{
    constexpr int PS = NT / TU;                           // points advanced per pass of the CTA
    constexpr int MB = 6;                                 // global loads in flight per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* lines = reinterpret_cast<cpx*>(smem_raw);
    const unsigned magic = plan.magic;
    cpx* twtab = lines + (MULTI ? 2 * TU : TU) * ldl;
    bp_fill_tw(twtab, plan, tw);
    const unsigned twsh = bp_tw_addr(twtab, plan);
    const int k = TU == 4 ? blockIdx.x : blockIdx.x >> 1;
    const int u0 = TU == 4 ? blockIdx.y * 4 : blockIdx.y * 4 + (blockIdx.x & 1) * 2;
    const int ncols = min(kcols[k], FW);
    const int fill = bp_fill_to(plan, ncols);
    const int l = threadIdx.x % TU, p0 = threadIdx.x / TU;
    cpx* ln = lines + l * ldl;
    cpx* an = ln + TU * ldl;                              // MULTI only: accumulator line

    for (int f = 0; f < F; ++f) {
        const cpx* Tp = T + ((size_t)(k * F + f) * maxcols4) * CHp + u0 + l;
        for (int xb = p0; xb < fill; xb += PS * 4) {        // 4 loads in flight (unconditional, clamped: they stay in registers)
            cpx v[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) v[b] = __ldcs(&Tp[(size_t)min(xb + b * PS, ncols - 1) * CHp]);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = xb + b * PS;
                if (x < fill) ln[bp_pidx(x, magic)] = x < ncols ? v[b] : make_float2(0.f, 0.f);
            }
        }
        __syncthreads();
        ip_forward(lines, TU, ldl, plan, tw, twsh, ncols);
        const cpx* Sf = Sp + (size_t)f * FW * CHp + u0 + l;     // rows already in digit-reversed order
        for (int pb = p0; pb < FW; pb += PS * MB) {
            cpx dsp[MB];
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                dsp[b] = __ldcs(&Sf[(size_t)min(pb + b * PS, FW - 1) * CHp]);
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int p = pb + b * PS;
                if (p < FW) {
                    const int o = bp_pidx(p, magic);
                    const cpx kx = ln[o];
                    cpx pr = CONJ ? cmulc(dsp[b], kx) : cmul(dsp[b], kx);
                    if (MULTI) {
                        if (f > 0) { const cpx a = an[o]; pr.x += a.x; pr.y += a.y; }
                        an[o] = pr;
                    } else {
                        ln[o] = pr;
                    }
                }
            }
        }
        __syncthreads();
    }
    ip_inverse(MULTI ? lines + TU * ldl : lines, TU, ldl, plan, tw, twsh);
    const cpx* rn = MULTI ? an : ln;
    cpx* Zk = Z + (size_t)k * FW * CHp + u0 + l;
    for (int x = p0; x < FW; x += PS) __stcs(&Zk[(size_t)x * CHp], rn[bp_pidx(x, magic)]);
}

// ------------------------------------------------------------------------------- bp_inv_h
// grid (FW / 4, nk) x 256; Z [k][FW][CHp] (already scaled) -> 4 real columns of plane k
template <int MINB>
__global__ void __launch_bounds__(256, MINB) bp_inv_h(const cpx* __restrict__ Z, int FH, int FW, int CH, int CHp,
                                                      const __grid_constant__ IpPlan plan, const cpx* __restrict__ tw,
                                                      const unsigned short* __restrict__ pos_of,
                                                      float* const* __restrict__ outs, int crop_h, int crop_w, int out_ld, int ldl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* lines = reinterpret_cast<cpx*>(smem_raw);
    const unsigned magic = plan.magic;
    const int k = blockIdx.y, x0 = blockIdx.x * 4;
    if (x0 >= crop_w) return;
    cpx* twtab = lines + 2 * ldl;
    bp_fill_tw(twtab, plan, tw);
    const unsigned twsh = bp_tw_addr(twtab, plan);
    const cpx* Zk = Z + ((size_t)k * FW + x0) * CHp;
    const int half = FH / 2;
    constexpr int MB = 5;                                   // 2 * MB global loads in flight per thread
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        cpx* ln = lines + l * ldl;
        const cpx* Za = Zk + (size_t)(2 * l) * CHp;
        const cpx* Zb = Za + CHp;
        for (int ub = threadIdx.x; ub < CH; ub += 256 * MB) {
            cpx za[MB], zb[MB];
            int pa[MB], pb[MB];
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int u = ub + b * 256;
                const int uc = min(u, CH - 1);
                za[b] = Za[uc]; zb[b] = Zb[uc]; pa[b] = pos_of[uc]; pb[b] = pos_of[uc == 0 ? 0 : FH - uc];
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int u = ub + b * 256;
                if (u < CH) {
                    if (u == 0 || u == half) {                          // C2R ignores Im of DC / Nyquist
                        ln[bp_pidx(pa[b], magic)] = make_float2(za[b].x, zb[b].x);
                    } else {
                        ln[bp_pidx(pa[b], magic)] = make_float2(za[b].x - zb[b].y, za[b].y + zb[b].x);     // za + i zb
                        ln[bp_pidx(pb[b], magic)] = make_float2(za[b].x + zb[b].y, zb[b].x - za[b].y);     // conj(za) + i conj(zb)
                    }
                }
            }
        }
    }
    __syncthreads();
    ip_inverse(lines, 2, ldl, plan, tw, twsh);
    float* o = outs[k];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const cpx* ln = lines + l * ldl;
        const int xa = x0 + 2 * l, xb = xa + 1;
        float* oa = o + (size_t)xa * out_ld;
        float* ob = o + (size_t)xb * out_ld;
        for (int y = threadIdx.x; y < crop_h; y += 256) {
            const cpx r = ln[bp_pidx(y, magic)];
            if (xa < crop_w) oa[y] = r.x;
            if (xb < crop_w) ob[y] = r.y;
        }
    }
}

}  // namespace fftconv
