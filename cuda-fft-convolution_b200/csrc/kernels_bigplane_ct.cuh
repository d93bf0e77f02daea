// Large-plane pipeline (path 4), size-specialised kernels: the same in-place transforms as kernels_bigplane.cuh with the
// whole radix plan known at COMPILE time.
//
// ncu of the run-time-plan kernels at BASELINE config 3 (profiles/r01e_bigplane_ncu.md) showed an issue-bound pipeline in
// which integer work (run-time strides, two integer divisions per butterfly, 64-bit address arithmetic behind the
// __noinline__ stage functions) is ~80 % of the floating-point work, plus 16.7 M local-memory loads of spilled butterfly
// state.  With the plan as a template parameter every shared-memory offset of a butterfly is an immediate of its LDS/STS,
// the index arithmetic is done once per butterfly in shifts, nothing is called, nothing spills, and neighbouring passes fuse:
//
//   bp_conv_w_ct   per template and 4 h-bins (one 32-byte sector of every column):
//                    pass A  template columns from HBM -> pruned first forward stage (a broadcast x * w^(jq)) -> smem
//                    pass B  middle forward stages
//                    pass C  last forward stage + product with the data spectrum (+ channel accumulation) + first inverse
//                            stage, all on the R points a thread already holds (neither stage has twiddles: stride 1)
//                    pass D  middle inverse stages
//                    pass E  last inverse stage -> Z in HBM straight from registers (natural order, whole sectors)
//                  5 shared-memory passes and 4 barriers instead of 8 and 8 (padData + cufftExecR2C + elementwiseProduct +
//                  cufftExecC2R along w, src/cudaConvFFTData.cuh:11-67, src/cudaConvFFTData.cu:233-271)
//   bp_inv_h_ct    C2R along h (two real columns per complex line), crop fused into the coalesced store
//
// A size without an instantiated plan runs on the run-time-plan kernels (kernels_bigplane.cuh); both use the same digit
// order (the host builds IpPlan from the same radices, see bp_ct_radices in fftconv.cu), so bp_kern_h / bp_repad_spec and
// the pos_of / nat_of tables are shared.
#pragma once
#include "kernels_bigplane.cuh"

namespace fftconv {

// ------------------------------------------------------------------------------- compile-time plan
template <int N_, int... Rs> struct CtPlan {
    static constexpr int N = N_;
    static constexpr int NS = sizeof...(Rs);
    static constexpr int Rarr[NS] = {Rs...};
    static constexpr int R(int s) { return Rarr[s]; }
    static constexpr int L(int s) { int l = N_; for (int i = 0; i < s; ++i) l /= Rarr[i]; return l; }   // sub-length entering stage s
    static constexpr int M(int s) { return L(s) / Rarr[s]; }                                            // butterfly stride of stage s
    static constexpr int M0 = N_ / Rarr[0];
    static_assert(L(NS - 1) == Rarr[NS - 1], "radices must multiply to N");
    static_assert(M0 % 16 == 0 && M0 <= BP_TW_SMEM_MAX, "first-stage stride: multiple of 16, later twiddles fit the smem table");
    // padded position of point i: one pad slot per 16 points, one more per M0 points (see bp_pidx)
    __host__ __device__ static constexpr int pidx(int i) { return i + (i >> 4) + i / M0; }
    // line stride (cpx) for NL lines walked line-fastest by the lanes: == 4 (mod 16) for 4 lines, == 8 (mod 16) for 2 --
    // a half-warp (NL lines x 16/NL neighbouring butterflies, 64-bit accesses) then covers all 32 banks exactly once
    template <int NL> static constexpr int ld() { return ((pidx(N_) + 15) / 16) * 16 + (NL >= 4 ? 4 : NL == 2 ? 8 : 0); }
    static constexpr int TWN = M0 + M0 / 16;                       // padded shared twiddle table (see ct_fill_tw)
};

// padded offset of sample r of a stage-S butterfly relative to its sample 0 (independent of the butterfly, see the
// derivation in DESIGN 3b: L % 16 == 0, or the whole butterfly lives inside one 16-point group)
template <class P, int S>
__host__ __device__ constexpr int ct_off(int r) {
    constexpr int m = P::M(S), L = P::L(S);
    static_assert(L % 16 == 0 || 16 % L == 0, "stage geometry");
    return r * m + (L % 16 == 0 ? ((r * m) >> 4) : 0) + (S == 0 ? r : 0);
}

// position of natural index v in the digit-reversed order the forward stages leave (the pos_of table, in arithmetic)
template <class P, int S = 0>
__device__ __forceinline__ int ct_pos(int v) {
    if constexpr (S == P::NS - 1) return v;
    else { const int q = v % P::R(S); return q * P::M(S) + ct_pos<P, S + 1>(v / P::R(S)); }
}

// rotations w_L^(j q), q = 1..R-1, as R/4 + 2 table reads (two-level split q = 4a + b) for R >= 16
template <int R, bool INV, class Fetch>
struct CtTw {
    static constexpr int G = (R >= 16 && R % 4 == 0) ? 4 : 1;
    cpx w1[G], wg[R / G];
    __device__ __forceinline__ void load(const Fetch& fetch) {
#pragma unroll
        for (int q = 1; q < G; ++q) w1[q] = twd<INV>(fetch(q));
#pragma unroll
        for (int a = 1; a < R / G; ++a) wg[a] = twd<INV>(fetch(G * a));
    }
    __device__ __forceinline__ cpx apply(cpx v, int r) const {
        if (r == 0) return v;
        const int a = r / G, q = r % G;
        return (G == 1 || q == 0) ? cmul(v, wg[a]) : (a == 0 ? cmul(v, w1[q]) : cmul(v, cmul(wg[a], w1[q])));
    }
};

// Rotations of the stride-M0 stage, w_N^(j q), q = 1..R-1, from TWO reads of the global table (w^j and w^4j) and products
// of depth <= 2 (rounding error <= 3 ulp of the table entries): this stage used to issue R - 1 scattered loads per butterfly,
// the top long-scoreboard stall of the pass.  R <= 9.
template <int R, bool INV>
struct CtTw0 {
    cpx w[R];
    __device__ __forceinline__ void load(const cpx* __restrict__ tw, int j) {
        static_assert(R <= 9, "power chain is laid out for R <= 9");
        w[1] = twd<INV>(__ldg(&tw[j]));
        if (R > 2) w[2] = cmul(w[1], w[1]);
        if (R > 3) w[3] = cmul(w[2], w[1]);
        if (R > 4) w[4] = twd<INV>(__ldg(&tw[4 * j]));
        if (R > 5) w[5] = cmul(w[4], w[1]);
        if (R > 6) w[6] = cmul(w[4], w[2]);
        if (R > 7) w[7] = cmul(w[4], w[3]);
        if (R > 8) w[8] = cmul(w[4], w[4]);
    }
    __device__ __forceinline__ cpx apply(cpx v, int r) const { return r == 0 ? v : cmul(v, w[r]); }
};

// One in-place stage over NL lines.  Item = (butterfly b, line l) with the LINE fastest: the lanes of a warp hold the
// same butterfly of 4 neighbouring lines x 8 neighbouring butterflies (conflict-free 64-bit accesses with LD == 4 mod 16).
//   forward (DIF):  DFT_R, then output q rotated by w_L^(j q)        inverse (DIT): the exact inverse
template <class P, int S, bool INV, int NL, int NT, bool CHAIN = false>
__device__ __forceinline__ void ct_stage(cpx* __restrict__ lines, const cpx* __restrict__ twg, const cpx* __restrict__ twsh) {
    constexpr int R = P::R(S), L = P::L(S), m = P::M(S), nb = P::N / R, LDL = P::template ld<NL>();
    constexpr bool rot = m > 1;
    for (int it = threadIdx.x; it < nb * NL; it += NT) {
        const int l = it % NL, b = it / NL;
        const int blk = b / m, j = b - blk * m;
        cpx* p = lines + l * LDL + P::pidx(blk * L + j);
        auto fetch = [&](int q) -> cpx {
            if constexpr (S == 0) return __ldg(&twg[j * q]);             // w_N^(j q): global table (L1-resident)
            else { const int t = j * q * (P::M0 / L); return twsh[t + (t >> 4)]; }   // w_L^(j q) = w_M0^(j q M0/L): shared table
        };
        std::conditional_t<(CHAIN && S == 0 && R <= 9), CtTw0<R, INV>, CtTw<R, INV, decltype(fetch)>> tw;
        if constexpr (CHAIN && S == 0 && R <= 9) tw.load(twg, j);
        else if (rot) tw.load(fetch);
        float re[R], im[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cpx v = p[ct_off<P, S>(r)];
            if (INV && rot) v = tw.apply(v, r);
            re[r] = v.x; im[r] = v.y;
        }
        dft_regs<R, INV>(re, im);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cpx v = make_float2(re[r], im[r]);
            if (!INV && rot) v = tw.apply(v, r);
            p[ct_off<P, S>(r)] = v;
        }
    }
}

template <class P, bool INV, int NL, int NT, int S0, int S1>
__device__ __forceinline__ void ct_stages(cpx* lines, const cpx* __restrict__ twg, const cpx* __restrict__ twsh) {
    // forward: S0, S0+1, ..., S1-1; inverse: S1-1, ..., S0.  Every stage is followed by a barrier.
    if constexpr (S0 < S1) {
        if constexpr (!INV) {
            ct_stage<P, S0, false, NL, NT>(lines, twg, twsh);
            __syncthreads();
            ct_stages<P, false, NL, NT, S0 + 1, S1>(lines, twg, twsh);
        } else {
            ct_stage<P, S1 - 1, true, NL, NT>(lines, twg, twsh);
            __syncthreads();
            ct_stages<P, true, NL, NT, S0, S1 - 1>(lines, twg, twsh);
        }
    }
}

// w_M0^t = w_N^(t R0), t < M0, one pad slot per 16 entries: a stage reads entries j*q*c for 8 neighbouring j, and with
// q*c a multiple of 16 the unpadded table put all of them on one bank
template <class P, int NT>
__device__ __forceinline__ void ct_fill_tw(cpx* tab, const cpx* __restrict__ tw) {
    for (int t = threadIdx.x; t < P::M0; t += NT) tab[t + (t >> 4)] = __ldg(&tw[t * P::R(0)]);
}

// ------------------------------------------------------------------------------- bp_conv_w_ct
// grid (TU == 4: nk, CHp / 4 | TU == 2: 2 nk, CHp / 4), NT threads, one CTA per SM.  Same tile ownership as bp_conv_w.
// FLAGS: 1 = first-stage rotations by power chain (CtTw0), 2 = all global loads of a pass issued before its first use
template <class P, bool CONJ, bool MULTI, int NT, int TU, int MINB, int FLAGS>
__global__ void __launch_bounds__(NT, MINB) bp_conv_w_ct(const cpx* __restrict__ T, const int* __restrict__ kcols, int maxcols4,
                                                      const cpx* __restrict__ Sp, int F, int CHp,
                                                      const cpx* __restrict__ tw, cpx* __restrict__ Z)
{
    constexpr int NS = P::NS, FW = P::N, M0 = P::M0, R0 = P::R(0), RL = P::R(NS - 1);
    static_assert(P::M(NS - 1) == 1, "last stage has stride 1");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* lines = reinterpret_cast<cpx*>(smem_raw);
    constexpr int LDL = P::template ld<TU>();
    cpx* acc = lines + TU * LDL;                           // MULTI only: channel accumulator lines
    cpx* twtab = lines + (MULTI ? 2 * TU : TU) * LDL;
    ct_fill_tw<P, NT>(twtab, tw);
    const int k = TU == 4 ? blockIdx.x : blockIdx.x >> 1;
    const int u0 = TU == 4 ? blockIdx.y * 4 : blockIdx.y * 4 + (blockIdx.x & 1) * 2;
    const int ncols = min(kcols[k], FW);
    const int l = threadIdx.x % TU, p0 = threadIdx.x / TU;
    constexpr int PS = NT / TU;
    static_assert(NT % TU == 0, "threads per line");
    cpx* ln = lines + l * LDL;

    for (int f = 0; f < F; ++f) {
        const cpx* Tp = T + ((size_t)(k * F + f) * maxcols4) * CHp + u0 + l;
        if (ncols <= M0) {
            // pass A: only the first M0 samples can be non-zero -> DFT_R0 of (x, 0, ..., 0) = broadcast, rotated by w_N^(j q)
            constexpr int NA = (FLAGS & 2) ? (M0 + PS - 1) / PS : 2;
            for (int jb = p0; jb < M0; jb += NA * PS) {
                cpx x[NA];
#pragma unroll
                for (int b = 0; b < NA; ++b) x[b] = __ldcs(&Tp[(size_t)min(jb + b * PS, ncols - 1) * CHp]);
#pragma unroll
                for (int b = 0; b < NA; ++b) {
                    const int j = jb + b * PS;
                    if (j < M0) {
                        const cpx v = j < ncols ? x[b] : make_float2(0.f, 0.f);
                        cpx* p = ln + P::pidx(j);
                        p[0] = v;
                        if constexpr ((FLAGS & 1) && R0 <= 9) {
                            CtTw0<R0, false> t;
                            t.load(tw, j);
#pragma unroll
                            for (int q = 1; q < R0; ++q) p[ct_off<P, 0>(q)] = t.apply(v, q);
                        } else {
#pragma unroll
                            for (int q = 1; q < R0; ++q) p[ct_off<P, 0>(q)] = cmul(v, __ldg(&tw[j * q]));
                        }
                    }
                }
            }
            __syncthreads();
        } else {
            for (int x = p0; x < FW; x += PS) ln[P::pidx(x)] = x < ncols ? __ldcs(&Tp[(size_t)x * CHp]) : make_float2(0.f, 0.f);
            __syncthreads();
            ct_stage<P, 0, false, TU, NT, (FLAGS & 1) != 0>(lines, tw, twtab);
            __syncthreads();
        }
        // pass B
        ct_stages<P, false, TU, NT, 1, NS - 1>(lines, tw, twtab);
        // pass C: forward DFT_RL, product with the data spectrum (rows already in digit-reversed order), channel sum,
        // and -- on the last channel -- the inverse DFT_RL
        const cpx* Sf = Sp + (size_t)f * FW * CHp + u0 + l;
        constexpr int NC = (FLAGS & 2) ? (FW / RL + PS - 1) / PS : 1;
        for (int bb = p0; bb < FW / RL; bb += NC * PS) {
            cpx d[NC][RL];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int b = min(bb + i * PS, FW / RL - 1);
#pragma unroll
                for (int r = 0; r < RL; ++r) d[i][r] = __ldcs(&Sf[(size_t)(b * RL + r) * CHp]);
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int b = bb + i * PS;
                if (b < FW / RL) {
                    cpx* p = ln + P::pidx(b * RL);
                    cpx* pa = p + TU * LDL;
                    float re[RL], im[RL];
#pragma unroll
                    for (int r = 0; r < RL; ++r) { const cpx v = p[ct_off<P, NS - 1>(r)]; re[r] = v.x; im[r] = v.y; }
                    dft_regs<RL, false>(re, im);
#pragma unroll
                    for (int r = 0; r < RL; ++r) {
                        const cpx kx = make_float2(re[r], im[r]);
                        cpx pr = CONJ ? cmulc(d[i][r], kx) : cmul(d[i][r], kx);
                        if (MULTI && f > 0) { const cpx a = pa[ct_off<P, NS - 1>(r)]; pr.x += a.x; pr.y += a.y; }
                        re[r] = pr.x; im[r] = pr.y;
                    }
                    if (!MULTI || f == F - 1) dft_regs<RL, true>(re, im);
                    cpx* po = MULTI ? pa : p;
#pragma unroll
                    for (int r = 0; r < RL; ++r) po[ct_off<P, NS - 1>(r)] = make_float2(re[r], im[r]);
                }
            }
        }
        __syncthreads();
    }
    cpx* res = MULTI ? acc : lines;
    // pass D
    ct_stages<P, true, TU, NT, 1, NS - 1>(res, tw, twtab);
    // pass E: last inverse stage (stride M0), natural order out -> HBM; a thread's R0 columns are M0 apart, the TU lanes
    // of a column write one whole sector (TU = 4) or half of it (TU = 2, the other half comes from the neighbour CTA)
    {
        constexpr int m = M0;
        cpx* Zk = Z + (size_t)k * FW * CHp + u0 + l;
        const cpx* rl = res + l * LDL;
        for (int j = p0; j < m; j += PS) {
            const cpx* p = rl + P::pidx(j);
            auto fetch = [&](int q) -> cpx { return __ldg(&tw[j * q]); };
            std::conditional_t<((FLAGS & 1) && R0 <= 9), CtTw0<R0, true>, CtTw<R0, true, decltype(fetch)>> t;
            if constexpr ((FLAGS & 1) && R0 <= 9) t.load(tw, j); else t.load(fetch);
            float re[R0], im[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r) { const cpx v = t.apply(p[ct_off<P, 0>(r)], r); re[r] = v.x; im[r] = v.y; }
            dft_regs<R0, true>(re, im);
#pragma unroll
            for (int r = 0; r < R0; ++r) __stcs(&Zk[(size_t)(j + r * m) * CHp], make_float2(re[r], im[r]));
        }
    }
}

// ------------------------------------------------------------------------------- bp_inv_h_ct
// grid (FW / (2 NLN), nk) x NT; Z [k][FW][CHp] (already scaled) -> 2 NLN real columns of plane k (NLN complex lines)
// FLAGS: 1 = first-stage rotations by power chain, 4 = digit-reversed positions in arithmetic instead of the pos_of table
template <class P, int NT, int NLN, int MINB, int FLAGS>
__global__ void __launch_bounds__(NT, MINB) bp_inv_h_ct(const cpx* __restrict__ Z, int FW, int CH, int CHp,
                                                        const cpx* __restrict__ tw, const unsigned short* __restrict__ pos_of,
                                                        float* const* __restrict__ outs, int crop_h, int crop_w, int out_ld)
{
    constexpr int FH = P::N, NS = P::NS, M0 = P::M0, R0 = P::R(0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* lines = reinterpret_cast<cpx*>(smem_raw);
    const int k = blockIdx.y, x0 = blockIdx.x * 2 * NLN;
    if (x0 >= crop_w) return;
    constexpr int LDL = P::template ld<NLN>();
    cpx* twtab = lines + NLN * LDL;
    ct_fill_tw<P, NT>(twtab, tw);
    const cpx* Zk = Z + ((size_t)k * FW + x0) * CHp;
    constexpr int half = FH / 2;
    constexpr int MB = 4;                                   // 2 * MB global loads in flight per thread
#pragma unroll 1
    for (int l = 0; l < NLN; ++l) {
        cpx* ln = lines + l * LDL;
        const cpx* Za = Zk + (size_t)(2 * l) * CHp;
        const cpx* Zb = Za + CHp;
        for (int ub = threadIdx.x; ub < CH; ub += NT * MB) {
            cpx za[MB], zb[MB];
            int pa[MB], pb[MB];
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int uc = min(ub + b * NT, CH - 1);
                za[b] = __ldcs(&Za[uc]); zb[b] = __ldcs(&Zb[uc]);
                if constexpr (FLAGS & 4) { pa[b] = ct_pos<P>(uc); pb[b] = ct_pos<P>(uc == 0 ? 0 : FH - uc); }
                else { pa[b] = pos_of[uc]; pb[b] = pos_of[uc == 0 ? 0 : FH - uc]; }
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int u = ub + b * NT;
                if (u < CH) {
                    if (u == 0 || u == half) {                          // C2R ignores Im of DC / Nyquist
                        ln[P::pidx(pa[b])] = make_float2(za[b].x, zb[b].x);
                    } else {
                        ln[P::pidx(pa[b])] = make_float2(za[b].x - zb[b].y, za[b].y + zb[b].x);     // za + i zb
                        ln[P::pidx(pb[b])] = make_float2(za[b].x + zb[b].y, zb[b].x - za[b].y);     // conj(za) + i conj(zb)
                    }
                }
            }
        }
    }
    __syncthreads();
    ct_stages<P, true, NLN, NT, 1, NS>(lines, tw, twtab);
    // last inverse stage (stride M0) -> plane straight from registers: lanes run along y (contiguous in the plane)
    {
        float* o = outs[k];
        for (int it = threadIdx.x; it < M0 * NLN; it += NT) {
            const int l = it / M0, j = it - l * M0;
            const cpx* p = lines + l * LDL + P::pidx(j);
            auto fetch = [&](int q) -> cpx { return __ldg(&tw[j * q]); };
            std::conditional_t<((FLAGS & 1) && R0 <= 9), CtTw0<R0, true>, CtTw<R0, true, decltype(fetch)>> t;
            if constexpr ((FLAGS & 1) && R0 <= 9) t.load(tw, j); else t.load(fetch);
            float re[R0], im[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r) { const cpx v = t.apply(p[ct_off<P, 0>(r)], r); re[r] = v.x; im[r] = v.y; }
            dft_regs<R0, true>(re, im);
            const int xa = x0 + 2 * l, xb = xa + 1;
            float* oa = o + (size_t)xa * out_ld;
            float* ob = o + (size_t)xb * out_ld;
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const int y = j + r * M0;
                if (y < crop_h) {
                    if (xa < crop_w) __stcs(&oa[y], re[r]);
                    if (xb < crop_w) __stcs(&ob[y], im[r]);
                }
            }
        }
    }
}

}  // namespace fftconv
