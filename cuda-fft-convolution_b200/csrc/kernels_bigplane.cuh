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
};

// one pad slot per 16 points keeps the short-stride stages (stride 1, 2, ... points between lanes) off a single bank
__device__ __forceinline__ int bp_pidx(int i) { return i + (i >> 4); }

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

// ------------------------------------------------------------------------------- in-place stages
// One work item = one radix-R butterfly of one line.  DIF (forward): DFT_R over the R samples  base + r*m, then the
// output q is rotated by w_L^(j q).  DIT (inverse): the exact inverse — rotate input q by conj(w_L^(j q)), then the
// inverse DFT_R.  Twiddles for R >= 16 come from a two-level split q = 4a + b (R/4 + 2 table reads instead of R - 1).
template <int R, bool INV>
__device__ __forceinline__ void ip_stage(cpx* __restrict__ lines, int nl, int ldl, int n, int L,
                                         const cpx* __restrict__ tw) {
    const int m = L / R, nb = n / R, step = n / L;
    const int items = nb * nl;
    constexpr int G = (R >= 16 && R % 4 == 0) ? 4 : 1;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int l = it / nb, b = it - l * nb;
        const int blk = b / m, j = b - blk * m;
        cpx* ln = lines + (size_t)l * ldl;
        const int base = blk * L + j;
        cpx w1[G], wg[R / G];
        const bool rot = m > 1;                       // last stage: j = 0, every rotation is 1
        if (rot) {
            const int js = j * step;
#pragma unroll
            for (int q = 1; q < G; ++q) w1[q] = twd<INV>(__ldg(&tw[js * q]));
#pragma unroll
            for (int a = 1; a < R / G; ++a) wg[a] = twd<INV>(__ldg(&tw[js * G * a]));
        }
        float re[R], im[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cpx v = ln[bp_pidx(base + r * m)];
            if (INV && rot && r > 0) {
                const int a = r / G, q = r % G;
                v = (G == 1 || q == 0) ? cmul(v, wg[a]) : (a == 0 ? cmul(v, w1[q]) : cmul(v, cmul(wg[a], w1[q])));
            }
            re[r] = v.x; im[r] = v.y;
        }
        dft_regs<R, INV>(re, im);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cpx v = make_float2(re[r], im[r]);
            if (!INV && rot && r > 0) {
                const int a = r / G, q = r % G;
                v = (G == 1 || q == 0) ? cmul(v, wg[a]) : (a == 0 ? cmul(v, w1[q]) : cmul(v, cmul(wg[a], w1[q])));
            }
            ln[bp_pidx(base + r * m)] = v;
        }
    }
}

template <bool INV>
__device__ __forceinline__ void ip_run_stage(int R, cpx* lines, int nl, int ldl, int n, int L, const cpx* __restrict__ tw) {
    switch (R) {
        case 2:  ip_stage<2, INV>(lines, nl, ldl, n, L, tw); break;
        case 3:  ip_stage<3, INV>(lines, nl, ldl, n, L, tw); break;
        case 4:  ip_stage<4, INV>(lines, nl, ldl, n, L, tw); break;
        case 5:  ip_stage<5, INV>(lines, nl, ldl, n, L, tw); break;
        case 7:  ip_stage<7, INV>(lines, nl, ldl, n, L, tw); break;
        case 8:  ip_stage<8, INV>(lines, nl, ldl, n, L, tw); break;
        case 9:  ip_stage<9, INV>(lines, nl, ldl, n, L, tw); break;
        case 11: ip_stage<11, INV>(lines, nl, ldl, n, L, tw); break;
        case 13: ip_stage<13, INV>(lines, nl, ldl, n, L, tw); break;
        case 16: ip_stage<16, INV>(lines, nl, ldl, n, L, tw); break;
        case 17: ip_stage<17, INV>(lines, nl, ldl, n, L, tw); break;
        default: ip_stage<32, INV>(lines, nl, ldl, n, L, tw); break;
    }
}

// forward first stage when only the first nz <= n/R samples of a line are non-zero (a zero-padded template):
// the DFT_R collapses to a broadcast,  Y_q[j] = x[j] * w_n^(j q).
__device__ __forceinline__ void ip_stage_pruned_fwd(cpx* __restrict__ lines, int nl, int ldl, int n, int R,
                                                    const cpx* __restrict__ tw) {
    const int m = n / R;
    const int items = m * nl;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int l = it / m, j = it - l * m;
        cpx* ln = lines + (size_t)l * ldl;
        const cpx x = ln[bp_pidx(j)];
        for (int q = 1; q < R; ++q) ln[bp_pidx(j + q * m)] = cmul(x, __ldg(&tw[j * q]));
    }
}

// natural order in (first nz samples non-zero, the caller zero-filled up to bp_fill_to) -> digit-reversed out
__device__ __forceinline__ void ip_forward(cpx* lines, int nl, int ldl, const IpPlan& P, const cpx* __restrict__ tw, int nz) {
    int s = 0;
    if (nz <= P.L[0] / P.R[0]) {
        ip_stage_pruned_fwd(lines, nl, ldl, P.n, P.R[0], tw);
        __syncthreads();
        s = 1;
    }
    for (; s < P.ns; ++s) {
        ip_run_stage<false>(P.R[s], lines, nl, ldl, P.n, P.L[s], tw);
        __syncthreads();
    }
}
__device__ __forceinline__ int bp_fill_to(const IpPlan& P, int nz) { return nz <= P.L[0] / P.R[0] ? P.L[0] / P.R[0] : P.n; }

// digit-reversed in -> natural order out (unnormalised)
__device__ __forceinline__ void ip_inverse(cpx* lines, int nl, int ldl, const IpPlan& P, const cpx* __restrict__ tw) {
    for (int s = P.ns - 1; s >= 0; --s) {
        ip_run_stage<true>(P.R[s], lines, nl, ldl, P.n, P.L[s], tw);
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
    const int pf = blockIdx.y, s = pf / F, f = pf - s * F;
    const SrcDesc d = srcs[s];
    const int x0 = blockIdx.x * 4;
    if (x0 >= d.cols) return;
    const int rows = min(d.rows, FH);
    const float* base = d.ptr + (size_t)f * d.rows * d.cols;
    const int fill = bp_fill_to(plan, rows);
    for (int idx = threadIdx.x; idx < 2 * fill; idx += blockDim.x) {
        const int l = idx / fill, y = idx - l * fill;
        const int xa = x0 + 2 * l, xb = xa + 1;
        cpx v = make_float2(0.f, 0.f);
        if (y < rows) {
            if (xa < d.cols) v.x = base[(size_t)xa * d.rows + y];
            if (xb < d.cols) v.y = base[(size_t)xb * d.rows + y];
        }
        lines[(size_t)l * ldl + bp_pidx(y)] = v;
    }
    __syncthreads();
    ip_forward(lines, 2, ldl, plan, tw, rows);
    for (int idx = threadIdx.x; idx < 2 * CHp; idx += blockDim.x) {
        const int l = idx / CHp, u = idx - l * CHp;
        cpx a = make_float2(0.f, 0.f), b = a;
        if (u < CH) {
            const cpx* ln = lines + (size_t)l * ldl;
            const cpx zu = ln[bp_pidx(pos_of[u])];
            const cpx zn = cconj(ln[bp_pidx(pos_of[u == 0 ? 0 : FH - u])]);
            a = make_float2(0.5f * (zu.x + zn.x), 0.5f * (zu.y + zn.y));
            const cpx dd = make_float2(0.5f * (zu.x - zn.x), 0.5f * (zu.y - zn.y));
            b = make_float2(dd.y, -dd.x);                                   // -i * dd
        }
        cpx* o = T + ((size_t)pf * maxcols4 + x0 + 2 * l) * CHp + u;
        o[0] = a;
        o[CHp] = b;
    }
}

// ------------------------------------------------------------------------------- bp_conv_w
// grid (nk, CHp / TU) x 512, template index fastest: the CTAs resident at any moment share a handful of h-bin tiles,
// so the data spectrum is read from HBM once per call, not once per template.
// TU = 4 lines (single channel, product in place) or 2 lines + 2 accumulator lines (MULTI: channel sum).
template <bool CONJ, bool MULTI, int NT>
__global__ void __launch_bounds__(NT, 1) bp_conv_w(const cpx* __restrict__ T, const int* __restrict__ kcols, int maxcols4,
                                                    const cpx* __restrict__ Sp, int F, int FW, int CHp,
                                                    const __grid_constant__ IpPlan plan, const cpx* __restrict__ tw,
                                                    cpx* __restrict__ Z, int ldl)
{
    constexpr int TU = MULTI ? 2 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* lines = reinterpret_cast<cpx*>(smem_raw);
    cpx* acc = lines + (size_t)TU * ldl;                  // MULTI only
    const int k = blockIdx.x, u0 = blockIdx.y * TU;
    const int ncols = min(kcols[k], FW);
    const int fill = bp_fill_to(plan, ncols);

    constexpr int MB = 6;                                   // global loads in flight per thread
    for (int f = 0; f < F; ++f) {
        const cpx* Tp = T + ((size_t)(k * F + f) * maxcols4) * CHp + u0;
        for (int i0 = threadIdx.x; i0 < fill * TU; i0 += NT * MB) {
            cpx v[MB];
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int idx = i0 + b * NT, x = idx / TU, l = idx - x * TU;
                v[b] = make_float2(0.f, 0.f);
                if (x < ncols) v[b] = Tp[(size_t)x * CHp + l];
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int idx = i0 + b * NT, x = idx / TU, l = idx - x * TU;
                if (idx < fill * TU) lines[(size_t)l * ldl + bp_pidx(x)] = v[b];
            }
        }
        __syncthreads();
        ip_forward(lines, TU, ldl, plan, tw, ncols);
        const cpx* Sf = Sp + (size_t)f * FW * CHp + u0;     // rows already in digit-reversed order
        for (int i0 = threadIdx.x; i0 < FW * TU; i0 += NT * MB) {
            cpx dsp[MB];
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int idx = i0 + b * NT, p = idx / TU, l = idx - p * TU;
                if (idx < FW * TU) dsp[b] = __ldg(&Sf[(size_t)p * CHp + l]);
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int idx = i0 + b * NT, p = idx / TU, l = idx - p * TU;
                if (idx < FW * TU) {
                    const size_t o = (size_t)l * ldl + bp_pidx(p);
                    const cpx kx = lines[o];
                    cpx pr = CONJ ? cmulc(dsp[b], kx) : cmul(dsp[b], kx);
                    if (MULTI) {
                        if (f > 0) { const cpx a = acc[o]; pr.x += a.x; pr.y += a.y; }
                        acc[o] = pr;
                    } else {
                        lines[o] = pr;
                    }
                }
            }
        }
        __syncthreads();
    }
    cpx* res = MULTI ? acc : lines;
    ip_inverse(res, TU, ldl, plan, tw);
    cpx* Zk = Z + (size_t)k * FW * CHp + u0;
    for (int idx = threadIdx.x; idx < FW * TU; idx += blockDim.x) {
        const int x = idx / TU, l = idx - x * TU;
        Zk[(size_t)x * CHp + l] = res[(size_t)l * ldl + bp_pidx(x)];
    }
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
    const int k = blockIdx.y, x0 = blockIdx.x * 4;
    if (x0 >= crop_w) return;
    const cpx* Zk = Z + ((size_t)k * FW + x0) * CHp;
    const int half = FH / 2;
    constexpr int MB = 5;                                   // 2 * MB global loads in flight per thread
    for (int i0 = threadIdx.x; i0 < 2 * CH; i0 += 256 * MB) {
        cpx za[MB], zb[MB];
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            const int idx = i0 + b * 256;
            if (idx < 2 * CH) {
                const int l = idx / CH, u = idx - l * CH;
                za[b] = Zk[(size_t)(2 * l) * CHp + u];
                zb[b] = Zk[(size_t)(2 * l + 1) * CHp + u];
            }
        }
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            const int idx = i0 + b * 256;
            if (idx < 2 * CH) {
                const int l = idx / CH, u = idx - l * CH;
                cpx* ln = lines + (size_t)l * ldl;
                if (u == 0 || u == half) {                          // C2R ignores Im of DC / Nyquist
                    ln[bp_pidx(pos_of[u])] = make_float2(za[b].x, zb[b].x);
                } else {
                    ln[bp_pidx(pos_of[u])] = make_float2(za[b].x - zb[b].y, za[b].y + zb[b].x);           // za + i zb
                    ln[bp_pidx(pos_of[FH - u])] = make_float2(za[b].x + zb[b].y, zb[b].x - za[b].y);      // conj(za) + i conj(zb)
                }
            }
        }
    }
    __syncthreads();
    ip_inverse(lines, 2, ldl, plan, tw);
    float* o = outs[k];
    for (int idx = threadIdx.x; idx < 2 * crop_h; idx += blockDim.x) {
        const int l = idx / crop_h, y = idx - l * crop_h;
        const cpx r = lines[(size_t)l * ldl + bp_pidx(y)];
        const int xa = x0 + 2 * l, xb = xa + 1;
        if (xa < crop_w) o[(size_t)xa * out_ld + y] = r.x;
        if (xb < crop_w) o[(size_t)xb * out_ld + y] = r.y;
    }
}

}  // namespace fftconv
