// libfftconv.so — C ABI + host schedule of the B200-native FFT-convolution engine.
// See include/fftconv.h for the contract and the reference lines each entry point replaces.
#include <cuda.h>            // CUtensorMap types only: the driver entry point is fetched through the runtime
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/fftconv.h"
#include "kernels_generic.cuh"
#include "kernels_tile16.cuh"
#include "kernels_osgemm.cuh"
#include "kernels_bigplane.cuh"
#include "kernels_bigplane_ct.cuh"

namespace fftconv {

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(FFTCONV_ERR_CUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e_,      \
                        cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

#define LAUNCH_CHECK()                                                                     \
    do {                                                                                   \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        CU(cudaGetLastError());                                                            \
    } while (0)

// Optional per-kernel timing (bench.py roofline leg): every launch of a profiled kind is
// bracketed by CUDA events on the stream it is launched on.
enum ProfKind { PK_RELAYOUT = 0, PK_KERN_H, PK_CONV, PK_C2R, PK_DATA_H, PK_DATA_W, PK_GEN_H, PK_GEN_W, PK_GEN_C2R,
                PK_OS_PLANE, PK_OS_DATA, PK_OS_KERN, PK_OS_GEMM, PK_OS_INV, PK_BP_REPAD, PK_BP_KERN_H, PK_BP_CONV_W, PK_BP_INV_H, PK_COUNT };
static const char* kProfNames[PK_COUNT] = {"tile16_relayout", "tile16_kern_hpass", "tile16_conv", "tile16_c2r",
                                           "fwd_h_pass(data)", "fwd_w_pass(data)", "fwd_h_pass(kernels)",
                                           "conv_w_pass_generic", "inv_h_pass",
                                           "os_spectrum_to_plane", "os_data_fft(tiles)",
                                           "os_kern_fft(templates)", "os_gemm", "os_inverse",
                                           "bp_repad_spec", "bp_kern_h", "bp_conv_w", "bp_inv_h"};
static bool g_prof_on = false;
static thread_local bool g_capturing = false;      // plan capture on this thread: no timing events inside a graph
struct ProfRec { int kind; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
struct ProfScope {
    int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s) {
        if (g_prof_on && !g_capturing) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (a) { cudaEventRecord(b, st); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back({kind, a, b}); }
    }
};

static const char* kMsgThread =
    "CUDA Thread Size must be 4 integers : THREAD_PER_BLOCK_H, THREAD_PER_BLOCK_W, "
    "THREAD_PER_BLOCK_D, THREAD_PER_BLOCK_2D\nYou must choose size such that total thread will "
    "not be larger than MaxThreadsPerBlock";
static const char* kMsgKernelShape =
    "Kernel and Data must have the same number of features and kernel size should be smaller "
    "than data size";

// ------------------------------------------------------------------------------ plans
static LinePlan make_line_plan(int n) {
    LinePlan p;
    p.n = n;
    p.nstages = 0;
    int m = n;
    auto push = [&](int r) { p.radix[p.nstages++] = r; m /= r; };
    while (m % 16 == 0) push(16);
    if (m % 8 == 0) push(8);
    if (m % 4 == 0) push(4);
    if (m % 2 == 0) push(2);
    while (m % 9 == 0) push(9);
    const int small[] = {3, 5, 7, 11, 13, 17};
    for (int r : small)
        while (m % r == 0) push(r);
    for (int r = 19; m > 1; r += 2)
        while (m % r == 0) push(r);
    return p;
}

static const size_t kMaxSmem = 227 * 1024;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Ctx {
    int dev = -1;
    bool inited = false;
    std::map<int, cpx*> tw;          // n -> device twiddle table e^{-2 pi i j / n}
    std::map<int, unsigned short*> ipmap;   // n -> digit-reversal tables of the in-place plan: pos_of[n], nat_of[n]
    DevBuf bpS;                      // large-plane path: re-padded, pre-scaled data spectrum
    DevBuf batchA;                   // fftconv_conv_batch: template spectra shared by the image groups of one call
    DevBuf T, Z, stage, desc, outstage, dspec, ddata, priv, Ag, Wg;
    DevBuf osA, osB, osP, osPlane, osZ, osPeaks;     // overlap-save / tcgen05 path scratch
    // plan capture (fftconv_plan_*): descriptor tables are staged through a PLAN-OWNED pinned buffer with bump allocation,
    // because a captured H2D copy reads its pinned source at every graph launch, long after this call returned
    char* plan_pin = nullptr;
    size_t plan_pin_cap = 0, plan_pin_off = 0;
    void* bounce = nullptr;          // pinned bounce ring for PAGEABLE host outputs (two halves of one chunk of planes)
    size_t bounce_cap = 0;
    cudaEvent_t evb[2] = {nullptr, nullptr};   // D2H into bounce half done
    void* pinned = nullptr;          // host staging (descriptors, packed kernels)
    size_t pinned_cap = 0;
    cudaEvent_t pinned_free = nullptr;   // recorded after the last async copy out of `pinned`
    cudaStream_t side = nullptr;         // copy stream for the chunk pipeline
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t side2 = nullptr;        // data-side transforms of the overlap-save path run here, next to os_kern_fft
    cudaEvent_t evf[4] = {nullptr, nullptr, nullptr, nullptr};   // [0..1] data side; [2..3] template transforms run ahead (OsAhead)
    int sm_count = 148;
    // Provenance of the last spectrum fftconv_fft_data produced on this device.  The overlap-save path works from the raw
    // data, not from the compat spectrum; a caller of the two-call interface hands the spectrum back, and inverting it to a
    // plane again was 6 % of the config-2 step.  fft_data keeps a private copy of the raw data and a 64-bit hash of the
    // spectrum it wrote; the convolution re-hashes what it is given ON THE DEVICE and its data-side kernels choose between
    // the kept raw data (hashes equal) and the inverse of the spectrum (anything else: a spectrum the caller modified,
    // another buffer that happens to live at the same address, ...).  No host synchronisation, no trust in pointers.
    struct SpecCache {
        bool valid = false;
        const void* spec = nullptr;
        int H = 0, W = 0, F = 0, FH = 0, FW = 0;
        DevBuf raw, hash;                // raw data [F][W][H]; hash[0] = at fft_data time, hash[1] = at convolution time
        // tile spectra (B images) computed next to the forward transform, on the data-side stream
        bool b_valid = false;
        int b_maxkh = 0, b_maxkw = 0;
        long long b_gen = -1;
        // geometry of the last spectrum-fed overlap-save convolution: decides whether the next fft_data tiles eagerly
        bool want = false;
        int w_F = 0, w_FH = 0, w_FW = 0, w_maxkh = 0, w_maxkw = 0;
        // the last spectrum-fed single-image convolution and whether it took the overlap-save path: a geometry that is
        // served by another pipeline (config 1: 3 extra launches = 8 of 67 us per call) does not record provenance at all
        bool hist = false, hist_os = false;
        int h_F = 0, h_FH = 0, h_FW = 0;
    } sc;
    long long osB_gen = 0;               // bumped whenever os_data_fft (re)writes osB
    cudaEvent_t spec_ready = nullptr;    // fftconv_spectrum_ready_event: one-shot dependency of the data-side work
    // One lock per device: a call holds it from its first touch of the cached scratch to its last enqueue (host-output
    // calls: to the final synchronisation), so calls on different devices never serialise each other (the reference's
    // own multi-GPU prototype is such a caller, src/cudaConvFFTDataStreams.cu:273-328).  Recursive: the one-shot entry
    // points nest the two-call ones.
    std::recursive_mutex mu;
    int depth = 0;
    cudaEvent_t last_use = nullptr;      // recorded behind the last enqueue of every call
    cudaStream_t last_stream = nullptr;
    bool used = false;
};

static std::mutex g_mu;                  // guards g_ctx (the map only; each Ctx has its own lock)
static std::map<int, Ctx> g_ctx;

static std::atomic<long long> g_scratch_gen{0};   // bumped whenever cached device scratch moves (invalidates captured plans)

static int dev_reserve(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    g_scratch_gen.fetch_add(1, std::memory_order_relaxed);
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t cap = bytes + bytes / 8 + 256;
    CU(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return 0;
}

static int pinned_reserve(Ctx& c, size_t bytes) {
    if (bytes <= c.pinned_cap) return 0;
    if (c.pinned) {
        CU(cudaEventSynchronize(c.pinned_free));
        CU(cudaFreeHost(c.pinned));
    }
    c.pinned = nullptr; c.pinned_cap = 0;
    size_t cap = bytes + bytes / 4 + 4096;
    CU(cudaMallocHost(&c.pinned, cap));
    c.pinned_cap = cap;
    return 0;
}

// Host staging for a descriptor table that an async H2D copy on `st` will read: *out = pinned memory of `bytes`.
// Normal calls: the shared staging buffer, once its previous copy has finished.  Plan capture: a fresh region of the
// plan's own buffer.  pinned_done marks the copy as enqueued.
static int pinned_get(Ctx& c, size_t bytes, void** out) {
    if (c.plan_pin) {
        const size_t off = (c.plan_pin_off + 63) & ~(size_t)63;
        if (off + bytes > c.plan_pin_cap) return fail(FFTCONV_ERR_UNSUPPORTED, "plan staging buffer too small");
        c.plan_pin_off = off + bytes;
        *out = c.plan_pin + off;
        return 0;
    }
    if (int e = pinned_reserve(c, bytes)) return e;
    CU(cudaEventSynchronize(c.pinned_free));
    *out = c.pinned;
    return 0;
}
static int pinned_done(Ctx& c, cudaStream_t st) {
    if (c.plan_pin) return 0;
    CU(cudaEventRecord(c.pinned_free, st));
    return 0;
}

template <typename K>
static int opt_in_smem(K kernel) {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kernel));                 // static shared memory counts against the 227 KB
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMaxSmem - fa.sharedSizeBytes)));
    return 0;
}

static int ctx_get(int device, Ctx** out) {
    Ctx* cp;
    { std::lock_guard<std::mutex> lk(g_mu); cp = &g_ctx[device]; }      // map nodes are stable
    Ctx& c = *cp;
    std::lock_guard<std::recursive_mutex> lkc(c.mu);
    if (!c.inited) {
        c.dev = device;
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(FFTCONV_ERR_CUDA, "fftconv-b200 is built for sm_100a only; device %d is sm_%d%d",
                        device, prop.major, prop.minor);
        c.sm_count = prop.multiProcessorCount;
        // odd-radix coefficient table (cos, sin)(2 pi j / R)
        static float2 h_tw[32][32];
        for (int r = 1; r < 32; ++r)
            for (int j = 0; j < 32; ++j) {
                const double a = 2.0 * M_PI * (double)(j % r) / (double)r;
                h_tw[r][j] = make_float2((float)cos(a), (float)sin(a));
            }
        CU(cudaMemcpyToSymbol(c_odd_tw, h_tw, sizeof h_tw));
        {   // w64^(r0 c) for the run-time-residue transform task of os_kern_fft (same constexpr series as the compile-time twiddles)
            static float2 h_w64[4][16];
            for (int r0 = 0; r0 < 4; ++r0)
                for (int cc = 0; cc < 16; ++cc)
                    h_w64[r0][cc] = make_float2((float)os_cos64d((r0 * cc) & 63), (float)os_sin64d((r0 * cc) & 63));
            CU(cudaMemcpyToSymbol(c_os_w64, h_w64, sizeof h_w64));
        }
        CU(cudaEventCreateWithFlags(&c.pinned_free, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c.last_use, cudaEventDisableTiming));
        for (auto& e : c.ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
        {   // data-side transforms are the critical path of a call: they outrank the template transforms they overlap
            int lo = 0, hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU(cudaStreamCreateWithPriority(&c.side2, cudaStreamNonBlocking, hi));
        }
        for (auto& e : c.evf) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : c.evb) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (opt_in_smem(fwd_h_pass<PAD_ZERO>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(fwd_h_pass<PAD_CLAMP>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(fwd_w_pass)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(conv_w_pass_generic<false>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(conv_w_pass_generic<true>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(inv_h_pass)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(tile16_conv<false>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(tile16_conv<true>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(tile16_c2r)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_kern_fft<1>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_kern_fft<2>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_data_fft)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_gemm)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_inverse)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_inverse_tma<0>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_inverse_tma<6>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(os_inverse_z)) return FFTCONV_ERR_CUDA;

        if (opt_in_smem(inv_w_pass)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(bp_conv_w<false, false, 512, 4, 1>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(bp_conv_w<true, false, 512, 4, 1>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(bp_conv_w<false, true, 512, 2, 1>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(bp_conv_w<true, true, 512, 2, 1>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(bp_kern_h<2>)) return FFTCONV_ERR_CUDA;
        if (opt_in_smem(bp_inv_h<2>)) return FFTCONV_ERR_CUDA;
        c.inited = true;
    }
    *out = &c;
    return 0;
}

static int get_twiddles(Ctx& c, int n, cudaStream_t st, const cpx** out) {
    auto it = c.tw.find(n);
    if (it == c.tw.end()) {
        std::vector<float2> h(n);
        for (int j = 0; j < n; ++j) {
            const double a = -2.0 * M_PI * (double)j / (double)n;
            h[j] = make_float2((float)cos(a), (float)sin(a));
        }
        cpx* d = nullptr;
        CU(cudaMalloc(&d, sizeof(float2) * (size_t)n));
        CU(cudaMemcpy(d, h.data(), sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice));
        it = c.tw.emplace(n, d).first;
    }
    (void)st;
    *out = it->second;
    return 0;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// A library call on `device` / `stream`: takes the device's lock, selects the device, and orders the call behind the
// previous call when that one ran on a DIFFERENT stream (the cached scratch -- operand images, product spectra, staging --
// is shared by all streams of a device: without this two device-output calls on two streams would race on it).
struct CtxScope {
    Ctx* c = nullptr;
    int err = 0;
    cudaStream_t st;
    std::unique_lock<std::recursive_mutex> lk;
    DeviceGuard guard;
    CtxScope(int device, cudaStream_t stream) : st(stream), guard(device) {
        if (!guard.ok) { err = fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device); return; }
        if ((err = ctx_get(device, &c))) { c = nullptr; return; }
        lk = std::unique_lock<std::recursive_mutex>(c->mu);
        if (c->depth++ == 0 && c->used && c->last_stream != st) {
            if (cudaStreamWaitEvent(st, c->last_use, 0) != cudaSuccess)
                err = fail(FFTCONV_ERR_CUDA, "cudaStreamWaitEvent failed");
        }
    }
    ~CtxScope() {
        if (!c || !lk.owns_lock()) return;
        if (--c->depth == 0) {
            cudaEventRecord(c->last_use, st);
            c->last_stream = st; c->used = true;
        }
    }
};

static inline int odd_ld(int n) { return n | 1; }   // line stride (float2) avoiding bank conflicts

// lines per CTA for the h passes, sized to ~64 KB of shared memory and <= 16
static int pick_lines(int n, int want_max) {
    const size_t per_line = 2 * (size_t)odd_ld(n) * sizeof(cpx);
    int nl = (int)((64 * 1024) / per_line);
    if (nl < 1) nl = 1;
    if (nl > want_max) nl = want_max;
    return nl;
}

// ------------------------------------------------------------------ data spectrum (cudaFFTData)
// d_data: device [F][W][H].  Writes the compat spectrum [F][FW][CH].
static int run_fft_data(Ctx& c, const float* d_data, int H, int W, int F, int FH, int FW,
                        int pad_mode, int kernel_y, int kernel_x, cpx* d_spec, cudaStream_t st) {
    const int CH = FH / 2 + 1;
    const cpx *twH, *twW;
    if (int e = get_twiddles(c, FH, st, &twH)) return e;
    if (int e = get_twiddles(c, FW, st, &twW)) return e;
    const LinePlan pH = make_line_plan(FH), pW = make_line_plan(FW);
    const int ldH = odd_ld(FH), ldW = odd_ld(FW);
    if (2 * (size_t)ldH * sizeof(cpx) > kMaxSmem || 2 * (size_t)ldW * sizeof(cpx) > kMaxSmem)
        return fail(FFTCONV_ERR_UNSUPPORTED, "FFT plane %dx%d exceeds the shared-memory line limit", FH, FW);

    const int ncols = pad_mode == PAD_CLAMP ? FW : W;
    if (int e = dev_reserve(c.T, sizeof(cpx) * (size_t)F * ncols * CH)) return e;
    // one SrcDesc through pinned staging
    if (int e = dev_reserve(c.desc, sizeof(SrcDesc))) return e;
    SrcDesc* hd;
    if (int e = pinned_get(c, sizeof(SrcDesc), (void**)&hd)) return e;
    hd->ptr = d_data; hd->rows = H; hd->cols = W;
    CU(cudaMemcpyAsync(c.desc.p, hd, sizeof(SrcDesc), cudaMemcpyHostToDevice, st));
    if (int e = pinned_done(c, st)) return e;

    const int NL = pick_lines(FH, 8);
    const long long nlines = (long long)F * ((ncols + 1) / 2);
    const unsigned grid = (unsigned)((nlines + NL - 1) / NL);
    const size_t smemH = 2 * (size_t)NL * ldH * sizeof(cpx);
    {
    ProfScope ps(PK_DATA_H, st);
    if (pad_mode == PAD_CLAMP)
        fwd_h_pass<PAD_CLAMP><<<grid, 256, smemH, st>>>((const SrcDesc*)c.desc.p, 1, F, W, FH, CH, pH, twH,
                                                         (cpx*)c.T.p, NL, ldH, kernel_y, kernel_x, FW);
    else
        fwd_h_pass<PAD_ZERO><<<grid, 256, smemH, st>>>((const SrcDesc*)c.desc.p, 1, F, W, FH, CH, pH, twH,
                                                        (cpx*)c.T.p, NL, ldH, 0, 0, W);
    LAUNCH_CHECK();
    }

    int TU = (int)((96 * 1024) / (2 * (size_t)ldW * sizeof(cpx)));
    TU = TU < 1 ? 1 : (TU > 16 ? 16 : TU);
    dim3 g2((CH + TU - 1) / TU, F);
    ProfScope ps(PK_DATA_W, st);
    fwd_w_pass<<<g2, 256, 2 * (size_t)TU * ldW * sizeof(cpx), st>>>((const cpx*)c.T.p, ncols, FW, CH, pW, twW,
                                                                     d_spec, TU, ldW);
    LAUNCH_CHECK();
    return 0;
}


// ------------------------------------------------------------------ tile16 fast path (host)
struct Tile16Cfg {
    int mh, mw, NT, nya, nxa, XC, XCP, KB, nstage, nmain, nextra;
    size_t dp_bytes, a_bytes, smem;
};

static bool tile16_config(int FH, int FW, int maxkh, int maxkw, Tile16Cfg& g) {
    g = Tile16Cfg{};
    g.mh = FH / 16; g.mw = FW / 16; g.NT = g.mh / 2 + 1;
    g.nya = (maxkh + 15) / 16; g.nxa = (maxkw + 15) / 16;
    g.XC = 16 * g.nxa; g.XCP = g.XC + 2;
    if (FW > 544 || g.nxa > 4 || g.nya > 16) return false;
    g.dp_bytes = (size_t)16 * g.mw * T16_PAD * sizeof(cpx);
    // pick the number of templates per CTA: 16 warps own `nmain` items in registers, the rest are
    // dealt round-robin as `nextra` warp-passes per channel; maximise issue balance over the 4
    // sub-partitions, then data-spectrum reuse (larger KB)
    double best = -1.0;
    for (int KB = 1; KB <= 8; ++KB) {
        const int N = KB * 16 * g.mw;
        int nmain, nextra, W;
        if (N <= T16_THREADS) { nmain = N; nextra = 0; W = (N + 31) / 32; }
        else { nmain = T16_THREADS; nextra = (N - T16_THREADS + 31) / 32; W = 16; }
        if (nextra > 6) break;
        const size_t a_bytes = (size_t)KB * 16 * g.XCP * sizeof(cpx);
        const size_t y_bytes = 2 * (size_t)KB * 256 * (g.mw | 1) * sizeof(cpx);
        int nstage = 0;
        size_t smem = 0;
        for (int ns = 4; ns >= 2; --ns) {
            const size_t pipe = (size_t)ns * (g.dp_bytes + a_bytes);
            const size_t tot = ((std::max(pipe, y_bytes) + 15) & ~(size_t)15) + 192 + (size_t)nextra * 2048 * sizeof(cpx);
            if (tot <= kMaxSmem) { nstage = ns; smem = tot; break; }
        }
        if (!nstage) continue;
        const double eff = (N / 32.0) / (4.0 * ((W + 3) / 4 + nextra / 4.0)) + 1e-3 * KB;
        if (eff > best) {
            best = eff;
            g.KB = KB; g.nmain = nmain; g.nextra = nextra; g.nstage = nstage; g.a_bytes = a_bytes; g.smem = smem;
        }
    }
    return best > 0.0;
}

static bool tile16_supported(int FH, int FW, int maxkh, int maxkw) {
    Tile16Cfg g;
    return tile16_config(FH, FW, maxkh, maxkw, g);
}

static size_t tile16_scratch_per_kernel(int FH, int FW, int F, int maxkh, int maxkw) {
    Tile16Cfg g;
    tile16_config(FH, FW, maxkh, maxkw, g);
    return sizeof(cpx) * ((size_t)g.NT * F * 16 * g.XCP + (size_t)FW * g.NT * 16);
}

static int tile16_prepare(Ctx& c, const cpx* d_spec, int FH, int FW, int F, int maxkh, int maxkw, int KC,
                          cudaStream_t st) {
    Tile16Cfg g;
    tile16_config(FH, FW, maxkh, maxkw, g);
    const int NG = (KC + g.KB - 1) / g.KB;
    if (int e = dev_reserve(c.Ag, (size_t)g.NT * NG * F * g.a_bytes)) return e;
    if (int e = dev_reserve(c.Wg, sizeof(cpx) * (size_t)KC * FW * g.NT * 16)) return e;
    const size_t dp_total = (size_t)g.NT * F * g.dp_bytes;
    if (int e = dev_reserve(c.priv, dp_total)) return e;
    const long long n = (long long)(dp_total / sizeof(cpx));
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)c.sm_count * 16);
    ProfScope ps(PK_RELAYOUT, st);
    tile16_relayout<<<grid, 256, 0, st>>>(d_spec, (cpx*)c.priv.p, F, FH, FW, FH / 2 + 1, g.mh, g.mw, g.NT);
    LAUNCH_CHECK();
    return 0;
}

static int tile16_chunk(Ctx& c, int FH, int FW, int F, int maxkh, int maxkw, const SrcDesc* d_descs, int nk,
                        float* const* d_outptrs, const fftconv_options& opt, cudaStream_t st) {
    Tile16Cfg g;
    tile16_config(FH, FW, maxkh, maxkw, g);
    const cpx *twH, *twW, *twMw, *twMh;
    if (int e = get_twiddles(c, FH, st, &twH)) return e;
    if (int e = get_twiddles(c, FW, st, &twW)) return e;
    if (int e = get_twiddles(c, g.mw, st, &twMw)) return e;
    if (int e = get_twiddles(c, g.mh, st, &twMh)) return e;
    const int NG = (nk + g.KB - 1) / g.KB;
    // 1. template h transforms into the tile-major private layout
    {
        const long long total = (long long)nk * F * g.NT * g.XC;
        ProfScope ps(PK_KERN_H, st);
        tile16_kern_hpass<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(d_descs, nk, F, FH, g.mh, g.NT, g.nya, g.XC,
                                                                           g.KB, NG, twH, (cpx*)c.Ag.p);
        LAUNCH_CHECK();
    }
    // 2. fused forward-w / multiply-accumulate / inverse-w / 16-point inverse-h
    {
        Tile16Params P;
        P.Dp = (const cpx*)c.priv.p; P.Ag = (const cpx*)c.Ag.p; P.Wg = (cpx*)c.Wg.p;
        P.twW = twW; P.twH = twH; P.twM = twMw; P.planM = make_line_plan(g.mw);
        P.F = F; P.FH = FH; P.FW = FW; P.mh = g.mh; P.mw = g.mw; P.NT = g.NT; P.NG = NG; P.KB = g.KB;
        P.nk = nk; P.nxa = g.nxa; P.nstage = g.nstage; P.nmain = g.nmain; P.nextra = g.nextra;
        const int threads = ((g.nmain + 31) / 32) * 32;
        dim3 grid(NG, g.NT);
        ProfScope ps(PK_CONV, st);
        if (opt.correlate) tile16_conv<true><<<grid, threads, g.smem, st>>>(P);
        else tile16_conv<false><<<grid, threads, g.smem, st>>>(P);
        LAUNCH_CHECK();
    }
    // 3. mh-point inverse along h, scale, crop, store
    {
        const int mhp = g.mh | 1;
        int NP = (int)((64 * 1024) / (2 * 16 * (size_t)mhp * sizeof(cpx)));
        NP = NP < 1 ? 1 : (NP > 8 ? 8 : NP);
        const long long nlines = (long long)nk * (FW / 2);
        const int crop_h = opt.crop_h > 0 ? opt.crop_h : FH;
        const int crop_w = opt.crop_w > 0 ? opt.crop_w : FW;
        const int out_ld = opt.out_ld > 0 ? opt.out_ld : crop_h;
        ProfScope ps(PK_C2R, st);
        tile16_c2r<<<(unsigned)((nlines + NP - 1) / NP), 256, 2 * (size_t)NP * 16 * mhp * sizeof(cpx), st>>>(
            (const cpx*)c.Wg.p, nk, FH, FW, g.mh, g.NT, make_line_plan(g.mh), twMh, 1.0f / ((float)FW * (float)FH),
            d_outptrs, crop_h, crop_w, out_ld, NP, mhp);
        LAUNCH_CHECK();
    }
    return 0;
}


// ------------------------------------------------------ overlap-save + tcgen05 GEMM path (host)
struct OsCfg {
    int F, FH, FW, maxkh, maxkw;
    int NFK, XCK;                 // template transform: 16*NFK leading samples per side
    int Sh, Sw, nth, ntw, NT;     // valid outputs per tile side, tile grid, tiles of the whole batch
    int nimg, NTimg;              // images in the batch, tiles per image
    int NKS, KC;                  // K stages per item, 16-byte k units per stage (even)
    int NNB, NTn, NMMA, RS, RSP;  // tile blocks, tiles per block, MMA N, P row length and row pitch (floats)
    int nsta;                     // A ring depth
    size_t gemm_smem, inv_smem;
    size_t a_stage, b_buf, p_blk; // bytes
    const OsLevel* d_levels = nullptr;   // pyramid batch (fftconv_conv_pyramid): per-level geometry on the device
    int nlevels = 0;
};

static bool os_config_tiles(OsCfg& g);
static bool os_config(int F, int FH, int FW, int maxkh, int maxkw, OsCfg& g, int nimg = 1) {
    g = OsCfg{};
    if (maxkh > 32 || maxkw > 32 || maxkh < 1 || maxkw < 1) return false;
    g.F = F; g.FH = FH; g.FW = FW; g.maxkh = maxkh; g.maxkw = maxkw;
    g.NFK = std::max(maxkh, maxkw) <= 16 ? 1 : 2;
    g.XCK = 16 * g.NFK;
    g.Sh = 65 - maxkh; g.Sw = 65 - maxkw;
    g.nth = (FH + g.Sh - 1) / g.Sh; g.ntw = (FW + g.Sw - 1) / g.Sw;
    g.nimg = nimg; g.NTimg = g.nth * g.ntw;
    g.NT = g.nimg * g.NTimg;
    return os_config_tiles(g);
}
// everything that follows from the NUMBER of tiles (the GEMM does not care where a tile comes from)
static bool os_config_tiles(OsCfg& g) {
    const int F = g.F;
    g.NKS = (2 * F + 31) / 32;
    const int units = (2 * F + 3) / 4;
    g.KC = ((units + g.NKS - 1) / g.NKS + 1) & ~1;
    g.NNB = (g.NT + 39) / 40;
    g.NTn = (g.NT + g.NNB - 1) / g.NNB;
    g.NMMA = (2 * g.NTn + 15) & ~15;
    g.RS = (2 * g.NTn + 7) & ~7;
    // Row pitch of the epilogue staging tile of os_gemm: pitch/4 odd puts the 8 lanes of a 128-bit store wavefront on 8 bank
    // groups (dense rows of 80 floats: 2).  P itself stays dense: the tile leaves through a TMA tensor store whose box is
    // RSP wide while the tensor is RS wide, so the pad columns are clipped.  (A padded P was tried first: os_gemm 7.09 ->
    // 5.56 ms at config 4, but the 32-byte boxes of the inverse then straddle sectors: 6.75 -> 7.19 ms.)
    g.RSP = ((g.RS >> 2) & 1) ? g.RS : g.RS + 4;
    g.a_stage = (size_t)2 * g.KC * OS_TM * 16;
    g.b_buf = (size_t)g.NKS * g.KC * g.NMMA * 16;                       // fp32 B image of one (tile block, bin)
    g.p_blk = (size_t)OS_TM * g.RS * 4;
    if (g.b_buf >= (1u << 20) || g.a_stage >= (1u << 20)) return false;       // mbarrier tx-count range
    g.nsta = 0;                                                           // raw A ring depth (fp32 K-stages in flight)
    for (int ns = 8; ns >= 2; --ns) {
        const size_t tot = (size_t)(ns + 2) * (g.a_stage / 2) + 4 * g.b_buf + (size_t)OS_TM * g.RSP * 4 + 512;   // + raw and lo images of B, double-buffered
        if (tot <= kMaxSmem) { g.nsta = ns; g.gemm_smem = tot; break; }
    }
    if (!g.nsta) return false;
    g.inv_smem = (size_t)OS_IG * OS_ITILE * sizeof(cpx);
    if (g.inv_smem > kMaxSmem) return false;
    return true;
}

// Debug / A-B switches of the overlap-save path, read ONCE (first use), never on the per-chunk path:
//   FFTCONV_OS_MIN_K      smallest bank that takes the overlap-save path automatically (default 64)
//   FFTCONV_OS_NTBLK      template blocks of 128 per chunk (default: 8 device outputs, 2 host outputs)
//   FFTCONV_OS_INV_TMA    0: per-thread cp.async gather in the inverse (os_inverse) instead of TMA tensor copies
//   FFTCONV_OS_GEMM       "simt": validation GEMM on the SIMT pipe
//   FFTCONV_OS_GEMM_TMAP  0: one bulk copy per template row in the GEMM epilogue instead of one bulk store per item
//   FFTCONV_OS_HI_INPLACE 1: rewrite the A stage as tf32(a) in shared memory instead of relying on the operand truncation
//   FFTCONV_OS_LBO_SWAP   swap the LBO / SBO fields of the shared-memory descriptors
//   FFTCONV_SPEC_CACHE    0: never reuse the raw data behind a spectrum of fftconv_fft_data (always invert the spectrum);
//                         1 (default): reuse the raw data; 2: also transform the tiles next to the forward transform
//   FFTCONV_BP_CT         -1: large-plane path on the run-time-plan kernels only; v >= 0 (default 0): variant v of the
//                         size-specialised kernels where one is instantiated for the line length (kernels_bigplane_ct.cuh)
//   FFTCONV_OS_AHEAD      1 / 2: with several chunks of device-resident templates, os_kern_fft of chunk i + 1 runs on a side stream
//                         (1: the high-priority one, 2: the copy stream) next to os_inverse_z of chunk i (default 0)
//   FFTCONV_OS_DATA       2: os_data_fft_occ (rows overlay the raw windows, channels of the w step one after the other) instead of os_data_fft
//   FFTCONV_OS_DATA_ST256 0: two 128-bit stores per 32-byte unit of the B image in os_data_fft instead of one 256-bit store (default 1:
//                         config 4 os_data_fft 13.2 -> 9.9 ms, the kernel was bound by its store requests)
//   FFTCONV_OS_PF         L2 prefetch distance of os_gemm's TMA producer in work items (default 0 = off: measured 0.21 -> 0.30 ms at config 2 with 6 items ahead, the prefetched lines fight the P stores for L2)
struct OsEnv { int min_k, ntblk, inv_tma, gemm_simt, gemm_tmap, hi_inplace, lbo_swap, dbg, pf, spec_cache, bp_ct, ahead, data_occ, data_st256; };
static const OsEnv& os_env() {
    static const OsEnv e = [] {
        auto geti = [](const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; };
        OsEnv x;
        x.min_k = geti("FFTCONV_OS_MIN_K", 64);
        x.ntblk = geti("FFTCONV_OS_NTBLK", 0);
        x.inv_tma = geti("FFTCONV_OS_INV_TMA", 1);
        const char* m = getenv("FFTCONV_OS_GEMM");
        x.gemm_simt = m && !strcmp(m, "simt");
        x.gemm_tmap = geti("FFTCONV_OS_GEMM_TMAP", 1);
        x.hi_inplace = geti("FFTCONV_OS_HI_INPLACE", 0);
        x.lbo_swap = geti("FFTCONV_OS_LBO_SWAP", 0);
        x.dbg = geti("FFTCONV_OS_DBG", 0);
        x.pf = geti("FFTCONV_OS_PF", 0);
        x.spec_cache = geti("FFTCONV_SPEC_CACHE", 1);
        x.bp_ct = geti("FFTCONV_BP_CT", 0);
        x.ahead = geti("FFTCONV_OS_AHEAD", 0);
        x.data_occ = geti("FFTCONV_OS_DATA", 0) == 2;
        x.data_st256 = geti("FFTCONV_OS_DATA_ST256", 1);
        return x;
    }();
    return e;
}

// CUtensorMap of P seen as {RS floats, 128 templates, bins}, box {8 floats, 1 template, 64 bins} (the inverse's gather).  cuTensorMapEncodeTiled is fetched through the
// runtime (cudaGetDriverEntryPoint): the library does not link libcuda.
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int os_make_p_tensor_map(const float* P, int RS, int RSP, unsigned long long nbins, OsTensorMap* out) {
    static PFN_tmapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return fail(FFTCONV_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = (PFN_tmapEncodeTiled)fn;
    }
    static_assert(sizeof(CUtensorMap) == sizeof(OsTensorMap), "tensor map size");
    const cuuint64_t gdim[3] = {(cuuint64_t)RS, OS_TM, (cuuint64_t)nbins};
    const cuuint64_t gstride[2] = {(cuuint64_t)RSP * 4, (cuuint64_t)RSP * 4 * OS_TM};
    const cuuint32_t box[3] = {8, 1, 64}, estride[3] = {1, 1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(P), gdim,
                              gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);   // (L2 promotion 64/128/256 B: no effect measured)
    if (r != CUDA_SUCCESS) return fail(FFTCONV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}
// P as the os_gemm epilogue stores it: {RS floats, 128 templates, items}, box {RSP, 128, 1} (RSP >= RS: the pad columns of
// the staging tile are outside the tensor and clipped)
static int os_make_p_store_map(float* P, int RS, int RSP, unsigned long long nitems, OsTensorMap* out) {
    static PFN_tmapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return fail(FFTCONV_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = (PFN_tmapEncodeTiled)fn;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)RS, OS_TM, (cuuint64_t)nitems};
    const cuuint64_t gstride[2] = {(cuuint64_t)RS * 4, (cuuint64_t)RS * 4 * OS_TM};
    const cuuint32_t box[3] = {(cuuint32_t)RSP, OS_TM, 1}, estride[3] = {1, 1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P, gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FFTCONV_ERR_CUDA, "cuTensorMapEncodeTiled (P store) failed (%d)", (int)r);
    return 0;
}
static int os_make_p_tensor_map5(const float* P, int RS, int RSP, unsigned long long nblk, unsigned vbox, OsTensorMap* out) {
    static PFN_tmapEncodeTiled encode = nullptr;   // P as {RS floats, 128 templates, 64 v, 33 u, template block x tile block}, box {8, 1, vbox, 33, 1}
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return fail(FFTCONV_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = (PFN_tmapEncodeTiled)fn;
    }
    static_assert(sizeof(CUtensorMap) == sizeof(OsTensorMap), "tensor map size");
    const cuuint64_t row = (cuuint64_t)RSP * 4;
    const cuuint64_t gdim[5] = {(cuuint64_t)RS, OS_TM, OS_T, OS_CH, (cuuint64_t)nblk};
    const cuuint64_t gstride[4] = {row, row * OS_TM, row * OS_TM * OS_T, row * OS_TM * OS_NBIN};
    const cuuint32_t box[5] = {8, 1, vbox, OS_CH, 1}, estride[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(P), gdim,
                              gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);   // (L2 promotion 64/128/256 B: no effect measured)
    if (r != CUDA_SUCCESS) return fail(FFTCONV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// Source plane -> B operand images (once per call).  If d_spec is given the plane is first recovered from
// the compat spectrum (inverse w, then C2R along h); otherwise the raw data [F][W][H] is tiled directly.
struct OsProv {                      // see Ctx::SpecCache
    const unsigned long long* hsel = nullptr;
    SrcDesc alt{};
    int alt_done = 0;
};
static unsigned os_hash_grid(const Ctx& c, size_t n) { return (unsigned)std::min<size_t>((n + 2047) / 2048, (size_t)c.sm_count * 4); }
static int os_prepare_data(Ctx& c, const OsCfg& g, const cpx* d_spec, const float* d_raw, int rawH, int rawW,
                           int correlate, cudaStream_t st, const OsProv* prov = nullptr) {
    const int F = g.F, FH = g.FH, FW = g.FW, CH = FH / 2 + 1;
    SrcDesc src;
    if (d_raw) {
        src.ptr = d_raw; src.rows = rawH; src.cols = rawW;
    } else {
        const cpx *twH, *twW;
        if (int e = get_twiddles(c, FH, st, &twH)) return e;
        if (int e = get_twiddles(c, FW, st, &twW)) return e;
        const LinePlan pH = make_line_plan(FH), pW = make_line_plan(FW);
        const int ldH = odd_ld(FH), ldW = odd_ld(FW);
        if (int e = dev_reserve(c.osZ, sizeof(cpx) * (size_t)F * FW * CH)) return e;
        if (int e = dev_reserve(c.osPlane, sizeof(float) * (size_t)F * FW * FH + sizeof(float*) * (size_t)F)) return e;
        float* plane = (float*)c.osPlane.p;
        float** d_planes = reinterpret_cast<float**>(plane + (size_t)F * FW * FH);
        float** h_planes;
        if (int e = pinned_get(c, sizeof(float*) * (size_t)F, (void**)&h_planes)) return e;
        for (int f = 0; f < F; ++f) h_planes[f] = plane + (size_t)f * FW * FH;
        CU(cudaMemcpyAsync(d_planes, h_planes, sizeof(float*) * (size_t)F, cudaMemcpyHostToDevice, st));
        if (int e = pinned_done(c, st)) return e;
        ProfScope ps(PK_OS_PLANE, st);
        const unsigned long long* skip = prov ? prov->hsel : nullptr;
        if (skip) {                                                  // hash of the spectrum the caller handed in -> hsel[1]
            unsigned long long* h1 = const_cast<unsigned long long*>(skip) + 1;
            const size_t n = (size_t)F * FW * CH;
            CU(cudaMemsetAsync(h1, 0, sizeof(unsigned long long), st));
            os_hash64<<<os_hash_grid(c, n), 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_spec), n, h1);
            LAUNCH_CHECK();
        }
        int TU = (int)((96 * 1024) / (2 * (size_t)ldW * sizeof(cpx)));
        TU = TU < 1 ? 1 : (TU > 16 ? 16 : TU);
        dim3 g2((CH + TU - 1) / TU, F);
        inv_w_pass<<<g2, 256, 2 * (size_t)TU * ldW * sizeof(cpx), st>>>(d_spec, FW, CH, pW, twW, (cpx*)c.osZ.p, TU, ldW, skip);
        LAUNCH_CHECK();
        const int NL = pick_lines(FH, 8);
        const long long nlines = (long long)F * (FW / 2);
        inv_h_pass<<<(unsigned)((nlines + NL - 1) / NL), 256, 2 * (size_t)NL * ldH * sizeof(cpx), st>>>(
            (const cpx*)c.osZ.p, F, FH, FW, CH, pH, twH, 1.0f / ((float)FW * (float)FH), d_planes, FH, FW, FH, NL, ldH, skip);
        LAUNCH_CHECK();
        src.ptr = plane; src.rows = FH; src.cols = FW;
    }
    if (int e = dev_reserve(c.osB, (size_t)g.NNB * OS_NBIN * g.b_buf)) return e;
    {
        OsDArgs a{};
        a.src = src; a.F = F; a.nth = g.nth; a.NTimg = g.NTimg; a.Sh = g.Sh; a.Sw = g.Sw; a.oy0 = g.maxkh - 1; a.ox0 = g.maxkw - 1;
        a.FH = FH; a.FW = FW; a.img = (float*)c.osB.p; a.NKS = g.NKS; a.KC = g.KC; a.NMMA = g.NMMA; a.NTn = g.NTn;
        a.correlate = correlate;
        a.st256 = os_env().data_st256;
        if (prov && !d_raw) { a.hsel = prov->hsel; a.alt = prov->alt; a.alt_done = prov->alt_done; }
        ++c.osB_gen;
        // channel pair fastest: the CTAs resident at any time complete whole (tile block, bin) blocks of the B image
        // together (tile-fastest order left every 16-byte row pair of a block to be written at 16 different times)
        const unsigned grid = (unsigned)g.NT * (unsigned)(g.NKS * g.KC);
        ProfScope ps(PK_OS_DATA, st);
        if (os_env().data_occ) os_data_fft_occ<<<grid, 128, OS_DATA_SMEM_OCC, st>>>(a);
        else os_data_fft<<<grid, 128, OS_DATA_SMEM, st>>>(a);
        LAUNCH_CHECK();
    }
    return 0;
}

static size_t os_kern_smem(int NF) {
    return (size_t)32 * (17 * 16 * NF + 2) * sizeof(cpx) + (NF == 1 ? (size_t)32 * (16 * 18 + 2) * sizeof(float) : 0);   // + staged raw planes
}

// Templates per chunk.  Device outputs: as large as the scratch budget allows (fewer launch tails, the B images are
// streamed once per chunk).  Host outputs: small chunks, so that the D2H of one chunk overlaps the next one's compute.
static int os_max_chunk(const OsCfg& g, bool out_on_device) {
    int ntblk = os_env().ntblk;
    if (ntblk <= 0) {
        ntblk = out_on_device ? 8 : 2;
        const size_t per_blk = (size_t)OS_NBIN * g.NKS * (g.a_stage / 2) + (size_t)g.NNB * OS_NBIN * g.p_blk;   // A + P
        while (ntblk > 1 && per_blk * ntblk > ((size_t)6 << 30)) ntblk >>= 1;
    }
    return ntblk * OS_TM;
}

static int os_reserve_chunk(Ctx& c, const OsCfg& g, int KC_templates, bool need_A = true) {
    const int ntblk = (KC_templates + OS_TM - 1) / OS_TM;
    if (need_A) if (int e = dev_reserve(c.osA, (size_t)ntblk * OS_NBIN * g.NKS * (g.a_stage / 2))) return e;     // fp32 images
    if (int e = dev_reserve(c.osP, (size_t)ntblk * g.NNB * OS_NBIN * g.p_blk)) return e;
    return 0;
}

// fused detection (fftconv_bank_conv_detect / _topk): per-template candidate keys, counters, thresholds, biases (device)
struct OsDetect {
    int mode = 0;                  // 0 off, 1 threshold, 3 top-k (candidate pass + threshold pass, both over the same P)
    int cap = 0, k = 0;
    unsigned long long* keys = nullptr;    // [K][cap]
    unsigned int* count = nullptr;         // [K]
    float* thr = nullptr;                  // [K]
    const float* bias = nullptr;           // [K] or nullptr
};
static int os_chunk_inverse(Ctx& c, const OsCfg& g, OsInvArgs a, int nk, cudaStream_t st);
// Template transforms of the NEXT chunk, run ahead of it: os_kern_fft(i + 1) only needs the A scratch, which is free as soon
// as os_gemm(i) has finished -- so it is enqueued on a side stream behind os_gemm(i) and shares the SMs with os_inverse_z(i)
// (one writes A at HBM speed, the other is bound by instruction issue and the wait for its boxes), and os_gemm(i + 1)
// finds its A images ready.  Only for templates that are already resident on the device.
struct OsAhead {
    bool kern_done = false;              // this chunk's A images were produced by the previous chunk's run-ahead
    const SrcDesc* next_descs = nullptr; // run ahead for the next chunk (nullptr: none)
    int next_nk = 0;
    cudaStream_t side = nullptr;
};
static int os_launch_kern_fft(Ctx& c, const OsCfg& g, const SrcDesc* d_descs, int nk, const fftconv_options& opt, cudaStream_t st) {
    const int ntblk = (nk + OS_TM - 1) / OS_TM;
    OsKArgs a{};
    a.descs = d_descs; a.nk = nk; a.F = g.F; a.img = (float*)c.osA.p; a.NKS = g.NKS; a.KC = g.KC;
    a.flip = opt.correlate ? 1 : 0;
    dim3 grid(ntblk * OS_TM / OS_KSL, g.NKS * g.KC);
    ProfScope ps(PK_OS_KERN, st);
    if (g.NFK == 1) os_kern_fft<1><<<grid, 256, os_kern_smem(1), st>>>(a);
    else os_kern_fft<2><<<grid, 256, os_kern_smem(2), st>>>(a);
    LAUNCH_CHECK();
    return 0;
}
static int os_chunk_detect(Ctx& c, const OsCfg& g, OsInvArgs a, int nk, const OsDetect& det, int k0, cudaStream_t st);
static int os_chunk(Ctx& c, const OsCfg& g, const SrcDesc* d_descs, int nk, float* const* d_outptrs,
                    const fftconv_options& opt, cudaStream_t st, int out_img_stride = 0, const float* bankA = nullptr,
                    unsigned long long* peak_keys = nullptr, const int2* khw = nullptr, int H = 0, int W = 0,
                    cudaEvent_t data_ready = nullptr, const OsDetect* det = nullptr, int det_k0 = 0,
                    const OsAhead* ahead = nullptr) {
    const int ntblk = (nk + OS_TM - 1) / OS_TM;
    if (!bankA && ahead && ahead->kern_done) {
        CU(cudaStreamWaitEvent(st, c.evf[3], 0));                    // the A images of this chunk were transformed ahead
    } else if (!bankA) {
        if (int e = os_launch_kern_fft(c, g, d_descs, nk, opt, st)) return e;
    }
    if (data_ready) CU(cudaStreamWaitEvent(st, data_ready, 0));      // B images of this call are complete
    {
        OsGemmArgs a{};
        a.Aimg = bankA ? bankA : (const float*)c.osA.p; a.Bimg = (const float*)c.osB.p; a.P = (float*)c.osP.p;
        a.NTBLK = ntblk; a.NNB = g.NNB; a.NKS = g.NKS; a.KC = g.KC; a.NMMA = g.NMMA; a.RS = g.RS; a.RSP = g.RSP;
        a.nitems = (long long)g.NNB * OS_NBIN * ntblk;
        a.nsta = g.nsta;
        if (const char* v = getenv("FFTCONV_OS_NSTA")) a.nsta = std::max(2, std::min(g.nsta, atoi(v)));   // timing experiments
        a.lbo_swap = os_env().lbo_swap;
        a.hi_inplace = os_env().hi_inplace;
        a.dbg = os_env().dbg;
        a.pf_items = os_env().pf;
        ProfScope ps(PK_OS_GEMM, st);
        if (os_env().gemm_simt) {                       // validation only, never the default
            os_gemm_simt<<<(unsigned)a.nitems, 128, 0, st>>>(a);
        } else {
            a.use_tmap = os_env().gemm_tmap;
            OsTensorMap pmap{};
            static const int use_pmap = []{ const char* v = getenv("FFTCONV_OS_GEMM_PMAP"); return v && *v ? atoi(v) : 1; }();
            a.use_pmap = (a.use_tmap && use_pmap && g.RSP != g.RS && os_make_p_store_map(a.P, g.RS, g.RSP, (unsigned long long)a.nitems, &pmap) == 0) ? 1 : 0;
            if (!a.use_pmap) { a.RSP = g.RS; g_err.clear(); }        // dense staging tile, plain bulk store
            const unsigned grid = (unsigned)std::min<long long>(a.nitems, (long long)c.sm_count);
            os_gemm<<<grid, 320, g.gemm_smem, st>>>(a, pmap);
        }
        LAUNCH_CHECK();
    }
    if (ahead && ahead->next_descs && !bankA) {
        CU(cudaEventRecord(c.evf[2], st));                           // os_gemm(i) done: the A scratch is free
        CU(cudaStreamWaitEvent(ahead->side, c.evf[2], 0));
        if (int e = os_launch_kern_fft(c, g, ahead->next_descs, ahead->next_nk, opt, ahead->side)) return e;
        CU(cudaEventRecord(c.evf[3], ahead->side));                  // joined by the next chunk before its os_gemm
    }
    {
        OsInvArgs a{};
        a.levels = g.d_levels; a.nlevels = g.nlevels;
        a.P = (const float*)c.osP.p; a.outs = d_outptrs; a.nk = nk; a.NNB = g.NNB; a.NTn = g.NTn; a.RS = g.RS;
        a.NT = g.NT; a.NTimg = g.NTimg; a.nth = g.nth; a.Sh = g.Sh; a.Sw = g.Sw; a.oy0 = g.maxkh - 1; a.ox0 = g.maxkw - 1;
        a.FH = g.FH; a.FW = g.FW; a.out_img_stride = out_img_stride;
        a.peak_keys = peak_keys; a.khw = khw; a.H = H; a.W = W;
        a.corr = (opt.correlate && !peak_keys) ? 1 : 0;
        a.dbg = os_env().dbg;
        a.crop_h = opt.crop_h > 0 ? opt.crop_h : g.FH;
        a.crop_w = opt.crop_w > 0 ? opt.crop_w : g.FW;
        a.out_ld = opt.out_ld > 0 ? opt.out_ld : a.crop_h;
        if (det && det->mode) return os_chunk_detect(c, g, a, nk, *det, det_k0, st);
        return os_chunk_inverse(c, g, a, nk, st);
    }
}

static int os_chunk_inverse(Ctx& c, const OsCfg& g, OsInvArgs a, int nk, cudaStream_t st) {
    ProfScope ps(PK_OS_INV, st);
    OsTensorMap tm;
    const int ntb = (nk + OS_TM - 1) / OS_TM;
    // (a driver without cuTensorMapEncodeTiled leaves the per-thread cp.async gather of os_inverse)
    if (os_env().inv_tma && os_make_p_tensor_map(a.P, g.RS, g.RS, (unsigned long long)ntb * g.NNB * OS_NBIN, &tm) == 0) {
        const long long nitems = (long long)g.NNB * (g.RS / 8) * nk;      // (template, tile block, group of 4 tiles)
        static const int use_z = []{ const char* v = getenv("FFTCONV_OS_INV_Z"); return v && *v ? atoi(v) : 1; }();
        OsTensorMap tm8, tm1;
        if (use_z && os_make_p_tensor_map5(a.P, g.RS, g.RS, (unsigned long long)ntb * g.NNB, 8, &tm8) == 0 &&
            os_make_p_tensor_map5(a.P, g.RS, g.RS, (unsigned long long)ntb * g.NNB, 1, &tm1) == 0) {
            // One CTA per item by default.  FFTCONV_OS_INV_PERSIST=1 runs 3 persistent CTAs per SM that request the boxes of
            // their next item while they store the current one: the wait for the boxes disappears (5 % of the stall samples
            // instead of 34 %), but with all 24 warps of an SM busy the 38 KB of live code thrash the 32 KB instruction cache
            // (`no_instruction` becomes the top stall) and the tile groups of a template drift apart in time (DRAM reads
            // 646 MB for the 608 MB of P): 0.254 ms against 0.231 ms at config 2 (profiles/r02_inverse_ab.md).
            static const int persist = []{ const char* v = getenv("FFTCONV_OS_INV_PERSIST"); return v && *v ? atoi(v) : 0; }();
            const unsigned grid = (unsigned)(persist ? std::min<long long>(nitems, 3LL * c.sm_count) : nitems);
            os_inverse_z<<<grid, 256, OS_IZ_SMEM, st>>>(a, tm8, tm1, (int)nitems);
        } else {
            if (a.levels) return fail(FFTCONV_ERR_UNSUPPORTED, "pyramid batch: the zone inverse (os_inverse_z) is not available");
            const dim3 g1(g.NNB * (g.RS / 8), nk);
            if (use_z == 0 && os_env().dbg & 128) os_inverse_tma<0><<<g1, 256, OS_ITMA_SMEM, st>>>(a, tm);
            else os_inverse_tma<6><<<g1, 256, OS_ITMA_SMEM, st>>>(a, tm);
        }
    } else {
        if (a.levels) return fail(FFTCONV_ERR_UNSUPPORTED, "pyramid batch: no tensor maps on this driver");
        dim3 grid((g.NT + OS_IG - 1) / OS_IG, nk);
        os_inverse<<<grid, OS_IG * 64, g.inv_smem, st>>>(a);
    }
    LAUNCH_CHECK();
    (void)c;
    return 0;
}

// Detection over the product spectra of one chunk (templates k0 .. k0 + nk - 1 of the call).
static int os_chunk_detect(Ctx& c, const OsCfg& g, OsInvArgs a, int nk, const OsDetect& det, int k0, cudaStream_t st) {
    a.outs = nullptr; a.peak_keys = nullptr; a.corr = 0;
    a.det_cap = det.cap;
    a.det_keys = det.keys + (size_t)k0 * det.cap;
    a.det_count = det.count + k0;
    a.det_thr = det.thr + k0;
    a.det_bias = det.bias ? det.bias + k0 : nullptr;
    if (det.mode == 3) {
        // pass 1: one candidate per lane (its own best response); the k-th largest candidate bounds the k-th largest
        // response from below and becomes the template's threshold for pass 2
        CU(cudaMemsetAsync(a.det_keys, 0, sizeof(unsigned long long) * (size_t)nk * det.cap, st));
        a.det_mode = 2;
        if (int e = os_chunk_inverse(c, g, a, nk, st)) return e;
        os_det_select<<<nk, 256, 0, st>>>(a.det_keys, nullptr, det.cap, det.k, nullptr, det.thr + k0, nullptr);
        LAUNCH_CHECK();
    }
    CU(cudaMemsetAsync(a.det_count, 0, sizeof(unsigned int) * (size_t)nk, st));
    a.det_mode = 1;
    return os_chunk_inverse(c, g, a, nk, st);
}

// ------------------------------------------------------------------ conv core
struct KernelRef {
    const float* ptr;
    int kh, kw;
    bool on_device;
};

struct ConvArgs {
    const cpx* d_spec;
    int CH, FW, F, K;
    const KernelRef* kernels;      // K entries
    float* const* outs;            // K entries (host or device pointers)
    bool out_on_device;
    fftconv_options opt;
    const float* d_raw = nullptr;  // one-shot entry point: the raw data [F][rawW][rawH] on the device
    int rawH = 0, rawW = 0;
    const float* bankA = nullptr;  // prepared bank (fftconv_bank_*): A operand images of all K templates, path 3 only
    int bank_maxkh = 0, bank_maxkw = 0;
    unsigned long long* peak_keys = nullptr;   // fused maximum (fftconv_bank_conv_max): K packed keys, no planes
    const int2* bank_khw = nullptr;            // (kh, kw) of every template of the bank (device)
    OsDetect det;                  // fused detection (no planes)
    cudaEvent_t spec_ready = nullptr;   // one-shot event handed over by fftconv_spectrum_ready_event (consumed by conv_impl)
    int nimg = 1;                  // batched entry point: nimg images [nimg][F][rawW][rawH] (raw, device, overlap-save path
                                   // only); outs then holds nimg*K device planes, image-major
};

enum ConvPath { PATH_AUTO = 0, PATH_GENERIC = 1, PATH_TILE16 = 2, PATH_OSGEMM = 3, PATH_BIGPLANE = 4 };

// Which pipeline serves this call.  The overlap-save / tensor-core path needs a bank large enough to fill
// 128-row MMA blocks; small banks stay on the SIMT pipelines.
static bool bigplane_supported(int FH, int FW, int F);
static int choose_path(const fftconv_options& opt, int F, int FH, int FW, int maxkh, int maxkw, int K) {
    OsCfg g;
    const bool os_ok = os_config(F, FH, FW, maxkh, maxkw, g);
    const bool t16_ok = tile16_supported(FH, FW, maxkh, maxkw);
    if (opt.force_generic || opt.path == PATH_GENERIC) return PATH_GENERIC;
    if (opt.path == PATH_OSGEMM) return os_ok ? PATH_OSGEMM : (t16_ok ? PATH_TILE16 : PATH_GENERIC);
    if (opt.path == PATH_TILE16) return t16_ok ? PATH_TILE16 : PATH_GENERIC;
    const bool bp_ok = bigplane_supported(FH, FW, F);
    if (opt.path == PATH_BIGPLANE) return bp_ok ? PATH_BIGPLANE : PATH_GENERIC;
    // very large planes with small templates: the tile spectra (B) and the product spectra of ONE template block (P) must
    // stay within a sane scratch budget, otherwise the plane goes to the large-plane / generic pipelines
    const bool os_fits = os_ok && (size_t)g.NNB * OS_NBIN * (g.b_buf + g.p_blk) <= ((size_t)24 << 30);
    if (os_fits && K >= os_env().min_k) return PATH_OSGEMM;
    if (t16_ok) return PATH_TILE16;
    // long lines: the in-place pipeline touches whole sectors in the strided pass (kernels_bigplane.cuh)
    return (bp_ok && FH >= 1024 && FW >= 1024) ? PATH_BIGPLANE : PATH_GENERIC;
}

static int conv_generic_chunk(Ctx& c, const ConvArgs& a, int FH, const SrcDesc* d_descs, const int* d_kcols,
                              int nk, int maxcols, float* const* d_outptrs, cudaStream_t st) {
    const int CH = a.CH, FW = a.FW, F = a.F;
    const cpx *twH, *twW;
    if (int e = get_twiddles(c, FH, st, &twH)) return e;
    if (int e = get_twiddles(c, FW, st, &twW)) return e;
    const LinePlan pH = make_line_plan(FH), pW = make_line_plan(FW);
    const int ldH = odd_ld(FH), ldW = odd_ld(FW);

    // 1. kernel half transforms (pad fused)
    {
        const int NL = pick_lines(FH, 8);
        const long long nlines = (long long)nk * F * ((maxcols + 1) / 2);
        const unsigned grid = (unsigned)((nlines + NL - 1) / NL);
        ProfScope ps(PK_GEN_H, st);
        fwd_h_pass<PAD_ZERO><<<grid, 256, 2 * (size_t)NL * ldH * sizeof(cpx), st>>>(
            d_descs, nk, F, maxcols, FH, CH, pH, twH, (cpx*)c.T.p, NL, ldH, 0, 0, maxcols);
        LAUNCH_CHECK();
    }
    // 2. w pass: forward, multiply-accumulate over channels, one inverse
    {
        int TU = (int)((160 * 1024) / (3 * (size_t)ldW * sizeof(cpx)));
        TU = TU < 1 ? 1 : (TU > 8 ? 8 : TU);
        dim3 grid((CH + TU - 1) / TU, nk);
        const size_t smem = 3 * (size_t)TU * ldW * sizeof(cpx);
        ProfScope ps(PK_GEN_W, st);
        if (a.opt.correlate)
            conv_w_pass_generic<true><<<grid, 256, smem, st>>>((const cpx*)c.T.p, d_kcols, maxcols, a.d_spec, F, FW, CH,
                                                               pW, twW, (cpx*)c.Z.p, TU, ldW);
        else
            conv_w_pass_generic<false><<<grid, 256, smem, st>>>((const cpx*)c.T.p, d_kcols, maxcols, a.d_spec, F, FW, CH,
                                                                pW, twW, (cpx*)c.Z.p, TU, ldW);
        LAUNCH_CHECK();
    }
    // 3. C2R along h + scale + (crop) store
    {
        const int NL = pick_lines(FH, 8);
        const long long nlines = (long long)nk * (FW / 2);
        const unsigned grid = (unsigned)((nlines + NL - 1) / NL);
        const int crop_h = a.opt.crop_h > 0 ? a.opt.crop_h : FH;
        const int crop_w = a.opt.crop_w > 0 ? a.opt.crop_w : FW;
        const int out_ld = a.opt.out_ld > 0 ? a.opt.out_ld : crop_h;
        ProfScope ps(PK_GEN_C2R, st);
        inv_h_pass<<<grid, 256, 2 * (size_t)NL * ldH * sizeof(cpx), st>>>(
            (const cpx*)c.Z.p, nk, FH, FW, CH, pH, twH, 1.0f / ((float)FW * (float)FH), d_outptrs, crop_h, crop_w,
            out_ld, NL, ldH, nullptr);
        LAUNCH_CHECK();
    }
    return 0;
}


// ------------------------------------------------------------------ large-plane path (kernels_bigplane.cuh), host side
// In-place plan: odd radices first (their stride is the longest, so a zero-padded template prunes the first stage),
// then the power of two split into radices 8 / 16 / 32.
// Size-specialised variants (kernels_bigplane_ct.cuh): X(line length, variant, threads and lines per CTA of bp_conv_w_ct,
// threads and complex lines per CTA of bp_inv_h_ct, FLAGS of both kernels, radices...).  Variant 0 of a length is the product; further variants are kept for A/B runs (FFTCONV_BP_CT=v).
// The run-time plan of a listed length is built from the SAME radices, so both kernel families share the digit order.
// Measured at config 3 (4608, 16 templates, B200): run-time plan 2.79 + 1.54 ms (bp_conv_w + bp_inv_h) -> 1.16 + 0.80 ms.
// Alternatives tried at 4608: radices [9,16,32] 1.62 ms and [9,8,8,8] 1.27 ms against 1.22 ms (bp_conv_w); 512 instead
// of 576 threads 1.28 ms; 2-line CTAs, two per SM 1.75 ms (half sectors); bp_inv_h with 2 lines x 288 threads, two CTAs per
// SM 0.88 ms against 0.80 ms (1 line x 160 threads, four per SM); FLAGS 0 -> 7: 1.22 -> 1.16 and 0.91 -> 0.88 ms.
#define BP_CT_LIST(X) \
    X(1152, 0, 288, 4, 160, 1, 7, 9, 8, 16) \
    X(2048, 0, 512, 4, 128, 1, 7, 8, 16, 16) \
    X(4096, 0, 512, 4, 128, 1, 7, 16, 16, 16) \
    X(4608, 0, 576, 4, 160, 1, 7, 9, 32, 16)

struct BpCtEntry { int n, var, ns; int R[BP_MAX_STAGES]; };
static const BpCtEntry* bp_ct_find(int n) {
#define BP_CT_ROW(N, V, NTW, TUW, NTH, NLH, FL, ...) {N, V, (int)(sizeof((int[]){__VA_ARGS__}) / sizeof(int)), {__VA_ARGS__}},
    static const BpCtEntry tab[] = {BP_CT_LIST(BP_CT_ROW)};
#undef BP_CT_ROW
    const int want = os_env().bp_ct;
    if (want < 0) return nullptr;
    const BpCtEntry* first = nullptr;
    for (const BpCtEntry& e : tab) {
        if (e.n != n) continue;
        if (e.var == want) return &e;
        if (e.var == 0) first = &e;
    }
    return first;
}

template <class P, int NTW, int TUW, int FL>
static int bp_ct_conv_w_launch(bool conj, bool multi, int nk, const cpx* T, const int* kcols, int maxcols4, const cpx* Sp, int F,
                               int CHp, const cpx* tw, cpx* Z, cudaStream_t st) {
    // single channel: TUW lines per CTA (4: whole sectors, one CTA per SM; 2: two CTAs per SM, the halves of a sector meet
    // in L2); several channels: 2 lines + 2 accumulator lines
    const bool two = multi || TUW == 2;
    const dim3 grid(two ? 2 * nk : nk, CHp / 4);
    constexpr size_t smem4 = (4 * (size_t)P::template ld<4>() + P::TWN) * sizeof(cpx);
    constexpr size_t smem2m = (4 * (size_t)P::template ld<2>() + P::TWN) * sizeof(cpx);
    constexpr size_t smem2 = (2 * (size_t)P::template ld<2>() + P::TWN) * sizeof(cpx);
    static_assert(smem4 <= kMaxSmem && smem2m <= kMaxSmem, "4 lines + twiddles must fit shared memory");
    constexpr int NTM = NTW * 2 / TUW;           // threads of the 2-line multi-channel CTA
#define BP_CT_GO(CONJ, MULTI, NT, TU, MINB, SM) do { \
        auto kern = bp_conv_w_ct<P, CONJ, MULTI, NT, TU, MINB, FL>; \
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM))); \
        kern<<<grid, NT, SM, st>>>(T, kcols, maxcols4, Sp, F, CHp, tw, Z); } while (0)
    if (multi) { if (conj) BP_CT_GO(true, true, NTM, 2, 1, smem2m); else BP_CT_GO(false, true, NTM, 2, 1, smem2m); }
    else if constexpr (TUW == 2) { if (conj) BP_CT_GO(true, false, NTW, 2, 2, smem2); else BP_CT_GO(false, false, NTW, 2, 2, smem2); }
    else { if (conj) BP_CT_GO(true, false, NTW, 4, 1, smem4); else BP_CT_GO(false, false, NTW, 4, 1, smem4); }
#undef BP_CT_GO
    return 0;
}
template <class P, int NT, int NLH, int FL>
static int bp_ct_inv_h_launch(int nk, const cpx* Z, int FW, int CH, int CHp, const cpx* tw, const unsigned short* pos_of,
                              float* const* outs, int crop_h, int crop_w, int out_ld, cudaStream_t st) {
    // NLH complex lines = 2 NLH real columns per CTA; as many CTAs per SM as shared memory holds (their load, transform and
    // store phases overlap)
    const dim3 grid(FW / (2 * NLH), nk);
    constexpr size_t smem = (NLH * (size_t)P::template ld<NLH>() + P::TWN) * sizeof(cpx);
    constexpr int MINB = (int)(kMaxSmem / (smem + 1024)) > 4 ? 4 : (int)(kMaxSmem / (smem + 1024));
    auto kern = bp_inv_h_ct<P, NT, NLH, MINB, FL>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(Z, FW, CH, CHp, tw, pos_of, outs, crop_h, crop_w, out_ld);
    return 0;
}
// returns -1 when no specialised kernel serves this length
static int bp_ct_conv_w(const BpCtEntry* e, bool conj, bool multi, int nk, const cpx* T, const int* kcols, int maxcols4,
                        const cpx* Sp, int F, int CHp, const cpx* tw, cpx* Z, cudaStream_t st) {
    if (!e) return -1;
#define BP_CT_CASE(N, V, NTW, TUW, NTH, NLH, FL, ...) \
    if (e->n == N && e->var == V) return bp_ct_conv_w_launch<CtPlan<N, __VA_ARGS__>, NTW, TUW, FL>(conj, multi, nk, T, kcols, maxcols4, Sp, F, CHp, tw, Z, st);
    BP_CT_LIST(BP_CT_CASE)
#undef BP_CT_CASE
    return -1;
}
static int bp_ct_inv_h(const BpCtEntry* e, int nk, const cpx* Z, int FW, int CH, int CHp, const cpx* tw,
                       const unsigned short* pos_of, float* const* outs, int crop_h, int crop_w, int out_ld, cudaStream_t st) {
    if (!e) return -1;
#define BP_CT_CASE(N, V, NTW, TUW, NTH, NLH, FL, ...) \
    if (e->n == N && e->var == V) return bp_ct_inv_h_launch<CtPlan<N, __VA_ARGS__>, NTH, NLH, FL>(nk, Z, FW, CH, CHp, tw, pos_of, outs, crop_h, crop_w, out_ld, st);
    BP_CT_LIST(BP_CT_CASE)
#undef BP_CT_CASE
    return -1;
}

static bool make_ip_plan(int n, IpPlan& p) {
    p = IpPlan{};
    p.n = n;
    if (n < 16 || n % 16 || n >= 65536) return false;
    int L = n;
    auto push = [&](int r) -> bool {
        if (p.ns >= BP_MAX_STAGES) return false;
        p.R[p.ns] = r; p.L[p.ns] = L; ++p.ns; L /= r;
        return true;
    };
    if (const BpCtEntry* e = bp_ct_find(n)) {       // a size-specialised variant exists: same radices, same digit order
        for (int i = 0; i < e->ns; ++i) if (!push(e->R[i])) return false;
        const int m0 = n / p.R[0];
        p.magic = m0 >= 32 ? (unsigned)((0x100000000ull + m0 - 1) / m0) : 0u;
        return L == 1;
    }
    int o = n, e = 0;
    while ((o & 1) == 0) { o >>= 1; ++e; }        // n = o * 2^e, e >= 4
    const int odd[] = {9, 17, 13, 11, 7, 5, 3};
    for (int r : odd)
        while (o % r == 0) { if (!push(r)) return false; o /= r; }
    if (o != 1) return false;                     // a prime factor above 17: stay on the generic path
    const int nst = (e + 4) / 5;                  // as few power-of-two stages as radices <= 32 allow, evenly split
                                                  // (measured at 4608 = 9 * 512: [9, 32, 16] 4.7 ms, [9, 8, 8, 8] 6.2 ms per 16 templates)
    for (int i = 0; i < nst; ++i) {
        const int bits = e / nst + (i < e % nst ? 1 : 0);
        if (!push(1 << bits)) return false;
    }
    const int m0 = n / p.R[0];
    p.magic = m0 >= 32 ? (unsigned)((0x100000000ull + m0 - 1) / m0) : 0u;
    return L == 1;
}

static inline int bp_ldl(int n) { return ((n + n / 16 + 40 + 15) / 16) * 16 + 4; }
static const size_t kBpTwBytes = BP_TW_SMEM_MAX * sizeof(cpx);   // shared twiddle table behind the lines   // == 4 (mod 16): the 4 lines of a tile hit disjoint banks

static bool bigplane_supported(int FH, int FW, int F) {
    IpPlan a, b;
    if (!make_ip_plan(FH, a) || !make_ip_plan(FW, b)) return false;
    const size_t conv_smem = 4 * (size_t)bp_ldl(FW) * sizeof(cpx) + kBpTwBytes;
    const size_t h_smem = 2 * (size_t)bp_ldl(FH) * sizeof(cpx) + kBpTwBytes;
    (void)F;
    return conv_smem <= kMaxSmem && h_smem <= kMaxSmem;
}

// pos_of[v]: position of natural index v in the digit-reversed order the forward stages leave; nat_of = its inverse
static int get_ip_tables(Ctx& c, int n, const IpPlan& P, const unsigned short** pos_of, const unsigned short** nat_of) {
    auto it = c.ipmap.find(n);
    if (it == c.ipmap.end()) {
        std::vector<unsigned short> h(2 * (size_t)n);
        for (int pos = 0; pos < n; ++pos) {
            int rem = pos, v = 0, mul = 1;
            for (int s = 0; s < P.ns; ++s) {
                const int m = P.L[s] / P.R[s];
                const int q = rem / m;
                rem -= q * m;
                v += q * mul;
                mul *= P.R[s];
            }
            h[(size_t)v] = (unsigned short)pos;
            h[(size_t)n + pos] = (unsigned short)v;
        }
        unsigned short* d = nullptr;
        CU(cudaMalloc(&d, sizeof(unsigned short) * 2 * (size_t)n));
        CU(cudaMemcpy(d, h.data(), sizeof(unsigned short) * 2 * (size_t)n, cudaMemcpyHostToDevice));
        it = c.ipmap.emplace(n, d).first;
    }
    *pos_of = it->second;
    *nat_of = it->second + n;
    return 0;
}

static inline int bp_chp(int CH) { return (CH + 3) & ~3; }

// once per call: compat spectrum -> re-padded, pre-scaled private copy
static int bigplane_prepare(Ctx& c, const cpx* d_spec, int CH, int FW, int F, int FH, cudaStream_t st) {
    const int CHp = bp_chp(CH);
    IpPlan pW;
    make_ip_plan(FW, pW);
    const unsigned short *posW, *natW;
    if (int e = get_ip_tables(c, FW, pW, &posW, &natW)) return e;
    if (int e = dev_reserve(c.bpS, sizeof(cpx) * (size_t)F * FW * CHp)) return e;
    const long long rows = (long long)F * FW;
    ProfScope ps(PK_BP_REPAD, st);
    bp_repad_spec<<<c.sm_count * 8, 256, 0, st>>>(d_spec, CH, CHp, FW, rows, 1.0f / ((float)FW * (float)FH), natW, (cpx*)c.bpS.p);
    LAUNCH_CHECK();
    return 0;
}

static int conv_bigplane_chunk(Ctx& c, const ConvArgs& a, int FH, const SrcDesc* d_descs, const int* d_kcols,
                               int nk, int maxcols, float* const* d_outptrs, cudaStream_t st) {
    const int CH = a.CH, FW = a.FW, F = a.F, CHp = bp_chp(CH);
    const int maxcols4 = (maxcols + 3) & ~3;
    IpPlan pH, pW;
    make_ip_plan(FH, pH); make_ip_plan(FW, pW);
    pH.tws = 0; pW.tws = 1;      // shared-memory twiddle table for the later stages of the w pass (neutral for the h passes)
    const cpx *twH, *twW;
    const unsigned short *posH, *natH, *posW, *natW;
    if (int e = get_twiddles(c, FH, st, &twH)) return e;
    if (int e = get_twiddles(c, FW, st, &twW)) return e;
    if (int e = get_ip_tables(c, FH, pH, &posH, &natH)) return e;
    if (int e = get_ip_tables(c, FW, pW, &posW, &natW)) return e;
    const int ldH = bp_ldl(FH), ldW = bp_ldl(FW);
    const size_t smemH = 2 * (size_t)ldH * sizeof(cpx) + kBpTwBytes;
    {
        dim3 grid(maxcols4 / 4, nk * F);
        ProfScope ps(PK_BP_KERN_H, st);
        bp_kern_h<2><<<grid, 256, smemH, st>>>(d_descs, F, maxcols4, FH, CH, CHp, pH, twH, posH, (cpx*)c.T.p, ldH);
        LAUNCH_CHECK();
    }
    {
        // single channel: 4 lines (whole 32-byte sectors), product in place; several channels: 2 lines + 2 accumulator
        // lines.  512 threads, one CTA per SM (measured alternatives at config 3, per 16 templates: 256 threads 3.8 ms,
        // two 2-line CTAs per SM 3.5 ms, against 3.0 ms)
        const bool multi = F > 1;
        dim3 grid(multi ? 2 * nk : nk, CHp / 4);
        const size_t smem = 4 * (size_t)ldW * sizeof(cpx) + kBpTwBytes;
        ProfScope ps(PK_BP_CONV_W, st);
        const cpx* T = (const cpx*)c.T.p; const cpx* Sp = (const cpx*)c.bpS.p; cpx* Z = (cpx*)c.Z.p;
#define BP_LAUNCH(CONJ, MULTI, TU) bp_conv_w<CONJ, MULTI, 512, TU, 1><<<grid, 512, smem, st>>>(T, d_kcols, maxcols4, Sp, F, FW, CHp, pW, twW, Z, ldW)
        const int ct = bp_ct_conv_w(bp_ct_find(FW), a.opt.correlate != 0, multi, nk, T, d_kcols, maxcols4, Sp, F, CHp, twW, Z, st);
        if (ct > 0) return ct;
        if (ct == 0) { /* size-specialised kernel launched */ }
        else if (multi) { if (a.opt.correlate) BP_LAUNCH(true, true, 2); else BP_LAUNCH(false, true, 2); }
        else { if (a.opt.correlate) BP_LAUNCH(true, false, 4); else BP_LAUNCH(false, false, 4); }
#undef BP_LAUNCH
        LAUNCH_CHECK();
    }
    {
        const int crop_h = a.opt.crop_h > 0 ? a.opt.crop_h : FH;
        const int crop_w = a.opt.crop_w > 0 ? a.opt.crop_w : FW;
        const int out_ld = a.opt.out_ld > 0 ? a.opt.out_ld : crop_h;
        dim3 grid(FW / 4, nk);
        ProfScope ps(PK_BP_INV_H, st);
        const int ct = bp_ct_inv_h(bp_ct_find(FH), nk, (const cpx*)c.Z.p, FW, CH, CHp, twH, posH, d_outptrs, crop_h, crop_w, out_ld, st);
        if (ct > 0) return ct;
        if (ct < 0)
            bp_inv_h<2><<<grid, 256, smemH, st>>>((const cpx*)c.Z.p, FH, FW, CH, CHp, pH, twH, posH, d_outptrs, crop_h, crop_w, out_ld, ldH);
        LAUNCH_CHECK();
    }
    return 0;
}


// ------------------------------------------------------------------ peer spectrum (CUDA IPC + NVLink)
__device__ unsigned int g_peer_timeouts = 0;

__global__ void peer_signal_kernel(unsigned long long* flag, unsigned long long value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

// one thread per flag; gives up after ~2 s so that a lost peer cannot wedge the device
__global__ void peer_wait_kernel(const unsigned long long* flags, int n, unsigned long long value) {
    const int i = threadIdx.x;
    if (i >= n) return;
    const long long t0 = clock64();
    unsigned long long v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + i) : "memory");
        if (v >= value) break;
        if (clock64() - t0 > 4000000000ll) { atomicAdd(&g_peer_timeouts, 1u); break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

// grid-stride 16-byte copy out of the mapped peer buffer (NVLink reads), tail in bytes
__global__ void peer_pull_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16,
                                 unsigned char* __restrict__ dtail, const unsigned char* __restrict__ stail, int ntail) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) dtail[threadIdx.x] = stail[threadIdx.x];
}

// All-gather of the spectrum slices through the peer mappings, ONE launch: block group p of rank r waits for rank p's
// ready flag (through its NVLink mapping), pulls slice p out of rank p's buffer into the same offset of its own buffer and
// -- the last block of the group -- acknowledges in rank p's memory.  Every buffer has the same layout: [spectrum | ready
// flag | n acknowledgement slots].  No rank's NVLink egress carries more than (n - 1) slices.
struct PeerAgArgs {
    unsigned char* base[16];          // base[p]: rank p's buffer (own pointer for p == rank, mapping otherwise)
    unsigned long long off[17];       // slice boundaries in bytes (multiples of 16)
    unsigned long long flag_off;
    unsigned long long step;
    int n, rank, bpg;
};
__device__ unsigned int g_peer_ag_done[16];

// One warp waits for the ready flags of all peers (lane p <-> rank p).  Runs in front of peer_allgather_kernel on the same
// stream: the 128 copy CTAs then start on flags that are already up instead of spinning -- while they spun they held a third
// of the GPU's thread slots and slowed the template transforms running next to them down (2 GPUs: +0.06 ms per step).
__global__ void peer_ag_wait_kernel(PeerAgArgs a) {
    const int p = threadIdx.x;
    if (p >= a.n || p == a.rank) return;
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(a.base[p] + a.flag_off);
    const long long t0 = clock64();
    unsigned long long v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= a.step) break;
        if (clock64() - t0 > 4000000000ll) { atomicAdd(&g_peer_timeouts, 1u); break; }
        __nanosleep(200);
    }
}

__global__ void __launch_bounds__(512) peer_allgather_kernel(PeerAgArgs a) {
    const int p = blockIdx.x / a.bpg, b = blockIdx.x - p * a.bpg;
    if (p == a.rank) return;
    if (threadIdx.x == 0) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(a.base[p] + a.flag_off);
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= a.step) break;
            if (clock64() - t0 > 4000000000ll) { atomicAdd(&g_peer_timeouts, 1u); break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    const size_t n16 = (size_t)(a.off[p + 1] - a.off[p]) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(a.base[p] + a.off[p]);
    uint4* dst = reinterpret_cast<uint4*>(a.base[a.rank] + a.off[p]);
    for (size_t i = (size_t)b * blockDim.x + threadIdx.x; i < n16; i += (size_t)a.bpg * blockDim.x) dst[i] = src[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(&g_peer_ag_done[p], 1u);
        if (old == (unsigned)a.bpg - 1) {
            g_peer_ag_done[p] = 0;
            __threadfence_system();
            unsigned long long* ack = reinterpret_cast<unsigned long long*>(a.base[p] + a.flag_off) + 1 + a.rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ack), "l"(a.step) : "memory");
        }
    }
}

static size_t plane_floats(const ConvArgs& a, int FH) {
    const int crop_h = a.opt.crop_h > 0 ? a.opt.crop_h : FH;
    const int crop_w = a.opt.crop_w > 0 ? a.opt.crop_w : a.FW;
    const int out_ld = a.opt.out_ld > 0 ? a.opt.out_ld : crop_h;
    return (size_t)crop_w * out_ld;
}

static int bounce_reserve(Ctx& c, size_t bytes) {
    if (bytes <= c.bounce_cap) return 0;
    if (c.bounce) { CU(cudaStreamSynchronize(c.side)); CU(cudaFreeHost(c.bounce)); }
    c.bounce = nullptr; c.bounce_cap = 0;
    CU(cudaMallocHost(&c.bounce, bytes));
    c.bounce_cap = bytes;
    return 0;
}

// true when the host pointer is page-locked (cudaMallocHost / cudaHostRegister): an async D2H into it runs at PCIe rate
static bool host_ptr_pinned(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// n planes of `bytes` each from the pinned bounce half to the caller's (pageable) planes, on a few host threads: first
// touch of freshly allocated output pages and the copy itself are both host-memory bound
static void host_copy_planes(float* const* dst, const char* src, size_t bytes, int n) {
    const size_t total = bytes * (size_t)n;
    unsigned nt = total < ((size_t)4 << 20) ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    nt = (unsigned)std::min<size_t>(nt, (size_t)n);
    auto work = [&](int b, int e) { for (int k = b; k < e; ++k) memcpy(dst[k], src + bytes * (size_t)k, bytes); };
    if (nt <= 1) { work(0, n); return; }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, (int)((size_t)n * t / nt), (int)((size_t)n * (t + 1) / nt));
    work(0, (int)((size_t)n / nt));
    for (auto& t : th) t.join();
}

// Convolve the spectrum with the whole bank, chunk by chunk.
static int run_conv(Ctx& c, const ConvArgs& a, cudaStream_t st) {
    const int CH = a.CH, FW = a.FW, F = a.F, K = a.K;
    const int FH = (CH - 1) * 2;                                   // src/cudaConvFFTData.cu:95
    if (K == 0) return 0;
    const int ldH = odd_ld(FH), ldW = odd_ld(FW);
    if (2 * (size_t)ldH * sizeof(cpx) > kMaxSmem || 3 * (size_t)ldW * sizeof(cpx) > kMaxSmem)
        return fail(FFTCONV_ERR_UNSUPPORTED, "FFT plane %dx%d exceeds the shared-memory line limit", FH, FW);
    if (a.opt.crop_h > FH || a.opt.crop_w > FW || (a.opt.out_ld > 0 && a.opt.out_ld < (a.opt.crop_h > 0 ? a.opt.crop_h : FH)))
        return fail(FFTCONV_ERR_INVALID_INPUT, "crop larger than the FFT plane or out_ld too small");

    int maxkh = 1, maxkw = 1;
    if (a.bankA) { maxkh = a.bank_maxkh; maxkw = a.bank_maxkw; }
    else for (int k = 0; k < K; ++k) {
        maxkh = std::max(maxkh, std::min(a.kernels[k].kh, FH));
        maxkw = std::max(maxkw, std::min(a.kernels[k].kw, FW));
    }
    // raw data in (no compat spectrum exists), batches and prepared banks are served by the overlap-save path only
    const int path = (a.nimg > 1 || a.bankA || a.d_raw) ? PATH_OSGEMM : choose_path(a.opt, F, FH, FW, maxkh, maxkw, K);
    const bool tile16 = path == PATH_TILE16;
    const bool osg = path == PATH_OSGEMM;
    const bool bigp = path == PATH_BIGPLANE;
    OsCfg og;
    if (osg && !os_config(F, FH, FW, maxkh, maxkw, og, a.nimg))
        return fail(FFTCONV_ERR_UNSUPPORTED, "batch outside the range of the overlap-save path");
    const size_t NO = (size_t)K * a.nimg;                          // output planes
    if (a.d_spec && !a.d_raw && a.nimg == 1 && !a.bankA) {
        c.sc.hist = true; c.sc.hist_os = osg; c.sc.h_F = F; c.sc.h_FH = FH; c.sc.h_FW = FW;
    }
    cudaEvent_t spec_ready = a.spec_ready;
    if (spec_ready && !osg) CU(cudaStreamWaitEvent(st, spec_ready, 0));   // only the overlap-save path has image-independent work to run ahead

    // ---- chunking: bound the scratch held per chunk
    const size_t plane = plane_floats(a, FH);
    size_t per_kernel = 0;
    if (tile16) per_kernel = tile16_scratch_per_kernel(FH, FW, F, maxkh, maxkw);
    else if (osg) per_kernel = 0;
    else if (bigp) per_kernel = sizeof(cpx) * ((size_t)F * ((maxkw + 3) & ~3) * bp_chp(CH) + (size_t)FW * bp_chp(CH));
    else per_kernel = sizeof(cpx) * ((size_t)F * maxkw * CH + (size_t)FW * CH);
    if (!a.out_on_device) per_kernel += plane * sizeof(float);
    // keep a chunk's intermediates L2-resident (126 MB L2); a large plane does not fit L2 anyway: there the chunk is
    // as large as memory comfortably allows, so that the CTAs of many templates share each data-spectrum tile
    const size_t budget = bigp ? (size_t)8 << 30 : (size_t)96 << 20;
    int KC = (int)std::max<size_t>(1, std::min<size_t>((size_t)K, budget / std::max<size_t>(per_kernel, 1)));
    if (osg) {
        KC = std::min(K, os_max_chunk(og, a.out_on_device));
        if (int e = os_reserve_chunk(c, og, KC, a.bankA == nullptr)) return e;
        // The data-side transforms (spectrum -> plane -> tile spectra: small, latency-bound grids) do not depend on the
        // bank: they run on a side stream in the shadow of the first os_kern_fft and join before the first GEMM.
        CU(cudaEventRecord(c.evf[0], st));
        CU(cudaStreamWaitEvent(c.side2, c.evf[0], 0));
        if (spec_ready) CU(cudaStreamWaitEvent(c.side2, spec_ready, 0));
        OsProv prov;
        const OsProv* pp = nullptr;
        Ctx::SpecCache& sc = c.sc;
        if (!a.d_raw && a.nimg == 1) {
            if (!g_capturing && !c.plan_pin && os_env().spec_cache && sc.valid && sc.spec == (const void*)a.d_spec && sc.F == F &&
                sc.FH == FH && sc.FW == FW) {
                prov.hsel = (const unsigned long long*)sc.hash.p;
                prov.alt.ptr = (const float*)sc.raw.p; prov.alt.rows = sc.H; prov.alt.cols = sc.W;
                prov.alt_done = (sc.b_valid && sc.b_maxkh == maxkh && sc.b_maxkw == maxkw && sc.b_gen == c.osB_gen) ? 1 : 0;
                pp = &prov;
            }
            sc.want = true; sc.w_F = F; sc.w_FH = FH; sc.w_FW = FW; sc.w_maxkh = maxkh; sc.w_maxkw = maxkw;
        }
        sc.b_valid = false;                                         // whatever happens next, osB belongs to this call
        if (int e = os_prepare_data(c, og, a.d_raw ? nullptr : a.d_spec, a.d_raw, a.rawH, a.rawW, 0, c.side2, pp)) return e;   // correlation = flipped templates + shifted store
        CU(cudaEventRecord(c.evf[1], c.side2));
        if (a.bankA) CU(cudaStreamWaitEvent(st, c.evf[1], 0));      // prepared bank: nothing to overlap with
    } else if (tile16) {
        // one CTA per SM: size the chunk so that NT * ceil(KC/KB) CTAs fill whole waves
        Tile16Cfg g;
        tile16_config(FH, FW, maxkh, maxkw, g);
        const int per_wave = c.sm_count;
        int best_kc = std::min(KC, K);
        if (best_kc < K) {
            for (int waves = 1; ; ++waves) {
                const int ng = (waves * per_wave) / g.NT;
                if (ng < 1) continue;
                const int kc = ng * g.KB;
                if (kc > KC) break;
                best_kc = kc;
            }
            best_kc = std::max(best_kc, g.KB);
        }
        KC = best_kc;
        if (int e = tile16_prepare(c, a.d_spec, FH, FW, F, maxkh, maxkw, KC, st)) return e;
    } else if (bigp) {
        if (int e = dev_reserve(c.T, sizeof(cpx) * (size_t)KC * F * ((maxkw + 3) & ~3) * bp_chp(CH))) return e;
        if (int e = dev_reserve(c.Z, sizeof(cpx) * (size_t)KC * FW * bp_chp(CH))) return e;
        if (int e = bigplane_prepare(c, a.d_spec, CH, FW, F, FH, st)) return e;
    } else {
        if (int e = dev_reserve(c.T, sizeof(cpx) * (size_t)KC * F * maxkw * CH)) return e;
        if (int e = dev_reserve(c.Z, sizeof(cpx) * (size_t)KC * FW * CH)) return e;
    }
    // Host outputs.  Page-locked planes: D2H straight into them.  PAGEABLE planes (what a MEX caller hands over: K separate
    // mxArrays, src/cudaConvFFTData.cu:275-277) would turn every cudaMemcpyAsync into a blocking staged copy and serialise
    // the chunk pipeline: they go through a library-owned pinned bounce ring instead, and this thread copies chunk i out
    // of the ring while the device computes chunk i + 1.
    bool bounce = false;
    if (!a.out_on_device) {
        if (int e = dev_reserve(c.outstage, sizeof(float) * plane * KC * 2)) return e;
        bounce = !(host_ptr_pinned(a.outs[0]) && host_ptr_pinned(a.outs[K - 1]));
        if (bounce) if (int e = bounce_reserve(c, sizeof(float) * plane * KC * 2)) return e;
    }

    // descriptors / kcols / out pointers for ALL kernels go through pinned staging once
    const size_t desc_bytes = (sizeof(SrcDesc) + sizeof(int2) + sizeof(int)) * (size_t)K + sizeof(float*) * NO + 64;
    size_t host_kernel_bytes = 0;
    for (int k = 0; k < K && !a.bankA; ++k)
        if (!a.kernels[k].on_device) host_kernel_bytes += sizeof(float) * (size_t)a.kernels[k].kh * a.kernels[k].kw * F;
    if (int e = dev_reserve(c.desc, desc_bytes)) return e;
    if (host_kernel_bytes)
        if (int e = dev_reserve(c.stage, host_kernel_bytes)) return e;

    SrcDesc* h_desc;
    if (int e = pinned_get(c, desc_bytes, (void**)&h_desc)) return e;
    float** h_outp = reinterpret_cast<float**>(h_desc + K);
    int2* h_khw = reinterpret_cast<int2*>(h_outp + NO);
    int* h_kcols = reinterpret_cast<int*>(h_khw + K);
    SrcDesc* d_desc = reinterpret_cast<SrcDesc*>(c.desc.p);
    float** d_outp = reinterpret_cast<float**>(d_desc + K);
    int2* d_khw = reinterpret_cast<int2*>(d_outp + NO);
    int* d_kcols = reinterpret_cast<int*>(d_khw + K);

    // chunk boundaries.  Host outputs: the D2H stream is the bottleneck (PCIe), so the first chunk is kept
    // small -- its results start flowing to the host while the rest of the bank is still being computed.
    std::vector<int> bounds{0};
    {
        int first = KC;
        if (!a.out_on_device && K > KC) first = std::max(1, osg ? std::min(KC, OS_TM) : KC / 4);
        for (int k = std::min(first, K); ; k = std::min(k + KC, K)) {
            bounds.push_back(k);
            if (k == K) break;
        }
    }
    // device address of every kernel (host kernels are staged; offsets fixed up front, copies issued per chunk)
    std::vector<size_t> stage_off((size_t)K + 1, 0);
    for (int k = 0; k < K && !a.bankA; ++k) {
        const size_t b = a.kernels[k].on_device ? 0 : sizeof(float) * (size_t)a.kernels[k].kh * a.kernels[k].kw * F;
        stage_off[k + 1] = stage_off[k] + b;
        h_desc[k].ptr = a.kernels[k].on_device ? a.kernels[k].ptr
                                               : reinterpret_cast<const float*>(reinterpret_cast<char*>(c.stage.p) + stage_off[k]);
        h_desc[k].rows = a.kernels[k].kh;
        h_desc[k].cols = a.kernels[k].kw;
        h_khw[k] = make_int2(a.kernels[k].kh, a.kernels[k].kw);
        h_kcols[k] = a.kernels[k].kw;
    }
    for (size_t ch = 0; ch + 1 < bounds.size(); ++ch)
        for (int k = bounds[ch]; k < bounds[ch + 1]; ++k)
            h_outp[k] = (a.peak_keys || a.det.mode) ? nullptr : a.out_on_device ? a.outs[k]
                                        : reinterpret_cast<float*>(c.outstage.p) + plane * (size_t)((k - bounds[ch]) + (ch & 1) * KC);
    for (size_t i = K; i < NO; ++i) h_outp[i] = a.outs[i];         // images 1.. of a batch (device planes)
    CU(cudaMemcpyAsync(c.desc.p, h_desc, desc_bytes, cudaMemcpyHostToDevice, st));
    if (int e = pinned_done(c, st)) return e;

    // template transforms of chunk i + 1 next to the inverse of chunk i (OsAhead): device-resident templates, device outputs
    const int ahead_mode = (osg && !a.bankA && a.out_on_device && host_kernel_bytes == 0 && !a.det.mode && bounds.size() > 2)
                               ? os_env().ahead : 0;
    // ---- chunk loop.  Device outputs: one stream.  Host outputs: the D2H of chunk i runs on the
    // side stream while chunk i+1 uploads its kernels and computes (double-buffered out staging).
    for (size_t chunk = 0; chunk + 1 < bounds.size(); ++chunk) {
        const int k0 = bounds[chunk], nk = bounds[chunk + 1] - k0;
        // upload this chunk's host kernels: contiguous runs are coalesced into single copies
        for (int k = k0; k < k0 + nk && !a.bankA;) {
            if (a.kernels[k].on_device) { ++k; continue; }
            const char* run_src = reinterpret_cast<const char*>(a.kernels[k].ptr);
            int j = k;
            while (j < k0 + nk && !a.kernels[j].on_device &&
                   reinterpret_cast<const char*>(a.kernels[j].ptr) == run_src + (stage_off[j] - stage_off[k]))
                ++j;
            CU(cudaMemcpyAsync(reinterpret_cast<char*>(c.stage.p) + stage_off[k], run_src, stage_off[j] - stage_off[k],
                               cudaMemcpyHostToDevice, st));
            k = j;
        }
        if (!a.out_on_device && chunk >= 2) CU(cudaStreamWaitEvent(st, c.ev[chunk & 1], 0));   // staging half free?
        int e;
        if (osg) {
            OsAhead ah;
            if (ahead_mode) {
                ah.kern_done = chunk > 0;
                ah.side = ahead_mode == 2 ? c.side : c.side2;
                if (chunk + 2 < bounds.size()) { ah.next_descs = d_desc + bounds[chunk + 1]; ah.next_nk = bounds[chunk + 2] - bounds[chunk + 1]; }
            }
            e = os_chunk(c, og, d_desc + k0, nk, d_outp + k0, a.opt, st, K,
                         a.bankA ? a.bankA + (size_t)(k0 / OS_TM) * OS_NBIN * og.NKS * (og.a_stage / 8) : nullptr,
                         a.peak_keys ? a.peak_keys + k0 : nullptr, (a.bank_khw ? a.bank_khw : d_khw) + k0, a.rawH, a.rawW,
                         (chunk == 0 && !a.bankA) ? c.evf[1] : nullptr, &a.det, k0, ahead_mode ? &ah : nullptr);
        }
        else if (tile16)
            e = tile16_chunk(c, FH, FW, F, maxkh, maxkw, d_desc + k0, nk, d_outp + k0, a.opt, st);
        else if (bigp)
            e = conv_bigplane_chunk(c, a, FH, d_desc + k0, d_kcols + k0, nk, maxkw, d_outp + k0, st);
        else
            e = conv_generic_chunk(c, a, FH, d_desc + k0, d_kcols + k0, nk, maxkw, d_outp + k0, st);
        if (e) return e;
        if (!a.out_on_device) {
            CU(cudaEventRecord(c.ev[2 + (chunk & 1)], st));
            CU(cudaStreamWaitEvent(c.side, c.ev[2 + (chunk & 1)], 0));
            if (bounce) {
                // (the host copy-out of chunk - 2, which used this ring half, finished in program order below)
                char* half = reinterpret_cast<char*>(c.bounce) + sizeof(float) * plane * (size_t)KC * (chunk & 1);
                CU(cudaMemcpyAsync(half, h_outp[k0], sizeof(float) * plane * (size_t)nk, cudaMemcpyDeviceToHost, c.side));
                CU(cudaEventRecord(c.evb[chunk & 1], c.side));
                CU(cudaEventRecord(c.ev[chunk & 1], c.side));
                if (chunk >= 1) {
                    const int p0 = bounds[chunk - 1], pn = k0 - p0;
                    CU(cudaEventSynchronize(c.evb[(chunk - 1) & 1]));
                    host_copy_planes(a.outs + p0, reinterpret_cast<char*>(c.bounce) + sizeof(float) * plane * (size_t)KC * ((chunk - 1) & 1),
                                     sizeof(float) * plane, pn);
                }
                continue;
            }
            // contiguous host planes -> one copy
            int k = k0;
            while (k < k0 + nk) {
                int j = k + 1;
                while (j < k0 + nk && a.outs[j] == a.outs[j - 1] + plane) ++j;
                CU(cudaMemcpyAsync(a.outs[k], h_outp[k], sizeof(float) * plane * (size_t)(j - k),
                                   cudaMemcpyDeviceToHost, c.side));
                k = j;
            }
            CU(cudaEventRecord(c.ev[chunk & 1], c.side));
        }
    }
    if (!a.out_on_device) {
        if (bounce) {
            const size_t last = bounds.size() - 2;
            const int p0 = bounds[last], pn = K - p0;
            CU(cudaEventSynchronize(c.evb[last & 1]));
            host_copy_planes(a.outs + p0, reinterpret_cast<char*>(c.bounce) + sizeof(float) * plane * (size_t)KC * (last & 1),
                             sizeof(float) * plane, pn);
        }
        CU(cudaStreamSynchronize(c.side));
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

static int check_threads(const double* threads, int nthreads) {
    if (threads != nullptr && nthreads != 4) return fail(FFTCONV_ERR_THREAD_SIZE, "%s", kMsgThread);
    if (threads == nullptr && nthreads != 0 && nthreads != 4) return fail(FFTCONV_ERR_THREAD_SIZE, "%s", kMsgThread);
    return 0;
}

static int build_kernel_refs(int K, const float* const* kernels, const int* kh, const int* kw, const int* kf,
                             const unsigned char* on_dev, int F, int FH, int FW, std::vector<KernelRef>& refs) {
    if (K < 0 || (K > 0 && (!kernels || !kh || !kw)))
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    refs.resize(K);
    for (int k = 0; k < K; ++k) {
        if (!kernels[k] || kh[k] <= 0 || kw[k] <= 0) return fail(FFTCONV_ERR_KERNEL_TYPE,
            "Kernels must be of type float and have features larger than 1");
        // src/cudaConvFFTData.cu:229
        if ((kf && kf[k] != F) || kw[k] > FW || kh[k] > FH) return fail(FFTCONV_ERR_KERNEL_SHAPE, "%s", kMsgKernelShape);
        refs[k].ptr = kernels[k];
        refs[k].kh = kh[k];
        refs[k].kw = kw[k];
        refs[k].on_device = on_dev ? on_dev[k] != 0 : false;
    }
    return 0;
}

}  // namespace fftconv

using namespace fftconv;

// =================================================================================== C ABI
extern "C" {

int fftconv_fft_size16(int n) {            // src/cudaConvFFTData.h:96-102
    const int mod = n / 16, rem = n % 16;
    return mod * 16 + (rem > 0 ? 16 : 0);
}

int fftconv_fft_size_pow2(int n) {         // src/cudaConvFFTData.h:67-94
    n = (n % 16 != 0) ? (n - n % 16 + 16) : n;
    int hi;
    for (hi = 31; hi >= 0; --hi)
        if ((unsigned)n & (1u << hi)) break;
    const unsigned low = 1u << hi;
    if (low == (unsigned)n) return n;
    return (int)(1u << (hi + 1));
}

// Provenance of a spectrum for the overlap-save path (Ctx::SpecCache): keep a copy of the raw data behind `d_spec` and the
// hash of the spectrum as it is NOW (stream order).  Called by fftconv_fft_data for the spectrum it just wrote, and by
// fftconv_spectrum_bind_raw for a spectrum the caller assembled elsewhere.
static int spec_cache_record(Ctx& c, const void* d_spec, const float* d_data, int H, int W, int F, int FH, int FW, bool zero_pad,
                             cudaStream_t st) {
    Ctx::SpecCache& sc = c.sc;
    const size_t raw_bytes = sizeof(float) * (size_t)H * W * F;
    if (sc.spec == d_spec) { sc.valid = false; sc.b_valid = false; }
    if (sc.hist && !sc.hist_os && sc.h_F == F && sc.h_FH == FH && sc.h_FW == FW) return 0;   // another pipeline serves this geometry
    if (zero_pad && !g_capturing && !c.plan_pin && os_env().spec_cache && raw_bytes <= ((size_t)512 << 20)) {
        sc.valid = false; sc.b_valid = false;
        if (int e = dev_reserve(sc.raw, raw_bytes)) return e;
        if (int e = dev_reserve(sc.hash, 2 * sizeof(unsigned long long))) return e;
        CU(cudaMemcpyAsync(sc.raw.p, d_data, raw_bytes, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemsetAsync(sc.hash.p, 0, 2 * sizeof(unsigned long long), st));
        const size_t n = (size_t)F * FW * (FH / 2 + 1);
        os_hash64<<<os_hash_grid(c, n), 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_spec), n,
                                                      (unsigned long long*)sc.hash.p);
        LAUNCH_CHECK();
        sc.valid = true; sc.spec = d_spec; sc.H = H; sc.W = W; sc.F = F; sc.FH = FH; sc.FW = FW;
        // the last convolution fed by a spectrum of this geometry took the overlap-save path: transform the tiles now, on
        // the data-side stream, next to whatever the caller does until it convolves
        OsCfg og;
        if (sc.want && sc.w_F == F && sc.w_FH == FH && sc.w_FW == FW && os_env().spec_cache > 1 &&
            os_config(F, FH, FW, sc.w_maxkh, sc.w_maxkw, og)) {
            CU(cudaEventRecord(c.evf[0], st));
            CU(cudaStreamWaitEvent(c.side2, c.evf[0], 0));
            if (int e = os_prepare_data(c, og, nullptr, (const float*)sc.raw.p, H, W, 0, c.side2)) return e;
            sc.b_valid = true; sc.b_maxkh = sc.w_maxkh; sc.b_maxkw = sc.w_maxkw; sc.b_gen = c.osB_gen;
        }
    }
    return 0;
}

int fftconv_spectrum_bind_raw(const fftconv_float2* d_spec, const float* d_raw, int H, int W, int F, int maxKH, int maxKW,
                              int device, void* stream) {
    g_err.clear();
    if (!d_spec || !d_raw || H <= 0 || W <= 0 || F <= 0 || maxKH <= 0 || maxKW <= 0)
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid data input");
    const int FH = fftconv_fft_size16(H + maxKH - 1), FW = fftconv_fft_size16(W + maxKW - 1);
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    return spec_cache_record(*cs.c, d_spec, d_raw, H, W, F, FH, FW, true, (cudaStream_t)stream);
}

static int fft_data_impl(const float* data, int data_on_device, int H, int W, int F, int KH, int KW,
                         int pad_mode, int kernel_y, int kernel_x, fftconv_float2* d_spec, int device, void* stream) {
    g_err.clear();
    if (!data || !d_spec || H <= 0 || W <= 0 || F <= 0 || KH <= 0 || KW <= 0)
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    cudaStream_t st = (cudaStream_t)stream;
    const int FH = fftconv_fft_size16(H + KH - 1), FW = fftconv_fft_size16(W + KW - 1);
    const float* d_data = data;
    if (!data_on_device) {
        const size_t bytes = sizeof(float) * (size_t)H * W * F;
        if (int e = dev_reserve(c->ddata, bytes)) return e;
        CU(cudaMemcpyAsync(c->ddata.p, data, bytes, cudaMemcpyHostToDevice, st));
        d_data = (const float*)c->ddata.p;
    }
    if (int e = run_fft_data(*c, d_data, H, W, F, FH, FW, pad_mode, kernel_y, kernel_x, (cpx*)d_spec, st)) return e;
    if (int e = spec_cache_record(*c, d_spec, d_data, H, W, F, FH, FW, pad_mode == PAD_ZERO, st)) return e;
    if (!data_on_device) CU(cudaStreamSynchronize(st));    // src/cudaFFTData.cu:147
    return 0;
}

int fftconv_fft_data(const float* data, int data_on_device, int H, int W, int F, int KH, int KW,
                     fftconv_float2* d_spec, int device, void* stream) {
    return fft_data_impl(data, data_on_device, H, W, F, KH, KW, PAD_ZERO, 0, 0, d_spec, device, stream);
}

int fftconv_fft_data_clamp(const float* data, int data_on_device, int H, int W, int F, int KH, int KW,
                           int kernel_y, int kernel_x, fftconv_float2* d_spec, int device, void* stream) {
    if (kernel_y < 0 || kernel_x < 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    return fft_data_impl(data, data_on_device, H, W, F, KH, KW, PAD_CLAMP, kernel_y, kernel_x, d_spec, device, stream);
}

struct ConvRaw { const float* d_data; int H, W; int nimg = 1; };

static int conv_impl(const fftconv_float2* d_spec, int CH, int FW, int F, int K, const float* const* kernels,
                     const int* kh, const int* kw, const int* kf, const unsigned char* kernel_on_device,
                     float* const* outs, int out_on_device, const double* threads, int nthreads,
                     const fftconv_options* opt, int device, void* stream, const ConvRaw* raw = nullptr) {
    g_err.clear();
    // the one-shot spectrum-ready event belongs to THIS call whatever its outcome (a validation error or K == 0 must
    // not leave a stale event pointer behind for a later call)
    cudaEvent_t spec_ready = nullptr;
    {
        Ctx* c0 = nullptr;
        { std::lock_guard<std::mutex> lk(g_mu); auto it = g_ctx.find(device); if (it != g_ctx.end()) c0 = &it->second; }
        if (c0) { std::lock_guard<std::recursive_mutex> lkc(c0->mu); spec_ready = c0->spec_ready; c0->spec_ready = nullptr; }
    }
    if (!d_spec && !raw) return fail(FFTCONV_ERR_NOT_GPU_ARRAY, "The data must be FFT-ed real array in GPU");
    if (CH < 2 || FW <= 0 || F <= 0 || (K > 0 && !outs)) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    if (int e = check_threads(threads, nthreads)) return e;
    const int FH = (CH - 1) * 2;
    if (FH % 16 != 0 || FW % 16 != 0)
        return fail(FFTCONV_ERR_INVALID_INPUT, "spectrum dims (%d x %d) do not come from computeFFTsize16", CH, FW);
    std::vector<KernelRef> refs;
    if (int e = build_kernel_refs(K, kernels, kh, kw, kf, kernel_on_device, F, FH, FW, refs)) return e;
    const size_t nout = (size_t)K * (raw ? raw->nimg : 1);
    for (size_t k = 0; k < nout; ++k)
        if (!outs[k]) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    ConvArgs a;
    a.d_spec = (const cpx*)d_spec; a.CH = CH; a.FW = FW; a.F = F; a.K = K;
    a.kernels = refs.data(); a.outs = outs; a.out_on_device = out_on_device != 0;
    a.opt = opt ? *opt : fftconv_options{};
    a.spec_ready = spec_ready;
    if (raw) { a.d_raw = raw->d_data; a.rawH = raw->H; a.rawW = raw->W; a.nimg = raw->nimg; }
    return run_conv(*c, a, (cudaStream_t)stream);
}

int fftconv_conv_fft_data(const fftconv_float2* d_spec, int CH, int FW, int F, int K, const float* const* kernels,
                          const int* kh, const int* kw, const int* kf, const unsigned char* kernel_on_device,
                          float* const* outs, int out_on_device, const double* threads, int nthreads,
                          const fftconv_options* opt, int device, void* stream) {
    return conv_impl(d_spec, CH, FW, F, K, kernels, kh, kw, kf, kernel_on_device, outs, out_on_device, threads,
                     nthreads, opt, device, stream);
}

int fftconv_conv_fft_data_streams(const fftconv_float2* d_spec, int CH, int FW, int F, int K,
                                  const float* const* kernels, const int* kh, const int* kw, const int* kf,
                                  float* const* outs, const double* threads, int nthreads,
                                  const fftconv_options* opt, int device) {
    return conv_impl(d_spec, CH, FW, F, K, kernels, kh, kw, kf, nullptr, outs, 0, threads, nthreads, opt, device,
                     nullptr);
}

int fftconv_convolution_fft(const float* data, int data_on_device, int H, int W, int F, int maxKH, int maxKW, int K,
                            const float* const* kernels, const int* kh, const int* kw, const int* kf,
                            const unsigned char* kernel_on_device, float* const* outs, int out_on_device,
                            const double* threads, int nthreads, const fftconv_options* opt, int device, void* stream) {
    g_err.clear();
    if (!data || H <= 0 || W <= 0 || F <= 0 || maxKH <= 0 || maxKW <= 0)
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid data input");
    if (int e = check_threads(threads, nthreads)) return e;
    const int FH = fftconv_fft_size16(H + maxKH - 1), FW = fftconv_fft_size16(W + maxKW - 1);   // :109-112
    const int CH = FH / 2 + 1;
    // which pipeline will serve the bank?  The overlap-save / tensor-core path tiles the raw data itself,
    // so the full-plane R2C transform of the data (src/cudaConvolutionFFT.cu:155-168) is not needed there.
    bool raw_path = false;
    if (K > 0 && kernels && kh && kw && outs) {
        int maxkh = 1, maxkw = 1;
        for (int k = 0; k < K; ++k) {
            maxkh = std::max(maxkh, std::min(kh[k], FH));
            maxkw = std::max(maxkw, std::min(kw[k], FW));
        }
        const fftconv_options o = opt ? *opt : fftconv_options{};
        raw_path = choose_path(o, F, FH, FW, maxkh, maxkw, K) == PATH_OSGEMM;
    }
    void* spec = nullptr;
    const float* d_data = data;
    // the device lock is held across the WHOLE call: the staged image (ddata / dspec) must not be overwritten by another
    // host thread between the data stage and the bank loop
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    {
        Ctx* c = cs.c;
        if (!raw_path) {
            if (int e = dev_reserve(c->dspec, sizeof(cpx) * (size_t)CH * FW * F)) return e;
            spec = c->dspec.p;
        } else if (!data_on_device) {
            const size_t bytes = sizeof(float) * (size_t)H * W * F;
            if (int e = dev_reserve(c->ddata, bytes)) return e;
            CU(cudaMemcpyAsync(c->ddata.p, data, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
            d_data = (const float*)c->ddata.p;
        }
    }
    if (raw_path) {
        ConvRaw raw{d_data, H, W};
        return conv_impl(nullptr, CH, FW, F, K, kernels, kh, kw, kf, kernel_on_device, outs, out_on_device, threads,
                         nthreads, opt, device, stream, &raw);
    }
    // stream-ordered: no sync between the data transform and the bank loop (the reference
    // synchronises the device here, src/cudaConvolutionFFT.cu:168)
    int e = fft_data_impl(data, data_on_device, H, W, F, maxKH, maxKW, PAD_ZERO, 0, 0, (fftconv_float2*)spec, device, stream);
    if (e) return e;
    return conv_impl((const fftconv_float2*)spec, CH, FW, F, K, kernels, kh, kw, kf, kernel_on_device, outs,
                     out_on_device, threads, nthreads, opt, device, stream);
}

struct fftconv_bank {
    int device, K, F, maxkh, maxkw, NKS, KC;
    float* A;             // [ceil(K/128)][bin][ks][kc][128][4] fp32
    size_t bytes;
    int2* khw;            // (kh, kw) per template, device
    bool owns;            // false: A lives in the cached workspace (the per-call bank of fftconv_conv_batch)
};

static int bank_conv_images(const fftconv_bank* b, const float* d_data, int nimg, int H, int W, float* const* outs,
                            int out_on_device, const fftconv_options& o, cudaStream_t st);
static int bank_create_impl(int K, const float* const* kernels, const int* kh, const int* kw, const int* kf,
                            const unsigned char* kernel_on_device, int F, int device, void* stream, fftconv_bank** out,
                            bool in_workspace);

int fftconv_conv_batch(const float* data, int data_on_device, int N, int H, int W, int F, int maxKH, int maxKW, int K,
                       const float* const* kernels, const int* kh, const int* kw, const int* kf,
                       const unsigned char* kernel_on_device, float* const* outs, int out_on_device,
                       const fftconv_options* opt, int device, void* stream) {
    g_err.clear();
    if (!data || N <= 0 || H <= 0 || W <= 0 || F <= 0 || maxKH <= 0 || maxKW <= 0)
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid data input");
    if (K > 0 && (!kernels || !kh || !kw || !outs)) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    const int FH = fftconv_fft_size16(H + maxKH - 1), FW = fftconv_fft_size16(W + maxKW - 1);
    const int CH = FH / 2 + 1;
    const size_t img = (size_t)H * W * F;
    const fftconv_options o = opt ? *opt : fftconv_options{};
    int maxkh = 1, maxkw = 1;
    for (int k = 0; k < K; ++k) {
        maxkh = std::max(maxkh, std::min(kh[k], FH));
        maxkw = std::max(maxkw, std::min(kw[k], FW));
    }
    CtxScope whole_call(device, (cudaStream_t)stream);      // see fftconv_convolution_fft
    if (whole_call.err) return whole_call.err;
    OsCfg g1;
    const bool batched = K > 0 && N > 1 && out_on_device && !o.correlate && !o.force_generic &&
                         (o.path == PATH_AUTO || o.path == PATH_OSGEMM) && os_config(F, FH, FW, maxkh, maxkw, g1, 1);
    if (!batched) {                     // image by image through the single-image entry point
        for (int n = 0; n < N; ++n)
            if (int e = fftconv_convolution_fft(data + (size_t)n * img, data_on_device, H, W, F, maxKH, maxKW, K, kernels, kh,
                                                kw, kf, kernel_on_device, outs + (size_t)n * K, out_on_device, nullptr, 0,
                                                opt, device, stream))
                return e;
        return 0;
    }
    // the images of a group only add tiles to the N dimension of the per-bin GEMM; the group is sized so that
    // the product spectra of one template chunk stay within a few GB
    const int G = std::max(1, std::min(N, 1280 / g1.NTimg));
    // several groups: the template spectra (A operand images) do not depend on the images -- transform the bank ONCE
    // (a prepared bank that lives for this call) instead of once per group
    fftconv_bank* tmp_bank = nullptr;
    if (N > G && fftconv_fft_size16(H + maxkh - 1) == FH && fftconv_fft_size16(W + maxkw - 1) == FW) {
        if (int e = bank_create_impl(K, kernels, kh, kw, kf, kernel_on_device, F, device, stream, &tmp_bank, true)) return e;
        if (tmp_bank->maxkh != maxkh || tmp_bank->maxkw != maxkw) { fftconv_bank_destroy(tmp_bank); tmp_bank = nullptr; }
    }
    struct BankGuard { fftconv_bank* b; ~BankGuard() { if (b) fftconv_bank_destroy(b); } } bank_guard{tmp_bank};
    for (int n0 = 0; n0 < N; n0 += G) {
        const int nimg = std::min(G, N - n0);
        const float* d_data = data + (size_t)n0 * img;
        if (!data_on_device) {
            CtxScope cs(device, (cudaStream_t)stream);
            if (cs.err) return cs.err;
            Ctx* c = cs.c;
            if (int e = dev_reserve(c->ddata, sizeof(float) * img * nimg)) return e;
            CU(cudaMemcpyAsync(c->ddata.p, d_data, sizeof(float) * img * nimg, cudaMemcpyHostToDevice, (cudaStream_t)stream));
            d_data = (const float*)c->ddata.p;
        }
        if (tmp_bank) {
            if (int e = bank_conv_images(tmp_bank, d_data, nimg, H, W, outs + (size_t)n0 * K, 1, o, (cudaStream_t)stream)) return e;
            continue;
        }
        ConvRaw raw{d_data, H, W, nimg};
        if (int e = conv_impl(nullptr, CH, FW, F, K, kernels, kh, kw, kf, kernel_on_device, outs + (size_t)n0 * K, 1,
                              nullptr, 0, opt, device, stream, &raw))
            return e;
    }
    return 0;
}

// ---- feature pyramid x one bank in one call (BASELINE config 5).  The levels differ in size, the templates do not: the
// tiles of ALL levels go through the per-bin GEMM as one N dimension (level-major tile numbering, OsLevel table), so the
// template spectra are computed ONCE and the A operand images are read once per template block instead of once per level,
// and a template chunk costs one GEMM launch and one inverse launch instead of one per level.
struct PyrLevel { const float* d_raw; const cpx* d_spec; int H, W, FH, FW; };

static int run_conv_pyramid(Ctx& c, int L, const PyrLevel* lv, int F, int maxkh, int maxkw, int K, const KernelRef* kernels,
                            float* const* outs, const fftconv_options& opt, cudaStream_t st, cudaEvent_t data_ready = nullptr) {
    OsCfg og;
    if (!os_config(F, 64, 64, maxkh, maxkw, og)) return fail(FFTCONV_ERR_UNSUPPORTED, "pyramid batch needs templates of at most 32 x 32");
    std::vector<OsLevel> hl((size_t)L);
    int NT = 0;
    size_t plane_floats_total = 0, z_max = 0;
    int nspec_planes = 0;
    for (int l = 0; l < L; ++l) {
        OsLevel& o = hl[(size_t)l];
        o.FH = lv[l].FH; o.FW = lv[l].FW;
        o.nth = (o.FH + og.Sh - 1) / og.Sh;
        const int ntw = (o.FW + og.Sw - 1) / og.Sw;
        o.m0 = NT; NT += o.nth * ntw;
        o.crop_h = o.FH; o.crop_w = o.FW; o.out_ld = o.FH;
        if (lv[l].d_raw) { o.src = lv[l].d_raw; o.rows = lv[l].H; o.cols = lv[l].W; }
        else {
            o.src = nullptr; o.rows = o.FH; o.cols = o.FW;             // plane recovered from the spectrum (below)
            plane_floats_total += (size_t)F * o.FW * o.FH;
            z_max = std::max(z_max, (size_t)F * o.FW * (o.FH / 2 + 1));
            nspec_planes += F;
        }
    }
    // same scratch bound as the image groups of fftconv_conv_batch: beyond it the levels go one by one
    if (NT > 1280) return fail(FFTCONV_ERR_UNSUPPORTED, "pyramid batch: %d tiles exceed one GEMM problem", NT);
    og.FH = 0; og.FW = 0; og.nimg = L; og.nth = 1; og.ntw = 1; og.NT = NT; og.NTimg = NT;
    if (!os_config_tiles(og)) return fail(FFTCONV_ERR_UNSUPPORTED, "pyramid batch outside the range of the overlap-save path");

    // ---- descriptor tables: [OsLevel L][SrcDesc K][float* L*K][int2 K][float* nspec_planes], one staged copy
    const size_t NO = (size_t)L * K;
    const size_t off_desc = (sizeof(OsLevel) * (size_t)L + 15) & ~(size_t)15;
    const size_t off_outp = off_desc + sizeof(SrcDesc) * (size_t)K;
    const size_t off_khw = off_outp + sizeof(float*) * NO;
    const size_t off_pl = (off_khw + sizeof(int2) * (size_t)K + 15) & ~(size_t)15;
    const size_t desc_bytes = off_pl + sizeof(float*) * (size_t)nspec_planes + 64;
    if (int e = dev_reserve(c.desc, desc_bytes)) return e;
    size_t host_kernel_bytes = 0;
    for (int k = 0; k < K; ++k)
        if (!kernels[k].on_device) host_kernel_bytes += sizeof(float) * (size_t)kernels[k].kh * kernels[k].kw * F;
    if (host_kernel_bytes) if (int e = dev_reserve(c.stage, host_kernel_bytes)) return e;
    if (plane_floats_total) {
        if (int e = dev_reserve(c.osPlane, sizeof(float) * plane_floats_total)) return e;
        if (int e = dev_reserve(c.osZ, sizeof(cpx) * z_max)) return e;
    }
    if (int e = dev_reserve(c.osB, (size_t)og.NNB * OS_NBIN * og.b_buf)) return e;
    const int KC = std::min(K, os_max_chunk(og, true));
    if (int e = os_reserve_chunk(c, og, KC, true)) return e;

    char* h_tab;
    if (int e = pinned_get(c, desc_bytes, (void**)&h_tab)) return e;
    char* d_tab = reinterpret_cast<char*>(c.desc.p);
    OsLevel* h_lv = reinterpret_cast<OsLevel*>(h_tab);
    SrcDesc* h_desc = reinterpret_cast<SrcDesc*>(h_tab + off_desc);
    float** h_outp = reinterpret_cast<float**>(h_tab + off_outp);
    int2* h_khw = reinterpret_cast<int2*>(h_tab + off_khw);
    float** h_pl = reinterpret_cast<float**>(h_tab + off_pl);
    float* plane = reinterpret_cast<float*>(c.osPlane.p);
    std::vector<float*> level_plane((size_t)L, nullptr);
    {
        size_t po = 0; int pi = 0;
        for (int l = 0; l < L; ++l) {
            if (!lv[l].d_raw) {
                level_plane[(size_t)l] = plane + po;
                hl[(size_t)l].src = plane + po;
                for (int f = 0; f < F; ++f) h_pl[pi++] = plane + po + (size_t)f * hl[(size_t)l].FW * hl[(size_t)l].FH;
                po += (size_t)F * hl[(size_t)l].FW * hl[(size_t)l].FH;
            }
            h_lv[l] = hl[(size_t)l];
        }
    }
    size_t so = 0;
    for (int k = 0; k < K; ++k) {
        const size_t b = kernels[k].on_device ? 0 : sizeof(float) * (size_t)kernels[k].kh * kernels[k].kw * F;
        h_desc[k].ptr = kernels[k].on_device ? kernels[k].ptr : reinterpret_cast<const float*>(reinterpret_cast<char*>(c.stage.p) + so);
        h_desc[k].rows = kernels[k].kh; h_desc[k].cols = kernels[k].kw;
        h_khw[k] = make_int2(kernels[k].kh, kernels[k].kw);
        if (b) CU(cudaMemcpyAsync(reinterpret_cast<char*>(c.stage.p) + so, kernels[k].ptr, b, cudaMemcpyHostToDevice, st));
        so += b;
    }
    for (size_t i = 0; i < NO; ++i) h_outp[i] = outs[i];
    CU(cudaMemcpyAsync(d_tab, h_tab, desc_bytes, cudaMemcpyHostToDevice, st));
    if (int e = pinned_done(c, st)) return e;
    og.d_levels = reinterpret_cast<const OsLevel*>(d_tab);
    og.nlevels = L;

    // ---- data side: spectrum -> plane for the levels that arrive as cudaFFTData spectra, then one tiling launch.  None of it
    // depends on the bank: it runs on the side stream next to the first os_kern_fft (as in run_conv) and is the only work
    // that waits for a pyramid still in flight on another stream (data_ready: fftconv_spectrum_ready_event).
    cudaStream_t ds = c.side2;
    CU(cudaEventRecord(c.evf[0], st));                                 // the descriptor tables are on their way
    CU(cudaStreamWaitEvent(ds, c.evf[0], 0));
    if (data_ready) CU(cudaStreamWaitEvent(ds, data_ready, 0));
    {
        float** d_pl = reinterpret_cast<float**>(d_tab + off_pl);
        int pi = 0;
        for (int l = 0; l < L; ++l) {
            if (lv[l].d_raw) continue;
            const int FH = lv[l].FH, FW = lv[l].FW, CH = FH / 2 + 1;
            const cpx *twH, *twW;
            if (int e = get_twiddles(c, FH, ds, &twH)) return e;
            if (int e = get_twiddles(c, FW, ds, &twW)) return e;
            const LinePlan pH = make_line_plan(FH), pW = make_line_plan(FW);
            const int ldH = odd_ld(FH), ldW = odd_ld(FW);
            ProfScope ps(PK_OS_PLANE, ds);
            int TU = (int)((96 * 1024) / (2 * (size_t)ldW * sizeof(cpx)));
            TU = TU < 1 ? 1 : (TU > 16 ? 16 : TU);
            dim3 g2((CH + TU - 1) / TU, F);
            inv_w_pass<<<g2, 256, 2 * (size_t)TU * ldW * sizeof(cpx), ds>>>(lv[l].d_spec, FW, CH, pW, twW, (cpx*)c.osZ.p, TU, ldW, nullptr);
            LAUNCH_CHECK();
            const int NL = pick_lines(FH, 8);
            const long long nlines = (long long)F * (FW / 2);
            inv_h_pass<<<(unsigned)((nlines + NL - 1) / NL), 256, 2 * (size_t)NL * ldH * sizeof(cpx), ds>>>(
                (const cpx*)c.osZ.p, F, FH, FW, CH, pH, twH, 1.0f / ((float)FW * (float)FH), d_pl + pi, FH, FW, FH, NL, ldH, nullptr);
            LAUNCH_CHECK();
            pi += F;
        }
        OsDArgs a{};
        a.levels = og.d_levels; a.nlevels = L;
        a.F = F; a.nth = 1; a.NTimg = NT; a.Sh = og.Sh; a.Sw = og.Sw; a.oy0 = maxkh - 1; a.ox0 = maxkw - 1;
        a.FH = 64; a.FW = 64; a.img = (float*)c.osB.p; a.NKS = og.NKS; a.KC = og.KC; a.NMMA = og.NMMA; a.NTn = og.NTn;
        a.correlate = 0;
        a.st256 = os_env().data_st256;
        c.sc.b_valid = false; ++c.osB_gen;                             // the B images now belong to this call
        const unsigned grid = (unsigned)NT * (unsigned)(og.NKS * og.KC);
        ProfScope ps(PK_OS_DATA, ds);
        if (os_env().data_occ) os_data_fft_occ<<<grid, 128, OS_DATA_SMEM_OCC, ds>>>(a);
        else os_data_fft<<<grid, 128, OS_DATA_SMEM, ds>>>(a);
        LAUNCH_CHECK();
    }
    CU(cudaEventRecord(c.evf[1], ds));                                 // joined by the first chunk in front of its os_gemm
    // ---- template chunks
    const SrcDesc* d_desc = reinterpret_cast<const SrcDesc*>(d_tab + off_desc);
    float* const* d_outp = reinterpret_cast<float* const*>(d_tab + off_outp);
    const int2* d_khw = reinterpret_cast<const int2*>(d_tab + off_khw);
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int nk = std::min(KC, K - k0);
        if (int e = os_chunk(c, og, d_desc + k0, nk, d_outp + k0, opt, st, K, nullptr, nullptr, d_khw + k0, 0, 0,
                             k0 == 0 ? c.evf[1] : nullptr, nullptr, 0))
            return e;
    }
    return 0;
}

int fftconv_conv_pyramid(int L, const float* const* level_data, const fftconv_float2* const* level_spec, const int* H,
                         const int* W, int F, int maxKH, int maxKW, int K, const float* const* kernels, const int* kh,
                         const int* kw, const int* kf, const unsigned char* kernel_on_device, float* const* outs,
                         const fftconv_options* opt, int device, void* stream) {
    g_err.clear();
    // one-shot event of fftconv_spectrum_ready_event: the levels are still arriving on another stream (an NCCL broadcast of
    // the packed pyramid); it belongs to THIS call whatever its outcome
    cudaEvent_t data_ready = nullptr;
    {
        Ctx* c0 = nullptr;
        { std::lock_guard<std::mutex> lk(g_mu); auto it = g_ctx.find(device); if (it != g_ctx.end()) c0 = &it->second; }
        if (c0) { std::lock_guard<std::recursive_mutex> lkc(c0->mu); data_ready = c0->spec_ready; c0->spec_ready = nullptr; }
    }
    if (L <= 0 || L > OS_MAX_LEVELS || !H || !W || F <= 0 || maxKH <= 0 || maxKW <= 0 || (!level_data && !level_spec))
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid data input");
    if (K > 0 && (!kernels || !kh || !kw || !outs)) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    if (K == 0) return 0;
    const fftconv_options o = opt ? *opt : fftconv_options{};
    std::vector<PyrLevel> lv((size_t)L);
    int minFH = 1 << 30, minFW = 1 << 30;
    for (int l = 0; l < L; ++l) {
        PyrLevel& p = lv[(size_t)l];
        p.d_raw = level_data ? level_data[l] : nullptr;
        p.d_spec = (!p.d_raw && level_spec) ? (const cpx*)level_spec[l] : nullptr;
        if ((!p.d_raw && !p.d_spec) || H[l] <= 0 || W[l] <= 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid data input");
        p.H = H[l]; p.W = W[l];
        p.FH = fftconv_fft_size16(H[l] + maxKH - 1); p.FW = fftconv_fft_size16(W[l] + maxKW - 1);   // src/cudaFFTData.cu:109-112
        minFH = std::min(minFH, p.FH); minFW = std::min(minFW, p.FW);
    }
    for (size_t i = 0; i < (size_t)L * K; ++i)
        if (!outs[i]) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    std::vector<KernelRef> refs;
    if (int e = build_kernel_refs(K, kernels, kh, kw, kf, kernel_on_device, F, minFH, minFW, refs)) return e;
    int maxkh = 1, maxkw = 1;
    for (int k = 0; k < K; ++k) { maxkh = std::max(maxkh, refs[k].kh); maxkw = std::max(maxkw, refs[k].kw); }
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    OsCfg g1;
    // the tiles must reproduce the planes the declared maxKernel sizes imply: an actual template larger than the declared
    // maximum wraps around the plane and is left to the per-level calls
    const bool batched = !o.correlate && !o.force_generic && (o.path == PATH_AUTO || o.path == PATH_OSGEMM) &&
                         o.crop_h <= 0 && o.crop_w <= 0 && o.out_ld <= 0 && maxkh <= maxKH && maxkw <= maxKW &&
                         os_config(F, 64, 64, maxkh, maxkw, g1);
    if (batched) {
        const int e = run_conv_pyramid(*cs.c, L, lv.data(), F, maxkh, maxkw, K, refs.data(), outs, o, (cudaStream_t)stream, data_ready);
        if (e != FFTCONV_ERR_UNSUPPORTED) return e;
        g_err.clear();
    }
    if (data_ready) CU(cudaStreamWaitEvent((cudaStream_t)stream, data_ready, 0));   // level by level: everything waits
    for (int l = 0; l < L; ++l) {                   // level by level through the single-image entry points
        const PyrLevel& p = lv[(size_t)l];
        int e;
        if (p.d_raw) e = fftconv_convolution_fft(p.d_raw, 1, p.H, p.W, F, maxKH, maxKW, K, kernels, kh, kw, kf, kernel_on_device,
                                                 outs + (size_t)l * K, 1, nullptr, 0, opt, device, stream);
        else e = conv_impl((const fftconv_float2*)p.d_spec, p.FH / 2 + 1, p.FW, F, K, kernels, kh, kw, kf, kernel_on_device,
                           outs + (size_t)l * K, 1, nullptr, 0, opt, device, stream);
        if (e) return e;
    }
    return 0;
}

// ---- prepared banks: the template spectra (tcgen05 A operand images) are independent of the image size on the
// overlap-save path (64 x 64 tiles), so they are computed once and stay resident in HBM.

// in_workspace: the A images go to the cached workspace and the call stays stream-ordered (no allocation, no host
// synchronisation): the per-call bank of fftconv_conv_batch.  Otherwise the bank owns its memory and the call returns
// when the transform is done.
static int bank_create_impl(int K, const float* const* kernels, const int* kh, const int* kw, const int* kf,
                            const unsigned char* kernel_on_device, int F, int device, void* stream, fftconv_bank** out,
                            bool in_workspace) {
    g_err.clear();
    if (!out || K <= 0 || F <= 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    *out = nullptr;
    std::vector<KernelRef> refs;
    if (int e = build_kernel_refs(K, kernels, kh, kw, kf, kernel_on_device, F, 1 << 30, 1 << 30, refs)) return e;
    int maxkh = 1, maxkw = 1;
    for (int k = 0; k < K; ++k) { maxkh = std::max(maxkh, refs[k].kh); maxkw = std::max(maxkw, refs[k].kw); }
    OsCfg g;
    if (!os_config(F, 64, 64, maxkh, maxkw, g))
        return fail(FFTCONV_ERR_UNSUPPORTED, "prepared banks need templates of at most 32 x 32 (got %d x %d)", maxkh, maxkw);
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    cudaStream_t st = (cudaStream_t)stream;
    const int ntblk = (K + OS_TM - 1) / OS_TM;
    std::unique_ptr<fftconv_bank> b(new fftconv_bank{device, K, F, maxkh, maxkw, g.NKS, g.KC, nullptr, 0, nullptr, !in_workspace});
    b->bytes = (size_t)ntblk * OS_NBIN * g.NKS * (g.a_stage / 2);
    if (in_workspace) {
        if (int e = dev_reserve(c->batchA, b->bytes)) return e;
        b->A = (float*)c->batchA.p;
    } else {
        CU(cudaMalloc(&b->A, b->bytes));
    }
    if (!in_workspace) {
        std::vector<int2> khw(K);
        for (int k = 0; k < K; ++k) khw[k] = make_int2(refs[k].kh, refs[k].kw);
        if (cudaMalloc(&b->khw, sizeof(int2) * (size_t)K) != cudaSuccess ||
            cudaMemcpy(b->khw, khw.data(), sizeof(int2) * (size_t)K, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(b->A); cudaFree(b->khw);
            return fail(FFTCONV_ERR_CUDA, "bank allocation failed");
        }
    }
    // stage host kernels + descriptors (the call is synchronous, so the shared staging buffers can be reused)
    size_t host_bytes = 0;
    for (int k = 0; k < K; ++k)
        if (!refs[k].on_device) host_bytes += sizeof(float) * (size_t)refs[k].kh * refs[k].kw * F;
    const size_t desc_bytes = sizeof(SrcDesc) * (size_t)K;
    int e = 0;
    if (!e) e = pinned_reserve(*c, desc_bytes);
    if (!e) e = dev_reserve(c->desc, desc_bytes);
    if (!e && host_bytes) e = dev_reserve(c->stage, host_bytes);
    if (!e && cudaEventSynchronize(c->pinned_free) != cudaSuccess) e = fail(FFTCONV_ERR_CUDA, "event sync failed");
    if (!e) {
        SrcDesc* h_desc = reinterpret_cast<SrcDesc*>(c->pinned);
        size_t off = 0;
        for (int k = 0; k < K && !e; ++k) {
            h_desc[k].rows = refs[k].kh; h_desc[k].cols = refs[k].kw;
            if (refs[k].on_device) { h_desc[k].ptr = refs[k].ptr; continue; }
            const size_t nb = sizeof(float) * (size_t)refs[k].kh * refs[k].kw * F;
            h_desc[k].ptr = reinterpret_cast<const float*>(reinterpret_cast<char*>(c->stage.p) + off);
            if (cudaMemcpyAsync(reinterpret_cast<char*>(c->stage.p) + off, refs[k].ptr, nb, cudaMemcpyHostToDevice, st) != cudaSuccess)
                e = fail(FFTCONV_ERR_CUDA, "kernel upload failed");
            off += nb;
        }
        if (!e && cudaMemcpyAsync(c->desc.p, c->pinned, desc_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
            e = fail(FFTCONV_ERR_CUDA, "descriptor upload failed");
    }
    if (!e) {
        OsKArgs a{};
        a.descs = (const SrcDesc*)c->desc.p; a.nk = K; a.F = F; a.img = b->A; a.NKS = g.NKS; a.KC = g.KC; a.flip = 0;
        dim3 grid(ntblk * OS_TM / OS_KSL, g.NKS * g.KC);
        ProfScope ps(PK_OS_KERN, st);
        if (g.NFK == 1) os_kern_fft<1><<<grid, 256, os_kern_smem(1), st>>>(a);
        else os_kern_fft<2><<<grid, 256, os_kern_smem(2), st>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) e = fail(FFTCONV_ERR_CUDA, "bank transform failed");
        if (!e && in_workspace) {
            if (cudaEventRecord(c->pinned_free, st) != cudaSuccess) e = fail(FFTCONV_ERR_CUDA, "event record failed");
        } else if (!e && cudaStreamSynchronize(st) != cudaSuccess) {
            e = fail(FFTCONV_ERR_CUDA, "bank transform failed");
        }
    }
    if (e) { if (b->owns) { cudaFree(b->A); cudaFree(b->khw); } return e; }
    *out = b.release();
    return 0;
}

int fftconv_bank_create(int K, const float* const* kernels, const int* kh, const int* kw, const int* kf,
                        const unsigned char* kernel_on_device, int F, int device, void* stream, fftconv_bank** out) {
    return bank_create_impl(K, kernels, kh, kw, kf, kernel_on_device, F, device, stream, out, false);
}

int fftconv_bank_info(const fftconv_bank* b, int* K, int* F, int* maxKH, int* maxKW, long long* bytes) {
    if (!b) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    if (K) *K = b->K;
    if (F) *F = b->F;
    if (maxKH) *maxKH = b->maxkh;
    if (maxKW) *maxKW = b->maxkw;
    if (bytes) *bytes = (long long)b->bytes;
    return 0;
}

// nimg device-resident images [nimg][F][W][H] against a prepared bank (caller holds the device lock and has selected the device: CtxScope);
// outs: nimg * K planes, image-major.
static int bank_conv_images(const fftconv_bank* b, const float* d_data, int nimg, int H, int W, float* const* outs,
                            int out_on_device, const fftconv_options& o, cudaStream_t st) {
    Ctx* c;
    if (int e = ctx_get(b->device, &c)) return e;
    const int FH = fftconv_fft_size16(H + b->maxkh - 1), FW = fftconv_fft_size16(W + b->maxkw - 1);
    ConvArgs a;
    a.d_spec = nullptr; a.CH = FH / 2 + 1; a.FW = FW; a.F = b->F; a.K = b->K;
    a.kernels = nullptr; a.outs = outs; a.out_on_device = out_on_device != 0;
    a.opt = o;
    a.d_raw = d_data; a.rawH = H; a.rawW = W; a.nimg = nimg;
    a.bankA = b->A; a.bank_maxkh = b->maxkh; a.bank_maxkw = b->maxkw;
    return run_conv(*c, a, st);
}

int fftconv_bank_conv(const fftconv_bank* b, const float* data, int data_on_device, int H, int W,
                      float* const* outs, int out_on_device, const fftconv_options* opt, void* stream) {
    g_err.clear();
    if (!b || !data || !outs || H <= 0 || W <= 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    const fftconv_options o = opt ? *opt : fftconv_options{};
    if (o.correlate) return fail(FFTCONV_ERR_UNSUPPORTED, "prepared banks do not serve correlation mode");
    const int F = b->F, K = b->K;
    const int FH = fftconv_fft_size16(H + b->maxkh - 1), FW = fftconv_fft_size16(W + b->maxkw - 1);
    for (int k = 0; k < K; ++k)
        if (!outs[k]) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    CtxScope cs(b->device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    cudaStream_t st = (cudaStream_t)stream;
    const float* d_data = data;
    if (!data_on_device) {
        const size_t bytes = sizeof(float) * (size_t)H * W * F;
        if (int e = dev_reserve(c->ddata, bytes)) return e;
        CU(cudaMemcpyAsync(c->ddata.p, data, bytes, cudaMemcpyHostToDevice, st));
        d_data = (const float*)c->ddata.p;
    }
    return bank_conv_images(b, d_data, 1, H, W, outs, out_on_device, o, st);
}

int fftconv_bank_conv_max(const fftconv_bank* b, const float* data, int data_on_device, int H, int W,
                          fftconv_peak* peaks, int peaks_on_device, void* stream) {
    g_err.clear();
    if (!b || !data || !peaks || H <= 0 || W <= 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    const int F = b->F, K = b->K;
    const int FH = fftconv_fft_size16(H + b->maxkh - 1), FW = fftconv_fft_size16(W + b->maxkw - 1);
    if (FH > 65535 || FW > 65535) return fail(FFTCONV_ERR_UNSUPPORTED, "plane too large for packed peak positions");
    CtxScope cs(b->device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    cudaStream_t st = (cudaStream_t)stream;
    const float* d_data = data;
    if (!data_on_device) {
        const size_t bytes = sizeof(float) * (size_t)H * W * F;
        if (int e = dev_reserve(c->ddata, bytes)) return e;
        CU(cudaMemcpyAsync(c->ddata.p, data, bytes, cudaMemcpyHostToDevice, st));
        d_data = (const float*)c->ddata.p;
    }
    if (int e = dev_reserve(c->osPeaks, (sizeof(unsigned long long) + sizeof(fftconv_peak)) * (size_t)K)) return e;
    unsigned long long* keys = (unsigned long long*)c->osPeaks.p;
    fftconv_peak* d_peaks = peaks_on_device ? peaks : reinterpret_cast<fftconv_peak*>(keys + K);
    CU(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)K, st));
    ConvArgs a;
    a.d_spec = nullptr; a.CH = FH / 2 + 1; a.FW = FW; a.F = F; a.K = K;
    a.kernels = nullptr; a.outs = nullptr; a.out_on_device = true;
    a.opt = fftconv_options{};
    a.d_raw = d_data; a.rawH = H; a.rawW = W;
    a.bankA = b->A; a.bank_maxkh = b->maxkh; a.bank_maxkw = b->maxkw;
    a.peak_keys = keys; a.bank_khw = b->khw;
    if (int e = run_conv(*c, a, st)) return e;
    static_assert(sizeof(fftconv_peak) == sizeof(fftconv_peak_dev), "peak layouts differ");
    os_peak_finalize<<<(K + 255) / 256, 256, 0, st>>>(keys, K, reinterpret_cast<fftconv_peak_dev*>(d_peaks));
    LAUNCH_CHECK();
    if (!peaks_on_device) {
        CU(cudaMemcpyAsync(peaks, d_peaks, sizeof(fftconv_peak) * (size_t)K, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

// shared by fftconv_bank_conv_detect (mode 1) and fftconv_bank_conv_topk (mode 3)
static int bank_detect_impl(const fftconv_bank* b, const float* data, int data_on_device, int H, int W, const float* bias,
                            int mode, float threshold, int nsel, fftconv_peak* dets, int* counts, int out_on_device, void* stream) {
    g_err.clear();
    if (!b || !data || !dets || H <= 0 || W <= 0 || nsel <= 0 || nsel > 4096)
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    const int F = b->F, K = b->K;
    const int FH = fftconv_fft_size16(H + b->maxkh - 1), FW = fftconv_fft_size16(W + b->maxkw - 1);
    if (FH > 65535 || FW > 65535) return fail(FFTCONV_ERR_UNSUPPORTED, "plane too large for packed peak positions");
    OsCfg g;
    if (!os_config(F, FH, FW, b->maxkh, b->maxkw, g)) return fail(FFTCONV_ERR_UNSUPPORTED, "plane outside the range of the overlap-save path");
    CtxScope cs(b->device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    cudaStream_t st = (cudaStream_t)stream;
    const float* d_data = data;
    if (!data_on_device) {
        const size_t bytes = sizeof(float) * (size_t)H * W * F;
        if (int e = dev_reserve(c->ddata, bytes)) return e;
        CU(cudaMemcpyAsync(c->ddata.p, data, bytes, cudaMemcpyHostToDevice, st));
        d_data = (const float*)c->ddata.p;
    }
    // candidate slots per template: one per lane of the candidate pass (64 per tile), at least 2048, at least 4 per pick
    const int cap = std::max(std::max(g.NT * 64, 2048), 4 * nsel);
    const size_t keys_b = sizeof(unsigned long long) * (size_t)K * cap;
    const size_t peaks_b = sizeof(fftconv_peak) * (size_t)K * nsel;
    const size_t total = keys_b + peaks_b + (sizeof(unsigned int) + sizeof(int) + 2 * sizeof(float)) * (size_t)K + 256;
    if (int e = dev_reserve(c->osPeaks, total)) return e;
    char* base = (char*)c->osPeaks.p;
    unsigned long long* keys = (unsigned long long*)base;
    fftconv_peak* d_dets = out_on_device ? dets : (fftconv_peak*)(base + keys_b);
    unsigned int* count = (unsigned int*)(base + keys_b + peaks_b);
    int* d_counts = (int*)(count + K);
    float* thr = (float*)(d_counts + K);
    float* d_bias = thr + K;
    if (bias) {                                       // K floats from the host through the pinned staging buffer
        if (int e = pinned_reserve(*c, sizeof(float) * (size_t)K)) return e;
        CU(cudaEventSynchronize(c->pinned_free));
        memcpy(c->pinned, bias, sizeof(float) * (size_t)K);
        CU(cudaMemcpyAsync(d_bias, c->pinned, sizeof(float) * (size_t)K, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->pinned_free, st));
    }
    if (mode == 1) {
        os_fill_f32<<<(K + 255) / 256, 256, 0, st>>>(thr, K, threshold);
        LAUNCH_CHECK();
    }
    ConvArgs a;
    a.d_spec = nullptr; a.CH = FH / 2 + 1; a.FW = FW; a.F = F; a.K = K;
    a.kernels = nullptr; a.outs = nullptr; a.out_on_device = true;
    a.opt = fftconv_options{};
    a.d_raw = d_data; a.rawH = H; a.rawW = W;
    a.bankA = b->A; a.bank_maxkh = b->maxkh; a.bank_maxkw = b->maxkw; a.bank_khw = b->khw;
    a.det.mode = mode; a.det.cap = cap; a.det.k = nsel; a.det.keys = keys; a.det.count = count; a.det.thr = thr;
    a.det.bias = bias ? d_bias : nullptr;
    if (int e = run_conv(*c, a, st)) return e;
    static_assert(sizeof(fftconv_peak) == sizeof(fftconv_peak_dev), "peak layouts differ");
    os_det_select<<<K, 256, 0, st>>>(keys, count, cap, nsel, reinterpret_cast<fftconv_peak_dev*>(d_dets), nullptr, d_counts);
    LAUNCH_CHECK();
    // host results (and the overflow check of the top-k mode) need the counts
    std::vector<int> h_counts;
    if (!out_on_device || mode == 3) {
        h_counts.resize(K);
        CU(cudaMemcpyAsync(h_counts.data(), d_counts, sizeof(int) * (size_t)K, cudaMemcpyDeviceToHost, st));
        if (!out_on_device) CU(cudaMemcpyAsync(dets, d_dets, peaks_b, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (mode == 3)
            for (int k = 0; k < K; ++k)
                if (h_counts[k] > cap)
                    return fail(FFTCONV_ERR_UNSUPPORTED, "top-k: template %d has %d responses above its candidate bound (capacity %d): "
                                "plateau-like response, use fftconv_bank_conv_detect with a threshold", k, h_counts[k], cap);
    }
    if (counts) {
        if (out_on_device) CU(cudaMemcpyAsync(counts, d_counts, sizeof(int) * (size_t)K, cudaMemcpyDeviceToDevice, st));
        else memcpy(counts, h_counts.data(), sizeof(int) * (size_t)K);
    }
    return 0;
}

int fftconv_bank_conv_detect(const fftconv_bank* b, const float* data, int data_on_device, int H, int W, const float* bias,
                             float threshold, int max_per_template, fftconv_peak* dets, int* counts, int out_on_device, void* stream) {
    return bank_detect_impl(b, data, data_on_device, H, W, bias, 1, threshold, max_per_template, dets, counts, out_on_device, stream);
}

int fftconv_bank_conv_topk(const fftconv_bank* b, const float* data, int data_on_device, int H, int W, const float* bias,
                           int k, fftconv_peak* dets, int out_on_device, void* stream) {
    if (k > 64) return fail(FFTCONV_ERR_INVALID_INPUT, "top-k: k must be at most 64");
    return bank_detect_impl(b, data, data_on_device, H, W, bias, 3, 0.f, k, dets, nullptr, out_on_device, stream);
}

void fftconv_bank_destroy(fftconv_bank* b) {
    if (!b) return;
    CtxScope cs(b->device, nullptr);
    if (b->owns) {
        cudaDeviceSynchronize();          // calls that read the bank may still be running on any stream
        cudaFree(b->A);
        cudaFree(b->khw);
    }
    delete b;
}

int fftconv_conv_bank(const fftconv_float2* d_spec, int CH, int FW, int F, int K, const float* d_bank, int kh, int kw,
                      float* d_out, const fftconv_options* opt, int device, void* stream) {
    g_err.clear();
    if (!d_bank || !d_out || K < 0 || kh <= 0 || kw <= 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    const int FH = (CH - 1) * 2;
    fftconv_options o = opt ? *opt : fftconv_options{};
    const int crop_h = o.crop_h > 0 ? o.crop_h : FH, crop_w = o.crop_w > 0 ? o.crop_w : FW;
    const size_t plane = (size_t)crop_w * (o.out_ld > 0 ? o.out_ld : crop_h);
    std::vector<const float*> kp(K);
    std::vector<float*> op(K);
    std::vector<int> khs(K, kh), kws(K, kw);
    std::vector<unsigned char> od(K, 1);
    for (int k = 0; k < K; ++k) {
        kp[k] = d_bank + (size_t)k * kh * kw * F;
        op[k] = d_out + (size_t)k * plane;
    }
    return conv_impl(d_spec, CH, FW, F, K, kp.data(), khs.data(), kws.data(), nullptr, od.data(), op.data(), 1,
                     nullptr, 0, &o, device, stream);
}

// ---- plans: the persistent graph schedule over the kernel bank (the role of the per-stream ConvPlans of
// src/cudaConvFFTDataStreams.cu:292-328,338-469).  A plan fixes the device buffers of a repeated call -- image (or
// spectrum), bank, output planes -- runs it once eagerly (which also sizes every cached scratch buffer) and captures the
// whole launch sequence (data transform, template transforms, per-bin GEMM, inverse; both internal streams) into ONE
// CUDA graph; every later execution is a single cudaGraphLaunch.  Buffer CONTENTS may change between executions.
struct fftconv_plan {
    int device;
    const float* d_data; int H, W, F, KH, KW;      // d_data == nullptr: the plan starts from the spectrum
    fftconv_float2* d_spec; int CH, FW;
    int K, kh, kw;
    const float* d_bank; float* d_out;
    fftconv_options opt;
    std::vector<KernelRef> refs;
    std::vector<float*> outs;
    char* pin = nullptr; size_t pin_cap = 0;
    cudaStream_t cap_stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    long long gen = -1;
    size_t nodes = 0;
};

static int plan_enqueue(Ctx& c, fftconv_plan* p, cudaStream_t st) {
    if (p->d_data)
        if (int e = run_fft_data(c, p->d_data, p->H, p->W, p->F, (p->CH - 1) * 2, p->FW, PAD_ZERO, 0, 0, (cpx*)p->d_spec, st)) return e;
    ConvArgs a;
    a.d_spec = (const cpx*)p->d_spec; a.CH = p->CH; a.FW = p->FW; a.F = p->F; a.K = p->K;
    a.kernels = p->refs.data(); a.outs = p->outs.data(); a.out_on_device = true;
    a.opt = p->opt;
    return run_conv(c, a, st);
}

// (re)capture; the caller holds the device lock.  `warm`: run once eagerly first so that no allocation happens in capture.
static int plan_capture(Ctx& c, fftconv_plan* p, cudaStream_t user_stream, bool warm) {
    if (warm) {
        if (int e = plan_enqueue(c, p, user_stream)) return e;
        CU(cudaStreamSynchronize(user_stream));
    }
    if (p->exec) { cudaGraphExecDestroy(p->exec); p->exec = nullptr; }
    if (p->graph) { cudaGraphDestroy(p->graph); p->graph = nullptr; }
    const long long gen0 = g_scratch_gen.load();
    c.plan_pin = p->pin; c.plan_pin_cap = p->pin_cap; c.plan_pin_off = 0;
    g_capturing = true;
    int e = 0;
    if (cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        e = fail(FFTCONV_ERR_CUDA, "cudaStreamBeginCapture failed");
    if (!e) {
        e = plan_enqueue(c, p, p->cap_stream);
        cudaGraph_t gr = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(p->cap_stream, &gr);
        if (!e && ce != cudaSuccess) e = fail(FFTCONV_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        if (e && gr) cudaGraphDestroy(gr);
        if (!e) p->graph = gr;
    }
    g_capturing = false;
    c.plan_pin = nullptr; c.plan_pin_cap = c.plan_pin_off = 0;
    if (e) { cudaGetLastError(); return e; }
    if (g_scratch_gen.load() != gen0) return fail(FFTCONV_ERR_CUDA, "scratch moved during plan capture");
    CU(cudaGraphInstantiate(&p->exec, p->graph, 0));
    CU(cudaGraphGetNodes(p->graph, nullptr, &p->nodes));
    p->gen = gen0;
    return 0;
}

int fftconv_plan_create(const float* d_data, int H, int W, int F, int KH, int KW, fftconv_float2* d_spec,
                        int K, const float* d_bank, int kh, int kw, float* d_out, const fftconv_options* opt,
                        int device, void* stream, fftconv_plan** out) {
    g_err.clear();
    if (!out) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    *out = nullptr;
    if (!d_spec || !d_bank || !d_out || H <= 0 || W <= 0 || F <= 0 || KH <= 0 || KW <= 0 || K <= 0 || kh <= 0 || kw <= 0)
        return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    const int FH = fftconv_fft_size16(H + KH - 1), FW = fftconv_fft_size16(W + KW - 1);
    if (kh > FH || kw > FW) return fail(FFTCONV_ERR_KERNEL_SHAPE, "%s", kMsgKernelShape);
    std::unique_ptr<fftconv_plan> p(new fftconv_plan{});
    p->device = device; p->d_data = d_data; p->H = H; p->W = W; p->F = F; p->KH = KH; p->KW = KW;
    p->d_spec = d_spec; p->CH = FH / 2 + 1; p->FW = FW; p->K = K; p->kh = kh; p->kw = kw; p->d_bank = d_bank; p->d_out = d_out;
    p->opt = opt ? *opt : fftconv_options{};
    const int crop_h = p->opt.crop_h > 0 ? p->opt.crop_h : FH, crop_w = p->opt.crop_w > 0 ? p->opt.crop_w : FW;
    const size_t plane = (size_t)crop_w * (p->opt.out_ld > 0 ? p->opt.out_ld : crop_h);
    p->refs.resize(K); p->outs.resize(K);
    for (int k = 0; k < K; ++k) {
        p->refs[k] = KernelRef{d_bank + (size_t)k * kh * kw * F, kh, kw, true};
        p->outs[k] = d_out + (size_t)k * plane;
    }
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    p->pin_cap = (sizeof(SrcDesc) + sizeof(int2) + sizeof(int) + sizeof(float*)) * (size_t)K + sizeof(float*) * (size_t)F + 4096;
    CU(cudaMallocHost((void**)&p->pin, p->pin_cap));
    if (cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaFreeHost(p->pin);
        return fail(FFTCONV_ERR_CUDA, "cudaStreamCreate failed");
    }
    if (int e = plan_capture(*cs.c, p.get(), (cudaStream_t)stream, true)) {
        cudaStreamDestroy(p->cap_stream); cudaFreeHost(p->pin);
        return e;
    }
    *out = p.release();
    return 0;
}

int fftconv_plan_execute(fftconv_plan* p, void* stream) {
    g_err.clear();
    if (!p) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    CtxScope cs(p->device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    if (p->gen != g_scratch_gen.load())               // another call grew (moved) the cached scratch: capture again
        if (int e = plan_capture(*cs.c, p, (cudaStream_t)stream, true)) return e;
    CU(cudaGraphLaunch(p->exec, (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int fftconv_plan_info(const fftconv_plan* p, int* graph_nodes, int* path) {
    if (!p) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    if (graph_nodes) *graph_nodes = (int)p->nodes;
    if (path) *path = choose_path(p->opt, p->F, (p->CH - 1) * 2, p->FW, std::min(p->kh, (p->CH - 1) * 2), std::min(p->kw, p->FW), p->K);
    return 0;
}

void fftconv_plan_destroy(fftconv_plan* p) {
    if (!p) return;
    {
        CtxScope cs(p->device, nullptr);
        cudaDeviceSynchronize();
        if (p->exec) cudaGraphExecDestroy(p->exec);
        if (p->graph) cudaGraphDestroy(p->graph);
        if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
        if (p->pin) cudaFreeHost(p->pin);
    }
    delete p;
}

int fftconv_modulate_and_normalize(fftconv_float2* d_a, const fftconv_float2* d_b, long long n, int device, void* stream) {
    g_err.clear();
    if (!d_a || !d_b || n < 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    if (n == 0) return 0;
    CtxScope cs(device, (cudaStream_t)stream);
    if (cs.err) return cs.err;
    Ctx* c = cs.c;
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)c->sm_count * 8);
    modulate_and_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((cpx*)d_a, (const cpx*)d_b, n, 1.0f / (float)n);
    LAUNCH_CHECK();
    return 0;
}

long long fftconv_launch_count(void) { return g_launches.load(); }

void fftconv_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
}

int fftconv_profile_kinds(void) { return PK_COUNT; }
const char* fftconv_profile_name(int kind) { return (kind >= 0 && kind < PK_COUNT) ? kProfNames[kind] : ""; }

int fftconv_profile_read(int kind, double* total_ms, long long* launches, int reset) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms = 0.0;
    long long n = 0;
    for (auto& r : g_prof) {
        if (r.kind != kind) continue;
        if (cudaEventSynchronize(r.b) != cudaSuccess) return fail(FFTCONV_ERR_CUDA, "profile event sync failed");
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return fail(FFTCONV_ERR_CUDA, "profile event read failed");
        ms += t; ++n;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (reset) {
        std::vector<ProfRec> keep;
        for (auto& r : g_prof) {
            if (r.kind == kind) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
            else keep.push_back(r);
        }
        g_prof.swap(keep);
    }
    return 0;
}

long long fftconv_workspace_bytes(int device) {
    Ctx* cp;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_ctx.find(device);
        if (it == g_ctx.end()) return 0;
        cp = &it->second;
    }
    Ctx& c = *cp;
    std::lock_guard<std::recursive_mutex> lkc(c.mu);
    size_t s = c.T.cap + c.Z.cap + c.stage.cap + c.desc.cap + c.outstage.cap + c.dspec.cap + c.ddata.cap +
               c.priv.cap + c.Ag.cap + c.Wg.cap + c.osA.cap + c.osB.cap + c.osP.cap +
               c.osPlane.cap + c.osZ.cap + c.osPeaks.cap + c.bpS.cap + c.batchA.cap;
    for (auto& kv : c.tw) s += sizeof(cpx) * (size_t)kv.first;
    return (long long)s;
}

void fftconv_release(void) {
    std::vector<Ctx*> all;
    { std::lock_guard<std::mutex> lk(g_mu); for (auto& kv : g_ctx) all.push_back(&kv.second); }
    int prev = -1;
    cudaGetDevice(&prev);
    for (Ctx* cp : all) {
        Ctx& c = *cp;
        std::lock_guard<std::recursive_mutex> lkc(c.mu);
        if (!c.inited) continue;
        cudaSetDevice(c.dev);
        cudaDeviceSynchronize();
        for (auto& t : c.tw) cudaFree(t.second);
        for (auto& t : c.ipmap) cudaFree(t.second);
        for (DevBuf* b : {&c.T, &c.Z, &c.stage, &c.desc, &c.outstage, &c.dspec, &c.ddata, &c.priv, &c.Ag, &c.Wg,
                          &c.osA, &c.osB, &c.osP, &c.osPlane, &c.osZ, &c.osPeaks, &c.bpS, &c.batchA})
            if (b->p) cudaFree(b->p);
        if (c.pinned) cudaFreeHost(c.pinned);
        if (c.bounce) cudaFreeHost(c.bounce);
        for (auto& e : c.evb) if (e) cudaEventDestroy(e);
        if (c.pinned_free) cudaEventDestroy(c.pinned_free);
        for (auto& e : c.ev) if (e) cudaEventDestroy(e);
        if (c.side) cudaStreamDestroy(c.side);
        if (c.side2) cudaStreamDestroy(c.side2);
        for (auto& e : c.evf) if (e) cudaEventDestroy(e);
        if (c.last_use) cudaEventDestroy(c.last_use);
        // the entry stays in the map (another thread may hold a pointer to it): reset it to its pristine state
        c.tw.clear(); c.ipmap.clear();
        for (DevBuf* b : {&c.T, &c.Z, &c.stage, &c.desc, &c.outstage, &c.dspec, &c.ddata, &c.priv, &c.Ag, &c.Wg,
                          &c.osA, &c.osB, &c.osP, &c.osPlane, &c.osZ, &c.osPeaks, &c.bpS, &c.batchA})
            *b = DevBuf{};
        c.pinned = nullptr; c.pinned_cap = 0; c.pinned_free = nullptr; c.side = c.side2 = nullptr;
        c.bounce = nullptr; c.bounce_cap = 0;
        for (auto& e : c.evb) e = nullptr;
        for (auto& e : c.ev) e = nullptr;
        for (auto& e : c.evf) e = nullptr;
        c.last_use = nullptr; c.last_stream = nullptr; c.used = false; c.spec_ready = nullptr;
        c.inited = false;
    }
    if (prev >= 0) cudaSetDevice(prev);
}

int fftconv_query_path(int H, int W, int F, int maxKH, int maxKW, int K, const fftconv_options* opt,
                       int* radices_h, int* radices_w) {
    g_err.clear();
    if (H <= 0 || W <= 0 || F <= 0 || maxKH <= 0 || maxKW <= 0 || K < 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid data input");
    const int FH = fftconv_fft_size16(H + maxKH - 1), FW = fftconv_fft_size16(W + maxKW - 1);
    const fftconv_options o = opt ? *opt : fftconv_options{};
    const int path = choose_path(o, F, FH, FW, std::min(maxKH, FH), std::min(maxKW, FW), K);
    for (int side = 0; side < 2; ++side) {
        int* out = side ? radices_w : radices_h;
        if (!out) continue;
        for (int i = 0; i < 8; ++i) out[i] = 0;
        IpPlan p;
        if (path == PATH_BIGPLANE && make_ip_plan(side ? FW : FH, p))
            for (int i = 0; i < p.ns && i < 8; ++i) out[i] = p.R[i];
    }
    return path;
}

int fftconv_spectrum_ready_event(int device, void* cuda_event) {
    DeviceGuard dg(device);
    if (!dg.ok) return fail(FFTCONV_ERR_CUDA, "cannot select device %d", device);
    Ctx* c;
    if (int e = ctx_get(device, &c)) return e;
    std::lock_guard<std::recursive_mutex> lkc(c->mu);
    c->spec_ready = reinterpret_cast<cudaEvent_t>(cuda_event);
    return 0;
}

// ---- peer spectrum
int fftconv_peer_alloc(size_t bytes, int device, void** ptr, unsigned char handle[64]) {
    g_err.clear();
    if (!ptr || !handle || bytes == 0) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        cudaFree(p);
        return fail(FFTCONV_ERR_CUDA, "CUDA IPC export failed: %s", cudaGetErrorString(e));
    }
    CU(cudaDeviceSynchronize());
    memcpy(handle, &h, 64);
    *ptr = p;
    return 0;
}
int fftconv_peer_open(const unsigned char handle[64], int device, void** ptr) {
    g_err.clear();
    if (!ptr || !handle) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int fftconv_peer_close(void* mapped_ptr, int device) {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    CU(cudaIpcCloseMemHandle(mapped_ptr));
    return 0;
}
int fftconv_peer_free(void* ptr, int device) {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    CU(cudaFree(ptr));
    return 0;
}
int fftconv_peer_signal(unsigned long long* flag, unsigned long long value, int device, void* stream) {
    DeviceGuard guard(device);
    if (!guard.ok || !flag) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    LAUNCH_CHECK();
    return 0;
}
int fftconv_peer_wait_all(const unsigned long long* flags, int n, unsigned long long value, int device, void* stream) {
    DeviceGuard guard(device);
    if (!guard.ok || !flags || n <= 0 || n > 1024) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    peer_wait_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(flags, n, value);
    LAUNCH_CHECK();
    return 0;
}
int fftconv_peer_wait(const unsigned long long* flag, unsigned long long value, int device, void* stream) {
    return fftconv_peer_wait_all(flag, 1, value, device, stream);
}
int fftconv_peer_pull(void* dst, const void* src_mapped, size_t bytes, int device, void* stream) {
    DeviceGuard guard(device);
    if (!guard.ok || !dst || !src_mapped) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    if (((uintptr_t)dst | (uintptr_t)src_mapped) & 15) {
        CU(cudaMemcpyAsync(dst, src_mapped, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    const size_t n16 = bytes / 16;
    const int ntail = (int)(bytes - n16 * 16);
    Ctx* c;
    if (int e = ctx_get(device, &c)) return e;
    peer_pull_kernel<<<c->sm_count * 2, 512, 0, (cudaStream_t)stream>>>((uint4*)dst, (const uint4*)src_mapped, n16,
                                                                      (unsigned char*)dst + n16 * 16,
                                                                      (const unsigned char*)src_mapped + n16 * 16, ntail);
    LAUNCH_CHECK();
    return 0;
}
int fftconv_peer_allgather(void* const* bases, int n, int rank, const unsigned long long* offs, unsigned long long flag_off,
                           unsigned long long step, int device, void* stream) {
    g_err.clear();
    if (!bases || !offs || n < 1 || n > 16 || rank < 0 || rank >= n) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    PeerAgArgs a{};
    for (int p = 0; p < n; ++p) {
        if (!bases[p] || (offs[p] & 15) || offs[p + 1] < offs[p]) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
        a.base[p] = (unsigned char*)bases[p];
        a.off[p] = offs[p];
    }
    if (offs[n] & 15) return fail(FFTCONV_ERR_INVALID_INPUT, "Invalid input to MEX file.");
    a.off[n] = offs[n];
    a.flag_off = flag_off; a.step = step; a.n = n; a.rank = rank;
    // the rank's own slice is complete (stream order): raise its ready flag and its own acknowledgement slot
    unsigned long long* ready = reinterpret_cast<unsigned long long*>(a.base[rank] + flag_off);
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ready, step);
    LAUNCH_CHECK();
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ready + 1 + rank, step);
    LAUNCH_CHECK();
    if (n > 1) {
        a.bpg = std::max(1, 128 / n);                       // all blocks resident at once: a waiting group never starves another
        peer_ag_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
        LAUNCH_CHECK();
        peer_allgather_kernel<<<n * a.bpg, 512, 0, (cudaStream_t)stream>>>(a);
        LAUNCH_CHECK();
    }
    return 0;
}

int fftconv_peer_status(int device) {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FFTCONV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    CU(cudaDeviceSynchronize());
    unsigned int n = 0;
    CU(cudaMemcpyFromSymbol(&n, g_peer_timeouts, sizeof n));
    if (n) return fail(FFTCONV_ERR_CUDA, "%u peer wait(s) timed out on device %d", n, device);
    return 0;
}

const char* fftconv_last_error(void) { return g_err.c_str(); }
const char* fftconv_version(void) { return "fftconv-b200 0.1.0 sm_100a"; }

}  // extern "C"
