// fftconv_bench — stand-alone C++ driver of the C ABI (include/fftconv.h); no MATLAB, no Python.
//
//   fftconv_bench [--config c1|c2|c3|c3s|c4|c4s|c5|c5s] [--H h --W w --F f --kh a --kw b --K k --N n] [--iters n]
//                 [--host] [--check n] [--device d]
//
// Builds the named synthetic workload, runs cudaConvolutionFFT-equivalent calls through
// libfftconv.so and prints conv outputs/s.  --host passes host buffers (MEX-style call, copies
// included); default keeps bank and planes resident on the device.  --check n verifies n sampled
// output pixels per template against a float64 direct convolution computed here on the CPU
// (verification of the driver's own run, not a compute path).
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/fftconv.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define FC(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "fftconv error %d: %s\n", r_, fftconv_last_error()); return 3; } } while (0)

struct Cfg { int H, W, F, kh, kw, K; const char* name; int N = 1; };   // N > 1: batched images (fftconv_conv_batch)

int main(int argc, char** argv) {
    Cfg c{256, 256, 31, 16, 16, 1000, "c2"};
    int iters = 10, check = 0, device = 0;
    bool host = false, pyramid = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? atoi(argv[++i]) : 0; };
        if (a == "--config" && i + 1 < argc) {
            std::string n = argv[++i];
            if (n == "c1") c = Cfg{64, 8, 5, 10, 4, 10, "c1"};
            else if (n == "c2") c = Cfg{256, 256, 31, 16, 16, 1000, "c2"};
            else if (n == "c3") c = Cfg{4096, 4096, 1, 512, 512, 64, "c3"};
            else if (n == "c3s") c = Cfg{1024, 1024, 1, 128, 128, 16, "c3s"};
            else if (n == "c4") c = Cfg{512, 512, 32, 32, 32, 256, "c4", 64};
            else if (n == "c4s") c = Cfg{512, 512, 32, 32, 32, 256, "c4s", 8};
            else if (n == "c5") { c = Cfg{256, 256, 31, 16, 16, 20000, "c5"}; pyramid = true; }
            else if (n == "c5s") { c = Cfg{256, 256, 31, 16, 16, 2000, "c5s"}; pyramid = true; }
            else { fprintf(stderr, "unknown config %s\n", n.c_str()); return 1; }
        } else if (a == "--H") c.H = next(); else if (a == "--W") c.W = next(); else if (a == "--F") c.F = next();
        else if (a == "--kh") c.kh = next(); else if (a == "--kw") c.kw = next(); else if (a == "--K") c.K = next();
        else if (a == "--N") c.N = next();
        else if (a == "--iters") iters = next(); else if (a == "--check") check = next();
        else if (a == "--device") device = next(); else if (a == "--host") host = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 1; }
    }
    const int FH = fftconv_fft_size16(c.H + c.kh - 1), FW = fftconv_fft_size16(c.W + c.kw - 1), CH = FH / 2 + 1;
    if (c.N < 1) c.N = 1;
    if (c.N > 1 && (host || check > 0)) { fprintf(stderr, "--host / --check are single-image options\n"); return 1; }
    const size_t nd = (size_t)c.H * c.W * c.F * c.N, nk1 = (size_t)c.kh * c.kw * c.F, plane = (size_t)FH * FW;
    printf("%s  %s: %d x data %dx%dx%d, %d templates %dx%dx%d, plane %dx%d\n", fftconv_version(), c.name, c.N, c.H, c.W, c.F,
           c.K, c.kh, c.kw, c.F, FH, FW);

    std::mt19937 rng(2);
    std::uniform_real_distribution<float> U(0.f, 0.2f);
    std::normal_distribution<float> Nn(0.f, 0.05f);
    std::vector<float> h_data(nd), h_bank(nk1 * c.K);
    for (auto& v : h_data) v = U(rng);
    for (auto& v : h_bank) v = Nn(rng);

    CK(cudaSetDevice(device));
    if (pyramid) {
        // BASELINE config 5 on one GPU: 10-level pyramid (side 256 * 2^(-l/5)) x one PREPARED bank
        // (fftconv_bank_create: the template spectra are transformed once and serve every level)
        int side[10];
        size_t out_floats = 0;
        for (int l = 0; l < 10; ++l) {
            side[l] = (int)std::lround(256.0 * std::pow(2.0, -l / 5.0));
            out_floats = std::max(out_floats, (size_t)fftconv_fft_size16(side[l] + c.kh - 1) * fftconv_fft_size16(side[l] + c.kw - 1));
        }
        float *d_lv, *d_bk, *d_o;
        CK(cudaMalloc(&d_lv, (size_t)256 * 256 * c.F * 4)); CK(cudaMalloc(&d_bk, nk1 * c.K * 4));
        CK(cudaMalloc(&d_o, out_floats * c.K * 4));
        CK(cudaMemcpy(d_lv, h_data.data(), (size_t)256 * 256 * c.F * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_bk, h_bank.data(), nk1 * c.K * 4, cudaMemcpyHostToDevice));
        std::vector<const float*> kp(c.K);
        std::vector<int> khs(c.K, c.kh), kws(c.K, c.kw);
        std::vector<unsigned char> ond(c.K, 1);
        for (int k = 0; k < c.K; ++k) kp[k] = d_bk + nk1 * k;
        cudaStream_t st;
        CK(cudaStreamCreate(&st));
        fftconv_bank* bank = nullptr;
        auto t0 = std::chrono::steady_clock::now();
        FC(fftconv_bank_create(c.K, kp.data(), khs.data(), kws.data(), nullptr, ond.data(), c.F, device, st, &bank));
        const double prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::vector<float*> op(c.K);
        double nout = 0;
        auto run = [&]() -> int {
            nout = 0;
            for (int l = 0; l < 10; ++l) {          // level l reuses the top-left side x side block of the buffer as its own image
                const int FHl = fftconv_fft_size16(side[l] + c.kh - 1), FWl = fftconv_fft_size16(side[l] + c.kw - 1);
                for (int k = 0; k < c.K; ++k) op[k] = d_o + (size_t)FHl * FWl * k;
                if (int r = fftconv_bank_conv(bank, d_lv, 1, side[l], side[l], op.data(), 1, nullptr, st)) return r;
                nout += (double)c.K * FHl * FWl;
            }
            return 0;
        };
        FC(run());
        CK(cudaStreamSynchronize(st));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) FC(run());
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
        printf("pyramid x prepared bank: bank transform %.1f ms (once), %.3f ms per pyramid, %.3e conv outputs/s\n", prep_ms, ms,
               nout / (ms * 1e-3));
        fftconv_bank_destroy(bank);
        // the same pyramid through fftconv_conv_pyramid: ONE call, the tiles of all levels share the per-bin GEMM and the
        // template spectra are computed once per call (no prepared bank needed)
        {
            CK(cudaFree(d_o));
            size_t tot = 0;
            std::vector<size_t> off(10);
            std::vector<int> Hs(10), Ws(10);
            std::vector<const float*> lvp(10, d_lv);
            for (int l = 0; l < 10; ++l) {
                Hs[l] = Ws[l] = side[l];
                off[l] = tot;
                tot += (size_t)fftconv_fft_size16(side[l] + c.kh - 1) * fftconv_fft_size16(side[l] + c.kw - 1) * c.K;
            }
            CK(cudaMalloc(&d_o, tot * 4));
            std::vector<float*> opl((size_t)10 * c.K);
            for (int l = 0; l < 10; ++l) {
                const size_t pl = (size_t)fftconv_fft_size16(side[l] + c.kh - 1) * fftconv_fft_size16(side[l] + c.kw - 1);
                for (int k = 0; k < c.K; ++k) opl[(size_t)l * c.K + k] = d_o + off[l] + pl * k;
            }
            auto run1 = [&]() {
                return fftconv_conv_pyramid(10, lvp.data(), nullptr, Hs.data(), Ws.data(), c.F, c.kh, c.kw, c.K, kp.data(), khs.data(),
                                            kws.data(), nullptr, ond.data(), opl.data(), nullptr, device, st);
            };
            FC(run1());
            CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(e0, st));
            for (int i = 0; i < iters; ++i) FC(run1());
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ms /= iters;
            printf("pyramid in one call (fftconv_conv_pyramid): %.3f ms per pyramid, %.3e conv outputs/s\n", ms, nout / (ms * 1e-3));
        }
        fftconv_release();
        return 0;
    }
    float *d_data, *d_bank, *d_out;
    fftconv_float2* d_spec;
    CK(cudaMalloc(&d_data, nd * 4)); CK(cudaMalloc(&d_bank, nk1 * c.K * 4));
    CK(cudaMalloc(&d_out, plane * c.K * c.N * 4)); CK(cudaMalloc(&d_spec, sizeof(fftconv_float2) * (size_t)CH * FW * c.F));
    CK(cudaMemcpy(d_data, h_data.data(), nd * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bank, h_bank.data(), nk1 * c.K * 4, cudaMemcpyHostToDevice));

    float* h_out = nullptr;
    std::vector<const float*> kp(c.K);
    std::vector<float*> op(c.K);
    std::vector<int> khs(c.K, c.kh), kws(c.K, c.kw);
    if (host) {
        CK(cudaMallocHost(&h_out, plane * c.K * 4));
        for (int k = 0; k < c.K; ++k) { kp[k] = h_bank.data() + nk1 * k; op[k] = h_out + plane * k; }
    }
    std::vector<const float*> bkp(c.K);
    std::vector<float*> bop((size_t)c.K * c.N);
    std::vector<unsigned char> ond(c.K, 1);
    for (int k = 0; k < c.K; ++k) bkp[k] = d_bank + nk1 * k;
    for (size_t i = 0; i < bop.size(); ++i) bop[i] = d_out + plane * i;
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    auto step = [&]() -> int {
        if (c.N > 1)       // batched images: the images add tiles to the per-bin tensor-core GEMM
            return fftconv_conv_batch(d_data, 1, c.N, c.H, c.W, c.F, c.kh, c.kw, c.K, bkp.data(), khs.data(), kws.data(),
                                      nullptr, ond.data(), bop.data(), 1, nullptr, device, st);
        if (host)
            return fftconv_convolution_fft(h_data.data(), 0, c.H, c.W, c.F, c.kh, c.kw, c.K, kp.data(), khs.data(),
                                           kws.data(), nullptr, nullptr, op.data(), 0, nullptr, 0, nullptr, device, st);
        int r = fftconv_fft_data(d_data, 1, c.H, c.W, c.F, c.kh, c.kw, d_spec, device, st);
        if (r) return r;
        return fftconv_conv_bank(d_spec, CH, FW, c.F, c.K, d_bank, c.kh, c.kw, d_out, nullptr, device, st);
    };
    for (int i = 0; i < 3; ++i) FC(step());
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto t0 = std::chrono::steady_clock::now();
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) FC(step());
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
    float ev_ms = 0.f;
    CK(cudaEventElapsedTime(&ev_ms, e0, e1));
    const double ms = host ? wall_ms : ev_ms / iters;
    printf("%s path: %.3f ms per call, %.3e conv outputs/s, %lld kernel launches so far\n", host ? "host->host" : "device-resident",
           ms, (double)c.N * c.K * plane / (ms * 1e-3), fftconv_launch_count());
    // compulsory bytes: every input read once, every output written once
    const double a_bytes = 4.0 * nd + 4.0 * nk1 * c.K + 4.0 * c.N * c.K * plane;
    printf("algorithmic bytes %.1f MB -> %.1f GB/s = %.1f %% of the 6.55 TB/s measured HBM copy bandwidth\n", a_bytes / 1e6,
           a_bytes / (ms * 1e-3) / 1e9, 100.0 * a_bytes / (ms * 1e-3) / 6.55e12);

    if (check > 0) {
        std::vector<float> out(plane * c.K);
        if (host) memcpy(out.data(), h_out, plane * c.K * 4);
        else CK(cudaMemcpy(out.data(), d_out, plane * c.K * 4, cudaMemcpyDeviceToHost));
        std::mt19937 pick(7);
        double num = 0, den = 0;
        const int OH = c.H + c.kh - 1, OW = c.W + c.kw - 1;
        for (int k = 0; k < c.K; k += std::max(1, c.K / 8))
            for (int s = 0; s < check; ++s) {
                const int oy = pick() % OH, ox = pick() % OW;
                double acc = 0;
                for (int f = 0; f < c.F; ++f)
                    for (int kx = 0; kx < c.kw; ++kx) {
                        const int dx = ox - kx;
                        if (dx < 0 || dx >= c.W) continue;
                        for (int ky = 0; ky < c.kh; ++ky) {
                            const int dy = oy - ky;
                            if (dy < 0 || dy >= c.H) continue;
                            acc += (double)h_data[((size_t)f * c.W + dx) * c.H + dy] *
                                   (double)h_bank[nk1 * k + ((size_t)f * c.kw + kx) * c.kh + ky];
                        }
                    }
                const double got = out[plane * k + (size_t)ox * FH + oy];
                num += (got - acc) * (got - acc); den += acc * acc;
            }
        const double rel = std::sqrt(num / std::max(den, 1e-300));
        printf("check: relative L2 on sampled pixels vs float64 direct convolution = %.3e (%s)\n", rel, rel < 1e-5 ? "ok" : "FAIL");
        if (!(rel < 1e-5)) return 4;
    }
    fftconv_release();
    return 0;
}
