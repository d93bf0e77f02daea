// TEST INFRASTRUCTURE: compiles the four MEX shims of cuda-fft-convolution_b200/mex/ against the stand-in MATLAB headers
// of this directory and drives them with fake mxArrays (tests/test_mex_shims.py runs the scenarios).
//   mex_driver errors    argument-error table: reference error ids / messages (src/cudaFFTData.cu:28-29,49-54,
//                        src/cudaConvFFTData.cu:47,69,72,107,198,230, src/cudaConvolutionFFT.cu:45-54); no GPU needed
//   mex_driver marshal   marshal_cell / alloc_out_cell on a mixed host / gpuArray cell; no GPU needed
//   mex_driver demo      the demo workload (demoCudaConvolutionFFT.m:37-61 shapes) through all four shims on the GPU,
//                        checked against a float64 direct convolution computed here
#include <cmath>
#include <cstdio>
#include <functional>
#include <random>
#include <string>
#include <vector>

#include "mex.h"
#include "gpu/mxGPUArray.h"
int g_mx_live_gpu_handles = 0;
bool g_mx_fake_gpu = true;

#define mexFunction mexFunction_cudaFFTData
#include "../../cuda-fft-convolution_b200/mex/cudaFFTData.cpp"
#undef mexFunction
#define mexFunction mexFunction_cudaConvFFTData
#include "../../cuda-fft-convolution_b200/mex/cudaConvFFTData.cpp"
#undef mexFunction
#define mexFunction mexFunction_cudaConvolutionFFT
#include "../../cuda-fft-convolution_b200/mex/cudaConvolutionFFT.cpp"
#undef mexFunction
#define mexFunction mexFunction_cudaConvFFTDataStreams
#include "../../cuda-fft-convolution_b200/mex/cudaConvFFTDataStreams.cpp"
#undef mexFunction

typedef void (*MexFn)(int, mxArray**, int, const mxArray**);

static mxArray* make_numeric(std::vector<mwSize> dims, mxClassID cls, const void* src = nullptr, bool gpu = false) {
    mxArray* a = new mxArray;
    a->cls = cls; a->dims = dims; mx_normalise_dims(a->dims); a->is_gpu = gpu;
    const size_t bytes = a->numel() * a->elsize();
    if (gpu && !g_mx_fake_gpu) {
        if (cudaMalloc(&a->data, bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(2); }
        if (src) cudaMemcpy(a->data, src, bytes, cudaMemcpyHostToDevice);
    } else {
        a->data = malloc(bytes + 16);
        if (src) memcpy(a->data, src, bytes); else memset(a->data, 0, bytes);
    }
    return a;
}
static mxArray* make_scalar(double v) { return make_numeric({1, 1}, mxDOUBLE_CLASS, &v); }
static mxArray* make_cell(std::vector<mxArray*> items) {
    mxArray* c = mxCreateCellMatrix(1, items.size());
    for (size_t i = 0; i < items.size(); ++i) mxSetCell(c, i, items[i]);
    return c;
}
static mxArray* make_doubles(std::vector<double> v) { return make_numeric({1, v.size()}, mxDOUBLE_CLASS, v.data()); }

// runs one shim; prints "name|id|message" (id/message empty on success) and the number of live gpu handles
static bool call(const char* name, MexFn fn, std::vector<const mxArray*> rhs, mxArray** out = nullptr) {
    mxArray* plhs[1] = {nullptr};
    const int before = g_mx_live_gpu_handles;
    bool ok = true;
    std::string id, msg;
    try { fn(1, plhs, (int)rhs.size(), rhs.data()); }
    catch (const MexError& e) { ok = false; id = e.id; msg = e.msg; }
    for (char& ch : msg) if (ch == '\n') ch = ' ';
    printf("%s|%s|%s|leaked_handles=%d\n", name, id.c_str(), msg.c_str(), g_mx_live_gpu_handles - before - ((ok && plhs[0] && plhs[0]->is_gpu) ? 1 : 0));
    if (out) *out = plhs[0];
    return ok;
}

static int scenario_errors() {
    g_mx_fake_gpu = true;
    mxArray* data = make_numeric({64, 8, 5}, mxSINGLE_CLASS);
    mxArray* data2d = make_numeric({64, 8, 1}, mxSINGLE_CLASS);              // MATLAB drops the trailing 1
    mxArray* data_dbl = make_numeric({64, 8, 5}, mxDOUBLE_CLASS);
    mxArray* data_gpu = make_numeric({64, 8, 5}, mxSINGLE_CLASS, nullptr, true);
    mxArray* spec = make_numeric({41, 16, 5}, mxSINGLE_CLASS, nullptr, true); spec->cplx = mxCOMPLEX;
    mxArray* ten = make_scalar(10), *four = make_scalar(4);
    mxArray* k_ok = make_numeric({10, 4, 5}, mxSINGLE_CLASS);
    mxArray* k_dbl = make_numeric({10, 4, 5}, mxDOUBLE_CLASS);
    mxArray* k_2d = make_numeric({10, 4, 1}, mxSINGLE_CLASS);
    mxArray* k_f3 = make_numeric({10, 4, 3}, mxSINGLE_CLASS);
    mxArray* k_big = make_numeric({100, 4, 5}, mxSINGLE_CLASS);
    mxArray* k_gpu = make_numeric({10, 4, 5}, mxSINGLE_CLASS, nullptr, true);
    call("fftdata_wrong_nargs", mexFunction_cudaFFTData, {data, ten});
    call("fftdata_2d_data", mexFunction_cudaFFTData, {data2d, ten, four});
    call("fftdata_double_data", mexFunction_cudaFFTData, {data_dbl, ten, four});
    call("fftdata_gpu_data", mexFunction_cudaFFTData, {data_gpu, ten, four});
    call("conv_host_spectrum", mexFunction_cudaConvFFTData, {data, make_cell({k_ok})});
    call("conv_wrong_nargs", mexFunction_cudaConvFFTData, {spec});
    call("conv_not_a_cell", mexFunction_cudaConvFFTData, {spec, k_ok});
    call("conv_double_kernel", mexFunction_cudaConvFFTData, {spec, make_cell({k_ok, k_dbl})});
    call("conv_2d_kernel", mexFunction_cudaConvFFTData, {spec, make_cell({k_2d})});
    call("conv_thread_vector_3", mexFunction_cudaConvFFTData, {spec, make_cell({k_ok}), make_doubles({8, 8, 8})});
    call("conv_feature_mismatch", mexFunction_cudaConvFFTData, {spec, make_cell({k_ok, k_f3})});
    call("conv_kernel_larger_than_plane", mexFunction_cudaConvFFTData, {spec, make_cell({k_big})});
    call("oneshot_wrong_nargs", mexFunction_cudaConvolutionFFT, {data, ten, four});
    call("oneshot_gpu_data", mexFunction_cudaConvolutionFFT, {data_gpu, ten, four, make_cell({k_ok})});
    call("oneshot_not_a_cell", mexFunction_cudaConvolutionFFT, {data, ten, four, k_ok});
    call("oneshot_thread_vector_5", mexFunction_cudaConvolutionFFT, {data, ten, four, make_cell({k_ok}), make_doubles({8, 8, 8, 16, 1})});
    call("streams_host_spectrum", mexFunction_cudaConvFFTDataStreams, {data, make_cell({k_ok})});
    call("streams_gpu_kernel", mexFunction_cudaConvFFTDataStreams, {spec, make_cell({k_gpu})});
    return 0;
}

static int scenario_marshal() {
    g_mx_fake_gpu = true;
    mxArray* a = make_numeric({10, 4, 5}, mxSINGLE_CLASS);
    mxArray* b = make_numeric({7, 3, 5}, mxSINGLE_CLASS, nullptr, true);
    mxArray* c3 = make_numeric({1, 1, 5}, mxSINGLE_CLASS);
    mxArray* cell = make_cell({a, b, c3});
    KernelCell kc;
    marshal_cell(cell, true, kc);
    bool ok = kc.ptr.size() == 3 && kc.ptr[0] == a->data && kc.ptr[1] == b->data && kc.ptr[2] == c3->data &&
              kc.kh[0] == 10 && kc.kw[0] == 4 && kc.kf[0] == 5 && kc.kh[1] == 7 && kc.kw[1] == 3 && kc.kf[1] == 5 &&
              kc.kh[2] == 1 && kc.kw[2] == 1 && kc.kf[2] == 5 &&
              kc.on_dev[0] == 0 && kc.on_dev[1] == 1 && kc.on_dev[2] == 0 && kc.handles.size() == 1 && g_mx_live_gpu_handles == 1;
    kc.release();
    ok = ok && g_mx_live_gpu_handles == 0;
    std::vector<float*> outs;
    mxArray* oc = alloc_out_cell(3, 80, 16, outs);
    ok = ok && oc->cls == mxCELL_CLASS && oc->cells.size() == 3 && outs.size() == 3;
    for (int k = 0; k < 3 && ok; ++k) {
        const mxArray* p = oc->cells[k];
        ok = p->cls == mxSINGLE_CLASS && p->dims.size() == 2 && p->dims[0] == 80 && p->dims[1] == 16 && outs[k] == p->data && !p->is_gpu;
    }
    int n = 0;
    const mxArray* rhs[3] = {a, cell, make_doubles({8, 8, 8, 16})};
    const double* t = thread_arg(3, rhs, 2, n);
    ok = ok && n == 4 && t[3] == 16 && thread_arg(2, rhs, 2, n) == nullptr && n == 0;
    int CH, FW, F;
    mxArray* s2 = make_numeric({41, 16, 1}, mxSINGLE_CLASS, nullptr, true);     // single-channel spectrum: 2 dims
    spectrum_dims(s2, CH, FW, F);
    ok = ok && CH == 41 && FW == 16 && F == 1;
    printf("marshal|%s\n", ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
}

// float64 direct full convolution summed over channels, embedded top-left in an FH x FW plane (column-major)
static std::vector<double> direct_conv(const std::vector<float>& d, int H, int W, int F, const std::vector<float>& k, int kh, int kw, int FH, int FW) {
    std::vector<double> o((size_t)FH * FW, 0.0);
    for (int f = 0; f < F; ++f)
        for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) {
            const double v = d[((size_t)f * W + x) * H + y];
            for (int j = 0; j < kw; ++j) for (int i = 0; i < kh; ++i)
                o[(size_t)(x + j) * FH + (y + i)] += v * k[((size_t)f * kw + j) * kh + i];
        }
    return o;
}
static double rel_l2(const float* a, const std::vector<double>& b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < b.size(); ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return std::sqrt(num / den);
}

static int scenario_demo() {
    g_mx_fake_gpu = false;
    const int H = 64, W = 8, F = 5, kh = 10, kw = 4, K = 3, FH = 80, FW = 16;
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<float> d((size_t)H * W * F);
    for (auto& v : d) v = U(rng);
    std::vector<std::vector<float>> ks(K, std::vector<float>((size_t)kh * kw * F));
    for (auto& k : ks) for (auto& v : k) v = U(rng);
    ks[2] = ks[0];                                                   // demoCudaConvolutionFFT.m:113 cell{1} == cell{3}
    mxArray* data = make_numeric({(mwSize)H, (mwSize)W, (mwSize)F}, mxSINGLE_CLASS, d.data());
    std::vector<mxArray*> host_k, mixed_k;
    for (int k = 0; k < K; ++k) {
        host_k.push_back(make_numeric({(mwSize)kh, (mwSize)kw, (mwSize)F}, mxSINGLE_CLASS, ks[k].data()));
        mixed_k.push_back(make_numeric({(mwSize)kh, (mwSize)kw, (mwSize)F}, mxSINGLE_CLASS, ks[k].data(), k == 1));
    }
    mxArray* spec = nullptr;
    bool ok = call("demo_cudaFFTData", mexFunction_cudaFFTData, {data, make_scalar(kh), make_scalar(kw)}, &spec);
    ok = ok && spec && spec->is_gpu && spec->cplx == mxCOMPLEX && spec->dims.size() == 3 && spec->dims[0] == 41 && spec->dims[1] == 16 && spec->dims[2] == 5;
    mxArray *o1 = nullptr, *o2 = nullptr, *o3 = nullptr;
    ok = ok && call("demo_cudaConvFFTData", mexFunction_cudaConvFFTData, {spec, make_cell(mixed_k), make_doubles({8, 8, 8, 16})}, &o1);
    ok = ok && call("demo_cudaConvolutionFFT", mexFunction_cudaConvolutionFFT, {data, make_scalar(kh), make_scalar(kw), make_cell(mixed_k), make_doubles({8, 8, 8, 16}), make_scalar(0)}, &o2);
    ok = ok && call("demo_cudaConvFFTDataStreams", mexFunction_cudaConvFFTDataStreams, {spec, make_cell(host_k)}, &o3);
    double worst = 0;
    for (int k = 0; k < K && ok; ++k) {
        const std::vector<double> ref = direct_conv(d, H, W, F, ks[k], kh, kw, FH, FW);
        for (mxArray* o : {o1, o2, o3}) {
            ok = ok && o && o->cells.size() == (size_t)K && o->cells[k]->dims[0] == (mwSize)FH && o->cells[k]->dims[1] == (mwSize)FW;
            if (ok) worst = std::max(worst, rel_l2((const float*)o->cells[k]->data, ref));
        }
    }
    ok = ok && worst < 1e-5;
    if (ok) ok = memcmp(o1->cells[0]->data, o1->cells[2]->data, sizeof(float) * FH * FW) == 0;      // determinism check of the demo
    printf("demo|%s|max_rel_l2=%.3e|live_handles=%d\n", ok ? "ok" : "FAILED", worst, g_mx_live_gpu_handles);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    const std::string s = argc > 1 ? argv[1] : "errors";
    if (s == "errors") return scenario_errors();
    if (s == "marshal") return scenario_marshal();
    if (s == "demo") return scenario_demo();
    fprintf(stderr, "usage: mex_driver errors|marshal|demo\n");
    return 2;
}
