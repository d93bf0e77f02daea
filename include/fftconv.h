/* fftconv.h — C ABI of the B200-native FFT-convolution engine (libfftconv.so).
 *
 * Drop-in boundary for the hot path of chrischoy/CUDA-FFT-Convolution.  Every entry point
 * below replaces one MEX entry point of the reference (file:line cited per function); the
 * MEX shims in cuda-fft-convolution_b200/mex/ and the stand-alone C++ harness only marshal
 * arguments and call these.  Plain pointers and sizes only — no torch / MATLAB types.
 *
 * Memory layouts are the reference's (MATLAB column-major, h contiguous,
 * src/cudaConvFFTData.cuh:26-27):
 *     data      H x W x F   fp32  = C array [F][W][H]
 *     kernel    kh x kw x F fp32  = C array [F][kw][kh]
 *     spectrum  CH x FW x F complex fp32 (interleaved) = C array [F][FW][CH], CH = FH/2+1
 *               (exactly what cufftPlanMany(n={FW,FH}, R2C, batch=F) writes,
 *                src/cudaFFTData.cu:128-146)
 *     output    FH x FW fp32 per kernel = C array [FW][FH]; the full linear convolution
 *               sum_f conv2(D_f, k_f) sits top-left, the rest of the plane is ~0
 *               (src/cudaConvFFTData.cu:111,186-188,275-279).  No flip, no crop.
 *     FH = fftconv_fft_size16(H + KH - 1), FW = fftconv_fft_size16(W + KW - 1).
 *
 * Return value: 0 on success, a negative FFTCONV_ERR_* otherwise; the message is available
 * from fftconv_last_error().  The library never calls exit() (the reference does,
 * src/cudaConvFFTData.h:6-29) and never falls back to a CPU path.
 *
 * Synchronisation: when any output (or the stream argument) lives on the host side, i.e.
 * out_on_device == 0, the call returns after the results are in the host buffers (MEX
 * semantics).  With device outputs the call is stream-ordered on `stream` (a cudaStream_t
 * passed as void*, NULL = legacy default stream) and returns without synchronising.
 */
#ifndef FFTCONV_H_
#define FFTCONV_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFTCONV_OK                 0
#define FFTCONV_ERR_INVALID_INPUT  (-1)  /* "Invalid input to MEX file." src/cudaFFTData.cu:28-29,49-54 */
#define FFTCONV_ERR_NOT_CELL       (-2)  /* "Kernel must be a cell array" src/cudaConvFFTData.cu:106-107 (raised by the shims) */
#define FFTCONV_ERR_THREAD_SIZE    (-3)  /* "CUDA Thread Size must be 4 integers ..." src/cudaConvFFTData.cu:71-72 */
#define FFTCONV_ERR_KERNEL_TYPE    (-4)  /* "Kernels must be of type float and have features larger than 1" :197-198 (shims) */
#define FFTCONV_ERR_KERNEL_SHAPE   (-5)  /* "Kernel and Data must have the same number of features and kernel size should be smaller than data size" :229-230 */
#define FFTCONV_ERR_NOT_GPU_ARRAY  (-6)  /* "The data must be FFT-ed real array in GPU" :68-69 */
#define FFTCONV_ERR_UNSUPPORTED    (-9)  /* plane larger than this build supports */
#define FFTCONV_ERR_CUDA           (-10) /* a CUDA runtime call failed (reference: printf + exit) */

typedef struct fftconv_float2 { float x, y; } fftconv_float2;

/* Extension block (all zero = exact reference behaviour). */
typedef struct fftconv_options {
    int correlate;   /* 1: multiply by conj(kernel spectrum) = built-in flip
                        (complexConjMulAndScale, src/cudaConvFFTData.cuh:42-45,63)            */
    int crop_h;      /* >0: store only the first crop_h rows ...                              */
    int crop_w;      /* ... and crop_w columns of each plane (demoCudaConvolutionFFT.m:149)   */
    int out_ld;      /* leading dimension (floats) of a stored column; 0 = crop_h or FH       */
    int force_generic; /* 1: bypass the fast paths (testing); same as path = 1               */
    int path;        /* 0: automatic; 1: generic line-FFT pipeline; 2: 16-point-tiled SIMT
                        pipeline; 3: overlap-save tiles + per-bin complex GEMM on tcgen05
                        (falls back to 2 / 1 when the shape is outside that path's range);
                        4: large-plane pipeline, in-place line transforms (automatic for planes
                        >= 1024 x 1024 that paths 2 / 3 do not serve, e.g. BASELINE config 3)    */
    int reserved[2];
} fftconv_options;

/* computeFFTsize16 — src/cudaConvFFTData.h:96-102.  Part of the API contract. */
int fftconv_fft_size16(int n);
/* computeFFTsize (power-of-two rule; unused by the reference) — src/cudaConvFFTData.h:67-94 */
int fftconv_fft_size_pow2(int n);

/* cudaFFTData — src/cudaFFTData.cu:18-160.
 * data: H x W x F fp32 (host if data_on_device == 0; the MEX requires host, :49-54).
 * d_spec: caller-allocated DEVICE buffer of CH*FW*F complex (the gpuArray of :90-103).  */
int fftconv_fft_data(const float* data, int data_on_device, int H, int W, int F,
                     int KH, int KW, fftconv_float2* d_spec, int device, void* stream);

/* Same, with the clamp-to-edge / wrap padding of the SDK padData
 * (src/convolutionFFTkernel.cu:46-76): index i < n -> i; n <= i < n+kernelOfs -> n-1; else 0. */
int fftconv_fft_data_clamp(const float* data, int data_on_device, int H, int W, int F,
                           int KH, int KW, int kernel_y, int kernel_x,
                           fftconv_float2* d_spec, int device, void* stream);

/* cudaConvFFTData — src/cudaConvFFTData.cu:24-306 (hot loop :191-282).
 * d_spec: device spectrum [F][FW][CH].  kernels[k]: kh[k] x kw[k] x F fp32, on the device
 * iff kernel_on_device[k] (NULL = all host).  kf: per-kernel feature count for the F check
 * of :229 (NULL = F).  outs[k]: FH*FW floats each, all host or all device.
 * threads: optional 4-vector accepted for call compatibility (:71-81); only its length is
 * validated, the engine picks its own tiling.  opt may be NULL. */
int fftconv_conv_fft_data(const fftconv_float2* d_spec, int CH, int FW, int F,
                          int K, const float* const* kernels, const int* kh, const int* kw,
                          const int* kf, const unsigned char* kernel_on_device,
                          float* const* outs, int out_on_device,
                          const double* threads, int nthreads,
                          const fftconv_options* opt, int device, void* stream);

/* cudaConvFFTDataStreams — src/cudaConvFFTDataStreams.cu:121-522.  Same contract as
 * cudaConvFFTData, host kernels only (:352-374).  The reference's 2-stream round-robin
 * (:292-328,:338-469) is rebuilt as a chunked copy/compute pipeline over the kernel bank (the same one
 * fftconv_conv_fft_data runs); its per-stream ConvPlan objects become the graph plans below (fftconv_plan_*). */
int fftconv_conv_fft_data_streams(const fftconv_float2* d_spec, int CH, int FW, int F,
                                  int K, const float* const* kernels, const int* kh, const int* kw,
                                  const int* kf, float* const* outs,
                                  const double* threads, int nthreads,
                                  const fftconv_options* opt, int device);

/* cudaConvolutionFFT — src/cudaConvolutionFFT.cu:27-311: both calls fused, data on the host
 * (:51-54), optional thread vector (:71-82) and 0-based GPU id (:84-89). */
int fftconv_convolution_fft(const float* data, int data_on_device, int H, int W, int F,
                            int maxKH, int maxKW,
                            int K, const float* const* kernels, const int* kh, const int* kw,
                            const int* kf, const unsigned char* kernel_on_device,
                            float* const* outs, int out_on_device,
                            const double* threads, int nthreads,
                            const fftconv_options* opt, int device, void* stream);

/* Extension: device-resident packed bank, K kernels of identical kh x kw x F stored back to
 * back; K output planes stored back to back in d_out (plane stride = FW*FH floats, or
 * crop_w*out_ld with opt->crop_*).  Stream-ordered, no host synchronisation. */
int fftconv_conv_bank(const fftconv_float2* d_spec, int CH, int FW, int F,
                      int K, const float* d_bank, int kh, int kw,
                      float* d_out, const fftconv_options* opt, int device, void* stream);

/* Extension (multi-GPU spectrum delivery): declares that the COMPLETE spectrum at d_spec (as of this point in stream
 * order) is the fftconv_fft_data transform -- zero pad, same maxKH x maxKW -- of the raw data d_raw [F][W][H] on the same
 * device.  For a spectrum this library did not produce in one piece: channel slices transformed on several GPUs and
 * all-gathered over NVLink (the one-process-per-GPU form of the peer copy in src/cudaConvFFTDataStreams.cu:279-289, see
 * sharding.PeerAllGatherSpectrum).  The overlap-save path then tiles the raw data instead of inverting the spectrum back
 * to a plane.  The spectrum is hashed now and re-hashed by the convolution on the device, so a buffer that changed after
 * the declaration is never served from the raw data; the declaration itself is the caller's contract.  The raw data is
 * copied (the caller may reuse d_raw at once).  Stream-ordered. */
int fftconv_spectrum_bind_raw(const fftconv_float2* d_spec, const float* d_raw, int H, int W, int F,
                              int maxKH, int maxKW, int device, void* stream);

/* Extension (BASELINE config "exemplar-SVM scale": feature pyramid x template bank): L levels of different sizes
 * against ONE bank in one call -- what a caller of the reference does with one cudaFFTData + one cudaConvFFTData per
 * level (demoCudaConvolutionFFT.m:111-129 is one level of it).  Level l is either raw data level_data[l] =
 * [F][W[l]][H[l]] on the device or, where level_data is NULL / level_data[l] is NULL, the spectrum level_spec[l] that
 * fftconv_fft_data produced for it with the same maxKH x maxKW ([F][FW_l][CH_l] on the device).  Plane (l, k) goes to
 * outs[l*K + k], a device buffer of FW_l*FH_l floats (FH_l = computeFFTsize16(H[l] + maxKH - 1), ...).  With templates
 * up to 32 x 32 the overlap-save tiles of ALL levels form one N dimension of the per-frequency-bin complex GEMM: the
 * template spectra are computed and streamed once per call instead of once per level.  Otherwise (or with
 * correlate / crop options) the call is L calls of the single-image entry points.  Stream-ordered.
 * Levels that are still being delivered on another stream (the NCCL broadcast of the packed pyramid in the multi-GPU
 * schedule, fftconv_b200/pyramid.py): hand the delivery's event to fftconv_spectrum_ready_event first -- only the data
 * side of this call waits for it, the template transforms of the first chunk start at once. */
int fftconv_conv_pyramid(int L, const float* const* level_data, const fftconv_float2* const* level_spec,
                         const int* H, const int* W, int F, int maxKH, int maxKW,
                         int K, const float* const* kernels, const int* kh, const int* kw,
                         const int* kf, const unsigned char* kernel_on_device,
                         float* const* outs, const fftconv_options* opt, int device, void* stream);

/* Extension (BASELINE config "batched"): N images of identical H x W x F stored back to back
 * (data: C array [N][F][W][H]) against ONE bank; plane (n, k) goes to outs[n*K + k].  With device outputs
 * and kernels up to 32 x 32 the images only add overlap-save tiles to the N dimension of the per-frequency-bin
 * complex GEMM (tcgen05, one kernel-spectrum pass for the whole group of images); otherwise the call is
 * N calls of fftconv_convolution_fft.  Same argument conventions as fftconv_convolution_fft
 * (src/cudaConvolutionFFT.cu:27-311 is its single-image counterpart). */
int fftconv_conv_batch(const float* data, int data_on_device, int N, int H, int W, int F,
                       int maxKH, int maxKW,
                       int K, const float* const* kernels, const int* kh, const int* kw,
                       const int* kf, const unsigned char* kernel_on_device,
                       float* const* outs, int out_on_device,
                       const fftconv_options* opt, int device, void* stream);

/* Extension (the counterpart of cudaFFTData for the kernel side, src/cudaFFTData.cu:1-160): a PREPARED
 * BANK.  On the overlap-save path the template spectra (64 x 64 tiles, tensor-core operand order) do not depend on
 * the image size, so a bank is transformed once, stays resident in HBM (256 KB per template at F = 31) and serves
 * every image / pyramid level / video frame after that: a call is then data tiles -> per-bin GEMM -> inverse.
 * Templates of at most 32 x 32; kernels host or device as in fftconv_conv_fft_data (src/cudaConvFFTData.cu:194-231).
 * fftconv_bank_conv: data H x W x F (host or device), outs[k] = FH x FW plane of template k with
 * FH = fft_size16(H + maxKH - 1), maxKH/maxKW = the largest template of the bank (fftconv_bank_info). */
typedef struct fftconv_bank fftconv_bank;
int fftconv_bank_create(int K, const float* const* kernels, const int* kh, const int* kw, const int* kf,
                        const unsigned char* kernel_on_device, int F, int device, void* stream,
                        fftconv_bank** out);
int fftconv_bank_info(const fftconv_bank* bank, int* K, int* F, int* maxKH, int* maxKW, long long* bytes);
int fftconv_bank_conv(const fftconv_bank* bank, const float* data, int data_on_device, int H, int W,
                      float* const* outs, int out_on_device, const fftconv_options* opt, void* stream);
void fftconv_bank_destroy(fftconv_bank* bank);

/* Extension: the detection consumers of the reference crop each plane and take its maximum on the host
 * (demoCudaConvolutionFFT.m:149 onward).  Here the reduction is fused into the store of the inverse transform: no
 * plane is written or copied; template k yields the maximum of its full linear convolution, i.e. of the
 * (H + kh_k - 1) x (W + kw_k - 1) top-left block of its plane, and its 0-based position (ties: smallest x, then y). */
typedef struct fftconv_peak { float value; int y; int x; int pad; } fftconv_peak;
int fftconv_bank_conv_max(const fftconv_bank* bank, const float* data, int data_on_device, int H, int W,
                          fftconv_peak* peaks, int peaks_on_device, void* stream);

/* Extension (the rest of the row): detections instead of planes.  A detection consumer of the reference adds
 * a per-template bias to each plane, crops it (demoCudaConvolutionFFT.m:149) and keeps the responses above a threshold or
 * the k best ones -- after the D2H of every plane (src/cudaConvFFTData.cu:277).  Both reductions are fused into the store
 * of the inverse transform: response r(y, x) = [full linear convolution of template k](y, x) + bias[k] (bias: K floats on
 * the HOST, NULL = 0), over the (H + kh_k - 1) x (W + kw_k - 1) block; no plane is written or copied.
 *   fftconv_bank_conv_detect: every response >= threshold.  dets[k * max_per_template + i], i < min(counts[k],
 *     max_per_template): the largest ones in descending order (ties: smallest x, then y); counts[k] = how many responses
 *     passed (may exceed max_per_template; at most max(64 * tiles, 2048) are ranked).  Unused slots: value -inf, y = x = -1.
 *   fftconv_bank_conv_topk: the k <= 64 largest responses of every template, exact: a candidate pass bounds the k-th
 *     largest response from below, a threshold pass over the same product spectra collects everything above the bound.
 * dets / counts on the device iff out_on_device (then the call is stream-ordered, except that the top-k mode reads the
 * counts back to detect a candidate overflow). */
int fftconv_bank_conv_detect(const fftconv_bank* bank, const float* data, int data_on_device, int H, int W,
                             const float* bias, float threshold, int max_per_template,
                             fftconv_peak* dets, int* counts, int out_on_device, void* stream);
int fftconv_bank_conv_topk(const fftconv_bank* bank, const float* data, int data_on_device, int H, int W,
                           const float* bias, int k, fftconv_peak* dets, int out_on_device, void* stream);

/* PLANS — the persistent graph schedule over the kernel bank.  The reference prototype keeps one ConvPlan per stream
 * (plans, scratch, stream; src/cudaConvFFTDataStreams.cu:124,292-328) and re-issues every launch of the per-kernel loop on
 * every call (:338-469).  Here a plan fixes the DEVICE buffers of a repeated call -- image H x W x F (or, with d_data ==
 * NULL, its spectrum), packed bank of K kernels kh x kw x F, K output planes back to back (as fftconv_conv_bank) -- and
 * captures the whole launch sequence (cudaFFTData + cudaConvFFTData on whichever pipeline serves the shape, including its
 * internal copy / data-side streams) into one CUDA graph.  fftconv_plan_execute is then a single graph launch on `stream`;
 * the contents of the buffers may change between executions (video frames, pyramid levels of one size).  d_spec is the
 * spectrum buffer (CH*FW*F complex): input when d_data == NULL, otherwise written by every execution.  The library
 * re-captures transparently when another call has grown its cached scratch. */
typedef struct fftconv_plan fftconv_plan;
int fftconv_plan_create(const float* d_data, int H, int W, int F, int KH, int KW, fftconv_float2* d_spec,
                        int K, const float* d_bank, int kh, int kw, float* d_out, const fftconv_options* opt,
                        int device, void* stream, fftconv_plan** out);
int fftconv_plan_execute(fftconv_plan* plan, void* stream);
int fftconv_plan_info(const fftconv_plan* plan, int* graph_nodes, int* path);
void fftconv_plan_destroy(fftconv_plan* plan);

/* modulateAndNormalize — src/convolutionFFTkernel.cu:84-100: in place a = a*b/dataN. */
int fftconv_modulate_and_normalize(fftconv_float2* d_a, const fftconv_float2* d_b,
                                   long long n, int device, void* stream);

/* Multi-GPU schedule (the N_GPU plans of src/cudaConvFFTDataStreams.cu:273-289, where the spectrum reaches GPU i by
 * cudaMemcpyPeerAsync): when the spectrum is being delivered by a collective on ANOTHER stream (an NCCL broadcast),
 * hand its completion event (cudaEvent_t) to the library.  The next convolution call on `device` then makes only its
 * data-side work wait for the event; the template transforms, which do not depend on the image, start at once on the
 * call stream, so the broadcast travels over NVLink in their shadow.  One-shot: consumed by that call
 * (fftconv_conv_fft_data / fftconv_conv_bank / fftconv_convolution_fft / fftconv_conv_pyramid). */
int fftconv_spectrum_ready_event(int device, void* cuda_event);

/* PEER SPECTRUM — the multi-GPU plans of src/cudaConvFFTDataStreams.cu:279-289 copy the spectrum GPU 0 -> GPU i with
 * cudaMemcpyPeerAsync.  One process per GPU cannot call that; the same copy is done here over CUDA IPC + NVLink, ordered
 * entirely on the device (no host synchronisation, no collective kernel):
 *   owner:  fftconv_peer_alloc (cudaMalloc + IPC handle; zero-filled), other ranks: fftconv_peer_open(handle);
 *   owner:  ... kernels that fill the buffer ...; fftconv_peer_signal(flag, step)          (release, system scope)
 *   peer:   fftconv_peer_wait(flag, step); fftconv_peer_pull(local, mapped, bytes); fftconv_peer_signal(ack_r, step)
 *   owner, before refilling: fftconv_peer_wait_all(acks, n, step)  — every peer has pulled the previous contents.
 * Flags are 64-bit counters living in the owner's allocation; every call is stream-ordered.  A wait gives up after
 * about two seconds of device time and records the failure (next fftconv_peer_status call returns non-zero), so a
 * lost peer can never wedge the GPU. */
int fftconv_peer_alloc(size_t bytes, int device, void** ptr, unsigned char handle[64]);
int fftconv_peer_open(const unsigned char handle[64], int device, void** ptr);
int fftconv_peer_close(void* mapped_ptr, int device);
int fftconv_peer_free(void* ptr, int device);
int fftconv_peer_signal(unsigned long long* flag, unsigned long long value, int device, void* stream);
int fftconv_peer_wait(const unsigned long long* flag, unsigned long long value, int device, void* stream);
int fftconv_peer_wait_all(const unsigned long long* flags, int n, unsigned long long value, int device, void* stream);
int fftconv_peer_pull(void* dst, const void* src_mapped, size_t bytes, int device, void* stream);
/* ALL-GATHER form (no single-source fan-out): every rank owns a peer buffer of the same layout [spectrum | ready flag | n
 * acknowledgement slots at flag_off], transforms ITS slice of the channels into its own buffer (fftconv_fft_data on a channel
 * range) and calls fftconv_peer_allgather: the rank raises its ready flag, then ONE kernel waits for every other rank's flag
 * through the NVLink mappings, pulls that rank's slice [offs[p], offs[p+1]) into the same offset of the local buffer and
 * acknowledges in the owner's memory.  bases[p] = rank p's buffer (own pointer for p == rank); n <= 16.  Before refilling its
 * slice for the next step a rank waits for the acknowledgements: fftconv_peer_wait_all(own acks, n, step). */
int fftconv_peer_allgather(void* const* bases, int n, int rank, const unsigned long long* offs,
                           unsigned long long flag_off, unsigned long long step, int device, void* stream);
int fftconv_peer_status(int device);        /* 0: no wait has timed out on this device (synchronises the device) */

/* Which pipeline would serve cudaConvolutionFFT(data H x W x F, declared maximum maxKH x maxKW, K kernels of that size)
 * under `opt` (NULL = defaults): returns fftconv_options.path numbering (1 generic, 2 16-point-tiled, 3 overlap-save +
 * tcgen05 GEMM, 4 large plane).  For path 4 the radices of the in-place line plans along h and w are written to
 * radices_h / radices_w (up to 8 entries, 0-terminated; either may be NULL).  Pure host logic: needs no device. */
int fftconv_query_path(int H, int W, int F, int maxKH, int maxKW, int K, const fftconv_options* opt,
                       int* radices_h, int* radices_w);

/* Number of kernel launches issued by this library since load (bench accounting). */
long long fftconv_launch_count(void);
/* Per-kernel device timing for the roofline leg of bench.py: while enabled, every kernel launch is
 * bracketed by CUDA events on the stream it is launched on.  fftconv_profile_read sums the elapsed
 * time and the launch count of one kernel kind (0 <= kind < fftconv_profile_kinds()). */
void fftconv_profile_enable(int on);
int fftconv_profile_kinds(void);
const char* fftconv_profile_name(int kind);
int fftconv_profile_read(int kind, double* total_ms, long long* launches, int reset);
/* Bytes of device scratch currently held by the cached workspace on `device`. */
long long fftconv_workspace_bytes(int device);
/* Drop cached plans / scratch (all devices). */
void fftconv_release(void);
/* Last error message of the calling thread ("" if none). */
const char* fftconv_last_error(void);
/* "fftconv-b200 <version> sm_100a" */
const char* fftconv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FFTCONV_H_ */
