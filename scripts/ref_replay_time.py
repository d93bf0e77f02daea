"""The reference's own GPU path (its kernels + host loop + cuFFT, oracle/ref_replay.cu) timed on this box on a sample of
the BASELINE config-2 bank, next to the product path on the same inputs.  Comparator only (never part of bench.py's
timed arms).  python scripts/ref_replay_time.py [n_templates] > profiles/<round>_ref_replay.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200"), os.path.join(ROOT, "tests", "golden")]
import numpy as np
import make_golden
import fftconv_b200 as fc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(2)
data = (rng.random((256, 256, 31), dtype=np.float32) * 0.2).astype(np.float32)
ks = [(rng.standard_normal((16, 16, 31)) * 0.05).astype(np.float32) for _ in range(n)]
L = make_golden.load_ref()
make_golden.ref_run(L, data, 16, 16, ks[:4])
outs, ms = make_golden.ref_run(L, data, 16, 16, ks)
t0 = time.perf_counter(); make_golden.ref_run(L, data, 16, 16, ks); wall = (time.perf_counter() - t0) * 1e3
import ctypes, torch
# the same call through the C ABI (fftconv_convolution_fft), host buffers pinned, copies inside the timed region
d = np.ascontiguousarray(data.transpose(2, 1, 0)); F, W, H = d.shape
h_data = torch.from_numpy(d).pin_memory()
h_bank = torch.from_numpy(np.stack([np.ascontiguousarray(k.transpose(2, 1, 0)) for k in ks])).pin_memory()
h_out = torch.empty((n, 272, 272), dtype=torch.float32).pin_memory()
kp = (ctypes.c_void_p * n)(*[h_bank.data_ptr() + 4 * k * F * 16 * 16 for k in range(n)])
op = (ctypes.c_void_p * n)(*[h_out.data_ptr() + 4 * k * 272 * 272 for k in range(n)])
khs = (ctypes.c_int * n)(*([16] * n)); kws = (ctypes.c_int * n)(*([16] * n))
Lf = fc.lib(); st = torch.cuda.current_stream().cuda_stream
def ours():
    rc = Lf.fftconv_convolution_fft(h_data.data_ptr(), 0, H, W, F, 16, 16, n, kp, khs, kws, None, None, op, 0, None, 0, None, 0, st)
    assert rc == 0, fc.last_error()
for _ in range(3): ours()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): ours()
torch.cuda.synchronize()
ours_ms = (time.perf_counter() - t0) * 1e2
err = max(float(np.linalg.norm(h_out[k].numpy() - outs[k]) / np.linalg.norm(outs[k])) for k in range(n))
print(json.dumps({"workload": f"256x256x31 map x {n} templates 16x16x31, host buffers in and out",
                  "reference_replay": {"kernel_loop_ms": ms, "call_wall_ms": wall, "outputs_per_s": n * 272 * 272 / (ms * 1e-3),
                                       "ms_per_template": ms / n, "note": "reference kernels + cuFFT 11.4, pageable host buffers as the MEX uses"},
                  "this_repo_same_call": {"call_wall_ms": ours_ms, "outputs_per_s": n * 272 * 272 / (ours_ms * 1e-3),
                                          "ms_per_template": ours_ms / n, "note": "fftconv_convolution_fft, pinned host buffers"},
                  "speedup": ms / ours_ms, "max_rel_l2_between_them": err}))
