# os_data_fft: one 256-bit store (STG.E.ENL2.256) per 32-byte operand unit instead of two 128-bit stores (FFTCONV_OS_DATA_ST256=1)
FFTCONV_OS_DATA_ST256=1 timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_spec_cache.py tests/test_gpu_vs_reference_replay.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for d in 0 1; do FFTCONV_OS_DATA_ST256=$d python scripts/config_time.py c4 2>&1 | grep -A2 "^\[c4\]"; done
