# os_data_fft_occ (FFTCONV_OS_DATA=2): rows overlay the raw windows, channels one after the other -> 96 registers, 35 KB, 5 CTAs / SM
FFTCONV_OS_DATA=2 timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_spec_cache.py tests/test_gpu_vs_reference_replay.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for d in 0 2; do FFTCONV_OS_DATA=$d python scripts/config_time.py c4 2>&1 | grep -A2 "^\[c4\]"; done
for d in 0 2; do FFTCONV_OS_DATA=$d python scripts/oneshot_time.py 1000 30 2>&1 | tail -1; done
