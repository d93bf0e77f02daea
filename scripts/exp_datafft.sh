# os_data_fft with rows 0 / 32 packed into one complex row (warps 1 and 3 unpack instead of transforming one row each)
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_spec_cache.py tests/test_gpu_vs_reference_replay.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python scripts/config_time.py c4 2>&1 | tail -8
python scripts/path_time.py c2 3 1000 2>&1 | tail -6
