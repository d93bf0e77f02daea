# chunking x run-ahead of the template transforms (OsAhead in fftconv.cu) on the C2 one-shot step
mkdir -p gpurun_out
( for cfg in "0 0" "4 0" "4 1" "4 2" "2 0" "2 1" "2 2" "3 1" "1 0" "1 1"; do set -- $cfg
    FFTCONV_OS_NTBLK=$1 FFTCONV_OS_AHEAD=$2 python scripts/oneshot_time.py 1000 30 2>&1 | tail -1
  done ) | tee gpurun_out/exp_ahead.txt
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -x -q -k "ragged or c2_full_bank_device" 2>&1 | tail -3
