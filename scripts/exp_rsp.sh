timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_baseline_shapes.py -m gpu -x -q 2>&1 | tail -3
FFTCONV_OS_INV_Z=0 timeout 300 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "not pyramid" 2>&1 | tail -2
FFTCONV_OS_INV_TMA=0 timeout 300 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "not pyramid" 2>&1 | tail -2
timeout 300 python scripts/c5_time.py 5000 2>&1 | grep -A8 "one_call=True" | grep "os_gemm\|os_inverse\|one_call"
timeout 300 python scripts/config_time.py c4 0.25 2>&1 | tail -5
python bench.py --no-cpu --no-configs --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],3) for k,v in d['roofline']['kernel_share_of_step'].items()})"
