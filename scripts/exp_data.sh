timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_spec_cache.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/config_time.py c4 0.25 2>&1 | tail -5
python bench.py --no-cpu --no-configs --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],3) for k,v in d['roofline']['kernel_share_of_step'].items()})"
