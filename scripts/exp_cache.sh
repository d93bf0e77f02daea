timeout 900 python -m pytest tests/test_gpu_spec_cache.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0 1 2; do echo "SPEC_CACHE=$c"; FFTCONV_SPEC_CACHE=$c python bench.py --no-cpu --no-configs --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], {k:round(v,3) for k,v in d['roofline']['kernel_share_of_step'].items()})"; done
