timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu --no-configs 2>gpurun_out/oneshot.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], {k:round(v*d['ms_per_step'],3) for k,v in d['roofline']['kernel_share_of_step'].items()}); print(d['extras']['two_call']); print(d['roofline']['frac'], d['config']['api'])" || tail -5 gpurun_out/oneshot.err
