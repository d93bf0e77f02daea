mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "pyramid" 2>&1 | tail -3
bash scripts/run_c5_n8.sh 2 | python -c "
import json,sys
t=sys.stdin.read(); d=json.loads(t[t.index('{'):])
print({k:v for k,v in d.items() if k!='ms'})
for k,v in d['ms'].items(): print(k, {a:b for a,b in v.items()})"
