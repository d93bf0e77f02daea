mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu --no-configs --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single gpu same box', d['ms_per_step'], d['value'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02c_bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('extras') or {}).get('c5',{}).get('ms'))" || tail -8 gpurun_out/r02c_bench_n2.err
