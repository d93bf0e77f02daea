mkdir -p gpurun_out
R=${1:-r02e}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/${R}_bench_n8.json 2> gpurun_out/${R}_bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench_n8.json').read().strip().splitlines()[-1]); print('n8', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config'].get('parallelism')); print(json.dumps((d.get('extras') or {}).get('c5'))[:700])" || tail -12 gpurun_out/${R}_bench_n8.err
