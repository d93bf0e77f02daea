mkdir -p gpurun_out
FFTCONV_BENCH_RAW_SCHEME=scatter timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 50 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_n8_scatter.json 2> gpurun_out/bench_n8_scatter.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_scatter.json').read().strip().splitlines()[-1]); print('n8 scatter', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -12 gpurun_out/bench_n8_scatter.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r02c_bench_n8.json 2> gpurun_out/r02c_bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r02c_bench_n8.json').read().strip().splitlines()[-1]); print('n8 pull', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config'].get('parallelism')); print(json.dumps((d.get('extras') or {}).get('c5'))[:400])" || tail -12 gpurun_out/r02c_bench_n8.err
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu --no-configs --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single gpu same box', d['ms_per_step'], d['value'])"
