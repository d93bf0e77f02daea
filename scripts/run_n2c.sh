mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "pyramid" 2>&1 | tail -2
CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/c5_time.py 20000 2>&1 | grep "one_call"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02c_bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('extras') or {}).get('c5',{}).get('ms'), (d.get('extras') or {}).get('c5',{}).get('rel_l2_vs_fp64'))" || tail -8 gpurun_out/r02c_bench_n2.err
