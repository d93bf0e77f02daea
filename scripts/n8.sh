mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1]); print('n8', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 gpurun_out/bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/c5_multi.py 20000 > gpurun_out/c5_n8_20000.json 2> gpurun_out/c5_n8.err; cat gpurun_out/c5_n8_20000.json; tail -2 gpurun_out/c5_n8.err
