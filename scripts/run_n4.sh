mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/r02c_bench_n4.json 2> gpurun_out/r02c_bench_n4.err
python -c "
import json; d=json.loads(open('gpurun_out/r02c_bench_n4.json').read().strip().splitlines()[-1]); print('n4', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('extras') or {}).get('c5',{}).get('ms'))" || tail -12 gpurun_out/r02c_bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
