mkdir -p gpurun_out
timeout 300 python scripts/c5_multi.py 10000 > gpurun_out/c5_n1_10000.json 2> gpurun_out/c5_n1.err; cat gpurun_out/c5_n1_10000.json; tail -2 gpurun_out/c5_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/c5_multi.py 20000 > gpurun_out/c5_n2_20000.json 2> gpurun_out/c5_n2.err; cat gpurun_out/c5_n2_20000.json; tail -2 gpurun_out/c5_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/c5_multi.py 10000 > gpurun_out/c5_n2_10000.json 2> gpurun_out/c5_n2b.err; cat gpurun_out/c5_n2_10000.json
timeout 120 python bench.py --workload c1 --steps 200 --no-cpu > gpurun_out/c1_bench.json 2> gpurun_out/c1.err; python -c "
import json; d=json.loads(open('gpurun_out/c1_bench.json').read().strip().splitlines()[-1]); print('c1', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
