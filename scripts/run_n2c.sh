mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -2
for sch in pull scatter; do
FFTCONV_BENCH_RAW_SCHEME=$sch timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_n2_x.json 2> gpurun_out/bench_n2_x.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_x.json').read().strip().splitlines()[-1]); print('n2 $sch', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -8 gpurun_out/bench_n2_x.err
done
