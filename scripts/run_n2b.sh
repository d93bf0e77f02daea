mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_spec_cache.py -m gpu -x -q 2>&1 | tail -3
for mode in allgather; do
FFTCONV_BENCH_BCAST=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_$mode.json').read().strip().splitlines()[-1]); print('n2 $mode', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['configs'].get('c5'))" || tail -8 gpurun_out/bench_n2_$mode.err
done
