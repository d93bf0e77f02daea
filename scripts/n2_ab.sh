mkdir -p gpurun_out
for mode in peer sync; do
FFTCONV_BENCH_BCAST=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/n2_$mode.json').read().strip().splitlines()[-1]); print('$mode', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['parallelism'][:70])" || tail -12 gpurun_out/n2_$mode.err
done
