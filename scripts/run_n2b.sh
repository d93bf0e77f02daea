mkdir -p gpurun_out
for mode in rawbcast; do
FFTCONV_BENCH_BCAST=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_$mode.json').read().strip().splitlines()[-1]); print('n2 $mode', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('extras') or {}).get('c5',{}).get('ms'))" || tail -8 gpurun_out/bench_n2_$mode.err
done
