mkdir -p gpurun_out
for mode in async sync; do
FFTCONV_BENCH_BCAST=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/n2_$mode.json').read().strip().splitlines()[-1]); print('$mode', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 gpurun_out/n2_$mode.err
done
timeout 200 python bench.py --steps 30 --no-cpu > gpurun_out/n1.json 2>gpurun_out/n1.err; python -c "
import json; d=json.loads(open('gpurun_out/n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
