# bench.py on N GPUs of one box, launched as the driver launches it:  bash scripts/run_nN.sh N [set]
mkdir -p gpurun_out
N=${1:-2}; R=${2:-r02e}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/${R}_bench_n${N}.json 2> gpurun_out/${R}_bench_n${N}.err
python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench_n${N}.json').read().strip().splitlines()[-1]); print('n$N', d['value'], d['ms_per_step'], d['e2e']['ms_per_step']); c5=(d.get('extras') or {}).get('c5') or {}; print({k: c5.get(k) for k in ('ms','value','rel_l2_vs_fp64','templates_per_gpu','error')})" || tail -12 gpurun_out/${R}_bench_n${N}.err
