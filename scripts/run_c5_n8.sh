mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/c5_n_breakdown.py > gpurun_out/c5_n${N}_breakdown.json 2> gpurun_out/c5_n${N}_breakdown.err || tail -20 gpurun_out/c5_n${N}_breakdown.err
cat gpurun_out/c5_n${N}_breakdown.json
