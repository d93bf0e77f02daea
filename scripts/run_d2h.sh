mkdir -p gpurun_out
: > gpurun_out/r02c_d2h_ceiling.jsonl
CUDA_VISIBLE_DEVICES=0 python scripts/d2h_ceiling.py >> gpurun_out/r02c_d2h_ceiling.jsonl 2>/dev/null
for n in 2 4 8; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n scripts/d2h_ceiling.py 2>/dev/null | grep n_ranks >> gpurun_out/r02c_d2h_ceiling.jsonl
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 scripts/d2h_ceiling.py --bind 2>/dev/null | grep n_ranks >> gpurun_out/r02c_d2h_ceiling.jsonl
cat gpurun_out/r02c_d2h_ceiling.jsonl
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" | head -8
