"""Raw ceiling of the host-output path on this box: N ranks copy 296 MB (the C2 planes of one step) device -> pinned host at the
same time, nothing else running.  torchrun --nproc-per-node N scripts/d2h_ceiling.py [--bind]   (or plain python for N = 1)
Prints per-rank and aggregate GB/s; `bench.py`'s e2e at N ranks cannot exceed the aggregate figure."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
bind = "--bind" in sys.argv
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
cpus = None
if bind:
    from fftconv_b200.sharding import bind_host_to_gpu
    cpus = bind_host_to_gpu(local)
nbytes = 1000 * 272 * 272 * 4
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h.zero_()
st = torch.cuda.current_stream()
for _ in range(3):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
iters = 20
t0 = time.perf_counter()
for _ in range(iters):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mine = nbytes * iters / dt / 1e9
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rates = [None] * world
    dist.all_gather_object(rates, round(mine, 1))
else:
    rates = [round(mine, 1)]
if rank == 0:
    agg = world * nbytes * iters / float(t.item()) / 1e9
    print(json.dumps({"n_ranks": world, "bound_to_numa_cores": bool(cpus), "bytes_per_copy": nbytes, "per_rank_GBps": rates,
                      "aggregate_GBps": round(agg, 1), "ms_per_296MB_step_at_aggregate": round(nbytes * world / agg / 1e6, 2)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
