"""BASELINE config 5 under torchrun: 10-level 31-channel HOG pyramid x K templates 16x16x31 sharded over the ranks,
level spectra broadcast by NCCL.  Reports whole-job outputs/s (max over ranks, CUDA events) and a parity spot check.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/c5_multi.py [K total]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import numpy as np, torch
import torch.distributed as dist
import fftconv_b200 as fc
from fftconv_b200.pyramid import pyramid_convolution_cuda, pyramid_sides, level_plane

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
F, kh, kw = 31, 16, 16
sides = pyramid_sides()
g = torch.Generator(device="cuda").manual_seed(5)
levels = [torch.rand((F, s, s), device="cuda", generator=g) * 0.2 for s in sides]          # same seed: identical everywhere
bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05                       # the full bank on every rank
shapes = [(s, s, F) for s in sides]
from fftconv_b200.sharding import shard_bank
b, e = shard_bank(None, world, K)[rank]
outs = [torch.empty((e - b,) + level_plane(s, s, kh, kw)[::-1], device="cuda") for s in sides]

def step():
    return pyramid_convolution_cuda(levels if rank == 0 else None, shapes, bank, kh, kw, outs)

for _ in range(2):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t.item()))
ms = float(np.median(ts))
# parity: level 3, first template of this rank's shard, float64 FFT convolution
FH, FW = level_plane(sides[3], sides[3], kh, kw)
ref = torch.fft.irfft2(torch.fft.rfft2(levels[3].double(), s=(FW, FH)) * torch.fft.rfft2(bank[b].double(), s=(FW, FH)), s=(FW, FH)).sum(0)
err = torch.tensor([float((outs[3][0].double() - ref).norm() / ref.norm())], device="cuda")
if world > 1: dist.all_reduce(err, op=dist.ReduceOp.MAX)
nout = sum(K * o.shape[1] * o.shape[2] for o in outs)
if rank == 0:
    print(json.dumps({"config": "C5: 10-level pyramid (256..74) x %d templates 16x16x31, bank sharded over %d GPU(s), NCCL spectrum broadcast" % (K, world),
                      "n_gpus": world, "ms": ms, "outputs_per_s": nout / (ms * 1e-3), "templates_per_gpu": e - b,
                      "max_rel_l2_vs_fp64": float(err.item())}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
