"""Config 5 under torchrun (N ranks): where the time of the sharded pyramid step goes.
   A  NCCL broadcast of the packed raw pyramid alone        B  peer pull of the same buffer (PeerBroadcastRaw) alone
   C  fftconv_conv_pyramid on the rank's shard alone (levels already local)
   D  the schedule as bench.py times it (pyramid_convolution_cuda: cat + NCCL broadcast + one call)
   E  the same with the peer-pull delivery
   CUDA events per rank, barrier in front of every repetition; prints per-rank medians and the max over ranks."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-fft-convolution_b200")]
import numpy as np, torch
import torch.distributed as dist
import fftconv_b200 as fc
from fftconv_b200.pyramid import pyramid_convolution_cuda, pyramid_sides, level_plane
from fftconv_b200.sharding import shard_bank, PeerBroadcastRaw

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
try:
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
except Exception:
    opts = None
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
K, F, kh, kw = 20000, 31, 16, 16
sides = pyramid_sides()
g = torch.Generator(device="cuda").manual_seed(5)
levels = [torch.rand((F, s, s), device="cuda", generator=g) * 0.2 for s in sides]
bank = torch.randn((K, F, kw, kh), device="cuda", generator=g) * 0.05
shapes = [(s, s, F) for s in sides]
b, e = shard_bank(None, world, K)[rank]
outs = [torch.empty((e - b,) + level_plane(s, s, kh, kw)[::-1], device="cuda") for s in sides]
packed_src = torch.cat([t.reshape(-1) for t in levels])
pad = (-packed_src.numel()) % 4                       # the peer buffer moves 16-byte units
if pad:
    packed_src = torch.cat([packed_src, packed_src.new_zeros(pad)])
nbytes = packed_src.numel() * 4
recv = torch.empty_like(packed_src)
bc = PeerBroadcastRaw(nbytes)


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def views(buf):
    offs = np.concatenate([[0], np.cumsum([s * s * F for s in sides])])
    return [buf[int(offs[l]):int(offs[l + 1])].view(F, s, s) for l, s in enumerate(sides)]


def a_nccl():
    buf = packed_src if rank == 0 else recv
    dist.broadcast(buf, src=0)


def b_peer():
    bc.begin()
    if rank == 0:
        bc.publish(packed_src)
    return bc.fetch()


def c_conv():
    fc.conv_pyramid(levels, bank[b:e], kh, kw, outs=outs)


def d_sched():
    pyramid_convolution_cuda(levels if rank == 0 else None, shapes, bank, kh, kw, outs)


def e_peer_sched():
    img = b_peer()
    fc.conv_pyramid(views(img.view(torch.float32)), bank[b:e], kh, kw, outs=outs)


FH, FW = level_plane(sides[3], sides[3], kh, kw)
ref = torch.fft.irfft2(torch.fft.rfft2(levels[3].double(), s=(FW, FH)).unsqueeze(0) *
                       torch.fft.rfft2(bank[b:b + 1].double(), s=(FW, FH)), s=(FW, FH)).sum(1)
ref9 = None


def parity():
    """level 3 / first template and level 9 / last template of the shard against the float64 FFT convolution"""
    global ref9
    if ref9 is None:
        fh, fw = level_plane(sides[9], sides[9], kh, kw)
        ref9 = torch.fft.irfft2(torch.fft.rfft2(levels[9].double(), s=(fw, fh)).unsqueeze(0) *
                                torch.fft.rfft2(bank[e - 1:e].double(), s=(fw, fh)), s=(fw, fh)).sum(1)
    r = max(float((outs[3][:1].double() - ref).norm() / ref.norm()), float((outs[9][-1:].double() - ref9).norm() / ref9.norm()))
    allr = [None] * world
    dist.all_gather_object(allr, r)
    return max(allr)


res = {}
for name, fn in (("A_nccl_bcast", a_nccl), ("B_peer_pull", b_peer), ("C_conv_only", c_conv), ("D_schedule_nccl", d_sched),
                 ("E_schedule_peer", e_peer_sched), ("C2_conv_only_again", c_conv)):
    for o in outs:
        o.fill_(float("nan"))
    ms = timed(fn)
    allms = [None] * world
    dist.all_gather_object(allms, ms)
    res[name] = {"max": max(allms), "per_rank": [round(x, 3) for x in allms]}
    if name[0] in "CDE":
        res[name]["rel_l2_max"] = parity()
        res[name]["nan"] = bool(any(bool(torch.isnan(o).any()) for o in outs))
rels = [res["D_schedule_nccl"]["rel_l2_max"]]
if rank == 0:
    print(json.dumps({"world": world, "peer_enabled": bc.enabled, "packed_MB": nbytes / 1e6, "templates_per_rank": e - b,
                      "ms": res, "rel_l2_level3_max": max(rels)}, indent=1))
bc.close()
dist.destroy_process_group()
