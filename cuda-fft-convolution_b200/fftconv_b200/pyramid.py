"""Feature-pyramid x template-bank schedule (BASELINE config 5, the exemplar-SVM / DPM detection loop).

What the reference's callers do by hand — call cudaFFTData once per pyramid level and cudaConvFFTData
with the whole cell of templates on every level (demoCudaConvolutionFFT.m:111-129 is one level of it) —
as one schedule over a process group: the template bank is sharded across the ranks (fftconv_b200.sharding),
rank 0 transforms every level, all level spectra are broadcast up front (NCCL over NVLink; the broadcasts
queue on NCCL's stream ahead of the compute), and every rank convolves its own shard of the bank with
every level.  Outputs stay sharded: plane (level l, template k) lives on the rank that owns k.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

from .sharding import shard_bank


def pyramid_sides(base: int = 256, levels: int = 10, per_octave: int = 5) -> List[int]:
    """Side of level l of a HOG pyramid with `per_octave` levels per octave: round(base * 2^(-l/per_octave))
    (256, 223, 194, 169, 147, 128, 111, 97, 84, 74 for the defaults)."""
    return [int(round(base * 2.0 ** (-l / per_octave))) for l in range(levels)]


def level_plane(side_h: int, side_w: int, kh: int, kw: int) -> Tuple[int, int]:
    """(FH, FW) of a level: computeFFTsize16 of side + kernel - 1 (src/cudaConvFFTData.h:96-102)."""
    f16 = lambda n: (n // 16) * 16 + (16 if n % 16 else 0)
    return f16(side_h + kh - 1), f16(side_w + kw - 1)


def pyramid_convolution(levels: Optional[Sequence], max_kh: int, max_kw: int, n_templates: int,
                        costs: Optional[Sequence[float]], fft_fn: Callable, alloc_spec: Callable,
                        conv_fn: Callable, level_shapes: Sequence[Tuple[int, int, int]], group=None):
    """Run the schedule on the calling rank.

    levels        rank 0: the level feature maps (any objects fft_fn understands); other ranks: None
    level_shapes  [(H, W, F)] of every level, known on every rank
    costs         per-template cost (kh*kw) for the shard balance, None = uniform
    fft_fn(level, max_kh, max_kw)  -> spectrum tensor [F][FW][CH]              (rank 0)
    alloc_spec(H, W, F)            -> empty spectrum tensor of that level        (other ranks)
    conv_fn(l, spec, begin, end)   -> whatever the caller keeps for templates begin..end on level l
    Returns (begin, end, [conv_fn result per level])."""
    import torch
    import torch.distributed as dist
    on = dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    b, e = shard_bank(list(costs) if costs is not None else None, world, n_templates)[rank]
    specs, pending = [], []
    for l, (H, W, F) in enumerate(level_shapes):
        spec = fft_fn(levels[l], max_kh, max_kw) if rank == 0 else alloc_spec(H, W, F)
        if world > 1:
            pending.append(dist.broadcast(torch.view_as_real(spec), src=0, group=group, async_op=True))
        specs.append(spec)
    results = []
    for l, spec in enumerate(specs):
        if world > 1:
            pending[l].wait()             # stream-ordered on CUDA: later levels keep arriving during compute
        results.append(conv_fn(l, spec, b, e))
    return b, e, results


_SIDE_STREAMS = {}


def _side_stream(dev):
    """One side stream per device for collectives that run next to the template transforms."""
    import torch
    key = str(dev)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def pyramid_convolution_cuda(level_tensors: Optional[Sequence], level_shapes: Sequence[Tuple[int, int, int]],
                             bank_t, kh: int, kw: int, outs: Optional[List] = None, options=None, group=None,
                             one_call: bool = True, raw: bool = True):
    """CUDA instance of the schedule: level_tensors[l] float32 [F][W][H] on this rank's device (rank 0),
    bank_t float32 [K][F][kw][kh] — the FULL bank, identical on every rank; each rank convolves its shard.
    one_call (default): all levels x the shard through fftconv_conv_pyramid; False: one cudaConvFFTData-style call per
    level (the reference caller's loop).  raw (default, with one_call): the raw levels are broadcast as one packed buffer and
    tiled directly; False: rank 0 runs cudaFFTData per level and the ten spectra are broadcast (the two-call form).
    Returns (begin, end, [out_l float32 [end-begin][FW_l][FH_l]])."""
    import torch
    import fftconv_b200 as fc
    K = int(bank_t.shape[0])
    dev = bank_t.device

    def fft_fn(level, mkh, mkw):
        F, W, H = (int(x) for x in level.shape)
        return fc.fft_data_device(level, H, W, F, mkh, mkw)

    def alloc_spec(H, W, F):
        FH, FW = level_plane(H, W, kh, kw)
        return torch.empty((F, FW, FH // 2 + 1), dtype=torch.complex64, device=dev)

    if one_call and raw:
        # the overlap-save path tiles raw data, so what travels is the raw pyramid: ONE packed buffer (smaller than the ten
        # spectra, one collective instead of ten), and no level is transformed to a full-plane spectrum anywhere
        import torch.distributed as dist
        from .sharding import shard_bank
        on = dist.is_initialized()
        world = dist.get_world_size(group) if on else 1
        rank = dist.get_rank(group) if on else 0
        b, e = shard_bank(None, world, K)[rank]
        sizes = [H * W * F for (H, W, F) in level_shapes]
        ready = None
        if world > 1:
            if rank == 0:
                packed = torch.cat([t.reshape(-1) for t in level_tensors])
            else:
                packed = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
            # the broadcast runs behind a side stream: the call below makes only its data-side work wait for it, so the
            # pyramid travels over NVLink while the first chunk of templates is being transformed
            cur = torch.cuda.current_stream(dev)
            side = _side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                dist.broadcast(packed, src=0, group=group)
                ready = torch.cuda.Event()
                ready.record(side)
            packed.record_stream(side)
            offs = [0]
            for n in sizes:
                offs.append(offs[-1] + n)
            lv = [packed[offs[l]:offs[l + 1]].view(F, W, H) for l, (H, W, F) in enumerate(level_shapes)]
        else:
            lv = list(level_tensors)
        if e == b:
            if ready is not None:
                torch.cuda.current_stream(dev).wait_event(ready)
            return b, e, [outs[i] if outs is not None else None for i in range(len(level_shapes))]
        return b, e, fc.conv_pyramid(lv, bank_t[b:e], kh, kw, outs=outs, options=options, data_ready=ready)

    if one_call:
        # fftconv_conv_pyramid: the spectra of all levels (transformed / received below) and the shard of the bank go
        # through ONE call -- the tiles of every level share one per-bin GEMM, the template spectra are computed once
        got = {}

        def conv_fn(l, spec, b, e):
            got[l] = spec
            if l + 1 < len(level_shapes):
                return None
            if e == b:
                return [outs[i] if outs is not None else None for i in range(len(level_shapes))]
            specs = [got[i] for i in range(len(level_shapes))]
            shapes = [(H, W) for (H, W, _) in level_shapes]
            return fc.conv_pyramid(None, bank_t[b:e], kh, kw, outs=outs, specs=specs, shapes=shapes, options=options)

        b, e, res = pyramid_convolution(level_tensors, kh, kw, K, None, fft_fn, alloc_spec, conv_fn, level_shapes, group)
        return b, e, res[-1]

    def conv_fn(l, spec, b, e):
        out = outs[l] if outs is not None else None
        if e == b:
            return out
        return fc.conv_bank(spec, bank_t[b:e], kh, kw, out, options=options)

    return pyramid_convolution(level_tensors, kh, kw, K, None, fft_fn, alloc_spec, conv_fn, level_shapes, group)


def pyramid_convolution_prepared(level_tensors: Optional[Sequence], level_shapes: Sequence[Tuple[int, int, int]],
                                 bank, outs: Optional[List] = None, options=None, group=None):
    """The same schedule over a PREPARED bank shard (fftconv_b200.Bank built by each rank from its own templates):
    the template spectra are resident, so a level costs data tiles -> per-bin GEMM -> inverse.  What travels is the
    raw level (smaller than its spectrum): rank 0 broadcasts every level up front, every rank convolves all levels
    with its shard.  level_tensors[l]: float32 [F][W][H] on the device (rank 0; allocated here on the others).
    Returns [out_l float32 [bank.K][FW_l][FH_l]]."""
    import torch
    import torch.distributed as dist
    on = dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    dev = torch.device("cuda", bank.device)
    levels, pending = [], []
    for l, (H, W, F) in enumerate(level_shapes):
        t = level_tensors[l] if rank == 0 else torch.empty((F, W, H), dtype=torch.float32, device=dev)
        if world > 1:
            pending.append(dist.broadcast(t, src=0, group=group, async_op=True))
        levels.append(t)
    results = []
    for l, t in enumerate(levels):
        if world > 1:
            pending[l].wait()
        results.append(bank.conv_device(t, outs[l] if outs is not None else None, options=options))
    return results
