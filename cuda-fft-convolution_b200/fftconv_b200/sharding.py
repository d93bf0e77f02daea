"""Multi-GPU schedule: the template bank is sharded over the ranks, the data spectrum is replicated.

The reference only sketches this (src/cudaConvFFTDataStreams.cu:219-289: spectrum copied GPU0 -> GPUi
with cudaMemcpyPeerAsync, kernels dealt round-robin to per-GPU plans; dead code, N_GPU = 1 at :271).
Here: one process per GPU, rank 0 transforms the data, the spectrum is broadcast with
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests), every rank convolves its own
contiguous shard; the outputs stay sharded (each (template, plane) depends on nothing else —
src/cudaConvFFTData.cu:191-282 has no cross-iteration state).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_bank(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Split templates 0..K-1 into `world` contiguous ranges of near-equal total cost.

    costs[k] ~ kh*kw of template k (uniform costs give K/world each).  Returns [(begin, end)] per rank;
    ranges are contiguous, ordered, cover 0..K exactly, and may be empty when K < world."""
    K = len(costs)
    if world <= 0:
        raise ValueError("world must be positive")
    total = float(sum(costs))
    bounds = [0]
    acc = 0.0
    k = 0
    for r in range(1, world):
        target = total * r / world
        while k < K and acc + costs[k] / 2.0 <= target:
            acc += costs[k]
            k += 1
        bounds.append(k)
    bounds.append(K)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def broadcast_spectrum(spec, src: int = 0, group=None):
    """Broadcast a complex64 spectrum tensor [F][FW][CH] from `src` (in place on the other ranks)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return spec
    dist.broadcast(torch.view_as_real(spec), src=src, group=group)
    return spec


_bcast_streams = {}


def broadcast_spectrum_async(spec, src: int = 0, group=None):
    """Broadcast on a side stream; returns a torch.cuda.Event to hand to conv_bank(spectrum_ready=...) (None when
    there is nothing to wait for).  The side stream first waits for the caller's current stream (rank `src` has just
    queued the transform that fills `spec`; on every rank the previous consumer of `spec` is ordered before), then the
    NCCL broadcast runs there, so the caller's stream is free to run the image-independent template transforms while
    the spectrum crosses NVLink."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    if not spec.is_cuda:                       # gloo / CPU tests: nothing to overlap
        dist.broadcast(torch.view_as_real(spec), src=src, group=group)
        return None
    dev = spec.device
    side = _bcast_streams.get(dev.index)
    if side is None:
        side = _bcast_streams[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        dist.broadcast(torch.view_as_real(spec), src=src, group=group)
        ev = torch.cuda.Event()
        ev.record(side)
    spec.record_stream(side)
    return ev


def bind_host_to_gpu(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPU cores NVML reports as local to the GPU (same NUMA node / PCIe root), BEFORE
    pinned host buffers are allocated: first touch then places them next to the GPU's PCIe link.  With one process per
    GPU this keeps the ranks' host<->device streams off the inter-socket link.  Best effort: returns the core list, or
    None when NVML / the affinity call is unavailable."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def sharded_convolution(data, max_kh: int, max_kw: int, kernels: Sequence, fft_fn: Callable, conv_fn: Callable,
                        alloc_spec: Callable, gather: bool = False, group=None):
    """cudaConvolutionFFT over a process group.

    fft_fn(data, max_kh, max_kw) -> spectrum tensor (called on rank 0 only)
    alloc_spec() -> empty spectrum tensor of the right shape/device (other ranks)
    conv_fn(spec, kernels_shard) -> list of planes for the shard
    Returns (begin, end, planes) for this rank, or the full list on every rank when gather=True."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    costs = [float(k.shape[0] * k.shape[1]) for k in kernels]
    b, e = shard_bank(costs, world)[rank]
    spec = fft_fn(data, max_kh, max_kw) if rank == 0 else alloc_spec()
    spec = broadcast_spectrum(spec, 0, group)
    planes = conv_fn(spec, list(kernels[b:e]))
    if not gather or world == 1:
        return (b, e, planes) if not gather else planes
    parts: List[Optional[list]] = [None] * world
    dist.all_gather_object(parts, planes, group=group)
    return [p for part in parts for p in part]
