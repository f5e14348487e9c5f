"""Multi-GPU driver: one process per GPU, samples-per-pixel split across ranks with the scene replicated,
fp32 accumulation buffers summed with ONE reduce over NCCL/NVLink (BASELINE.json north_star).

torch is plumbing here (device memory, current stream, torch.distributed); the rendering is libzrt's
zrt_render_device.  The RNG is keyed on the GLOBAL sample index, so the union of the paths traced by all
ranks is the same set of paths for any world size: the u64 counters are exactly world-size invariant and
the image differs only by the order of f32 additions."""
import copy

import numpy as np

from . import _abi as A


def sample_range(samples_per_pixel, rank, world_size):
    """Global sample indices [begin, end) traced by `rank`."""
    return (rank * samples_per_pixel) // world_size, ((rank + 1) * samples_per_pixel) // world_size


def rank_params(params, rank, world_size):
    p = copy.copy(params)
    p.sample_begin, p.sample_end = sample_range(params.samples_per_pixel, rank, world_size)
    p.flags = params.flags | A.ZRT_FLAG_RAW_SUM
    return p


def color_scale(samples_per_pixel):
    """raytrace.zig:157 `1.0 / @intToFloat(f32, samples_per_pixel)` in f32"""
    return np.float32(1.0) / np.float32(samples_per_pixel)


def reduce_and_scale(accum, counters, samples_per_pixel, group=None, dst=0):
    """accum: float tensor of raw sums, counters: int64[6] tensor.  One SUM-reduce each to `dst`, then the
    reference's final `* (1/spp)` (raytrace.zig:182) on the destination rank."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
        dist.reduce(counters, dst=dst, op=dist.ReduceOp.SUM, group=group)
        is_dst = dist.get_rank(group) == dst
    else:
        is_dst = True
    if is_dst:
        accum.mul_(float(color_scale(samples_per_pixel)))
    return is_dst


def create_group(built_scene, device, group=None):
    """One process per GPU (torchrun): every rank of the torch.distributed group gets the same NCCL id - made by
    libzrt on rank 0, broadcast as 128 bytes through torch.distributed - and joins libzrt's OWN communicator
    (zrt_multi_create_rank).  From then on the per-step path is libzrt only: render, ncclReduce, 1/spp, copy home."""
    import torch
    import torch.distributed as dist

    from . import lib as Z

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return Z.MultiScene(built_scene, devices=[device])
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = [Z.comm_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    return Z.MultiScene(built_scene, device=device, comm_id=ids[0], rank=rank, world=world)


def render_distributed(dev_scene, camera, params, accum=None, counters=None, group=None):
    """Render this rank's share on its GPU and reduce to rank 0.

    dev_scene: zraytrace_b200.lib.Scene resident on this rank's device.
    Returns (accum tensor [H][W][3] on device, counters tensor int64[6]); only meaningful on rank 0.
    Everything is enqueued on torch's current CUDA stream; nothing synchronises the host."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = torch.device("cuda", dev_scene.device)
    if accum is None:
        accum = torch.empty((params.height, params.width, 3), dtype=torch.float32, device=dev)
    if counters is None:
        counters = torch.zeros(6, dtype=torch.int64, device=dev)
    p = rank_params(params, rank, world)
    stream = torch.cuda.current_stream(dev).cuda_stream
    if p.sample_begin == p.sample_end:
        # fewer samples than ranks: this rank has nothing to trace.  (0, 0) must not reach the C ABI, where it
        # means "all samples" (zrt.h: sample_begin / sample_end)
        accum.zero_()
        counters.zero_()
    else:
        dev_scene.render_device(camera, p, accum.data_ptr(), counters.data_ptr(), stream)
    reduce_and_scale(accum, counters, params.samples_per_pixel, group)
    return accum, counters
