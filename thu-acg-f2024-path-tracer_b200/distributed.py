"""Multi-GPU plumbing: one process per GPU, samples-per-pixel split, ONE reduce(sum) of the fp32 accumulators.

The path shards by independent units: rank g renders sample indices {g, g+G, g+2G, ...} of every
pixel; the counter-based RNG is keyed by (seed, pixel, sample), so the union over ranks is exactly the 1-GPU
sample set.  There is no data-path collective; the only exchange is the final reduce of W*H*3 fp32 sums
(24.9 MB at FHD) over NCCL/NVLink, done with torch.distributed (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def partition_samples(spp, rank, world):
    """(sample_begin, sample_count, sample_stride) for this rank: indices rank, rank+world, ... below spp."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    count = (spp - rank + world - 1) // world if spp > rank else 0
    return rank, count, world


def reduce_accumulators(accum, dst=0, group=None):
    """Sum the per-rank radiance accumulators onto rank `dst` (a single collective)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def render_distributed(dev_scene, camera, spp, seed, nan_policy=0, pool_paths=0, group=None):
    """Each rank adds its share of samples into a torch CUDA tensor through pt_render_accumulate, then one reduce.
    Returns (accum tensor [H,W,3] fp32 holding radiance SUMS — complete on rank 0, stats of this rank)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    begin, count, stride = partition_samples(spp, rank, world)
    h = dev_scene.ctx.lib.pt_camera_image_height(camera)
    accum = torch.zeros((h, camera.image_width, 3), dtype=torch.float32, device="cuda")
    dev_scene.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    stats = None
    if count:
        stats = dev_scene.render_accumulate(accum.data_ptr(), camera=camera, spp=count, seed=seed, sample_begin=begin,
                                            sample_stride=stride, nan_policy=nan_policy, pool_paths=pool_paths)
    reduce_accumulators(accum, 0, group)
    return accum, stats
