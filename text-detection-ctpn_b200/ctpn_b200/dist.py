"""Multi-GPU plumbing: images are independent, so the batch is cut into contiguous per-rank shards and
the only collective on the path is the gather of the fixed-shape per-image results.  Backend-agnostic
(NCCL on GPUs; the CPU tests run it over gloo with world_size 2)."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(int(n_images), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(rois, count):
    """rois [B,post,5] f32, count [B] i32 of this rank -> ([world*B,post,5], [world*B]) in rank order.
    Every rank must contribute the same B (pad the last shard with empty images if needed)."""
    world = dist.get_world_size()
    rois, count = rois.contiguous(), count.contiguous()
    all_r = torch.empty((world * rois.shape[0],) + tuple(rois.shape[1:]), dtype=rois.dtype, device=rois.device)
    all_c = torch.empty((world * count.shape[0],), dtype=count.dtype, device=count.device)
    dist.all_gather_into_tensor(all_r, rois)
    dist.all_gather_into_tensor(all_c, count)
    return all_r, all_c
