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


def gather_packed(packed):
    """packed: this rank's result buffer [L] float32 (rois followed by the int32 counts' bit patterns, see
    Engine.detect_packed) -> [world, L] in rank order.  ONE collective per batch.  Issued on the caller's current
    stream (Engine.rois_batches calls it on its result stream, off the compute stream).  The collective is enqueued
    asynchronously and the CURRENT STREAM is made to wait for it (no host block: the synchronous form was seen to hold
    the host for up to 10 ms per call on 2 GPUs, tools/dbg_gather.py)."""
    world = dist.get_world_size()
    packed = packed.contiguous()
    out = torch.empty((world * packed.numel(),), dtype=packed.dtype, device=packed.device)   # flat: gloo wants [world * L]
    work = dist.all_gather_into_tensor(out, packed.reshape(-1), async_op=True)
    work.wait()        # NCCL: stream-level wait on the current stream; gloo (CPU tests): blocks until done
    return out.view((world,) + tuple(packed.shape))


def gather_results(rois, count):
    """rois [B,post,5] f32, count [B] i32 of this rank -> ([world*B,post,5], [world*B]) in rank order.
    Every rank must contribute the same B (pad the last shard with empty images if needed).  The counts ride in the
    same buffer as the rois (bit patterns), so this is a single all-gather."""
    B = rois.shape[0]
    n = rois.numel()
    packed = torch.cat([rois.reshape(-1), count.contiguous().view(torch.float32).reshape(-1)])
    out = gather_packed(packed)
    all_r = out[:, :n].reshape((out.shape[0] * B,) + tuple(rois.shape[1:]))
    all_c = out[:, n:].contiguous().view(torch.int32).reshape(-1)
    return all_r, all_c
