"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink; gloo on CPU for the host-logic tests).

The path shards by independent units: calibration samples per rank with ONE all-reduce of dL/dWq per
iteration, evaluation images round-robin per rank with a 3-number all-reduce.  No other collective exists."""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if (dist.is_available() and dist.is_initialized()) else (0, 1)


def shard_indices(n_items: int, rank: int = None, world_size: int = None):
    """Round-robin assignment of evaluation images (rank r takes r, r+G, ...)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n_items, world_size))


def allreduce_flat_(tensors, group=None, average: bool = False):
    """Bucket a list of gradient tensors into one flat buffer, all-reduce (sum) once, scatter back in place."""
    _, w = world()
    if w == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= w
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return tensors


def reduce_metrics(psnr_sum: float, bpp_sum: float, count: int, device="cpu"):
    """(sum psnr, sum bpp, count) -> global averages; the only collective evaluation needs."""
    acc = torch.tensor([psnr_sum, bpp_sum, float(count)], dtype=torch.float64, device=device)
    _, w = world()
    if w > 1:
        dist.all_reduce(acc)
    p, b, c = acc.tolist()
    return p / max(c, 1), b / max(c, 1), int(c)
