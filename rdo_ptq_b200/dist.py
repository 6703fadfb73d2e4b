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


class PeerLayer:
    """Symmetric (peer-mapped) buffers of one layer for the fused multi-GPU tail (b200lic_xgpu_reduce_adam_sched):
    every rank allocates [gradient | alpha | flags] through torch's symmetric-memory allocator, the rendezvous exchanges
    the CUDA IPC handles, and the kernel receives device arrays of the ranks' pointers.  Plumbing only: allocation,
    handle exchange and pointer tables; the reduction, the Adam step and the barriers are in the kernel."""

    def __init__(self, alpha: torch.nn.Parameter, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        n = alpha.numel()
        if n % 4 != 0:
            raise ValueError("layer size must be a multiple of four elements")
        dev = alpha.device
        flag_words = 64 * ((2 * self.world + 63) // 64)
        self.buf = symm.empty(2 * n + flag_words, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.grad = self.buf[:n]
        self.alpha = self.buf[n:2 * n]
        self.alpha.copy_(alpha.data.reshape(-1))
        alpha.data = self.alpha.view(alpha.shape)             # the Parameter now lives in peer-mapped memory
        self.grad_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.alpha_ptrs = torch.tensor([p + 4 * n for p in ptrs], dtype=torch.int64, device=dev)
        self.flag_ptrs = torch.tensor([p + 8 * n for p in ptrs], dtype=torch.int64, device=dev)
        self.state = torch.zeros(2, dtype=torch.int32, device=dev)
        per = (n // 4 + self.world - 1) // self.world * 4
        self.lo, self.hi = min(n, self.rank * per), min(n, (self.rank + 1) * per)
        torch.cuda.synchronize(dev)
        dist.barrier(group)                                    # every rank's buffers are zeroed before anyone signals
