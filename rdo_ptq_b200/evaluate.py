"""Evaluation entry of the hot path: drop-in for task-oriented-PTQ/test_datasets.py (:21-33 PSNR/bpp, :45-73
pad/crop, :76-117 Test_kodak), losses/losses.py (:15-35 RateDistortionLoss, :38-84 Metrics) and the PSNR / MS-SSIM /
bpp report of light-uniform-PTQ/quantize.py:58-92, on libb200lic reductions (K11) and the MS-SSIM kernels (msssim.cu).

Images are independent, so `evaluate` shards them round-robin over the ranks of the default process group and
all-reduces (sum psnr, sum bpp, [sum ms-ssim,] count): the only collective evaluation needs.
"""
import math

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def pad(x, p=2 ** 6):
    """Centred zero pad to a multiple of p (test_datasets.py:45-58).  Buffer plumbing, not arithmetic."""
    h, w = x.size(2), x.size(3)
    H, W = (h + p - 1) // p * p, (w + p - 1) // p * p
    l, t = (W - w) // 2, (H - h) // 2
    return F.pad(x, (l, W - w - l, t, H - h - t), mode="constant", value=0)


def crop(x, size):
    """Inverse of `pad` (test_datasets.py:61-73)."""
    H, W = x.size(2), x.size(3)
    h, w = size
    l, t = (W - w) // 2, (H - h) // 2
    return x[:, :, t:t + h, l:l + w].contiguous()


def squared_error_sums(a, b):
    """Device tensor [sum (a-b)^2, sum (clamp(a,0,1)-b)^2]."""
    return ops.sq_err_sum(a, b)


def compute_psnr(a, b, clamp=False):
    """-10 log10(mean((a-b)^2)) (test_datasets.py:21-23); `clamp` folds the reference's `rec.clamp_(0,1)` (:98)."""
    s = squared_error_sums(a, b)
    return -10 * math.log10(float(s[1 if clamp else 0]) / a.numel())


def compute_msssim(a, b):
    """pytorch_msssim.ms_ssim(a, b, data_range=1.).item() (LU quantize.py:55-56, dataset_test.py:60-61)."""
    return float(ops.ms_ssim(a, b, data_range=1.0))


def total_bits(out_net):
    """sum over likelihood tensors of sum(-log2 lik), as a device scalar (one K11 pass per tensor)."""
    tot = None
    for lik in out_net["likelihoods"].values():
        b = ops.bits_sum(lik)
        tot = b if tot is None else ops.add_act(tot, b)
    return tot


def compute_bpp(out_net):
    """test_datasets.py:29-33: bits / pixels of the (padded) x_hat."""
    n, _, h, w = out_net["x_hat"].shape
    return float(total_bits(out_net)) / (n * h * w)


class GraphedForward:
    """`model.forward` + the bit count, captured once per input shape as a CUDA graph and replayed.

    The quantised forward of one image is ~250 short kernels (weight re-quantisation on every forward like the
    reference, operand staging, conv, dynamic A8 statistics + apply per layer ...): issued eagerly it is bound by launch
    latency, replayed as a graph it is bound by the kernels.  The first call on a shape runs eagerly (scale
    initialisation and other first-forward state), the second captures.  Returned tensors are the graph's
    static outputs: consume them before the next call."""

    def __init__(self, model):
        self.model, self.cache, self.seen = model, {}, set()

    @torch.no_grad()
    def __call__(self, x):
        key = (tuple(x.shape), x.device.index)
        ent = self.cache.get(key)
        if ent is None:
            if key not in self.seen:
                self.seen.add(key)
                with ops.defer_actq():
                    out = self.model.forward(x)
                return out, total_bits(out)
            static_x = x.clone()
            with ops.defer_actq():
                self.model.forward(static_x)     # eager: builds every lazily prepared operand before the capture
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), ops.defer_actq():
                out = self.model.forward(static_x)
                bits = total_bits(out)
            ent = self.cache[key] = (g, static_x, out, bits)
        g, static_x, out, bits = ent
        static_x.copy_(x)
        g.replay()
        return out, bits


@torch.no_grad()
def evaluate(model, images, p=256, shard=True, graph=True, ms_ssim=False):
    """Test_kodak (test_datasets.py:76-117) over a list of [1,3,h,w] CUDA tensors.
    Returns dict(psnr, bpp, count, per_image=[(psnr, bpp), ...] for this rank's images); with `ms_ssim` (the third
    number of LU quantize.py:58-92 / TO Metrics) also ms_ssim and per_image_ms_ssim."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if (shard and dist.is_initialized()) else (0, 1)
    per, per_ms = [], []
    fwd = GraphedForward(model) if graph else None
    for i, x in enumerate(images):
        if i % world != rank:
            continue
        h, w = x.size(2), x.size(3)
        if fwd is not None:
            out, bits = fwd(pad(x, p))
            n_, _, hh, ww = out["x_hat"].shape
            rec = crop(out["x_hat"], (h, w))
            per.append((compute_psnr(rec, x, clamp=True), float(bits) / (n_ * hh * ww)))
        else:
            with ops.defer_actq():
                out = model.forward(pad(x, p))
            rec = crop(out["x_hat"], (h, w))
            per.append((compute_psnr(rec, x, clamp=True), compute_bpp(out)))
        if ms_ssim:
            per_ms.append(compute_msssim(rec.clamp(0, 1), x))      # the reconstruction is clamped first (:98)
    acc = torch.tensor([sum(v[0] for v in per), sum(v[1] for v in per), float(len(per)), sum(per_ms)],
                       dtype=torch.float64, device=images[0].device)
    if world > 1:
        dist.all_reduce(acc)
    psnr_sum, bpp_sum, cnt, ms_sum = acc.tolist()
    res = dict(psnr=psnr_sum / max(cnt, 1), bpp=bpp_sum / max(cnt, 1), count=int(cnt), per_image=per)
    if ms_ssim:
        res.update(ms_ssim=ms_sum / max(cnt, 1), per_image_ms_ssim=per_ms)
    return res


class RateDistortionLoss(nn.Module):
    """losses/losses.py:8-35: bpp_loss + lmbda * 255^2 * mse (metric='mse') or bpp_loss + lmbda * (1 - MS-SSIM)
    (metric='ms-ssim'); values only: PTQ never back-props it.  `ms_ssim_loss` is reported for both metrics like the
    reference whenever the image is large enough for five levels (sides > 160 px; the reference raises below that)."""

    def __init__(self, lmbda=1e-2, metric='mse'):
        super().__init__()
        if metric not in ('mse', 'ms-ssim'):
            raise ValueError(f"metric {metric!r}")
        self.lmbda, self.metric = lmbda, metric

    def forward(self, output, target):
        N, _, H, W = target.size()
        bpp = float(total_bits(output)) / (N * H * W)
        mse = float(squared_error_sums(output["x_hat"], target)[0]) / target.numel()
        out = {"bpp_loss": bpp, "mse_loss": mse}
        if min(H, W) > 160:
            out["ms_ssim_loss"] = 1.0 - compute_msssim(output["x_hat"], target)
        if self.metric == 'mse':
            out["loss"] = self.lmbda * 255 ** 2 * mse + bpp
        else:
            if "ms_ssim_loss" not in out:
                raise ValueError("metric='ms-ssim' needs image sides > 160 px")
            out["loss"] = self.lmbda * out["ms_ssim_loss"] + bpp
        return out
