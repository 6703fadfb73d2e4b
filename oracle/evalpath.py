"""Oracle restatement of the evaluation entry (test infrastructure, CPU fp32).

Follows /root/reference/task-oriented-PTQ/test_datasets.py:21-33,45-117 (pad-256 / crop / PSNR / bpp,
MS-SSIM dropped: pytorch_msssim is not installable here) and losses/losses.py:15-28.
"""
import math

import torch
import torch.nn.functional as F


def pad(x, p=256):                         # test_datasets.py:45-58
    h, w = x.size(2), x.size(3)
    H, W = (h + p - 1) // p * p, (w + p - 1) // p * p
    l, t = (W - w) // 2, (H - h) // 2
    return F.pad(x, (l, W - w - l, t, H - h - t), mode="constant", value=0)


def crop(x, size):                         # test_datasets.py:61-73
    H, W = x.size(2), x.size(3)
    h, w = size
    l, t = (W - w) // 2, (H - h) // 2
    return F.pad(x, (-l, -(W - w - l), -t, -(H - h - t)), mode="constant", value=0)


def compute_psnr(a, b):                    # :21-23
    return -10 * math.log10(torch.mean((a - b) ** 2).item())


def compute_bpp(out_net):                  # :29-33 (pixels of the PADDED x_hat)
    n, _, h, w = out_net["x_hat"].shape
    return sum(torch.log(l).sum() / (-math.log(2) * n * h * w) for l in out_net["likelihoods"].values()).item()


def evaluate(model, images):
    """Test_kodak (:76-117) over a list of [1,3,h,w] tensors.  Returns per-image (psnr, bpp) lists."""
    ps, bs = [], []
    for x in images:
        h, w = x.size(2), x.size(3)
        with torch.no_grad():
            out = model(pad(x, 256))
        rec = crop(out["x_hat"], (h, w)).clamp_(0, 1)
        ps.append(compute_psnr(x, rec))
        bs.append(compute_bpp(out))
    return ps, bs


def rate_distortion_loss(out, target, lmbda=1e-2):
    """losses.py:15-28, metric='mse' (MS-SSIM term dropped)."""
    n, _, h, w = target.shape
    bpp = sum((-torch.log2(l).sum() / (n * h * w)) for l in out["likelihoods"].values())
    mse = F.mse_loss(out["x_hat"], target)
    return {"bpp_loss": bpp, "mse_loss": mse, "loss": lmbda * 255 ** 2 * mse + bpp}
