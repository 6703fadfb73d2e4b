"""Deterministic synthetic inputs and random-init weights.

No checkpoints or datasets can be fetched, so the benchmark and the parity tests run on seeded synthetic images
"in Kodak / CLIC shape" and on random-init weights.  Everything is generated on the CPU (seed 1005 = the
reference's default, main2.py:27) so the CPU oracle and the GPU path see identical bits.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

SEED = 1005


def synthetic_image(h, w, index=0, seed=SEED):
    """[1,3,h,w] fp32 in [0,1] on the 8-bit grid: smooth low-pass noise + fine noise (not white noise)."""
    g = torch.Generator().manual_seed(seed * 7919 + index)
    sigma, rad = 8.0, 24
    t = torch.arange(-rad, rad + 1, dtype=torch.float32)
    k = torch.exp(-t * t / (2 * sigma * sigma))
    k = (k / k.sum()).view(1, 1, -1)
    n = torch.randn(3, 1, h + 2 * rad, w + 2 * rad, generator=g)
    low = F.conv2d(F.conv2d(n, k.unsqueeze(2)), k.unsqueeze(3))           # separable Gaussian, sigma = 8
    low = low / low.std()
    fine = torch.randn(3, 1, h, w, generator=g)
    x = (0.5 + 0.25 * low + 0.05 * fine).clamp(0, 1)
    return (torch.round(x * 255) / 255).permute(1, 0, 2, 3).contiguous()


def synthetic_images(count, h, w, seed=SEED):
    return [synthetic_image(h, w, i, seed) for i in range(count)]


def calibration_patches(count, size=256, seed=SEED):
    """[count,3,size,size] stand-in for the RandomCrop(256) calibration patches (datasets/dataset.py:45-54)."""
    return torch.cat([synthetic_image(size, size, 10_000 + i, seed) for i in range(count)], dim=0)


@torch.no_grad()
def init_weights(model, seed=SEED, gain=1.0):
    """Random init of every Conv2d / ConvTranspose2d: uniform with variance gain/fan so activations stay O(1)
    through 20+ random layers (PyTorch's default kaiming(a=sqrt 5) shrinks them by 3x per layer, which would make
    every latent round to 0 and the entropy path degenerate).  GDN and EntropyBottleneck keep compressai's init.
    """
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.ConvTranspose2d):
            fan = m.in_channels * m.kernel_size[0] * m.kernel_size[1] / (m.stride[0] * m.stride[1])
        elif isinstance(m, nn.Conv2d):
            fan = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
        else:
            continue
        bound = math.sqrt(3.0 * gain / fan)
        m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)
        if m.bias is not None:
            m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * 0.05)
    for m in model.modules():
        if type(m).__name__ == "GDN":
            # compressai's init (beta=1, gamma=0.1*I) puts every gamma entry exactly on a quantisation grid point
            # (0 or the row maximum), which makes AdaRound's gradient degenerate; a trained GDN has dense gamma.
            c = m.beta.numel()
            gam = 0.1 * torch.eye(c) + 0.04 * torch.rand(c, c, generator=g) / max(1.0, c / 16.0)
            bet = 1.0 + 0.3 * torch.rand(c, generator=g)
            m.gamma.copy_(m.gamma_reparam.init(gam))
            m.beta.copy_(m.beta_reparam.init(bet))
    for m in model.modules():
        if type(m).__name__ == "EntropyBottleneck":
            for i in range(5):
                b = getattr(m, f"_bias{i:d}")
                b.copy_(torch.rand(b.shape, generator=g) - 0.5)
    return model
