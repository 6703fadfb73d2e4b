"""Oracle restatement of MS-SSIM (test infrastructure only, CPU; never imported by the product).

The reference takes MS-SSIM from the third-party package `pytorch_msssim==1.0.0` (task-oriented-PTQ/requirements.txt;
call sites: TO/losses/losses.py:5,26,31,49-52, LU/quantize.py:13,89, LU/quant.py:13,86, LU/single_test.py:59-60,
LU/dataset_test.py:60-61 -- always `ms_ssim(a, b, data_range=1.)` on [N,3,H,W] tensors in [0,1]).  The package is not in
/root/reference and not installable here, so this file restates its published algorithm (Wang, Simoncelli, Bovik 2003 as
implemented by pytorch_msssim 1.0.0: `_fspecial_gauss_1d`, `gaussian_filter`, `_ssim`, `ms_ssim`):

  * window: 11 taps, sigma 1.5, normalised; separable VALID filtering, the height axis first, then the width axis;
  * per level: mu, sigma from the filtered x, y, x*x, y*y, x*y; cs = (2 s12 + C2) / (s1 + s2 + C2),
    ssim = (2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1) * cs, C1 = (0.01 L)^2, C2 = (0.03 L)^2; per-(image, channel) means;
  * five levels with weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333); between levels 2x2 average pooling with
    padding = size % 2 per axis (zeros, counted in the average); relu on cs of levels 0..3 and on ssim of level 4;
  * result = mean over (image, channel) of prod_l value_l ^ weight_l.

PARITY UNPINNED against the package itself (dependency absent); pinned here by closed forms and by an independent
float64 direct (non-separable, 121-tap) evaluation in tests/test_oracle_closed_forms.py.
"""
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
WIN_SIZE, WIN_SIGMA, K1, K2 = 11, 1.5, 0.01, 0.03


def gauss_1d(size=WIN_SIZE, sigma=WIN_SIGMA, dtype=torch.float32):
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_filter(x, win):
    """Depth-wise VALID filtering, height axis first (pytorch_msssim.gaussian_filter walks input.shape[2:] in order)."""
    c = x.shape[1]
    out = F.conv2d(x, win.view(1, 1, -1, 1).repeat(c, 1, 1, 1), groups=c)
    return F.conv2d(out, win.view(1, 1, 1, -1).repeat(c, 1, 1, 1), groups=c)


def ssim_level(x, y, data_range=1.0, win=None):
    """(ssim, cs) per (image, channel) of one level."""
    win = gauss_1d(dtype=x.dtype) if win is None else win
    c1, c2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    mu1, mu2 = gaussian_filter(x, win), gaussian_filter(y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = gaussian_filter(x * x, win) - mu1_sq
    s2 = gaussian_filter(y * y, win) - mu2_sq
    s12 = gaussian_filter(x * y, win) - mu12
    cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def downsample(x):
    return F.avg_pool2d(x, kernel_size=2, padding=[s % 2 for s in x.shape[2:]])


def ms_ssim(x, y, data_range=1.0, size_average=True):
    if min(x.shape[-2:]) <= (WIN_SIZE - 1) * 2 ** 4:
        raise ValueError("ms_ssim: image side must exceed 160 px (four 2x downsamplings, 11-tap window)")
    w = torch.tensor(WEIGHTS, dtype=x.dtype)
    vals = []
    for lvl in range(len(WEIGHTS)):
        ssim_pc, cs = ssim_level(x, y, data_range)
        if lvl < len(WEIGHTS) - 1:
            vals.append(torch.relu(cs))
            x, y = downsample(x), downsample(y)
    vals.append(torch.relu(ssim_pc))
    out = torch.prod(torch.stack(vals, 0) ** w.view(-1, 1, 1), dim=0)
    return out.mean() if size_average else out.mean(1)
