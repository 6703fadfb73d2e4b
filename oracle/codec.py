"""Oracle restatement of the `compressai==1.2.4` pieces on the hot path (test infrastructure).

compressai is a pinned, un-vendored dependency of the reference (requirements.txt:1) and cannot
be installed here, so its published semantics are restated.  Parameter
names match compressai's state_dict (`gdn.beta/gamma`, `entropy_bottleneck._matrix{i}` ..., cf. the
key renamer at /root/reference/task-oriented-PTQ/ckpts/pretrained.py:47-56).  Reference call sites:
GDN quant_layer.py:7,51-56,142-154; entropy models nic_cvt.py:221-222,297-308; residual blocks
quant_block.py:8,219-328.  PARITY UNPINNED for this file (no reference fixtures exist); checked
against closed forms in tests/test_oracle_closed_forms.py.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------- GDN (a15, a7)
class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)).type_as(g) * g, None


class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum: float = 0.0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + pedestal) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(nn.Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        return f_gdn(x, self.gamma, self.beta, self.inverse, self.gamma_reparam, self.beta_reparam)


def f_gdn(x, gamma, beta, inverse, gamma_reparam, beta_reparam):
    """TO quant_layer.py:142-154 (identical to compressai GDN.forward)."""
    C = x.shape[1]
    norm = F.conv2d(x ** 2, gamma_reparam(gamma).reshape(C, C, 1, 1), beta_reparam(beta))
    return x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))


# ---------------------------------------------------------------------------- entropy models (a16, a17)
def quantize_latent(inputs, mode, means=None, ste=False):
    """compressai EntropyModel.quantize (the reference carries a copy at TO quantizer.py:19-48).
    `ste` (extension, not compressai): round_ste of TO quantizer.py:64-68 instead of torch.round -- the straight-through
    rounding the reference applies to y in fp_out (layer_opt.py:69); forward values are identical."""
    if mode == "noise":
        return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
    out = inputs.clone()
    if means is not None:
        out -= means
    out = (torch.round(out) - out).detach() + out if ste else torch.round(out)
    if mode == "dequantize":
        if means is not None:
            out += means
        return out
    assert mode == "symbols"
    return out.int()


class EntropyBottleneck(nn.Module):
    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), likelihood_bound=1e-9):
        super().__init__()
        self.channels, self.filters = int(channels), tuple(int(f) for f in filters)
        self.init_scale, self.tail_mass = float(init_scale), float(tail_mass)
        self.likelihood_bound = float(likelihood_bound)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / f[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, f[i + 1], f[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, f[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        q = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles = nn.Parameter(q.repeat(channels, 1, 1))
        t = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-t, 0, t]))
        self.ste_round = False

    def likelihood_lower_bound(self, lik):
        """compressai's LowerBound on the likelihood (applied functionally: no extra state_dict key)."""
        return _LowerBoundFn.apply(lik, lik.new_tensor([self.likelihood_bound]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, v, stop_gradient=False):
        for i in range(len(self.filters) + 1):
            m, b = getattr(self, f"_matrix{i:d}"), getattr(self, f"_bias{i:d}")
            if stop_gradient:
                m, b = m.detach(), b.detach()
            v = torch.matmul(F.softplus(m), v) + b
            if i < len(self.filters):
                fac = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    fac = fac.detach()
                v = v + torch.tanh(fac) * torch.tanh(v)
        return v

    def _likelihood(self, v):
        lower, upper = self._logits_cumulative(v - 0.5), self._logits_cumulative(v + 0.5)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def loss(self):
        return torch.abs(self._logits_cumulative(self.quantiles, True) - self.target).sum()

    def quantize(self, inputs, mode, means=None):
        return quantize_latent(inputs, mode, means, self.ste_round)

    def forward(self, x, training=None):
        training = self.training if training is None else training
        xp = x.transpose(0, 1).contiguous()
        shape = xp.shape
        v = xp.reshape(shape[0], 1, -1)
        out = quantize_latent(v, "noise" if training else "dequantize", self._get_medians(), self.ste_round)
        lik = self.likelihood_lower_bound(self._likelihood(out))
        out = out.reshape(shape).transpose(0, 1).contiguous()
        lik = lik.reshape(shape).transpose(0, 1).contiguous()
        return out, lik


class GaussianConditional(nn.Module):
    def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9, likelihood_bound=1e-9):
        super().__init__()
        self.likelihood_bound = float(likelihood_bound)
        self.lower_bound_scale = LowerBound(scale_bound)
        self.ste_round = False

    def likelihood_lower_bound(self, lik):
        return _LowerBoundFn.apply(lik, lik.new_tensor([self.likelihood_bound]))

    def quantize(self, inputs, mode, means=None):
        return quantize_latent(inputs, mode, means, self.ste_round)

    @staticmethod
    def _standardized_cumulative(t):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * t)

    def _likelihood(self, inputs, scales, means=None):
        v = torch.abs(inputs - means if means is not None else inputs)
        s = self.lower_bound_scale(scales)
        return self._standardized_cumulative((0.5 - v) / s) - self._standardized_cumulative((-0.5 - v) / s)

    def forward(self, inputs, scales, means=None, training=None):
        training = self.training if training is None else training
        out = quantize_latent(inputs, "noise" if training else "dequantize", means, self.ste_round)
        lik = self.likelihood_lower_bound(self._likelihood(out, scales, means))
        return out, lik


# ---------------------------------------------------------------------------- layers (a18, a19)
class MaskedConv2d(nn.Conv2d):
    def __init__(self, *args, mask_type="A", **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.shape
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask        # bakes the mask in
        return super().forward(x)


def conv(i, o, kernel_size=5, stride=2):
    return nn.Conv2d(i, o, kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(i, o, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(i, o, kernel_size, stride=stride, output_padding=stride - 1, padding=kernel_size // 2)


def conv3x3(i, o, stride=1):
    return nn.Conv2d(i, o, 3, stride=stride, padding=1)


def conv1x1(i, o, stride=1):
    return nn.Conv2d(i, o, 1, stride=stride)


def subpel_conv3x3(i, o, r=1):
    return nn.Sequential(nn.Conv2d(i, o * r ** 2, 3, padding=1), nn.PixelShuffle(r))


class ResidualBlockWithStride(nn.Module):
    def __init__(self, i, o, stride=2):
        super().__init__()
        self.conv1 = conv3x3(i, o, stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(o, o)
        self.gdn = GDN(o)
        self.skip = conv1x1(i, o, stride) if stride != 1 or i != o else None

    def forward(self, x):
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        return out + (self.skip(x) if self.skip is not None else x)


class ResidualBlockUpsample(nn.Module):
    def __init__(self, i, o, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(i, o, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(o, o)
        self.igdn = GDN(o, inverse=True)
        self.upsample = subpel_conv3x3(i, o, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return out + self.upsample(x)


class ResidualBlock(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.conv1 = conv3x3(i, o)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(o, o)
        self.skip = conv1x1(i, o) if i != o else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        return out + (self.skip(x) if self.skip is not None else x)


class _ResidualUnit(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                  nn.ReLU(inplace=True), conv1x1(N // 2, N))
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.conv(x) + x)


class AttentionBlock(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv_a = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N))
        self.conv_b = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N), conv1x1(N, N))

    def forward(self, x):
        return self.conv_a(x) * torch.sigmoid(self.conv_b(x)) + x


# ---------------------------------------------------------------------------- model graphs
class ScaleHyperprior(nn.Module):
    """bmshj2018-hyperprior (BASELINE config 1)."""

    def __init__(self, N=128, M=192):
        super().__init__()
        self.N, self.M = N, M
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales = self.h_s(z_hat)
        y_hat, y_lik = self.gaussian_conditional(y, scales)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}

    # -- tail of the forward from one sub-network's output: what the (commented-out) R + lambda*D task criterion of
    #    layer_opt.py:146-148 needs; restates the same graph as forward() --------------------------------------------
    def _hyper_in(self, y):
        return torch.abs(y)

    def _gauss(self, y, params):
        return self.gaussian_conditional(y, params, training=False)

    def latents(self, x):
        y = self.g_a(x)
        return y, self.h_a(self._hyper_in(y))

    def forward_from(self, coder, value, ctx=None):
        ctx = ctx or {}
        y = value if coder == "g_a" else ctx["y"]
        z = self.h_a(self._hyper_in(y)) if coder == "g_a" else (value if coder == "h_a" else ctx["z"])
        z_hat, z_lik = self.entropy_bottleneck(z, training=False)     # evaluation-mode rounding, whatever .training says
        params = value if coder == "h_s" else self.h_s(z_hat)
        y_hat, y_lik = self._gauss(y, params)
        x_hat = value if coder == "g_s" else self.g_s(y_hat)
        bits = -torch.log2(y_lik).sum() - torch.log2(z_lik).sum()
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "bits": bits}


class MeanScaleHyperprior(ScaleHyperprior):
    """mbt2018-mean (BASELINE config 2)."""

    def __init__(self, N=128, M=192):
        super().__init__(N, M)
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True), conv(N, N),
                                 nn.LeakyReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, M), nn.LeakyReLU(inplace=True), deconv(M, M * 3 // 2),
                                 nn.LeakyReLU(inplace=True), conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales, means = self.h_s(z_hat).chunk(2, 1)
        y_hat, y_lik = self.gaussian_conditional(y, scales, means=means)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}

    def _hyper_in(self, y):
        return y

    def _gauss(self, y, params):
        scales, means = params.chunk(2, 1)
        return self.gaussian_conditional(y, scales, means=means, training=False)


class Cheng2020Attention(nn.Module):
    """cheng2020-attn as shipped by compressai: single Gaussian + parallel masked context (configs 3, 4)."""

    def __init__(self, N=192):
        super().__init__()
        self.N = self.M = M = N
        RBWS, RBU, RB, AB = ResidualBlockWithStride, ResidualBlockUpsample, ResidualBlock, AttentionBlock
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(RBWS(3, N, 2), RB(N, N), RBWS(N, N, 2), AB(N), RB(N, N), RBWS(N, N, 2),
                                 RB(N, N), conv3x3(N, N, stride=2), AB(N))
        self.h_a = nn.Sequential(conv3x3(N, N), nn.LeakyReLU(inplace=True), conv3x3(N, N), nn.LeakyReLU(inplace=True),
                                 conv3x3(N, N, stride=2), nn.LeakyReLU(inplace=True), conv3x3(N, N),
                                 nn.LeakyReLU(inplace=True), conv3x3(N, N, stride=2))
        self.h_s = nn.Sequential(conv3x3(N, N), nn.LeakyReLU(inplace=True), subpel_conv3x3(N, N, 2),
                                 nn.LeakyReLU(inplace=True), conv3x3(N, N * 3 // 2), nn.LeakyReLU(inplace=True),
                                 subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), nn.LeakyReLU(inplace=True),
                                 conv3x3(N * 3 // 2, N * 2))
        self.g_s = nn.Sequential(AB(N), RB(N, N), RBU(N, N, 2), RB(N, N), RBU(N, N, 2), AB(N), RB(N, N),
                                 RBU(N, N, 2), RB(N, N), subpel_conv3x3(N, 3, 2))
        self.entropy_parameters = nn.Sequential(nn.Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
                                                nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
                                                nn.Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(y, "noise" if self.training else "dequantize")
        ctx = self.context_prediction(y_hat)
        scales, means = self.entropy_parameters(torch.cat((params, ctx), dim=1)).chunk(2, 1)
        _, y_lik = self.gaussian_conditional(y, scales, means=means)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


ARCHS = {"bmshj2018-hyperprior": ScaleHyperprior, "mbt2018-mean": MeanScaleHyperprior,
         "cheng2020-attn": Cheng2020Attention}
