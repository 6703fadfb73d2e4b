"""compressai-style layers whose forward runs on libb200lic kernels.

Class / parameter names follow compressai 1.2.4 (the reference's un-vendored dependency) so that pickled
models and `state_dict`s line up (`gdn.beta`, `gdn.gamma`, cf. task-oriented-PTQ/ckpts/pretrained.py:47-56) and
`QuantModel`'s isinstance-based module swap (quant_model.py:37-56) keeps working: `Conv2d`/`ConvTranspose2d`
subclass the torch containers, only `forward` is replaced.
"""
import torch
import torch.nn as nn

from .. import ops


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters, sm_100a forward (K1)."""

    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class ConvTranspose2d(nn.ConvTranspose2d):
    """nn.ConvTranspose2d parameters, sm_100a forward (K2)."""

    def forward(self, x, output_size=None):
        return ops.conv_transpose2d(x, self.weight, self.bias, self.stride, self.padding, self.output_padding,
                                    self.groups, self.dilation)


class LeakyReLU(nn.LeakyReLU):
    def forward(self, x):
        return ops.add_act_fn(x, None, ops.ACT_LEAKY_RELU, float(self.negative_slope))


class ReLU(nn.ReLU):
    def forward(self, x):
        return ops.add_act_fn(x, None, ops.ACT_RELU, 0.0)


class PixelShuffle(nn.PixelShuffle):
    def forward(self, x):
        return ops.pixel_shuffle(x, self.upscale_factor)


class LowerBound(nn.Module):
    """compressai.ops.LowerBound: max(x, bound) with the 'pass if moving towards the bound' gradient."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):      # used by GaussianConditional through the fused K9 kernel; kept for API parity
        raise RuntimeError("LowerBound is evaluated inside the fused kernels (K3 reparam / K9 scale bound)")


class NonNegativeParametrizer(nn.Module):
    """compressai.ops.NonNegativeParametrizer; forward = max(x, bound)^2 - pedestal via K3's reparam kernel."""

    def __init__(self, minimum: float = 0.0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.bound_value = (self.minimum + pedestal) ** 0.5
        self.pedestal_value = pedestal
        self.lower_bound = LowerBound(self.bound_value)

    def init(self, x):          # construction-time only (CPU)
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return ops.gdn_reparam_fn(x, self.bound_value, self.pedestal_value)


class GDN(nn.Module):
    """compressai.layers.GDN: y = x * (beta' + gamma' . x^2)^-1/2 (inverse: ^+1/2), fused in K3."""

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        return f_gdn(x, self.gamma, self.beta, self.inverse, self.gamma_reparam, self.beta_reparam)


def f_gdn(x, gamma, beta, inverse, gamma_reparam, beta_reparam):
    """Functional GDN with the reference's signature (quant_layer.py:142-154)."""
    return ops.gdn(x, gamma_reparam(gamma), beta_reparam(beta), inverse)


class MaskedConv2d(Conv2d):
    """compressai.layers.MaskedConv2d: the forward bakes the causal mask into weight.data."""

    def __init__(self, *args, mask_type: str = "A", **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.shape
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask     # parameter plumbing, once per forward, not a data-path op
        return super().forward(x)


def conv(i, o, kernel_size=5, stride=2):
    return Conv2d(i, o, kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(i, o, kernel_size=5, stride=2):
    return ConvTranspose2d(i, o, kernel_size, stride=stride, output_padding=stride - 1, padding=kernel_size // 2)


def conv3x3(i, o, stride=1):
    return Conv2d(i, o, 3, stride=stride, padding=1)


def conv1x1(i, o, stride=1):
    return Conv2d(i, o, 1, stride=stride)


def subpel_conv3x3(i, o, r=1):
    return nn.Sequential(Conv2d(i, o * r ** 2, 3, padding=1), PixelShuffle(r))


class ResidualBlockWithStride(nn.Module):
    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride)
        self.leaky_relu = LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride) if stride != 1 or in_ch != out_ch else None

    def forward(self, x):
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        return ops.add_act_fn(out, self.skip(x) if self.skip is not None else x)


class ResidualBlockUpsample(nn.Module):
    def __init__(self, in_ch, out_ch, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return ops.add_act_fn(out, self.upsample(x))


class ResidualBlock(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        return ops.add_act_fn(out, self.skip(x) if self.skip is not None else x)


class _ResidualUnit(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), ReLU(inplace=True), conv3x3(N // 2, N // 2), ReLU(inplace=True),
                                  conv1x1(N // 2, N))
        self.relu = ReLU(inplace=True)

    def forward(self, x):
        return ops.add_act_fn(self.conv(x), x, ops.ACT_RELU, 0.0)


class AttentionBlock(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv_a = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N))
        self.conv_b = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N), conv1x1(N, N))

    def forward(self, x):
        return ops.attn_gate(self.conv_a(x), self.conv_b(x), x)


class GELU(nn.Module):
    """nn.GELU() (exact form) on b200lic_gelu_fwd."""

    def forward(self, x):
        return ops.gelu(x)


class Mlp(nn.Module):
    """task-oriented-PTQ/models/layers.py:35-52 (the feed-forward half of a Swin block): fc1 -> GELU -> fc2 over token
    tensors [..., C]; dropout is the identity at evaluation.  The Linear layers run on the tcgen05 conv engine
    (ops.linear).  Forward only: see SURVEY 8(f) N4."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        x = ops.linear(x, self.fc1.weight, self.fc1.bias)
        x = self.act(x)
        return ops.linear(x, self.fc2.weight, self.fc2.bias)
