"""The three BASELINE architectures (compressai 1.2.4 graphs: bmshj2018-hyperprior, mbt2018-mean,
cheng2020-attn) assembled from the sm_100a layers.  Module / parameter names follow compressai so the reference's
`QuantModel` walk (`g_a`, `g_s`, `h_a`, `h_s`, `entropy_bottleneck`, `gaussian_conditional`,
`context_prediction`, `entropy_parameters`; main2.py:256-263) applies unchanged.
"""
import torch
import torch.nn as nn

from .. import ops
from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import (GDN, MaskedConv2d, AttentionBlock, ResidualBlock, ResidualBlockUpsample,
                     ResidualBlockWithStride, Conv2d, LeakyReLU, ReLU, conv, deconv, conv3x3, subpel_conv3x3)


class ScaleHyperprior(nn.Module):
    def __init__(self, N=128, M=192):
        super().__init__()
        self.N, self.M = N, M
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), ReLU(inplace=True), conv(N, N),
                                 ReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), ReLU(inplace=True), deconv(N, N), ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(ops.abs_(y))
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales = self.h_s(z_hat)
        y_hat, y_lik = self.gaussian_conditional(y, scales)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}

    # -- real entropy coding (compressai CompressionModel.update / ScaleHyperprior.compress / decompress) ----------------
    def update(self, scale_table=None, force=False):
        from .coding import get_scale_table
        a = self.gaussian_conditional.update_scale_table(get_scale_table() if scale_table is None else scale_table, force)
        b = self.entropy_bottleneck.update(force)
        return a or b

    def _scales_means(self, params):
        return params, None

    @torch.no_grad()
    def compress(self, x):
        """{"strings": [y_strings, z_strings], "shape": z spatial size}; one string per image and latent."""
        y = self.g_a(x)
        z = self.h_a(self._hyper_in(y))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.shape[-2:])
        scales, means = self._scales_means(self.h_s(z_hat))
        indexes = self.gaussian_conditional.build_indexes(scales)
        y_strings = self.gaussian_conditional.compress(y, indexes, means=means)
        return {"strings": [y_strings, z_strings], "shape": tuple(z.shape[-2:])}

    @torch.no_grad()
    def decompress(self, strings, shape):
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        scales, means = self._scales_means(self.h_s(z_hat))
        indexes = self.gaussian_conditional.build_indexes(scales)
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, means=means)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}

    # -- tail of the forward from the output of one sub-network (R + lambda*D task criterion of the calibration) -------
    def _hyper_in(self, y):
        return ops.abs_fn(y)

    def _gauss(self, y, params):
        return self.gaussian_conditional(y, params, training=False)

    def latents(self, x):
        """(y, z) of an image batch: the upstream context `forward_from` needs for units in h_a / h_s / g_s."""
        y = self.g_a(x)
        return y, self.h_a(self._hyper_in(y))

    def forward_from(self, coder, value, ctx=None):
        """Continue the forward with `value` standing in for the OUTPUT of sub-network `coder` ('g_a': y, 'h_a': z,
        'h_s': the Gaussian parameters, 'g_s': x_hat); whatever lies upstream comes from ctx['y'] / ctx['z'].
        Returns the usual dict plus 'bits' = sum(-log2 likelihood) over y and z as a (differentiable) device scalar."""
        ctx = ctx or {}
        if coder not in ("g_a", "h_a", "h_s", "g_s"):
            raise NotImplementedError(f"forward_from: no tail defined after {coder!r}")
        y = value if coder == "g_a" else ctx["y"]
        z = self.h_a(self._hyper_in(y)) if coder == "g_a" else (value if coder == "h_a" else ctx["z"])
        z_hat, z_lik = self.entropy_bottleneck(z, training=False)     # evaluation-mode rounding, whatever .training says
        bits_z = self.entropy_bottleneck.last_bits
        params = value if coder == "h_s" else self.h_s(z_hat)
        y_hat, y_lik = self._gauss(y, params)
        bits = ops.add_act_fn(self.gaussian_conditional.last_bits, bits_z)
        x_hat = value if coder == "g_s" else self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "bits": bits}


class MeanScaleHyperprior(ScaleHyperprior):
    def __init__(self, N=128, M=192):
        super().__init__(N, M)
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), LeakyReLU(inplace=True), conv(N, N),
                                 LeakyReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, M), LeakyReLU(inplace=True), deconv(M, M * 3 // 2),
                                 LeakyReLU(inplace=True), conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales, means = self.h_s(z_hat).chunk(2, 1)           # strided views; K9 reads them in place
        y_hat, y_lik = self.gaussian_conditional(y, scales, means=means)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}

    def _hyper_in(self, y):
        return y

    def _scales_means(self, params):
        return params.chunk(2, 1)

    def _gauss(self, y, params):
        scales, means = params.chunk(2, 1)
        return self.gaussian_conditional(y, scales, means=means, training=False)


class Cheng2020Attention(nn.Module):
    """compressai's cheng2020-attn: single Gaussian (mean, scale) with the masked 5x5 context model evaluated over
    the whole latent in parallel (no GMM in the reference path)."""

    def __init__(self, N=192):
        super().__init__()
        self.N = self.M = M = N
        RBWS, RBU, RB, AB = ResidualBlockWithStride, ResidualBlockUpsample, ResidualBlock, AttentionBlock
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(RBWS(3, N, 2), RB(N, N), RBWS(N, N, 2), AB(N), RB(N, N), RBWS(N, N, 2), RB(N, N),
                                 conv3x3(N, N, stride=2), AB(N))
        self.h_a = nn.Sequential(conv3x3(N, N), LeakyReLU(inplace=True), conv3x3(N, N), LeakyReLU(inplace=True),
                                 conv3x3(N, N, stride=2), LeakyReLU(inplace=True), conv3x3(N, N),
                                 LeakyReLU(inplace=True), conv3x3(N, N, stride=2))
        self.h_s = nn.Sequential(conv3x3(N, N), LeakyReLU(inplace=True), subpel_conv3x3(N, N, 2),
                                 LeakyReLU(inplace=True), conv3x3(N, N * 3 // 2), LeakyReLU(inplace=True),
                                 subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), LeakyReLU(inplace=True),
                                 conv3x3(N * 3 // 2, N * 2))
        self.g_s = nn.Sequential(AB(N), RB(N, N), RBU(N, N, 2), RB(N, N), RBU(N, N, 2), AB(N), RB(N, N),
                                 RBU(N, N, 2), RB(N, N), subpel_conv3x3(N, 3, 2))
        self.entropy_parameters = nn.Sequential(Conv2d(M * 12 // 3, M * 10 // 3, 1), LeakyReLU(inplace=True),
                                                Conv2d(M * 10 // 3, M * 8 // 3, 1), LeakyReLU(inplace=True),
                                                Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)

    def compress(self, x):
        raise NotImplementedError("cheng2020-attn: y is coded under the masked 5x5 context model, whose decoder is serial "
                                  "in raster order (compressai runs it on the CPU); the chunked rANS coder covers the "
                                  "hyperprior families (bmshj2018-hyperprior, mbt2018-mean) -- DESIGN.md section 8")

    def decompress(self, strings, shape):
        raise NotImplementedError("cheng2020-attn: see compress()")

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(y, "dequantize")
        ctx = self.context_prediction(y_hat)
        # channel concat is buffer plumbing (a strided copy), not arithmetic
        scales, means = self.entropy_parameters(torch.cat((params, ctx), dim=1)).chunk(2, 1)
        _, y_lik = self.gaussian_conditional(y, scales, means=means)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


ARCHS = {"bmshj2018-hyperprior": ScaleHyperprior, "mbt2018-mean": MeanScaleHyperprior,
         "cheng2020-attn": Cheng2020Attention}
