"""The Lu2022 codec graph (task-oriented-PTQ/models/nic_cvt.py:21-330: `NIC`) over libb200lic modules, FORWARD ONLY
(SURVEY 8(f) N4).  Attribute names (`g_a0` ... `g_s7`, `h_a0` ... `h_s3`, `entropy_bottleneck`, `gaussian_conditional`,
`context_prediction`, `entropy_parameters`) and parameter shapes follow the reference so that its checkpoints load and
`QuantModel` rewrites the tree the same way (RSTB -> QuantRSTB, convolutions -> QuantModule).

State of this file: the graph and its rewrite are pinned on the CPU against the reference's own `NIC`
(tests/test_host_logic.py: state-dict keys and shapes, QuantModel module tree); every module it is composed of is verified
on the GPU against the reference (tests/test_gpu_tokens.py, tests/test_gpu_model.py), but the composed forward has not
been run on hardware in this round -- it has no benchmark leg and no parity number yet.
"""
import torch
import torch.nn as nn

from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import Conv2d, ConvTranspose2d, MaskedConv2d
from .swin import RSTB

DEPTHS = (2, 4, 6, 2, 2, 2, 2, 2, 2, 6, 4, 2)
HEADS = (4, 8, 8, 16, 16, 16, 16, 16, 16, 8, 8, 4)


class NIC(nn.Module):
    def __init__(self, config):
        super().__init__()
        H, W = config['height'], config['width']
        E, M, ws = config['embed_dim'], config['latent_dim'], config['window_size']
        cin = config['in_chans']
        self.M = M
        kw = dict(mlp_ratio=config['mlp_ratio'], qkv_bias=config['qkv_bias'], qk_scale=config['qk_scale'])

        def rstb(i, dim, div, window):
            return RSTB(dim=dim, input_resolution=(H // div, W // div), depth=DEPTHS[i], num_heads=HEADS[i],
                        window_size=window, **kw)

        def down(i, o, k=3):
            return Conv2d(i, o, kernel_size=k, stride=2, padding=k // 2)

        def up(i, o, k=3):
            return ConvTranspose2d(i, o, kernel_size=k, stride=2, padding=k // 2, output_padding=1)

        self.g_a0, self.g_a1 = down(cin, E, 5), rstb(0, E, 2, ws)
        self.g_a2, self.g_a3 = down(E, E), rstb(1, E, 4, ws)
        self.g_a4, self.g_a5 = down(E, E), rstb(2, E, 8, ws)
        self.g_a6, self.g_a7 = down(E, M), rstb(3, M, 16, ws)
        self.h_a0, self.h_a1 = down(M, E), rstb(4, E, 32, ws // 2)
        self.h_a2, self.h_a3 = down(E, E), rstb(5, E, 64, ws // 2)
        self.h_s0, self.h_s1 = rstb(6, E, 64, ws // 2), up(E, E)
        self.h_s2, self.h_s3 = rstb(7, E, 32, ws // 2), up(E, M * 2)
        self.g_s0, self.g_s1 = rstb(8, M, 16, ws), up(M, E)
        self.g_s2, self.g_s3 = rstb(9, E, 8, ws), up(E, E)
        self.g_s4, self.g_s5 = rstb(10, E, 4, ws), up(E, E)
        self.g_s6, self.g_s7 = rstb(11, E, 2, ws), up(E, cin, 5)
        self.entropy_bottleneck = EntropyBottleneck(E)
        self.gaussian_conditional = GaussianConditional(None)
        self.context_prediction = MaskedConv2d(M, M * 2, kernel_size=5, padding=2, stride=1)
        self.entropy_parameters = nn.Sequential(
            Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
            Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
            Conv2d(M * 8 // 3, M * 6 // 3, 1))
        for m in self.modules():                                   # nic_cvt.py:284-291
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    @staticmethod
    def _chain(x, size, steps):
        """steps: (module, divisor) pairs; divisor None = a (transposed) convolution, else an RSTB on a size // divisor map."""
        for m, div in steps:
            x = m(x) if div is None else m(x, (size[0] // div, size[1] // div))
        return x

    def g_a(self, x, x_size=None):
        s = x.shape[2:4] if x_size is None else x_size
        return self._chain(x, s, [(self.g_a0, None), (self.g_a1, 2), (self.g_a2, None), (self.g_a3, 4),
                                  (self.g_a4, None), (self.g_a5, 8), (self.g_a6, None), (self.g_a7, 16)])

    def g_s(self, x, x_size=None):
        s = (x.shape[2] * 16, x.shape[3] * 16) if x_size is None else x_size
        return self._chain(x, s, [(self.g_s0, 16), (self.g_s1, None), (self.g_s2, 8), (self.g_s3, None),
                                  (self.g_s4, 4), (self.g_s5, None), (self.g_s6, 2), (self.g_s7, None)])

    def h_a(self, x, x_size=None):
        s = (x.shape[2] * 16, x.shape[3] * 16) if x_size is None else x_size
        return self._chain(x, s, [(self.h_a0, None), (self.h_a1, 32), (self.h_a2, None), (self.h_a3, 64)])

    def h_s(self, x, x_size=None):
        s = (x.shape[2] * 64, x.shape[3] * 64) if x_size is None else x_size
        return self._chain(x, s, [(self.h_s0, 64), (self.h_s1, None), (self.h_s2, 32), (self.h_s3, None)])

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def forward(self, x):
        size = (x.shape[2], x.shape[3])
        y = self.g_a(x, size)
        z = self.h_a(y, size)
        z_hat, z_lik = self.entropy_bottleneck(z)
        params = self.h_s(z_hat, size)
        y_hat = self.gaussian_conditional.quantize(y, "dequantize")
        ctx = self.context_prediction(y_hat)
        scales, means = self.entropy_parameters(torch.cat((params, ctx), dim=1)).chunk(2, 1)
        _, y_lik = self.gaussian_conditional(y, scales, means=means)
        return {"x_hat": self.g_s(y_hat, size), "likelihoods": {"y": y_lik, "z": z_lik}}
