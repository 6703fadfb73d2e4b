"""B200 drop-in for light-uniform-PTQ/quant_int/quantizer.py (:63-183): the weight quantiser returns
(uint8-range codes, delta); activations are static Q8.8 fixed point (:120-128)."""
import torch
import torch.nn as nn

from .. import ops
from ..quantization.quantizer import StraightThrough, round_ste      # noqa: F401


def ActQuantizer(x, a_l=8, a_r=8):
    """reference :120-128: round(clamp(x, -2^(a_l-1), 2^(a_l-1)) * 2^a_r) / 2^a_r."""
    return ops.fixed_point(x, a_l, a_r)


class UniformAffineQuantizer(nn.Module):
    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False, scale_method: str = 'max',
                 leaf_param: bool = False, tconv: bool = False, act: bool = False, prob: float = 1.0):
        super().__init__()
        self.sym, self.n_bits, self.n_levels = symmetric, n_bits, 2 ** n_bits
        self.delta = torch.empty(0)
        self.zero_point = torch.empty(0)
        self.inited = False
        self.leaf_param, self.channel_wise, self.scale_method = leaf_param, channel_wise, scale_method
        self.tconv, self.act, self.prob, self.is_training = tconv, act, prob, False

    def channel_axis(self, x):
        if not self.channel_wise or x.dim() == 1:
            return None
        return 1 if (self.tconv and x.dim() == 4) else 0

    def forward(self, x: torch.Tensor, act: bool = False):
        if self.inited is False:
            if self.leaf_param:                       # activations: never initialised => always Q8.8 (SURVEY a3')
                return ActQuantizer(x, a_l=8, a_r=8)
            if 'max' in self.scale_method:
                self.delta, self.zero_point = ops.wq_init_minmax(x.detach(), self.channel_axis(x), self.n_bits,
                                                                 'scale' in self.scale_method, self.sym)
            elif self.scale_method == 'mse':          # reference :265-278: 80 shrink steps of 1 %, squared-error score
                self.delta, self.zero_point = ops.wq_init_search(x.detach(), self.channel_axis(x), self.n_bits, 'mse',
                                                                 80, 0.01, 2.0, self.sym)
            else:
                raise NotImplementedError(self.scale_method)                 # as the reference, :279-280
            self.inited = True
        codes = ops.wq_fake_quant(x.detach(), self.delta, self.zero_point, self.channel_axis(x), self.n_levels,
                                  want=("codes",))
        return codes, self.delta
