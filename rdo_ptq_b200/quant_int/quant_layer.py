"""B200 drop-in for light-uniform-PTQ/quant_int/quant_layer.py, channel-wise branch (:115-137): the first forward
stores uint8 codes in `weight.data`; every forward de-quantises `(codes - zp) * delta` (K7), runs the conv /
transposed conv with the absorbed activation and the Q8.8 activation quantiser fused in the epilogue (K1/K2)."""
import torch
import torch.nn as nn

from .. import ops
from .quantizer import StraightThrough, UniformAffineQuantizer


class QuantModule(nn.Module):
    def __init__(self, org_module, weight_quant_params: dict = {}, act_quant_params: dict = {},
                 disable_act_quant: bool = False, se_module=None):
        super().__init__()
        self.if_layer_norm = False
        self.if_tconv = isinstance(org_module, nn.ConvTranspose2d)
        if self.if_tconv:
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   output_padding=org_module.output_padding, dilation=org_module.dilation,
                                   groups=org_module.groups)
            self.fwd_func = ops.conv_transpose2d
        elif isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = ops.conv2d
        else:
            raise NotImplementedError("only Conv2d / ConvTranspose2d wrappers are on the B200 hot path "
                                      "(Linear/LayerNorm belong to TinyLIC, SURVEY.md section 2 row 19)")
        if not weight_quant_params.get("channel_wise", False):
            raise NotImplementedError("layer-wise branch is dead in the reference (channel_wise=True is hard-coded, "
                                      "quantize.py:144)")
        self.weight, self.bias = org_module.weight, org_module.bias
        self.use_weight_quant = False
        self.use_act_quant = False
        self.disable_act_quant = disable_act_quant
        self.weight_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **weight_quant_params)
        self.bias_quantizer = UniformAffineQuantizer(**weight_quant_params)
        self.act_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **act_quant_params)
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False
        self.se_module = se_module
        self.trained = False

    def forward(self, input: torch.Tensor):
        wq = self.weight_quantizer
        if not self.trained:
            codes_u8 = ops.wq_fake_quant(self.weight.data, *self._scale(), wq.channel_axis(self.weight), wq.n_levels,
                                         want=("u8",))
            self.weight.requires_grad_(False)
            self.weight.data = codes_u8
            self.trained = True
        w = ops.wq_dequant_u8(self.weight.data, wq.delta, wq.zero_point, wq.channel_axis(self.weight))
        act, slope = ops._act_id(self.activation_function)
        fuse_q88 = int(self.use_act_quant and not self.disable_act_quant and self.act_quantizer.leaf_param
                       and not self.act_quantizer.inited)
        out = self.fwd_func(input, w, self.bias.data if self.bias is not None else None, act=act, slope=slope,
                            fixed_pt=fuse_q88, **self.fwd_kwargs)
        if self.disable_act_quant or fuse_q88:
            return out
        if self.use_act_quant:
            out = self.act_quantizer(out, True)
        return out

    def _scale(self):
        wq = self.weight_quantizer
        if not wq.inited:
            wq(self.weight.data)        # initialises delta / zero_point (K7)
        return wq.delta, wq.zero_point

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
