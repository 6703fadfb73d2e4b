"""B200 drop-in for task-oriented-PTQ/quantization/quant_layer.py (QuantModule, f_gdn).

Same constructor, attributes and state machine as the reference (quant_layer.py:11-139); the forward issues
libb200lic kernels: weight quantiser (K7/K6) -> implicit-GEMM conv / transposed conv / fused GDN (K1/K2/K3) with the
absorbed activation applied in the conv epilogue -> dynamic activation quantiser (K8).
"""
from typing import Union

import torch
import torch.nn as nn

from .. import ops
from ..codec.layers import GDN, f_gdn      # noqa: F401  (f_gdn is re-exported like the reference module)
from .quantizer import StraightThrough, UniformAffineQuantizer


class QuantModule(nn.Module):
    def __init__(self, org_module: Union[nn.Conv2d, nn.ConvTranspose2d, GDN, nn.PixelShuffle],
                 weight_quant_params: dict = {}, act_quant_params: dict = {}, disable_act_quant: bool = False,
                 se_module=None):
        super().__init__()
        self.if_layer_norm = False
        self.if_tconv = False
        self.is_ps = False
        self.is_gdn = False
        if isinstance(org_module, nn.ConvTranspose2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   output_padding=org_module.output_padding, dilation=org_module.dilation,
                                   groups=org_module.groups)
            self.fwd_func = ops.conv_transpose2d
            self.if_tconv = True
        elif isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = ops.conv2d
        elif isinstance(org_module, GDN):
            self.fwd_kwargs = dict(inverse=org_module.inverse, gamma_reparam=org_module.gamma_reparam,
                                   beta_reparam=org_module.beta_reparam)
            self.fwd_func = f_gdn
            self.is_gdn = True
        elif isinstance(org_module, nn.PixelShuffle):
            self.fwd_kwargs = org_module.upscale_factor
            self.fwd_func = ops.pixel_shuffle
            self.is_ps = True
        elif isinstance(org_module, (nn.Linear, nn.LayerNorm)):
            raise NotImplementedError("Linear/LayerNorm wrappers belong to the Lu2022 Swin codec, which is outside the "
                                      "B200 hot path (SURVEY.md 8(f) N4)")
        else:
            raise ValueError('Not supported modules: {}'.format(org_module))

        if self.is_gdn:
            self.weight, self.bias = org_module.gamma, org_module.beta
        elif self.is_ps:
            self.weight = self.bias = None
        else:
            self.weight, self.bias = org_module.weight, org_module.bias
        self.org_weight = None if self.weight is None else self.weight.data.clone()
        self.org_bias = None if self.bias is None else self.bias.data.clone()

        self.use_weight_quant = False
        self.use_act_quant = False
        self.disable_act_quant = disable_act_quant
        self.weight_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **weight_quant_params)
        self.act_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **act_quant_params)
        self.activation_function = nn.LeakyReLU(inplace=True) if self.is_ps else StraightThrough()   # reference :100
        self.ignore_reconstruction = False
        self.se_module = se_module
        self.trained = False
        self._repr = org_module.extra_repr()

    def extra_repr(self):
        return self._repr

    def forward(self, input: torch.Tensor):
        act, slope = ops._act_id(self.activation_function)
        if self.is_ps:
            return ops.pixel_shuffle(input, self.fwd_kwargs, act, slope)
        out = None
        if self.use_weight_quant and not self.is_gdn and not torch.is_grad_enabled():
            # hard-quantised weight, no gradient wanted (evaluation): integer weights + per-channel scale in the conv
            # epilogue, two tensor-core passes instead of three (b200lic_conv_fwd_wq); same value up to fp32 rounding
            iw = getattr(self.weight_quantizer, "int_weights", lambda _w: None)(self.weight)
            if iw is not None:
                out = ops.conv_wq(input, iw[0], iw[1], self.bias, transposed=self.if_tconv, act=act, slope=slope,
                                  **self.fwd_kwargs)
        if out is not None:
            weight = bias = None
        elif self.use_weight_quant:
            weight, bias = self.weight_quantizer(self.weight), self.bias
        else:
            weight, bias = self.org_weight, self.org_bias
        if out is not None:
            pass
        elif self.is_gdn:
            out = self.fwd_func(input, weight, bias, **self.fwd_kwargs)
            if act != ops.ACT_NONE:
                out = ops.add_act_fn(out, None, act, slope)
        else:
            out = self.fwd_func(input, weight, bias, act=act, slope=slope, **self.fwd_kwargs)
        if self.se_module is not None:
            raise NotImplementedError("se_module is never set on the LIC hot path")
        if self.disable_act_quant:
            return out
        if self.use_act_quant and self.trained:
            out = self.act_quantizer(out, True)
        return out

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
