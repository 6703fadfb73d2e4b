"""Oracle restatement of the reference's quant-module wrappers (test infrastructure, CPU fp32).

TO = /root/reference/task-oriented-PTQ/quantization, LU = /root/reference/light-uniform-PTQ/quant_int.
The reference files import compressai/timm, so they are restated here over `oracle.codec`; the quantizer arithmetic
they call is the pinned `oracle.quantizers`.  PINNED: `oracle/make_golden.py::wrap_vectors` imports the unmodified
reference packages through `oracle/_ref_shim.py` and asserts that every class below reproduces them bit for bit
(`tests/golden/wrap_ref.pt`, `tests/test_oracle_wrap_golden.py`).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import codec
from .quantizers import (UniformAffineQuantizer, LUUniformAffineQuantizer, act_quant, lu_act_quantizer)


class StraightThrough(nn.Module):
    def forward(self, x):
        return x


class QuantModule(nn.Module):
    """TO quant_layer.py:11-139."""

    def __init__(self, org_module, weight_quant_params=None, act_quant_params=None, disable_act_quant=False):
        super().__init__()
        wq, aq = dict(weight_quant_params or {}), dict(act_quant_params or {})
        self.if_tconv = self.is_ps = self.is_gdn = False
        if isinstance(org_module, nn.ConvTranspose2d):                      # :30-37
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   output_padding=org_module.output_padding, dilation=org_module.dilation,
                                   groups=org_module.groups)
            self.fwd_func, self.if_tconv = F.conv_transpose2d, True
        elif isinstance(org_module, nn.Conv2d):                             # :23-28
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = F.conv2d
        elif isinstance(org_module, nn.Linear):                             # :39-42
            self.fwd_kwargs, self.fwd_func = {}, F.linear
        elif isinstance(org_module, codec.GDN):                             # :51-56
            self.fwd_kwargs = dict(inverse=org_module.inverse, gamma_reparam=org_module.gamma_reparam,
                                   beta_reparam=org_module.beta_reparam)
            self.fwd_func, self.is_gdn = codec.f_gdn, True
        elif isinstance(org_module, nn.PixelShuffle):                       # :58-61
            self.fwd_kwargs, self.fwd_func, self.is_ps = org_module.upscale_factor, F.pixel_shuffle, True
        else:
            raise ValueError(f"Not supported modules: {org_module}")
        if self.is_gdn:                                                     # :67-75
            self.weight, self.bias = org_module.gamma, org_module.beta
        elif self.is_ps:
            self.weight = self.bias = None
        else:
            self.weight, self.bias = org_module.weight, org_module.bias
        self.org_weight = None if self.weight is None else self.weight.data.clone()
        self.org_bias = None if self.bias is None else self.bias.data.clone()
        self.use_weight_quant = self.use_act_quant = False
        self.disable_act_quant = disable_act_quant
        self.weight_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **wq)
        self.act_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **aq)
        self.activation_function = nn.LeakyReLU(inplace=True) if self.is_ps else StraightThrough()   # :100 (Q4)
        self.ignore_reconstruction = False
        self.trained = False

    def forward(self, x):                                                   # :107-134
        if self.is_ps:
            return self.activation_function(self.fwd_func(x, self.fwd_kwargs))
        if self.use_weight_quant:
            w, b = self.weight_quantizer(self.weight), self.bias
        else:
            w, b = self.org_weight, self.org_bias
        out = self.activation_function(self.fwd_func(x, w, b, **self.fwd_kwargs))
        if self.disable_act_quant:
            return out
        if self.use_act_quant and self.trained:
            out = self.act_quantizer(out, True)
        return out

    def set_quant_state(self, weight_quant=False, act_quant=False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant


class BaseQuantBlock(nn.Module):
    """TO quant_block.py:77-102."""

    def __init__(self, act_quant_params=None):
        super().__init__()
        self.use_weight_quant = self.use_act_quant = False
        self.trained = False
        self.act_quantizer = UniformAffineQuantizer(act=True, **dict(act_quant_params or {}))
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False

    def set_quant_state(self, weight_quant=False, act_quant=False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)

    def _aq(self, t):
        return self.act_quantizer(t, True) if (self.use_act_quant and self.trained) else t


class QuantRBWS(BaseQuantBlock):
    """TO quant_block.py:219-250."""

    def __init__(self, blk, wq=None, aq=None):
        super().__init__(aq)
        self.conv1 = QuantModule(blk.conv1, wq, aq, disable_act_quant=True)
        self.leaky_relu = blk.leaky_relu
        self.conv2 = QuantModule(blk.conv2, wq, aq)
        self.gdn = QuantModule(blk.gdn, wq, aq)
        self.skip = QuantModule(blk.skip, wq, aq) if blk.skip is not None else None

    def forward(self, x):
        out = self._aq(self.leaky_relu(self.conv1(x)))
        out = self.gdn(self.conv2(out))
        out = out + (self.skip(x) if self.skip is not None else x)
        return self._aq(out)


class QuantRBU(BaseQuantBlock):
    """TO quant_block.py:253-284."""

    def __init__(self, blk, wq=None, aq=None):
        super().__init__(aq)
        self.subpel_conv = nn.Sequential(QuantModule(blk.subpel_conv[0], wq, aq, disable_act_quant=True),
                                         blk.subpel_conv[1])
        self.leaky_relu = blk.leaky_relu
        self.conv = QuantModule(blk.conv, wq, aq)
        self.igdn = QuantModule(blk.igdn, wq, aq)
        self.upsample = nn.Sequential(QuantModule(blk.upsample[0], wq, aq), blk.upsample[1])

    def forward(self, x):
        out = self._aq(self.leaky_relu(self.subpel_conv(x)))
        out = self.igdn(self.conv(out))
        return self._aq(out + self.upsample(x))


class QuantRB(BaseQuantBlock):
    """TO quant_block.py:286-313."""

    def __init__(self, blk, wq=None, aq=None):
        super().__init__(aq)
        self.conv1 = QuantModule(blk.conv1, wq, aq, disable_act_quant=True)
        self.leaky_relu = blk.leaky_relu
        self.conv2 = QuantModule(blk.conv2, wq, aq, disable_act_quant=True)
        self.skip = QuantModule(blk.skip, wq, aq) if blk.skip is not None else None

    def forward(self, x):
        out = self._aq(self.leaky_relu(self.conv1(x)))
        out = self._aq(self.leaky_relu(self.conv2(out)))
        out = out + (self.skip(x) if self.skip is not None else x)
        return self._aq(out)


specials = {codec.ResidualBlockWithStride: QuantRBWS, codec.ResidualBlockUpsample: QuantRBU,
            codec.ResidualBlock: QuantRB}        # TO quant_block.py:645-657 (QuantSC is unreachable, Q4)


class QuantModel(nn.Module):
    """TO quant_model.py:10-98."""

    def __init__(self, model, weight_quant_params=None, act_quant_params=None, is_fusing=True, is_cheng=False):
        super().__init__()
        self.model = model
        self._refactor(model, weight_quant_params, act_quant_params)

    def _refactor(self, module, wq, aq):                                    # :23-62
        prev = None
        for name, child in module.named_children():
            if type(child) in specials:
                setattr(module, name, specials[type(child)](child, wq, aq))
            elif isinstance(child, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, codec.GDN, nn.PixelShuffle)):
                prev = QuantModule(child, wq, aq)
                setattr(module, name, prev)
            elif isinstance(child, (nn.LeakyReLU, nn.GELU, nn.ReLU, nn.ReLU6)):
                if prev is not None:
                    prev.activation_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, StraightThrough):
                continue
            else:
                self._refactor(child, wq, aq)

    def set_quant_state(self, weight_quant=False, act_quant=False):         # :64-67
        for m in self.model.modules():
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                m.set_quant_state(weight_quant, act_quant)

    def forward(self, x):
        return self.model(x)

    def quant_modules(self):
        return [m for m in self.model.modules() if isinstance(m, QuantModule)]

    def set_first_last_layer_to_8bit(self):                                 # :81-91
        ml = self.quant_modules()
        ml[0].weight_quantizer.bitwidth_refactor(8)
        ml[0].act_quantizer.bitwidth_refactor(8)
        ml[-1].weight_quantizer.bitwidth_refactor(8)
        ml[-2].act_quantizer.bitwidth_refactor(8)

    def disable_network_output_quantization(self):                          # :93-98
        self.quant_modules()[-1].disable_act_quant = True


# ---------------------------------------------------------------------------- light-uniform-PTQ
class LUQuantModule(nn.Module):
    """LU quant_layer.py:10-141, channel-wise branch (:115-137; layer-wise branch is dead, SURVEY a6')."""

    def __init__(self, org_module, weight_quant_params=None, act_quant_params=None, disable_act_quant=False):
        super().__init__()
        wq, aq = dict(weight_quant_params or {}), dict(act_quant_params or {})
        self.if_tconv = isinstance(org_module, nn.ConvTranspose2d)
        if self.if_tconv:
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   output_padding=org_module.output_padding, dilation=org_module.dilation,
                                   groups=org_module.groups)
            self.fwd_func = F.conv_transpose2d
        elif isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = F.conv2d
        elif isinstance(org_module, nn.Linear):
            self.fwd_kwargs, self.fwd_func = {}, F.linear
        else:
            raise ValueError(f"Not supported modules: {org_module}")
        self.weight, self.bias = org_module.weight, org_module.bias
        self.use_weight_quant = self.use_act_quant = False
        self.disable_act_quant = disable_act_quant
        self.weight_quantizer = LUUniformAffineQuantizer(tconv=self.if_tconv, **wq)
        self.act_quantizer = LUUniformAffineQuantizer(tconv=self.if_tconv, **aq)
        self.activation_function = StraightThrough()
        self.trained = False

    def forward(self, x):
        if not self.trained:                                     # :116-119 store uint8 codes in place
            codes, _ = self.weight_quantizer(self.weight)
            self.weight.requires_grad_(False)
            self.weight.data = codes.to(torch.uint8)
            self.trained = True
        w = (self.weight.type_as(x) - self.weight_quantizer.zero_point) * self.weight_quantizer.delta
        out = self.activation_function(self.fwd_func(x, w, self.bias.type_as(x), **self.fwd_kwargs))
        if self.disable_act_quant:
            return out
        if self.use_act_quant:
            out = self.act_quantizer(out, True)                  # leaf_param & never inited -> Q8.8 (a3')
        return out

    def set_quant_state(self, weight_quant=False, act_quant=False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant


class LUQuantModel(nn.Module):
    """LU quant_model.py:9-78 with a 1-argument forward; GDN stays fp32 (:23)."""

    def __init__(self, model, weight_quant_params=None, act_quant_params=None, skip_prefixes=()):
        super().__init__()
        self.model = model
        self._refactor(model, weight_quant_params, act_quant_params, skip_prefixes)

    def _refactor(self, module, wq, aq, skip):
        prev = None
        for name, child in module.named_children():
            if any(name.startswith(p) for p in skip):           # QuantCodingModel: skips g_a*/g_s* (:23-26)
                continue
            if isinstance(child, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
                prev = LUQuantModule(child, wq, aq)
                setattr(module, name, prev)
            elif isinstance(child, (nn.LeakyReLU, nn.GELU, nn.ReLU, nn.ReLU6)):
                if prev is not None:
                    prev.activation_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, StraightThrough):
                continue
            else:
                self._refactor(child, wq, aq, ())

    def set_quant_state(self, weight_quant=False, act_quant=False):
        for m in self.model.modules():
            if isinstance(m, LUQuantModule):
                m.set_quant_state(weight_quant, act_quant)

    def forward(self, x):
        return self.model(x)

    def disable_network_output_quantization(self):
        [m for m in self.model.modules() if isinstance(m, LUQuantModule)][-1].disable_act_quant = True
