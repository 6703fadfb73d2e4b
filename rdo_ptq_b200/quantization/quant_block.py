"""B200 drop-in for the Cheng2020 part of task-oriented-PTQ/quantization/quant_block.py (:77-102, :219-328, :645-657).

The Lu2022 Swin blocks of that file (QuantMlp :330-350, QuantWindowAttention :353-423, QuantSwinTransformerBlock :426-553,
QuantBasicLayer :556-598, QuantRSTB :601-641) are carried FORWARD ONLY: block reconstruction through
them, and the NIC / TinyLIC graphs that use them, are not built.  Residual add, LeakyReLU and the block-level dynamic activation quantiser run as
libb200lic kernels (`add_act`, K8).
"""
import torch
import torch.nn as nn

from .. import ops
from ..codec.layers import ResidualBlockWithStride, ResidualBlockUpsample, ResidualBlock, subpel_conv3x3, Mlp
from ..codec import swin
from .quant_layer import QuantModule
from .quantizer import StraightThrough, UniformAffineQuantizer, ActQuantizer


class BaseQuantBlock(nn.Module):
    def __init__(self, act_quant_params: dict = {}):
        super().__init__()
        self.use_weight_quant = False
        self.use_act_quant = False
        self.trained = False
        self.act_quantizer = UniformAffineQuantizer(act=True, **act_quant_params)
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)

    def _aq(self, t, consumer=None):
        """Block-level dynamic activation quantiser.  `consumer`: the QuantModule that is the ONLY reader of the result --
        in an evaluation forward the quantiser is then deferred into that module's operand staging (ops.DEFER_ACTQ: the
        statistics are taken here, the codes are applied while the consumer stages its operand; same values)."""
        if not (self.use_act_quant and self.trained):
            return t
        if (consumer is not None and ops.DEFER_ACTQ and not torch.is_grad_enabled() and t.dim() == 4
                and 4 * t.numel() >= ops.DEFER_ACTQ_MIN_BYTES and not consumer._forward_hooks
                and not consumer._forward_pre_hooks):
            bits = self.act_quantizer.n_bits if self.act_quantizer.act_bits_follow_n_bits else 8
            t = t.detach()
            t._b200_actq = (ops.act_quant_stats(t), bits)
            return t
        return self.act_quantizer(t, True)

    def _lrelu(self, t):
        return ops.add_act_fn(t, None, ops.ACT_LEAKY_RELU, float(self.leaky_relu.negative_slope))

    def _conv_lrelu(self, conv, x):
        """LeakyReLU(conv(x)): without a gradient the activation rides in the convolution's epilogue."""
        if torch.is_grad_enabled() or conv._forward_hooks:
            return self._lrelu(conv(x))
        return conv(x, act_override=(ops.ACT_LEAKY_RELU, float(self.leaky_relu.negative_slope)))


class QuantRBWS(BaseQuantBlock):
    """ResidualBlockWithStride (reference :219-250)."""

    def __init__(self, basic_block: ResidualBlockWithStride, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.conv1 = QuantModule(basic_block.conv1, weight_quant_params, act_quant_params, disable_act_quant=True)
        self.leaky_relu = basic_block.leaky_relu
        self.conv2 = QuantModule(basic_block.conv2, weight_quant_params, act_quant_params)
        self.gdn = QuantModule(basic_block.gdn, weight_quant_params, act_quant_params)
        self.skip = (QuantModule(basic_block.skip, weight_quant_params, act_quant_params)
                     if basic_block.skip is not None else None)

    def forward(self, x):
        out = self._aq(self._conv_lrelu(self.conv1, x), consumer=self.conv2)
        out = self.gdn(self.conv2(out))
        out = ops.add_act_fn(out, self.skip(x) if self.skip is not None else x)
        return self._aq(out)


class QuantRBU(BaseQuantBlock):
    """ResidualBlockUpsample (reference :253-284)."""

    def __init__(self, basic_block: ResidualBlockUpsample, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.subpel_conv = nn.Sequential(
            QuantModule(basic_block.subpel_conv[0], weight_quant_params, act_quant_params, disable_act_quant=True),
            basic_block.subpel_conv[1])
        self.leaky_relu = basic_block.leaky_relu
        self.conv = QuantModule(basic_block.conv, weight_quant_params, act_quant_params)
        self.igdn = QuantModule(basic_block.igdn, weight_quant_params, act_quant_params)
        self.upsample = nn.Sequential(QuantModule(basic_block.upsample[0], weight_quant_params, act_quant_params),
                                      basic_block.upsample[1])

    def forward(self, x):
        out = self._aq(self._lrelu(self.subpel_conv(x)), consumer=self.conv)
        out = self.igdn(self.conv(out))
        return self._aq(ops.add_act_fn(out, self.upsample(x)))


class QuantRB(BaseQuantBlock):
    """ResidualBlock (reference :286-313)."""

    def __init__(self, basic_block: ResidualBlock, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.conv1 = QuantModule(basic_block.conv1, weight_quant_params, act_quant_params, disable_act_quant=True)
        self.leaky_relu = basic_block.leaky_relu
        self.conv2 = QuantModule(basic_block.conv2, weight_quant_params, act_quant_params, disable_act_quant=True)
        self.skip = (QuantModule(basic_block.skip, weight_quant_params, act_quant_params)
                     if basic_block.skip is not None else None)

    def forward(self, x):
        out = self._aq(self._conv_lrelu(self.conv1, x), consumer=self.conv2)
        out = self._aq(self._conv_lrelu(self.conv2, out))
        out = ops.add_act_fn(out, self.skip(x) if self.skip is not None else x)
        return self._aq(out)


class QuantSC(BaseQuantBlock):
    """subpel_conv3x3 wrapper (reference :315-328).  Unreachable through `specials` in the reference too: the key is
    a function, the lookup is by type(); kept for API parity."""

    def __init__(self, basic_block, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.subpel_conv = nn.Sequential(
            QuantModule(basic_block[0], weight_quant_params, act_quant_params, disable_act_quant=True),
            basic_block[1], nn.LeakyReLU(inplace=True))

    def forward(self, x):
        out = self.subpel_conv[1](self.subpel_conv[0](x))
        return ops.add_act_fn(out, None, ops.ACT_LEAKY_RELU, 0.01)


class QuantMlp(BaseQuantBlock):
    """quant_block.py:330-350: fc1 (its own activation quantiser disabled) -> act -> [A8] -> fc2.  Values only."""

    def __init__(self, basic_block: Mlp, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.fc1 = QuantModule(basic_block.fc1, weight_quant_params, act_quant_params, disable_act_quant=True)
        self.act = basic_block.act
        self.fc2 = QuantModule(basic_block.fc2, weight_quant_params, act_quant_params)

    def forward(self, x):
        x = self.fc1(x)
        x = ops.gelu(x) if isinstance(self.act, nn.GELU) or type(self.act).__name__ == "GELU" else self.act(x)
        if self.use_act_quant and self.trained:
            x = ActQuantizer(x)
        return self.fc2(x)


class QuantWindowAttention(BaseQuantBlock):
    """quant_block.py:353-423: qkv and proj as QuantModules; the dynamic quantiser sits on the attention matrix (per head,
    4-D) and on the attention output (per channel, 3-D) once the block is trained."""

    def __init__(self, basic_block: swin.WindowAttention, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.dim, self.window_size, self.num_heads = basic_block.dim, basic_block.window_size, basic_block.num_heads
        self.scale = basic_block.scale
        self.qkv = QuantModule(basic_block.qkv, weight_quant_params, act_quant_params)
        self.proj = QuantModule(basic_block.proj, weight_quant_params, act_quant_params)
        self.relative_position_bias_table = basic_block.relative_position_bias_table
        self.relative_position_index = basic_block.relative_position_index

    def forward(self, x, mask=None):
        qkv = self.qkv(x)
        bias = swin.gathered_bias(self.relative_position_bias_table, self.relative_position_index, x.shape[1])
        attn = ops.window_attn_softmax(qkv, bias, mask, self.num_heads, self.scale)
        quantise = self.use_act_quant and self.trained
        if quantise:
            attn = ActQuantizer(attn)
        out = ops.window_attn_apply(attn, qkv)
        if quantise:
            out = ActQuantizer(out)
        return self.proj(out)


class QuantSwinTransformerBlock(BaseQuantBlock):
    """quant_block.py:426-553: LayerNorms as QuantModules, quantised attention and Mlp, block output quantised when
    trained."""

    def __init__(self, basic_block: swin.SwinTransformerBlock, weight_quant_params: dict = {},
                 act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.geom = basic_block            # window / shift geometry and the cached mask (no parameters of its own are used)
        self.norm1 = QuantModule(basic_block.norm1, weight_quant_params, act_quant_params)
        self.attn = QuantWindowAttention(basic_block.attn, weight_quant_params, act_quant_params)
        self.norm2 = QuantModule(basic_block.norm2, weight_quant_params, act_quant_params)
        self.mlp = QuantMlp(basic_block.mlp, weight_quant_params, act_quant_params)

    def forward(self, x, x_size):
        y = self.geom.attend(self.attn, self.norm1(x), x_size)
        x = ops.add_act(x, y.contiguous())
        x = ops.add_act(x, self.mlp(self.norm2(x)))
        if self.use_act_quant and self.trained:
            x = ActQuantizer(x)
        return x


class QuantBasicLayer(BaseQuantBlock):
    """quant_block.py:556-598."""

    def __init__(self, basic_block: swin.BasicLayer, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.blocks = nn.ModuleList([QuantSwinTransformerBlock(b, weight_quant_params, act_quant_params)
                                     for b in basic_block.blocks])

    def forward(self, x, x_size):
        for blk in self.blocks:
            x = blk(x, x_size)
        return x


class QuantRSTB(BaseQuantBlock):
    """quant_block.py:601-641: tokens through the quantised BasicLayer, back to a feature map, plus the input; the sum is
    quantised (4-D, per channel) when the block is trained."""

    def __init__(self, basic_block: swin.RSTB, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.dim, self.input_resolution = basic_block.dim, basic_block.input_resolution
        self.residual_group = QuantBasicLayer(basic_block.residual_group, weight_quant_params, act_quant_params)
        self.patch_embed, self.patch_unembed = swin.PatchEmbed(), swin.PatchUnEmbed()

    def forward(self, x, x_size):
        y = self.patch_unembed(self.residual_group(self.patch_embed(x), x_size), x_size)
        out = ops.add_act(y, x)
        if self.use_act_quant and self.trained:
            out = ActQuantizer(out)
        return out


specials = {
    swin.RSTB: QuantRSTB,
    ResidualBlockWithStride: QuantRBWS,
    ResidualBlockUpsample: QuantRBU,
    ResidualBlock: QuantRB,
    subpel_conv3x3: QuantSC,
}
