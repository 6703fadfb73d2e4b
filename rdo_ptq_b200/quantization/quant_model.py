"""B200 drop-in for task-oriented-PTQ/quantization/quant_model.py (:10-98): host-side graph rewrite only."""
import torch.nn as nn

from ..codec.layers import GDN
from ..codec.entropy_models import EntropyBottleneck
from .quant_block import specials, BaseQuantBlock, QuantRBWS, QuantRBU
from .quant_layer import QuantModule, StraightThrough


def link_sequential_consumers(model: nn.Module):
    """Inside an nn.Sequential the output of child i is read by child i+1 only.  Where both are QuantModules (identity
    StraightThrough placeholders in between are skipped) the producer learns its consumer (`_defer_to`), which lets the
    evaluation forward defer the dynamic activation quantiser into the consumer's operand staging (ops.DEFER_ACTQ).
    Host-side wiring only; no effect unless that switch is on."""
    for seq in model.modules():
        if not isinstance(seq, nn.Sequential):
            continue
        kids = [m for m in seq.children() if not isinstance(m, StraightThrough)]
        for a, b in zip(kids, kids[1:]):
            if isinstance(a, QuantModule) and isinstance(b, QuantModule) and not b.is_ps:
                a.__dict__["_defer_to"] = b
    # the same inside the hand-written Cheng2020 blocks: conv -> (I)GDN with nothing else reading the conv's output
    for m in model.modules():
        if isinstance(m, QuantRBWS):
            m.conv2.__dict__["_defer_to"] = m.gdn
        elif isinstance(m, QuantRBU):
            m.conv.__dict__["_defer_to"] = m.igdn


class QuantModel(nn.Module):
    def __init__(self, model: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {}, is_fusing=True,
                 is_cheng=False):
        super().__init__()
        # BN folding (reference fold_bn.py) is a no-op for the LIC graphs: they contain no BatchNorm.
        self.model = model
        self.quant_module_refactor(self.model, weight_quant_params, act_quant_params, is_cheng)
        link_sequential_consumers(self.model)

    def quant_module_refactor(self, module: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {},
                              is_cheng=False):
        """Recursive swap with the reference's rules and order (quant_model.py:35-62)."""
        prev_quantmodule = None
        for name, child in module.named_children():
            if type(child) in specials:
                setattr(module, name, specials[type(child)](child, weight_quant_params, act_quant_params))
            elif isinstance(child, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, nn.LayerNorm, GDN, nn.PixelShuffle)):
                prev_quantmodule = QuantModule(child, weight_quant_params, act_quant_params)
                setattr(module, name, prev_quantmodule)
            elif isinstance(child, (nn.LeakyReLU, nn.GELU, nn.ReLU, nn.ReLU6)):
                if prev_quantmodule is not None:
                    prev_quantmodule.activation_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, StraightThrough):
                continue
            else:
                self.quant_module_refactor(child, weight_quant_params, act_quant_params, is_cheng)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        for m in self.model.modules():
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                m.set_quant_state(weight_quant, act_quant)

    def forward(self, input):
        return self.model(input)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def _quant_modules(self):
        return [m for m in self.model.modules() if isinstance(m, QuantModule)]

    def set_first_last_layer_to_8bit(self):
        ml = self._quant_modules()
        ml[0].weight_quantizer.bitwidth_refactor(8)
        ml[0].act_quantizer.bitwidth_refactor(8)
        ml[-1].weight_quantizer.bitwidth_refactor(8)
        ml[-2].act_quantizer.bitwidth_refactor(8)

    def disable_network_output_quantization(self):
        self._quant_modules()[-1].disable_act_quant = True
