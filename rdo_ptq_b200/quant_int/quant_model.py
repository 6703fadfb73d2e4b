"""B200 drop-in for light-uniform-PTQ/quant_int/{quant_model,quant_coding_model}.py (host-side graph rewrite)."""
import torch.nn as nn

from ..codec.entropy_models import EntropyBottleneck
from .quant_layer import QuantModule, StraightThrough


class QuantModel(nn.Module):
    """reference quant_model.py:9-78.  `forward(input)` is the 1-argument pass-through a Balle2018 model needs; the
    reference's `forward(input, lamda)` (TinyLIC) is kept when `lamda` is given."""

    skip_prefixes = ()

    def __init__(self, model: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {}, is_fusing=False):
        super().__init__()
        self.model = model
        self.quant_module_refactor(self.model, weight_quant_params, act_quant_params, top=True)

    def quant_module_refactor(self, module, weight_quant_params: dict = {}, act_quant_params: dict = {}, top=False):
        prev_quantmodule = None
        for name, child in module.named_children():
            if top and any(name.startswith(p) for p in self.skip_prefixes):
                continue
            if isinstance(child, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, nn.LayerNorm)):
                prev_quantmodule = QuantModule(child, weight_quant_params, act_quant_params)
                setattr(module, name, prev_quantmodule)
            elif isinstance(child, (nn.LeakyReLU, nn.GELU, nn.ReLU, nn.ReLU6)):
                if prev_quantmodule is not None:
                    prev_quantmodule.activation_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, StraightThrough):
                continue
            else:
                self.quant_module_refactor(child, weight_quant_params, act_quant_params)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        for m in self.model.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)

    def forward(self, input, lamda=None):
        return self.model(input) if lamda is None else self.model(input, lamda)

    def compress(self, *input):
        return self.model.compress(*input)

    def decompress(self, input, shape, *params):
        return self.model.decompress(input, shape, *params)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def disable_network_output_quantization(self):
        [m for m in self.model.modules() if isinstance(m, QuantModule)][-1].disable_act_quant = True


class QuantCodingModel(QuantModel):
    """reference quant_coding_model.py:23-26: leaves the g_a*/g_s* transforms in fp32, quantises hyper/entropy nets."""
    skip_prefixes = ("g_a", "g_s")
