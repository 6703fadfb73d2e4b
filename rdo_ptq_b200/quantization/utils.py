"""B200 drop-in for the hot-path part of task-oriented-PTQ/quantization/utils.py:
`LinearTempDecay` (:37-54) and the hooked input/output caching `save_inp_oup_data` / `GetLayerInpOut` /
`DataSaverHook` (:92-139, :176-258).  The Fisher-gradient helpers (:142-173, :285-335) are dead code in the
reference (`opt_mode='mse'` is hard-coded, main2.py:225) and are not carried.
"""
from typing import Union

import torch

from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel


class StopForwardException(Exception):
    """Raised by the data-saver hook to cut the forward right after the unit of interest."""


def set_mode(model, act_quant):
    """reference utils.py:28-35: re-enable quantisation on every already-trained QuantModule."""
    for _, module in model.named_children():
        if isinstance(module, QuantModule):
            if module.trained:
                module.set_quant_state(True, act_quant)
        else:
            set_mode(module, act_quant)


class LinearTempDecay:
    def __init__(self, t_max: int, rel_start_decay: float = 0.2, start_b: int = 10, end_b: int = 2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b = start_b
        self.end_b = end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, (1 - rel_t))


class DataSaverHook:
    def __init__(self, store_input=False, store_output=False, stop_forward=False):
        self.store_input, self.store_output, self.stop_forward = store_input, store_output, stop_forward
        self.input_store = None
        self.output_store = None

    def __call__(self, module, input_batch, output_batch):
        if self.store_input:
            self.input_store = input_batch
        if self.store_output:
            self.output_store = output_batch
        if self.stop_forward:
            raise StopForwardException


class GetLayerInpOut:
    """reference utils.py:195-258: one FP pass (input_sym, fp output) + one pass with trained units quantised."""

    def __init__(self, model: QuantModel, layer: Union[QuantModule, BaseQuantBlock], device, asym: bool = False,
                 act_quant: bool = False, input_prob: bool = False):
        self.model, self.layer, self.asym, self.device = model, layer, asym, device
        self.act_quant, self.input_prob = act_quant, input_prob
        self.data_saver = DataSaverHook(store_input=True, store_output=True, stop_forward=True)

    def _run(self, x):
        try:
            self.model(x)
        except StopForwardException:
            pass

    def __call__(self, model_input):
        self.model.eval()
        self.model.set_quant_state(False, False)
        handle = self.layer.register_forward_hook(self.data_saver)
        x = model_input.to(self.device)
        with torch.no_grad():
            self._run(x)
            input_sym = self.data_saver.input_store[0].detach() if self.input_prob else None
            if self.asym:
                self.data_saver.store_output = False
                set_mode(self.model, self.act_quant)
                self._run(x)
            self.data_saver.store_output = True
        handle.remove()
        self.model.set_quant_state(False, False)
        set_mode(self.model, self.act_quant)
        self.layer.set_quant_state(True, self.act_quant)
        self.model.train()
        inp, out = self.data_saver.input_store[0].detach(), self.data_saver.output_store.detach()
        return (inp, out, input_sym) if self.input_prob else (inp, out)


def save_inp_oup_data(model: QuantModel, layer: Union[QuantModule, BaseQuantBlock], cali_data: torch.Tensor,
                      asym: bool = False, act_quant: bool = False, batch_size: int = 32, keep_gpu: bool = True,
                      input_prob: bool = False):
    """reference utils.py:92-139.  The caches stay in HBM (the reference's `.cpu()` round trip is dropped; a
    [12,192,128,128] fp32 cache is 150 MB of 180 GB)."""
    device = next(model.parameters()).device
    get_inp_out = GetLayerInpOut(model, layer, device=device, asym=asym, act_quant=act_quant, input_prob=input_prob)
    rows = [get_inp_out(cali_data[i * batch_size:(i + 1) * batch_size])
            for i in range(int(cali_data.size(0) / batch_size))]
    cached_inps = torch.cat([r[0] for r in rows])
    cached_outs = torch.cat([r[1] for r in rows])
    if input_prob:
        return (cached_inps, torch.cat([r[2] for r in rows])), cached_outs
    return (cached_inps,), cached_outs
