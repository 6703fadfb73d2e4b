"""B200 drop-in for task-oriented-PTQ/quantization/block_opt.py: `block_reconstruction` (:176-323) for the Cheng2020
residual blocks; all QuantModules of the block are optimised jointly by `recon.run_reconstruction`.
"""
import logging
import time

import torch

from .layer_opt import find_unquantized_module, _task_p, _rd_task
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel
from .quantizer import StraightThrough
from .recon import DrawPlan, UnitTrainer, run_reconstruction
from .utils import save_inp_oup_data


def set_mode(model, act_quant):
    """reference block_opt.py:15-22: here trained *blocks* are switched on too."""
    for _, module in model.named_children():
        if isinstance(module, (QuantModule, BaseQuantBlock)):
            if module.trained:
                module.set_quant_state(True, act_quant)
        else:
            set_mode(module, act_quant)


def block_reconstruction(model: QuantModel, block: BaseQuantBlock, block_name: str, cali_data: torch.Tensor,
                         batch_size: int = 32, iters: int = 20000, weight: float = 0.01, opt_mode: str = 'mse',
                         asym: bool = False, include_act_func: bool = True, b_range: tuple = (20, 2),
                         warmup: float = 0.0, input_prob: float = 1.0, act_quant: bool = False, lr: float = 4e-5,
                         p: float = 2.0, config=None, args=None, plan: DrawPlan = None, unit_id: int = 0, trace=None,
                         log_every: int = 500, graph: bool = True, process_group=None, task: str = None,
                         lmbda: float = None, unit_path: str = None):
    """Same arguments as the reference; `plan` / `unit_id` / `trace` / `task` / `lmbda` / `unit_path` are additive, as in
    `layer_reconstruction`."""
    if opt_mode != 'mse':
        raise NotImplementedError("only opt_mode='mse' is reachable in the reference (main2.py:225)")
    t0 = time.time()
    cached_inps, cached_outs = save_inp_oup_data(model, block, cali_data, asym, act_quant, batch_size=1,
                                                 input_prob=True)
    logging.info('Cached init time: {}'.format(time.time() - t0))
    module_list, name_list = find_unquantized_module(model, block_name, [], [])
    logging.info(name_list)
    if module_list:
        raise NotImplementedError("fp_out tail over later modules only triggers for Lu2022-style names (out of scope)")
    model.set_quant_state(False, False)
    set_mode(model, act_quant)
    block.set_quant_state(True, act_quant)
    org_act_func = None
    if not include_act_func:
        org_act_func, block.activation_function = block.activation_function, StraightThrough()
    rd = _rd_task(model, unit_path, cali_data, args, task, lmbda, cached_outs)
    trainer = UnitTrainer(block, iters, weight, b_range, warmup, p, _task_p(args), process_group=process_group,
                          rd_task=rd)
    losses = run_reconstruction(trainer, cached_inps, cached_outs, batch_size, input_prob, unit_id, plan, trace=trace,
                                log_every=log_every, graph=graph)
    if org_act_func is not None:
        block.activation_function = org_act_func
    return losses
