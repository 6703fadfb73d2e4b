"""B200 drop-in for task-oriented-PTQ/quantization/layer_opt.py: `layer_reconstruction` with the reference's
signature and bookkeeping (:175-319); the 20 000-iteration loop runs in `recon.run_reconstruction`.
"""
import logging
import time

import torch

from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel
from .quantizer import StraightThrough
from .recon import CoderTask, DrawPlan, RDTask, UnitTrainer, run_reconstruction
from .utils import save_inp_oup_data, set_mode


def find_unquantized_module(model: torch.nn.Module, _name_: str = "g_a", module_list: list = [], name_list: list = []):
    """reference :15-43.  Collects later untrained units whose *child name* contains the coder tag; for compressai
    Sequential children ("0", "1", ...) this never matches, so the result is empty."""
    for name, module in model.named_children():
        if isinstance(module, (QuantModule, BaseQuantBlock)):
            if not module.trained:
                module.set_quant_state(False, False)
                for tag in ("g_a", "h_a", "h_s", "g_s"):
                    if tag in _name_ and tag in name:
                        name_list.append(name)
                        module_list.append(module)
        else:
            find_unquantized_module(module, _name_, module_list, name_list)
    return module_list[1:], name_list[1:]


def _task_p(args, default=2.0):
    v = getattr(args, "task_loss", default) if args is not None else default
    return default if v == "rd" else float(v)


def _rd_task(model, unit_path, cali_data, args, task, lmbda, fp_unit_out=None):
    """`task='rd'` (or args.task_loss == 'rd'): the R + lambda*D task criterion (recon.RDTask); lambda from `lmbda` or
    args.lmbda (main2.py's --lmbda).  `task='coder'`: the reference's fp_out tail by the unit's real position
    (recon.CoderTask)."""
    if task is None and args is not None and getattr(args, "task_loss", None) == "rd":
        task = "rd"
    if task in ("rd", "coder") and unit_path is not None:
        coder, _, rest = unit_path.partition(".")
        if coder not in ("g_a", "h_a", "h_s", "g_s") or not rest.isdigit():
            # a unit nested inside a composite child (e.g. the convolutions of Cheng2020's attention blocks): the codec's
            # forward cannot be continued from its output by position
            raise NotImplementedError(f"task={task!r}: {unit_path!r} is not a direct child of g_a / h_a / h_s / g_s")
    if task == "coder":
        if unit_path is None:
            raise ValueError("task='coder' needs unit_path (the unit's path inside the codec, e.g. 'g_a.2')")
        return CoderTask(model, unit_path, fp_unit_out, _task_p(args))
    if task != "rd":
        return None
    if unit_path is None:
        raise ValueError("task='rd' needs unit_path (the unit's path inside the codec, e.g. 'g_a.2')")
    lm = lmbda if lmbda is not None else getattr(args, "lmbda", None)
    if lm is None:
        raise ValueError("task='rd' needs lmbda (or args.lmbda)")
    return RDTask(model, unit_path, cali_data, lm)


def layer_reconstruction(model: QuantModel, layer: QuantModule, layer_name: str, cali_data: torch.Tensor,
                         batch_size: int = 32, iters: int = 20000, weight: float = 0.001, opt_mode: str = 'mse',
                         asym: bool = False, include_act_func: bool = True, b_range: tuple = (20, 2),
                         warmup: float = 0.0, input_prob: float = 1.0, act_quant: bool = False, lr: float = 4e-5,
                         p: float = 2.0, config=None, args=None, plan: DrawPlan = None, unit_id: int = 0, trace=None,
                         log_every: int = 500, graph: bool = True, process_group=None, task: str = None,
                         lmbda: float = None, unit_path: str = None, learn_delta: bool = False):
    """Same arguments as the reference; `plan` / `unit_id` / `trace` are additive (deterministic replays, tests), and
    so is `task='rd'` with `lmbda` / `unit_path`: the R + lambda*D task criterion instead of the reference's live
    lp_loss(quant_out, fp_out, args.task_loss).
    As in the reference, `lr` is accepted and ignored: Adam runs at its default 1e-3 (layer_opt.py:254)."""
    if opt_mode != 'mse':
        raise NotImplementedError("only opt_mode='mse' is reachable in the reference (main2.py:225)")
    t0 = time.time()
    cached_inps, cached_outs = save_inp_oup_data(model, layer, cali_data, asym, act_quant, batch_size=1,
                                                 input_prob=True)
    logging.info('Cached init time: {}'.format(time.time() - t0))
    module_list, name_list = find_unquantized_module(model, layer_name, [], [])
    logging.info(name_list)
    if module_list:
        raise NotImplementedError("fp_out tail over later modules only triggers for Lu2022-style names (out of scope)")
    model.set_quant_state(False, False)
    set_mode(model, act_quant)
    if "7" in layer_name:                       # reference :227-235 (`"g_s" and "7" in name`)
        logging.info("=======last layer, close activation quantization=======")
        layer.set_quant_state(True, False)
    else:
        layer.set_quant_state(True, act_quant)
    org_act_func = None
    if not include_act_func:
        org_act_func, layer.activation_function = layer.activation_function, StraightThrough()
    if layer.org_weight is None:                # PixelShuffle wrapper: nothing to learn (reference :245-246)
        return None
    rd = _rd_task(model, unit_path, cali_data, args, task, lmbda, cached_outs)
    trainer = UnitTrainer(layer, iters, weight, b_range, warmup, p, _task_p(args), process_group=process_group,
                          rd_task=rd, learn_delta=learn_delta)
    losses = run_reconstruction(trainer, cached_inps, cached_outs, batch_size, input_prob, unit_id, plan, trace=trace,
                                log_every=log_every, graph=graph)
    if org_act_func is not None:
        layer.activation_function = org_act_func
    return losses
