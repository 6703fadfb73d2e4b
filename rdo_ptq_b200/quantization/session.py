"""Device-resident calibration session: caches the (quant_in, fp_in, fp_out) triplets of every
reconstruction unit with two batched hooked forwards instead of the reference's 24 truncated batch-1 forwards per
unit (utils.py:92-139), and steps all units' AdaRound problems from HBM- or host-resident caches.

Used by bench.py (the throughput workload) and available to callers that want the whole-model sweep; the
reference-faithful sequential path stays `layer_reconstruction` / `block_reconstruction`.
"""
from typing import Dict, List

import torch

from .. import ops
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel
from .recon import UnitTrainer


def reconstruction_units(qnn: QuantModel):
    """Units in `recon_model` order (main2.py:227-253): QuantModules and BaseQuantBlocks, depth first, blocks opaque."""
    units = []

    def walk(module, prefix):
        for name, m in module.named_children():
            full = f"{prefix}.{name}" if prefix else name
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                if not m.ignore_reconstruction and not (isinstance(m, QuantModule) and m.org_weight is None):
                    units.append((full, m))
            else:
                walk(m, full)

    walk(qnn.model, "")
    return units


@torch.no_grad()
def cache_all_units(qnn: QuantModel, cali: torch.Tensor, units, batch: int = 8, to_host: bool = False):
    """Returns {name: (quant_in, fp_in, fp_out)}; fp_* from an all-FP pass, quant_in from a pass with every weight
    quantiser on (the state the last unit sees in the sequential procedure)."""
    store: Dict[str, List[List[torch.Tensor]]] = {n: [[], [], []] for n, _ in units}

    def run(slot_in, slot_out):
        hooks = []
        for n, m in units:
            def hook(_m, inp, out, n=n):
                store[n][slot_in].append(inp[0].detach().cpu() if to_host else inp[0].detach())
                if slot_out is not None:
                    store[n][slot_out].append(out.detach().cpu() if to_host else out.detach())
            hooks.append(m.register_forward_hook(hook))
        for i in range(0, cali.size(0), batch):
            qnn(cali[i:i + batch])
        for h in hooks:
            h.remove()

    qnn.eval()
    qnn.set_quant_state(False, False)
    run(1, 2)
    qnn.set_quant_state(True, False)
    run(0, None)
    out = {}
    for n, _ in units:
        q_in, fp_in, fp_out = (torch.cat(v) for v in store[n])
        if to_host:
            q_in, fp_in, fp_out = q_in.pin_memory(), fp_in.pin_memory(), fp_out.pin_memory()
        out[n] = (q_in, fp_in, fp_out)
    return out


class CalibrationSession:
    """All units' AdaRound problems side by side; `sweep()` = one fused iteration on every unit."""

    def __init__(self, qnn: QuantModel, cali: torch.Tensor, batch_size: int = 8, iters: int = 20000,
                 weight: float = 0.01, b_range=(20, 2), warmup: float = 0.2, input_prob: float = 0.5, p: float = 2.0,
                 task_p: float = 2.0, host_caches: bool = False, seed: int = 1005):
        self.qnn, self.batch_size, self.input_prob, self.seed = qnn, batch_size, input_prob, seed
        self.units = reconstruction_units(qnn)
        self.host = host_caches
        self.caches = cache_all_units(qnn, cali, self.units, batch=batch_size, to_host=host_caches)
        qnn.set_quant_state(False, False)
        self.trainers = {}
        for n, u in self.units:
            u.set_quant_state(True, False)
            self.trainers[n] = UnitTrainer(u, iters, weight, b_range, warmup, p, task_p)
        self.n_samples = cali.size(0)
        self.it = 0
        dev = next(qnn.parameters()).device
        self.dev = dev
        g = torch.Generator().manual_seed(seed)
        # pre-drawn batch picks (randperm rows), resident on the device: no host RNG inside the timed loop
        self._perm = torch.stack([torch.randperm(self.n_samples, generator=g)[:batch_size] for _ in range(4096)])
        self._perm_dev = self._perm.to(dev)
        self.h2d_bytes = 0

    def _batch(self, name, k):
        q_in, fp_in, fp_out = self.caches[name]
        row = k % self._perm.size(0)
        seed = (self.seed * 2654435761 + k) & 0xFFFFFFFFFFFF
        if self.host:
            idx = self._perm[row]
            qi = q_in[idx].pin_memory().to(self.dev, non_blocking=True)
            fi = fp_in[idx].pin_memory().to(self.dev, non_blocking=True)
            tgt = fp_out[idx].pin_memory().to(self.dev, non_blocking=True)
            self.h2d_bytes += 4 * (qi.numel() + fi.numel() + tgt.numel())
            cur = ops.gather_mix(qi, fi, None, prob=self.input_prob, seed=seed)
            return cur, tgt
        idx = self._perm_dev[row]
        cur = ops.gather_mix(q_in, fp_in, idx, prob=self.input_prob, seed=seed)
        tgt = ops.gather_mix(fp_out, fp_out, idx, prob=1.0)
        return cur, tgt

    def sweep(self, only=None):
        """One AdaRound iteration (fwd + loss + bwd(alpha) + Adam [+ all-reduce]) on every unit."""
        for j, (n, _) in enumerate(self.units):
            if only is not None and n not in only:
                continue
            cur, tgt = self._batch(n, self.it * len(self.units) + j)
            self.trainers[n].step(cur, tgt)
        self.it += 1

    def losses(self):
        return {n: t.read_losses(max(t.count, 1)) for n, t in self.trainers.items()}

    def finish(self):
        for t in self.trainers.values():
            t.finish()
        self.qnn.set_quant_state(True, False)
