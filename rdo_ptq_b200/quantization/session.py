"""Device-resident calibration session: caches the (quant_in, fp_in, fp_out) triplets of every
reconstruction unit with two batched hooked forwards instead of the reference's 24 truncated batch-1 forwards per
unit (utils.py:92-139), and steps all units' AdaRound problems from HBM- or host-resident caches.

Per unit, one iteration (batch pick + QDrop mix -> soft weights -> forward -> loss -> wgrad/dgrad -> [all-reduce] ->
STE/regulariser/Adam) is captured ONCE as a CUDA graph and replayed: everything that changes between iterations
(step counter, Adam bias corrections, temperature b, batch-pick row, QDrop seed) lives in a device-resident
`b200lic_calib_sched` advanced by a one-thread tick kernel, so the replayed kernels' arguments never change.
With `host_caches=True` the caches stay in pinned host memory and every iteration's batch rows are copied host->device
on a copy stream into per-unit staging buffers, overlapped with the previous unit's graph (the `e2e` leg of bench.py).

Used by bench.py (the throughput workload) and available to callers that want the whole-model sweep; the
reference-faithful sequential path stays `layer_reconstruction` / `block_reconstruction`.
"""
from typing import Dict, List

import torch
import torch.distributed as dist

from .. import ops, _lib
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel
from .recon import UnitTrainer

PERM_ROWS = 4096


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
                 task_p: float = 2.0, host_caches: bool = False, seed: int = 1005, graph: bool = True,
                 graph_warmup: int = 2, lr: float = 1e-3, process_group=None):
        self.qnn, self.batch_size, self.input_prob, self.seed = qnn, batch_size, input_prob, seed
        self.units = reconstruction_units(qnn)
        self.host = host_caches
        self.caches = cache_all_units(qnn, cali, self.units, batch=batch_size, to_host=host_caches)
        qnn.set_quant_state(False, False)
        self.trainers = {}
        for n, u in self.units:
            u.set_quant_state(True, False)
            self.trainers[n] = UnitTrainer(u, iters, weight, b_range, warmup, p, task_p, lr=lr,
                                           process_group=process_group)
        self.n_samples = cali.size(0)
        self.it = 0
        dev = next(qnn.parameters()).device
        self.dev = dev
        g = torch.Generator().manual_seed(seed)
        # pre-drawn batch picks (randperm rows), resident on the device: no host RNG inside the timed loop
        self._perm = torch.stack([torch.randperm(self.n_samples, generator=g)[:batch_size] for _ in range(PERM_ROWS)])
        self._perm_dev = self._perm.to(dev)
        self.seed_base = (seed * 2654435761) & 0xFFFFFFFFFFFF
        # device-resident schedule shared by all units (they advance in lock step)
        self.sched = ops.new_sched(dev)
        self._tick = (iters, warmup, b_range[0], b_range[1], lr)
        self.use_graph, self.graph_warmup = graph, graph_warmup
        self._graphs, self._graph_launches, self._pool, self._captured_launches = {}, {}, None, 0
        self._since = 0
        self.h2d_bytes = 0
        self.replayed_launches = 0           # kernels launched by graph replays (not seen by the library's host counter)
        if self.host:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage, self._ready, self._consumed = {}, {}, {}
            for n, _ in self.units:
                q_in, fp_in, fp_out = self.caches[n]
                self._stage[n] = tuple(torch.empty((batch_size,) + tuple(t.shape[1:]), device=dev) for t in
                                       (q_in, fp_in, fp_out))
                self._ready[n] = torch.cuda.Event()
                self._consumed[n] = None

    # -- one unit, one iteration: the capturable body -----------------------------------------------------------------
    def _body(self, j, name):
        U, bs = len(self.units), self.batch_size
        if self.host:
            qi, fi, tgt = self._stage[name]
            cur = ops.gather_mix_sched(qi, fi, None, bs, self.input_prob, self.seed_base, U, j, self.sched)
        else:
            q_in, fp_in, fp_out = self.caches[name]
            cur = ops.gather_mix_sched(q_in, fp_in, self._perm_dev, bs, self.input_prob, self.seed_base, U, j,
                                       self.sched)
            tgt = ops.gather_mix_sched(fp_out, fp_out, self._perm_dev, bs, 1.0, self.seed_base, U, j, self.sched)
        self.trainers[name].step(cur, tgt, sched=self.sched)

    def _upload(self, j, name):
        """Host-cache mode: copy this iteration's batch rows of (quant_in, fp_in, fp_out) from pinned host memory into
        the unit's staging buffers on the copy stream (one cudaMemcpyAsync per row; no host-side gather)."""
        k = self.it * len(self.units) + j
        idx = self._perm[k % PERM_ROWS].tolist()
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[name] is not None:
                self._copy_stream.wait_event(self._consumed[name])      # previous sweep's graph has read the staging
            for src, dst in zip(self.caches[name], self._stage[name]):
                for b, i in enumerate(idx):
                    dst[b].copy_(src[i], non_blocking=True)
                self.h2d_bytes += 4 * dst.numel()
            self._ready[name].record(self._copy_stream)

    def _capture(self, j, name):
        n0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        kw = dict(pool=self._pool) if self._pool is not None else {}
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            kw["capture_error_mode"] = "thread_local"      # NCCL's watchdog thread polls events during capture
        with torch.cuda.graph(g, **kw):
            self._body(j, name)
        if self._pool is None:
            self._pool = g.pool()
        self._graphs[name] = g
        self._graph_launches[name] = _lib.launch_count() - n0
        self._captured_launches += self._graph_launches[name]      # counted by the library while capturing, not executed

    def sweep(self, only=None):
        """One AdaRound iteration (fwd + loss + bwd(alpha) + Adam [+ all-reduce]) on every unit."""
        cur_stream = torch.cuda.current_stream()
        ops.sched_tick(self.sched, *self._tick)
        graphed = self.use_graph and self.it >= self.graph_warmup
        for j, (n, _) in enumerate(self.units):
            if only is not None and n not in only:
                continue
            if self.host:
                self._upload(j, n)
                cur_stream.wait_event(self._ready[n])
            if graphed:
                if n not in self._graphs:
                    self._capture(j, n)
                self._graphs[n].replay()
                self.replayed_launches += self._graph_launches[n]
            else:
                self._body(j, n)
            if self.host:
                ev = torch.cuda.Event()
                ev.record(cur_stream)
                self._consumed[n] = ev
        self.it += 1
        self._since += 1

    def launch_total(self):
        """Library kernel launches so far, including those inside graph replays."""
        return _lib.launch_count() + self.replayed_launches - self._captured_launches

    def losses(self):
        out = {n: t.read_losses(max(self._since, 1)) for n, t in self.trainers.items()}
        self._since = 0
        return out

    def finish(self):
        for t in self.trainers.values():
            t.finish()
        self.qnn.set_quant_state(True, False)
