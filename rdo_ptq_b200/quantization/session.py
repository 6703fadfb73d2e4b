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


def unit_macs(u, cache):
    """Algorithmic MACs per sample of a unit's forward: conv Ho*Wo*Cout*Cin*k*k, transposed conv
    Hin*Win*Cin*Cout*k*k, GDN H*W*C^2; blocks: sum over their QuantModules is approximated by the first rule on the
    block's input/output (only used to balance streams)."""
    q_in, _, fp_out = cache
    w = getattr(u, "weight", None)
    if getattr(u, "is_gdn", False):
        C_, H, W = q_in.shape[1:]
        return H * W * C_ * C_
    if w is None or w.dim() != 4:
        return float(sum(m.weight.numel() for m in u.modules() if getattr(m, "weight", None) is not None)) * \
            fp_out.shape[2] * fp_out.shape[3]
    if getattr(u, "if_tconv", False):
        Cin, H, W = q_in.shape[1:]
        return H * W * Cin * w.shape[1] * w.shape[2] * w.shape[3]
    Cout, Ho, Wo = fp_out.shape[1:]
    return Ho * Wo * Cout * w.shape[1] * w.shape[2] * w.shape[3]


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


class _FixedWeight(torch.nn.Module):
    """Stand-in weight quantiser that returns a precomputed (nearest-rounded) weight: used by the streaming forwards."""

    def __init__(self, w, iw=None):
        super().__init__()
        self.w, self.iw = w, iw
        self.fixed_weight = (w, iw)       # QuantModule._prepared packs it once (quant_layer.py)

    def forward(self, _x):
        return self.w

    def int_weights(self, _x):
        return self.iw


class CalibrationSession:
    """All units' AdaRound problems side by side; `sweep()` = one fused iteration on every unit."""

    def __init__(self, qnn: QuantModel, cali: torch.Tensor, batch_size: int = 8, iters: int = 20000,
                 weight: float = 0.01, b_range=(20, 2), warmup: float = 0.2, input_prob: float = 0.5, p: float = 2.0,
                 task_p: float = 2.0, host_caches: bool = False, seed: int = 1005, graph: bool = True,
                 graph_warmup: int = 2, lr: float = 1e-3, process_group=None, overlap_update: bool = True,
                 n_streams: int = 3, learn_delta: bool = False):
        self.qnn, self.batch_size, self.input_prob, self.seed = qnn, batch_size, input_prob, seed
        self.units = reconstruction_units(qnn)
        dev = next(qnn.parameters()).device
        self.dev = dev
        # host_caches: False = caches in HBM; True = caches in pinned host memory, batch rows copied every iteration;
        # "stream" = nothing cached: the calibration IMAGES stay in pinned host memory, each iteration copies its batch
        # of images (6 MB for 8 patches) and recomputes every unit's (quant_in, fp_in, fp_out) on the device with two
        # captured forwards (all-FP, and all weights nearest-rounded: the same definition the cached modes use).
        self.stream = (host_caches == "stream")
        self.host = bool(host_caches)
        if self.stream:
            self._cali_host = cali.detach().cpu().contiguous().pin_memory()
            self._img = torch.empty((batch_size,) + tuple(cali.shape[1:]), device=dev)
            self._img.copy_(self._cali_host[:batch_size])
            self.caches = self._stream_setup()
        else:
            self.caches = cache_all_units(qnn, cali, self.units, batch=batch_size, to_host=bool(host_caches))
        qnn.set_quant_state(False, False)
        self.trainers = {}
        for n, u in self.units:
            u.set_quant_state(True, False)
            self.trainers[n] = UnitTrainer(u, iters, weight, b_range, warmup, p, task_p, lr=lr,
                                           process_group=process_group, learn_delta=learn_delta)
        # all units' (rec, task, round) accumulators live in one [U, 3] buffer: one device->host read per report
        self._loss_all = torch.zeros(len(self.units), 3, device=dev)
        self._loss_host = None                           # pinned double buffer of losses(lag=True)
        for i, (n, _) in enumerate(self.units):
            self.trainers[n].loss_buf = self._loss_all[i]
        self.n_samples = cali.size(0)
        self.it = 0
        g = torch.Generator().manual_seed(seed)
        # pre-drawn batch picks (randperm rows), resident on the device: no host RNG inside the timed loop
        # (row k = (step - 1) * units + unit; PERM_ROWS rows, or one per (step, unit) of the whole run when that is more,
        # capped at 2^19 rows, so the picks of a full-length calibration do not repeat)
        rows = min(max(PERM_ROWS, iters * len(self.units)), 1 << 19)
        if rows <= PERM_ROWS:
            self._perm = torch.stack([torch.randperm(self.n_samples, generator=g)[:batch_size] for _ in range(rows)])
        else:
            self._perm = torch.cat([torch.rand(min(1 << 15, rows - r0), self.n_samples, generator=g).argsort(dim=1)[:, :batch_size]
                                    for r0 in range(0, rows, 1 << 15)]).contiguous()
        self._perm_rows = rows
        self._perm_dev = self._perm.to(dev)
        self.seed_base = (seed * 2654435761) & 0xFFFFFFFFFFFF
        # device-resident schedule shared by all units (they advance in lock step)
        self.sched = ops.new_sched(dev)
        self._tick = (iters, warmup, b_range[0], b_range[1], lr)
        self.use_graph, self.graph_warmup = graph, graph_warmup
        self._graphs, self._graph_launches, self._captured_launches = {}, {}, 0
        self._tail_graphs, self._tail_launches, self._grads = {}, {}, {}
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        # multi-GPU: the all-reduce + Adam tail of each unit runs on ONE side stream (same collective order on every
        # rank) under the compute of the following units
        self._side = torch.cuda.Stream(device=dev) if (self.world > 1 and graph and overlap_update) else None
        # Sequential multi-GPU sweep (one stream, tails in line): the peer-memory tail of unit k needs no exit barrier --
        # the entry barrier of unit k+1's tail, which every rank reaches before unit k runs again, is a system-scope
        # release / acquire over all ranks that covers unit k's alpha stores (b200lic_xgpu_reduce_adam_sched).  One
        # cross-GPU barrier per unit instead of two.
        self._peer_deferred = (self.world > 1 and self._side is None and max(1, n_streams if graph else 1) == 1 and
                               len(self.units) >= 2 and all(t.peer is not None for t in self.trainers.values()))
        if self._peer_deferred:
            for t in self.trainers.values():
                t.peer_exit_barrier = False
        # Units are independent problems, and most of them (hyperprior layers, 32x32 / 16x16 stages) launch
        # 1-64 CTAs: their graphs are spread over `n_streams` streams (longest-processing-time-first by a cost estimate)
        # so the small units fill the SMs the large ones leave idle.  Each stream owns its graph memory pool.
        self.n_streams = max(1, n_streams if graph else 1)
        self._streams = [None] + [torch.cuda.Stream(device=dev) for _ in range(self.n_streams - 1)]   # None = caller's
        self._pools = [None] * self.n_streams
        cost = {n: 120.0 + 6.0 * unit_macs(u, self.caches[n]) * batch_size / 8e8 for n, u in self.units}   # ~us
        load = [0.0] * self.n_streams
        self._sid = {}
        for n in sorted(cost, key=cost.get, reverse=True):
            k = min(range(self.n_streams), key=load.__getitem__)
            self._sid[n] = k
            load[k] += cost[n]
        # Multi-GPU tails (all-reduce + Adam) share ONE side stream, which runs them in issue order: issued in model
        # order, the tail of a unit queued behind the big layers of its stream held back every later tail until the end
        # of the sweep (N=2: +0.44 ms per step).  They are issued in the order the units are expected to FINISH instead
        # (same cost model on every rank, so NCCL sees the same collective order everywhere).
        acc = [0.0] * self.n_streams
        finish = {}
        for n, _ in self.units:
            acc[self._sid[n]] += cost[n]
            finish[n] = acc[self._sid[n]]
        self._tail_order = sorted((n for n, _ in self.units), key=finish.get)
        self._unit_done = {n: torch.cuda.Event() for n, _ in self.units}
        self._since = 0
        self.h2d_bytes = 0
        self.replayed_launches = 0           # kernels launched by graph replays (not seen by the library's host counter)
        if self.stream:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = self.caches                       # the hooked tensors of the streaming forwards (batch rows)
            self._img_ready, self._img_used, self._fwd_graph, self._fwd_launches = torch.cuda.Event(), None, None, 0
        elif self.host:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage, self._ready, self._consumed = {}, {}, {}
            for n, _ in self.units:
                q_in, fp_in, fp_out = self.caches[n]
                self._stage[n] = tuple(torch.empty((batch_size,) + tuple(t.shape[1:]), device=dev) for t in
                                       (q_in, fp_in, fp_out))
                self._ready[n] = torch.cuda.Event()
                self._consumed[n] = None

    # -- streaming mode: recompute the units' inputs / targets from the batch of images ------------------------------
    def _stream_setup(self):
        """Scale initialisation (main2.py:194-198), nearest-rounded weights frozen per QuantModule, and one eager
        streaming forward so that self._stage_bufs exist (shapes, cost estimates)."""
        qnn = self.qnn
        qnn.eval()
        qnn.set_quant_state(True, False)
        with torch.no_grad():
            qnn(self._img)                                   # first forward initialises every weight quantiser
        self._fixed_w = {}
        self._fwd_stream = torch.cuda.Stream(device=self.dev)
        for m in qnn.modules():
            if isinstance(m, QuantModule) and m.org_weight is not None:
                with torch.no_grad():
                    q = m.weight_quantizer
                    iw = None if m.is_gdn else getattr(q, "int_weights", lambda _w: None)(m.weight)
                    # ONE stand-in object per module for the whole session: QuantModule caches the packed operand of a
                    # constant weight per quantiser object, so the captured forwards reuse it (no per-step packing)
                    self._fixed_w[m] = _FixedWeight(q(m.weight).detach().clone(), iw)
        return self._stream_forward()

    @torch.no_grad()
    def _stream_forward(self):
        """All-FP pass (fp_in, fp_out of every unit) + pass with every weight nearest-rounded (quant_in) on self._img.
        Returns {unit: (quant_in, fp_in, fp_out)}; tensors of block units are cloned (blocks may run in-place ops)."""
        qnn, store = self.qnn, {n: [None, None, None] for n, _ in self.units}

        def run(slot_in, slot_out):
            hooks = []
            for n, m in self.units:
                clone = isinstance(m, BaseQuantBlock)

                def hook(_m, inp, out, n=n, clone=clone):
                    # the first unit's input IS the image buffer, which the copy stream refills for the next step as soon
                    # as this forward has run (self._img_used) -- possibly while that unit's iteration still reads it:
                    # the unit gets its own copy (6 MB device-to-device at the benchmark's size)
                    own = clone or inp[0].data_ptr() == self._img.data_ptr()
                    store[n][slot_in] = inp[0].detach().clone() if own else inp[0].detach()
                    if slot_out is not None:
                        store[n][slot_out] = out.detach().clone() if clone else out.detach()
                hooks.append(m.register_forward_hook(hook))
            qnn(self._img)
            for h in hooks:
                h.remove()

        states = [(m, m.use_weight_quant, m.use_act_quant) for m in qnn.modules()
                  if isinstance(m, (QuantModule, BaseQuantBlock))]
        # the two passes are independent: the quantised one runs on a forked stream (also inside graph capture), so the
        # many 1-64 CTA layers of one pass fill the SMs the other leaves idle
        cur = torch.cuda.current_stream()
        fork = self._fwd_stream
        fork.wait_stream(cur)
        qnn.set_quant_state(False, False)
        run(1, 2)
        swapped = []
        for m, fq in self._fixed_w.items():                 # nearest-rounded weights, whatever quantiser is installed
            swapped.append((m, m.weight_quantizer))
            m.weight_quantizer = fq
        qnn.set_quant_state(True, False)
        with torch.cuda.stream(fork):
            run(0, None)
        cur.wait_stream(fork)
        for m, q in swapped:
            m.weight_quantizer = q
        for m, w_, a_ in states:
            m.use_weight_quant, m.use_act_quant = w_, a_
        return {n: tuple(v) for n, v in store.items()}

    def _stream_step(self, main):
        """Copy this iteration's batch of images host->device (copy stream) and replay the two streaming forwards."""
        idx = self._perm[self.it % self._perm_rows].tolist()
        with torch.cuda.stream(self._copy_stream):
            if self._img_used is not None:
                self._copy_stream.wait_event(self._img_used)
            for b, i in enumerate(idx):
                self._img[b].copy_(self._cali_host[i], non_blocking=True)
            self.h2d_bytes += 4 * self._img.numel()
            self._img_ready.record(self._copy_stream)
        main.wait_event(self._img_ready)
        if self.use_graph and self.it >= self.graph_warmup:
            if self._fwd_graph is None:
                torch.cuda.synchronize()
                n0 = _lib.launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._stage = self._stream_forward()
                self._fwd_graph = g
                self._fwd_launches = _lib.launch_count() - n0
                self._captured_launches += self._fwd_launches
                self._graphs.clear()                         # unit graphs must read the captured forward's buffers
                self._tail_graphs.clear()
            self._fwd_graph.replay()
            self.replayed_launches += self._fwd_launches
        else:
            self._stage = self._stream_forward()
        ev = torch.cuda.Event()
        ev.record(main)
        self._img_used = ev

    # -- one unit, one iteration: the capturable bodies ---------------------------------------------------------------
    def _compute(self, j, name):
        """pick + QDrop -> soft weights -> forward -> loss -> wgrad; leaves the unit's dL/dWq in self._grads[name]."""
        U, bs = len(self.units), self.batch_size
        t = self.trainers[name]
        src = self._stage[name] if self.host else self.caches[name]
        fused = t.fused_plan(src[0].shape[1:], bs)
        if fused is not None:
            # prepared-operand iteration: pick + QDrop -> staged operand, quantiser -> packed weight, GEMM, loss -> staged
            # dY, wgrad (+ Adam tail at world size 1); host / streaming modes hold the batch rows already (identity pick)
            table = None if self.host else self._perm_dev
            self._grads[name] = t.fused_compute(fused, src[0], src[1], src[2], table, U, j, self.sched, self.input_prob,
                                                self.seed_base)
            return
        if self.host:
            qi, fi, tgt = self._stage[name]
            cur = ops.gather_mix_sched(qi, fi, None, bs, self.input_prob, self.seed_base, U, j, self.sched)
        else:
            q_in, fp_in, fp_out = self.caches[name]
            cur = ops.gather_mix_sched(q_in, fp_in, self._perm_dev, bs, self.input_prob, self.seed_base, U, j,
                                       self.sched)
            t = self.trainers[name]
            if t.task_p is not None and float(t.task_p) == float(t.p):
                # the loss kernel reads the target rows straight from the cache (same pick as the input rows)
                _, self._grads[name] = t.step_compute(cur, fp_out, tgt_pick=(self._perm_dev, U, j, self.sched))
                return
            tgt = ops.gather_mix_sched(fp_out, fp_out, self._perm_dev, bs, 1.0, self.seed_base, U, j, self.sched)
        _, self._grads[name] = self.trainers[name].step_compute(cur, tgt)

    def _update(self, j, name):
        """[all-reduce] -> STE / regulariser / Adam for the unit (reads self._grads[name])."""
        self.trainers[name].step_update(self._grads[name], sched=self.sched)

    def _body(self, j, name):
        self._compute(j, name)
        self._update(j, name)

    def _upload(self, j, name):
        """Host-cache mode: copy this iteration's batch rows of (quant_in, fp_in, fp_out) from pinned host memory into
        the unit's staging buffers on the copy stream (one cudaMemcpyAsync per row; no host-side gather)."""
        k = self.it * len(self.units) + j
        idx = self._perm[k % self._perm_rows].tolist()
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[name] is not None:
                self._copy_stream.wait_event(self._consumed[name])      # previous sweep's graph has read the staging
            for src, dst in zip(self.caches[name], self._stage[name]):
                for b, i in enumerate(idx):
                    dst[b].copy_(src[i], non_blocking=True)
                self.h2d_bytes += 4 * dst.numel()
            self._ready[name].record(self._copy_stream)

    def _capture(self, j, name):
        """Capture the unit's iteration on its stream.  Single GPU: one graph.  Multi-GPU: a compute graph and an update
        graph (side stream) so the all-reduce + Adam tail overlaps the compute of the following units."""
        n0 = _lib.launch_count()
        sid = self._sid[name]
        kw = dict(pool=self._pools[sid]) if self._pools[sid] is not None else {}
        if self._streams[sid] is not None:
            kw["stream"] = self._streams[sid]
        if self.world > 1:
            kw["capture_error_mode"] = "thread_local"      # NCCL's watchdog thread polls events during capture
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, **kw):
            if self._side is None:
                self._body(j, name)
            else:
                self._compute(j, name)
        if self._pools[sid] is None:
            self._pools[sid] = g.pool()
        n1 = _lib.launch_count()
        if self._side is not None:
            gt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gt, stream=self._side, capture_error_mode="thread_local"):   # no allocations: own pool
                self._update(j, name)
            self._tail_graphs[name] = gt
            self._tail_launches[name] = _lib.launch_count() - n1
        self._graphs[name] = g
        self._graph_launches[name] = n1 - n0
        self._captured_launches += _lib.launch_count() - n0         # counted by the library while capturing, not executed

    def sweep(self, only=None):
        """One AdaRound iteration (fwd + loss + bwd(alpha) + Adam [+ all-reduce]) on every unit."""
        if only is not None and self._peer_deferred:
            raise ValueError("sweep(only=...) with the deferred peer barrier: the tails of ALL units order each other")
        main = torch.cuda.current_stream()
        ops.sched_tick(self.sched, *self._tick)
        graphed = self.use_graph and self.it >= self.graph_warmup
        if self.stream:
            self._stream_step(main)
        if graphed:
            for st in self._streams[1:]:
                st.wait_stream(main)               # the schedule has been advanced
        used = set()
        for j, (n, _) in enumerate(self.units):
            if only is not None and n not in only:
                continue
            st = (self._streams[self._sid[n]] if graphed else None) or main
            if self.host and not self.stream:
                self._upload(j, n)
                st.wait_event(self._ready[n])
            if graphed:
                if n not in self._graphs:
                    torch.cuda.synchronize()       # capture starts from a quiet device (other streams are mid-sweep)
                    self._capture(j, n)
                with torch.cuda.stream(st):
                    self._graphs[n].replay()
                used.add(st)
                self.replayed_launches += self._graph_launches[n]
                if self._side is not None:
                    self._unit_done[n].record(st)
            else:
                self._body(j, n)
            if self.host and not self.stream:
                ev = torch.cuda.Event()
                ev.record(st)
                self._consumed[n] = ev
        if self._side is not None and graphed:     # tails in expected completion order (see __init__)
            for n in self._tail_order:
                if (only is not None and n not in only) or n not in self._tail_graphs:
                    continue
                self._side.wait_event(self._unit_done[n])
                with torch.cuda.stream(self._side):
                    self._tail_graphs[n].replay()
                self.replayed_launches += self._tail_launches[n]
        for st in used:                            # join: the sweep is complete on the caller's stream
            if st is not main:
                main.wait_stream(st)
        if self._side is not None and graphed:
            main.wait_stream(self._side)
        self.it += 1
        self._since += 1

    def launch_total(self):
        """Library kernel launches so far, including those inside graph replays."""
        return _lib.launch_count() + self.replayed_launches - self._captured_launches

    def losses(self, lag: bool = False):
        """Per-unit mean (rec, task, round, total) since the last call: ONE device->host copy of the [U, 3] buffer.

        `lag=True`: the copy goes to pinned host memory without blocking and the call returns the values of the PREVIOUS
        call's window (empty dict the first time), so a loop that logs every step never drains the device queue: the
        host enqueues step k+1 while step k runs (a blocking read left the GPU idle for the ~0.3 ms of launch work at
        the start of every step)."""
        if not lag:
            vals = (self._loss_all / max(self._since, 1)).tolist()
            self._loss_all.zero_()
            self._since = 0
            return self._loss_dict(vals)
        if self._loss_host is None:
            self._loss_host = [torch.empty(self._loss_all.shape, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_ev, self._loss_k = [None, None], 0
        k = self._loss_k
        snap = self._loss_all / max(self._since, 1)
        self._loss_host[k % 2].copy_(snap, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._loss_ev[k % 2] = ev
        self._loss_all.zero_()
        self._since = 0
        self._loss_k = k + 1
        if k == 0:
            return {}
        self._loss_ev[(k - 1) % 2].synchronize()
        return self._loss_dict(self._loss_host[(k - 1) % 2].tolist())

    def _loss_dict(self, vals):
        out = {}
        for (n, _), (rec, task, rnd) in zip(self.units, vals):
            t = self.trainers[n]
            if getattr(t, "_same", False):
                task = rec
            t.last = out[n] = dict(rec=rec, task=task, round=rnd, total=rec + task + rnd)
        return out

    def finish(self):
        for t in self.trainers.values():
            t.finish()
        self.qnn.set_quant_state(True, False)
