"""Oracle restatement of the AdaRound layer/block reconstruction loop (test infrastructure, CPU fp32).

Follows /root/reference/task-oriented-PTQ/quantization/{layer_opt.py,block_opt.py,utils.py} and the
`recon_model` walk of main2.py:227-264.  The only deviation: the per-iteration randomness
(`torch.randperm` batch pick and the QDrop mask, layer_opt.py:289-292) is drawn from an explicit
`DrawPlan` so the CUDA path can replay the identical draws.
PINNED: `oracle/make_golden.py::wrap_vectors` runs the reference's own loops on CPU (through `oracle/_ref_shim.py`, which
serves their `torch.randperm` / `torch.rand_like` calls from the same `DrawPlan`) and asserts bit-exact alpha and
hardened weights for every unit of a mbt2018-mean walk and of the first 14 cheng2020-attn units.
"""
import torch
import torch.nn as nn

from .quantizers import AdaRoundQuantizer, LinearTempDecay, lp_loss, round_ste
from .quant_wrap import QuantModule, BaseQuantBlock, QuantModel


class DrawPlan:
    """Deterministic stand-in for layer_opt.py:289-292: idx = randperm(n)[:bs]; mask = rand_like(x) < p."""

    def __init__(self, seed: int = 1005):
        self.seed = seed

    def draw(self, unit: int, it: int, n: int, bs: int, shape, prob: float):
        g = torch.Generator().manual_seed(self.seed * 1000003 + unit * 100003 + it)
        idx = torch.randperm(n, generator=g)[:bs]
        keep = None
        if prob < 1.0:
            keep = torch.rand((bs,) + tuple(shape), generator=g) < prob
        return idx, keep


class StopForward(Exception):
    pass


def set_mode(model, act_quant, blocks=False):
    """layer_opt.py:77-84 (QuantModule only) / block_opt.py:15-22 (also blocks)."""
    kinds = (QuantModule, BaseQuantBlock) if blocks else (QuantModule,)
    for _, m in model.named_children():
        if isinstance(m, kinds):
            if m.trained:
                m.set_quant_state(True, act_quant)
        else:
            set_mode(m, act_quant, blocks)


def get_inp_out(model: QuantModel, unit, x, act_quant, blocks):
    """utils.py:207-258 (asym=True, input_prob=True): FP pass stores (fp_in, fp_out); a second pass with
    the already-trained units quantised stores quant_in."""
    store = {}

    def hook(_m, inp, out):
        store["in"], store["out"] = inp[0].detach(), out.detach()
        raise StopForward

    model.eval()
    model.set_quant_state(False, False)
    h = unit.register_forward_hook(hook)
    with torch.no_grad():
        try:
            model(x)
        except StopForward:
            pass
        fp_in, fp_out = store["in"], store["out"]
        set_mode(model, act_quant, False)           # utils.py:28-35 only ever matches QuantModule
        try:
            model(x)
        except StopForward:
            pass
        q_in = store["in"]
    h.remove()
    model.set_quant_state(False, False)
    set_mode(model, act_quant, False)
    unit.set_quant_state(True, act_quant)
    model.train()
    return q_in, fp_out, fp_in


def save_inp_oup_data(model, unit, cali_data, act_quant, blocks):
    """utils.py:92-139 with batch_size=1 (layer_opt.py:213)."""
    rows = [get_inp_out(model, unit, cali_data[i:i + 1], act_quant, blocks) for i in range(cali_data.size(0))]
    return (torch.cat([r[0] for r in rows]), torch.cat([r[2] for r in rows])), torch.cat([r[1] for r in rows])


class LossFunction:
    """layer_opt.py:87-173 / block_opt.py:87-173 with rec_loss='mse'.  `task_p` is args.task_loss."""

    def __init__(self, unit, weight, max_count, b_range, warmup, p=2.0, task_p=2.0):
        self.unit, self.weight, self.p, self.task_p = unit, weight, p, task_p
        self.loss_start = max_count * warmup
        self.temp_decay = LinearTempDecay(max_count, rel_start_decay=warmup, start_b=b_range[0], end_b=b_range[1])
        self.count = 0
        self.last = {}

    def round_modules(self):
        if isinstance(self.unit, QuantModule):
            return [self.unit]
        return [m for _, m in self.unit.named_modules() if isinstance(m, QuantModule)]

    def __call__(self, pred, tgt, quant_net_out, fp_net_out):
        self.count += 1
        rec = lp_loss(pred, tgt, p=self.p)
        task = lp_loss(quant_net_out, fp_net_out, p=self.task_p) if quant_net_out is not None else 0.0
        b = self.temp_decay(self.count)
        if self.count < self.loss_start:
            b = rnd = 0
        else:
            rnd = 0
            for m in self.round_modules():
                h = m.weight_quantizer.get_soft_targets()
                rnd = rnd + self.weight * (1 - ((h - .5).abs() * 2).pow(b)).sum()
        self.last = dict(rec=float(rec.detach()), task=float(task.detach()) if torch.is_tensor(task) else task,
                         round=float(rnd.detach()) if torch.is_tensor(rnd) else rnd, b=b)
        return rnd + rec + task


class RDTask:
    """The R + lambda*D task criterion the reference keeps commented out (layer_opt.py:146-148:
    `criterion = RateDistortionLoss(...); task_loss = criterion(quant_net_out, cali_data)['loss']`), made runnable:
    quant_net_out = the codec's forward continued from the unit's output through the not-yet-trained modules (the role
    of fp_out, layer_opt.py:45-75, whose round_ste on y becomes straight-through rounding of both latents), and
    losses.py:15-28 (MSE metric) on it.  `unit_path` is the unit's path inside the codec, e.g. "g_a.2"."""

    def __init__(self, qnn: QuantModel, unit_path: str, cali_data, lmbda: float):
        from .evalpath import rate_distortion_loss
        self.rd = rate_distortion_loss
        codec = qnn.model
        coder, _, rest = unit_path.partition(".")
        k = int(rest.split(".")[0])
        self.codec, self.coder, self.lmbda = codec, coder, lmbda
        self.tail = list(getattr(codec, coder).children())[k + 1:]
        codec.entropy_bottleneck.ste_round = codec.gaussian_conditional.ste_round = True
        with torch.no_grad():
            y, z = codec.latents(cali_data)
        self.ctx = {"x": cali_data, "y": y, "z": z}

    def __call__(self, out, idx):
        v = out
        for m in self.tail:
            v = m(v)
        o = self.codec.forward_from(self.coder, v, {k: t[idx] for k, t in self.ctx.items()})
        return self.rd(o, self.ctx["x"][idx], self.lmbda)["loss"]

    def close(self):
        self.codec.entropy_bottleneck.ste_round = self.codec.gaussian_conditional.ste_round = False


class CoderTask:
    """layer_opt.py:45-75 (fp_out), :219-224 (fp_net_out) and :296-299 (quant_net_out) with `module_list` = the later,
    untrained modules of the unit's own sub-network taken by position (what `find_unquantized_module` intends; for
    compressai child names its string test never matches)."""

    def __init__(self, qnn: QuantModel, unit_path: str, fp_unit_out, p=2.0):
        coder, _, rest = unit_path.partition(".")
        k = int(rest.split(".")[0])
        self.tail = list(getattr(qnn.model, coder).children())[k + 1:]
        self.round, self.p = coder == "g_a", p
        with torch.no_grad():
            self.target = self._run(fp_unit_out)

    def _run(self, v):
        for m in self.tail:
            v = m(v)
        return round_ste(v) if self.round else v         # layer_opt.py:67-70

    def __call__(self, out, idx):
        return lp_loss(self._run(out), self.target[idx], p=self.p)

    def close(self):
        pass


def reconstruct(model: QuantModel, unit, unit_id: int, unit_name: str, cali_data, batch_size=4, iters=20000,
                weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5, act_quant=False, p=2.0, task_p=2.0,
                plan: DrawPlan = None, trace=None, rd_task: RDTask = None, learn_delta=False,
                delta_lr_scale=0.1):
    """layer_reconstruction (layer_opt.py:175-319) / block_reconstruction (block_opt.py:176-323).
    For compressai-style models `find_unquantized_module` returns [] so fp_out is the identity
    and the task term equals lp_loss(out_quant, fp_out, task_p)."""
    plan = plan or DrawPlan()
    blocks = isinstance(unit, BaseQuantBlock)
    (q_in, fp_in), fp_out = save_inp_oup_data(model, unit, cali_data, act_quant, blocks)
    model.set_quant_state(False, False)
    set_mode(model, act_quant, blocks)
    if not blocks and "7" in unit_name:                 # layer_opt.py:227-235 (Q2)
        unit.set_quant_state(True, False)
    else:
        unit.set_quant_state(True, act_quant)
    mods = [unit] if not blocks else [m for _, m in unit.named_modules() if isinstance(m, QuantModule)]
    if not blocks and unit.org_weight is None:          # PixelShuffle wrapper (layer_opt.py:245-246)
        return None
    for m in mods:
        if m.org_weight is None:
            continue
        m.weight_quantizer = AdaRoundQuantizer(m.weight_quantizer, m.org_weight.data)
        m.weight_quantizer.soft_targets = True
    params = [m.weight_quantizer.alpha for m in mods if m.org_weight is not None]
    if trace is not None:
        trace["h0_all"] = [m.weight_quantizer.get_soft_targets().detach().clone() for m in mods if m.org_weight is not None]
        trace["h0"] = trace["h0_all"][0]
    opt = torch.optim.Adam(params)                      # lr 1e-3 (layer_opt.py:254)
    deltas = []
    if learn_delta:
        # the option the reference keeps commented out (layer_opt.py:259-265): delta becomes a leaf with its own Adam
        # group; autograd of quantizer.py:437-449 supplies d loss / d delta (floor() has zero gradient)
        for m in mods:
            if m.org_weight is not None:
                m.weight_quantizer.delta = nn.Parameter(m.weight_quantizer.delta.detach().clone())
                deltas.append(m.weight_quantizer.delta)
        opt = torch.optim.Adam([{"params": params}, {"params": deltas, "lr": 1e-3 * delta_lr_scale}])
    loss_fn = LossFunction(unit, weight, iters, b_range, warmup, p, task_p)
    losses = []
    for it in range(iters):
        idx, keep = plan.draw(unit_id, it, q_in.size(0), batch_size, q_in.shape[1:], input_prob)
        cur = q_in[idx]
        if keep is not None:
            cur = torch.where(keep, cur, fp_in[idx])
        opt.zero_grad()
        out = unit(cur)
        if rd_task is not None:                         # task = R + lambda*D instead of lp(out, fp_out, task_p)
            err = loss_fn(out, fp_out[idx], None, None)
            task = rd_task(out, idx)
            loss_fn.last["task"] = float(task.detach())
            err = err + task
        else:
            err = loss_fn(out, fp_out[idx], out, fp_out[idx])
        err.backward()
        if trace is not None and it == 0:
            trace["grad0"] = [p_.grad.detach().clone() for p_ in params]
            trace["out0"] = out.detach().clone()
            trace["d_delta0"] = [d.grad.detach().clone().reshape(-1) for d in deltas]
        opt.step()
        for d in deltas:
            d.data.clamp_(min=1e-8)
        losses.append(float(err.detach()))
    for m in mods:
        if m.org_weight is not None:
            m.weight_quantizer.soft_targets = False
    for m in ([unit] if not blocks else [m for _, m in unit.named_modules() if isinstance(m, (QuantModule, BaseQuantBlock))]):
        m.trained = True
    return losses


def recon_model(qnn: QuantModel, cali_data, **kw):
    """main2.py:227-264 depth-first walk.  Returns {unit_name: loss trace}."""
    traces, counter = {}, [0]

    def walk(module, prefix):
        for name, m in module.named_children():
            full = f"{prefix}.{name}" if prefix else name
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                if not m.ignore_reconstruction:
                    traces[full] = reconstruct(qnn, m, counter[0], name, cali_data, **kw)
                counter[0] += 1
            else:
                walk(m, full)

    walk(qnn, "")
    return traces
