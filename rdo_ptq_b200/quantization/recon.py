"""Fused AdaRound reconstruction engine shared by `layer_reconstruction` and `block_reconstruction`.

One iteration of the reference loop (layer_opt.py:287-309 / block_opt.py:287-311) is issued as:
  gather_mix (batch pick + QDrop)  ->  K6 adaround_fwd per QuantModule (soft weight, a leaf tensor)
  -> unit forward on K1/K2/K3 kernels  ->  K11 lp_loss_fwd_bwd (loss value + dL/dout in one pass)
  -> wgrad/dgrad kernels (K4/K5, driven by the autograd tape of the unit)  ->  [NCCL all-reduce of dL/dWq]
  -> K6 adaround_bwd_adam per QuantModule (STE masks + rounding regulariser + Adam in one kernel).
No `.item()`/`float()` syncs inside the loop; the loss is read back only every `log_every` iterations.
"""
import logging
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import ops
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quantizer import AdaRoundQuantizer
from .utils import LinearTempDecay


class DrawPlan:
    """Source of the per-iteration randomness (layer_opt.py:289-292).  The default draws the batch indices with
    torch's device generator and the QDrop mask inside the gather_mix kernel (counter-based hash).  Tests pass a plan
    that replays explicit (idx, mask) pairs so the CPU oracle sees the same draws."""

    def __init__(self, seed: int = 1005):
        self.seed = seed

    def draw(self, unit_id: int, it: int, n: int, bs: int, shape, prob: float, device):
        g = torch.Generator(device=device)
        g.manual_seed(self.seed * 1000003 + unit_id * 100003 + it)
        idx = torch.randperm(n, generator=g, device=device)[:bs]
        return idx, None, (self.seed * 2654435761 + unit_id * 40503 + it) & 0xFFFFFFFFFFFF


class RDTask:
    """Task criterion R + lambda*D (BASELINE north_star; the reference keeps it commented out at layer_opt.py:146-148 /
    block_opt.py:145-147): the codec's forward is continued from the unit's output through the not-yet-trained modules
    (the role of `fp_out`, layer_opt.py:45-75) and RateDistortionLoss (losses/losses.py:15-28, MSE metric) is evaluated
    on it.  Both latents are rounded straight-through (round_ste, which fp_out applies to y at layer_opt.py:69), so the
    rate term back-propagates through the likelihood kernels (b200lic_gaussian_lik_bwd / _factorized_lik_bwd) and the
    distortion term through the dgrad kernels.  `unit_path` is the unit's path inside the codec, e.g. "g_a.2"."""

    def __init__(self, qnn, unit_path: str, cali_data: torch.Tensor, lmbda: float, batch: int = 8):
        codec = qnn.model
        coder, _, rest = unit_path.partition(".")
        if coder not in ("g_a", "h_a", "h_s", "g_s") or not hasattr(codec, "forward_from"):
            raise NotImplementedError(f"RDTask: no forward tail defined after {unit_path!r}")
        k = int(rest.split(".")[0])
        self.codec, self.coder, self.lmbda = codec, coder, float(lmbda)
        self.tail = list(getattr(codec, coder).children())[k + 1:]
        codec.entropy_bottleneck.ste_round = codec.gaussian_conditional.ste_round = True
        ys, zs = [], []
        with torch.no_grad():
            for i in range(0, cali_data.size(0), batch):
                y, z = codec.latents(cali_data[i:i + batch])
                ys.append(y)
                zs.append(z)
        self.ctx = {"x": cali_data, "y": torch.cat(ys), "z": torch.cat(zs)}

    def __call__(self, out: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        v = out
        for m in self.tail:
            v = m(v)
        ctx = {k: t[idx] for k, t in self.ctx.items()}          # batch rows: buffer plumbing
        o = self.codec.forward_from(self.coder, v, ctx)
        return ops.rd_loss(o["x_hat"], ctx["x"], o["bits"], self.lmbda)

    def close(self):
        self.codec.entropy_bottleneck.ste_round = self.codec.gaussian_conditional.ste_round = False


class CoderTask:
    """The task term the reference computes when `find_unquantized_module` finds later modules (layer_opt.py:45-75,
    219-224, 296-299): the unit's output is pushed through the not-yet-trained modules of the SAME sub-network in FP32
    (plus round_ste when that sub-network is g_a) and compared with the same tail applied to the FP unit output:
    task = lp_loss(tail(out_quant), tail(out_fp), p=args.task_loss).  With compressai's "0", "1", ... child names the
    reference's name test never fires and the tail is empty; `task="coder"` applies the intended rule by
    the unit's real position (`unit_path`, e.g. "g_a.2")."""

    def __init__(self, qnn, unit_path: str, fp_unit_out: torch.Tensor, p: float = 2.0, batch: int = 8):
        coder, _, rest = unit_path.partition(".")
        k = int(rest.split(".")[0])
        self.tail = list(getattr(qnn.model, coder).children())[k + 1:]
        self.round, self.p = coder == "g_a", float(p)
        outs = []
        with torch.no_grad():
            for i in range(0, fp_unit_out.size(0), batch):
                outs.append(self._run(fp_unit_out[i:i + batch]))
        self.target = torch.cat(outs)

    def _run(self, v):
        for m in self.tail:
            v = m(v)
        if self.round:
            v = ops.round_latent_ste(v) if (torch.is_grad_enabled() and v.requires_grad) else ops.round_latent(v)
        return v

    def __call__(self, out: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        v = self._run(out)
        return ops.lp_loss_fn(v, self.target[idx], self.p, 1.0 / (v.numel() // v.shape[1]))

    def close(self):
        pass


FUSED_DEFAULT = os.environ.get("B200LIC_FUSED", "1") != "0"
# multi-GPU tail: the fused peer-memory kernel (1) or ncclAllReduce + Adam (0)
XGPU_DEFAULT = os.environ.get("B200LIC_XGPU", "1") != "0"
# the folded-tap 3-channel layers through the fused chain as 1x1 problems (FusedFolded); A/B switch like FUSED_DEFAULT
FOLDED_DEFAULT = os.environ.get("B200LIC_FUSED_FOLDED", "1") != "0"


class FusedLayer:
    """Static buffers of the FUSED iteration of a conv / transposed-conv unit (prepared operands, include/b200lic.h):
    stage_mix (pick + QDrop -> staged activation operand) -> quant_pack (AdaRound soft weight -> packed weight operand)
    -> GEMM -> loss_stage (loss + gradient -> staged dY operand) -> wgrad with the STE / regulariser / Adam tail fused
    behind its split-K reduction.  Seven launches; no fp32 batch, soft weight, dL/dout or dL/dWq tensor is written.
    Not applicable (-> None from `build`): blocks, PixelShuffle, the SIMT engine; GDN units take `FusedGdn`, the folded-tap
    3-channel layers `FusedFolded`."""

    def __init__(self):
        self.d = self.tr = self.packed = self.ws_f = self.ws_w = self.x_slot = self.dy_slot = self.y = self.dw = None

    @staticmethod
    def build(m, in_shape, batch, world):
        if not isinstance(m, QuantModule) or m.is_gdn or m.is_ps or m.org_weight is None or m.se_module is not None:
            return None
        kw = m.fwd_kwargs
        if ops._sq(kw["dilation"], "dilation") != 1 or kw["groups"] != 1 or ops.DEFAULT_ENGINE == ops.ENGINE_SIMT:
            return None
        f = FusedLayer()
        act, slope = ops._act_id(m.activation_function)
        f.tr = m.if_tconv
        f.d = ops.conv_desc((batch,) + tuple(in_shape), m.weight.shape, kw["stride"], kw["padding"], f.tr,
                            kw.get("output_padding", 0), act=act, slope=slope)
        dev = m.weight.device
        f.packed = ops.new_packed(f.d, f.tr, dev)
        if f.packed is None:
            return None
        f.ws_f = ops._workspace(f.d, ops.fwd_op(f.tr), dev)
        f.ws_w = ops._workspace(f.d, ops.wgrad_op(f.tr), dev)
        f.x_slot = ops.conv_x_slot(f.d, f.tr, f.ws_f)
        f.dy_slot = ops.conv_dy_slot(f.d, f.tr, f.ws_w)
        if f.x_slot is None or f.dy_slot is None:
            return None
        f.y = torch.empty((f.d.N, f.d.Cout, f.d.Ho, f.d.Wo), device=dev, dtype=torch.float32)
        f.dw = torch.empty_like(m.weight.data) if world > 1 else None
        return f

    def use_peer(self, peer, weight):
        """N > 1 with the peer-memory tail: the weight gradient is written straight into the symmetric buffer."""
        self.dw = peer.grad.view(weight.shape)


class FusedFolded:
    """Fused iteration of the two folded-tap layers (3 -> N analysis conv, N -> 3 synthesis transposed conv): the engine
    runs them as 1x1 problems over an im2col'ed operand (conv_tc_smallc.cu), so they take the prepared-operand chain of
    `FusedLayer` on the 1x1 descriptor `d1`:
      conv : pick + QDrop (3-channel batch) -> im2col_stage (the staged activation operand, K = Cin*k*k channels; the SAME
             buffer is the weight gradient's operand -- the un-fused path builds it twice) -> quant_pack -> GEMM ->
             loss_stage -> wgrad + STE / regulariser / Adam
      tconv: stage_mix -> quant_pack -> GEMM (col = x . W'') -> col2im (+ bias) -> loss on the 3-channel output ->
             im2col_stage of dL/dy as the staged dY operand -> wgrad + tail (the staged x is reused; the un-fused path
             splits it a second time)."""

    @staticmethod
    def build(m, in_shape, batch, world):
        if not isinstance(m, QuantModule) or m.is_gdn or m.is_ps or m.org_weight is None or m.se_module is not None:
            return None
        kw = m.fwd_kwargs
        if ops._sq(kw["dilation"], "dilation") != 1 or kw["groups"] != 1 or ops.DEFAULT_ENGINE == ops.ENGINE_SIMT:
            return None
        act, slope = ops._act_id(m.activation_function)
        f = FusedFolded()
        f.tr = m.if_tconv
        f.full = ops.conv_desc((batch,) + tuple(in_shape), m.weight.shape, kw["stride"], kw["padding"], f.tr,
                               kw.get("output_padding", 0), act=act, slope=slope)
        g = f.full
        kk = g.KH * g.KW
        if kk <= 1 or g.stride > 4:
            return None
        dev = m.weight.device
        if not f.tr:
            K = g.Cin * kk
            if K > 128:
                return None
            f.d1 = ops.conv_desc((batch, K, g.Ho, g.Wo), (g.Cout, K, 1, 1), 1, 0, False, act=act, slope=slope)
        else:
            K = g.Cout * kk
            if K > 128 or act != ops.ACT_NONE:
                return None
            f.d1 = ops.conv_desc((batch, g.Cin, g.H, g.W), (g.Cin, K, 1, 1), 1, 0, True)
            f.col = torch.empty((batch, K, g.H, g.W), device=dev, dtype=torch.float32)
        f.packed = ops.new_packed(f.d1, f.tr, dev)
        if f.packed is None:
            return None
        f.ws_f = ops._workspace(f.d1, ops.fwd_op(f.tr), dev)
        f.ws_w = ops._workspace(f.d1, ops.wgrad_op(f.tr), dev)
        f.x_slot = ops.conv_x_slot(f.d1, f.tr, f.ws_f)
        f.dy_slot = ops.conv_dy_slot(f.d1, f.tr, f.ws_w)
        if f.x_slot is None or f.dy_slot is None or (not f.tr and f.x_slot[2] < K) or (f.tr and f.dy_slot[2] < K):
            return None
        f.xb = None if f.tr else torch.empty((batch,) + tuple(in_shape), device=dev, dtype=torch.float32)
        f.y = torch.empty((g.N, g.Cout, g.Ho, g.Wo), device=dev, dtype=torch.float32)
        f.dw = torch.empty_like(m.weight.data) if world > 1 else None
        return f

    def use_peer(self, peer, weight):
        self.dw = peer.grad.view(weight.shape)

    def d1_wshape(self, m):
        """The weight (gradient) seen through the 1x1 descriptor: same memory, taps folded into a channel axis."""
        w = m.weight.shape
        return (w[0], w[1] * w[2] * w[3], 1, 1)


class FusedGdn:
    """Static buffers of the fused iteration of a GDN / IGDN unit (f_gdn, quant_layer.py:142-154):
    stage_mix (pick + QDrop -> the fp32 batch x for the epilogue AND the staged x*x operand) -> AdaRound soft gamma ->
    re-parametrise -> pack -> GEMM (the accumulator norm = beta + gamma . x*x only) -> loss_stage with the GDN forward tail
    and backward folded in (y = x * norm^-+1/2 recomputed, dL/dy -> dL/dnorm -> staged dY operand; y, dL/dy and dL/dnorm
    are never written) -> wgrad of gamma on the staged x*x -> re-parametrisation backward -> STE / regulariser / Adam."""

    @staticmethod
    def build(m, in_shape, batch, world):
        if not isinstance(m, QuantModule) or not m.is_gdn or ops.DEFAULT_ENGINE == ops.ENGINE_SIMT:
            return None
        if ops._act_id(m.activation_function)[0] != ops.ACT_NONE:
            return None
        f = FusedGdn()
        f.inverse = bool(m.fwd_kwargs["inverse"])
        f.d = ops.gdn_desc((batch,) + tuple(in_shape), f.inverse)
        f.d.gdn_mode = 0        # calibration only needs the accumulator norm = beta + gamma . x^2: a plain 1x1 conv of x*x
        Cc = in_shape[0]
        f.dw = ops.ConvDesc(f.d.N, Cc, f.d.H, f.d.W, Cc, f.d.H, f.d.W, 1, 1, 1, 0, ops.ACT_NONE, 0.0, f.d.engine, 1, 0, 0)
        dev = m.weight.device
        f.packed = ops.new_packed(f.d, False, dev)
        if f.packed is None:
            return None
        f.ws_f = ops._workspace(f.d, ops.fwd_op(False), dev)
        f.ws_w = ops._workspace(f.dw, ops.wgrad_op(False), dev)
        f.x_slot = ops.conv_x_slot(f.d, False, f.ws_f)
        f.dy_slot = ops.conv_dy_slot(f.dw, False, f.ws_w)
        if f.x_slot is None or f.dy_slot is None:
            return None
        shape = (batch,) + tuple(in_shape)
        f.x = torch.empty(shape, device=dev, dtype=torch.float32)
        f.norm = torch.empty_like(f.x)
        f.leaf = torch.empty_like(m.weight.data)
        f.g_eff, f.dgamma, f.dleaf = torch.empty_like(f.leaf), torch.empty_like(f.leaf), torch.empty_like(f.leaf)
        f.b_eff = torch.empty_like(m.bias.data)
        f.gr, f.br = m.fwd_kwargs["gamma_reparam"], m.fwd_kwargs["beta_reparam"]
        return f


class UnitTrainer:
    """State of one reconstruction problem: the QuantModules whose alpha is trained, Adam moments, schedules."""

    def __init__(self, unit, iters: int, weight: float, b_range, warmup: float, p: float, task_p: Optional[float],
                 lr: float = 1e-3, process_group=None, rd_task: Optional[RDTask] = None, learn_delta: bool = False,
                 delta_lr_scale: float = 0.1):
        self.unit = unit
        self.rd_task = rd_task
        # learned per-channel step sizes (LSQ-style; the reference keeps the option as commented-out code,
        # layer_opt.py:259-265): delta gets its own fused Adam at delta_lr_scale * lr, zero_point stays fixed
        self.learn_delta, self.delta_lr_scale = learn_delta, delta_lr_scale
        self.mods: List[QuantModule] = ([unit] if isinstance(unit, QuantModule) else
                                        [m for _, m in unit.named_modules() if isinstance(m, QuantModule)])
        self.mods = [m for m in self.mods if m.org_weight is not None]
        self.iters, self.weight, self.p, self.task_p, self.lr = iters, weight, p, task_p, lr
        self.loss_start = iters * warmup
        self.temp_decay = LinearTempDecay(iters, rel_start_decay=warmup, start_b=b_range[0], end_b=b_range[1])
        self.count = 0
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        for m in self.mods:
            if not isinstance(m.weight_quantizer, AdaRoundQuantizer):
                m.weight_quantizer = AdaRoundQuantizer(uaq=m.weight_quantizer, round_mode='learned_hard_sigmoid',
                                                       weight_tensor=m.org_weight.data)
            m.weight_quantizer.soft_targets = True
        # Only alpha is optimised (layer_opt.py:253-254); the fused loop never needs autograd gradients of the unit's own
        # parameters (bias, GDN beta, un-quantised weights), so they are frozen for the duration: no dbeta reduction, no
        # reparametrisation backward.  finish() restores the flags.
        self._frozen = [p for p in unit.parameters() if p.requires_grad]
        for p in self._frozen:
            p.requires_grad_(False)
        self.exp_avg = [torch.zeros_like(m.weight_quantizer.alpha.data) for m in self.mods]
        self.exp_avg_sq = [torch.zeros_like(m.weight_quantizer.alpha.data) for m in self.mods]
        if learn_delta:
            for m in self.mods:        # own storage: the AdaRound quantiser shares delta with the quantiser it wraps
                m.weight_quantizer.delta = m.weight_quantizer.delta.clone().contiguous()
            self.d_exp_avg = [torch.zeros(m.weight_quantizer.delta.numel(), device=m.weight.device) for m in self.mods]
            self.d_exp_avg_sq = [torch.zeros_like(t) for t in self.d_exp_avg]
        dev = self.mods[0].weight.device if self.mods else None
        self.loss_buf = torch.zeros(3, device=dev)      # [rec, task, round] accumulated since the last read
        self.last = {}
        self.fused = FUSED_DEFAULT                      # A/B switch: the fused iteration where it applies
        self._fused_plan = False                        # False = not built yet; None = not applicable
        # N > 1: gradients are summed and alpha is updated by ONE kernel over NVLink peer memory
        # (b200lic_xgpu_reduce_adam_sched) instead of ncclAllReduce + Adam; alpha moves into peer-mapped buffers.  Every
        # rank takes the same decision (same shapes, same environment).
        self.peer = None
        # the kernel's exit barrier (peers' alpha stores have landed); a sequential multi-unit session turns it off: the
        # next unit's entry barrier orders the same stores before this unit runs again (b200lic_xgpu_reduce_adam_sched)
        self.peer_exit_barrier = True
        if self.world > 1 and XGPU_DEFAULT and not learn_delta and self.mods and \
                all(m.weight.numel() % 4 == 0 for m in self.mods):
            from .. import dist as _dist
            self.peer = [_dist.PeerLayer(m.weight_quantizer.alpha, process_group) for m in self.mods]

    # -- the fused iteration (single conv / transposed-conv unit, reference-default loss) --------------------------------
    def fused_plan(self, in_shape, batch):
        if self._fused_plan is False:
            ok = (self.fused and self.rd_task is None and not self.learn_delta and self.task_p is not None and
                  float(self.task_p) == float(self.p) and len(self.mods) == 1 and self.mods[0] is self.unit)
            self._fused_plan = None
            if ok:
                self._fused_plan = (FusedGdn.build(self.unit, in_shape, batch, self.world) if self.unit.is_gdn else
                                    FusedLayer.build(self.unit, in_shape, batch, self.world))
                if self._fused_plan is None and not self.unit.is_gdn and FOLDED_DEFAULT:
                    self._fused_plan = FusedFolded.build(self.unit, in_shape, batch, self.world)
                if self._fused_plan is not None and self.peer is not None:
                    if isinstance(self._fused_plan, FusedGdn):
                        self._fused_plan.dleaf = self.peer[0].grad.view(self._fused_plan.leaf.shape)
                    else:
                        self._fused_plan.use_peer(self.peer[0], self.unit.weight.data)
        return self._fused_plan

    def _fused_gdn(self, f, q_in, fp_in, tgt_cache, idx_table, units, unit, sched, prob, seed_base):
        m = self.unit
        q = m.weight_quantizer
        C_ = f.leaf.shape[0]
        ops.stage_mix_sched(q_in, fp_in, idx_table, f.d.N, prob, seed_base, units, unit, sched, f.x_slot, square=True,
                            out=f.x)
        ops.adaround_fwd(m.weight.data, q.alpha.data, q.delta, q.zero_point, q.axis, q.n_levels, True, out=f.leaf)
        ops.gdn_reparam(f.leaf, f.gr.bound_value, f.gr.pedestal_value, out=f.g_eff)
        ops.gdn_reparam(m.bias.data, f.br.bound_value, f.br.pedestal_value, out=f.b_eff)
        ops.pack_weights(f.g_eff.view(C_, C_, 1, 1), f.d, False, out=f.packed)
        ops.conv_fwd_packed(None, f.packed, f.d, False, bias=f.b_eff, ws=f.ws_f, y=f.norm)
        denom = f.norm.numel() // f.norm.shape[1]
        ops.lp_loss_stage_sched(None, tgt_cache, idx_table, units, unit, sched, self.p, 1.0 / denom, 2.0 / denom,
                                ops.ACT_NONE, 0.0, self.loss_buf[0:1], f.dy_slot, gdn=(f.x, f.norm, f.inverse))
        ops.conv_wgrad_prepared(f.dw, False, f.x_slot, None, f.dgamma.view(C_, C_, 1, 1), f.ws_w)
        ops.gdn_reparam_bwd(f.leaf, f.dgamma, f.gr.bound_value, out=f.dleaf)
        self._same = True
        if self.world > 1:
            self._flat = f.dleaf.view(-1)
            return [f.dleaf]
        ops.adaround_bwd_adam_sched(m.weight.data, q.alpha.data, q.delta, q.zero_point, f.dleaf, self.exp_avg[0],
                                    self.exp_avg_sq[0], q.axis, q.n_levels, sched, reg_weight=self.weight,
                                    reg_loss=self.loss_buf[2:3])
        return []

    def _fused_folded(self, f, q_in, fp_in, tgt_cache, idx_table, units, unit, sched, prob, seed_base):
        m = self.unit
        q = m.weight_quantizer
        g = f.full
        if not f.tr:
            ops.gather_mix_sched(q_in, fp_in, idx_table, g.N, prob, seed_base, units, unit, sched, out=f.xb)
            ops.im2col_stage(f.xb, g.KH, g.KW, g.stride, g.pad, g.Ho, g.Wo, f.x_slot)
        else:
            ops.stage_mix_sched(q_in, fp_in, idx_table, g.N, prob, seed_base, units, unit, sched, f.x_slot)
        ops.quant_pack_weights(m.weight.data, q.alpha.data, q.delta, q.zero_point, q.axis, q.n_levels, True, f.d1, f.tr,
                               out=f.packed)
        denom = f.y.numel() // f.y.shape[1]
        if not f.tr:
            ops.conv_fwd_packed(None, f.packed, f.d1, False, bias=m.bias, ws=f.ws_f, y=f.y)
            ops.lp_loss_stage_sched(f.y, tgt_cache, idx_table, units, unit, sched, self.p, 1.0 / denom, 2.0 / denom,
                                    g.act, g.act_slope, self.loss_buf[0:1], f.dy_slot)
        else:
            ops.conv_fwd_packed(None, f.packed, f.d1, True, ws=f.ws_f, y=f.col)
            ops.col2im(f.col, m.bias, g.Cout, g.KH, g.KW, g.stride, g.pad, g.Ho, g.Wo, y=f.y)
            _, d_y = ops.lp_loss_fwd_bwd(f.y, tgt_cache, self.p, scale=1.0 / denom, grad_scale=2.0 / denom,
                                         loss=self.loss_buf[0:1],
                                         pick=None if idx_table is None else (idx_table, units, unit, sched))
            # dL/dcol = im2col(dL/dy) over the conv geometry whose output grid is the transposed conv's input grid
            ops.im2col_stage(d_y, g.KH, g.KW, g.stride, g.pad, g.H, g.W, f.dy_slot)
        self._same = True
        if self.world > 1:
            ops.conv_wgrad_prepared(f.d1, f.tr, f.x_slot, None, f.dw.view(f.d1_wshape(m)), f.ws_w)
            self._flat = f.dw.view(-1)
            return [f.dw]
        ops.conv_wgrad_adam_sched(f.d1, f.tr, f.x_slot, None, f.ws_w, m.weight.data, q.alpha.data, q.delta, q.zero_point,
                                  self.exp_avg[0], self.exp_avg_sq[0], q.axis, q.n_levels, sched, reg_weight=self.weight,
                                  reg_loss=self.loss_buf[2:3])
        return []

    def fused_compute(self, f, q_in, fp_in, tgt_cache, idx_table, units, unit, sched, prob, seed_base):
        """One fused iteration up to (world == 1: and including) the Adam step.  Returns the list of dL/dWq tensors for
        `step_update` (empty when the tail ran fused)."""
        if isinstance(f, FusedGdn):
            return self._fused_gdn(f, q_in, fp_in, tgt_cache, idx_table, units, unit, sched, prob, seed_base)
        if isinstance(f, FusedFolded):
            return self._fused_folded(f, q_in, fp_in, tgt_cache, idx_table, units, unit, sched, prob, seed_base)
        m = self.unit
        q = m.weight_quantizer
        ops.stage_mix_sched(q_in, fp_in, idx_table, f.d.N, prob, seed_base, units, unit, sched, f.x_slot)
        ops.quant_pack_weights(m.weight.data, q.alpha.data, q.delta, q.zero_point, q.axis, q.n_levels, True, f.d, f.tr,
                               out=f.packed)
        ops.conv_fwd_packed(None, f.packed, f.d, f.tr, bias=m.bias, ws=f.ws_f, y=f.y)
        denom = f.y.numel() // f.y.shape[1]
        ops.lp_loss_stage_sched(f.y, tgt_cache, idx_table, units, unit, sched, self.p, 1.0 / denom, 2.0 / denom, f.d.act,
                                f.d.act_slope, self.loss_buf[0:1], f.dy_slot)
        self._same = True
        if self.world > 1:
            ops.conv_wgrad_prepared(f.d, f.tr, f.x_slot, None, f.dw, f.ws_w)
            self._flat = f.dw.view(-1)
            return [f.dw]
        ops.conv_wgrad_adam_sched(f.d, f.tr, f.x_slot, None, f.ws_w, m.weight.data, q.alpha.data, q.delta, q.zero_point,
                                  self.exp_avg[0], self.exp_avg_sq[0], q.axis, q.n_levels, sched, reg_weight=self.weight,
                                  reg_loss=self.loss_buf[2:3])
        return []

    # -- one iteration ---------------------------------------------------------------------------------------
    def step(self, cur_inp: torch.Tensor, tgt: torch.Tensor, trace: Optional[dict] = None, sched=None, idx=None):
        """One fused iteration.  With `sched` (a device-resident b200lic_calib_sched, already ticked for this
        iteration) no host scalar that changes between iterations enters a kernel argument, so the call sequence can be
        captured once as a CUDA graph and replayed (session.py).  `idx` = the batch pick (the R + lambda*D task needs
        the images and upstream latents of the same samples)."""
        out, grads = self.step_compute(cur_inp, tgt, idx=idx)
        self.step_update(grads, trace=trace, sched=sched)
        if trace is not None:
            trace["out"] = out.detach()
        return out

    def step_compute(self, cur_inp: torch.Tensor, tgt, tgt_pick=None, idx=None):
        """Soft weights -> unit forward -> loss value + dL/dout -> wgrad/dgrad.  Returns (out, [dL/dWq per module]).
        `tgt` is the target batch, or (with `tgt_pick = (idx_table, units, unit, sched)`) the whole cached-output tensor
        whose rows the loss kernel picks itself, so the target batch is never materialised."""
        leaves = []
        for m in self.mods:
            q = m.weight_quantizer
            leaf = ops.adaround_fwd(m.weight.data, q.alpha.data, q.delta, q.zero_point, q.axis, q.n_levels, True)
            leaf.requires_grad_(True)
            q._leaf = leaf
            leaves.append(leaf)
        try:
            with torch.enable_grad():
                out = self.unit(cur_inp)
            # rec + task (with compressai-style names fp_out is the identity, so task == lp(out, tgt, task_p))
            denom = out.numel() // out.shape[1]
            rec = self.loss_buf[0:1]
            if self.rd_task is not None:
                # rec gradient from the loss kernel, task gradient by autograd through the codec's tail
                _, d_out = ops.lp_loss_fwd_bwd(out.detach(), tgt, self.p, scale=1.0 / denom, loss=rec)
                task = self.rd_task(out, idx)
                (g_task,) = torch.autograd.grad(task, out, retain_graph=True)
                ops.add_act(self.loss_buf[1:2], task.detach().reshape(1), out=self.loss_buf[1:2])
                d_out = ops.add_act(d_out, g_task)
                same = False
            elif self.task_p is not None and float(self.task_p) == float(self.p):
                _, d_out = ops.lp_loss_fwd_bwd(out.detach(), tgt, self.p, scale=1.0 / denom, grad_scale=2.0 / denom,
                                               loss=rec, pick=tgt_pick)
                same = True
            else:
                if tgt_pick is not None:
                    raise NotImplementedError("picked targets are only wired for task_p == p (the reference default)")
                _, d_out = ops.lp_loss_fwd_bwd(out.detach(), tgt, self.p, scale=1.0 / denom, loss=rec)
                if self.task_p is not None:
                    _, d2 = ops.lp_loss_fwd_bwd(out.detach(), tgt, self.task_p, scale=1.0 / denom,
                                                loss=self.loss_buf[1:2])
                    d_out = ops.add_act(d_out, d2)
                same = False
            self._same = same
            out.backward(d_out)
        finally:
            for m in self.mods:
                m.weight_quantizer._leaf = None
        grads = [l.grad for l in leaves]
        if self.world > 1:          # one flat bucket per unit for the all-reduce of step_update
            flat = torch.cat([g.reshape(-1) for g in grads])
            grads = [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)]
            self._flat = flat
        return out, grads

    def step_update(self, grads, trace: Optional[dict] = None, sched=None):
        """[NCCL all-reduce of the unit's dL/dWq bucket] -> STE masks + rounding regulariser + Adam on alpha.  Touches
        only this unit's state, so a session may run it on a side stream under the next unit's step_compute."""
        if sched is None:
            self.count += 1
            b = self.temp_decay(self.count)
            reg_b = 0.0 if self.count < self.loss_start else float(b)
        if not grads:                                # the fused tail already applied this iteration's Adam step
            return
        if self.peer is not None and sched is not None:
            for i, m in enumerate(self.mods):
                q, pl = m.weight_quantizer, self.peer[i]
                if grads[i].data_ptr() != pl.grad.data_ptr():
                    pl.grad.copy_(grads[i].reshape(-1))       # autograd-tape units: stage the gradient for the peers
                ops.xgpu_reduce_adam_sched(pl, m.weight.data, q.delta, q.zero_point, self.exp_avg[i], self.exp_avg_sq[i],
                                           q.axis, q.n_levels, sched, grad_scale=1.0 / self.world, reg_weight=self.weight,
                                           reg_loss=self.loss_buf[2:3], exit_barrier=self.peer_exit_barrier)
            return
        if self.world > 1:
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.pg)
        for i, m in enumerate(self.mods):
            q = m.weight_quantizer
            if self.learn_delta:        # before the alpha step: both gradients belong to the same (alpha, delta) point
                if sched is not None:
                    ops.lsq_delta_grad(m.weight.data, q.delta, q.zero_point, grads[i], q.axis, q.n_levels,
                                       alpha=q.alpha.data, soft=True, grad_scale=1.0 / self.world,
                                       adam=(self.d_exp_avg[i], self.d_exp_avg_sq[i]), sched=sched,
                                       lr_scale=self.delta_lr_scale)
                else:
                    d_delta = ops.lsq_delta_grad(m.weight.data, q.delta, q.zero_point, grads[i], q.axis, q.n_levels,
                                                 alpha=q.alpha.data, soft=True, grad_scale=1.0 / self.world,
                                                 adam=(self.d_exp_avg[i], self.d_exp_avg_sq[i], self.count,
                                                       self.lr * self.delta_lr_scale))
                    if trace is not None:
                        trace.setdefault("d_delta", []).append(d_delta)
            if sched is not None:
                ops.adaround_bwd_adam_sched(m.weight.data, q.alpha.data, q.delta, q.zero_point, grads[i],
                                            self.exp_avg[i], self.exp_avg_sq[i], q.axis, q.n_levels, sched,
                                            grad_scale=1.0 / self.world, reg_weight=self.weight,
                                            reg_loss=self.loss_buf[2:3])
                continue
            d_alpha = torch.empty_like(q.alpha.data) if trace is not None else None
            ops.adaround_bwd_adam(m.weight.data, q.alpha.data, q.delta, q.zero_point, grads[i], self.exp_avg[i],
                                  self.exp_avg_sq[i], q.axis, q.n_levels, self.count, lr=self.lr,
                                  grad_scale=1.0 / self.world, reg_weight=self.weight, reg_b=reg_b,
                                  reg_loss=self.loss_buf[2:3], d_alpha_out=d_alpha)
            if trace is not None:
                trace.setdefault("d_alpha", []).append(d_alpha)

    def read_losses(self, since: int):
        """One device->host read of the accumulated (rec, task, round) sums; returns per-iteration means."""
        v = (self.loss_buf / max(since, 1)).tolist()
        self.loss_buf.zero_()
        rec, task, rnd = v
        if getattr(self, "_same", False):
            task = rec
        self.last = dict(rec=rec, task=task, round=rnd, total=rec + task + rnd)
        return self.last

    def finish(self):
        for p in self._frozen:
            p.requires_grad_(True)
        self._frozen = []
        for m in self.mods:
            m.weight_quantizer.soft_targets = False
            m.invalidate_prepared()             # alpha was written by kernels: any operand cached before is stale
        marks = ([self.unit] if isinstance(self.unit, QuantModule) else
                 [m for _, m in self.unit.named_modules() if isinstance(m, (QuantModule, BaseQuantBlock))])
        for m in marks:
            m.act_quantizer.is_training = False
            m.trained = True
        if self.rd_task is not None:
            self.rd_task.close()
        if self.peer is not None:           # alpha leaves the peer-mapped buffers (they die with the trainer)
            if not self.peer_exit_barrier:  # the last updates' peer stores: complete on every rank before alpha is read
                torch.cuda.synchronize()
                dist.barrier(self.pg)
            for m in self.mods:
                a = m.weight_quantizer.alpha
                a.data = a.data.clone()
            self.peer = None


def run_reconstruction(trainer: UnitTrainer, cached_inps, cached_outs, batch_size: int, input_prob: float,
                       unit_id: int = 0, plan: Optional[DrawPlan] = None, log_every: int = 500, trace=None,
                       graph: bool = True, graph_warmup: int = 2):
    """The hot loop: `iters` fused iterations over the cached (quant_in, fp_in) -> fp_out pairs.

    Default (no explicit `plan`): batch picks are pre-drawn into a device table, the QDrop mask comes from the
    counter-based hash inside gather_mix, the iteration-dependent scalars live in a device `b200lic_calib_sched`, and
    after `graph_warmup` eager iterations ONE captured CUDA graph of the whole iteration is replayed -- the reference's
    ~60 launches + 2 host syncs per iteration (layer_opt.py:287-309) become one graph launch.
    With an explicit `plan` (tests replaying the oracle's draws) every iteration is issued eagerly."""
    q_in, fp_in = cached_inps[0], (cached_inps[1] if len(cached_inps) > 1 else cached_inps[0])
    n = q_in.size(0)
    batch_size = min(batch_size, n)      # the reference's randperm(n)[:batch_size] just truncates (layer_opt.py:289)
    losses, since = [], 0

    def log(it):
        nonlocal since
        l = trainer.read_losses(since)
        since = 0
        losses.append(l)
        cnt = it + 1
        logging.info('Total loss:\t{:.3f} ( task:{:.3f}, rec:{:.3f}, round:{:.3f})\tb={:.2f}\tcount={}'.format(
            l["total"], l["task"], l["rec"], l["round"], trainer.temp_decay(cnt), cnt))

    if plan is None and trainer.rd_task is not None:
        plan = DrawPlan()                       # the rate-distortion tail is issued eagerly (autograd through the codec)
    if plan is not None:
        for it in range(trainer.iters):
            idx, mask, seed = plan.draw(unit_id, it, n, batch_size, q_in.shape[1:], input_prob, q_in.device)
            cur_inp = ops.gather_mix(q_in, fp_in, idx, prob=input_prob, seed=seed, mask=mask)
            tgt = ops.gather_mix(cached_outs, cached_outs, idx, prob=1.0)
            trainer.step(cur_inp, tgt, trace=trace if (trace is not None and it == 0) else None, idx=idx)
            since += 1
            if trainer.count % log_every == 0 or it == trainer.iters - 1:
                log(it)
        trainer.finish()
        return losses

    if trainer.rd_task is not None:
        raise NotImplementedError("the R + lambda*D task runs on the explicit-plan (eager) path")
    dev = q_in.device
    seed = DrawPlan().seed
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed * 1000003 + unit_id * 100003)
    # one fresh random subset per iteration, like the reference's randperm inside the loop (no wrap-around): up to 4096
    # rows drawn one randperm at a time, longer runs in one shot (argsort of uniform keys = a uniform random permutation)
    rows = max(trainer.iters, 1)
    if rows <= 4096:
        table = torch.stack([torch.randperm(n, generator=gen, device=dev)[:batch_size] for _ in range(rows)])
    else:
        table = torch.cat([torch.rand(min(8192, rows - r0), n, generator=gen, device=dev).argsort(dim=1)[:, :batch_size]
                           for r0 in range(0, rows, 8192)]).contiguous()
    seed_base = (seed * 2654435761 + unit_id * 40503) & 0xFFFFFFFFFFFF
    sched = ops.new_sched(dev)
    tick = (trainer.iters, trainer.loss_start / trainer.iters if trainer.iters else 0.0, trainer.temp_decay.start_b,
            trainer.temp_decay.end_b, trainer.lr)

    fused = trainer.fused_plan(q_in.shape[1:], batch_size) if table.size(1) == batch_size else None

    def body():
        if fused is not None:
            grads = trainer.fused_compute(fused, q_in, fp_in, cached_outs, table, 1, 0, sched, input_prob, seed_base)
            trainer.step_update(grads, sched=sched)
            return
        cur_inp = ops.gather_mix_sched(q_in, fp_in, table, batch_size, input_prob, seed_base, 1, 0, sched)
        tgt = ops.gather_mix_sched(cached_outs, cached_outs, table, batch_size, 1.0, seed_base, 1, 0, sched)
        trainer.step(cur_inp, tgt, sched=sched)

    g = None
    for it in range(trainer.iters):
        ops.sched_tick(sched, *tick)
        if graph and it >= graph_warmup:
            if g is None:
                g = torch.cuda.CUDAGraph()
                kw = {"capture_error_mode": "thread_local"} if trainer.world > 1 else {}
                with torch.cuda.graph(g, **kw):
                    body()
            g.replay()
        else:
            body()
        trainer.count += 1
        since += 1
        if trainer.count % log_every == 0 or it == trainer.iters - 1:
            log(it)
    trainer.finish()
    return losses
