"""B200 drop-in for task-oriented-PTQ/quantization/quant_layer.py (QuantModule, f_gdn).

Same constructor, attributes and state machine as the reference (quant_layer.py:11-139); the forward issues
libb200lic kernels: weight quantiser (K7/K6) -> implicit-GEMM conv / transposed conv / fused GDN (K1/K2/K3) with the
absorbed activation applied in the conv epilogue -> dynamic activation quantiser (K8).
"""
import os
from typing import Union

import torch
import torch.nn as nn

from .. import ops
from ..codec.layers import GDN, f_gdn      # noqa: F401  (f_gdn is re-exported like the reference module)
from .quantizer import AdaRoundQuantizer, StraightThrough, UniformAffineQuantizer


# Masked context convolution (compressai MaskedConv2d; the wrap drops the mask, quant_model.py:45-48): at
# evaluation the taps behind the mask are contracted only when they hold something.  B200LIC_MASKED_TAPS=0: always dense.
MASKED_TAPS = os.environ.get("B200LIC_MASKED_TAPS", "1") != "0"


def _live_tap_prefix(mask):
    """Length L of the raster-order prefix of live taps when `mask` [Cout,Cin,KH,KW] is the same 1...10...0 pattern for
    every channel pair (mask 'A' of a 5x5 kernel: 12; mask 'B': 13), else 0."""
    if mask is None or mask.dim() != 4:
        return 0
    flat = mask.reshape(mask.shape[0] * mask.shape[1], -1)
    L = int(flat[0].sum().item())
    pattern = torch.zeros_like(flat[0])
    pattern[:L] = 1
    return L if 0 < L < flat.shape[1] and bool((flat == pattern).all()) else 0


class QuantModule(nn.Module):
    def __init__(self, org_module: Union[nn.Conv2d, nn.ConvTranspose2d, GDN, nn.PixelShuffle],
                 weight_quant_params: dict = {}, act_quant_params: dict = {}, disable_act_quant: bool = False,
                 se_module=None):
        super().__init__()
        self.if_layer_norm = False
        self.is_linear = False
        self.if_tconv = False
        self.is_ps = False
        self.is_gdn = False
        if isinstance(org_module, nn.ConvTranspose2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   output_padding=org_module.output_padding, dilation=org_module.dilation,
                                   groups=org_module.groups)
            self.fwd_func = ops.conv_transpose2d
            self.if_tconv = True
        elif isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = ops.conv2d
            self.mask_live_taps = _live_tap_prefix(getattr(org_module, "mask", None))
        elif isinstance(org_module, GDN):
            self.fwd_kwargs = dict(inverse=org_module.inverse, gamma_reparam=org_module.gamma_reparam,
                                   beta_reparam=org_module.beta_reparam)
            self.fwd_func = f_gdn
            self.is_gdn = True
        elif isinstance(org_module, nn.PixelShuffle):
            self.fwd_kwargs = org_module.upscale_factor
            self.fwd_func = ops.pixel_shuffle
            self.is_ps = True
        elif isinstance(org_module, nn.Linear):
            # quant_layer.py:38-41; forward only (evaluation): the Swin blocks that train through these wrappers
            # (quant_block.py:330-641) are not built, SURVEY 8(f) N4
            self.fwd_kwargs = dict()
            self.fwd_func = ops.linear
            self.is_linear = True
        elif isinstance(org_module, nn.LayerNorm):
            # quant_layer.py:43-48: gamma goes through the weight quantiser (a 1-D weight: one range for the tensor)
            self.fwd_kwargs = dict(normalized_shape=org_module.normalized_shape, eps=org_module.eps)
            self.fwd_func = ops.layer_norm
            self.if_layer_norm = True
            if len(org_module.normalized_shape) != 1:
                raise NotImplementedError("LayerNorm over more than the last axis")
        else:
            raise ValueError('Not supported modules: {}'.format(org_module))

        if self.is_gdn:
            self.weight, self.bias = org_module.gamma, org_module.beta
        elif self.is_ps:
            self.weight = self.bias = None
        else:
            self.weight, self.bias = org_module.weight, org_module.bias
        self.org_weight = None if self.weight is None else self.weight.data.clone()
        self.org_bias = None if self.bias is None else self.bias.data.clone()

        self.use_weight_quant = False
        self.use_act_quant = False
        self.disable_act_quant = disable_act_quant
        self.weight_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **weight_quant_params)
        self.act_quantizer = UniformAffineQuantizer(tconv=self.if_tconv, **act_quant_params)
        self.activation_function = nn.LeakyReLU(inplace=True) if self.is_ps else StraightThrough()   # reference :100
        self.ignore_reconstruction = False
        self.se_module = se_module
        self.trained = False
        self.last_k_taps = 0
        self._repr = org_module.extra_repr()

    def extra_repr(self):
        return self._repr

    # -- prepared weight operand (evaluation) ------------------------------------------------------------------------
    # The reference re-quantises every weight on every forward (quant_layer.py:113-115).  Without a gradient the weight is
    # a constant of the layer, so its tensor-core operand -- quantiser + packing in one kernel
    # (b200lic_quant_pack_weights); for GDN: quantise gamma, re-parametrise, pack -- is built once and reused until
    # anything it depends on changes (quantiser object, its state, alpha / delta / weight versions).
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_prep", None)                 # device scratch, not part of the pickled QuantModel (main2.py:288)
        state.pop("_defer_to", None)             # wiring hint of QuantModel (re-derived by link_sequential_consumers)
        state.pop("_stat_keys", None)
        return state

    def _prep_key(self):
        q = self.weight_quantizer

        def ver(t):
            return None if t is None else (t.data_ptr(), t._version)
        bias = self.bias if self.use_weight_quant else self.org_bias
        return (self.use_weight_quant, id(q), getattr(q, "inited", None), getattr(q, "n_levels", None),
                getattr(q, "soft_targets", None), ver(self.weight), ver(self.org_weight), ver(getattr(q, "alpha", None)),
                ver(getattr(q, "delta", None)), ver(getattr(q, "zero_point", None)), ver(bias), ops.DEFAULT_ENGINE)

    def _prepared(self, d, fold=1):
        """(packed weight operand, w_scale or None, bias) for descriptor `d`, or None when this call has to take the
        general path (quantiser not initialised yet, layer without a packed operand).  fold = k*k when `d` is the 1x1
        form of a folded-tap transposed conv (ops.folded_deconv_desc): same weight memory, taps folded into the output
        channel axis."""
        q = self.weight_quantizer
        # a stand-in quantiser that returns a constant weight (the streaming session's nearest-rounded weights,
        # session._FixedWeight): `fixed_weight` = (dequantised weight, (n_int, scale) or None)
        fixed = getattr(q, "fixed_weight", None) if self.use_weight_quant else None
        if self.use_weight_quant and fixed is None and (
                not isinstance(q, (UniformAffineQuantizer, AdaRoundQuantizer))
                or getattr(q, "_leaf", None) is not None or getattr(q, "soft_targets", False)
                or (hasattr(q, "inited") and not q.inited) or q.delta is None):
            return None
        key = self._prep_key()
        slots = self.__dict__.setdefault("_prep", {})          # one operand per quant state: FP and quantised forwards
        ent = slots.get(self.use_weight_quant)                 # alternate while the calibration caches are built
        if ent is not None and ent[0] == key:
            return ent[1]
        if torch.cuda.is_current_stream_capturing():
            return None                # operands are built eagerly (a host read decides the integer form); general path
        dev = self.weight.device
        if ops.packed_weight_bytes(d, self.if_tconv) == 0:
            val = None
        elif self.is_gdn:
            gamma = q(self.weight).detach() if self.use_weight_quant else self.org_weight
            beta = self.bias if self.use_weight_quant else self.org_bias
            g_eff = self.fwd_kwargs["gamma_reparam"](gamma).detach()
            b_eff = self.fwd_kwargs["beta_reparam"](beta).detach()
            val = (ops.pack_weights(g_eff.reshape(g_eff.shape[0], g_eff.shape[1], 1, 1), d, False), None, b_eff)
        elif not self.use_weight_quant:
            w = self.org_weight if fold == 1 else self.org_weight.reshape(self.org_weight.shape[0], -1, 1, 1)
            val = (ops.pack_weights(w, d, self.if_tconv), None, self.org_bias)
        elif fixed is not None:
            # same operands as the general path builds per call (ops.conv_wq on the integer weights, or the dequantised
            # weight when there is no integer form), packed once
            w, iw = fixed
            scale = None
            if iw is not None:
                w, scale = iw
                if fold != 1:
                    scale = scale.repeat_interleave(fold).contiguous()
            if fold != 1:
                w = w.reshape(w.shape[0], -1, 1, 1)
            val = (ops.pack_weights(w, d, self.if_tconv), scale, self.bias)
        else:
            axis = q.axis if hasattr(q, "axis") else q.channel_axis(self.weight)
            alpha = q.alpha.detach() if hasattr(q, "alpha") else None
            out_axis = 1 if self.if_tconv else 0
            # two-pass integer form: n = code - zero_point must be exact in bf16 (|n| <= 256) and the scale must sit on
            # the output channels; a searched range that excludes 0 puts zero_point outside [0, levels-1]
            integer = (q.n_levels <= 256 and axis in (None, out_axis) and
                       bool(((q.zero_point >= 0) & (q.zero_point <= q.n_levels - 1)).all()))
            packed = ops.quant_pack_weights(self.weight.detach(), alpha, q.delta, q.zero_point, axis, q.n_levels, False,
                                            d, self.if_tconv, integer_mode=integer)
            scale = None
            if integer:
                cout = self.weight.shape[out_axis]
                scale = q.delta.reshape(-1)
                scale = (scale.expand(cout) if scale.numel() == 1 else scale).contiguous()
                if fold != 1:
                    scale = scale.repeat_interleave(fold).contiguous()
            val = (packed, scale, self.bias)
        if val is not None:
            # masked context convolution: when every tap behind the mask is exactly zero in the weight this operand was
            # built from (always, unless AdaRound un-masked one: alpha of a masked tap is trainable), the
            # engine contracts the live prefix only (b200lic_conv_desc::k_taps)
            k_taps, L = 0, getattr(self, "mask_live_taps", 0)
            if MASKED_TAPS and L and not self.is_gdn and not self.if_tconv and fold == 1:
                w_eff = self.weight_quantizer(self.weight).detach() if self.use_weight_quant else self.org_weight
                if bool((w_eff.reshape(w_eff.shape[0], w_eff.shape[1], -1)[:, :, L:] == 0).all()):
                    k_taps = L
            val = val + (k_taps,)
        slots[self.use_weight_quant] = (key, val)
        return val

    def _stats_wanted(self):
        """The conditions of the deferred activation quantiser that are known BEFORE the layer runs: when they hold, the
        convolution's epilogue takes the per-channel statistics of its own output (ops.conv_stats_arm)."""
        nxt = self.__dict__.get("_defer_to")
        return (ops.FUSE_ACTQ_STATS and ops.DEFER_ACTQ and not self.is_ps
                and not self.disable_act_quant and self.use_act_quant and self.trained and nxt is not None
                and not torch.is_grad_enabled() and not nxt._forward_hooks and not nxt._forward_pre_hooks
                and not self._forward_hooks)

    def _arm_stats(self, fuse, d=None, input=None):
        """Right before the launch of this module's convolution (nothing else may run on the engine in between).  Only
        for outputs the deferred quantiser will actually take (ops.DEFER_ACTQ_MIN_BYTES): smaller ones go through the
        one-launch cluster quantiser, which takes its own statistics."""
        if not fuse:
            return
        if d is None:
            kw = self.fwd_kwargs
            d = ops.conv_desc(input.shape, self.weight.shape, kw["stride"], kw["padding"], self.if_tconv,
                              kw.get("output_padding", 0))
        if 4 * d.N * d.Cout * d.Ho * d.Wo >= ops.DEFER_ACTQ_MIN_BYTES:
            self._stat_keys = ops.conv_stats_arm(d.Cout, self.weight.device)

    def _forward_prepared(self, input, act, slope, fuse=False):
        if self.is_gdn:
            d = ops.gdn_desc(input.shape, self.fwd_kwargs["inverse"])
        else:
            kw = self.fwd_kwargs
            if ops._sq(kw["dilation"], "dilation") != 1 or kw["groups"] != 1:
                return None
            d = ops.conv_desc(input.shape, self.weight.shape, kw["stride"], kw["padding"], self.if_tconv,
                              kw.get("output_padding", 0), act=act, slope=slope)
        if d.engine == ops.ENGINE_SIMT:
            return None
        d1 = ops.folded_deconv_desc(d) if (self.if_tconv and not self.is_gdn) else None
        if d1 is not None:
            return self._forward_folded_deconv(input, d, d1, act, slope)
        ws = ops._workspace(d, ops.fwd_op(self.if_tconv), input.device)
        if ws[0] is None:                        # shape the tensor-core path rejects: the general path picks the engine
            return None
        prep = self._prepared(d)
        if prep is None:
            return None
        packed, scale, bias, k_taps = prep
        d.k_taps = k_taps
        self.last_k_taps = k_taps                # what the last prepared forward contracted (0 = all taps; tests)
        x = ops._c(input, "input")
        pend = getattr(input, "_b200_actq", None)
        if self.is_gdn and (pend is None or pend[1] <= 8) and ops.gdn_fused_ok(x.shape[1], x.shape[2] * x.shape[3]):
            # one kernel over the raw tensor: quantiser (if deferred to us), square, GEMM and epilogue on chip
            self._arm_stats(fuse and act == ops.ACT_NONE, d)
            out = ops.gdn_fwd_fused(x, packed, bias, self.fwd_kwargs["inverse"], pending=pend)
            return ops.add_act(out, None, act, slope) if act != ops.ACT_NONE else out
        if pend is not None:
            # the producer deferred its activation quantiser to us: quantise while staging our own operand
            slot = ops.conv_x_slot(d, self.if_tconv, ws)
            if slot is None:
                x, pend = ops.act_quant_apply(x, pend[0], pend[1]), None
            else:
                xq = torch.empty_like(x) if self.is_gdn else None           # GDN's epilogue reads the quantised x itself
                ops.act_quant_apply_stage(x, pend[0], pend[1], slot, square=self.is_gdn, out=xq)
                if self.is_gdn:
                    out = ops.conv_fwd_packed(None, packed, d, False, bias=bias, gdn_x=xq, ws=ws)
                    return ops.add_act(out, None, act, slope) if act != ops.ACT_NONE else out
                self._arm_stats(fuse, d)
                return ops.conv_fwd_packed(None, packed, d, self.if_tconv, bias=bias, w_scale=scale, ws=ws)
        if self.is_gdn:
            out = ops.conv_fwd_packed(x, packed, d, False, bias=bias, gdn_x=x, ws=ws)
            return ops.add_act(out, None, act, slope) if act != ops.ACT_NONE else out
        self._arm_stats(fuse, d)
        return ops.conv_fwd_packed(x, packed, d, self.if_tconv, bias=bias, w_scale=scale, ws=ws)

    def _forward_folded_deconv(self, input, d, d1, act, slope):
        """N -> 3 synthesis layer without a gradient: the 1x1 GEMM on prepared operands (a deferred quantiser of the input is
        applied while staging it) + the col2im gather with bias / activation."""
        ws = ops._workspace(d1, ops.fwd_op(True), input.device)
        if ws[0] is None:
            return None
        prep = self._prepared(d1, fold=d.KH * d.KW)
        if prep is None:
            return None
        packed, scale, bias, _ = prep
        x = ops._c(input, "input")
        pend = getattr(input, "_b200_actq", None)
        if pend is not None:
            slot = ops.conv_x_slot(d1, True, ws)
            if slot is None:
                x, pend = ops.act_quant_apply(x, pend[0], pend[1]), None
            else:
                ops.act_quant_apply_stage(x, pend[0], pend[1], slot)
                x = None
        col = ops.conv_fwd_packed(x, packed, d1, True, w_scale=scale, ws=ws)
        return ops.col2im(col, bias, d.Cout, d.KH, d.KW, d.stride, d.pad, d.Ho, d.Wo, act, slope)

    def forward(self, input: torch.Tensor, act_override=None):
        """act_override = (act id, slope): an activation the CALLER applies right after this module (the hand-written
        Cheng2020 blocks, quant_block.py) folded into this module's epilogue; same values, one launch less."""
        act, slope = act_override if act_override is not None else ops._act_id(self.activation_function)
        if self.is_ps:
            return ops.pixel_shuffle(input, self.fwd_kwargs, act, slope)
        if self.is_linear or self.if_layer_norm:
            return self._forward_tokens(input)
        out = None
        fuse = self._stats_wanted()
        self._stat_keys = None
        if not torch.is_grad_enabled():
            out = self._forward_prepared(input, act, slope, fuse)
        if out is None:
            input = ops.resolve_actq(input)      # general path: a deferred activation quantisation is materialised
            # (transposed convs may run as a folded GEMM + col2im there: its GEMM output is not the layer output)
            fuse = fuse and not self.if_tconv and not self.is_gdn
        if out is None and self.use_weight_quant and not self.is_gdn and not torch.is_grad_enabled():
            # hard-quantised weight, no gradient wanted (evaluation): integer weights + per-channel scale in the conv
            # epilogue, two tensor-core passes instead of three (b200lic_conv_fwd_wq); same value up to fp32 rounding
            iw = getattr(self.weight_quantizer, "int_weights", lambda _w: None)(self.weight)
            if iw is not None:
                self._arm_stats(fuse, input=input)
                out = ops.conv_wq(input, iw[0], iw[1], self.bias, transposed=self.if_tconv, act=act, slope=slope,
                                  **self.fwd_kwargs)
                if out is None and self._stat_keys is not None:
                    ops.conv_stats_taken()       # no launch happened: disarm
                    self._stat_keys = None
        if out is not None:
            weight = bias = None
        elif self.use_weight_quant:
            weight, bias = self.weight_quantizer(self.weight), self.bias
        else:
            weight, bias = self.org_weight, self.org_bias
        if out is not None:
            pass
        elif self.is_gdn:
            out = self.fwd_func(input, weight, bias, **self.fwd_kwargs)
            if act != ops.ACT_NONE:
                out = ops.add_act_fn(out, None, act, slope)
        else:
            if self._stat_keys is None:
                self._arm_stats(fuse, input=input)
            out = self.fwd_func(input, weight, bias, act=act, slope=slope, **self.fwd_kwargs)
        keys = None
        if self._stat_keys is not None:          # armed: did a tensor-core launch take the statistics?
            keys = self._stat_keys if ops.conv_stats_taken() else None
            self._stat_keys = None
        if self.se_module is not None:
            raise NotImplementedError("se_module is never set on the LIC hot path")
        if self.disable_act_quant:
            return out
        if self.use_act_quant and self.trained:
            nxt = self.__dict__.get("_defer_to")
            if (ops.DEFER_ACTQ and nxt is not None and not torch.is_grad_enabled() and out.dim() == 4
                    and 4 * out.numel() >= ops.DEFER_ACTQ_MIN_BYTES and not nxt._forward_hooks and not nxt._forward_pre_hooks and not self._forward_hooks):
                # evaluation inside an nn.Sequential: the only reader is the next QuantModule, which applies the
                # quantiser while staging its operand (ops.DEFER_ACTQ); the statistics are taken here
                bits = self.act_quantizer.n_bits if self.act_quantizer.act_bits_follow_n_bits else 8
                out = out.detach()
                out._b200_actq = (keys if keys is not None else ops.act_quant_stats(out), bits)
                return out
            out = self.act_quantizer(out, True)
        return out

    def _forward_tokens(self, input):
        """nn.Linear / nn.LayerNorm over token tensors (quant_layer.py:113-134 with fwd_func = F.linear / F.layer_norm):
        the GEMM of a Linear runs on the tcgen05 conv engine (ops.linear), LayerNorm and the token-major activation
        quantiser are one kernel each.  Values only: the reference trains alpha of these layers inside QuantRSTB block
        reconstruction, which is not built."""
        if torch.is_grad_enabled() and (input.requires_grad or any(
                p.requires_grad for p in self.weight_quantizer.parameters())):
            raise NotImplementedError("Linear / LayerNorm wrappers are forward-only (SURVEY.md 8(f) N4)")
        if self.use_weight_quant:
            weight, bias = self.weight_quantizer(self.weight).detach(), self.bias
        else:
            weight, bias = self.org_weight, self.org_bias
        if self.is_linear:
            out = ops.linear(input.detach(), weight, bias)
        else:
            out = ops.layer_norm(input.detach(), weight, bias, self.fwd_kwargs["eps"])
        af = self.activation_function                  # an activation QuantModel absorbed from the nn.Sequential (:51-56)
        if isinstance(af, nn.GELU) or type(af).__name__ == "GELU":
            out = ops.gelu(out)
        elif not isinstance(af, StraightThrough):
            act, slope = ops._act_id(af)
            if act != ops.ACT_NONE:
                out = ops.add_act(out, None, act, slope)
        if self.disable_act_quant:
            return out
        if self.use_act_quant and self.trained:
            out = self.act_quantizer(out, True)
        return out

    def invalidate_prepared(self):
        """Drop the cached weight operand.  Version counters catch in-place torch ops on weight / alpha / delta; writes
        through `.data` or by a kernel (the AdaRound loop) do not bump them, so the AdaRound loop calls this when it hardens a unit."""
        self.__dict__.pop("_prep", None)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
