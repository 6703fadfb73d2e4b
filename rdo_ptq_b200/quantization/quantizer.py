"""B200 drop-in for task-oriented-PTQ/quantization/quantizer.py.

Same public names and state (`UniformAffineQuantizer.delta/zero_point/n_bits/n_levels/sym/inited`,
`AdaRoundQuantizer.alpha/soft_targets/get_soft_targets`), but the arithmetic runs in libb200lic
kernels: K7 (range + fake-quant), K6 (AdaRound) and K8 (dynamic activation quant).  No CPU path.
"""
import torch
import torch.nn as nn

from .. import ops


class StraightThrough(nn.Module):
    """reference quantizer.py:12-17."""

    def __init__(self, channel_num: int = 1):
        super().__init__()

    def forward(self, input):
        return input


def round_ste(x: torch.Tensor):
    """reference quantizer.py:64-68.  Only the calibration surface uses it (never the fused loop)."""
    return x + (ops.round_latent(x) - x).detach()


def lp_loss(pred, tgt, p=2.0, reduction='none'):
    """reference quantizer.py:71-79.  Differentiable w.r.t. `pred` when it requires a gradient (one K11 pass gives value
    and gradient); the fused loop calls ops.lp_loss_fwd_bwd directly."""
    denom = pred.numel() // pred.shape[1] if reduction == 'none' else pred.numel()
    if torch.is_grad_enabled() and pred.requires_grad:
        return ops.lp_loss_fn(pred, tgt.detach(), p=p, scale=1.0 / denom).reshape(())
    loss, _ = ops.lp_loss_fwd_bwd(pred.detach(), tgt.detach(), p=p, scale=1.0 / denom, want_grad=False)
    return loss.reshape(())


def ActQuantizer(x: torch.Tensor, n_bits: int = 8):
    """reference quantizer.py:99-121: dynamic per-channel fake-quant, detached.  The reference hard-wires 8 bit
    (`Handle_Parameter(b_w=8)`); `n_bits` is the additive knob BASELINE config 4 (W10A10) needs.
    4-D: per channel of axis 1; 3-D (token tensors of the Swin blocks, quantizer.py:107-109): per channel of the LAST axis."""
    if x.dim() == 3:
        return ops.act_quant_tokens(x, n_bits)
    return ops.act_quant(x, n_bits)


def ActQuant(x: torch.Tensor):
    return ActQuantizer(x)


class _FakeQuantSTE(torch.autograd.Function):
    """(clamp(rint(x/d)+zp, 0, L-1) - zp) * d with the straight-through gradient of quantizer.py:175-177."""

    @staticmethod
    def forward(ctx, x, delta, zp, axis, n_levels):
        ctx.save_for_backward(x, delta, zp)
        ctx.n_levels = n_levels
        return ops.wq_fake_quant(x, delta, zp, axis, n_levels, want=("dq",))

    @staticmethod
    def backward(ctx, g):          # off the hot path: nothing in RDO-PTQ trains through this quantizer
        x, delta, zp = ctx.saved_tensors
        xi = torch.round(x / delta) + zp
        return g * ((xi >= 0) & (xi <= ctx.n_levels - 1)).to(g.dtype), None, None, None, None


class UniformAffineQuantizer(nn.Module):
    """reference quantizer.py:123-393.  Scale methods: 'max' / 'max_scale' (the default `--init max`, K7a) and the
    search- / moment-based 'mse', 'l1', 'l2', 'gaussian' (quantizer.py:300-370, K7c), all per channel on the GPU."""

    # additive switch: thread n_bits into the activation quantiser (config 4, W10A10); default = reference (8 bit)
    act_bits_follow_n_bits = False

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False, scale_method: str = 'max',
                 leaf_param: bool = False, tconv: bool = False, act: bool = False, prob: float = 1.0):
        super().__init__()
        self.sym = symmetric
        assert 2 <= n_bits <= 16, 'bitwidth not supported'      # reference asserts <= 8; lifted for W10A10 (Q6)
        self.n_bits = n_bits
        self.n_levels = 2 ** self.n_bits
        self.delta = None
        self.zero_point = None
        self.inited = False
        self.leaf_param = leaf_param
        self.channel_wise = channel_wise
        self.scale_method = scale_method
        self.tconv = tconv
        self.act = act
        self.prob = prob
        self.is_training = False

    def channel_axis(self, x):
        """Axis that carries per-channel scales (quantizer.py:237-279); None = per tensor."""
        if not self.channel_wise or x.dim() == 1:
            return None
        return 1 if (self.tconv and x.dim() == 4) else 0

    def forward(self, x: torch.Tensor, act: bool = False):
        if act:
            return ActQuantizer(x, self.n_bits if self.act_bits_follow_n_bits else 8)
        if self.inited is False:
            if self.leaf_param:
                return x
            self.delta, self.zero_point = self.init_quantization_scale(x, self.channel_wise)
            self.inited = True
        return _FakeQuantSTE.apply(x, self.delta, self.zero_point, self.channel_axis(x), self.n_levels)

    def codes(self, x):
        """Integer codes (fp32-valued) of the forward above: the bit-exact contract."""
        return ops.wq_fake_quant(x.detach(), self.delta, self.zero_point, self.channel_axis(x), self.n_levels,
                                 want=("codes",))

    def int_weights(self, x):
        """(code - zero_point, delta per output channel) of a conv / transposed-conv weight, or None when the two-pass
        integer-weight forward does not apply (not initialised, more than 256 levels: |n| must be exact in bf16)."""
        if not self.inited or self.n_levels > 256 or x.dim() != 4 or self.delta is None:
            return None
        if not _zp_in_grid(self):
            return None
        return _int_weights(x, self.delta, self.zero_point, self.channel_axis(x), self.n_levels, self.tconv, None)

    def init_quantization_scale(self, x: torch.Tensor, channel_wise: bool = False):
        axis = self.channel_axis(x) if channel_wise else None
        if 'max' in self.scale_method:
            delta, zp = ops.wq_init_minmax(x.detach(), axis, self.n_bits, 'scale' in self.scale_method, self.sym)
        elif self.scale_method in ops.SCALE_METHODS:      # 10 shrink steps of 5 %, L3.5 / L1 / L2 score; mean -+ 6 var
            delta, zp = ops.wq_init_search(x.detach(), axis, self.n_bits, self.scale_method, 10, 0.05, 3.5, self.sym)
        else:
            raise NotImplementedError(f"scale_method {self.scale_method!r}")        # as quantizer.py:372
        if channel_wise and x.dim() == 1:
            delta, zp = delta.view(-1), zp.view(-1)
        return delta, zp

    def bitwidth_refactor(self, refactored_bit: int):
        assert 2 <= refactored_bit <= 16, 'bitwidth not supported'
        self.n_bits = refactored_bit
        self.n_levels = 2 ** self.n_bits

    def extra_repr(self):
        return (f'bit={self.n_bits}, scale_method={self.scale_method}, symmetric={self.sym}, '
                f'channel_wise={self.channel_wise}, leaf_param={self.leaf_param}')


def _zp_in_grid(q):
    """n = code - zero_point is exact in bf16 only for |n| <= 256, i.e. 0 <= zero_point <= levels - 1.  'max' ranges
    always contain 0; the searched ranges ('mse' / 'l1' / 'l2') need not (reference quantizer.py:300-370), and a
    single-signed channel then gets a zero point far outside the grid.  One device read per (delta, zero_point) pair."""
    key = (q.zero_point.data_ptr(), q.zero_point._version, q.n_levels)
    cached = q.__dict__.get("_zp_ok")
    if cached is None or cached[0] != key:
        if torch.cuda.is_current_stream_capturing():
            return False                  # no host read inside a graph capture: take the always-valid three-pass form
        ok = bool(((q.zero_point >= 0) & (q.zero_point <= q.n_levels - 1)).all())
        q.__dict__["_zp_ok"] = cached = (key, ok)
    return cached[1]


def _int_weights(x, delta, zp, axis, n_levels, tconv, alpha):
    cout = x.shape[1] if tconv else x.shape[0]
    if axis is not None and axis != (1 if tconv else 0):
        return None                                   # scale not along the output channels
    n = ops.wq_int_weights(x.detach(), delta, zp, axis, n_levels, alpha)
    scale = delta.reshape(-1)
    if scale.numel() == 1:
        scale = scale.expand(cout)
    return n, scale.contiguous()


class _AdaRoundFn(torch.autograd.Function):
    """AdaRound forward with gradient to alpha (autograd surface for reference-style training loops;
    the fused calibration loop bypasses it and calls K6 bwd+Adam directly)."""

    @staticmethod
    def forward(ctx, x, alpha, q):
        ctx.q = q
        ctx.save_for_backward(x, alpha)
        return ops.adaround_fwd(x, alpha, q.delta, q.zero_point, q.axis, q.n_levels, True)

    @staticmethod
    def backward(ctx, g):
        x, alpha = ctx.saved_tensors
        q = ctx.q
        d_alpha = torch.empty_like(alpha)
        ops.adaround_bwd_adam(x, alpha, q.delta, q.zero_point, g.contiguous(), None, None, q.axis, q.n_levels, 1,
                              d_alpha_out=d_alpha)
        return None, d_alpha, None


class AdaRoundQuantizer(nn.Module):
    """reference quantizer.py:397-470 ('learned_hard_sigmoid')."""

    def __init__(self, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor, round_mode='learned_round_sigmoid'):
        super().__init__()
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta
        self.zero_point = uaq.zero_point
        self.n_levels = uaq.n_levels
        self.axis = uaq.channel_axis(weight_tensor)
        self.tconv = uaq.tconv
        self.round_mode = round_mode
        self.alpha = None
        self.soft_targets = False
        self.gamma, self.zeta = -0.1, 1.1
        self.beta = 2 / 3
        self._leaf = None          # soft-quantised weight materialised by the fused loop (leaf for dWq)
        self.init_alpha(x=weight_tensor)

    def init_alpha(self, x: torch.Tensor):
        if self.round_mode != 'learned_hard_sigmoid':
            raise NotImplementedError
        self.alpha = nn.Parameter(ops.adaround_init_alpha(x.detach(), self.delta, self.axis))

    def get_soft_targets(self):
        """h(alpha) = clamp(sigmoid(alpha)*(zeta-gamma)+gamma, 0, 1) (quantizer.py:451-452).  Compatibility
        surface only: the fused loop evaluates h and the regulariser inside K6."""
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def forward(self, x):
        if self.round_mode != 'learned_hard_sigmoid':
            raise ValueError('Wrong rounding mode')
        if self._leaf is not None:
            return self._leaf
        if self.soft_targets:
            if torch.is_grad_enabled() and self.alpha.requires_grad:
                return _AdaRoundFn.apply(x, self.alpha, self)
            return ops.adaround_fwd(x.detach(), self.alpha.detach(), self.delta, self.zero_point, self.axis,
                                    self.n_levels, True)
        return ops.adaround_fwd(x.detach(), self.alpha.detach(), self.delta, self.zero_point, self.axis, self.n_levels,
                                False)

    def codes(self, x):
        return ops.adaround_fwd(x.detach(), self.alpha.detach(), self.delta, self.zero_point, self.axis, self.n_levels,
                                False, want_codes=True)[1]

    def int_weights(self, x):
        """See UniformAffineQuantizer.int_weights; only the hardened quantiser has integer weights."""
        if self.soft_targets or self._leaf is not None or self.n_levels > 256 or x.dim() != 4:
            return None
        if not _zp_in_grid(self):
            return None
        return _int_weights(x, self.delta, self.zero_point, self.axis, self.n_levels, getattr(self, "tconv", False),
                            self.alpha.detach())

    def extra_repr(self):
        return 'bit={}'.format(self.n_bits)
