"""Oracle restatement of the reference quantizers (test infrastructure, CPU fp32).

Follows, function by function:
  TO = /root/reference/task-oriented-PTQ/quantization/quantizer.py
  LU = /root/reference/light-uniform-PTQ/quant_int/quantizer.py
Every function cites the lines it restates.  `oracle/make_golden.py` checks each one
bit-for-bit against the imported reference module.
"""
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

ZETA, GAMMA = 1.1, -0.1          # TO quantizer.py:423 (rectified-sigmoid stretch)


def round_ste(x: torch.Tensor) -> torch.Tensor:
    """TO quantizer.py:64-68 / LU :63-67 -- value rint(x) (half-to-even), identity gradient."""
    return x + (torch.round(x) - x).detach()


def lp_loss(pred, tgt, p=2.0, reduction="none"):
    """TO quantizer.py:71-79: sum over dim 1, mean over the rest ('none'); plain mean otherwise."""
    e = (pred - tgt).abs().pow(p)
    return e.sum(1).mean() if reduction == "none" else e.mean()


# --------------------------------------------------------------------------------------
# dynamic per-channel activation fake-quant (TO quantizer.py:81-121)
# --------------------------------------------------------------------------------------
def handle_parameter(t: torch.Tensor, n_bits: int = 8) -> torch.Tensor:
    """TO quantizer.py:81-96 on one slice: m=min, r=max(max|t-m|,1e-6), q=rint(clamp((t-m)/r,-1,1)*L),
    out=(q/L)*r+m.  The reference hard-wires n_bits=8; n_bits is threaded for W10A10."""
    levels = float(2 ** n_bits - 1)
    m = t.min()
    s = t - m
    r = torch.max(s.abs().max(), torch.tensor(1e-6, dtype=torch.float32))
    q = torch.round(torch.clamp(s / r, -1, 1) * levels)
    return (q / levels) * r + m


def act_quant_loop(x: torch.Tensor, n_bits: int = 8) -> torch.Tensor:
    """TO quantizer.py:99-117 exactly as shipped: a Python loop over channels ("verbatim" CPU timing)."""
    y = x.detach().clone()
    if y.dim() == 4:
        for c in range(y.shape[1]):
            y[:, c] = handle_parameter(y[:, c], n_bits)
    elif y.dim() == 3:
        for c in range(y.shape[2]):
            y[:, :, c] = handle_parameter(y[:, :, c], n_bits)
    elif y.dim() == 2:
        for c in range(y.shape[1]):
            y[:, c] = handle_parameter(y[:, c], n_bits)
    else:
        y = handle_parameter(y, n_bits)
    return y


def act_quant(x: torch.Tensor, n_bits: int = 8, return_codes: bool = False):
    """Vectorised-equivalent of TO quantizer.py:99-117 (same per-element fp32 op order)."""
    levels = float(2 ** n_bits - 1)
    y = x.detach()
    if y.dim() == 4:
        dims = (0, 2, 3)
    elif y.dim() == 3:
        dims = (0, 1)
    elif y.dim() == 2:
        dims = (0,)
    else:
        dims = tuple(range(y.dim()))
    m = y.amin(dim=dims, keepdim=True)
    s = y - m
    r = torch.clamp_min(s.abs().amax(dim=dims, keepdim=True), 1e-6)
    q = torch.round(torch.clamp(s / r, -1, 1) * levels)
    out = (q / levels) * r + m
    return (out, q) if return_codes else out


def lu_act_quantizer(x: torch.Tensor, a_l: int = 8, a_r: int = 8) -> torch.Tensor:
    """LU quantizer.py:120-128: static Q(a_l).(a_r) fixed point."""
    lo, hi, mult = -(2 ** (a_l - 1)), 2 ** (a_l - 1), 2 ** a_r
    return torch.round(torch.clamp(x, lo, hi) * mult) / mult


# --------------------------------------------------------------------------------------
# per-channel asymmetric weight range (TO quantizer.py:233-298, scale_method 'max')
# --------------------------------------------------------------------------------------
def channel_axis(shape, tconv: bool) -> Optional[int]:
    """Which axis carries the per-channel scale: TO quantizer.py:237-279."""
    if len(shape) == 1:
        return None                       # :257-258 1-D tensors fall back to per-tensor
    return 1 if tconv else 0              # :260-265 (tconv applies only to 4-D weights in practice)


def minmax_scale(x_min: float, x_max: float, n_levels: int, n_bits: int, method: str, sym: bool
                 ) -> Tuple[float, float]:
    """TO quantizer.py:281-298 for one slice.  x_min/x_max arrive as Python floats (`.item()`), the
    range arithmetic is float64, delta is rounded to fp32 by `torch.tensor`, and
    `(-x_min / delta)` is `delta.reciprocal() * (-x_min)` in fp32 (torch `Tensor.__rdiv__`)."""
    x_min, x_max = min(x_min, 0.0), max(x_max, 0.0)
    if "scale" in method:
        x_min = x_min * (n_bits + 2) / 8
        x_max = x_max * (n_bits + 2) / 8
    if sym:
        a = max(abs(x_min), x_max)
        x_min, x_max = (-a if x_min < 0 else 0), a
    delta = torch.tensor((x_max - x_min) / (n_levels - 1), dtype=torch.float32)
    delta = torch.max(delta, torch.tensor(1e-8, dtype=torch.float32))
    zp = torch.round(delta.reciprocal() * torch.tensor(-x_min, dtype=torch.float32))
    return float(delta), float(zp)


class UniformAffineQuantizer(nn.Module):
    """TO quantizer.py:123-393 (weights) restated; scale methods 'max', 'max_scale', 'mse', 'l1', 'l2',
    'gaussian' (all the reference has).  `n_bits` up to 16 is accepted (Q6: the reference
    asserts <= 8, config 4 needs 10)."""

    # Extension switch (not in the reference): thread n_bits into the dynamic activation quantiser, which the
    # reference hard-wires to 8 bit (quantizer.py:81-121); BASELINE config 4 (W10A10) sets it.  Default = reference.
    act_bits_follow_n_bits = False
    # CPU-baseline switch: run the dynamic activation quantiser as the reference ships it (a Python loop over the
    # channels, quantizer.py:99-117) instead of the vectorised equivalent; the two are bit-identical (make_golden.py)
    act_verbatim_loop = False

    def __init__(self, n_bits=8, symmetric=False, channel_wise=False, scale_method="max",
                 leaf_param=False, tconv=False, act=False, prob=1.0):
        super().__init__()
        assert 2 <= n_bits <= 16
        self.sym, self.n_bits, self.n_levels = symmetric, n_bits, 2 ** n_bits
        self.delta = self.zero_point = None
        self.inited = False
        self.leaf_param, self.channel_wise, self.scale_method = leaf_param, channel_wise, scale_method
        self.tconv, self.act, self.prob, self.is_training = tconv, act, prob, False

    def bitwidth_refactor(self, bits):                     # :385-388
        self.n_bits, self.n_levels = bits, 2 ** bits

    def _slice_scale(self, t: torch.Tensor):
        if "max" in self.scale_method:
            return minmax_scale(t.min().item(), t.max().item(), self.n_levels, self.n_bits,
                                self.scale_method, self.sym)
        if self.scale_method in ("mse", "l1", "l2"):      # :300-316, :339-370
            score_fn = {"mse": lambda a, b: lp_loss(a, b, p=3.5, reduction="all"),
                        "l1": torch.nn.functional.l1_loss, "l2": torch.nn.functional.mse_loss}[self.scale_method]
            return self._shrink_search(t, 10, 0.05, score_fn)
        if self.scale_method == "gaussian":               # :318-336 (mean -+ 6 * unbiased variance, fp32 tensors)
            mu, var = torch.mean(t), torch.var(t)
            lo, hi = torch.clamp_max(mu - 6 * var, 0), torch.clamp_min(mu + 6 * var, 0)
            if self.sym:
                a = torch.max(lo.abs(), hi)
                lo, hi = (-a if lo < 0 else torch.zeros(())), a
            d = torch.max((hi - lo) / (self.n_levels - 1), torch.tensor(1e-8))
            return float(d), float((-lo / d).round())
        raise NotImplementedError(self.scale_method)

    def _shrink_search(self, t, steps, shrink, score_fn):
        """:300-316 / LU quantizer.py:265-278: shrink (max, min) step by step, keep the first strictly best score."""
        best, out = 1e10, None
        hi, lo = t.max(), t.min()
        eps = torch.tensor(1e-8)
        for i in range(steps):
            nh, nl = hi * (1.0 - i * shrink), lo * (1.0 - i * shrink)
            d = torch.max((nh - nl) / (2 ** self.n_bits - 1), eps)
            z = (-nl / d).round()
            tq = (torch.clamp(torch.round(t / d) + z, 0, self.n_levels - 1) - z) * d      # :375-382
            score = score_fn(t, tq)
            if score < best:
                best, out = score, (float(d), float(z))
        return out

    def init_quantization_scale(self, x: torch.Tensor, channel_wise: bool = False):
        """:233-298.  Returns (delta, zero_point) shaped for broadcasting against x."""
        x = x.detach()
        ax = channel_axis(x.shape, self.tconv) if channel_wise else None
        if ax is None:
            d, z = self._slice_scale(x)
            d, z = torch.tensor(d, dtype=x.dtype), torch.tensor(z, dtype=x.dtype)
            if channel_wise:                                # 1-D: .view(-1) (:274-276)
                d, z = d.view(-1), z.view(-1)
            return d.to(x.device), z.to(x.device)            # (`type_as(x)` in the reference also moves to x's device)
        n = x.shape[ax]
        d = torch.empty(n, dtype=x.dtype)
        z = torch.empty(n, dtype=x.dtype)
        for c in range(n):
            d[c], z[c] = self._slice_scale(x.select(ax, c))
        shape = [1] * x.dim()
        shape[ax] = n
        if x.dim() != 4:                                    # :277-279 2-D/3-D -> [n,1]
            shape = [n, 1]
        return d.view(shape).to(x.device), z.view(shape).to(x.device)

    def forward(self, x: torch.Tensor, act: bool = False):
        """:156-184."""
        if act:
            fn = act_quant_loop if self.act_verbatim_loop else act_quant
            return fn(x, self.n_bits if self.act_bits_follow_n_bits else 8)
        if not self.inited:
            if self.leaf_param:
                return x
            self.delta, self.zero_point = self.init_quantization_scale(x, self.channel_wise)
            self.inited = True
        x_int = round_ste(x / self.delta) + self.zero_point
        x_q = torch.clamp(x_int, 0, self.n_levels - 1)
        return (x_q - self.zero_point) * self.delta

    def codes(self, x):
        """Integer codes of the forward above (the bit-exact contract)."""
        return torch.clamp(torch.round(x / self.delta) + self.zero_point, 0, self.n_levels - 1)


class AdaRoundQuantizer(nn.Module):
    """TO quantizer.py:397-470, round_mode 'learned_hard_sigmoid'."""

    def __init__(self, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor,
                 round_mode="learned_hard_sigmoid"):
        super().__init__()
        assert round_mode == "learned_hard_sigmoid"
        self.n_bits, self.sym, self.n_levels = uaq.n_bits, uaq.sym, uaq.n_levels
        self.delta, self.zero_point = uaq.delta, uaq.zero_point
        self.round_mode, self.soft_targets = round_mode, False
        self.gamma, self.zeta = GAMMA, ZETA
        t = weight_tensor.detach().clone() / self.delta           # :454-462
        rest = t - torch.floor(t)
        self.alpha = nn.Parameter(-torch.log((ZETA - GAMMA) / (rest - GAMMA) - 1))

    def get_soft_targets(self):                                   # :451-452
        return torch.clamp(torch.sigmoid(self.alpha) * (ZETA - GAMMA) + GAMMA, 0, 1)

    def forward(self, x):                                         # :437-449
        base = torch.floor(x / self.delta)
        up = self.get_soft_targets() if self.soft_targets else (self.alpha >= 0).float()
        x_q = torch.clamp(base + up + self.zero_point, 0, self.n_levels - 1)
        return (x_q - self.zero_point) * self.delta

    def codes(self, x):
        return torch.clamp(torch.floor(x / self.delta) + (self.alpha >= 0).float() + self.zero_point,
                           0, self.n_levels - 1)


class LUUniformAffineQuantizer(UniformAffineQuantizer):
    """LU quantizer.py:130-183: forward returns (codes, delta); leaf_param => static Q8.8."""

    def _slice_scale(self, t):
        if self.scale_method == "mse":                    # LU :265-278: 80 steps of 1 %, lp_loss p=2 ('none')
            return self._shrink_search(t, 80, 0.01, lambda a, b: lp_loss(a, b))
        return super()._slice_scale(t)

    def forward(self, x, act=False):
        if not self.inited:
            if self.leaf_param:
                return lu_act_quantizer(x, 8, 8)
            self.delta, self.zero_point = self.init_quantization_scale(x, self.channel_wise)
            self.inited = True
        return self.codes(x), self.delta


class LinearTempDecay:
    """TO quantization/utils.py:37-54."""

    def __init__(self, t_max, rel_start_decay=0.2, start_b=10, end_b=2):
        self.t_max, self.start_decay = t_max, rel_start_decay * t_max
        self.start_b, self.end_b = start_b, end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, 1 - rel)
