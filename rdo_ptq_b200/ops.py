"""Tensor-level entry points over the C ABI (include/b200lic.h) + the autograd glue.

PyTorch is plumbing here: it owns device memory and streams; every arithmetic op on the hot path is a
libb200lic kernel enqueued on torch's current CUDA stream.  Tensors must be CUDA fp32; there is no CPU path.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ConvDesc, ACT_NONE, ACT_RELU, ACT_LEAKY_RELU, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC

_ENGINES = {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tc": ENGINE_TC}
DEFAULT_ENGINE = _ENGINES[os.environ.get("B200LIC_ENGINE", "auto").lower()]


def set_default_engine(name: str):
    global DEFAULT_ENGINE
    DEFAULT_ENGINE = _ENGINES[name]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c(t, name="tensor"):
    """Validate a tensor for the ABI: CUDA, fp32 (or uint8), contiguous."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"rdo_ptq_b200: `{name}` must be a CUDA tensor -- the hot path has no CPU fallback")
    if t.dtype not in (torch.float32, torch.uint8, torch.int32, torch.int64):
        raise TypeError(f"rdo_ptq_b200: `{name}` must be float32 (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def call(name, *args):
    _lib.call(name, *args, stream=_stream())


def channel_view(shape, axis):
    """(outer, ch, inner) view of a weight for per-channel quantisation along `axis` (None = per tensor)."""
    numel = 1
    for s in shape:
        numel *= s
    if axis is None:
        return 1, 1, numel
    outer = 1
    for s in shape[:axis]:
        outer *= s
    return outer, shape[axis], numel // (outer * shape[axis])


def _bshape(shape, axis, ch):
    """Broadcast shape the reference gives delta/zero_point (quantizer.py:267-279)."""
    if axis is None:
        return ()
    if len(shape) == 4:
        s = [1, 1, 1, 1]
        s[axis] = ch
        return tuple(s)
    return (ch, 1)


# ------------------------------------------------------------------------------------------------ K7 / K6
def wq_init_minmax(w, axis, n_bits=8, scale_variant=False, symmetric=False):
    w = _c(w, "weight")
    outer, ch, inner = channel_view(w.shape, axis)
    delta = torch.empty(ch, device=w.device, dtype=torch.float32)
    zp = torch.empty_like(delta)
    call("wq_init_minmax", _p(w), outer, ch, inner, n_bits, int(scale_variant), int(symmetric), _p(delta), _p(zp))
    bs = _bshape(w.shape, axis, ch)
    return delta.view(bs), zp.view(bs)


SCALE_METHODS = {"mse": 1, "l1": 2, "l2": 3, "gaussian": 4}


def wq_init_search(w, axis, n_bits=8, method="mse", n_steps=10, shrink=0.05, p=3.5, symmetric=False):
    """Search-based ('mse' / 'l1' / 'l2') and moment-based ('gaussian') ranges: quantizer.py:300-370."""
    w = _c(w, "weight")
    outer, ch, inner = channel_view(w.shape, axis)
    delta = torch.empty(ch, device=w.device, dtype=torch.float32)
    zp = torch.empty_like(delta)
    call("wq_init_search", _p(w), outer, ch, inner, n_bits, SCALE_METHODS[method], n_steps, shrink, p, int(symmetric),
         _p(delta), _p(zp))
    bs = _bshape(w.shape, axis, ch)
    return delta.view(bs), zp.view(bs)


def wq_fake_quant(w, delta, zp, axis, n_levels, want=("dq",)):
    w = _c(w, "weight")
    outer, ch, inner = channel_view(w.shape, axis)
    dq = torch.empty_like(w) if "dq" in want else None
    codes = torch.empty_like(w) if "codes" in want else None
    u8 = torch.empty(w.shape, device=w.device, dtype=torch.uint8) if "u8" in want else None
    call("wq_fake_quant", _p(w), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))), outer, ch, inner, n_levels,
         _p(dq), _p(codes), _p(u8))
    out = {"dq": dq, "codes": codes, "u8": u8}
    return out[want[0]] if len(want) == 1 else tuple(out[k] for k in want)


def wq_dequant_u8(u8, delta, zp, axis):
    u8 = _c(u8, "codes")
    outer, ch, inner = channel_view(u8.shape, axis)
    out = torch.empty(u8.shape, device=u8.device, dtype=torch.float32)
    call("wq_dequant_u8", _p(u8), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))), outer, ch, inner, _p(out))
    return out


def adaround_init_alpha(w, delta, axis):
    w = _c(w, "weight")
    outer, ch, inner = channel_view(w.shape, axis)
    alpha = torch.empty_like(w)
    call("adaround_init_alpha", _p(w), _p(_c(delta.reshape(-1))), outer, ch, inner, _p(alpha))
    return alpha


def adaround_fwd(w, alpha, delta, zp, axis, n_levels, soft, want_codes=False, out=None):
    w, alpha = _c(w, "weight"), _c(alpha, "alpha")
    outer, ch, inner = channel_view(w.shape, axis)
    wq = out if out is not None else torch.empty_like(w)
    codes = torch.empty_like(w) if want_codes else None
    call("adaround_fwd", _p(w), _p(alpha), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))), outer, ch, inner,
         n_levels, int(soft), _p(wq), _p(codes))
    return (wq, codes) if want_codes else wq


def adaround_bwd_adam(w, alpha, delta, zp, d_wq, exp_avg, exp_avg_sq, axis, n_levels, step, lr=1e-3, beta1=0.9,
                      beta2=0.999, eps=1e-8, grad_scale=1.0, reg_weight=0.0, reg_b=0.0, reg_loss=None,
                      d_alpha_out=None):
    outer, ch, inner = channel_view(w.shape, axis)
    call("adaround_bwd_adam", _p(_c(w)), _p(_c(alpha)), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))),
         _p(_c(d_wq)), _p(exp_avg), _p(exp_avg_sq), outer, ch, inner, n_levels, int(step), lr, beta1, beta2, eps,
         grad_scale, reg_weight, reg_b, _p(reg_loss), _p(d_alpha_out))


# ------------------------------------------------------------------------------------------------ K8
# One cluster launch (b200lic_actq_fused) instead of stats_init + stats + apply; both paths are bit-identical.  Measured
# (scripts/actq_time.py): the cluster kernel wins on single-image activations up to ~75 MB (launch latency, one HBM
# read), the three-launch path on batches and on 2K-size tensors (more CTAs in flight), so the wrapper picks by shape.
ACTQ_FUSED = True
ACTQ_FUSED_MAX_BYTES = 80 * 1000 * 1000


def act_quant(x, n_bits=8, want_codes=False):
    """Dynamic per-channel fake-quant of a [N,C,H,W] (or [N,C]) activation; result is detached (quantizer.py:100)."""
    x = _c(x.detach(), "activation")
    if x.dim() == 4:
        N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    elif x.dim() == 2:
        N, Cc, HW = x.shape[0], x.shape[1], 1
    else:
        raise NotImplementedError("act_quant: only NCHW / NC activations are on the hot path")
    out = torch.empty_like(x)
    codes = torch.empty_like(x) if want_codes else None
    if ACTQ_FUSED and (ACTQ_FUSED == "always" or (N == 1 and 4 * x.numel() <= ACTQ_FUSED_MAX_BYTES)):
        call("actq_fused", _p(x), N, Cc, HW, n_bits, _p(out), _p(codes))
        return (out, codes) if want_codes else out
    keys = torch.empty(2 * Cc, device=x.device, dtype=torch.int32)
    call("actq_stats_init", _p(keys), Cc)
    call("actq_stats", _p(x), N, Cc, HW, _p(keys))
    call("actq_apply", _p(x), _p(keys), N, Cc, HW, n_bits, _p(out), _p(codes))
    return (out, codes) if want_codes else out


def act_quant_stats(x):
    """Per-channel (min, max) keys of a [N,C,H,W] activation: the first half of `act_quant` (three-launch form)."""
    x = _c(x.detach(), "activation")
    N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    keys = torch.empty(2 * Cc, device=x.device, dtype=torch.int32)
    call("actq_stats_init", _p(keys), Cc)
    call("actq_stats", _p(x), N, Cc, HW, _p(keys))
    return keys


# Statistics from the producer's epilogue (b200lic_conv_stats_once): FUSE_ACTQ_STATS=False runs the separate pass (A/B).
FUSE_ACTQ_STATS = os.environ.get("B200LIC_FUSE_ACTQ_STATS", "1") != "0"


def conv_plan_info(d, transposed=False):
    """Host-side plan of the generic conv engine for descriptor `d` (b200lic_conv_plan_info) as a dict."""
    info = (C.c_int * 12)()
    rc = _lib.lib().b200lic_conv_plan_info(C.byref(d), fwd_op(transposed), info)
    if rc != 0:
        raise _lib.B200LicError("conv_plan_info", rc, _lib.lib().b200lic_last_error_string().decode())
    keys = ("eligible", "BN", "n_tiles", "MT", "m_tiles", "items", "pair", "stream_k", "grid", "stages", "epi_warps",
            "tmem_cols")
    return dict(zip(keys, list(info)))


def conv_stats_arm(channels, device):
    """Fresh (min, max) keys for `channels` output channels and a request that the next conv-engine forward fills them."""
    keys = torch.empty(2 * channels, device=device, dtype=torch.int32)
    call("actq_stats_init", _p(keys), channels)
    _lib.lib().b200lic_conv_stats_once(_p(keys))      # replaces a request an aborted caller may have left behind
    return keys


def conv_stats_taken():
    """True when a launch consumed the request of `conv_stats_arm` (its keys are valid once that launch completes)."""
    return _lib.lib().b200lic_conv_stats_pending() == 0


def act_quant_apply(x, keys, n_bits=8):
    """The second half of `act_quant`."""
    x = _c(x.detach(), "activation")
    N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    out = torch.empty_like(x)
    call("actq_apply", _p(x), _p(keys), N, Cc, HW, n_bits, _p(out), None)
    return out


def act_quant_apply_stage(x, keys, n_bits, slot, square=False, out=None):
    """`act_quant_apply` whose result leaves as the staged operand `slot` of the next layer's GEMM (and as fp32 in `out`
    when given): b200lic_actq_apply_stage."""
    x = _c(x.detach(), "activation")
    N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    hi, lo, cpad = slot
    call("actq_apply_stage", _p(x), _p(keys), N, Cc, HW, n_bits, int(bool(square)), hi, lo, cpad, _p(out))


# Deferred activation quantisation (evaluation only; switched on by evaluate.GraphedForward / evaluate.evaluate): a
# QuantModule whose output goes to exactly one consumer -- the next QuantModule of the same nn.Sequential -- returns its
# RAW output tagged with the per-channel statistics, and the consumer quantises it while staging its own GEMM operand
# (one pass instead of apply + NHWC split).  Any other reader calls `resolve_actq` first.
DEFER_ACTQ = False
# below this size the one-launch cluster kernel (b200lic_actq_fused) + the consumer's own split is cheaper than statistics
# + apply_stage: deferring trades a pass over the tensor for one more launch
# (B200LIC_DEFER_MIN_BYTES for A/B runs: 1 MB and 0.25 MB measure within 1 % of the default at 768x512 and 2K)
DEFER_ACTQ_MIN_BYTES = int(os.environ.get("B200LIC_DEFER_MIN_BYTES", 4 * 1000 * 1000))


class defer_actq:
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global DEFER_ACTQ
        self.prev, DEFER_ACTQ = DEFER_ACTQ, self.on

    def __exit__(self, *a):
        global DEFER_ACTQ
        DEFER_ACTQ = self.prev


def resolve_actq(x):
    """Materialise a deferred activation quantisation (no-op for ordinary tensors)."""
    pend = getattr(x, "_b200_actq", None)
    if pend is None:
        return x
    return act_quant_apply(x, pend[0], pend[1])


def fixed_point(x, a_l=8, a_r=8):
    x = _c(x, "activation")
    out = torch.empty_like(x)
    call("fixed_point", _p(x), x.numel(), a_l, a_r, _p(out))
    return out


# ------------------------------------------------------------------------------------------------ K9 / K10 / K11
def gaussian_lik(y, scales, means=None, scale_bound=0.11, lik_bound=1e-9, want_lik=True):
    """Returns (y_hat, lik, bits) with bits = sum(-log2 lik) as a device scalar."""
    y = _c(y, "y")
    N, Cc = y.shape[0], y.shape[1]
    HW = y.numel() // (N * Cc)
    chw = Cc * HW

    def strided_ok(t):
        return t.shape == y.shape and t.stride()[1:] == y.stride()[1:] and t.stride(0) >= chw

    if means is not None and not (strided_ok(scales) and strided_ok(means) and scales.stride(0) == means.stride(0)):
        scales, means = scales.contiguous(), means.contiguous()
    elif means is None and not strided_ok(scales):
        scales = scales.contiguous()
    for t in (scales, means):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("gaussian_lik: parameters must be CUDA fp32")
    y_hat = torch.empty_like(y)
    lik = torch.empty_like(y) if want_lik else None
    bits = torch.zeros(1, device=y.device, dtype=torch.float32)
    call("gaussian_lik_fwd", _p(y), _p(scales), _p(means), N, Cc, HW, scales.stride(0) if N > 1 else chw,
         scale_bound, lik_bound, _p(y_hat), _p(lik), _p(bits))
    return y_hat, lik, bits


def round_latent(y, means=None):
    y = _c(y, "y")
    out = torch.empty_like(y)
    call("round_latent", _p(y), _p(_c(means)), y.numel(), _p(out))
    return out


def factorized_table(packed_params, medians, lik_bound=1e-9):
    """Per-channel symbol tables of the factorised prior (b200lic_factorized_table): [C, 2, 2R+1]."""
    packed_params, medians = _c(packed_params), _c(medians)
    Cc = packed_params.shape[0]
    per = int(_lib.lib().b200lic_factorized_table_floats())
    table = torch.empty(Cc, 2, per // 2, device=packed_params.device, dtype=torch.float32)
    call("factorized_table", _p(packed_params), _p(medians), Cc, lik_bound, _p(table))
    return table


def factorized_lik(z, packed_params, medians, lik_bound=1e-9, want_lik=True, table=None):
    z = _c(z, "z")
    N, Cc = z.shape[0], z.shape[1]
    HW = z.numel() // (N * Cc)
    z_hat = torch.empty_like(z)
    lik = torch.empty_like(z) if want_lik else None
    bits = torch.zeros(1, device=z.device, dtype=torch.float32)
    call("factorized_lik_fwd", _p(z), _p(_c(packed_params)), _p(_c(medians)), _p(table), N, Cc, HW, lik_bound,
         _p(z_hat), _p(lik), _p(bits))
    return z_hat, lik, bits


def _param_views(y, scales, means):
    """scales / means may be the two halves of one [N,2C,H,W] tensor (chunk(2,1)): keep them strided when the rows of a
    sample are contiguous, else materialise.  Returns (scales, means, param_batch_stride)."""
    chw = y[0].numel()

    def strided_ok(t):
        return t.shape == y.shape and t.stride()[1:] == y.stride()[1:] and t.stride(0) >= chw

    if means is not None and not (strided_ok(scales) and strided_ok(means) and scales.stride(0) == means.stride(0)):
        scales, means = scales.contiguous(), means.contiguous()
    elif means is None and not strided_ok(scales):
        scales = scales.contiguous()
    return scales, means, (scales.stride(0) if y.shape[0] > 1 else chw)


class _GaussianLikFn(torch.autograd.Function):
    """K9 with its backward (b200lic_gaussian_lik_bwd).  `ste`: straight-through latent rounding (round_ste) instead of
    compressai's zero-gradient torch.round."""

    @staticmethod
    def forward(ctx, y, scales, means, scale_bound, lik_bound, ste):
        y_hat, lik, bits = gaussian_lik(y, scales, means, scale_bound, lik_bound)
        ctx.save_for_backward(y_hat, scales, means if means is not None else y_hat.new_empty(0))
        ctx.cfg = (float(scale_bound), float(lik_bound), bool(ste), means is not None)
        ctx.set_materialize_grads(False)
        return y_hat, lik, bits

    @staticmethod
    def backward(ctx, g_yhat, g_lik, g_bits):
        y_hat, scales, means = ctx.saved_tensors
        scale_bound, lik_bound, ste, has_means = ctx.cfg
        means = means if has_means else None
        scales, means, pstride = _param_views(y_hat, scales.detach(), None if means is None else means.detach())
        N, Cc = y_hat.shape[0], y_hat.shape[1]
        HW = y_hat.numel() // (N * Cc)
        d_y = torch.empty_like(y_hat) if ctx.needs_input_grad[0] else None
        d_s = torch.empty_like(y_hat)
        d_m = torch.empty_like(y_hat) if has_means else None
        g_yhat = None if g_yhat is None else _c(g_yhat)
        g_lik = None if g_lik is None else _c(g_lik)
        g_bits = None if g_bits is None else _c(g_bits.reshape(1))
        call("gaussian_lik_bwd", _p(y_hat), _p(scales), _p(means), _p(g_lik), _p(g_bits), _p(g_yhat), N, Cc, HW,
             pstride, Cc * HW, scale_bound, lik_bound, int(ste), _p(d_y), _p(d_s), _p(d_m))
        return d_y, d_s, (d_m if ctx.needs_input_grad[2] else None), None, None, None


def gaussian_lik_fn(y, scales, means=None, scale_bound=0.11, lik_bound=1e-9, ste=False):
    """Differentiable (y_hat, lik, bits)."""
    return _GaussianLikFn.apply(y, scales, means, scale_bound, lik_bound, ste)


class _FactorizedLikFn(torch.autograd.Function):
    """K10 with its latent gradient (b200lic_factorized_lik_bwd); the prior's parameters are frozen in PTQ."""

    @staticmethod
    def forward(ctx, z, packed_params, medians, lik_bound, ste, table=None):
        z_hat, lik, bits = factorized_lik(z, packed_params, medians, lik_bound, table=table)
        ctx.save_for_backward(z_hat, packed_params, medians)
        ctx.cfg = (float(lik_bound), bool(ste))
        ctx.set_materialize_grads(False)
        return z_hat, lik, bits

    @staticmethod
    def backward(ctx, g_zhat, g_lik, g_bits):
        z_hat, packed, med = ctx.saved_tensors
        lik_bound, ste = ctx.cfg
        N, Cc = z_hat.shape[0], z_hat.shape[1]
        HW = z_hat.numel() // (N * Cc)
        d_z = torch.empty_like(z_hat)
        g_zhat = None if g_zhat is None else _c(g_zhat)
        g_lik = None if g_lik is None else _c(g_lik)
        g_bits = None if g_bits is None else _c(g_bits.reshape(1))
        call("factorized_lik_bwd", _p(z_hat), _p(_c(packed)), _p(_c(med)), _p(g_lik), _p(g_bits), _p(g_zhat), N, Cc, HW,
             lik_bound, int(ste), _p(d_z))
        return d_z, None, None, None, None, None


def factorized_lik_fn(z, packed_params, medians, lik_bound=1e-9, ste=False, table=None):
    return _FactorizedLikFn.apply(z, packed_params, medians, lik_bound, ste, table)


class _RoundLatentSTE(torch.autograd.Function):
    """rint(y - mu) + mu with the straight-through gradient of round_ste (quantizer.py:64-68): d/dy = 1, d/dmu = 0."""

    @staticmethod
    def forward(ctx, y, means):
        return round_latent(y, means)

    @staticmethod
    def backward(ctx, g):
        return g, None


def round_latent_ste(y, means=None):
    return _RoundLatentSTE.apply(y, means)


class _LpLossFn(torch.autograd.Function):
    """scale * sum |pred - tgt|^p with its gradient w.r.t. pred, both from one K11 pass."""

    @staticmethod
    def forward(ctx, pred, tgt, p, scale):
        val, g = lp_loss_fwd_bwd(pred, tgt, p, scale=scale, grad_scale=scale)
        ctx.save_for_backward(g)
        return val.reshape(())

    @staticmethod
    def backward(ctx, g_out):
        (g,) = ctx.saved_tensors
        return g * g_out, None, None, None


def lp_loss_fn(pred, tgt, p=2.0, scale=1.0):
    return _LpLossFn.apply(pred, tgt, float(p), float(scale))


class _RDLossFn(torch.autograd.Function):
    """RateDistortionLoss (losses/losses.py:20-28, MSE metric): loss = lmbda*255^2*mean((x_hat-x)^2) + bits/pixels.
    One K11 pass yields the distortion value and its gradient; bits come from the likelihood kernels' reductions."""

    @staticmethod
    def forward(ctx, x_hat, target, bits, lmbda, pixels):
        k = float(lmbda) * 255.0 ** 2 / x_hat.numel()
        dist, g = lp_loss_fwd_bwd(x_hat, target, 2.0, scale=k, grad_scale=k)     # kernel applies p*|d|^(p-1)
        ctx.save_for_backward(g)
        ctx.inv_px = 1.0 / float(pixels)
        return add_act(dist, bits.reshape(1) * ctx.inv_px).reshape(())

    @staticmethod
    def backward(ctx, g_out):
        (g,) = ctx.saved_tensors
        return g * g_out, None, (g_out * ctx.inv_px).reshape(1), None, None


def rd_loss(x_hat, target, bits, lmbda=1e-2, pixels=None):
    """Differentiable R + lambda*D; `bits` = device scalar sum of -log2 likelihoods."""
    N, _, H, W = target.shape
    return _RDLossFn.apply(x_hat, target, bits, lmbda, N * H * W if pixels is None else pixels)


def lsq_delta_grad(w, delta, zp, d_wq, axis, n_levels, alpha=None, soft=True, grad_scale=1.0, adam=None, sched=None,
                   lr_scale=1.0):
    """Per-channel d loss / d delta from dL/dWq (learned step size).  adam = (exp_avg, exp_avg_sq, step, lr) updates
    delta in place; with `sched` (device-resident schedule) adam = (exp_avg, exp_avg_sq) and lr = lr_scale * the
    schedule's.  Returns d_delta ([ch], always written)."""
    w, d_wq = _c(w, "w"), _c(d_wq, "d_wq")
    outer, ch, inner = channel_view(w.shape, axis)
    dflat = delta.view(-1)
    if not dflat.is_contiguous() or dflat.numel() != ch:
        raise ValueError("lsq_delta_grad: delta must hold one contiguous value per quantisation channel")
    d_delta = torch.empty(ch, device=w.device, dtype=torch.float32)
    m = v = None
    step, lr = 0, 0.0
    if sched is not None:
        m, v = adam[0], adam[1]
        call("lsq_delta_grad_sched", _p(w), _p(None if alpha is None else _c(alpha)), _p(dflat), _p(_c(zp.view(-1))),
             _p(d_wq), outer, ch, inner, int(n_levels), int(bool(soft)), float(grad_scale), _p(d_delta), _p(m), _p(v),
             _p(sched), float(lr_scale), 0.9, 0.999, 1e-8)
        return d_delta
    if adam is not None:
        m, v, step, lr = adam
    call("lsq_delta_grad", _p(w), _p(None if alpha is None else _c(alpha)), _p(dflat), _p(_c(zp.view(-1))), _p(d_wq),
         outer, ch, inner, int(n_levels), int(bool(soft)), float(grad_scale), _p(d_delta), _p(m), _p(v), int(step),
         float(lr), 0.9, 0.999, 1e-8)
    return d_delta


def lp_loss_fwd_bwd(pred, tgt, p=2.0, scale=1.0, grad_scale=None, loss=None, want_grad=True, pick=None):
    """loss += scale * sum|pred-tgt|^p ; returns (loss, d_pred).
    pick = (idx_table, units, unit, sched): `tgt` is a [samples, ...] cache and row b of `pred` is compared with
    tgt[idx_table[k % rows][b]], k = (sched.step - 1) * units + unit (the pick gather_mix_sched makes)."""
    pred, tgt = _c(pred, "pred"), _c(tgt, "tgt")
    if loss is None:
        loss = torch.zeros(1, device=pred.device, dtype=torch.float32)
    g = torch.empty_like(pred) if want_grad else None
    gs = float(scale if grad_scale is None else grad_scale)
    if pick is None:
        call("lp_loss_fwd_bwd", _p(pred), _p(tgt), pred.numel(), float(p), float(scale), gs, _p(loss), _p(g))
    else:
        idx_table, units, unit, sched = pick
        rows, row = pred.shape[0], pred[0].numel()
        if idx_table.dtype != torch.int64 or idx_table.dim() != 2 or idx_table.size(1) != rows or tgt[0].numel() != row:
            raise ValueError("lp_loss_fwd_bwd: pick table / target cache do not match the prediction batch")
        call("lp_loss_fwd_bwd_sched", _p(pred), _p(tgt), _p(idx_table.contiguous()), idx_table.size(0), rows, row,
             int(units), int(unit), _p(sched), float(p), float(scale), gs, _p(loss), _p(g))
    return loss, g


def sq_err_sum(a, b):
    a, b = _c(a), _c(b)
    out = torch.zeros(2, device=a.device, dtype=torch.float32)
    call("sq_err_sum", _p(a), _p(b), a.numel(), _p(out))
    return out


def bits_sum(lik):
    lik = _c(lik)
    out = torch.zeros(1, device=lik.device, dtype=torch.float32)
    call("bits_sum", _p(lik), lik.numel(), _p(out))
    return out


# ------------------------------------------------------------------------------------------------ token-major ops
def linear(x, w, bias=None, packed=None):
    """F.linear(x, w, bias) for x [..., Cin], w [Cout, Cin] on the tcgen05 conv engine: the tokens are `rows` one-pixel
    images, the token matrix is already the engine's NHWC operand (include/b200lic.h, "Token-major pieces").  `packed`:
    a prepared weight operand of `linear_desc(rows, Cin, Cout)` (pack_weights of w.view(Cout, Cin, 1, 1))."""
    x = _c(x, "input")
    Cout, Cin = w.shape
    if x.shape[-1] != Cin:
        raise ValueError(f"linear: weight {tuple(w.shape)} does not match input features {x.shape[-1]}")
    rows = x.numel() // Cin
    d = linear_desc(rows, Cin, Cout)
    ws = _workspace(d, _lib.OP_CONV_FWD, x.device)
    slot = None if ws[0] is None else conv_x_slot(d, False, ws)
    if slot is None:
        raise _lib.B200LicError("linear", -4, f"no tensor-core plan for {rows} x {Cin} -> {Cout}")
    hi, lo, cpad = slot
    call("stage_tokens", _p(x), rows, Cin, cpad, hi, lo)
    if packed is None:
        packed = pack_weights(_c(w, "weight").view(Cout, Cin, 1, 1), d, False)
    y = conv_fwd_packed(None, packed, d, False, bias=bias, ws=ws)
    return y.view(*x.shape[:-1], Cout)


def linear_desc(rows, Cin, Cout):
    return conv_desc((rows, Cin, 1, 1), (Cout, Cin, 1, 1), 1, 0)


def layer_norm(x, weight=None, bias=None, eps=1e-5):
    """F.layer_norm over the last axis (b200lic_layernorm_fwd)."""
    x = _c(x, "input")
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    call("layernorm_fwd", _p(x), _p(_c(weight)), _p(_c(bias)), x.numel() // Cc, Cc, float(eps), _p(y))
    return y


def act_quant_tokens(x, n_bits=8):
    """ActQuantizer of a [..., C] token tensor: per last-axis channel over everything else (quantizer.py:81-121, 3-D branch)."""
    x = _c(x.detach(), "activation")
    Cc = x.shape[-1]
    keys = torch.empty(2 * Cc, device=x.device, dtype=torch.int32)
    out = torch.empty_like(x)
    call("actq_tokens", _p(x), x.numel() // Cc, Cc, int(n_bits), _p(keys), _p(out))
    return out


def window_attn_softmax(qkv, bias, mask, num_heads, scale):
    """softmax((q * scale) @ k^T + bias [+ mask]) of window attention: qkv [B_, N, 3C] -> P [B_, nH, N, N]."""
    qkv, bias = _c(qkv, "qkv"), _c(bias, "bias")
    B_, N, C3 = qkv.shape
    Cc = C3 // 3
    P = torch.empty(B_, num_heads, N, N, device=qkv.device, dtype=torch.float32)
    nW = 1 if mask is None else mask.shape[0]
    call("window_attn_softmax", _p(qkv), _p(bias), _p(_c(mask)), B_, N, Cc, num_heads, nW, float(scale), _p(P))
    return P


def window_attn_apply(P, qkv):
    """(P @ v).transpose(1, 2).reshape(B_, N, C) with v taken from qkv [B_, N, 3C]."""
    P, qkv = _c(P, "attn"), _c(qkv, "qkv")
    B_, nH, N, _ = P.shape
    Cc = qkv.shape[2] // 3
    out = torch.empty(B_, N, Cc, device=qkv.device, dtype=torch.float32)
    call("window_attn_apply", _p(P), _p(qkv), B_, N, Cc, nH, _p(out))
    return out


def gelu(x):
    x = _c(x, "input")
    y = torch.empty_like(x)
    call("gelu_fwd", _p(x), x.numel(), _p(y))
    return y


_SSIM_WIN = {}


def ms_ssim(x, y, data_range=1.0, size_average=True):
    """pytorch_msssim.ms_ssim(x, y, data_range) of two [N,C,H,W] tensors (reference call sites: include/b200lic.h,
    "MS-SSIM"): five levels of b200lic_ssim_level with b200lic_avg_pool2 between them, b200lic_msssim_combine.
    Returns a device scalar (size_average) or the per-image means [N]."""
    x, y = _c(x, "x"), _c(y, "y")
    if x.shape != y.shape or x.dim() != 4:
        raise ValueError(f"ms_ssim: shapes {tuple(x.shape)} / {tuple(y.shape)}")
    N, Cc, H, W = x.shape
    if min(H, W) <= (11 - 1) * 2 ** 4:
        raise ValueError("ms_ssim: image side must exceed 160 px (four 2x downsamplings, 11-tap window)")
    dev = x.device
    win = _SSIM_WIN.get(dev)
    if win is None:
        c = torch.arange(11, dtype=torch.float32) - 5
        g = torch.exp(-(c ** 2) / (2 * 1.5 ** 2))
        win = _SSIM_WIN[dev] = (g / g.sum()).to(dev)
    planes, levels = N * Cc, 5
    sums = torch.zeros(levels, planes, 2, device=dev, dtype=torch.float64)
    inv = []
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    for lvl in range(levels):
        call("ssim_level", _p(x), _p(y), _p(win), planes, H, W, c1, c2, _p(sums[lvl]))
        inv.append(1.0 / ((H - 10) * (W - 10)))
        if lvl < levels - 1:
            ph, pw = H % 2, W % 2
            Ho, Wo = (H + 2 * ph - 2) // 2 + 1, (W + 2 * pw - 2) // 2 + 1
            xn = torch.empty(N, Cc, Ho, Wo, device=dev, dtype=torch.float32)
            yn = torch.empty_like(xn)
            call("avg_pool2", _p(x), planes, H, W, ph, pw, _p(xn))
            call("avg_pool2", _p(y), planes, H, W, ph, pw, _p(yn))
            x, y, H, W = xn, yn, Ho, Wo
    inv_t = torch.tensor(inv, dtype=torch.float64).to(dev)
    per_plane = torch.empty(planes, device=dev, dtype=torch.float32)
    mean = torch.empty(1, device=dev, dtype=torch.float32)
    call("msssim_combine", _p(sums), _p(inv_t), levels, planes, _p(per_plane), _p(mean))
    return mean[0] if size_average else per_plane.view(N, Cc).mean(1)


# ------------------------------------------------------------------------------------------------ convolutions
def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _sq(v, what):
    a, b = _pair(v)
    if a != b:
        raise NotImplementedError(f"rdo_ptq_b200: non-square {what} {v}")
    return a


def conv_desc(x_shape, w_shape, stride, padding, transposed=False, output_padding=0, act=ACT_NONE, slope=0.01,
              engine=None, in_square=0, gdn_mode=0, fixed_pt=0, k_taps=0):
    N, Cin, H, W = x_shape
    st, pd = _sq(stride, "stride"), _sq(padding, "padding")
    KH, KW = w_shape[2], w_shape[3]
    if transposed:
        op = _sq(output_padding, "output_padding")
        if w_shape[0] != Cin:
            raise ValueError(f"conv_transpose2d: weight {tuple(w_shape)} does not match input channels {Cin}")
        Cout = w_shape[1]
        Ho, Wo = (H - 1) * st - 2 * pd + KH + op, (W - 1) * st - 2 * pd + KW + op
    else:
        if w_shape[1] != Cin:
            raise ValueError(f"conv2d: weight {tuple(w_shape)} does not match input channels {Cin}")
        Cout = w_shape[0]
        Ho, Wo = (H + 2 * pd - KH) // st + 1, (W + 2 * pd - KW) // st + 1
    return ConvDesc(N, Cin, H, W, Cout, Ho, Wo, KH, KW, st, pd, act, slope,
                    DEFAULT_ENGINE if engine is None else engine, in_square, gdn_mode, fixed_pt, k_taps)


def _act_id(module):
    """Map an absorbed activation module (quant_model.py:51-56) to (act id, slope)."""
    import torch.nn as nn
    if module is None or type(module).__name__ == "StraightThrough":
        return ACT_NONE, 0.0
    if isinstance(module, nn.LeakyReLU):
        return ACT_LEAKY_RELU, float(module.negative_slope)
    if isinstance(module, nn.ReLU):
        return ACT_RELU, 0.0
    raise NotImplementedError(f"rdo_ptq_b200: activation {module} is not on the hot path")


def _workspace(d, op, device):
    """Scratch for the tensor-core engine (caller-owned, C ABI contract); empty when the engine/shape is SIMT-only."""
    if d.engine == ENGINE_SIMT:
        return None, 0
    n = int(_lib.lib().b200lic_conv_workspace_bytes(C.byref(d), op))
    if n == 0:
        return None, 0
    return torch.empty(n, dtype=torch.uint8, device=device), n


def conv2d_raw(x, w, bias, d, gdn_x=None, want_norm=False, want_ws=False):
    y = torch.empty((d.N, d.Cout, d.Ho, d.Wo), device=x.device, dtype=torch.float32)
    norm = torch.empty_like(y) if want_norm else None
    ws, nws = _workspace(d, _lib.OP_CONV_FWD, x.device)
    call("conv_fwd", C.byref(d), _p(x), _p(w), _p(bias), _p(gdn_x), _p(norm), _p(y), _p(ws), nws)
    out = (y, norm) if want_norm else y
    return (out, (ws, nws)) if want_ws else out


def deconv2d_raw(x, w, bias, d, want_ws=False):
    y = torch.empty((d.N, d.Cout, d.Ho, d.Wo), device=x.device, dtype=torch.float32)
    ws, nws = _workspace(d, _lib.OP_DECONV_FWD, x.device)
    call("deconv_fwd", C.byref(d), _p(x), _p(w), _p(bias), _p(y), _p(ws), nws)
    return (y, (ws, nws)) if want_ws else y


def _staged_x(d, op, fwd_ws):
    """(x_hi, x_lo) device pointers of the split-bf16 NHWC copy of x the forward call left in its workspace, or None
    when the wgrad engine cannot consume it (b200lic_conv_staged_view)."""
    ws, nws = fwd_ws
    if ws is None:
        return None
    hi, lo = C.c_void_p(), C.c_void_p()
    rc = _lib.lib().b200lic_conv_staged_view(C.byref(d), op, _p(ws), nws, C.byref(hi), C.byref(lo))
    if rc != 0 or not hi.value:
        return None
    return hi, lo


def _wgrad(d, transposed, x, dy, dw, fwd_ws):
    """dW through the tensor-core engine, reusing the forward's staged copy of x when there is one."""
    op = _lib.OP_DECONV_WGRAD if transposed else _lib.OP_CONV_WGRAD
    ws, nws = _workspace(d, op, dy.device)
    st = _staged_x(d, _lib.OP_DECONV_FWD if transposed else _lib.OP_CONV_FWD, fwd_ws) if ws is not None else None
    if st is not None:
        call("conv_wgrad_staged", C.byref(d), int(transposed), st[0], st[1], _p(dy), _p(dw), _p(ws), nws)
    else:
        call("deconv_wgrad" if transposed else "conv_wgrad", C.byref(d), _p(x), _p(dy), _p(dw), _p(ws), nws)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, act, slope, fixed_pt):
        x, w, bias = _c(x, "input"), _c(w, "weight"), _c(bias, "bias")
        d = conv_desc(x.shape, w.shape, stride, padding, act=act, slope=slope, fixed_pt=fixed_pt)
        need_dw = ctx.needs_input_grad[1]
        res = conv2d_raw(x, w, bias, d, want_ws=need_dw)
        y, ctx.fwd_ws = res if need_dw else (res, (None, 0))
        ctx.d, ctx.act, ctx.slope = d, act, slope
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy, "grad")
        if ctx.act != ACT_NONE:
            dy = act_bwd(y, dy, ctx.act, ctx.slope)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ws, nws = _workspace(ctx.d, _lib.OP_CONV_DGRAD, dy.device)
            call("conv_dgrad", C.byref(ctx.d), _p(dy), _p(w), _p(dx), _p(ws), nws)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            _wgrad(ctx.d, False, x, dy, dw, ctx.fwd_ws)
        # bias gradient (never needed by the AdaRound loop, which freezes everything but alpha; kept so that autograd use
        # of the op outside that loop is complete): a [Cout] reduction of dL/d(pre-activation)
        db = dy.sum((0, 2, 3)) if ctx.needs_input_grad[2] else None
        ctx.fwd_ws = (None, 0)
        return dx, dw, db, None, None, None, None, None


class _DeconvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, output_padding, act, slope, fixed_pt):
        x, w, bias = _c(x, "input"), _c(w, "weight"), _c(bias, "bias")
        d = conv_desc(x.shape, w.shape, stride, padding, True, output_padding, act=act, slope=slope, fixed_pt=fixed_pt)
        need_dw = ctx.needs_input_grad[1]
        res = deconv2d_raw(x, w, bias, d, want_ws=need_dw)
        y, ctx.fwd_ws = res if need_dw else (res, (None, 0))
        ctx.d, ctx.act, ctx.slope = d, act, slope
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy, "grad")
        if ctx.act != ACT_NONE:
            dy = act_bwd(y, dy, ctx.act, ctx.slope)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ws, nws = _workspace(ctx.d, _lib.OP_DECONV_DGRAD, dy.device)
            call("deconv_dgrad", C.byref(ctx.d), _p(dy), _p(w), _p(dx), _p(ws), nws)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            _wgrad(ctx.d, True, x, dy, dw, ctx.fwd_ws)
        db = dy.sum((0, 2, 3)) if ctx.needs_input_grad[2] else None
        ctx.fwd_ws = (None, 0)
        return dx, dw, db, None, None, None, None, None, None


def wq_int_weights(w, delta, zp, axis, n_levels, alpha=None):
    """n = code - zero_point (integer-valued fp32): nearest rounding, or hardened AdaRound when `alpha` is given."""
    w = _c(w, "weight")
    outer, ch, inner = channel_view(w.shape, axis)
    out = torch.empty_like(w)
    call("wq_int_weights", _p(w), _p(None if alpha is None else _c(alpha)), _p(_c(delta.reshape(-1))),
         _p(_c(zp.reshape(-1))), outer, ch, inner, int(n_levels), _p(out))
    return out


def conv_wq(x, w_int, w_scale, bias=None, stride=1, padding=0, output_padding=0, dilation=1, groups=1, transposed=False,
            act=ACT_NONE, slope=0.01, fixed_pt=0):
    """Inference-only forward with integer-valued weights and a per-output-channel scale (two tensor-core passes):
    act(conv(x, w_int) * w_scale + bias).  Returns None when the shape has no two-pass kernel (the caller then runs the
    regular op on the dequantised weight)."""
    if _sq(dilation, "dilation") != 1 or groups != 1:
        return None
    x, w_int, w_scale, bias = _c(x, "input"), _c(w_int, "weight"), _c(w_scale, "scale"), _c(bias, "bias")
    d = conv_desc(x.shape, w_int.shape, stride, padding, transposed, output_padding, act=act, slope=slope,
                  fixed_pt=fixed_pt)
    ws, nws = _workspace(d, _lib.OP_DECONV_FWD if transposed else _lib.OP_CONV_FWD, x.device)
    if ws is None:
        return None
    y = torch.empty((d.N, d.Cout, d.Ho, d.Wo), device=x.device, dtype=torch.float32)
    try:
        call("deconv_fwd_wq" if transposed else "conv_fwd_wq", C.byref(d), _p(x), _p(w_int), _p(w_scale), _p(bias),
             _p(y), _p(ws), nws)
    except _lib.B200LicError as e:
        if e.code == -4:                 # B200LIC_ERR_UNSUPPORTED: folded-tap layers keep the three-pass engine
            return None
        raise
    return y


# ------------------------------------------------------------------------------------------------ prepared operands
def fwd_op(transposed):
    return _lib.OP_DECONV_FWD if transposed else _lib.OP_CONV_FWD


def wgrad_op(transposed):
    return _lib.OP_DECONV_WGRAD if transposed else _lib.OP_CONV_WGRAD


def packed_weight_bytes(d, transposed):
    """Bytes of the packed tensor-core weight operand of this layer; 0 = no prepared-operand path (folded-tap layer, SIMT)."""
    return int(_lib.lib().b200lic_conv_packed_weight_bytes(C.byref(d), fwd_op(transposed)))


def new_packed(d, transposed, device):
    n = packed_weight_bytes(d, transposed)
    return None if n == 0 else torch.empty(n, dtype=torch.uint8, device=device)


def pack_weights(w, d, transposed, out=None):
    """fp32 weight -> packed split-bf16 operand (b200lic_conv_pack_weights)."""
    w = _c(w, "weight")
    out = new_packed(d, transposed, w.device) if out is None else out
    if out is None:
        return None
    call("conv_pack_weights", C.byref(d), fwd_op(transposed), _p(w), _p(out), out.numel())
    return out


def quant_pack_weights(w, alpha, delta, zp, axis, n_levels, soft, d, transposed, integer_mode=False, out=None, w_q=None):
    """Weight quantiser (nearest when alpha is None, else AdaRound soft / hard) fused with the packing."""
    w = _c(w, "weight")
    outer, ch, inner = channel_view(w.shape, axis)
    out = new_packed(d, transposed, w.device) if out is None else out
    if out is None:
        return None
    call("quant_pack_weights", C.byref(d), fwd_op(transposed), _p(w), _p(None if alpha is None else _c(alpha)),
         _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))), outer, ch, inner, int(n_levels), int(bool(soft)),
         int(bool(integer_mode)), _p(out), out.numel(), _p(w_q))
    return out


def conv_fwd_packed(x, packed, d, transposed, bias=None, w_scale=None, gdn_x=None, want_norm=False, ws=None, y=None,
                    norm=None):
    """Forward with a prepared weight operand; x=None: the activation operand is already staged in `ws`.
    Returns y, or (y, norm) with want_norm (`norm`: caller-owned buffer for it)."""
    if ws is None:
        ws = _workspace(d, fwd_op(transposed), packed.device)
    wsb, nws = ws
    if y is None:
        y = torch.empty((d.N, d.Cout, d.Ho, d.Wo), device=packed.device, dtype=torch.float32)
    if want_norm and norm is None:
        norm = torch.empty_like(y)
    call("conv_fwd_packed", C.byref(d), fwd_op(transposed), _p(None if x is None else _c(x, "input")), _p(packed),
         _p(_c(w_scale)), _p(_c(bias)), _p(gdn_x), _p(norm), _p(y), _p(wsb), nws)
    return (y, norm) if want_norm else y


def folded_deconv_desc(d):
    """The 1x1 transposed-conv descriptor a folded-tap transposed conv (few output channels: Cout*k*k <= 128) runs as --
    col[n, (co,r,s), h, w] = sum_ci x[n,ci,h,w] * w[ci,co,r,s], finished by `col2im` -- or None (conv_tc_smallc.cu)."""
    kk = d.KH * d.KW
    if kk <= 1 or d.Cout * kk > 128 or d.stride > 4 or d.in_square or d.gdn_mode or d.engine == ENGINE_SIMT:
        return None
    return ConvDesc(d.N, d.Cin, d.H, d.W, d.Cout * kk, d.H, d.W, 1, 1, 1, 0, ACT_NONE, 0.0, d.engine, 0, 0, 0)


def im2col_stage(x, kh, kw, stride, pad, ho, wo, slot):
    """im2col of x [N,C,H,W] written as the staged split-bf16 operand `slot` (b200lic_im2col_stage)."""
    x = _c(x, "input")
    N, Cc, H, W = x.shape
    call("im2col_stage", _p(x), N, Cc, H, W, kh, kw, stride, pad, ho, wo, slot[0], slot[1], slot[2])


def col2im(col, bias, cout, kh, kw, stride, pad, ho, wo, act=ACT_NONE, slope=0.0, y=None):
    """Tail of the folded transposed conv: y [N,cout,ho,wo] = act(bias + scatter-add of col [N,cout*kh*kw,H,W])."""
    N, _, H, W = col.shape
    if y is None:
        y = torch.empty((N, cout, ho, wo), device=col.device, dtype=torch.float32)
    call("col2im", _p(col), _p(None if bias is None else _c(bias)), N, cout, H, W, kh, kw, stride, pad, ho, wo, int(act),
         float(slope), 0, _p(y))
    return y


GDN_FUSED = True           # evaluation GDN / IGDN through b200lic_gdn_fwd_fused when the shape allows (tests flip it for A/B)


def gdn_fused_ok(C_, HW):
    return GDN_FUSED and bool(_lib.lib().b200lic_gdn_fused_ok(int(C_), int(HW)))


def gdn_fwd_fused(x, packed, beta, inverse, pending=None, y=None):
    """GDN / IGDN forward in one kernel over the raw NCHW tensor: b200lic_gdn_fwd_fused.  `pending` = (minmax keys,
    n_bits) of a deferred activation quantiser on x (applied on chip), or None."""
    x = _c(x.detach(), "input")
    N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    if y is None:
        y = torch.empty_like(x)
    keys, bits = pending if pending is not None else (None, 8)
    call("gdn_fwd_fused", _p(x), _p(keys), int(bits), _p(packed), _p(_c(beta)), N, Cc, HW, int(bool(inverse)), _p(y))
    return y


def _slot(name, d, op, ws):
    wsb, nws = ws
    if wsb is None:
        return None
    hi, lo, cpad = C.c_void_p(), C.c_void_p(), C.c_int()
    rc = getattr(_lib.lib(), "b200lic_" + name)(C.byref(d), op, _p(wsb), nws, C.byref(hi), C.byref(lo), C.byref(cpad))
    if rc != 0 or not hi.value:
        return None
    return hi, lo, cpad.value


def conv_x_slot(d, transposed, ws):
    """(hi, lo, cpad) of the staged activation operand inside a forward workspace, or None."""
    return _slot("conv_x_slot", d, fwd_op(transposed), ws)


def conv_dy_slot(d, transposed, ws):
    """(hi, lo, cpad) of the staged dY operand inside a weight-gradient workspace, or None."""
    return _slot("conv_dy_slot", d, wgrad_op(transposed), ws)


def stage_mix_sched(q, fp, idx_table, rows, prob, seed_base, units, unit, sched, slot, square=False, out=None):
    """gather_mix_sched whose result is written as the staged (split-bf16 NHWC) activation operand `slot`."""
    q, fp = _c(q), _c(fp)
    Cc, HW = q.shape[1], q[0, 0].numel()
    hi, lo, cpad = slot
    call("stage_mix_sched", _p(q), _p(fp), _p(idx_table), 0 if idx_table is None else idx_table.size(0), int(rows), Cc, HW,
         float(prob), int(seed_base) & 0xFFFFFFFFFFFFFFFF, int(units), int(unit), _p(sched), int(bool(square)), hi, lo,
         cpad, _p(out))


def lp_loss_stage_sched(pred, tgt_cache, idx_table, units, unit, sched, p, scale, grad_scale, act, slope, loss, slot,
                        d_pred=None, gdn=None):
    """lp_loss value + gradient with the gradient written as the staged dY operand `slot` of the weight gradient.
    gdn = (x, norm, inverse): `pred` is a GDN unit's output; the gradient is carried on to the norm accumulator."""
    tgt_cache = _c(tgt_cache)
    gx, gn, ginv = (None, None, 0) if gdn is None else (_c(gdn[0]), _c(gdn[1]), int(bool(gdn[2])))
    pred = None if pred is None else _c(pred)          # None (GDN only): y = x * norm^-+1/2 is recomputed in the kernel
    shp = (pred if pred is not None else gx).shape
    rows, Cc, HW = shp[0], shp[1], shp[2] * shp[3]
    hi, lo, cpad = slot
    call("lp_loss_stage_sched", _p(pred), _p(tgt_cache), _p(idx_table), 0 if idx_table is None else idx_table.size(0),
         rows, Cc, HW, int(units), int(unit), _p(sched), float(p), float(scale), float(grad_scale), int(act), float(slope),
         _p(gx), _p(gn), ginv, _p(loss), hi, lo, cpad, _p(d_pred))


def conv_wgrad_prepared(d, transposed, x_slot, dy, dw, ws):
    """Weight gradient from a pre-staged x (and, with dy=None, a pre-staged dY in `ws`)."""
    wsb, nws = ws
    call("conv_wgrad_prepared", C.byref(d), int(transposed), x_slot[0], x_slot[1], x_slot[2], _p(dy), _p(dw), _p(wsb), nws)


def conv_wgrad_adam_sched(d, transposed, x_slot, dy, ws, w, alpha, delta, zp, exp_avg, exp_avg_sq, axis, n_levels, sched,
                          beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, reg_weight=0.0, reg_loss=None, dw_out=None):
    """Weight gradient with the AdaRound backward + Adam step fused behind its split-K reduction."""
    outer, ch, inner = channel_view(w.shape, axis)
    wsb, nws = ws
    call("conv_wgrad_adam_sched", C.byref(d), int(transposed), x_slot[0], x_slot[1], x_slot[2], _p(dy), _p(wsb), nws,
         _p(_c(w)), _p(_c(alpha)), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))), _p(exp_avg), _p(exp_avg_sq), outer,
         ch, inner, int(n_levels), _p(sched), beta1, beta2, eps, float(grad_scale), float(reg_weight), _p(reg_loss),
         _p(dw_out))


def conv2d(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1, act=ACT_NONE, slope=0.01, fixed_pt=0):  # noqa
    """Drop-in for F.conv2d on the hot path (dilation 1, groups 1), optional fused activation."""
    if _sq(dilation, "dilation") != 1 or groups != 1:
        raise NotImplementedError("rdo_ptq_b200.conv2d: dilation/groups != 1 are not on the hot path")
    return _ConvFn.apply(x, w, bias, stride, padding, act, slope, fixed_pt)


def conv_transpose2d(x, w, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1, act=ACT_NONE,
                     slope=0.01, fixed_pt=0):
    """Drop-in for F.conv_transpose2d on the hot path."""
    if _sq(dilation, "dilation") != 1 or groups != 1:
        raise NotImplementedError("rdo_ptq_b200.conv_transpose2d: dilation/groups != 1 are not on the hot path")
    return _DeconvFn.apply(x, w, bias, stride, padding, output_padding, act, slope, fixed_pt)


# ------------------------------------------------------------------------------------------------ GDN
def gdn_reparam(p, bound, pedestal, out=None):
    p = _c(p)
    out = torch.empty_like(p) if out is None else out
    call("gdn_reparam_fwd", _p(p), p.numel(), bound, pedestal, _p(out))
    return out


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, bound, pedestal):
        p = _c(p)
        ctx.bound = bound
        ctx.save_for_backward(p)
        return gdn_reparam(p, bound, pedestal)

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        g = _c(g)
        dp = torch.empty_like(p)
        call("gdn_reparam_bwd", _p(p), _p(g), p.numel(), ctx.bound, _p(dp))
        return dp, None, None


def gdn_reparam_bwd(p, g, bound, out=None):
    """LowerBound-gated backward of `gdn_reparam` (b200lic_gdn_reparam_bwd)."""
    p, g = _c(p), _c(g)
    out = torch.empty_like(p) if out is None else out
    call("gdn_reparam_bwd", _p(p), _p(g), p.numel(), float(bound), _p(out))
    return out


def gdn_desc(x_shape, inverse, want_norm=False):
    N, Cc, H, W = x_shape
    return ConvDesc(N, Cc, H, W, Cc, H, W, 1, 1, 1, 0, ACT_NONE, 0.0, DEFAULT_ENGINE, 1, 2 if inverse else 1, 0)


class _GdnFn(torch.autograd.Function):
    """y = x * (beta + gamma . x^2)^(-1/2 | +1/2) with effective (re-parametrised) gamma [C,C], beta [C]."""

    @staticmethod
    def forward(ctx, x, gamma_eff, beta_eff, inverse):
        x, gamma_eff, beta_eff = _c(x), _c(gamma_eff), _c(beta_eff)
        d = gdn_desc(x.shape, inverse)
        need_bwd = any(ctx.needs_input_grad[:2])
        res, ctx.fwd_ws = conv2d_raw(x, gamma_eff, beta_eff, d, gdn_x=x, want_norm=need_bwd, want_ws=True)
        y, norm = res if need_bwd else (res, None)
        if not ctx.needs_input_grad[1]:
            ctx.fwd_ws = (None, 0)
        ctx.d, ctx.inverse = d, inverse
        ctx.save_for_backward(x, gamma_eff, norm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma_eff, norm = ctx.saved_tensors
        dy = _c(dy)
        need_dx = ctx.needs_input_grad[0]
        d_norm = torch.empty_like(x)
        dx_direct = torch.empty_like(x) if need_dx else None
        call("gdn_bwd_prep", _p(x), _p(norm), _p(dy), x.numel(), int(ctx.inverse), _p(d_norm), _p(dx_direct))
        d = ConvDesc(ctx.d.N, ctx.d.Cin, ctx.d.H, ctx.d.W, ctx.d.Cout, ctx.d.Ho, ctx.d.Wo, 1, 1, 1, 0, ACT_NONE, 0.0,
                     ctx.d.engine, 1, 0, 0)
        dx = dgamma = None
        if ctx.needs_input_grad[1]:
            dgamma = torch.empty_like(gamma_eff)
            _wgrad(d, False, x, d_norm, dgamma, ctx.fwd_ws)                              # in_square: sum d_norm * x^2
            ctx.fwd_ws = (None, 0)
        if need_dx:
            d.in_square = 0
            t = torch.empty_like(x)
            ws, nws = _workspace(d, _lib.OP_CONV_DGRAD, x.device)
            call("conv_dgrad", C.byref(d), _p(d_norm), _p(gamma_eff), _p(t), _p(ws), nws)
            dx = torch.empty_like(x)
            call("gdn_bwd_finish", _p(x), _p(t), _p(dx_direct), x.numel(), _p(dx))
        dbeta = None
        if ctx.needs_input_grad[2]:
            dbeta = d_norm.sum(dim=(0, 2, 3))
        return dx, dgamma, dbeta, None


def gdn(x, gamma_eff, beta_eff, inverse):
    return _GdnFn.apply(x, gamma_eff, beta_eff, bool(inverse))


def gdn_reparam_fn(p, bound, pedestal):
    return _ReparamFn.apply(p, float(bound), float(pedestal))


# ------------------------------------------------------------------------------------------------ elementwise
def add_act(a, b=None, act=ACT_NONE, slope=0.01, out=None):
    a = _c(a)
    if out is None:
        out = torch.empty_like(a)
    call("add_act", _p(a), _p(_c(b)), a.numel(), act, slope, _p(out))
    return out


def act_bwd(y, dy, act, slope):
    out = torch.empty_like(dy)
    call("act_bwd", _p(_c(y)), _p(_c(dy)), dy.numel(), act, slope, _p(out))
    return out


class _AddActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, act, slope):
        y = add_act(a, b, act, slope)
        ctx.act, ctx.slope, ctx.has_b = act, slope, b is not None
        ctx.save_for_backward(y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        g = act_bwd(y, dy, ctx.act, ctx.slope) if ctx.act != ACT_NONE else dy
        return g, (g if ctx.has_b else None), None, None


def add_act_fn(a, b=None, act=ACT_NONE, slope=0.01):
    return _AddActFn.apply(a, b, act, slope)


def gather_mix(q, fp, idx=None, prob=1.0, seed=0, mask=None, out=None):
    """Batch pick + QDrop mix: out[b] = keep ? q[idx[b]] : fp[idx[b]]  (layer_opt.py:289-292)."""
    q, fp = _c(q), _c(fp)
    rows = q.shape[0] if idx is None else idx.numel()
    row = q[0].numel()
    if idx is not None and (idx.dtype != torch.int64 or not idx.is_cuda):
        raise TypeError("gather_mix: idx must be a CUDA int64 tensor")
    if mask is not None:
        mask = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
        if not mask.is_cuda or mask.numel() != rows * row:
            raise ValueError("gather_mix: mask must be a CUDA tensor with rows*row elements")
        mask = mask.contiguous()
    if out is None:
        out = torch.empty((rows,) + tuple(q.shape[1:]), device=q.device, dtype=torch.float32)
    call("gather_mix", _p(q), _p(fp), _p(idx), rows, row, float(prob), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(mask),
         _p(out))
    return out


# ------------------------------------------------------------------------------------------------ device schedule
def new_sched(device):
    """Zero-initialised `b200lic_calib_sched` in device memory (4 x 32-bit words)."""
    return torch.zeros(4, dtype=torch.int32, device=device)


def sched_tick(sched, iters, warmup, b_start, b_end, lr=1e-3, beta1=0.9, beta2=0.999):
    call("calib_sched_tick", _p(sched), int(iters), float(warmup), float(b_start), float(b_end), lr, beta1, beta2)


def read_sched(sched):
    """Host copy of the schedule (synchronises): dict(step, lr_over_bc1, inv_sqrt_bc2, reg_b)."""
    raw = sched.cpu()
    f = raw.view(torch.float32)
    return dict(step=int(raw[0]), lr_over_bc1=float(f[1]), inv_sqrt_bc2=float(f[2]), reg_b=float(f[3]))


def adaround_bwd_adam_sched(w, alpha, delta, zp, d_wq, exp_avg, exp_avg_sq, axis, n_levels, sched, beta1=0.9,
                            beta2=0.999, eps=1e-8, grad_scale=1.0, reg_weight=0.0, reg_loss=None):
    outer, ch, inner = channel_view(w.shape, axis)
    call("adaround_bwd_adam_sched", _p(_c(w)), _p(_c(alpha)), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))),
         _p(_c(d_wq)), _p(exp_avg), _p(exp_avg_sq), outer, ch, inner, n_levels, _p(sched), beta1, beta2, eps,
         grad_scale, reg_weight, _p(reg_loss))


def xgpu_reduce_adam_sched(peer, w, delta, zp, exp_avg, exp_avg_sq, axis, n_levels, sched, beta1=0.9, beta2=0.999,
                           eps=1e-8, grad_scale=1.0, reg_weight=0.0, reg_loss=None, exit_barrier=True):
    """Multi-GPU tail over NVLink peer memory (dist.PeerLayer): shard-wise sum of the ranks' gradients + STE masks +
    regulariser + Adam + broadcast of the new alpha, one kernel, no collective launch.  exit_barrier=False: the caller
    guarantees that another call (any layer) precedes the next use of this layer's alpha / gradient buffer on every rank."""
    outer, ch, inner = channel_view(w.shape, axis)
    call("xgpu_reduce_adam_sched", _p(peer.grad_ptrs), _p(peer.alpha_ptrs), _p(peer.flag_ptrs), _p(peer.state), peer.rank,
         peer.world, peer.lo, peer.hi, _p(_c(w)), _p(_c(delta.reshape(-1))), _p(_c(zp.reshape(-1))), _p(exp_avg),
         _p(exp_avg_sq), outer, ch, inner, int(n_levels), _p(sched), beta1, beta2, eps, float(grad_scale),
         float(reg_weight), _p(reg_loss), int(bool(exit_barrier)))


def gather_mix_sched(q, fp, idx_table, rows, prob, seed_base, units, unit, sched, out=None):
    """gather_mix whose batch pick (row `k % table_rows` of the int64 `idx_table`, None = identity) and QDrop seed
    follow the device schedule: k = (step - 1) * units + unit."""
    q, fp = _c(q), _c(fp)
    row = q[0].numel()
    if idx_table is not None:
        if idx_table.dtype != torch.int64 or not idx_table.is_cuda or idx_table.dim() != 2 or idx_table.size(1) != rows:
            raise TypeError("gather_mix_sched: idx_table must be a CUDA int64 [table_rows, rows] tensor")
        idx_table = idx_table.contiguous()
    if out is None:
        out = torch.empty((rows,) + tuple(q.shape[1:]), device=q.device, dtype=torch.float32)
    call("gather_mix_sched", _p(q), _p(fp), _p(idx_table), 0 if idx_table is None else idx_table.size(0), rows, row,
         float(prob), int(seed_base) & 0xFFFFFFFFFFFFFFFF, int(units), int(unit), _p(sched), _p(out))
    return out


def attn_gate(a, b, c):
    a = _c(a)
    out = torch.empty_like(a)
    call("attn_gate", _p(a), _p(_c(b)), _p(_c(c)), a.numel(), _p(out))
    return out


def abs_(x):
    x = _c(x)
    out = torch.empty_like(x)
    call("abs", _p(x), x.numel(), _p(out))
    return out


class _AbsFn(torch.autograd.Function):
    """|x| with gradient sign(x)*g (the LeakyReLU-derivative kernel with slope -1; differs from torch.abs only at
    exactly 0, where torch returns 0)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return abs_(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return act_bwd(x, g, ACT_LEAKY_RELU, -1.0)


def abs_fn(x):
    return _AbsFn.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else abs_(x)


class _PixelShuffleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, r, act, slope):
        x = _c(x)
        N, Crr, H, W = x.shape
        Cc = Crr // (r * r)
        out = torch.empty((N, Cc, H * r, W * r), device=x.device, dtype=torch.float32)
        call("pixel_shuffle", _p(x), N, Cc, H, W, r, act, slope, _p(out))
        ctx.r, ctx.act, ctx.slope, ctx.shape = r, act, slope, (N, Cc, H, W)
        ctx.save_for_backward(out if act != ACT_NONE else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        if ctx.act != ACT_NONE:
            dy = act_bwd(y, dy, ctx.act, ctx.slope)
        N, Cc, H, W = ctx.shape
        dx = torch.empty((N, Cc * ctx.r * ctx.r, H, W), device=dy.device, dtype=torch.float32)
        call("pixel_unshuffle", _p(dy), N, Cc, H, W, ctx.r, _p(dx))
        return dx, None, None, None


def pixel_shuffle(x, r, act=ACT_NONE, slope=0.01):
    return _PixelShuffleFn.apply(x, r, act, slope)
