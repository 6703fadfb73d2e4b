"""Imports the reference's OWN wrapper packages in this container.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`task-oriented-PTQ/quantization` and `light-uniform-PTQ/quant_int` import `compressai`, `timm` and `pytorch_msssim`,
none of which is installed (or installable: no network).  What they take from those packages is small:

  compressai.layers.gdn.GDN, compressai.layers.{MaskedConv2d, layers.ResidualBlock*, layers.subpel_conv3x3},
  compressai.entropy_models.{EntropyBottleneck, GaussianConditional}             -> `oracle.codec` (the restatement)
  compressai.ans.{BufferedRansEncoder, RansDecoder}                              -> placeholders (rANS is off the path)
  compressai.models.utils.update_registered_buffers                              -> placeholder
  timm.models.layers.{DropPath, to_2tuple, trunc_normal_}                        -> three-line equivalents
  pytorch_msssim.ms_ssim                                                         -> placeholder that raises

With these registered in `sys.modules` the UNMODIFIED reference files are imported from where they lie under
/root/reference (nothing is copied); `oracle/make_golden.py` then runs the reference's `QuantModel`, `QuantModule`,
blocks, `save_inp_oup_data`, `LossFunction`, `layer_reconstruction` and `block_reconstruction` on CPU and pins
`oracle.quant_wrap` / `oracle.calib` against them bit for bit (`tests/golden/wrap_ref.pt`).

Two run-time patches are needed to execute the reference's loops without a GPU; neither touches arithmetic:
  * `layer_opt.py:211` / `block_opt.py:211` hard-code `device = 'cuda'` -> `cpu_device()` rewrites that one string in
    `Tensor.to(...)` to 'cpu' while the loop runs;
  * the loops draw `torch.randperm` / `torch.rand_like` from the global generator (`layer_opt.py:289-292`) ->
    `replay_draws()` serves them from an `oracle.calib.DrawPlan`, so the oracle and the CUDA path can replay them.
This module only works where /root/reference exists (this container, not the GPU box).
"""
import contextlib
import importlib
import itertools
import os
import sys
import types

import torch
import torch.nn as nn

from . import codec

REF = "/root/reference"
TO_ROOT = os.path.join(REF, "task-oriented-PTQ")
LU_ROOT = os.path.join(REF, "light-uniform-PTQ")
_OWN_TOPLEVEL = ("quantization", "quant_int", "models", "losses", "utils", "datasets", "ckpts")


def available() -> bool:
    return os.path.isdir(TO_ROOT) and os.path.isdir(LU_ROOT)


class _DropPath(nn.Module):                      # timm.models.layers.DropPath (identity in eval mode)
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _unavailable(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is outside the hot path (SURVEY 8f); placeholder of oracle/_ref_shim.py")
    return f


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []                              # lets `import a.b` resolve through sys.modules
    sys.modules[name] = m
    return m


def install_dependencies():
    """Registers the stand-ins for compressai / timm / pytorch_msssim (idempotent)."""
    if "compressai" in sys.modules and getattr(sys.modules["compressai"], "_oracle_shim", False):
        return
    ans = _module("compressai.ans", BufferedRansEncoder=_unavailable("BufferedRansEncoder"),
                  RansDecoder=_unavailable("RansDecoder"))
    ent = _module("compressai.entropy_models", EntropyBottleneck=codec.EntropyBottleneck,
                  GaussianConditional=codec.GaussianConditional)
    gdn = _module("compressai.layers.gdn", GDN=codec.GDN)
    lay = _module("compressai.layers.layers", ResidualBlockWithStride=codec.ResidualBlockWithStride,
                  ResidualBlockUpsample=codec.ResidualBlockUpsample, ResidualBlock=codec.ResidualBlock,
                  subpel_conv3x3=codec.subpel_conv3x3, AttentionBlock=codec.AttentionBlock,
                  MaskedConv2d=codec.MaskedConv2d, conv3x3=codec.conv3x3, conv1x1=codec.conv1x1)
    layers = _module("compressai.layers", GDN=codec.GDN, MaskedConv2d=codec.MaskedConv2d, gdn=gdn, layers=lay,
                     ResidualBlockWithStride=codec.ResidualBlockWithStride,
                     ResidualBlockUpsample=codec.ResidualBlockUpsample, ResidualBlock=codec.ResidualBlock,
                     subpel_conv3x3=codec.subpel_conv3x3, AttentionBlock=codec.AttentionBlock)
    mutils = _module("compressai.models.utils", update_registered_buffers=_unavailable("update_registered_buffers"))
    cmodels = _module("compressai.models", utils=mutils)
    _module("compressai", ans=ans, entropy_models=ent, layers=layers, models=cmodels, _oracle_shim=True,
            __version__="1.2.4-restated")
    tl = _module("timm.models.layers", DropPath=_DropPath, to_2tuple=_to_2tuple,
                 trunc_normal_=torch.nn.init.trunc_normal_)
    _module("timm.models.layers.helpers", to_2tuple=_to_2tuple)
    tm = _module("timm.models", layers=tl)
    _module("timm", models=tm)
    _module("pytorch_msssim", ms_ssim=_unavailable("ms_ssim"))


def _purge():
    for k in list(sys.modules):
        if k.split(".")[0] in _OWN_TOPLEVEL and getattr(sys.modules[k], "__file__", None) \
                and str(sys.modules[k].__file__).startswith(REF):
            del sys.modules[k]


@contextlib.contextmanager
def _on_path(root):
    _purge()
    sys.path.insert(0, root)
    try:
        yield
    finally:
        sys.path.remove(root)


def import_task_oriented():
    """-> the reference's `quantization` package (QuantModel, QuantModule, BaseQuantBlock, layer_reconstruction,
    block_reconstruction) plus its submodules, imported unmodified from TO_ROOT."""
    install_dependencies()
    with _on_path(TO_ROOT):
        pkg = importlib.import_module("quantization")
        for sub in ("quantizer", "quant_layer", "quant_block", "quant_model", "utils", "layer_opt", "block_opt"):
            importlib.import_module(f"quantization.{sub}")
        losses = importlib.import_module("losses.losses")
    pkg.losses_module = losses
    return pkg


def import_light_uniform():
    """-> the reference's `quant_int` package (QuantModule, QuantModel, QuantCodingModel), unmodified from LU_ROOT."""
    install_dependencies()
    with _on_path(LU_ROOT):
        pkg = importlib.import_module("quant_int")
        for sub in ("quantizer", "quant_layer", "quant_model", "quant_coding_model"):
            importlib.import_module(f"quant_int.{sub}")
    return pkg


@contextlib.contextmanager
def cpu_device():
    """While active, `tensor.to('cuda')` means `tensor.to('cpu')` (layer_opt.py:211, block_opt.py:211)."""
    orig = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(v, str) and v.startswith("cuda")) else v for v in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return orig(self, *a, **k)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


@contextlib.contextmanager
def replay_draws(plan, unit_id, n, batch_size, prob):
    """Serves the loop's `torch.randperm(n)[:batch_size]` and `torch.rand_like(cur_inp)` (layer_opt.py:289-292) from
    `plan.draw(unit_id, it, ...)`: randperm returns the planned pick padded to a permutation, rand_like returns a
    tensor that is < prob exactly where the planned keep-mask is true."""
    orig_perm, orig_rand = torch.randperm, torch.rand_like
    it = itertools.count()
    state = {}

    def randperm(m, *a, **k):
        if k.get("generator") is not None:          # the plan's own draws
            return orig_perm(m, *a, **k)
        assert m == n, (m, n)
        state["it"] = next(it)
        idx, _ = plan.draw(unit_id, state["it"], n, batch_size, (), 1.0)       # the pick does not depend on the shape
        rest = torch.tensor([j for j in range(n) if j not in set(idx.tolist())], dtype=idx.dtype)
        return torch.cat([idx, rest])

    def rand_like(t, *a, **k):
        _, keep = plan.draw(unit_id, state["it"], n, batch_size, tuple(t.shape[1:]), prob)
        assert keep.shape == t.shape, (keep.shape, t.shape)
        return torch.where(keep, torch.zeros_like(t), torch.ones_like(t))

    torch.randperm, torch.rand_like = randperm, rand_like
    try:
        yield
    finally:
        torch.randperm, torch.rand_like = orig_perm, orig_rand
