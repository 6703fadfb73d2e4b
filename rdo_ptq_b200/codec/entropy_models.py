"""compressai-style entropy models on libb200lic kernels (K9 Gaussian conditional, K10 factorised prior).

Evaluation-mode semantics of compressai 1.2.4 (`quantize(..., "dequantize")`, likelihood lower bound 1e-9, scale
lower bound 0.11), reached in the reference from task-oriented-PTQ/models/nic_cvt.py:297-308 and
quant_model.py:72-79.  Each forward also leaves `last_bits` = sum(-log2 likelihood) (device scalar) so the
bpp reduction (losses/losses.py:20-23) needs no second pass over the likelihood tensor.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from . import coding
from .layers import LowerBound


class EntropyBottleneck(nn.Module):
    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), likelihood_bound=1e-9):
        super().__init__()
        self.channels, self.filters = int(channels), tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("K10 is specialised for compressai's default filters (3,3,3,3)")
        self.init_scale, self.tail_mass = float(init_scale), float(tail_mass)
        self.likelihood_bound = float(likelihood_bound)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / f[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, f[i + 1], f[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, f[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(channels, 1, 1))
        t = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-t, 0, t]))
        self.last_bits = None
        # additive switch for the R + lambda*D task criterion: straight-through latent rounding (round_ste,
        # quantizer.py:64-68) instead of compressai's zero-gradient torch.round; forward values are identical
        self.ste_round = False
        self._prior_cache = None         # (key, packed params, medians, symbol tables): see _prior()

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_prior_cache"] = None     # derived from the parameters; rebuilt on first use after un-pickling
        state.pop("_tables", None)       # coding tables: rebuilt by update()
        return state

    def _prior(self):
        """Packed parameters, medians and per-channel symbol tables (b200lic_factorized_table) of the prior.  PTQ never
        trains the prior, so they are built once and reused by every forward; the key (storage address and in-place
        version of every parameter) rebuilds them after load_state_dict / .to() / an optimiser step."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.likelihood_bound,)
        if self._prior_cache is None or self._prior_cache[0] != key:
            packed = self.packed_params()
            med = self._get_medians().detach().reshape(-1).contiguous()
            self._prior_cache = (key, packed, med, ops.factorized_table(packed, med, self.likelihood_bound))
        return self._prior_cache[1:]

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def packed_params(self):
        """[C,58] = matrices 3+9+9+9+3 | biases 3+3+3+3+1 | factors 3+3+3+3 (layout of b200lic_factorized_lik_fwd)."""
        C_ = self.channels
        parts = [getattr(self, f"_matrix{i:d}").detach().reshape(C_, -1) for i in range(5)]
        parts += [getattr(self, f"_bias{i:d}").detach().reshape(C_, -1) for i in range(5)]
        parts += [getattr(self, f"_factor{i:d}").detach().reshape(C_, -1) for i in range(4)]
        return torch.cat(parts, dim=1).contiguous()

    def quantize(self, inputs, mode, means=None):
        if mode != "dequantize":
            raise NotImplementedError("only the evaluation-mode quantiser is on the hot path")
        if self.ste_round and torch.is_grad_enabled() and inputs.requires_grad:
            return ops.round_latent_ste(inputs, means)
        return ops.round_latent(inputs, means)

    def forward(self, x, training=None):
        training = self.training if training is None else training
        if training:
            raise NotImplementedError("EntropyBottleneck: the additive-noise training path is not on the PTQ hot path "
                                      "(call .eval(); the reference evaluates under model.eval())")
        packed, med, table = self._prior()
        if torch.is_grad_enabled() and x.requires_grad:
            z_hat, lik, bits = ops.factorized_lik_fn(x, packed, med, self.likelihood_bound, self.ste_round, table)
        else:
            z_hat, lik, bits = ops.factorized_lik(x, packed, med, self.likelihood_bound, table=table)
        self.last_bits = bits
        return z_hat, lik

    # -- real entropy coding (compressai EntropyBottleneck.update / compress / decompress) -------------
    def update(self, force=False):
        """Build the quantised CDF of every channel from its quantiles (compressai EntropyBottleneck.update)."""
        if getattr(self, "_tables", None) is not None and not force:
            return False
        pmf, tail, length, offset = coding.bottleneck_pmf(self)
        cdf = coding.quantized_cdf_rows(pmf, tail, length)
        self._tables = coding.Tables(cdf, length + 2, offset, self.quantiles.device)
        return True

    def _coding_tables(self):
        if getattr(self, "_tables", None) is None:
            raise RuntimeError("EntropyBottleneck: call update() before compress() / decompress()")
        if self._tables.cdf.device != self.quantiles.device:
            self._tables = coding.Tables(*self._tables.host, self.quantiles.device)
        return self._tables

    def compress(self, x, chunk=coding.DEFAULT_CHUNK):
        """[B,C,H,W] -> one string per image: symbols = round(x - median), table row = channel."""
        t = self._coding_tables()
        med = self._get_medians().detach().reshape(-1)
        sym, idx = coding.symbols_and_indexes(x, means=med)
        return [coding.encode(sym[b], idx[b], t, chunk) for b in range(x.shape[0])]

    def decompress(self, strings, size):
        t = self._coding_tables()
        dev = self.quantiles.device
        H, W = int(size[0]), int(size[1])
        idx = torch.arange(self.channels, dtype=torch.int32, device=dev).view(-1, 1, 1).expand(-1, H, W).contiguous()
        med = self._get_medians().detach().reshape(1, -1, 1, 1)
        out = torch.stack([coding.decode(s, idx, t) for s in strings])
        return out.to(torch.float32) + med

    def loss(self):
        """aux loss |logits(quantiles) - target| (quant_model.py:72-79).  Parameter-sized (C x 3) host-side helper,
        not a data-path op; evaluated with plain tensor ops."""
        import torch.nn.functional as F
        v = self.quantiles
        for i in range(5):
            v = torch.matmul(F.softplus(getattr(self, f"_matrix{i:d}").detach()), v) + getattr(self, f"_bias{i:d}").detach()
            if i < 4:
                v = v + torch.tanh(getattr(self, f"_factor{i:d}").detach()) * torch.tanh(v)
        return torch.abs(v - self.target).sum()


class GaussianConditional(nn.Module):
    def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9, likelihood_bound=1e-9):
        super().__init__()
        self.scale_bound, self.likelihood_bound = float(scale_bound), float(likelihood_bound)
        self.lower_bound_scale = LowerBound(scale_bound)
        self.last_bits = None
        self.ste_round = False           # see EntropyBottleneck.ste_round
        self.tail_mass = float(tail_mass)
        self.scale_table = None if scale_table is None else torch.as_tensor(scale_table, dtype=torch.float32)
        self._tables = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_tables"] = None
        return state

    # -- real entropy coding (compressai GaussianConditional.update_scale_table / build_indexes / compress / decompress) ----
    def update_scale_table(self, scale_table, force=False):
        if self._tables is not None and not force:
            return False
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32)
        pmf, tail, length, offset = coding.gaussian_pmf(self.scale_table, self.tail_mass)
        cdf = coding.quantized_cdf_rows(pmf, tail, length)
        self._tables = coding.Tables(cdf, length + 2, offset, "cpu")
        return True

    def _coding_tables(self, device):
        if self._tables is None:
            raise RuntimeError("GaussianConditional: call update_scale_table() before compress() / decompress()")
        if self._tables.cdf.device != torch.device(device):
            self._tables = coding.Tables(*self._tables.host, device)
            self.scale_table = self.scale_table.to(device)
        return self._tables

    def build_indexes(self, scales):
        self._coding_tables(scales.device)
        return coding.symbols_and_indexes(scales, scales=scales, scale_table=self.scale_table, bound=self.scale_bound)[1]

    def compress(self, inputs, indexes, means=None, chunk=coding.DEFAULT_CHUNK):
        t = self._coding_tables(inputs.device)
        sym, _ = coding.symbols_and_indexes(inputs, means=means)
        return [coding.encode(sym[b], indexes[b], t, chunk) for b in range(inputs.shape[0])]

    def decompress(self, strings, indexes, means=None):
        t = self._coding_tables(indexes.device)
        out = torch.stack([coding.decode(s, indexes[b], t) for b, s in enumerate(strings)]).to(torch.float32)
        return out if means is None else out + means

    def quantize(self, inputs, mode, means=None):
        if mode != "dequantize":
            raise NotImplementedError("only the evaluation-mode quantiser is on the hot path")
        if self.ste_round and torch.is_grad_enabled() and inputs.requires_grad:
            return ops.round_latent_ste(inputs, means)
        return ops.round_latent(inputs, means)

    def forward(self, inputs, scales, means=None, training=None):
        training = self.training if training is None else training
        if training:
            raise NotImplementedError("GaussianConditional: the additive-noise training path is not on the PTQ hot "
                                      "path (call .eval())")
        if torch.is_grad_enabled() and (inputs.requires_grad or scales.requires_grad or
                                        (means is not None and means.requires_grad)):
            y_hat, lik, bits = ops.gaussian_lik_fn(inputs, scales, means, self.scale_bound, self.likelihood_bound,
                                                   self.ste_round)
        else:
            y_hat, lik, bits = ops.gaussian_lik(inputs, scales, means, self.scale_bound, self.likelihood_bound)
        self.last_bits = bits
        return y_hat, lik
