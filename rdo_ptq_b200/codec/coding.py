"""Real entropy coding of the latents on libb200lic: table build, chunked rANS encode / decode.

The reference reaches this through compressai 1.2.4's `update()` / `compress()` / `decompress()` (task-oriented-PTQ/
models/nic_cvt.py:426-570, light-uniform-PTQ/models/tinylic.py:236-367, timed by light-uniform-PTQ/dataset_test.py:159-180);
the surface here keeps those names and argument meanings.  One string per image:

    uint32 magic "B2RA" | n_symbols | chunk | n_chunks | uint32 offset[n_chunks + 1] (words) | rans64 streams

Every chunk is a complete stream of the reference coder's format (ryg rans64, 16-bit CDFs, 4-bit bypass digits for symbols
outside a table's support), encoded and decoded by one GPU thread; `chunk` symbols per stream, 12 bytes of framing each.
"""
import ctypes as C
import math
import struct

import numpy as np
import torch

from .. import _lib, ops

MAGIC = 0x41523242
DEFAULT_CHUNK = 2048
SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256.0, 64


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS):
    """compressai.models.google.get_scale_table"""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


class Tables:
    """Quantised CDFs of one entropy model: cdf [rows, max_length + 2], cdf_length [rows], offset [rows] (int32)."""

    def __init__(self, cdf, cdf_length, offset, device):
        self.host = (np.ascontiguousarray(cdf, np.int32), np.ascontiguousarray(cdf_length, np.int32),
                     np.ascontiguousarray(offset, np.int32))
        self.cdf, self.cdf_length, self.offset = (torch.from_numpy(a).to(device) for a in self.host)
        self.stride = int(self.cdf.shape[1])


def quantized_cdf_rows(pmf, tail_mass, pmf_length):
    """EntropyModel._pmf_to_cdf: b200lic_pmf_to_quantized_cdf (host-side table build; exact integer arithmetic)."""
    pmf = np.ascontiguousarray(pmf, np.float32)
    tail = np.ascontiguousarray(tail_mass, np.float32).reshape(-1)
    length = np.ascontiguousarray(pmf_length, np.int32).reshape(-1)
    rows, max_length = pmf.shape
    out = np.zeros((rows, max_length + 2), np.int32)
    rc = _lib.lib().b200lic_pmf_to_quantized_cdf(pmf.ctypes.data_as(C.c_void_p), tail.ctypes.data_as(C.c_void_p),
                                                 length.ctypes.data_as(C.c_void_p), rows, max_length,
                                                 out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise _lib.B200LicError("pmf_to_quantized_cdf", rc, _lib.lib().b200lic_last_error_string().decode())
    return out


def symbols_and_indexes(x, means=None, scales=None, scale_table=None, bound=SCALES_MIN):
    """round(x - means) and the table row of every element (channel id, or the scale-table index of `scales`)."""
    x = ops._c(x.detach(), "latent")
    n = x.numel()
    Cc, HW = x.shape[1], x.shape[2] * x.shape[3]
    per_channel = 0
    if means is not None:
        means = means.detach()
        if means.numel() == Cc:
            means, per_channel = means.reshape(-1).contiguous(), 1
        else:
            means = means.expand_as(x).contiguous()
    if scales is not None:
        scales = scales.detach().expand_as(x).contiguous()
    sym = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    idx = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    ops.call("rans_symbols", ops._p(x), ops._p(means), per_channel, ops._p(scales), ops._p(scale_table),
             0 if scale_table is None else scale_table.numel(), float(bound), Cc, HW, n, ops._p(sym), ops._p(idx))
    return sym, idx


def encode(symbols, indexes, tables, chunk=DEFAULT_CHUNK):
    """int32 symbols + table rows of ONE image (device) -> bytes."""
    sym, idx = symbols.reshape(-1).contiguous(), indexes.reshape(-1).contiguous()
    n = sym.numel()
    n_chunks = (n + chunk - 1) // chunk
    sizes = torch.empty(n_chunks, dtype=torch.int32, device=sym.device)
    ops.call("rans_encode_sizes", ops._p(sym), ops._p(idx), n, chunk, ops._p(tables.cdf), ops._p(tables.cdf_length),
             ops._p(tables.offset), tables.stride, ops._p(sizes))
    offs = torch.zeros(n_chunks + 1, dtype=torch.int32, device=sym.device)
    torch.cumsum(sizes, 0, out=offs[1:])
    total = int(offs[-1].item())
    words = torch.empty(max(total, 1), dtype=torch.int32, device=sym.device)
    ops.call("rans_encode_write", ops._p(sym), ops._p(idx), n, chunk, ops._p(tables.cdf), ops._p(tables.cdf_length),
             ops._p(tables.offset), tables.stride, ops._p(offs), ops._p(words))
    return (struct.pack("<4I", MAGIC, n, chunk, n_chunks) + offs.cpu().numpy().astype("<u4").tobytes() +
            words[:total].cpu().numpy().astype("<u4").tobytes())


def decode(blob, indexes, tables):
    """bytes + the table row of every symbol (device int32) -> int32 symbols (device), shaped like `indexes`."""
    magic, n, chunk, n_chunks = struct.unpack_from("<4I", blob, 0)
    if magic != MAGIC:
        raise ValueError("not a b200lic rANS string")
    idx = indexes.reshape(-1).contiguous()
    if idx.numel() != n:
        raise ValueError(f"string holds {n} symbols, indexes describe {idx.numel()}")
    dev = idx.device
    payload = np.frombuffer(blob, dtype="<u4", offset=16)
    if payload.size < n_chunks + 1 or int(payload[n_chunks]) != payload.size - (n_chunks + 1):
        raise ValueError("truncated rANS string")
    buf = torch.from_numpy(payload.astype(np.int32)).to(dev)
    offs, words = buf[:n_chunks + 1], buf[n_chunks + 1:]
    out = torch.empty(n, dtype=torch.int32, device=dev)
    ops.call("rans_decode", ops._p(words), ops._p(offs), n, chunk, ops._p(idx), ops._p(tables.cdf),
             ops._p(tables.cdf_length), ops._p(tables.offset), tables.stride, ops._p(out))
    return out.view(indexes.shape)


# -- PMF supports (EntropyBottleneck.update / GaussianConditional.update), parameter-sized host math in CPU fp32 ------------
def bottleneck_pmf(eb):
    import torch.nn.functional as F
    q = eb.quantiles.detach().float().cpu()
    medians = q[:, 0, 1]
    minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
    maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    samples = torch.arange(int(pmf_length.max())).float()[None, None, :] + pmf_start[:, None, None]

    def logits(v):
        for i in range(5):
            v = torch.matmul(F.softplus(getattr(eb, f"_matrix{i:d}").detach().float().cpu()), v) + \
                getattr(eb, f"_bias{i:d}").detach().float().cpu()
            if i < 4:
                v = v + torch.tanh(getattr(eb, f"_factor{i:d}").detach().float().cpu()) * torch.tanh(v)
        return v

    lower, upper = logits(samples - 0.5), logits(samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = (torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:]))[:, 0]
    return pmf.numpy(), tail.numpy(), pmf_length.numpy(), (-minima).numpy()


def gaussian_pmf(scale_table, tail_mass=1e-9):
    from scipy.stats import norm
    st = torch.as_tensor(scale_table, dtype=torch.float32).cpu()
    center = torch.ceil(st * (-float(norm.ppf(tail_mass / 2)))).int()
    length = 2 * center + 1
    samples = torch.abs(torch.arange(int(length.max())).int() - center[:, None]).float()

    def cum(v):
        return 0.5 * torch.erfc(-(2 ** -0.5) * v)

    upper, lower = cum((0.5 - samples) / st.unsqueeze(1)), cum((-0.5 - samples) / st.unsqueeze(1))
    return (upper - lower).numpy(), (2 * lower[:, 0]).numpy(), length.numpy(), (-center).numpy()


def string_bytes(strings):
    """Total payload of a compress() result: sum of len(s) over every string of every latent."""
    return sum(len(s) for group in strings for s in group)
