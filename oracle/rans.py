"""TEST INFRASTRUCTURE -- CPU restatement of the entropy-coding step that follows the hot path.

PARITY UNPINNED.  The algorithm lives in the reference's third-party dependency compressai==1.2.4 (requirements.txt),
which is not vendored under /root/reference and not installed here; its published algorithm is restated from the
reference's call sites (task-oriented-PTQ/models/nic_cvt.py:426-570 `compress` / `decompress`,
light-uniform-PTQ/models/tinylic.py:236-367, light-uniform-PTQ/dataset_test.py:159-180) and from the library's
documented behaviour:

  * `pmf_to_quantized_cdf`  (compressai/cpp_exts/ops/ops.cpp): 16-bit CDF with every symbol kept codable
  * `EntropyBottleneck.update`, `GaussianConditional.update_scale_table` / `update` / `build_indexes`
    (compressai/entropy_models/entropy_models.py): PMF supports, tail mass, offsets
  * `BufferedRansEncoder.encode_with_indexes` / `RansDecoder.decode_with_indexes` (compressai/cpp_exts/rans/
    rans_interface.cpp on ryg_rans rans64.h): 64-bit state, 32-bit renormalisation, 16-bit precision, out-of-range symbols
    escaped through the last CDF entry and sent as 4-bit bypass digits

so what is pinned here is self-consistency (round trips, CDF invariants, code length against the model's entropy), not
byte equality with compressai's files.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.

The product's stream is this coder applied to fixed-size CHUNKS of the symbol sequence (one independent rANS stream per
chunk, one GPU thread each) behind a small offset table: `encode_chunked` / `decode_chunked` restate that container.
With a single chunk the payload is the sequential stream itself.
"""
import math
import struct

import numpy as np
import torch

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
MAGIC = 0x41523242           # "B2RA"


# ---------------------------------------------------------------------------------------------------------------------
# CDF tables
def pmf_to_quantized_cdf(pmf, precision=PRECISION):
    """ops.cpp pmf_to_quantized_cdf: round(p * 2^precision), renormalise to the total, prefix-sum, then give every
    zero-width symbol one count stolen from the least frequent symbol that can spare it."""
    pmf = np.asarray(pmf, dtype=np.float32)
    cdf = [0] + [int(math.floor(float(np.float32(p) * np.float32(1 << precision)) + 0.5)) for p in pmf]   # std::round, p >= 0
    total = sum(cdf)
    if total == 0:
        raise ValueError("pmf sums to zero")
    cdf = [((1 << precision) * c) // total for c in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(len(cdf) - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best = None, -1
            for j in range(len(cdf) - 1):
                f = cdf[j + 1] - cdf[j]
                if f > 1 and (best_freq is None or f < best_freq):
                    best_freq, best = f, j
            assert best != -1
            if best < i:
                for j in range(best + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best + 1):
                    cdf[j] += 1
    return np.asarray(cdf, dtype=np.int32)


def pmf_rows_to_cdf(pmf, tail_mass, pmf_length, max_length):
    """EntropyModel._pmf_to_cdf: one row per channel / scale: [pmf[:len], tail_mass] -> cdf, zero padded to max_length + 2."""
    out = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
    for i, n in enumerate(pmf_length):
        prob = np.concatenate([np.asarray(pmf[i][:n], np.float32), np.asarray(tail_mass[i], np.float32).reshape(1)])
        c = pmf_to_quantized_cdf(prob)
        out[i, :len(c)] = c
    return out


def _logits_cumulative(eb, v):
    """EntropyBottleneck._logits_cumulative on [C,1,L] inputs (parameters detached, CPU fp32)."""
    import torch.nn.functional as F
    for i in range(5):
        m = getattr(eb, f"_matrix{i:d}").detach().float().cpu()
        v = torch.matmul(F.softplus(m), v) + getattr(eb, f"_bias{i:d}").detach().float().cpu()
        if i < 4:
            v = v + torch.tanh(getattr(eb, f"_factor{i:d}").detach().float().cpu()) * torch.tanh(v)
    return v


def eb_pmf(eb):
    """EntropyBottleneck.update(): the PMF support of every channel from its quantiles -> (pmf [C,L], tail_mass [C],
    pmf_length [C], offset [C], medians [C])."""
    q = eb.quantiles.detach().float().cpu()
    medians = q[:, 0, 1]
    minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
    maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
    offset = -minima
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length).float()[None, None, :] + pmf_start[:, None, None]
    lower = _logits_cumulative(eb, samples - 0.5)
    upper = _logits_cumulative(eb, samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = (torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:]))[:, 0]
    # the tail of channel c is read at ITS last sample in compressai only when all lengths agree; it evaluates
    # upper[:, 0, -1:] over the padded support exactly as written here
    return pmf.numpy(), tail.numpy(), pmf_length.numpy(), offset.numpy(), medians.numpy()


def eb_tables(eb):
    pmf, tail, length, offset, medians = eb_pmf(eb)
    return pmf_rows_to_cdf(pmf, tail, length, int(length.max())), (length + 2).astype(np.int32), offset.astype(np.int32), medians


def get_scale_table(lo=0.11, hi=256.0, levels=64):
    """compressai.models.google.get_scale_table"""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def _std_cumulative(x):
    return 0.5 * torch.erfc(-(2 ** -0.5) * x)


def gc_pmf(scale_table, tail_mass=1e-9):
    """GaussianConditional.update(): zero-mean Gaussian PMF per table scale -> (pmf, tail_mass, pmf_length, offset)."""
    from scipy.stats import norm
    st = torch.as_tensor(scale_table, dtype=torch.float32)
    multiplier = -float(norm.ppf(tail_mass / 2))
    center = torch.ceil(st * multiplier).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    scale = st.unsqueeze(1)
    upper = _std_cumulative((0.5 - samples) / scale)
    lower = _std_cumulative((-0.5 - samples) / scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    return pmf.numpy(), tail[:, 0].numpy(), length.numpy(), (-center).numpy()


def gc_tables(scale_table, tail_mass=1e-9):
    pmf, tail, length, offset = gc_pmf(scale_table, tail_mass)
    return pmf_rows_to_cdf(pmf, tail, length, int(length.max())), (length + 2).astype(np.int32), offset.astype(np.int32)


def build_indexes(scales, scale_table, bound=0.11):
    """GaussianConditional.build_indexes: number of table entries (but the last) that are >= the bounded scale, counted
    from the top: index = levels - 1 - #{s in table[:-1] : scale <= s}."""
    s = np.maximum(np.asarray(scales, np.float32), np.float32(bound))
    idx = np.full(s.shape, len(scale_table) - 1, dtype=np.int32)
    for t in np.asarray(scale_table, np.float32)[:-1]:
        idx -= (s <= t).astype(np.int32)
    return idx


def symbols_of(x, means=None):
    """EntropyModel.quantize(x, "symbols", means): round half to even of (x - means), int32."""
    x = np.asarray(x, np.float32)
    if means is not None:
        x = x - np.asarray(means, np.float32)
    return np.rint(x).astype(np.int32)


# ---------------------------------------------------------------------------------------------------------------------
# rans64 (ryg_rans) as driven by rans_interface.cpp
def _expand(symbols, indexes, cdf, cdf_len, offset):
    """encode_with_indexes, first half: (start, range, bypass) per coded item, in forward order."""
    items = []
    for s, k in zip(symbols, indexes):
        max_value = int(cdf_len[k]) - 2
        value = int(s) - int(offset[k])
        raw = 0
        if value < 0:
            raw = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw = 2 * (value - max_value)
            value = max_value
        items.append((int(cdf[k][value]), int(cdf[k][value + 1]) - int(cdf[k][value]), False))
        if value == max_value:
            n_bypass = 0
            while (raw >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS:
                items.append((MAX_BYPASS, 0, True))
                val -= MAX_BYPASS
            items.append((val, 0, True))
            for j in range(n_bypass):
                items.append(((raw >> (j * BYPASS_PRECISION)) & MAX_BYPASS, 0, True))
    return items


def rans64_encode(symbols, indexes, cdf, cdf_len, offset):
    """-> np.uint32 words of the stream (BufferedRansEncoder.flush order: state low word, state high word, then the
    renormalisation words in decoding order)."""
    x = RANS64_L
    out = []
    for start, rng, bypass in reversed(_expand(symbols, indexes, cdf, cdf_len, offset)):
        if bypass:                                  # Rans64EncPutBits(val, 4)
            freq, bits = 1 << (32 - BYPASS_PRECISION), BYPASS_PRECISION
            if x >= ((RANS64_L >> BYPASS_PRECISION) << 32):
                out.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << bits) | start
            del freq
        else:                                       # Rans64EncPut(start, range, 16)
            if x >= ((RANS64_L >> PRECISION) << 32) * rng:
                out.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
    out.append((x >> 32) & 0xFFFFFFFF)
    out.append(x & 0xFFFFFFFF)
    return np.asarray(out[::-1], dtype=np.uint32)


def rans64_decode(words, n, indexes, cdf, cdf_len, offset):
    """RansDecoder.decode_with_indexes -> int32 symbols."""
    words = [int(w) for w in words]
    x = words[0] | (words[1] << 32)
    pos = 2
    mask = (1 << PRECISION) - 1
    out = np.empty(n, dtype=np.int32)

    def get_bits():
        nonlocal x, pos
        v = x & MAX_BYPASS
        x >>= BYPASS_PRECISION
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        return v

    for i in range(n):
        k = int(indexes[i])
        c = cdf[k]
        max_value = int(cdf_len[k]) - 2
        cum = x & mask
        # first entry of the row that is > cum, minus one (std::find_if in the interface)
        s = 0
        while int(c[s + 1]) <= cum:
            s += 1
        start, rng = int(c[s]), int(c[s + 1]) - int(c[s])
        x = rng * (x >> PRECISION) + (x & mask) - start
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        value = s
        if value == max_value:
            val = get_bits()
            n_bypass = val
            while val == MAX_BYPASS:
                val = get_bits()
                n_bypass += val
            raw = 0
            for j in range(n_bypass):
                raw |= get_bits() << (j * BYPASS_PRECISION)
            value = raw >> 1
            if raw & 1:
                value = -value - 1
            else:
                value += max_value
        out[i] = value + int(offset[k])
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the product's container: header | chunk offsets | per-chunk rans64 streams
def encode_chunked(symbols, indexes, cdf, cdf_len, offset, chunk):
    symbols, indexes = np.asarray(symbols).reshape(-1), np.asarray(indexes).reshape(-1)
    n = len(symbols)
    n_chunks = (n + chunk - 1) // chunk
    streams = [rans64_encode(symbols[i * chunk:(i + 1) * chunk], indexes[i * chunk:(i + 1) * chunk], cdf, cdf_len, offset)
               for i in range(n_chunks)]
    offs = np.zeros(n_chunks + 1, dtype=np.uint32)
    for i, s in enumerate(streams):
        offs[i + 1] = offs[i] + len(s)
    head = struct.pack("<4I", MAGIC, n, chunk, n_chunks)
    body = np.concatenate(streams) if streams else np.zeros(0, np.uint32)
    return head + offs.tobytes() + body.astype("<u4").tobytes()


def decode_chunked(blob, indexes, cdf, cdf_len, offset):
    magic, n, chunk, n_chunks = struct.unpack_from("<4I", blob, 0)
    if magic != MAGIC:
        raise ValueError("not a chunked rANS stream")
    offs = np.frombuffer(blob, dtype="<u4", count=n_chunks + 1, offset=16)
    words = np.frombuffer(blob, dtype="<u4", offset=16 + 4 * (n_chunks + 1))
    indexes = np.asarray(indexes).reshape(-1)
    out = np.empty(n, dtype=np.int32)
    for i in range(n_chunks):
        lo, hi = i * chunk, min(n, (i + 1) * chunk)
        out[lo:hi] = rans64_decode(words[offs[i]:offs[i + 1]], hi - lo, indexes[lo:hi], cdf, cdf_len, offset)
    return out


def ideal_bits(symbols, indexes, cdf, cdf_len, offset):
    """Code length of the quantised model itself: sum -log2(freq / 2^16) (+ bypass digits), the figure the stream size is
    compared with."""
    bits = 0.0
    for _, rng, bypass in _expand(symbols, indexes, cdf, cdf_len, offset):
        bits += BYPASS_PRECISION if bypass else PRECISION - math.log2(rng)
    return bits
