"""CPU: the entropy-coding oracle (oracle/rans.py) against its own invariants, and the library's host-side
table build (b200lic_pmf_to_quantized_cdf) against the oracle.  compressai is not installed, so no byte-level fixture of
the reference's files exists: parity of this row is pinned by round trips, CDF invariants and code length (oracle header)."""
import numpy as np
import pytest
import torch

from oracle import rans as R


def _check_cdf(c, n):
    assert c[0] == 0 and c[n] == 1 << 16
    assert np.all(np.diff(c[:n + 1]) >= 1)


@pytest.mark.parametrize("seed", range(4))
def test_quantized_cdf_keeps_every_symbol_codable(seed):
    g = np.random.default_rng(seed)
    n = int(g.integers(3, 300))
    pmf = g.random(n).astype(np.float32) ** 8                 # many near-zero entries: the stealing loop runs
    pmf[g.integers(0, n, n // 3)] = 0.0
    pmf /= max(pmf.sum(), 1e-9)
    c = R.pmf_to_quantized_cdf(pmf)
    _check_cdf(c, n)
    big = np.argsort(pmf)[-3:]
    assert np.all(np.abs(np.diff(c)[big] / 65536.0 - pmf[big]) < 0.02)


def test_gaussian_tables_and_library_table_build_agree():
    from rdo_ptq_b200.codec import coding
    st, cdf, cdf_len, off = gaussian_tables()
    pmf, tail, length, offset = R.gc_pmf(st)
    assert cdf.shape == (64, int(length.max()) + 2) and np.array_equal(off, offset)
    for r in range(64):
        _check_cdf(cdf[r], int(cdf_len[r]) - 1)
    p2, t2, l2, o2 = coding.gaussian_pmf(st)
    assert np.array_equal(p2, pmf) and np.array_equal(t2, tail) and np.array_equal(l2, length) and np.array_equal(o2, offset)
    lib_cdf = coding.quantized_cdf_rows(p2, t2, l2)             # host function of libb200lic: runs without a GPU
    assert np.array_equal(lib_cdf, cdf)


def test_bottleneck_tables_and_library_table_build_agree():
    from rdo_ptq_b200.codec import coding
    from rdo_ptq_b200.codec.entropy_models import EntropyBottleneck
    torch.manual_seed(3)
    eb = EntropyBottleneck(24)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= torch.rand(24) * 20
        eb.quantiles[:, 0, 2] += torch.rand(24) * 30
    cdf, cdf_len, offset, med = R.eb_tables(eb)
    for r in range(24):
        _check_cdf(cdf[r], int(cdf_len[r]) - 1)
    p, t, l, o = coding.bottleneck_pmf(eb)
    assert np.array_equal(o, offset) and np.array_equal(l + 2, cdf_len)
    assert np.array_equal(coding.quantized_cdf_rows(p, t, l), cdf)


def test_build_indexes_is_the_count_of_larger_table_entries():
    st = R.get_scale_table().numpy()
    g = np.random.default_rng(0)
    s = np.concatenate([g.random(500).astype(np.float32) * 300, st, st * np.float32(1.0000001), [0.0, 0.05, 1e6]]).astype(np.float32)
    idx = R.build_indexes(s, st)
    sb = np.maximum(s, np.float32(0.11))
    want = np.array([min(int(np.searchsorted(st, v, side="left")), 63) for v in sb])
    assert np.array_equal(idx, want) and idx.min() == 0 and idx.max() == 63


from _rans_cases import _case, gaussian_tables  # noqa: E402


@pytest.mark.parametrize("n", [0, 1, 2, 17, 1000])
def test_sequential_round_trip_with_escapes(n):
    sym, idx, cdf, cdf_len, off = _case(n, 7 + n)
    words = R.rans64_encode(sym, idx, cdf, cdf_len, off)
    assert len(words) >= 2
    out = R.rans64_decode(words, n, idx, cdf, cdf_len, off)
    assert np.array_equal(out, sym)


def test_code_length_tracks_the_model_entropy():
    sym, idx, cdf, cdf_len, off = _case(4000, 11, escapes=False)
    bits = R.ideal_bits(sym, idx, cdf, cdf_len, off)
    words = R.rans64_encode(sym, idx, cdf, cdf_len, off)
    assert 0 <= 32 * len(words) - bits < 64 + 32                 # flushed 64-bit state + one partial word


@pytest.mark.parametrize("chunk", [1, 64, 1000, 5000])
def test_chunked_container_round_trip(chunk):
    sym, idx, cdf, cdf_len, off = _case(1000, 5)
    blob = R.encode_chunked(sym, idx, cdf, cdf_len, off, chunk)
    assert np.array_equal(R.decode_chunked(blob, idx, cdf, cdf_len, off), sym)
    if chunk >= 1000:                                            # one chunk: the payload IS the sequential stream
        assert blob[16 + 8:] == R.rans64_encode(sym, idx, cdf, cdf_len, off).astype("<u4").tobytes()
