"""Shared inputs of the entropy-coding tests: the 64-scale Gaussian tables (built once) and symbol streams with escapes."""
import numpy as np

from oracle import rans as R

_GC = []


def gaussian_tables():
    if not _GC:                                                  # the pure-Python stealing loop takes seconds: build once
        st = R.get_scale_table()
        _GC.append((st,) + tuple(R.gc_tables(st)))
    return _GC[0]


def _case(n, seed, escapes=True):
    g = np.random.default_rng(seed)
    st, cdf, cdf_len, off = gaussian_tables()
    idx = g.integers(0, 64, n).astype(np.int32)
    sym = np.rint(g.standard_normal(n) * st.numpy()[idx]).astype(np.int32)
    if escapes and n > 8:
        where = g.integers(0, n, max(1, n // 50))
        sym[where] = (g.integers(-1, 2, len(where)) * g.integers(1, 1 << 20, len(where))).astype(np.int32)
    return sym, idx, cdf, cdf_len, off
