"""Gelman-Rubin convergence check on the gathered log-likelihood traces (host logic).

``gelman_rubin`` is the statistic the reference's comment block describes
(check_converg.c:100-153): m chains of n retained draws, R = V/W with
V = W (n-1)/n + B/n.  ``gelman_rubin_ref_compat`` reproduces what the reference actually
computes: chain_converg passes totrep = ckrep although the trace holds n_chain*ckrep
values (check_converg.c:67), so it compares n_chain consecutive segments of CHAIN 0 and
ignores the other chains.  The CLI reports the former and can print
the latter for byte parity.
"""
from __future__ import annotations

import numpy as np


def _gr(segments: np.ndarray) -> float:
    m, n = segments.shape
    mu = segments.mean(axis=1)
    W = (((segments - mu[:, None]) ** 2).sum(axis=1) / (n - 1)).mean()
    B = n * ((mu - mu.mean()) ** 2).sum() / (m - 1)
    V = W * (n - 1) / n + B / n
    return float(V / W)


def gelman_rubin(traces: np.ndarray) -> float:
    """traces: [n_chain][ckrep]."""
    t = np.asarray(traces, dtype=np.float64)
    if t.ndim != 2 or t.shape[0] < 2 or t.shape[1] < 2:
        raise ValueError("need at least 2 chains of at least 2 draws")
    return _gr(t)


def gelman_rubin_ref_compat(convg_ld: np.ndarray, n_chain: int, ckrep: int) -> float:
    v = np.asarray(convg_ld, dtype=np.float64).reshape(-1)
    per = ckrep // n_chain
    return _gr(v[: n_chain * per].reshape(n_chain, per))
