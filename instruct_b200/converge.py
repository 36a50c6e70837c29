"""Gelman-Rubin convergence check on the gathered log-likelihood traces (host logic).

``gelman_rubin`` is the statistic the reference's comment block describes
(check_converg.c:100-153): m chains of n retained draws, R = V/W with
V = W (n-1)/n + B/n.  ``gelman_rubin_ref_compat`` reproduces what the reference actually
computes: chain_converg passes totrep = ckrep although the trace holds n_chain*ckrep
values (check_converg.c:67), so it compares n_chain consecutive segments of CHAIN 0 and
ignores the other chains.  The CLI reports the former and can print
the latter for byte parity.

SURVEY.md section 8f rank 4 asks for more than the log-likelihood: ``gelman_rubin_params`` gives R per
parameter for traces of S, Q or anything else, and ``align_labels`` removes the label switching between
chains first (clusters are exchangeable, so chain c's cluster 2 may be chain 0's cluster 0).
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


def gelman_rubin_params(traces: np.ndarray) -> np.ndarray:
    """traces: [n_chain][n_draws][...]; returns R for every trailing index (nan where a parameter never moved)."""
    t = np.asarray(traces, dtype=np.float64)
    if t.ndim < 3 or t.shape[0] < 2 or t.shape[1] < 2:
        raise ValueError("need [n_chain >= 2][n_draws >= 2][params...]")
    m, n = t.shape[:2]
    mu = t.mean(axis=1)
    W = (((t - mu[:, None]) ** 2).sum(axis=1) / (n - 1)).mean(axis=0)
    B = n * ((mu - mu.mean(axis=0)) ** 2).sum(axis=0) / (m - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(W > 0, (W * (n - 1) / n + B / n) / W, np.nan)


def align_labels(q_ref: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Permutation p of the clusters of one chain such that q[:, p] matches q_ref ([N][K] posterior-mean admixture
    proportions) best in squared error: exhaustive for K <= 7 (5040 permutations), greedy beyond."""
    import itertools

    q_ref, q = np.asarray(q_ref, dtype=np.float64), np.asarray(q, dtype=np.float64)
    K = q_ref.shape[1]
    cost = ((q_ref[:, :, None] - q[:, None, :]) ** 2).sum(axis=0)          # cost[a][b]: ref cluster a against cluster b
    if K <= 7:
        best = min(itertools.permutations(range(K)), key=lambda p: sum(cost[a, p[a]] for a in range(K)))
        return np.array(best)
    perm, free = np.full(K, -1), set(range(K))
    for a in np.argsort(cost.min(axis=1)):
        b = min(free, key=lambda j: cost[a, j])
        perm[a] = b
        free.remove(b)
    return perm


def chain_diagnostics(ll: np.ndarray, S: np.ndarray, Q: np.ndarray) -> dict:
    """ll [n_chain][n], S [n_chain][n][K] (population rates) and Q [n_chain][n][N][K] traces of several chains:
    aligns every chain's clusters to chain 0 and returns R for the log-likelihood, every S_k and the cluster
    sizes sum_i Q_ik."""
    ll, S, Q = np.asarray(ll, float), np.asarray(S, float), np.asarray(Q, float)
    S, Q = S.copy(), Q.copy()
    perms = [np.arange(Q.shape[3])]
    for c in range(1, Q.shape[0]):
        p = align_labels(Q[0].mean(axis=0), Q[c].mean(axis=0))
        perms.append(p)
        Q[c] = Q[c][:, :, p]
        if S.shape[2] == Q.shape[3]:
            S[c] = S[c][:, p]
    return {"R_loglik": gelman_rubin(ll), "R_S": gelman_rubin_params(S), "R_cluster_size": gelman_rubin_params(Q.sum(axis=2)),
            "permutations": np.array(perms)}
