"""The algebra behind the sharded update_S_POP (ig_kernels.cu local_sums_kernel / spop_decide_kernel).

update_S_POP (mcmc.c:913-983) is K sequential Metropolis steps; step j proposes a new S_j and compares proposal()
(mcmc.c:1630) at the current vector with proposal() at the vector whose j-th entry is replaced.  Step j's PROPOSAL depends on
S_j alone, which only step j changes -- so all K proposals are known before the first decision, and the sums the decisions
need are indexed by the subset B of rates replaced so far: current = T[B], proposed = T[B | 1 << j].  The GPU path takes T
for all 2^K subsets in one launch over the local individuals, all-reduces it, and walks it.  Here the same walk is checked
against the sequential loop in numpy, with the sums in the same fixed point (2^-30 per term, -inf terms counted), over
shuffled and split orders of the individuals: the decisions and the integer table must not depend on either."""
import numpy as np

FX = 2.0 ** 30


def _log_geom(s, g):
    with np.errstate(divide="ignore", invalid="ignore"):
        l1 = np.log(1.0 - s)
        return np.where(g > 1, (g - 1) * np.log(s) + l1, l1)


def _fixed_sum(terms):
    fin = np.isfinite(terms)
    return int(np.rint(terms[fin] * FX).astype(np.int64).sum()), int((~fin).sum())


def _value(acc, ninf):
    return -np.inf if ninf else acc / FX


def _sum_at(Q, G, S):
    return _fixed_sum(_log_geom(Q @ S, G))


def _sequential(Q, G, S, props, us):
    S = S.copy()
    cur = _value(*_sum_at(Q, G, S))
    took = []
    for j in range(len(S)):
        Sp = S.copy(); Sp[j] = props[j]
        pl = _value(*_sum_at(Q, G, Sp))
        with np.errstate(invalid="ignore", over="ignore"):
            ratio = np.exp(pl - cur)
        ok = bool(np.isnan(ratio) or us[j] < min(1.0, ratio))        # MIN2(1, NaN) == 1, mcmc.h:10
        took.append(ok)
        if ok:
            S, cur = Sp, pl
    return S, took


def _table(Q, G, S, props, parts):
    K = len(S)
    T = np.zeros((1 << K, 2), dtype=object)
    for B in range(1 << K):
        Sb = np.where([(B >> k) & 1 for k in range(K)], props, S)
        acc = ninf = 0
        for idx in parts:                                            # one "rank" per part: integer sums add up exactly
            a, n = _sum_at(Q[idx], G[idx], Sb)
            acc += a; ninf += n
        T[B] = (acc, ninf)
    return T


def _walk(T, S, props, us):
    B, S = 0, S.copy()
    cur = _value(*T[0])
    took = []
    for j in range(len(S)):
        Bp = B | (1 << j)
        pl = _value(*T[Bp])
        with np.errstate(invalid="ignore", over="ignore"):
            ratio = np.exp(pl - cur)
        ok = bool(np.isnan(ratio) or us[j] < min(1.0, ratio))
        took.append(ok)
        if ok:
            B, cur = Bp, pl
            S[j] = props[j]
    return S, took


def _case(seed, N, K, edge=False):
    r = np.random.default_rng(seed)
    Q = r.dirichlet(np.full(K, 0.4), size=N)
    G = r.integers(1, 6, size=N)
    S = r.uniform(0.05, 0.95, size=K)
    props = np.clip(S + r.uniform(-0.05, 0.05, size=K), 0.0, 1.0)
    if edge:                                                         # a rate of exactly 0 / 1 and pure individuals: log 0 terms
        props[0] = 0.0
        S[K - 1] = 1.0
        Q[:3] = np.eye(K)[[0, K - 1, 0]]
        G[:3] = [3, 1, 1]
    us = r.uniform(size=K)
    return Q, G, S, props, us


def test_subset_walk_equals_the_sequential_steps():
    flips = 0
    for seed in range(40):
        Q, G, S, props, us = _case(seed, N=257, K=1 + seed % 6, edge=seed % 5 == 0)
        want_S, want_took = _sequential(Q, G, S, props, us)
        parts = np.array_split(np.random.default_rng(seed + 1000).permutation(len(G)), 1 + seed % 4)
        T = _table(Q, G, S, props, parts)
        got_S, got_took = _walk(T, S, props, us)
        assert got_took == want_took, seed
        assert np.array_equal(got_S, want_S), seed
        flips += sum(want_took)
    assert flips > 20                                                # the cases do accept and reject


def test_table_does_not_depend_on_the_split():
    Q, G, S, props, us = _case(7, N=500, K=4)
    whole = _table(Q, G, S, props, [np.arange(500)])
    for nparts in (2, 3, 8):
        parts = np.array_split(np.random.default_rng(nparts).permutation(500), nparts)
        split = _table(Q, G, S, props, parts)
        assert all(tuple(whole[B]) == tuple(split[B]) for B in range(16)), nparts
