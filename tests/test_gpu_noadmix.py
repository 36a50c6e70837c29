"""GPU parity for mode 0, the no-admixture model (mcmc_POP_no_admixture, mcmc.c:90-131;
SURVEY.md section 8f rank 1): tallies by cluster label bit-exact, log_ld_indv_K to 1e-6, the
label draw against its exact conditional, posterior summaries against the compiled reference."""
import os

import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-6
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _freq(o, rng):
    K = o.K
    f = rng.dirichlet(np.ones(o.Amax), size=(K, o.L))
    for l in range(o.L):
        a = o.allelenum[l]
        f[:, l, a:] = 0
        f[:, l, :a] /= f[:, l, :a].sum(axis=1, keepdims=True)
    return f.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("N,L,K,A,miss", [(300, 40, 2, 2, 0.0), (257, 33, 5, 6, 0.05), (64, 130, 8, 2, 0.1), (70, 21, 12, 3, 0.02)])
def test_mode0_pieces(N, L, K, A, miss):
    d = make_dataset(N=N, L=L, K=K, A=A, miss=miss, seed=3)
    sd = SeqData(d.x, d.allelenum, K, mode=0)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K, mode=0)
    rng = np.random.default_rng(5)
    o.zz[...] = rng.integers(0, K, size=N)
    o.freq[...] = _freq(o, rng)
    s.set(_lib.STATE_ITER, [1])
    s.set(_lib.STATE_P, o.freq)
    s.set(_lib.STATE_G, o.zz.astype(np.int32))            # mode 0: the G slot carries the cluster labels
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())      # update_P's zz branch, mcmc.c:825-831
    s.run_phase(_lib.PHASE_ZQ | _lib.PHASE_ALPHA)
    zz = s.get(_lib.STATE_G)
    assert zz.min() >= 0 and zz.max() < K
    o.zz[...] = zz
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    q = s.get(_lib.STATE_Q)
    assert np.array_equal(q, np.eye(K)[zz])
    lk = s.get(_lib.STATE_INDVLKH)
    want = np.array([o.log_ld_indv_K(i, zz[i]) for i in range(N)])
    assert np.max(np.abs(lk - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    tot = s.get(_lib.STATE_TOTALLKH)[0]
    assert abs(tot - want.sum()) <= RTOL * abs(want.sum())
    s.close()


def test_mode0_label_draw_matches_exact_conditional():
    """update_Z (mcmc.c:1094-1119): P(zz_i = k) = exp(ll_ik) / sum_m exp(ll_im), pooled chi-square."""
    N, L, K, A = 48, 3, 4, 4
    d = make_dataset(N=N, L=L, K=K, A=A, miss=0.0, seed=6)
    sd = SeqData(d.x, d.allelenum, K, mode=0)
    s = Sampler(sd, seed=9)
    o = Oracle(d.x, d.allelenum, K, mode=0)
    rng = np.random.default_rng(8)
    o.freq[...] = _freq(o, rng)
    s.set(_lib.STATE_ITER, [1])
    s.set(_lib.STATE_P, o.freq)
    ll = np.array([[o.log_ld_indv_K(i, k) for k in range(K)] for i in range(N)])
    p = np.exp(ll - ll.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    reps = 600
    cnt = np.zeros((N, K))
    for r in range(reps):
        s.set(_lib.STATE_ITER, [r + 1])
        s.run_phase(_lib.PHASE_ZQ)
        cnt[np.arange(N), s.get(_lib.STATE_G)] += 1
    exp = p * reps
    m = exp > 5
    chi2 = ((cnt - exp) ** 2 / np.maximum(exp, 1e-12))[m].sum()
    dof = m.sum() - N
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)
    s.close()


def test_mode0_chain_deterministic_and_graph_matches_direct():
    d = make_dataset(N=150, L=30, K=3, A=4, miss=0.03, seed=8, pure=True)
    sd = SeqData(d.x, d.allelenum, 3, mode=0)
    out = []
    for ug in (0, 2, 0):
        s = Sampler(sd, update=60, burnin=20, thinning=4, ckrep=4, seed=11, use_graph=ug)
        ch, cv = s.run_chain(0)
        out.append((ch, cv))
        s.close()
    a, b, c = out
    for x, y in ((a, b), (a, c)):
        assert x[0].totallkh == y[0].totallkh and np.array_equal(x[0].qq, y[0].qq) and np.array_equal(x[1], y[1])
    np.testing.assert_allclose(a[0].qq.sum(axis=1), 1.0, rtol=1e-12)      # CHAIN.z / steps
    assert np.isfinite(a[0].totallkh) and a[0].step == a[0].steps == 10


def _z(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(axis=0, ddof=1) / len(a) + b.var(axis=0, ddof=1) / len(b))
    return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-12)


def test_mode0_posterior_matches_reference_within_mcse():
    g = np.load(os.path.join(GOLD, "posterior_mode0.npz"))
    K = int(g["K"])
    R = g["LL"].shape[0]
    sd = SeqData(g["x"], g["allelenum"], K, mode=0)
    pop = g["pop"]
    LL, M = [], []
    for rep in range(R):
        s = Sampler(sd, update=int(g["update"]), burnin=int(g["burnin"]), thinning=int(g["thinning"]), ckrep=5, seed=7000 + rep)
        ch, _ = s.run_chain(rep)
        s.close()
        o = np.argsort(ch.qq[pop == 0].mean(axis=0))[::-1]
        LL.append(ch.totallkh)
        M.append([ch.qq[pop == p][:, o[0]].mean() for p in range(K)])
    LL, M = np.array(LL), np.array(M)
    refM = np.stack([g["Z"][:, pop == p, 0].mean(axis=1) for p in range(K)], axis=1).astype(np.float64)
    zLL = _z(LL[:, None], g["LL"][:, None])
    msg = f"zLL={zLL} LL={LL.mean()} ref={g['LL'].mean()} M={M.mean(0)} refM={refM.mean(0)}"
    assert np.all(np.abs(zLL) < 3.5), msg
    assert abs(LL.mean() - g["LL"].mean()) < 5.0, msg
    assert np.all(np.abs(M.mean(0) - refM.mean(0)) < 0.01), msg
