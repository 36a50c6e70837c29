"""GPU parity for the inbreeding-coefficient modes:
mode 4 = mcmc_POP_inbreedcoff (mcmc.c:242), F per population; mode 5 = mcmc_INDV_inbreedcoff
(mcmc.c:386, uniform prior), F per individual.  Same three levels as test_gpu_parity.py:
integer work bit-exact, log-likelihood pieces on identical states to 1e-6, posterior summaries
against chains of the compiled reference (tests/golden/posterior_mode{4,5}.npz)."""
import os

import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-6
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inject(s, o, rng, F):
    K = o.K
    o.z[...] = rng.integers(0, K, size=o.z.shape)
    o.qq[...] = rng.dirichlet(np.ones(K) * 0.8, size=o.N)
    f = rng.dirichlet(np.ones(o.Amax), size=(K, o.L))
    for l in range(o.L):
        a = o.allelenum[l]
        f[:, l, a:] = 0
        f[:, l, :a] /= f[:, l, :a].sum(axis=1, keepdims=True)
    o.freq[...] = f.astype(np.float32).astype(np.float64)
    o.alpha = 0.9
    o.self_rates[...] = F
    s.set(_lib.STATE_ITER, [1])
    s.set(_lib.STATE_Z, o.z)
    s.set(_lib.STATE_Q, o.qq)
    s.set(_lib.STATE_P, o.freq)
    s.set(_lib.STATE_ALPHA, [o.alpha])
    s.set(_lib.STATE_S, o.self_rates)


SHAPES = [(300, 40, 2, 2, 0.0), (257, 33, 5, 6, 0.05), (64, 130, 8, 2, 0.1), (70, 21, 12, 3, 0.02)]


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_mode5_fused_pieces(N, L, K, A, miss):
    """update_F_IND + update_ZQ + cal_lkh in one pass: the per-individual old-Z ratio piece and
    the new-Z likelihood under F and F' against log_ld_F_indv (mcmc.c:1812) of the oracle."""
    d = make_dataset(N=N, L=L, K=K, A=A, miss=miss, seed=3)
    sd = SeqData(d.x, d.allelenum, K, mode=5)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K, mode=5)
    rng = np.random.default_rng(5)
    F = rng.uniform(0.02, 0.98, N)
    _inject(s, o, rng, F)
    z_old = o.z.copy()
    before = s.get(_lib.STATE_TALLY)
    s.run_phase(_lib.PHASE_UPDATE_S | _lib.PHASE_ZQ | _lib.PHASE_ALPHA)
    Fp = s.get(_lib.STATE_FPROP)
    assert np.all((Fp >= 0) & (Fp <= 1)) and np.all(np.abs(Fp - F) <= 0.05 + 1e-12)      # reflected random walk, mcmc.c:897-903
    ll_old = np.array([o.log_ld_F(F[i:i + 1], 0, i) for i in range(N)])
    ll_old_p = np.array([o.log_ld_F(Fp[i:i + 1], 0, i) for i in range(N)])
    z_new = s.get(_lib.STATE_Z)
    usable = ~(d.x < 0).any(axis=2)
    assert np.array_equal(z_new[~usable], z_old[~usable])
    o.z[...] = z_new
    assert np.array_equal(s.get(_lib.STATE_TALLY) - before, o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    parts = s.get(_lib.STATE_LLPARTS)
    ll_new = np.array([o.log_ld_F(F[i:i + 1], 0, i) for i in range(N)])
    ll_new_p = np.array([o.log_ld_F(Fp[i:i + 1], 0, i) for i in range(N)])
    assert np.max(np.abs(parts[:, 0] - (ll_old_p - ll_old)) / np.maximum(np.abs(ll_old), 1.0)) <= RTOL
    assert np.max(np.abs(parts[:, 1] + parts[:, 2] - ll_new) / np.maximum(np.abs(ll_new), 1.0)) <= RTOL
    assert np.max(np.abs(parts[:, 1] + parts[:, 3] - ll_new_p) / np.maximum(np.abs(ll_new_p), 1.0)) <= RTOL
    Fa = s.get(_lib.STATE_S)                       # the accepted coefficient is the old or the proposed one
    took = Fa == Fp
    assert np.all(took | (Fa == F))
    lk = s.get(_lib.STATE_INDVLKH)
    want = np.where(took, ll_new_p, ll_new)
    assert np.max(np.abs(lk - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    tot = s.get(_lib.STATE_TOTALLKH)[0]
    assert abs(tot - lk.sum()) <= 1e-9 * abs(tot) + len(lk) * 2.0 ** -24     # totallkh is a 2^-24 fixed-point sum
    s.close()


@pytest.mark.parametrize("back_refl", [1, 0])
@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_mode4_fused_pieces(N, L, K, A, miss, back_refl):
    """update_inbreedcoff_POP + update_ZQ + cal_lkh in one pass: per population, the old-Z
    difference summed over individuals is log_ld_F_total(F with F'_k) - log_ld_F_total(F)
    (mcmc.c:1038), and the likelihood held afterwards is log_ld_F_pop at the accepted F."""
    d = make_dataset(N=N, L=L, K=K, A=A, miss=miss, seed=4)
    sd = SeqData(d.x, d.allelenum, K, mode=4, back_refl=back_refl)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K, mode=4, back_refl=back_refl)
    rng = np.random.default_rng(6)
    F = rng.uniform(0.05, 0.95, K)
    _inject(s, o, rng, F)
    if back_refl == 0:
        s.set(_lib.STATE_STATE, np.ones(K, dtype=np.int32))
    before = s.get(_lib.STATE_TALLY)
    s.run_phase(_lib.PHASE_UPDATE_S | _lib.PHASE_ZQ | _lib.PHASE_ALPHA)
    Fp = s.get(_lib.STATE_FPROP)
    assert np.all((Fp >= 0) & (Fp <= 1))
    base_old = o.log_ld_F_total(F)
    fk = s.get(_lib.STATE_FK)                      # [N][2][K]
    D = fk[:, 0, :].sum(axis=0)
    for k in range(K):
        if not (0.0 < Fp[k] < 1.0):
            continue                               # -e 0 proposals at exactly 0 or 1: log 0 terms, compared below through the accept only
        t = F.copy(); t[k] = Fp[k]
        want = o.log_ld_F_total(t) - base_old
        assert abs(D[k] - want) <= RTOL * max(abs(base_old), 1.0), (k, D[k], want)
    z_new = s.get(_lib.STATE_Z)
    o.z[...] = z_new
    assert np.array_equal(s.get(_lib.STATE_TALLY) - before, o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    Fa = s.get(_lib.STATE_S)
    assert np.all((Fa == F) | (Fa == Fp))
    ok = np.isfinite(fk[:, 1, :]).all()
    lk = s.get(_lib.STATE_INDVLKH)
    want = np.array([o.log_ld_F(Fa, 1, i) for i in range(N)])
    if ok:
        assert np.max(np.abs(lk - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
        tot = s.get(_lib.STATE_TOTALLKH)[0]
        assert abs(tot - lk.sum()) <= 1e-9 * abs(tot) + len(lk) * 2.0 ** -24     # totallkh is a 2^-24 fixed-point sum
    s.close()


@pytest.mark.parametrize("mode", [4, 5])
def test_inbreeding_chain_runs_deterministically_and_graph_matches_direct(mode):
    d = make_dataset(N=150, L=30, K=3, A=4, miss=0.03, seed=8)
    sd = SeqData(d.x, d.allelenum, 3, mode=mode)
    out = []
    for ug in (0, 2, 0):
        s = Sampler(sd, update=60, burnin=20, thinning=4, ckrep=4, seed=11, use_graph=ug)
        ch, cv = s.run_chain(0, initd=[0.2, 0.5, 0.8])
        out.append((ch, cv))
        s.close()
    a, b, c = out
    for x, y in ((a, b), (a, c)):
        assert x[0].totallkh == y[0].totallkh and np.array_equal(x[0].qq, y[0].qq) and np.array_equal(x[1], y[1])
        assert np.array_equal(x[0].self_rates, y[0].self_rates)
    ns = 3 if mode == 4 else 150
    assert a[0].self_rates.shape == (ns,) and np.all((a[0].self_rates >= 0) & (a[0].self_rates <= 1))
    assert np.isfinite(a[0].totallkh) and a[0].step == a[0].steps == 10


def _z(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(axis=0, ddof=1) / len(a) + b.var(axis=0, ddof=1) / len(b))
    return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-12)


@pytest.mark.parametrize("mode", [4, 5])
def test_inbreeding_posterior_matches_reference_within_mcse(mode):
    g = np.load(os.path.join(GOLD, f"posterior_mode{mode}.npz"))
    K = int(g["K"])
    R = g["LL"].shape[0]
    sd = SeqData(g["x"], g["allelenum"], K, mode=mode)
    pop = g["pop"]
    LL, M, F = [], [], []
    for rep in range(R):
        s = Sampler(sd, update=int(g["update"]), burnin=int(g["burnin"]), thinning=int(g["thinning"]), ckrep=5, seed=5000 + rep)
        ch, _ = s.run_chain(rep, initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
        s.close()
        assert ch.flag_empty_cluster == 0 and ch.step == ch.steps
        # label switching: order the clusters by their inbreeding coefficient (mode 4) or by who holds population 0 (mode 5)
        o = np.argsort(ch.self_rates) if mode == 4 else np.argsort(ch.qq[pop == 0].mean(axis=0))[::-1]
        LL.append(ch.totallkh)
        M.append([ch.qq[pop == p][:, o[0]].mean() for p in range(K)])
        F.append(ch.self_rates[o] if mode == 4 else [ch.self_rates[pop == p].mean() for p in range(K)])
    LL, M, F = np.array(LL), np.array(M), np.array(F)
    refQ, refF, refLL = g["Q"].astype(np.float64), g["F"], g["LL"]
    # store_chn's running mean m*((step + x/m)/(1+step)) (mcmc.c:1327) turns into inf once a mean of q
    # underflows; such a reference chain (one of the ten mode-4 chains) is left out of the comparison
    good = np.isfinite(refQ).all(axis=(1, 2))
    assert good.sum() >= R - 2
    refQ, refF, refLL = refQ[good], refF[good], refLL[good]
    if mode == 4:
        ro = np.argsort(refF, axis=1)
        refQ = np.stack([refQ[r][:, ro[r]] for r in range(len(refQ))])
        refF = np.take_along_axis(refF, ro, axis=1)
    else:
        refF = np.stack([refF[:, pop == p].mean(axis=1) for p in range(K)], axis=1)
    refM = np.stack([refQ[:, pop == p, 0].mean(axis=1) for p in range(K)], axis=1)
    zLL, zQ, zF = _z(LL[:, None], refLL[:, None]), _z(M, refM), _z(F, refF)
    msg = f"zLL={zLL} zQ={zQ} zF={zF} LL={LL.mean()} ref={refLL.mean()} F={F.mean(0)} refF={refF.mean(0)} M={M.mean(0)} refM={refM.mean(0)}"
    assert np.all(np.abs(zLL) < 3.5), msg
    assert np.all(np.abs(zQ) < 3.5), msg
    assert np.all(np.abs(zF) < 3.5), msg
    assert np.all(np.abs(F.mean(0) - refF.mean(0)) < 0.03), msg
    assert abs(LL.mean() - refLL.mean()) < 10.0, msg
