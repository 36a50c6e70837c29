"""GPU parity tests of the ALLOTETRAPLOID sweep (-p 4 -ap 0), through the
C-ABI, against oracle/tetra_oracle.c with autopoly = 0 (itself pinned bit for bit to the compiled
reference, tests/test_tetra_oracle_vs_reference.py):

  integer work, bit-exact   the two tallies of update_P_allo (copies 0,1 -> n, copies 2,3 -> n2,
                            poly_geno.c:441-489) and the ancestry counts given identical (z, geno);
  floating point            calc_exfreq_allo / allo_genfreq tables (float in the reference: 1e-5),
                            the S statistics, indvlkh / totallkh (1e-6);
  distributions             the 7 / 12 / 6-way dosage resolutions (choose_two_allo :962,
                            choose_tri_allo :1043, choose_tetra_allo :1144) against the oracle's exact
                            conditionals, chi-square.
"""
import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_tetra_dataset
from oracle.pytetra import TetraOracle

pytestmark = pytest.mark.gpu

RES2 = [(0, 0, 0, 1), (0, 1, 0, 0), (0, 0, 1, 1), (1, 1, 0, 0), (0, 1, 1, 1), (1, 1, 0, 1), (0, 1, 0, 1)]
RES3 = [(0, 0, 1, 2), (1, 2, 0, 0), (1, 1, 0, 2), (0, 2, 1, 1), (2, 2, 0, 1), (0, 1, 2, 2),
        (0, 1, 1, 2), (1, 2, 0, 1), (1, 2, 0, 2), (0, 2, 1, 2), (0, 2, 0, 1), (0, 1, 0, 2)]
RES4 = [(0, 1, 2, 3), (2, 3, 0, 1), (0, 2, 1, 3), (1, 3, 0, 2), (0, 3, 1, 2), (1, 2, 0, 3)]
RES = {2: RES2, 3: RES3, 4: RES4}

SHAPES = [
    # N,  L,  K, A, miss
    (70, 13, 3, 4, 0.05),     # KP=4, ragged L
    (300, 9, 6, 4, 0.02),     # config-5 shape in small: K=6 (KP=8), several individual passes
    (40, 22, 2, 2, 0.1),      # biallelic: only the 7-way resolution
    (33, 8, 5, 5, 0.0),       # largest catalogue an 8-bit index held (225 genotypes)
    (40, 6, 3, 7, 0.02),      # 784 genotypes: 16-bit catalogue indices, two-dp4a codes
    (30, 4, 2, 10, 0.0),      # the limit: 3025 genotypes
    (50, 12, 2, 3, 0.3),      # heavy missingness
]


def _mk(N, L, K, A, miss, seed):
    d = make_tetra_dataset(N=N, L=L, K=K, A=A, miss=miss, seed=seed)
    return d, SeqData(d.x, d.allelenum, K, ploid=4, autopoly=0)


def _freq(rng, o):
    f = rng.dirichlet(np.ones(o.Amax), size=(o.K, o.L))
    for l in range(o.L):
        a = o.allelenum[l]
        f[:, l, a:] = 0
        f[:, l, :a] /= f[:, l, :a].sum(axis=1, keepdims=True)
    return f.astype(np.float32).astype(np.float64)


def _inject(s: Sampler, o: TetraOracle, rng):
    K = o.K
    o.initial_geno()
    o.geno[o.nd == 0] = -1
    o.z[...] = rng.integers(0, K, size=o.z.shape)
    same = rng.random(o.z.shape[:2]) < 0.4
    o.z[same] = o.z[same][:, :1]
    o.qq[...] = rng.dirichlet(np.ones(K) * 0.8, size=o.N)
    o.freq[...] = _freq(rng, o)
    o.freq2[...] = _freq(rng, o)
    o.alpha = 0.9
    o.self_rates[...] = rng.uniform(0.1, 0.9, size=K)
    sprop = np.clip(o.self_rates + rng.uniform(-0.05, 0.05, size=K), 0.01, 0.99)
    s.set(_lib.STATE_ITER, [1])
    s.set(_lib.STATE_GENO, o.geno)
    s.set(_lib.STATE_Z, o.z)
    s.set(_lib.STATE_Q, o.qq)
    s.set(_lib.STATE_P, o.freq)
    s.set(_lib.STATE_P2, o.freq2)
    s.set(_lib.STATE_ALPHA, [o.alpha])
    s.set(_lib.STATE_S, o.self_rates)
    s.set(_lib.STATE_SPROP, sprop)
    s.refresh_tables()
    o.tables()
    return sprop


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_allo_tables_match_oracle(N, L, K, A, miss):
    d, sd = _mk(N, L, K, A, miss, seed=2)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    sprop = _inject(s, o, np.random.default_rng(3))
    assert s.gmax() == o.Gmax
    assert np.array_equal(s.get(_lib.STATE_P2), o.freq2)
    ex, cur, prop = s.get(_lib.STATE_EXFREQ), s.get(_lib.STATE_TABLES), s.get(_lib.STATE_TABLES_PROP)
    want_prop = np.stack([o.calc_genofreq(k, sprop[k]) for k in range(K)])
    for l in range(d.L):
        n = len(o.genolist(l))
        np.testing.assert_allclose(ex[:, l, :n], o.exfreq[:, l, :n], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cur[:, l, :n], o.genofreq[:, l, :n], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(prop[:, l, :n], want_prop[:, l, :n], rtol=1e-5, atol=2e-6)
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_allo_pass_a_statistics_counts_and_z(N, L, K, A, miss):
    d, sd = _mk(N, L, K, A, miss, seed=4)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    sprop = _inject(s, o, np.random.default_rng(5))
    z_old = o.z.copy()
    base = o.cal_lkd()
    want_D = np.array([o.cal_lkd_props(k, o.calc_genofreq(k, sprop[k])) - base for k in range(K)])
    s.run_phase(_lib.PHASE_ZQ)
    D = s.get(_lib.STATE_DSTAT)
    n_same = int(((z_old == z_old[:, :, :1]).all(axis=2) & (d.nd > 0)).sum())
    assert np.max(np.abs(D - want_D)) <= 2e-6 * max(n_same, 1) + 1e-9 * abs(base)
    z_new = s.get(_lib.STATE_Z)
    usable = d.nd > 0
    assert np.array_equal(z_new[~usable], z_old[~usable])
    assert z_new[usable].min() >= 0 and z_new[usable].max() < K
    o.z[...] = z_new
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_allo_pass_b_geno_likelihood_and_tallies(N, L, K, A, miss):
    """PASS B on an injected state: a legal resolution was written (each pair ascending, a catalogue
    genotype), the likelihood is that of the written state, both tallies are those of (z, new geno)."""
    d, sd = _mk(N, L, K, A, miss, seed=6)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    _inject(s, o, np.random.default_rng(7))
    b1, b2 = s.get(_lib.STATE_TALLY), s.get(_lib.STATE_TALLY2)
    w1, w2 = o.tally_allo()
    assert np.array_equal(b1, w1) and np.array_equal(b2, w2)       # set(Z/GENO) keeps the tallies in step
    s.run_phase(_lib.PHASE_GENO)
    g_new = s.get(_lib.STATE_GENO)
    usable = d.nd > 0
    assert (g_new[~usable] == -1).all()
    for l in range(d.L):
        for i in range(d.N):
            nd = int(d.nd[l, i])
            if nd == 0:
                continue
            a = [int(v) for v in d.x[l, i, :nd]]
            gg = tuple(int(v) for v in g_new[l, i])
            if nd == 1:
                assert gg == (a[0],) * 4
            else:
                assert gg in [tuple(a[q] for q in r) for r in RES[nd]], (l, i, gg, a)
            assert o.geno_index(l, g_new[l, i]) >= 0
    o.geno[...] = g_new
    tot = o.cal_lkd()
    lk = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(lk - o.indvlkh) / np.maximum(np.abs(o.indvlkh), 1.0)) <= 1e-6
    assert abs(float(s.get(_lib.STATE_TOTALLKH)[0]) - tot) <= 1e-6 * abs(tot)
    w1, w2 = o.tally_allo()
    assert np.array_equal(s.get(_lib.STATE_TALLY) - b1, w1)
    assert np.array_equal(s.get(_lib.STATE_TALLY2) - b2, w2)
    s.close()


def test_allo_resolution_matches_exact_conditionals():
    """Chi-square of the 7 / 12 / 6-way dosage resolutions against the oracle's exact conditionals."""
    N, L, K, A = 24, 8, 3, 4
    d, sd = _mk(N, L, K, A, 0.0, seed=8)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    _inject(s, o, np.random.default_rng(9))
    geno0, z0, Q = o.geno.copy(), o.z.copy(), o.qq.copy()
    reps = 600
    gc = np.zeros((L, N, 12))
    for r in range(reps):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_GENO, geno0)
        s.set(_lib.STATE_Z, z0)
        s.set(_lib.STATE_Q, Q)
        s.run_phase(_lib.PHASE_GENO)
        g = s.get(_lib.STATE_GENO)
        for l in range(L):
            for i in range(N):
                nd = int(d.nd[l, i])
                if nd >= 2:
                    a = d.x[l, i]
                    gg = tuple(int(v) for v in g[l, i])
                    gc[l, i, [tuple(int(a[q]) for q in rr) for rr in RES[nd]].index(gg)] += 1
    o.geno[...] = geno0
    o.z[...] = z0
    o.qq[...] = Q
    c2 = dof = 0
    seen = set()
    for l in range(L):
        for i in range(N):
            nd = int(d.nd[l, i])
            if nd < 2:
                continue
            p = o.geno_conditional_allo(i, l)
            assert len(p) == len(RES[nd])
            e = p * reps
            m = e > 5
            if m.sum() < 2:
                continue
            ee = np.append(e[m], e[~m].sum())
            oo = np.append(gc[l, i, : len(p)][m], gc[l, i, : len(p)][~m].sum())
            ok = ee > 0
            c2 += (((oo - ee) ** 2)[ok] / ee[ok]).sum(); dof += ok.sum() - 1
            seen.add(nd)
    assert seen == {2, 3, 4}
    assert dof > 100 and abs(c2 - dof) < 5 * np.sqrt(2 * dof), ("geno", c2, dof)
    s.close()


def test_allo_update_p_draws_both_subgenomes():
    """update_P_allo: freq ~ Dirichlet(n + 1) and freq2 ~ Dirichlet(n2 + 1), two different streams."""
    N, L, K, A = 200, 6, 2, 3
    d, sd = _mk(N, L, K, A, 0.0, seed=10)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    _inject(s, o, np.random.default_rng(11))
    n1, n2 = o.tally_allo()
    reps = 300
    m1, m2 = np.zeros((K, L, A)), np.zeros((K, L, A))
    for r in range(reps):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_Z, o.z)                     # restores both tallies (update_P consumes them)
        s.run_phase(_lib.PHASE_UPDATE_P)
        p1, p2 = s.get(_lib.STATE_P), s.get(_lib.STATE_P2)
        assert not np.array_equal(p1, p2)
        np.testing.assert_allclose(p1.sum(axis=2), 1.0, rtol=1e-5)
        np.testing.assert_allclose(p2.sum(axis=2), 1.0, rtol=1e-5)
        m1 += p1; m2 += p2
    for m, n in ((m1 / reps, n1), (m2 / reps, n2)):
        a = n + 1.0
        for l in range(L):
            a[:, l, o.allelenum[l]:] = 0
        a0 = a.sum(axis=2, keepdims=True)
        mean, var = a / a0, a * (a0 - a) / (a0 * a0 * (a0 + 1))
        z = (m - mean) / np.sqrt(np.maximum(var, 1e-12) / reps)
        assert np.abs(z[var > 0]).max() < 5.0
    s.close()


def test_allo_chain_runs_and_state_is_consistent():
    N, L, K, A = 120, 40, 3, 4
    d, sd = _mk(N, L, K, A, 0.03, seed=12)
    sd.nstep_check_empty_cluster = 1 << 30
    s = Sampler(sd, update=60, burnin=20, thinning=2, ckrep=5, seed=5)
    ch, cv = s.run_chain(0, initd=[0.3, 0.5, 0.7])
    assert ch.step == 20 and np.isfinite(ch.totallkh) and ch.totallkh < 0
    np.testing.assert_allclose(ch.qq.sum(axis=1), 1.0, rtol=1e-9)
    assert ((ch.self_rates > 0) & (ch.self_rates < 1)).all()
    assert np.isfinite(cv).all()
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    o.z[...] = s.get(_lib.STATE_Z)
    o.geno[...] = s.get(_lib.STATE_GENO)
    w1, w2 = o.tally_allo()
    assert np.array_equal(s.get(_lib.STATE_TALLY), w1) and np.array_equal(s.get(_lib.STATE_TALLY2), w2)
    o.qq[...] = s.get(_lib.STATE_Q)
    o.freq[...] = s.get(_lib.STATE_P)
    o.freq2[...] = s.get(_lib.STATE_P2)
    o.self_rates[...] = s.get(_lib.STATE_S)
    o.tables()
    tot = o.cal_lkd()
    assert abs(float(s.get(_lib.STATE_TOTALLKH)[0]) - tot) <= 2e-6 * abs(tot)
    s2 = Sampler(sd, update=60, burnin=20, thinning=2, ckrep=5, seed=5)
    ch2, cv2 = s2.run_chain(0, initd=[0.3, 0.5, 0.7])
    assert np.array_equal(cv, cv2) and np.array_equal(ch.qq, ch2.qq)
    s.close(); s2.close()
