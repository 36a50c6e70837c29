"""GPU parity tests of the AUTOTETRAPLOID sweep, through the C-ABI,
against oracle/tetra_oracle.c (itself pinned bit for bit to the compiled reference):

  integer work, bit-exact   genotype store / mask, the tally over the latent genotype and the
                            per-individual ancestry counts given identical (z, geno);
  floating point            log genotype-frequency tables (float in the reference: 1e-5 relative,
                            the fp32 gate of BASELINE.json is 1e-4), the S statistics
                            cal_lkd_props(k) - cal_lkd(), indvlkh / totallkh (1e-6);
  distributions             the z draw and the dosage resolution against the oracle's exact
                            conditionals (chi-square; RNG streams cannot match).
"""
import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_tetra_dataset
from oracle.pytetra import TetraOracle

pytestmark = pytest.mark.gpu

SHAPES = [
    # N,  L,  K, A, miss
    (70, 13, 3, 4, 0.05),     # KP=4, ragged L
    (300, 9, 6, 4, 0.02),     # config-5 shape in small: K=6 (KP=8), several individual passes
    (40, 22, 2, 2, 0.1),      # biallelic: no tri / quadri classes
    (33, 8, 5, 6, 0.0),       # largest catalogue whose code is one dp4a (126 genotypes)
    (40, 6, 3, 8, 0.02),      # 330 genotypes: 16-bit catalogue indices, two-dp4a codes
    (36, 5, 2, 10, 0.0),      # microsatellite-like, the limit: 715 genotypes
    (50, 12, 2, 3, 0.3),      # heavy missingness
]


def _mk(N, L, K, A, miss, seed):
    d = make_tetra_dataset(N=N, L=L, K=K, A=A, miss=miss, seed=seed)
    sd = SeqData(d.x, d.allelenum, K, ploid=4, autopoly=1)
    return d, sd


def _inject(s: Sampler, o: TetraOracle, rng):
    """Random valid state on both sides (device P is fp32, so the oracle gets the rounded values)."""
    K = o.K
    o.initial_geno()
    o.geno[o.nd == 0] = -1                           # the library's convention for missing genotypes
    o.z[...] = rng.integers(0, K, size=o.z.shape)
    # make a good share of genotypes single-population, the case the selfing tables serve
    same = rng.random(o.z.shape[:2]) < 0.4
    o.z[same] = o.z[same][:, :1]
    o.qq[...] = rng.dirichlet(np.ones(K) * 0.8, size=o.N)
    f = rng.dirichlet(np.ones(o.Amax), size=(K, o.L))
    for l in range(o.L):
        a = o.allelenum[l]
        f[:, l, a:] = 0
        f[:, l, :a] /= f[:, l, :a].sum(axis=1, keepdims=True)
    o.freq[...] = f.astype(np.float32).astype(np.float64)
    o.alpha = 0.9
    o.self_rates[...] = rng.uniform(0.1, 0.9, size=K)
    sprop = np.clip(o.self_rates + rng.uniform(-0.05, 0.05, size=K), 0.01, 0.99)
    s.set(_lib.STATE_ITER, [1])
    s.set(_lib.STATE_GENO, o.geno)
    s.set(_lib.STATE_Z, o.z)
    s.set(_lib.STATE_Q, o.qq)
    s.set(_lib.STATE_P, o.freq)
    s.set(_lib.STATE_ALPHA, [o.alpha])
    s.set(_lib.STATE_S, o.self_rates)
    s.set(_lib.STATE_SPROP, sprop)
    s.refresh_tables()
    o.tables()
    return sprop


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_store_and_mask_bit_exact(N, L, K, A, miss):
    d, sd = _mk(N, L, K, A, miss, seed=1)
    s = Sampler(sd)
    assert np.array_equal(s.get(_lib.STATE_X), d.x)
    assert np.array_equal(s.get(_lib.STATE_MASK) != 0, d.nd == 0)
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_tables_match_oracle(N, L, K, A, miss):
    d, sd = _mk(N, L, K, A, miss, seed=2)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K)
    sprop = _inject(s, o, np.random.default_rng(3))
    assert s.gmax() == o.Gmax
    ex, cur, prop = s.get(_lib.STATE_EXFREQ), s.get(_lib.STATE_TABLES), s.get(_lib.STATE_TABLES_PROP)
    want_prop = np.stack([o.calc_genofreq(k, sprop[k]) for k in range(K)])
    for l in range(d.L):
        n = len(o.genolist(l))
        np.testing.assert_allclose(ex[:, l, :n], o.exfreq[:, l, :n], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cur[:, l, :n], o.genofreq[:, l, :n], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(prop[:, l, :n], want_prop[:, l, :n], rtol=1e-5, atol=1e-6)
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_pass_a_statistics_counts_and_z(N, L, K, A, miss):
    """PASS A on an injected state: the S statistics on the OLD z, the new z it wrote and the
    ancestry counts accumulated from that z."""
    d, sd = _mk(N, L, K, A, miss, seed=4)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K)
    sprop = _inject(s, o, np.random.default_rng(5))
    z_old = o.z.copy()
    base = o.cal_lkd()
    want_D = np.array([o.cal_lkd_props(k, o.calc_genofreq(k, sprop[k])) - base for k in range(K)])
    s.run_phase(_lib.PHASE_ZQ)                       # no PHASE_UPDATE_S: statistics only, S untouched
    D = s.get(_lib.STATE_DSTAT)
    n_same = int(((z_old == z_old[:, :, :1]).all(axis=2) & (d.nd > 0)).sum())
    assert np.max(np.abs(D - want_D)) <= 2e-6 * max(n_same, 1) + 1e-9 * abs(base)
    assert np.array_equal(s.get(_lib.STATE_S), o.self_rates)
    z_new = s.get(_lib.STATE_Z)
    usable = d.nd > 0
    assert np.array_equal(z_new[~usable], z_old[~usable])
    assert z_new[usable].min() >= 0 and z_new[usable].max() < K
    o.z[...] = z_new
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    q = s.get(_lib.STATE_Q)
    np.testing.assert_allclose(q.sum(axis=1), 1.0, rtol=1e-12)
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_pass_b_geno_likelihood_and_tally(N, L, K, A, miss):
    """PASS B on an injected state: the dosage resolution it wrote is one of the legal ones, the
    likelihood is that of the written state (1e-6), and the tally is that of (z, new geno)."""
    d, sd = _mk(N, L, K, A, miss, seed=6)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K)
    _inject(s, o, np.random.default_rng(7))
    before = s.get(_lib.STATE_TALLY)                 # set(Z/GENO) keeps tally(z, injected geno) in the buffer
    assert np.array_equal(before, o.tally())
    s.run_phase(_lib.PHASE_GENO)
    g_new = s.get(_lib.STATE_GENO)
    usable = d.nd > 0
    assert (g_new[~usable] == -1).all()
    for l in range(d.L):
        for i in range(d.N):
            nd = int(d.nd[l, i])
            if nd == 0:
                continue
            obs = sorted(int(v) for v in d.x[l, i, :nd])
            assert sorted(set(int(v) for v in g_new[l, i])) == obs
            o.geno_index(l, g_new[l, i])              # aborts if not a catalogue genotype in canonical writing
    o.geno[...] = g_new
    tot = o.cal_lkd()
    lk = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(lk - o.indvlkh) / np.maximum(np.abs(o.indvlkh), 1.0)) <= 1e-6
    assert abs(float(s.get(_lib.STATE_TOTALLKH)[0]) - tot) <= 1e-6 * abs(tot)
    assert np.array_equal(s.get(_lib.STATE_TALLY) - before, o.tally())
    s.close()


def test_z_draw_and_resolution_match_exact_conditionals():
    """Chi-square of the per-copy z draw (poly_geno.c:766-779) and of the three-way dosage
    resolution (choose_two_auto / choose_tri_auto) against the oracle's exact conditionals."""
    N, L, K, A = 24, 8, 3, 4
    d, sd = _mk(N, L, K, A, 0.0, seed=8)
    s = Sampler(sd)
    o = TetraOracle(d.x, d.nd, d.allelenum, K)
    _inject(s, o, np.random.default_rng(9))
    geno0, z0, Q = o.geno.copy(), o.z.copy(), o.qq.copy()
    reps = 500
    zc = np.zeros((L, N, 4, K))
    gc = np.zeros((L, N, 3))
    li, ni, ci = np.meshgrid(np.arange(L), np.arange(N), np.arange(4), indexing="ij")
    for r in range(reps):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_GENO, geno0)
        s.set(_lib.STATE_Z, z0)
        s.set(_lib.STATE_Q, Q)
        s.run_phase(_lib.PHASE_GENO)                 # resolution | z0, Q
        g = s.get(_lib.STATE_GENO)
        for l in range(L):
            for i in range(N):
                nd = int(d.nd[l, i])
                if nd in (2, 3):
                    a = d.x[l, i]
                    gg = tuple(int(v) for v in g[l, i])
                    if nd == 2:
                        opts = [(a[0], a[0], a[0], a[1]), (a[1], a[1], a[1], a[0]), (a[0], a[0], a[1], a[1])]
                    else:
                        opts = [(a[0], a[0], a[1], a[2]), (a[1], a[1], a[0], a[2]), (a[2], a[2], a[0], a[1])]
                    gc[l, i, [tuple(int(v) for v in op) for op in opts].index(gg)] += 1
        s.set(_lib.STATE_GENO, geno0)
        s.set(_lib.STATE_Q, Q)
        s.run_phase(_lib.PHASE_ZQ)                   # z | geno0, Q
        z = s.get(_lib.STATE_Z)
        zc[li, ni, ci, z] += 1
    o.geno[...] = geno0
    o.z[...] = z0
    o.qq[...] = Q

    def chi(obs, p):
        e = p * reps
        m = e > 5
        if m.sum() < 2:
            return 0.0, 0
        ee = np.append(e[m], e[~m].sum())
        oo = np.append(obs[m], obs[~m].sum())
        ok = ee > 0
        return (((oo - ee) ** 2)[ok] / ee[ok]).sum(), ok.sum() - 1
    c2 = dof = 0
    for l in range(L):
        for i in range(N):
            for c in range(4):
                a, b = chi(zc[l, i, c], o.z_conditional(i, l, c))
                c2 += a; dof += b
    assert abs(c2 - dof) < 5 * np.sqrt(2 * dof), ("z", c2, dof)
    c2 = dof = 0
    for l in range(L):
        for i in range(N):
            if d.nd[l, i] in (2, 3):
                a, b = chi(gc[l, i], o.geno_conditional(i, l))
                c2 += a; dof += b
    assert dof > 20 and abs(c2 - dof) < 5 * np.sqrt(2 * dof), ("geno", c2, dof)
    s.close()


def test_chain_runs_and_moments_are_consistent():
    """A whole chain through ig_run_chain: finite moments, Q rows sum to one, rates inside (0,1),
    the tally stays in step with (z, geno), and the stored likelihood is the state's."""
    N, L, K, A = 120, 40, 3, 4
    d, sd = _mk(N, L, K, A, 0.03, seed=12)
    sd.nstep_check_empty_cluster = 1 << 30
    s = Sampler(sd, update=60, burnin=20, thinning=2, ckrep=5, seed=5)
    ch, cv = s.run_chain(0, initd=[0.3, 0.5, 0.7])
    assert ch.step == 20 and np.isfinite(ch.totallkh) and ch.totallkh < 0
    np.testing.assert_allclose(ch.qq.sum(axis=1), 1.0, rtol=1e-9)
    assert ((ch.self_rates > 0) & (ch.self_rates < 1)).all()
    assert np.isfinite(cv).all()
    o = TetraOracle(d.x, d.nd, d.allelenum, K)
    o.z[...] = s.get(_lib.STATE_Z)
    o.geno[...] = s.get(_lib.STATE_GENO)
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    o.qq[...] = s.get(_lib.STATE_Q)
    o.freq[...] = s.get(_lib.STATE_P)
    o.self_rates[...] = s.get(_lib.STATE_S)
    o.tables()
    tot = o.cal_lkd()
    assert abs(float(s.get(_lib.STATE_TOTALLKH)[0]) - tot) <= 2e-6 * abs(tot)
    # same seed, same chain
    s2 = Sampler(sd, update=60, burnin=20, thinning=2, ckrep=5, seed=5)
    ch2, cv2 = s2.run_chain(0, initd=[0.3, 0.5, 0.7])
    assert np.array_equal(cv, cv2) and np.array_equal(ch.qq, ch2.qq)
    s.close(); s2.close()
