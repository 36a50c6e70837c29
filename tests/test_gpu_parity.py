"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the
same seeded inputs (BASELINE.json north_star, three levels):

  1. integer work bit-exact      -- genotype store / missing mask, n[k][l][a] tallies and
                                    per-individual ancestry counts given identical Z;
  2. floating point on identical -- log-likelihood pieces, proposal(), alpha ratio:
     states                         |rel err| <= 1e-6 (device P is fp32, accumulation fp64);
  3. distributions               -- the categorical Z draw, Dirichlet P and Q draws, G
                                    proposal/accept: chi-square / moment tests against the
                                    oracle's exact conditionals (RNG streams cannot match).
"""
import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-6      # north_star: fp64 accumulation gate


def _mk(N, L, K, A, miss, seed, mode=2, **kw):
    d = make_dataset(N=N, L=L, K=K, A=A, miss=miss, seed=seed)
    sd = SeqData(d.x, d.allelenum, K, mode=mode, **kw)
    return d, sd


def _inject(s: Sampler, o: Oracle, rng, mode=2):
    """Random but valid state, identical on both sides (device P is fp32, so the oracle gets
    the fp32-rounded values)."""
    K = o.K
    o.z[...] = rng.integers(0, K, size=o.z.shape)
    o.qq[...] = rng.dirichlet(np.ones(K) * 0.8, size=o.N)
    f = rng.dirichlet(np.ones(o.Amax), size=(K, o.L))
    for l in range(o.L):                       # alleles beyond allelenum[l] do not exist
        a = o.allelenum[l]
        f[:, l, a:] = 0
        f[:, l, :a] /= f[:, l, :a].sum(axis=1, keepdims=True)
    o.freq[...] = f.astype(np.float32).astype(np.float64)
    o.gen[...] = rng.integers(1, 9, size=o.N)
    o.alpha = 0.9
    o.self_rates[...] = rng.uniform(0.1, 0.9, size=o.self_rates.shape)
    s.set(_lib.STATE_ITER, [1])
    s.set(_lib.STATE_Z, o.z)
    s.set(_lib.STATE_Q, o.qq)
    s.set(_lib.STATE_P, o.freq)
    s.set(_lib.STATE_G, o.gen)
    s.set(_lib.STATE_ALPHA, [o.alpha])
    s.set(_lib.STATE_S, o.self_rates)


SHAPES = [
    # N,   L,  K, A, miss
    (300, 40, 2, 2, 0.0),       # KP=4, SNP
    (257, 33, 5, 6, 0.05),      # KP=8, ragged N and L, microsatellite-like
    (64, 130, 8, 2, 0.1),       # KP=8, SNP, several micro-tiles
    (70, 21, 12, 3, 0.02),      # KP=16
    (530, 9, 3, 12, 0.3),       # many alleles, heavy missingness, several individual passes
    (70, 1300, 8, 2, 0.05),     # biallelic path (zq_snp.cu), 1024-locus chunks: one full chunk and a partial one
    (45, 600, 3, 2, 0.02),      # biallelic path, KP=4, one partial 1024-locus chunk
    (40, 300, 7, 2, 0.5),       # biallelic path, 256-locus chunks, half of the genotypes missing
]


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_store_and_mask_bit_exact(N, L, K, A, miss):
    d, sd = _mk(N, L, K, A, miss, seed=1)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K)
    x = s.get(_lib.STATE_X)
    usable = ~(d.x < 0).any(axis=2)
    assert np.array_equal(x[usable], d.x[usable])
    assert (x[~usable] < 0).all()              # a genotype with ANY missing copy is dropped whole
    assert np.array_equal(s.get(_lib.STATE_MASK), o.missing_mask())
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_tally_bit_exact_given_z(N, L, K, A, miss):
    d, sd = _mk(N, L, K, A, miss, seed=2)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(0)
    _inject(s, o, rng)
    assert np.array_equal(s.get(_lib.STATE_Z), o.z)
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    s.close()


@pytest.mark.parametrize("type_freq", [1, 0])
@pytest.mark.parametrize("N,L,K,A,miss", SHAPES)
def test_fused_sweep_pieces(N, L, K, A, miss, type_freq):
    """One fused pass on an injected state: the new Z it wrote, the tally and counts it
    accumulated from that Z (bit-exact), and the four log-likelihood pieces (1e-6)."""
    d, sd = _mk(N, L, K, A, miss, seed=3, type_freq=type_freq)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K, type_freq=type_freq)
    rng = np.random.default_rng(5)
    _inject(s, o, rng)
    z_old = o.z.copy()
    g_old = o.gen.copy()
    gprop = rng.integers(1, 12, size=o.N).astype(np.int32)
    s.set(_lib.STATE_GPROP, gprop)
    # the tally buffer is cleared by update_P in a real sweep; here it still holds tally(z_old)
    before = s.get(_lib.STATE_TALLY)
    s.run_phase(_lib.PHASE_ZQ)
    z_new = s.get(_lib.STATE_Z)
    usable = ~(d.x < 0).any(axis=2)
    assert np.array_equal(z_new[~usable], z_old[~usable])          # missing genotypes keep their z (mcmc.c:1137)
    assert z_new.min() >= 0 and z_new.max() < K
    # ---- integer work, bit-exact given the Z the kernel drew
    ll_old_g = np.array([o.log_ld_indv(g_old[i], i) for i in range(o.N)])
    ll_old_p = np.array([o.log_ld_indv(gprop[i], i) for i in range(o.N)])
    o.z[...] = z_new
    assert np.array_equal(s.get(_lib.STATE_TALLY) - before, o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    # ---- floating point, identical state
    parts = s.get(_lib.STATE_LLPARTS)
    ll_new_g = np.array([o.log_ld_indv(g_old[i], i) for i in range(o.N)])
    ll_new_p = np.array([o.log_ld_indv(gprop[i], i) for i in range(o.N)])
    scale = np.maximum(np.abs(ll_old_g), 1.0)
    assert np.max(np.abs(parts[:, 0] - (ll_old_p - ll_old_g)) / scale) <= RTOL
    assert np.max(np.abs(parts[:, 1] + parts[:, 2] - ll_new_g) / np.maximum(np.abs(ll_new_g), 1.0)) <= RTOL
    assert np.max(np.abs(parts[:, 1] + parts[:, 3] - ll_new_p) / np.maximum(np.abs(ll_new_p), 1.0)) <= RTOL
    # ---- accept bookkeeping: G is either the old or the proposed value and indvlkh matches it
    g_new = s.get(_lib.STATE_G)
    assert np.all((g_new == g_old) | (g_new == gprop))
    lk = s.get(_lib.STATE_INDVLKH)
    want = np.where(g_new == gprop, ll_new_p, ll_new_g)
    same = (gprop == g_old)
    want[same] = ll_new_g[same]
    assert np.max(np.abs(lk - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    s.close()


def test_standalone_loglik_proposal_alpha():
    N, L, K, A = 120, 64, 4, 5
    d, sd = _mk(N, L, K, A, 0.05, seed=4)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(7)
    _inject(s, o, rng)
    for gens in (o.gen.copy(), np.full(N, 1), np.full(N, 50)):
        got = s.loglik(gens)
        want = np.array([o.log_ld_indv(gens[i], i) for i in range(N)])
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    for _ in range(3):
        S = rng.uniform(0.02, 0.98, K)
        a, b = s.proposal_loglik(S), o.proposal(S)
        assert abs(a - b) <= RTOL * abs(b)
    for ra in (0.3, 1.7):
        a, b = s.alpha_logratio(ra), o.alpha_logratio(ra)
        assert abs(a - b) <= RTOL * max(abs(b), 1.0)
    s.close()


@pytest.mark.parametrize("K,A", [(5, 4), (6, 2), (3, 2)])
def test_z_draw_matches_exact_conditional(K, A):
    """Chi-square of the categorical draw (update_ZQ, mcmc.c:1139-1153) against the oracle's
    exact P(z = k) = Q_ik P_k,l,x / sum, pooled over repeated sweeps of a fixed state.  (K, A) = (5, 4): the generic
    kernel; A = 2: the biallelic path.  Both take 16 random bits per allele copy."""
    N, L = 32, 8
    d, sd = _mk(N, L, K, A, 0.0, seed=6)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(11)
    _inject(s, o, rng)
    reps = 600
    counts = np.zeros((L, N, 2, K))
    Q, P = o.qq.copy(), o.freq.copy()
    li, ni, ci = np.meshgrid(np.arange(L), np.arange(N), np.arange(2), indexing="ij")
    for r in range(reps):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_Q, Q)                 # the epilogue redraws Q after every pass
        s.run_phase(_lib.PHASE_ZQ)
        z = s.get(_lib.STATE_Z)
        counts[li, ni, ci, z] += 1
    chi2, dof = 0.0, 0
    for l in range(L):
        for i in range(N):
            for c in range(2):
                p = o.z_conditional(i, l, c)
                e = p * reps
                m = e > 5
                if m.sum() < 2:
                    continue
                obs = counts[l, i, c]
                # pool the low-expectation cells
                ee = np.append(e[m], e[~m].sum())
                oo = np.append(obs[m], obs[~m].sum())
                ok = ee > 0
                chi2 += (((oo - ee) ** 2)[ok] / ee[ok]).sum()
                dof += ok.sum() - 1
    # chi2 ~ N(dof, 2 dof) for large dof
    assert abs(chi2 - dof) < 5 * np.sqrt(2 * dof), (chi2, dof)
    s.close()


def test_p_and_q_draw_moments():
    """update_P: P[k][l][.] ~ Dir(n + 1) (mcmc.c:846-857); Q_i ~ Dir(cnt + alpha) (mcmc.c:1196-1198)."""
    N, L, K, A = 96, 16, 3, 4
    d, sd = _mk(N, L, K, A, 0.05, seed=8)
    s = Sampler(sd)
    o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(13)
    _inject(s, o, rng)
    n = o.tally().astype(np.float64)
    reps = 400
    accP, accP2 = np.zeros_like(n), np.zeros_like(n)
    for r in range(reps):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_Z, o.z)               # restores the tally that update_P consumes
        s.run_phase(_lib.PHASE_UPDATE_P)
        P = s.get(_lib.STATE_P)
        accP += P
        accP2 += P * P
    a = n + 1.0
    for l in range(L):
        a[:, l, o.allelenum[l]:] = 0
    a0 = a.sum(axis=2, keepdims=True)
    mean = a / a0
    var = a * (a0 - a) / (a0 * a0 * (a0 + 1))
    zscore = (accP / reps - mean) / np.sqrt(np.maximum(var, 1e-30) / reps)
    live = a > 0
    assert np.abs(zscore[live]).max() < 5.0
    assert abs(np.mean(zscore[live] ** 2) - 1.0) < 0.2
    # ---- Q: E[Q | cnt] = (cnt + alpha) / (sum cnt + K alpha) is linear in cnt, and sum cnt is
    # fixed per individual, so averaging over passes (whose Z draws and counts vary) is exact
    accQ = np.zeros((N, K))
    cnts = []
    for r in range(reps):
        s.set(_lib.STATE_ITER, [1000 + r])
        s.set(_lib.STATE_Q, o.qq)
        s.set(_lib.STATE_ALPHA, [o.alpha])
        s.run_phase(_lib.PHASE_ZQ)
        accQ += s.get(_lib.STATE_Q)
        cnts.append(s.get(_lib.STATE_CNT).astype(np.float64))
    # E[Q | cnt] averaged over the (varying) counts of each pass
    cm = np.mean(cnts, axis=0) + o.alpha
    want = cm / cm.sum(axis=1, keepdims=True)
    assert np.abs(accQ / reps - want).max() < 0.03
    s.close()


def test_same_seed_same_chain_and_launch_shape_independence():
    """Counter-based RNG: a chain is a pure function of (seed, chain id) -- identical when
    repeated -- and different for another chain id."""
    d, sd = _mk(200, 40, 3, 4, 0.05, seed=9)
    out = []
    for chain in (0, 0, 1):
        s = Sampler(sd, update=30, burnin=10, thinning=2, ckrep=5, seed=77)
        ch, cv = s.run_chain(chain, initd=[0.2, 0.5, 0.8])
        out.append((ch, cv))
        s.close()
    a, b, c = out
    assert a[0].totallkh == b[0].totallkh and np.array_equal(a[0].qq, b[0].qq) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0].self_rates, b[0].self_rates) and np.array_equal(a[0].gen, b[0].gen)
    assert a[0].totallkh != c[0].totallkh
    assert a[0].step == a[0].steps == 10


@pytest.mark.parametrize("mode,prior", [(2, 0), (3, 0), (1, 0)])
def test_graph_replay_is_bit_identical_to_direct_launches(mode, prior):
    """A sweep replayed as a CUDA graph (the default where it is eligible) reads its RNG counter
    from device memory; the chain must equal the directly launched one bit for bit, also when
    state hooks and single phases run between sweeps."""
    d, sd = _mk(210, 37, 3, 4, 0.05, seed=19, mode=mode, prior_flag=prior)
    res = []
    for ug in (0, 2):
        s = Sampler(sd, update=40, burnin=12, thinning=3, ckrep=4, seed=31, use_graph=ug)
        ch, cv = s.run_chain(0, initd=[0.2, 0.5, 0.8])
        s.chain_init(1, initd=[0.3, 0.4, 0.6])
        s.sweep(3)
        s.run_phase(_lib.PHASE_UPDATE_P)            # a hook between replays must not desynchronise the counter
        s.sweep(2)
        s.set(_lib.STATE_ITER, [11])
        s.sweep(2)
        res.append((ch, cv, s.get(_lib.STATE_Z), s.get(_lib.STATE_Q), s.get(_lib.STATE_INDVLKH), s.get(_lib.STATE_ITER)))
        s.close()
    a, b = res
    assert a[0].totallkh == b[0].totallkh and np.array_equal(a[0].qq, b[0].qq) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0].gen, b[0].gen) and np.array_equal(a[0].indvlkh, b[0].indvlkh)
    for u, v in zip(a[2:], b[2:]):
        assert np.array_equal(u, v)


def test_running_moments_match_direct_average():
    """store_chn (mcmc.c:1320): the device running means equal the plain average of the
    retained states read back sweep by sweep."""
    d, sd = _mk(90, 24, 3, 3, 0.0, seed=10)
    upd, burn, thin = 24, 8, 4
    s = Sampler(sd, update=upd, burnin=burn, thinning=thin, ckrep=4, seed=5)
    ch, cv = s.run_chain(0, initd=[0.3, 0.5, 0.7])
    s.close()
    s2 = Sampler(sd, update=upd, burnin=burn, thinning=thin, ckrep=4, seed=5)
    s2.chain_init(0, initd=[0.3, 0.5, 0.7])
    qs, ss, gs, ts, ls = [], [], [], [], []
    for step in range(upd):
        s2.sweep(1)
        if step >= burn and (step + 1 - burn) % thin == 0:
            qs.append(s2.get(_lib.STATE_Q)); ss.append(s2.get(_lib.STATE_S)); gs.append(s2.get(_lib.STATE_G))
            ts.append(s2.get(_lib.STATE_TOTALLKH)[0]); ls.append(s2.get(_lib.STATE_INDVLKH))
    s2.close()
    assert len(qs) == ch.steps == 4
    np.testing.assert_allclose(ch.qq, np.mean(qs, axis=0), rtol=1e-12)
    np.testing.assert_allclose(ch.qq2, np.mean(np.square(qs), axis=0), rtol=1e-12)
    np.testing.assert_allclose(ch.self_rates, np.mean(ss, axis=0), rtol=1e-12)
    np.testing.assert_allclose(ch.gen, np.mean(gs, axis=0), rtol=1e-12)
    np.testing.assert_allclose(ch.gen2, np.mean(np.square(np.array(gs, dtype=float)), axis=0), rtol=1e-12)
    np.testing.assert_allclose(ch.indvlkh, np.mean(ls, axis=0), rtol=1e-12)
    assert abs(ch.totallkh - np.mean(ts)) <= 1e-12 * abs(np.mean(ts))
    np.testing.assert_allclose(cv, ts[:4], rtol=0, atol=0)
    # totallkh is the sum of indvlkh (cal_lkh, mcmc.c:1940)
    assert abs(ts[-1] - ls[-1].sum()) <= 1e-9 * abs(ts[-1])


def test_sweep_keeps_tally_in_step_with_z():
    """After whole sweeps (update_P clears n, the fused pass refills it) the held tally is the
    tally of the held Z, and indvlkh is log_ld_indv of the held (Z, G, P)."""
    d, sd = _mk(300, 50, 4, 3, 0.05, seed=12)
    s = Sampler(sd, seed=3)
    s.chain_init(0, initd=[0.2, 0.4, 0.6, 0.8])
    s.sweep(5)
    o = Oracle(d.x, d.allelenum, 4)
    o.z[...] = s.get(_lib.STATE_Z)
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(float), o.count_z())
    o.freq[...] = s.get(_lib.STATE_P)
    g = s.get(_lib.STATE_G)
    want = np.array([o.log_ld_indv(g[i], i) for i in range(o.N)])
    got = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    q = s.get(_lib.STATE_Q)
    np.testing.assert_allclose(q.sum(axis=1), 1.0, rtol=1e-12)
    s.close()


@pytest.mark.parametrize("mode,prior", [(3, 0), (3, 1)])
def test_mode3_runs_and_is_consistent(mode, prior):
    d = make_dataset(N=150, L=40, K=3, A=4, miss=0.03, seed=14, s_atoms=[0.05, 0.5, 0.9])
    sd = SeqData(d.x, d.allelenum, 3, mode=mode, prior_flag=prior, alpha_dpm=2.0)
    s = Sampler(sd, update=40, burnin=10, thinning=3, ckrep=5, seed=21)
    ch, cv = s.run_chain(0)
    assert ch.self_rates.shape == (150,)
    assert np.all((ch.self_rates >= 0) & (ch.self_rates <= 1))
    assert np.all(ch.gen >= 1) and np.all(ch.gen <= 50)
    assert np.isfinite(ch.totallkh) and ch.step == 10
    s.close()


def test_errors_are_reported_not_fatal():
    d, sd = _mk(40, 10, 2, 2, 0.0, seed=15)
    import instruct_b200
    bad = SeqData(d.x, d.allelenum, 40)
    with pytest.raises(instruct_b200.InstructError):
        Sampler(bad)
    s = Sampler(sd)
    with pytest.raises(instruct_b200.InstructError):
        s.sweep(1)                             # chain not initialised
    s.close()


def test_mode1_sweep_is_the_no_selfing_model():
    """Mode 1 (mcmc_POP_admixture, mcmc.c:135-180): update_P, update_ZQ, update_alpha, cal_lkh with
    log_ld_noselfing_indv (mcmc.c:1869).  After a few sweeps the stored likelihood is that of the
    state, the tally and counts are in step with Z, and G stays 1."""
    d, sd = _mk(257, 33, 5, 6, 0.05, seed=21, mode=1)
    s = Sampler(sd, seed=3)
    s.chain_init(0)
    s.sweep(4)
    o = Oracle(d.x, d.allelenum, 5, mode=1)
    o.z[...] = s.get(_lib.STATE_Z)
    o.freq[...] = s.get(_lib.STATE_P)
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    assert (s.get(_lib.STATE_G) == 1).all()
    o.cal_lkh()
    lk = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(lk - o.indvlkh) / np.maximum(np.abs(o.indvlkh), 1.0)) <= RTOL
    assert abs(float(s.get(_lib.STATE_TOTALLKH)[0]) - o.totallkh) <= RTOL * abs(o.totallkh)
    s.close()
