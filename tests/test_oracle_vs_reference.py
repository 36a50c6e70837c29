"""Pins the CPU restatement (oracle/instruct_oracle.c) to the UNMODIFIED reference compiled
into oracle/_ref (oracle/Makefile) -- SURVEY.md section 8c.  Integer work bit-exact, whole
chains bit-exact on identical Wichmann-Hill seeds, floating-point evaluators to 1e-12.
The reference has no golden vectors of its own; tests/test_golden.py additionally checks
the oracle against fixtures that tools/make_golden.py generated from the same reference."""
import numpy as np
import pytest

from instruct_b200.synth import make_dataset
from oracle import pyoracle
from oracle.pyoracle import Oracle, Reference

pytestmark = pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _state(o, rng, K):
    o.z[...] = rng.integers(0, K, size=o.z.shape)
    o.qq[...] = rng.dirichlet(np.ones(K) * 0.7, size=o.N)
    f = rng.dirichlet(np.ones(o.Amax), size=(K, o.L))
    o.freq[...] = f
    o.gen[...] = rng.integers(1, 8, size=o.N)
    o.alpha = 0.8


def _push(o, r):
    r.set_z(o.z); r.set_qq(o.qq); r.set_freq(o.freq); r.set_gen(o.gen); r.set_alpha(o.alpha)
    r.set_self(o.self_rates)


@pytest.mark.parametrize("K,A,miss", [(2, 3, 0.0), (5, 6, 0.05), (8, 2, 0.1)])
def test_tally_mask_and_counts_bit_exact(K, A, miss):
    d = make_dataset(N=70, L=33, K=K, A=A, miss=miss, seed=3)
    o, r = Oracle(d.x, d.allelenum, K), Reference(d.x, d.allelenum, K)
    rng = np.random.default_rng(0)
    _state(o, rng, K)
    o.self_rates[...] = rng.uniform(0.1, 0.9, K)
    _push(o, r)
    assert np.array_equal(o.missing_mask(), r.missindx())
    ref_tally = r.update_P(want_tally=True)           # the reference's own loop nest, mcmc.c:815-845
    assert np.array_equal(o.tally(), ref_tally)
    # per-individual counts: the reference fills qqnum inside update_ZQ from the z it just drew
    r.setseeds(5, 6, 7)
    r.update_ZQ(0)
    o.z[...] = r.get_z()
    assert np.array_equal(o.count_z(), r.get_qqnum())


@pytest.mark.parametrize("type_freq", [1, 0])
def test_loglik_and_proposal(type_freq):
    K = 4
    d = make_dataset(N=40, L=50, K=K, A=5, miss=0.05, seed=9)
    o = Oracle(d.x, d.allelenum, K, type_freq=type_freq)
    r = Reference(d.x, d.allelenum, K, type_freq=type_freq)
    rng = np.random.default_rng(1)
    _state(o, rng, K)
    o.self_rates[...] = rng.uniform(0.05, 0.95, K)
    _push(o, r)
    for i in range(0, 40, 3):
        for g in (1, 2, 7, 50):
            a, b = o.log_ld_indv(g, i), r.log_ld_indv(g, i)
            assert a == b or abs(a - b) <= 1e-12 * abs(b)
    S2 = rng.uniform(0.05, 0.95, K)
    assert o.proposal(S2) == r.proposal(S2)
    r.cal_lkh(); o.cal_lkh()
    ind, tot = r.get_lkh()
    assert np.array_equal(ind, o.indvlkh) and tot == o.totallkh
    # log form of the alpha ratio agrees with the reference's product form where that is finite
    prod = o.alpha_ratio_product(1.3)
    if np.isfinite(prod) and prod > 0:
        assert abs(np.log(prod) - o.alpha_logratio(1.3)) <= 1e-9 * abs(o.alpha_logratio(1.3))


@pytest.mark.parametrize("mode,prior,back_refl", [(2, 0, 1), (2, 0, 0), (3, 0, 1), (3, 1, 1)])
def test_whole_chain_bit_exact(mode, prior, back_refl):
    K = 3
    d = make_dataset(N=45, L=21, K=K, A=4, miss=0.04, seed=11)
    o = Oracle(d.x, d.allelenum, K, mode=mode, prior_flag=prior, back_refl=back_refl)
    r = Reference(d.x, d.allelenum, K, mode=mode, prior_flag=prior, back_refl=back_refl)
    o.setseeds(13, 4, 1972); r.setseeds(13, 4, 1972)
    kw = dict(update=160, burnin=60, thinning=5, ckrep=8, nstep_check_empty=10, initd=[0.3, 0.5, 0.7])
    co = o.run_chain(**kw)
    cr = r.mcmc_updating(**kw)
    assert co["flag_empty_cluster"] == cr["flag_empty_cluster"] == 0
    for k in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "gen", "gen2", "convg"]:
        assert np.array_equal(np.asarray(co[k]), np.asarray(cr[k])), k


def test_whole_chain_bit_exact_mode1():
    """Mode 1, the CLI default (mcmc_POP_admixture, mcmc.c:135-180): admixture without selfing."""
    K = 3
    d = make_dataset(N=45, L=21, K=K, A=4, miss=0.04, seed=12)
    o = Oracle(d.x, d.allelenum, K, mode=1)
    r = Reference(d.x, d.allelenum, K, mode=1)
    o.setseeds(13, 4, 1972); r.setseeds(13, 4, 1972)
    kw = dict(update=160, burnin=60, thinning=5, ckrep=8, nstep_check_empty=10)
    co = o.run_chain(**kw)
    cr = r.mcmc_updating(**kw)
    assert co["flag_empty_cluster"] == cr["flag_empty_cluster"] == 0
    for k in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "convg"]:
        assert np.array_equal(np.asarray(co[k]), np.asarray(cr[k])), k


@pytest.mark.parametrize("mode,back_refl", [(4, 1), (4, 0), (5, 1)])
def test_whole_chain_bit_exact_inbreeding_modes(mode, back_refl):
    """Modes 4/5 (mcmc_POP_inbreedcoff mcmc.c:242, mcmc_INDV_inbreedcoff :386 with the uniform
    prior): inbreeding coefficients per population / per individual instead of selfing rates."""
    K = 3
    d = make_dataset(N=40, L=19, K=K, A=4, miss=0.04, seed=13)
    o = Oracle(d.x, d.allelenum, K, mode=mode, back_refl=back_refl)
    r = Reference(d.x, d.allelenum, K, mode=mode, back_refl=back_refl)
    o.setseeds(13, 4, 1972); r.setseeds(13, 4, 1972)
    kw = dict(update=120, burnin=40, thinning=5, ckrep=8, nstep_check_empty=10, initd=[0.3, 0.5, 0.7])
    co = o.run_chain(**kw)
    cr = r.mcmc_updating(**kw)
    assert co["flag_empty_cluster"] == cr["flag_empty_cluster"] == 0
    for k in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "convg"]:
        assert np.array_equal(np.asarray(co[k]), np.asarray(cr[k])), k


@pytest.mark.parametrize("mode", [4, 5])
def test_inbreeding_loglik(mode):
    K = 4
    d = make_dataset(N=30, L=40, K=K, A=5, miss=0.05, seed=10)
    o, r = Oracle(d.x, d.allelenum, K, mode=mode), Reference(d.x, d.allelenum, K, mode=mode)
    rng = np.random.default_rng(3)
    _state(o, rng, K)
    o.self_rates[...] = rng.uniform(0.05, 0.95, o.self_rates.shape)
    r.set_z(o.z); r.set_qq(o.qq); r.set_freq(o.freq); r.set_alpha(o.alpha); r.set_self(o.self_rates)
    F = rng.uniform(0.0, 1.0, K)
    for i in range(0, 30, 4):
        if mode == 4:
            assert o.log_ld_F(F, 1, i) == r.log_ld_F(F, 1, i)
        else:
            assert o.log_ld_F(F[:1], 0, i) == r.log_ld_F(F[:1], 0, i)
    Ft = rng.uniform(0, 1, o.self_rates.shape)
    assert o.log_ld_F_total(Ft) == r.log_ld_F_total(Ft)
    r.cal_lkh(); o.cal_lkh()
    ind, tot = r.get_lkh()
    assert np.array_equal(ind, o.indvlkh) and tot == o.totallkh


def test_whole_chain_bit_exact_mode0():
    """Mode 0 (mcmc_POP_no_admixture, mcmc.c:90-131): every individual wholly in one cluster;
    CHAIN.z counts the retained samples per (individual, cluster)."""
    K = 3
    d = make_dataset(N=45, L=21, K=K, A=4, miss=0.04, seed=14, pure=True)
    o = Oracle(d.x, d.allelenum, K, mode=0)
    r = Reference(d.x, d.allelenum, K, mode=0)
    o.setseeds(13, 4, 1972); r.setseeds(13, 4, 1972)
    kw = dict(update=160, burnin=60, thinning=5, ckrep=8, nstep_check_empty=10)
    co = o.run_chain(**kw)
    cr = r.mcmc_updating(**kw)
    for k in ["totallkh", "totallkh2", "indvlkh", "qq", "convg"]:
        assert np.array_equal(np.asarray(co[k]), np.asarray(cr[k])), k
    assert np.array_equal(co["qq"].sum(axis=1), np.full(45, 20.0))      # 20 retained samples per individual


def test_mode0_updates_follow_reference_stream():
    K = 4
    d = make_dataset(N=30, L=17, K=K, A=3, miss=0.05, seed=22)
    o, r = Oracle(d.x, d.allelenum, K, mode=0), Reference(d.x, d.allelenum, K, mode=0)
    rng = np.random.default_rng(4)
    o.zz[...] = rng.integers(0, K, size=o.N)
    o.freq[...] = rng.dirichlet(np.ones(o.Amax), size=(K, o.L))
    r.set_zz(o.zz); r.set_freq(o.freq)
    assert np.array_equal(o.tally(), r.update_P(want_tally=True))     # mcmc.c:825-831: counts by zz
    r.set_freq(o.freq)
    for i in range(0, 30, 5):
        for k in range(K):
            assert o.log_ld_indv_K(i, k) == r.log_ld_indv_K(i, k)
    o.setseeds(7, 8, 9); r.setseeds(7, 8, 9)
    o.update_Z(0); r.update_Z(0)
    assert np.array_equal(o.zz, r.get_zz())
    o.cal_lkh(); r.cal_lkh()
    ind, tot = r.get_lkh()
    assert np.array_equal(ind, o.indvlkh) and tot == o.totallkh


def test_single_updates_follow_reference_stream():
    """Each conditional update consumes the RNG like the reference (state compared after each)."""
    K = 3
    d = make_dataset(N=30, L=17, K=K, A=3, miss=0.05, seed=21)
    o, r = Oracle(d.x, d.allelenum, K), Reference(d.x, d.allelenum, K)
    rng = np.random.default_rng(2)
    _state(o, rng, K)
    o.self_rates[...] = [0.2, 0.5, 0.8]
    _push(o, r)
    o.setseeds(101, 202, 303); r.setseeds(101, 202, 303)
    o.update_P(); r.update_P(want_tally=False)
    assert np.array_equal(o.freq, r.get_freq())
    o.update_S_POP(); r.update_S_POP()
    assert np.array_equal(o.self_rates, r.get_self())
    o.update_G(); r.update_G()
    assert np.array_equal(o.gen, r.get_gen())
    o.update_ZQ(0); r.update_ZQ(0)
    assert np.array_equal(o.z, r.get_z()) and np.array_equal(o.qq, r.get_qq())
    o.update_alpha(); r.update_alpha()
    assert o.alpha == r.get_alpha()


def test_reference_reader_matches_generator(tmp_path):
    """The reference's own text reader recodes our synthetic text to the same dense store."""
    from instruct_b200.synth import write_reference_text
    d = make_dataset(N=25, L=12, K=2, A=4, miss=0.06, seed=5)
    p = str(tmp_path / "d.txt")
    write_reference_text(p, d.x, pop=d.pop)
    x, an, mi, _ = pyoracle.ref_read_data(p, 2, d.N, 2, d.L)
    assert np.array_equal(an, d.allelenum)
    assert np.array_equal(x, d.x)
    assert np.array_equal(mi, (d.x < 0).any(axis=2).astype(np.int32))


@pytest.mark.parametrize("case", range(16))
def test_whole_chain_bit_exact_fuzz(case):
    """Random shapes, seeds, schedules and models (modes 1-5, uniform / DP prior where the reference has one, both
    -e settings, both -y settings): the restatement reproduces the reference's chain bit for bit, or both report the
    same empty-cluster retry."""
    rng = np.random.default_rng(3000 + case)
    mode = int(rng.choice([1, 2, 2, 3, 3, 4, 5]))
    prior = int(rng.integers(0, 2)) if mode == 3 else 0
    back_refl = int(rng.integers(0, 2)) if mode in (2, 4) else 1
    K, A = int(rng.integers(2, 5)), int(rng.integers(2, 7))
    N, L = int(rng.integers(12, 60)), int(rng.integers(3, 30))
    d = make_dataset(N=N, L=L, K=K, A=A, miss=float(rng.choice([0.0, 0.05, 0.2])), seed=4000 + case)
    o = Oracle(d.x, d.allelenum, K, mode=mode, prior_flag=prior, back_refl=back_refl)
    r = Reference(d.x, d.allelenum, K, mode=mode, prior_flag=prior, back_refl=back_refl)
    seeds = [int(v) for v in rng.integers(1, 30000, size=3)]
    o.setseeds(*seeds); r.setseeds(*seeds)
    burnin = int(rng.integers(5, 40))
    kw = dict(update=burnin + int(rng.integers(20, 90)), burnin=burnin, thinning=int(rng.integers(1, 6)), ckrep=5,
              nstep_check_empty=int(rng.choice([3, 10 ** 6])))
    if mode in (2, 4):
        kw["initd"] = [float(v) for v in rng.uniform(0.05, 0.95, size=K)]
    co = o.run_chain(**kw)
    cr = r.mcmc_updating(**kw)
    assert co["flag_empty_cluster"] == cr["flag_empty_cluster"]
    if co["flag_empty_cluster"]:
        return
    keys = ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "convg"] + (["self_rates", "self_rates2"] if mode != 1 else []) + \
           (["gen", "gen2"] if mode in (2, 3) else [])
    for k in keys:
        assert np.array_equal(np.asarray(co[k]), np.asarray(cr[k]), equal_nan=True), (k, mode, prior, back_refl)
