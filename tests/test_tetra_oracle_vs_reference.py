"""Pins oracle/tetra_oracle.c (the autotetraploid CPU restatement) to the UNMODIFIED reference
compiled in oracle/_ref/ (poly_geno.c through oracle/ref_harness_poly.c): genotype catalogues,
the float log genotype-frequency tables and whole-chain running moments, bit for bit on
identical Wichmann-Hill seeds.  CPU only."""
import numpy as np
import pytest

from instruct_b200.synth import make_tetra_dataset
from oracle.pyoracle import have_ref
from oracle.pytetra import RefTetra, TetraOracle

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference once)")


def _random_freq(rng, K, L, allelenum, Amax):
    f = rng.dirichlet(np.ones(Amax), size=(K, L))
    for l in range(L):
        a = allelenum[l]
        f[:, l, a:] = 0
        f[:, l, :a] /= f[:, l, :a].sum(axis=1, keepdims=True)
    return f


@pytest.mark.parametrize("A,K", [(2, 2), (3, 3), (4, 3), (6, 2), (8, 2), (10, 2)])
def test_catalogue_and_tables_bit_exact(A, K):
    d = make_tetra_dataset(N=30, L=9, K=K, A=A, miss=0.05, seed=A)
    o = TetraOracle(d.x, d.nd, d.allelenum, K)
    r = RefTetra(d.x, d.nd, d.allelenum, K)
    assert o.Gmax == r.Gmax
    for l in range(d.L):
        assert np.array_equal(o.genolist(l), r.genolist(l))
    rng = np.random.default_rng(A)
    f = _random_freq(rng, K, d.L, d.allelenum, o.Amax)
    S = rng.uniform(0.02, 0.98, size=K)
    o.freq[...] = f
    o.self_rates[...] = S
    o.tables()
    ex, gf = r.tables(f, S)
    assert np.array_equal(o.exfreq, ex)
    assert np.array_equal(o.genofreq, gf)
    # log frequencies are <= 0 (the reference aborts otherwise, poly_geno.c:2015-2019).  They do
    # NOT sum to one: the monoallelic class re-uses a stale index (poly_geno.c:1984-1989), which
    # the oracle reproduces as written.
    for l in range(d.L):
        n = len(o.genolist(l))
        assert (gf[:, l, :n] <= 0).all()


@pytest.mark.parametrize("A,K,miss,back_refl", [(4, 3, 0.05, 1), (3, 2, 0.0, 1), (2, 2, 0.1, 1), (5, 2, 0.02, 0), (9, 2, 0.02, 1)])
def test_whole_chain_bit_exact(A, K, miss, back_refl):
    d = make_tetra_dataset(N=36, L=10, K=K, A=A, miss=miss, seed=10 + A)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, back_refl=back_refl)
    r = RefTetra(d.x, d.nd, d.allelenum, K, back_refl=back_refl)
    o.setseeds(13, 4, 1972)
    r.setseeds(13, 4, 1972)
    initd = np.linspace(0.3, 0.7, K)
    a = o.run_chain(50, 20, 3, ckrep=4, initd=initd)
    b = r.run_chain(50, 20, 3, ckrep=4, initd=initd)
    assert a["flag"] == b["flag"] == 0
    for key in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "convg"]:
        # -e 0 drives a rate to exactly 0 or 1, where the reference's tables are log(0): its
        # likelihoods turn NaN, and the restatement follows it there (equal_nan)
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key]), equal_nan=True), key


# ---- allotetraploid (-p 4 -ap 0): two subgenomes, freq / freq2 ----------

@pytest.mark.parametrize("A,K", [(2, 2), (3, 3), (4, 3), (5, 2), (8, 2), (10, 2)])
def test_allo_catalogue_and_tables_bit_exact(A, K):
    d = make_tetra_dataset(N=30, L=9, K=K, A=A, miss=0.05, seed=20 + A)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0)
    r = RefTetra(d.x, d.nd, d.allelenum, K, autopoly=0)
    assert o.Gmax == r.Gmax
    for l in range(d.L):
        assert np.array_equal(o.genolist(l), r.genolist(l))          # allo_geno_list, poly_geno.c:2050
    rng = np.random.default_rng(A)
    f = _random_freq(rng, K, d.L, d.allelenum, o.Amax)
    f2 = _random_freq(rng, K, d.L, d.allelenum, o.Amax)
    S = rng.uniform(0.02, 0.98, size=K)
    o.freq[...] = f
    o.freq2[...] = f2
    o.self_rates[...] = S
    o.tables()
    ex, gf = r.tables(f, S, freq2=f2)
    assert np.array_equal(o.exfreq, ex)                              # calc_exfreq_allo :1592
    assert np.array_equal(o.genofreq, gf)                            # allo_genfreq :2122
    for l in range(d.L):
        n = len(o.genolist(l))
        assert (gf[:, l, :n] <= 0).all()
        # unlike the autotetraploid tables these ARE a distribution over the catalogue
        np.testing.assert_allclose(np.exp(gf[:, l, :n].astype(np.float64)).sum(axis=1), 1.0, atol=2e-5)


@pytest.mark.parametrize("A,K,miss,back_refl", [(4, 3, 0.05, 1), (3, 2, 0.0, 1), (2, 2, 0.1, 1), (5, 2, 0.02, 1), (8, 2, 0.02, 1)])
def test_allo_whole_chain_bit_exact(A, K, miss, back_refl):
    d = make_tetra_dataset(N=36, L=10, K=K, A=A, miss=miss, seed=30 + A)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, back_refl=back_refl, autopoly=0)
    r = RefTetra(d.x, d.nd, d.allelenum, K, back_refl=back_refl, autopoly=0)
    o.setseeds(13, 4, 1972)
    r.setseeds(13, 4, 1972)
    initd = np.linspace(0.3, 0.7, K)
    a = o.run_chain(50, 20, 3, ckrep=4, initd=initd)
    b = r.run_chain(50, 20, 3, ckrep=4, initd=initd)
    assert a["flag"] == b["flag"] == 0
    for key in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "convg"]:
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key]), equal_nan=True), key


@pytest.mark.parametrize("case", range(10))
def test_tetra_whole_chain_bit_exact_fuzz(case):
    """Random shapes, allele counts, missingness, seeds and schedules, both tetraploid models: bit for bit."""
    rng = np.random.default_rng(5000 + case)
    autopoly = int(rng.integers(0, 2))
    A = int(rng.integers(2, 6 if autopoly == 0 else 7))
    K = int(rng.integers(2, 4))
    N, L = int(rng.integers(10, 40)), int(rng.integers(2, 14))
    d = make_tetra_dataset(N=N, L=L, K=K, A=A, miss=float(rng.choice([0.0, 0.05, 0.25])), seed=6000 + case)
    o = TetraOracle(d.x, d.nd, d.allelenum, K, back_refl=1, autopoly=autopoly)
    r = RefTetra(d.x, d.nd, d.allelenum, K, back_refl=1, autopoly=autopoly)
    seeds = [int(v) for v in rng.integers(1, 30000, size=3)]
    o.setseeds(*seeds)
    r.setseeds(*seeds)
    initd = rng.uniform(0.1, 0.9, size=K)
    burnin = int(rng.integers(3, 15))
    upd, thin = burnin + int(rng.integers(10, 40)), int(rng.integers(1, 5))
    a = o.run_chain(upd, burnin, thin, ckrep=4, initd=initd)
    b = r.run_chain(upd, burnin, thin, ckrep=4, initd=initd)
    assert a["flag"] == b["flag"]
    for key in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "convg"]:
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key]), equal_nan=True), (key, autopoly, A, K)
