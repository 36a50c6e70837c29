"""The oracle against the committed golden fixtures (tests/golden/*.npz), which
tools/make_golden.py generated from the UNMODIFIED reference compiled in oracle/_ref.
Runs anywhere (no reference, no GPU needed): this is what keeps the oracle pinned on the
GPU box, where /root/reference does not exist."""
import glob
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "state_K*.npz"))))
def test_state_fixture(path):
    g = np.load(path)
    K = int(g["K"])
    o = Oracle(g["x"], g["allelenum"], K)
    o.z[...] = g["z"]; o.qq[...] = g["qq"]; o.freq[...] = g["freq"]; o.gen[...] = g["gen"]
    o.self_rates[...] = g["S"]; o.alpha = 0.8
    assert np.array_equal(o.tally(), g["tally"])                      # integer: bit-exact
    assert np.array_equal(o.missing_mask(), g["missindx"])
    for j, gen in enumerate(g["ll_g"]):
        got = np.array([o.log_ld_indv(int(gen), i) for i in range(o.N)])
        np.testing.assert_allclose(got, g["ll"][:, j], rtol=1e-12, atol=0)
    o.cal_lkh()
    np.testing.assert_allclose(o.indvlkh, g["indvlkh"], rtol=1e-12)
    assert abs(o.totallkh - float(g["totallkh"])) <= 1e-12 * abs(float(g["totallkh"]))
    assert abs(o.proposal(g["S"]) - float(g["proposal_S"])) <= 1e-12 * abs(float(g["proposal_S"]))
    assert abs(o.proposal(g["S2"]) - float(g["proposal_S2"])) <= 1e-12 * abs(float(g["proposal_S2"]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "chain_mode*.npz"))))
def test_chain_fixture_bit_exact(path):
    """Whole chains from fixed Wichmann-Hill seeds reproduce the reference's CHAIN moments."""
    g = np.load(path)
    K = int(g["K"])
    o = Oracle(g["x"], g["allelenum"], K, mode=int(g["mode"]), prior_flag=int(g["prior"]), alpha_dpm=float(g["alpha_dpm"]))
    o.setseeds(*[int(v) for v in g["seeds"]])
    c = o.run_chain(update=int(g["kw_update"]), burnin=int(g["kw_burnin"]), thinning=int(g["kw_thinning"]),
                    ckrep=int(g["kw_ckrep"]), nstep_check_empty=int(g["kw_nstep_check_empty"]), initd=g["kw_initd"])
    assert c["flag_empty_cluster"] == int(g["flag_empty_cluster"]) == 0
    keys = ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "gen", "gen2", "convg"]
    if int(g["mode"]) == 0:
        keys = ["totallkh", "totallkh2", "indvlkh", "qq", "convg"]   # qq carries CHAIN.z, the per-cluster sample counts
    if int(g["mode"]) in (4, 5):
        keys = [k for k in keys if k not in ("gen", "gen2")]       # the inbreeding modes keep no generations (mcmc.c:517)
    for k in keys:
        # libm differences between hosts could in principle move the last bit; allow 1e-12
        np.testing.assert_allclose(np.asarray(c[k]), g[k], rtol=1e-12, atol=0, err_msg=k)


def test_gelman_rubin_variants():
    from oracle import pyoracle
    from instruct_b200.converge import gelman_rubin, gelman_rubin_ref_compat
    rng = np.random.default_rng(0)
    tr = rng.normal(size=(4, 12)) + np.array([0, 0.5, 1.0, 1.5])[:, None]
    flat = tr.reshape(-1)
    assert abs(gelman_rubin(tr) - pyoracle.gelman_rubin(flat, 4, 12)) < 1e-12
    # what the reference computes (check_converg.c:67): segments of chain 0 only
    assert abs(gelman_rubin_ref_compat(flat, 4, 12) - pyoracle.gelman_rubin_ref(flat, 4, 12)) < 1e-12
    assert gelman_rubin(tr) > 1.2      # four chains with shifted means do not look converged


def test_tetra_fixture():
    """Autotetraploid: catalogues, float tables and one whole chain against the golden outputs of
    the reference's poly_geno.c (tools/make_golden.py tetra)."""
    from oracle.pytetra import TetraOracle
    g = np.load(os.path.join(GOLD, "tetra_chain.npz"))
    K = int(g["K"])
    o = TetraOracle(g["x"], g["nd"], g["allelenum"], K)
    for l in range(o.L):
        assert np.array_equal(o.genolist(l), g["codes"][l][: len(o.genolist(l))])
    o.freq[...] = g["freq"]
    o.self_rates[...] = g["S"]
    o.tables()
    # float tables: same binary libm here, but allow one float ulp across hosts
    np.testing.assert_allclose(o.exfreq, g["exfreq"], rtol=2e-7, atol=0)
    np.testing.assert_allclose(o.genofreq, g["genofreq"], rtol=2e-7, atol=0)
    o.setseeds(*[int(v) for v in g["seeds"]])
    c = o.run_chain(update=int(g["kw_update"]), burnin=int(g["kw_burnin"]), thinning=int(g["kw_thinning"]),
                    ckrep=int(g["kw_ckrep"]), initd=g["kw_initd"])
    assert c["flag"] == int(g["flag"]) == 0
    for k in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "convg"]:
        np.testing.assert_allclose(np.asarray(c[k]), g[k], rtol=1e-12, atol=0, err_msg=k)


def test_allo_fixture():
    """Allotetraploid (-ap 0): catalogue, two-subgenome float tables and one whole chain against the
    golden outputs of the reference's poly_geno.c (tools/make_golden.py allo)."""
    from oracle.pytetra import TetraOracle
    g = np.load(os.path.join(GOLD, "allo_chain.npz"))
    K = int(g["K"])
    o = TetraOracle(g["x"], g["nd"], g["allelenum"], K, autopoly=0)
    for l in range(o.L):
        assert np.array_equal(o.genolist(l), g["codes"][l][: len(o.genolist(l))])
    o.freq[...] = g["freq"]
    o.freq2[...] = g["freq2"]
    o.self_rates[...] = g["S"]
    o.tables()
    np.testing.assert_allclose(o.exfreq, g["exfreq"], rtol=2e-7, atol=0)
    np.testing.assert_allclose(o.genofreq, g["genofreq"], rtol=2e-7, atol=0)
    o.setseeds(*[int(v) for v in g["seeds"]])
    c = o.run_chain(update=int(g["kw_update"]), burnin=int(g["kw_burnin"]), thinning=int(g["kw_thinning"]),
                    ckrep=int(g["kw_ckrep"]), initd=g["kw_initd"])
    assert c["flag"] == int(g["flag"]) == 0
    for k in ["totallkh", "totallkh2", "indvlkh", "qq", "qq2", "self_rates", "self_rates2", "convg"]:
        np.testing.assert_allclose(np.asarray(c[k]), g[k], rtol=1e-12, atol=0, err_msg=k)


def test_gelman_rubin_params_and_label_alignment():
    """R per parameter and the label-switching alignment it needs."""
    from instruct_b200.converge import align_labels, chain_diagnostics, gelman_rubin, gelman_rubin_params
    rng = np.random.default_rng(1)
    m, n, N, K = 4, 60, 30, 3
    home = rng.integers(0, K, size=N)
    Qtrue = np.full((N, K), 0.05); Qtrue[np.arange(N), home] = 0.9
    Strue = np.array([0.2, 0.5, 0.8])
    perms = [np.arange(K), np.array([2, 0, 1]), np.array([1, 2, 0]), np.array([0, 2, 1])]
    ll = rng.normal(-1000, 3, size=(m, n))
    S = np.empty((m, n, K)); Q = np.empty((m, n, N, K))
    for c in range(m):
        inv = np.argsort(perms[c])
        S[c] = (Strue + rng.normal(0, 0.02, size=(n, K)))[:, inv]          # chain c calls true cluster a "inv[a]"...
        Q[c] = (Qtrue[None] + rng.normal(0, 0.01, size=(n, N, K)))[:, :, inv]
    # per-parameter R agrees with the scalar statistic
    assert abs(gelman_rubin_params(ll[:, :, None])[0] - gelman_rubin(ll)) < 1e-12
    # without alignment the switched labels look unconverged, with it they do not
    assert np.nanmax(gelman_rubin_params(S)) > 2.0
    for c in range(m):
        p = align_labels(Q[0].mean(axis=0), Q[c].mean(axis=0))
        assert np.array_equal(Q[c].mean(axis=0)[:, p].argmax(axis=1), home)
    d = chain_diagnostics(ll, S, Q)
    assert d["R_loglik"] < 1.1 and np.all(d["R_S"] < 1.1) and np.all(d["R_cluster_size"] < 1.1)
    # a chain stuck elsewhere is flagged on S only
    S2 = S.copy(); S2[3, :, np.argsort(perms[3])[1]] += 0.3
    assert chain_diagnostics(ll, S2, Q)["R_S"].max() > 1.5
    # the greedy branch (K > 7) finds a planted permutation too
    K2 = 9
    q0 = np.eye(K2)[rng.integers(0, K2, size=200)] * 0.9 + 0.01
    pp = rng.permutation(K2)
    got = align_labels(q0, q0[:, pp])
    assert np.allclose(q0[:, pp][:, got], q0)
