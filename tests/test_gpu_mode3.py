"""BASELINE.json configs[2] -- selfing rates per individual (`-v 3`), uniform prior (`-f 0`, update_S_IND mcmc.c:864-886)
and Dirichlet-process prior (`-f 1`, update_DP DPMM.c:165-199, gen_post_prob :361-377, sample_poster :392-398) -- on the
CUDA path through the C-ABI, against the oracle on the same injected state and against chains of the compiled reference.

  * the weights the DP step hands to disc_unif, alpha/((G+1)G) and num * dgeom(S_c, G): 1e-6 on an identical state;
  * one Chinese-restaurant scan from an injected clustering: where individuals end up, how many clusters there are and
    the Beta(G, 2) draws of new clusters, against many scans of orc_update_DP from the same state;
  * update_S_IND from an injected (S, G): acceptance rates and means against the exact Metropolis kernel (quadrature of
    min(1, dgeom(s', G) / dgeom(s, G)) over the reflected proposal) and against orc_update_S_IND;
  * posterior means of S, G, log-likelihood and the number of clusters against R chains of the reference
    (tests/golden/posterior_mode3_prior{0,1}.npz, tools/make_golden.py mode3)."""
import os

import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-6


def _dp_state(N, rng):
    atoms = np.array([0.07, 0.33, 0.5, 0.81, 0.93])
    S = atoms[rng.integers(0, len(atoms), N)]
    S[7] = 0.61                                     # a singleton: its cluster disappears when it leaves (delete, DPMM.c:280)
    G = rng.integers(1, 12, N).astype(np.int32)
    G[3] = 50
    G[11] = 1
    return S, G


def _mk(prior, N=150, L=24, seed=14):
    d = make_dataset(N=N, L=L, K=3, A=4, miss=0.03, seed=seed, s_atoms=[0.05, 0.5, 0.9])
    sd = SeqData(d.x, d.allelenum, 3, mode=3, prior_flag=prior, alpha_dpm=2.0)
    return d, sd


def test_dp_weights_match_gen_post_prob():
    """gen_post_prob, DPMM.c:367-377, on an identical state: new-cluster weight alpha * B(G, 2) = alpha / ((G+1) G) and
    num_c * dgeom(S_c, G) for every other cluster in value order, with j itself taken out first."""
    d, sd = _mk(1)
    s = Sampler(sd, seed=3)
    s.chain_init(0)
    o = Oracle(d.x, d.allelenum, 3, mode=3, prior_flag=1, alpha_dpm=2.0)
    rng = np.random.default_rng(5)
    S, G = _dp_state(d.N, rng)
    s.set(_lib.STATE_G, G)
    s.set(_lib.STATE_S, S)
    assert int(s.get(_lib.STATE_DPCLUSTERS)[0]) == len(np.unique(S))
    w = s.get(_lib.STATE_DPWEIGHTS)
    for j in range(d.N):
        vals, cnt = np.unique(np.delete(S, j), return_counts=True)            # ascending = the list order find/insert keep
        want = [2.0 / (G[j] + 1) / G[j]] + [c * o.dgeom(v, G[j]) for v, c in zip(vals, cnt)]
        got = w[j, : len(want)]
        assert np.all(w[j, len(want):] == 0)
        assert np.max(np.abs(got - np.array(want)) / np.maximum(np.abs(want), 1e-300)) <= RTOL, (j, got, want)
    s.close()


def test_dp_scan_matches_oracle_from_identical_state():
    """One update_DP scan (DPMM.c:165-199) from the same clustering and the same G, many times on both sides (the streams
    cannot match): per-individual mean of the new S, how often an individual opens a new cluster, the number of clusters
    after the scan, and the Beta(G, 2) law of the values of new clusters (sample_poster, DPMM.c:395)."""
    d, sd = _mk(1)
    N = d.N
    s = Sampler(sd, seed=9)
    s.chain_init(0)
    o = Oracle(d.x, d.allelenum, 3, mode=3, prior_flag=1, alpha_dpm=2.0)
    rng = np.random.default_rng(6)
    S0, G = _dp_state(N, rng)
    atoms = set(np.unique(S0).tolist())
    R = 300
    gS, gNew, gNC, beta_z = np.zeros((R, N)), np.zeros((R, N)), np.zeros(R), []
    for r in range(R):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_G, G)
        s.set(_lib.STATE_S, S0)
        s.run_phase(_lib.PHASE_UPDATE_S)
        S1 = s.get(_lib.STATE_S)
        gS[r] = S1
        gNew[r] = [v not in atoms for v in S1]
        gNC[r] = int(s.get(_lib.STATE_DPCLUSTERS)[0])
        assert gNC[r] == len(np.unique(S1))
        # the creator of a new value is the first individual (scan order) that carries it: its value ~ Beta(G_creator, 2)
        seen = set()
        for j in range(N):
            v = S1[j]
            if v in atoms or v in seen:
                continue
            seen.add(v)
            g = float(G[j])
            beta_z.append((v - g / (g + 2)) / np.sqrt(2 * g / ((g + 2) ** 2 * (g + 3))))
    oS, oNew, oNC = np.zeros((R, N)), np.zeros((R, N)), np.zeros(R)
    o.gen[...] = G
    o.setseeds(13, 4, 1972)               # seeded once: the Wichmann-Hill streams of neighbouring seeds are linearly related
    for r in range(R):
        o.self_rates[...] = S0
        o.dp_from_values()
        o.update_DP()
        oS[r] = o.self_rates
        oNew[r] = [v not in atoms for v in o.self_rates]
        oNC[r] = o.dp_nclusters()

    def z(a, b):
        se = np.sqrt(a.var(axis=0, ddof=1) / len(a) + b.var(axis=0, ddof=1) / len(b))
        return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-9)
    zS = z(gS, oS)
    assert np.abs(zS).max() < 4.8, (np.abs(zS).max(), np.argmax(np.abs(zS)))
    assert np.mean(zS ** 2) < 1.5, np.mean(zS ** 2)                # no systematic shift hiding under the noise
    zNC = z(gNC[:, None], oNC[:, None])
    assert abs(zNC[0]) < 4.0, (gNC.mean(), oNC.mean())
    # opening a new cluster: pooled over individuals with the same G (the weight alpha/((G+1)G) depends on G only)
    for g in np.unique(G):
        m = G == g
        a, b = gNew[:, m].mean(axis=1), oNew[:, m].mean(axis=1)
        zz = (a.mean() - b.mean()) / max(np.sqrt(a.var(ddof=1) / R + b.var(ddof=1) / R), 1e-9)
        assert abs(zz) < 4.5, (g, a.mean(), b.mean())
    bz = np.array(beta_z)
    assert len(bz) > 300
    assert abs(bz.mean()) * np.sqrt(len(bz)) < 4.5 and abs(bz.var() - 1.0) < 0.25, (bz.mean(), bz.var(), len(bz))
    s.close()


def _mh_kernel(s, g, npts=4001):
    """Exact acceptance probability and mean of update_S_IND's Metropolis step at (s, g): proposal s + U(-0.05, 0.05)
    reflected into [0, 1] (mcmc.c:872-875), ratio dgeom(s', g) / dgeom(s, g) (mcmc.c:877-879)."""
    t = s + np.linspace(-0.05, 0.05, npts)
    t = np.where(t <= 0, -t, t)
    t = np.where(t >= 1, 2.0 - t, t)
    dg = lambda v: v ** (g - 1) * (1 - v)
    acc = np.minimum(1.0, dg(t) / dg(s))
    w = np.full(npts, 1.0); w[0] = w[-1] = 0.5; w /= w.sum()       # trapezoid
    pa = float((acc * w).sum())
    mean = float(s + ((t - s) * acc * w).sum())
    return pa, mean


def test_update_s_ind_matches_exact_metropolis_kernel():
    d, sd = _mk(0, N=200)
    N = d.N
    s = Sampler(sd, seed=4)
    s.chain_init(0)
    o = Oracle(d.x, d.allelenum, 3, mode=3, prior_flag=0)
    rng = np.random.default_rng(8)
    S0 = rng.uniform(0.05, 0.95, N)
    S0[:6] = [0.004, 0.03, 0.97, 0.996, 0.5, 0.049]                # both reflections
    G = rng.integers(1, 10, N).astype(np.int32)
    G[:6] = [1, 6, 2, 1, 50, 3]
    R = 400
    acc, mean = np.zeros(N), np.zeros(N)
    for r in range(R):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_G, G)
        s.set(_lib.STATE_S, S0)
        s.run_phase(_lib.PHASE_UPDATE_S)
        S1 = s.get(_lib.STATE_S)
        assert np.all((S1 >= 0) & (S1 <= 1))
        moved = S1 != S0
        assert np.all(np.abs(S1 - S0) <= 0.05 + 1e-12)                 # a reflected proposal stays within the step of s as well
        acc += moved
        mean += S1
    acc /= R
    mean /= R
    oacc, omean = np.zeros(N), np.zeros(N)
    o.gen[...] = G
    o.setseeds(13, 4, 1972)
    for r in range(R):
        o.self_rates[...] = S0
        o.update_S_IND()
        oacc += o.self_rates != S0
        omean += o.self_rates
    oacc /= R
    omean /= R
    zs, zo = [], []
    for i in range(N):
        pa, mu = _mh_kernel(S0[i], int(G[i]))
        se = np.sqrt(max(pa * (1 - pa), 1e-6) / R)
        zs.append((acc[i] - pa) / se)
        zo.append((oacc[i] - pa) / se)
        assert abs(mean[i] - mu) < 0.006, (i, S0[i], G[i], mean[i], mu)
    zs, zo = np.array(zs), np.array(zo)
    assert np.abs(zs).max() < 4.8, (np.abs(zs).max(), np.argmax(np.abs(zs)))
    assert np.mean(zs ** 2) < 1.5, np.mean(zs ** 2)
    assert np.abs(zo).max() < 4.8                                   # the oracle obeys the same kernel: the quadrature is the right yardstick
    assert abs(acc.mean() - oacc.mean()) < 0.0065, (acc.mean(), oacc.mean())     # pooled over all individuals: 80 000 decisions a side
    s.close()


@pytest.mark.parametrize("prior", [0, 1])
def test_mode3_posterior_matches_reference_within_mcse(prior):
    g = np.load(os.path.join(GOLD, f"posterior_mode3_prior{prior}.npz"))
    K, R = int(g["K"]), g["S"].shape[0]
    upd, burn, thin = int(g["update"]), int(g["burnin"]), int(g["thinning"])
    sd = SeqData(g["x"], g["allelenum"], K, mode=3, prior_flag=prior, alpha_dpm=float(g["alpha_dpm"]))
    atoms = np.unique(g["S_true"])
    grp = [g["S_true"] == a for a in atoms]
    S, LL, G, NC = [], [], [], []
    for rep in range(R):
        s = Sampler(sd, update=upd, burnin=burn, thinning=thin, ckrep=5, seed=4000 + rep)
        ch, _ = s.run_chain(rep)
        assert ch.flag_empty_cluster == 0 and ch.step == ch.steps
        S.append([ch.self_rates[m].mean() for m in grp]); LL.append(ch.totallkh); G.append(ch.gen.mean())
        if prior == 1 and rep < 6:                          # the number of clusters is not in CHAIN: step a second chain by hand
            s.chain_init(100 + rep)
            s.sweep(burn)
            nc = []
            for _ in range((upd - burn) // thin // 2):
                s.sweep(thin)
                nc.append(int(s.get(_lib.STATE_DPCLUSTERS)[0]))
            NC.append(np.mean(nc))
        s.close()
    S, LL, G = np.array(S), np.array(LL), np.array(G)
    refS = np.stack([g["S"][:, m].mean(axis=1) for m in grp], axis=1)

    def z(a, b):
        a, b = np.asarray(a, float), np.asarray(b, float)
        se = np.sqrt(a.var(axis=0, ddof=1) / len(a) + b.var(axis=0, ddof=1) / len(b))
        return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-12)
    zS, zLL, zG = z(S, refS), z(LL[:, None], g["LL"][:, None]), z(G[:, None], g["G"].mean(axis=1)[:, None])
    msg = f"zS={zS} zLL={zLL} zG={zG} S_gpu={S.mean(0)} S_ref={refS.mean(0)} LL_gpu={LL.mean()} LL_ref={g['LL'].mean()} G_gpu={G.mean()} G_ref={g['G'].mean()}"
    assert np.all(np.abs(zS) < 3.5), msg
    assert np.all(np.abs(zLL) < 3.5), msg
    assert np.all(np.abs(zG) < 3.5), msg
    assert np.all(np.abs(S.mean(0) - refS.mean(0)) < 0.03), msg
    assert abs(LL.mean() - g["LL"].mean()) < 5.0, msg
    if prior == 1:
        zNC = z(np.array(NC)[:, None], g["NC"][:, None])
        assert abs(zNC[0]) < 3.5 and abs(np.mean(NC) - g["NC"].mean()) < 0.15 * g["NC"].mean(), (np.mean(NC), g["NC"].mean(), zNC)
