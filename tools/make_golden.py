#!/usr/bin/env python
"""Generate the committed golden fixtures in tests/golden/ from the UNMODIFIED reference
(the include-the-.c harness in oracle/_ref, built by oracle/Makefile from /root/reference).

The reference ships no golden vectors of its own, so these are the
"outputs of the reference itself run here" that pin the oracle and give the GPU path its
statistical target.  Run in the authoring container only:

    python tools/make_golden.py            # writes tests/golden/*.npz (small, committed)

Fixtures
  state_K{K}.npz      injected state + the reference's tally / missing mask / log_ld_indv /
                      proposal / cal_lkh on it                       (parity levels 1 and 2)
  chain_mode{m}.npz   CHAIN moments of one short chain from fixed Wichmann-Hill seeds through
                      the reference's own mcmc_updating()            (whole-chain pin)
  posterior_c1.npz    config-1 shaped data (K=2 N=200 L=10), posterior means of S, Q, log-lik
                      from R independent reference chains            (parity level 3)
  posterior_mode3_prior{0,1}.npz  BASELINE configs[2] in small: selfing rate per individual under the uniform and the
                      Dirichlet-process prior; posterior means from R reference chains, and the number of DP clusters
  tetra_chain.npz     autotetraploid: genotype catalogues, float tables on an injected state and
                      the CHAIN moments of one short chain through mcmc_POP_tetra_selfing
  posterior_tetra.npz autotetraploid posterior means from R independent reference chains
  allo_chain.npz / posterior_allo.npz   the same two for the allotetraploid model (-ap 0)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from instruct_b200.synth import make_dataset, make_tetra_dataset  # noqa: E402
from oracle.pyoracle import Reference  # noqa: E402
from oracle.pytetra import RefTetra  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def state_fixture(K, A, miss, seed):
    d = make_dataset(N=48, L=29, K=K, A=A, miss=miss, seed=seed)
    rng = np.random.default_rng(seed + 100)
    r = Reference(d.x, d.allelenum, K)
    z = rng.integers(0, K, size=d.x.shape).astype(np.int8)
    qq = rng.dirichlet(np.ones(K) * 0.7, size=d.N)
    freq = rng.dirichlet(np.ones(r.Amax), size=(K, d.L))
    gen = rng.integers(1, 9, size=d.N).astype(np.int32)
    S = rng.uniform(0.1, 0.9, K)
    r.set_z(z); r.set_qq(qq); r.set_freq(freq); r.set_gen(gen); r.set_self(S); r.set_alpha(0.8)
    tally = r.update_P(want_tally=True)          # also redraws freq: restore it
    r.set_freq(freq)
    ll = np.array([[r.log_ld_indv(g, i) for g in (1, 2, 5, 50)] for i in range(d.N)])
    r.cal_lkh()
    indv, tot = r.get_lkh()
    S2 = rng.uniform(0.1, 0.9, K)
    np.savez_compressed(os.path.join(OUT, f"state_K{K}.npz"), x=d.x, allelenum=d.allelenum, K=K, z=z, qq=qq,
                        freq=freq, gen=gen, S=S, S2=S2, tally=tally, missindx=r.missindx().astype(np.uint8),
                        ll_g=np.array([1, 2, 5, 50]), ll=ll, indvlkh=indv, totallkh=tot,
                        proposal_S=r.proposal(S), proposal_S2=r.proposal(S2))


def chain_fixture(mode, prior):
    K = 3
    d = make_dataset(N=40, L=18, K=K, A=4, miss=0.05, seed=31, s_atoms=[0.1, 0.5, 0.9] if mode == 3 else None)
    r = Reference(d.x, d.allelenum, K, mode=mode, prior_flag=prior, alpha_dpm=2.0)
    r.setseeds(13, 4, 1972)
    kw = dict(update=120, burnin=40, thinning=4, ckrep=6, nstep_check_empty=10, initd=[0.25, 0.5, 0.75])
    c = r.mcmc_updating(**kw)
    np.savez_compressed(os.path.join(OUT, f"chain_mode{mode}_prior{prior}.npz"), x=d.x, allelenum=d.allelenum, K=K,
                        mode=mode, prior=prior, seeds=[13, 4, 1972], alpha_dpm=2.0, **{f"kw_{k}": v for k, v in kw.items()},
                        **{k: np.asarray(v) for k, v in c.items()})


def posterior_fixture(R=12):
    """BASELINE.json configs[0] shape: K=2 N=200 L=10 microsatellite, mode 2, uniform prior."""
    K = 2
    d = make_dataset(N=200, L=10, K=K, A=8, miss=0.0, seed=1001, pure=True)
    S, Qm, LL, G = [], [], [], []
    for rep in range(R):
        r = Reference(d.x, d.allelenum, K, mode=2)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        c = r.mcmc_updating(update=6000, burnin=2000, thinning=10, ckrep=5, nstep_check_empty=20,
                            initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
        # label switching: order clusters by their mean selfing rate
        o = np.argsort(c["self_rates"])
        S.append(c["self_rates"][o]); Qm.append(c["qq"][:, o]); LL.append(c["totallkh"]); G.append(c["gen"])
        print("posterior rep", rep, c["self_rates"][o], c["totallkh"], flush=True)
    np.savez_compressed(os.path.join(OUT, "posterior_c1.npz"), x=d.x, allelenum=d.allelenum, K=K, S=np.array(S),
                        Q=np.array(Qm).astype(np.float32), LL=np.array(LL), G=np.array(G).astype(np.float32),
                        S_true=d.S_true, pop=d.pop, update=6000, burnin=2000, thinning=10)


def posterior_mode1_fixture(R=10):
    """Mode 1 (the CLI default, mcmc_POP_admixture mcmc.c:135): admixture without selfing."""
    K = 2
    d = make_dataset(N=200, L=10, K=K, A=8, miss=0.0, seed=1001, pure=False, own=0.85)
    Qm, LL = [], []
    for rep in range(R):
        r = Reference(d.x, d.allelenum, K, mode=1)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        c = r.mcmc_updating(update=4000, burnin=1500, thinning=10, ckrep=5, nstep_check_empty=20)
        o = np.argsort(c["qq"][d.pop == 0].mean(axis=0))[::-1]
        Qm.append(c["qq"][:, o]); LL.append(c["totallkh"])
        print("mode-1 posterior rep", rep, c["totallkh"], c["qq"][d.pop == 0].mean(axis=0)[o], flush=True)
    np.savez_compressed(os.path.join(OUT, "posterior_mode1.npz"), x=d.x, allelenum=d.allelenum, K=K, Q=np.array(Qm).astype(np.float32),
                        LL=np.array(LL), pop=d.pop, update=4000, burnin=1500, thinning=10)


def tetra_fixtures(R=12):
    """poly_geno.c through the harness: tables + one whole chain, and the posterior target."""
    K = 3
    d = make_tetra_dataset(N=36, L=10, K=K, A=4, miss=0.05, seed=21)
    r = RefTetra(d.x, d.nd, d.allelenum, K)
    rng = np.random.default_rng(22)
    f = rng.dirichlet(np.ones(r.Amax), size=(K, d.L))
    S = rng.uniform(0.05, 0.95, size=K)
    ex, gf = r.tables(f, S)
    r.setseeds(13, 4, 1972)
    kw = dict(update=50, burnin=20, thinning=3, ckrep=4, initd=[0.3, 0.5, 0.7])
    c = r.run_chain(**kw)
    np.savez_compressed(os.path.join(OUT, "tetra_chain.npz"), x=d.x, nd=d.nd, allelenum=d.allelenum, K=K, freq=f, S=S,
                        exfreq=ex, genofreq=gf, codes=np.stack([r.genolist(l) for l in range(d.L)]), seeds=[13, 4, 1972],
                        **{f"kw_{k}": v for k, v in kw.items()}, **{k: np.asarray(v) for k, v in c.items()})
    K = 2
    d = make_tetra_dataset(N=120, L=30, K=K, A=4, miss=0.02, seed=23)
    Ss, Qm, LL, alphas = [], [], [], []
    for rep in range(R):
        r = RefTetra(d.x, d.nd, d.allelenum, K)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        # the tetraploid driver draws alpha = ran1() * 10 once (poly_geno.c:386) and never updates it,
        # so every chain samples a DIFFERENT posterior: record it (first Wichmann-Hill draw, random.c:34-47)
        s1, s2, s3 = (171 * (13 + 7 * rep)) % 30269, (172 * (4 + 3 * rep)) % 30307, (170 * (1972 + 11 * rep)) % 30323
        alphas.append(10 * ((s1 / 30269.0 + s2 / 30307.0 + s3 / 30323.0) % 1.0))
        c = r.run_chain(update=1500, burnin=500, thinning=5, ckrep=5, initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
        o = np.argsort(c["qq"][d.pop == 0].mean(axis=0))[::-1]        # label switching: cluster 0 = home of pop 0
        Ss.append(c["self_rates"][o]); Qm.append(c["qq"][:, o]); LL.append(c["totallkh"])
        print("tetra posterior rep", rep, c["self_rates"][o], c["totallkh"], flush=True)
    np.savez_compressed(os.path.join(OUT, "posterior_tetra.npz"), x=d.x, nd=d.nd, allelenum=d.allelenum, K=K, S=np.array(Ss),
                        Q=np.array(Qm).astype(np.float32), LL=np.array(LL), pop=d.pop, update=1500, burnin=500, thinning=5,
                        alpha=np.array(alphas))


def allo_fixtures(R=12):
    """Allotetraploid (-p 4 -ap 0) through the same harness: catalogue, tables with two subgenomes,
    one whole chain, and the posterior target (poly_geno.c *_allo)."""
    K = 3
    d = make_tetra_dataset(N=36, L=10, K=K, A=4, miss=0.05, seed=41)
    r = RefTetra(d.x, d.nd, d.allelenum, K, autopoly=0)
    rng = np.random.default_rng(42)
    f = rng.dirichlet(np.ones(r.Amax), size=(K, d.L))
    f2 = rng.dirichlet(np.ones(r.Amax), size=(K, d.L))
    S = rng.uniform(0.05, 0.95, size=K)
    ex, gf = r.tables(f, S, freq2=f2)
    r.setseeds(13, 4, 1972)
    kw = dict(update=50, burnin=20, thinning=3, ckrep=4, initd=[0.3, 0.5, 0.7])
    c = r.run_chain(**kw)
    np.savez_compressed(os.path.join(OUT, "allo_chain.npz"), x=d.x, nd=d.nd, allelenum=d.allelenum, K=K, freq=f, freq2=f2, S=S,
                        exfreq=ex, genofreq=gf, codes=np.stack([r.genolist(l) for l in range(d.L)]), seeds=[13, 4, 1972],
                        **{f"kw_{k}": v for k, v in kw.items()}, **{k: np.asarray(v) for k, v in c.items()})
    K = 2
    d = make_tetra_dataset(N=120, L=30, K=K, A=4, miss=0.02, seed=43)
    Ss, Qm, LL, alphas = [], [], [], []
    for rep in range(R):
        r = RefTetra(d.x, d.nd, d.allelenum, K, autopoly=0)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        s1, s2, s3 = (171 * (13 + 7 * rep)) % 30269, (172 * (4 + 3 * rep)) % 30307, (170 * (1972 + 11 * rep)) % 30323
        alphas.append(10 * ((s1 / 30269.0 + s2 / 30307.0 + s3 / 30323.0) % 1.0))      # poly_geno.c:386, as above
        c = r.run_chain(update=1500, burnin=500, thinning=5, ckrep=5, initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
        o = np.argsort(c["qq"][d.pop == 0].mean(axis=0))[::-1]
        Ss.append(c["self_rates"][o]); Qm.append(c["qq"][:, o]); LL.append(c["totallkh"])
        print("allo posterior rep", rep, c["self_rates"][o], c["totallkh"], flush=True)
    np.savez_compressed(os.path.join(OUT, "posterior_allo.npz"), x=d.x, nd=d.nd, allelenum=d.allelenum, K=K, S=np.array(Ss),
                        Q=np.array(Qm).astype(np.float32), LL=np.array(LL), pop=d.pop, update=1500, burnin=500, thinning=5,
                        alpha=np.array(alphas))


def posterior_mode0_fixture(R=10):
    """Mode 0 (mcmc_POP_no_admixture, mcmc.c:90): whole-individual assignment; CHAIN.z / steps."""
    K = 2
    d = make_dataset(N=200, L=10, K=K, A=8, miss=0.0, seed=1001, pure=True)
    Zm, LL = [], []
    for rep in range(R):
        r = Reference(d.x, d.allelenum, K, mode=0)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        c = r.mcmc_updating(update=3000, burnin=1000, thinning=10, ckrep=5, nstep_check_empty=20)
        z = c["qq"] / c["qq"].sum(axis=1, keepdims=True)
        o = np.argsort(z[d.pop == 0].mean(axis=0))[::-1]
        Zm.append(z[:, o]); LL.append(c["totallkh"])
        print("mode-0 posterior rep", rep, c["totallkh"], z[d.pop == 0].mean(axis=0)[o], flush=True)
    np.savez_compressed(os.path.join(OUT, "posterior_mode0.npz"), x=d.x, allelenum=d.allelenum, K=K, Z=np.array(Zm).astype(np.float32),
                        LL=np.array(LL), pop=d.pop, update=3000, burnin=1000, thinning=10)


def posterior_inbreeding_fixture(mode, R=10):
    """Modes 4/5 (mcmc_POP_inbreedcoff mcmc.c:242, mcmc_INDV_inbreedcoff :386, uniform prior):
    inbreeding coefficients per population / per individual on the config-1 shaped data set."""
    K = 2
    d = make_dataset(N=200, L=10, K=K, A=8, miss=0.0, seed=1001, pure=True)
    F, Qm, LL = [], [], []
    for rep in range(R):
        r = Reference(d.x, d.allelenum, K, mode=mode)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        c = r.mcmc_updating(update=5000, burnin=2000, thinning=10, ckrep=5, nstep_check_empty=20,
                            initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
        o = np.argsort(c["qq"][d.pop == 0].mean(axis=0))[::-1]          # cluster that holds population 0 first
        F.append(c["self_rates"][o] if mode == 4 else c["self_rates"])
        Qm.append(c["qq"][:, o]); LL.append(c["totallkh"])
        print(f"mode-{mode} posterior rep", rep, c["totallkh"], F[-1][:4], flush=True)
    np.savez_compressed(os.path.join(OUT, f"posterior_mode{mode}.npz"), x=d.x, allelenum=d.allelenum, K=K, F=np.array(F),
                        Q=np.array(Qm).astype(np.float32), LL=np.array(LL), pop=d.pop, update=5000, burnin=2000, thinning=10)


def posterior_mode3_fixture(prior, R=10):
    """Mode 3 (mcmc_INDV_selfing, mcmc.c:297-383): selfing rate per individual under the uniform prior (update_S_IND,
    mcmc.c:864) or the Dirichlet-process prior (update_DP, DPMM.c:165) -- BASELINE configs[2] in small.  CHAIN moments of R
    chains through the reference's own driver; for the DP prior also the number of clusters, from R runs stepped through
    the reference's own update functions in the driver's order (refh_sweeps), since CHAIN does not carry it."""
    K = 2
    d = make_dataset(N=200, L=10, K=K, A=8, miss=0.0, seed=1003, pure=True, s_atoms=[0.05, 0.5, 0.9])
    upd, burn, thin = 5000, 2000, 10
    S, Qm, LL, G, NC = [], [], [], [], []
    for rep in range(R):
        r = Reference(d.x, d.allelenum, K, mode=3, prior_flag=prior, alpha_dpm=2.0)
        r.setseeds(13 + 7 * rep, 4 + 3 * rep, 1972 + 11 * rep)
        c = r.mcmc_updating(update=upd, burnin=burn, thinning=thin, ckrep=5, nstep_check_empty=20)
        o = np.argsort(c["qq"][d.pop == 0].mean(axis=0))[::-1]
        S.append(c["self_rates"]); Qm.append(c["qq"][:, o]); LL.append(c["totallkh"]); G.append(c["gen"])
        print(f"mode-3 prior {prior} rep", rep, c["totallkh"], c["self_rates"][:4], flush=True)
        if prior == 1:
            r2 = Reference(d.x, d.allelenum, K, mode=3, prior_flag=1, alpha_dpm=2.0)
            r2.setseeds(17 + 5 * rep, 9 + 2 * rep, 1900 + 13 * rep)
            rng = np.random.default_rng(50 + rep)
            r2.init_DP()                                             # mcmc.c:318-324
            s0 = r2.get_self()
            r2.set_gen(np.minimum(rng.geometric(1.0 - np.clip(s0, 1e-9, 1 - 1e-9)), 50).astype(np.int32))   # mcmc.c:326-331
            r2.set_alpha(10.0 * rng.random())                        # initial_chn, mcmc.c:479
            r2.update_ZQ(1)
            r2.sweeps(burn)
            nc = []
            for _ in range((upd - burn) // thin):
                r2.sweeps(thin)
                nc.append(r2.dp_nclusters())
            NC.append(np.mean(nc))
            print("   clusters", NC[-1], flush=True)
    np.savez_compressed(os.path.join(OUT, f"posterior_mode3_prior{prior}.npz"), x=d.x, allelenum=d.allelenum, K=K, S=np.array(S),
                        Q=np.array(Qm).astype(np.float32), LL=np.array(LL), G=np.array(G).astype(np.float32), NC=np.array(NC),
                        S_true=d.S_true, pop=d.pop, update=upd, burnin=burn, thinning=thin, alpha_dpm=2.0)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mode3":
        posterior_mode3_fixture(0)
        posterior_mode3_fixture(1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mode0":
        chain_fixture(0, 0)
        posterior_mode0_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "inbreeding":
        chain_fixture(4, 0)
        chain_fixture(5, 0)
        posterior_inbreeding_fixture(4)
        posterior_inbreeding_fixture(5)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mode1":
        posterior_mode1_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tetra":
        os.makedirs(OUT, exist_ok=True)
        tetra_fixtures()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "allo":
        os.makedirs(OUT, exist_ok=True)
        allo_fixtures()
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    state_fixture(2, 3, 0.0, 1)
    state_fixture(5, 6, 0.06, 2)
    state_fixture(8, 2, 0.1, 3)
    chain_fixture(2, 0)
    chain_fixture(3, 0)
    chain_fixture(3, 1)
    chain_fixture(0, 0)
    posterior_mode0_fixture()
    chain_fixture(4, 0)
    chain_fixture(5, 0)
    posterior_inbreeding_fixture(4)
    posterior_inbreeding_fixture(5)
    posterior_mode3_fixture(0)
    posterior_mode3_fixture(1)
    posterior_fixture()
    posterior_mode1_fixture()
    tetra_fixtures()
    allo_fixtures()
