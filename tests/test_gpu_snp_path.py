"""The opt-in biallelic sweep path (zq_snp.cu, IG_SNP_PATH=1): the same parity gates as the default kernel -- store and
tally bit-exact, the four log-likelihood pieces to 1e-6, the chi-square of the Z draw, whole sweeps in step with the
oracle -- on shapes that cover both chunk lengths (256 and 1024 loci), KP = 4 and 8, a partly filled last chunk and heavy
missingness.  The path is not the default (it measured 5 % slower at config 4, profiles/r2_zq_levers.md) but it is kept
working."""
import numpy as np
import pytest

import test_gpu_parity as P
from instruct_b200 import Sampler, SeqData, _lib
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu

SNP_SHAPES = [(300, 40, 2, 2, 0.0), (64, 130, 8, 2, 0.1), (70, 1300, 8, 2, 0.05), (45, 600, 3, 2, 0.02), (40, 300, 7, 2, 0.5)]


@pytest.fixture(autouse=True)
def _snp_path(monkeypatch):
    monkeypatch.setenv("IG_SNP_PATH", "1")


@pytest.mark.parametrize("N,L,K,A,miss", SNP_SHAPES)
def test_path_is_taken_and_store_roundtrips(N, L, K, A, miss):
    d, sd = P._mk(N, L, K, A, miss, seed=1)
    s = Sampler(sd)
    geo = s.geometry()
    assert geo["TL"] in (256, 1024) and geo["R"] == 1 and geo["A"] == 2      # the biallelic path's geometry
    o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(0)
    P._inject(s, o, rng)
    assert np.array_equal(s.get(_lib.STATE_Z), o.z)                          # canonical -> class-sorted -> canonical
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    s.close()


@pytest.mark.parametrize("N,L,K,A,miss", SNP_SHAPES)
def test_fused_sweep_pieces(N, L, K, A, miss):
    P.test_fused_sweep_pieces(N, L, K, A, miss, 1)


@pytest.mark.parametrize("K", [6, 3])
def test_z_draw(K):
    P.test_z_draw_matches_exact_conditional(K, 2)


def test_whole_sweeps_keep_tally_and_likelihood_in_step():
    d, sd = P._mk(300, 1100, 8, 2, 0.03, seed=12)
    s = Sampler(sd, seed=3)
    assert s.geometry()["TL"] == 1024
    s.chain_init(0, initd=np.linspace(0.2, 0.8, 8))
    s.sweep(5)
    o = Oracle(d.x, d.allelenum, 8)
    o.z[...] = s.get(_lib.STATE_Z)
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(float), o.count_z())
    o.freq[...] = s.get(_lib.STATE_P)
    g = s.get(_lib.STATE_G)
    want = np.array([o.log_ld_indv(g[i], i) for i in range(o.N)])
    got = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= P.RTOL
    s.close()
