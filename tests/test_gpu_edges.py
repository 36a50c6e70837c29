"""Edge cases of the hot path through the C-ABI (empty and ragged inputs,
maximum sizes, missing data): degenerate shapes, the largest supported K, many alleles (the
un-replicated histogram path), individuals and loci with nothing but missing data, monomorphic
loci.  Every case is checked against the oracle on the identical state."""
import numpy as np
import pytest

import instruct_b200
from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _check_state(s, d, K, sweeps=3, mode=2):
    s.chain_init(0, initd=np.linspace(0.2, 0.8, K))
    s.sweep(sweeps)
    o = Oracle(d.x, d.allelenum, K, mode=mode)
    z = s.get(_lib.STATE_Z)
    assert z.min() >= 0 and z.max() < K
    o.z[...] = z
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    o.freq[...] = s.get(_lib.STATE_P)
    g = s.get(_lib.STATE_G)
    want = np.array([o.log_ld_indv(g[i], i) for i in range(o.N)])
    got = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    q = s.get(_lib.STATE_Q)
    np.testing.assert_allclose(q.sum(axis=1), 1.0, rtol=1e-12)
    assert np.isfinite(s.get(_lib.STATE_TOTALLKH)[0])


@pytest.mark.parametrize("N,L,K,A", [(1, 1, 1, 2), (1, 37, 3, 4), (513, 1, 2, 3), (2, 8, 16, 2), (33, 9, 16, 5), (300, 7, 2, 40),
                                     (40, 3, 3, 120)])
def test_degenerate_and_extreme_shapes(N, L, K, A):
    # built by hand: the generator (like the reference's reader) drops loci that come out monomorphic,
    # which at N = 1 is all of them; allelenum is what the reader would report for A observed alleles
    rng = np.random.default_rng(N + L + K + A)

    class D:
        pass
    d = D()
    d.x = rng.integers(0, A, size=(L, N, 2)).astype(np.int16)
    d.allelenum = np.full(L, A, dtype=np.int32)
    s = Sampler(SeqData(d.x, d.allelenum, K), seed=3)
    _check_state(s, d, K)
    s.close()


def test_empty_data_set_is_an_error():
    with pytest.raises(instruct_b200.InstructError):
        Sampler(SeqData(np.zeros((0, 5, 2), dtype=np.int16), np.zeros(0, dtype=np.int32), 2))


def test_all_missing_individuals_and_loci():
    K = 3
    d = make_dataset(N=70, L=40, K=K, A=4, miss=0.05, seed=41)
    x = d.x.copy()
    x[:, 5, :] = -9                     # an individual with no data at all
    x[:, 69, :] = -9                    # ... the last one too (partly filled warp)
    x[7, :, :] = -9                     # a locus nobody was typed at
    x[39, :, 0] = -9                    # last locus: one copy missing everywhere drops the whole genotype (data_interface.c:828)
    s = Sampler(SeqData(x, d.allelenum, K), seed=4)
    s.chain_init(0, initd=[0.2, 0.5, 0.8])
    s.sweep(4)
    o = Oracle(x, d.allelenum, K)
    o.z[...] = s.get(_lib.STATE_Z)
    assert np.array_equal(s.get(_lib.STATE_TALLY), o.tally())
    cnt = s.get(_lib.STATE_CNT)
    assert np.array_equal(cnt.astype(np.float64), o.count_z())
    assert cnt[5].sum() == 0 and cnt[69].sum() == 0
    t = s.get(_lib.STATE_TALLY)
    assert t[:, 7, :].sum() == 0 and t[:, 39, :].sum() == 0
    lk = s.get(_lib.STATE_INDVLKH)
    assert lk[5] == 0.0 and lk[69] == 0.0            # log-likelihood of no data
    o.freq[...] = s.get(_lib.STATE_P)
    g = s.get(_lib.STATE_G)
    want = np.array([o.log_ld_indv(g[i], i) for i in range(o.N)])
    assert np.max(np.abs(lk - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    s.close()


def test_monomorphic_loci_are_skipped():
    """allelenum[l] == 1: the reference skips the locus everywhere (mcmc.c:817,1137,1737)."""
    K = 2
    d = make_dataset(N=50, L=16, K=K, A=3, miss=0.0, seed=42)
    x, an = d.x.copy(), d.allelenum.copy()
    x[3, :, :] = 0; an[3] = 1
    x[15, :, :] = 0; an[15] = 1
    s = Sampler(SeqData(x, an, K), seed=5)
    s.chain_init(0, initd=[0.3, 0.6])
    s.sweep(3)
    o = Oracle(x, an, K)
    o.z[...] = s.get(_lib.STATE_Z)
    t = s.get(_lib.STATE_TALLY)
    assert np.array_equal(t, o.tally()) and t[:, 3, :].sum() == 0 and t[:, 15, :].sum() == 0
    assert np.array_equal(s.get(_lib.STATE_CNT).astype(np.float64), o.count_z())
    o.freq[...] = s.get(_lib.STATE_P)
    g = s.get(_lib.STATE_G)
    want = np.array([o.log_ld_indv(g[i], i) for i in range(o.N)])
    lk = s.get(_lib.STATE_INDVLKH)
    assert np.max(np.abs(lk - want) / np.maximum(np.abs(want), 1.0)) <= RTOL
    s.close()


def test_limits_are_reported():
    d = make_dataset(N=20, L=5, K=2, A=2, miss=0.0, seed=43)
    for bad in (dict(popnum=17), dict(popnum=0)):
        with pytest.raises(instruct_b200.InstructError):
            Sampler(SeqData(d.x, d.allelenum, bad["popnum"]))
    with pytest.raises(instruct_b200.InstructError):
        Sampler(SeqData(d.x, d.allelenum, 2, mode=5, prior_flag=1))      # mode 5 with the DP prior is not built
    with pytest.raises(instruct_b200.InstructError):
        Sampler(SeqData(d.x, d.allelenum, 2, mode=6))
