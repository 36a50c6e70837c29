"""Local scalar updates of a sharded chain (ig_kernels.cu spop_tree_kernel / spop_decide_kernel / post_local_kernel;
DESIGN.md section 8): update_S_POP (mcmc.c:913-983) taken from the 2^K subset sums of proposal() (mcmc.c:1630) over the
LOCAL individuals + one int64 all-reduce, update_alpha / cal_lkh totals (mcmc.c:1244-1263, :1940) the same way, moments
of the local individuals gathered at the chain's end.

On one GPU the whole sharded code path runs over a one-rank NCCL communicator (IG_COMM_SINGLE=1) and has to reproduce
the plain one-GPU chain bit for bit -- the same claim tests/test_shard_multi.py makes across two real GPUs."""
import os

import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset

pytestmark = pytest.mark.gpu

FIELDS = ("qq", "qq2", "self_rates", "self_rates2", "gen", "gen2", "indvlkh")


def _chain(d, K, mode, back_refl, env, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        s = Sampler(SeqData(d.x, d.allelenum, K, mode=mode, back_refl=back_refl), **kw)
        if env.get("IG_COMM_SINGLE"):
            s.comm_init(Sampler.unique_id())
        ch, cv = s.run_chain(0, initd=np.linspace(0.2, 0.8, K) if mode == 2 else None)
        st = {"S": s.get(_lib.STATE_S) if mode == 2 else None, "Q": s.get(_lib.STATE_Q), "alpha": s.get(_lib.STATE_ALPHA),
              "Z": s.get(_lib.STATE_Z), "lkh": s.get(_lib.STATE_TOTALLKH)}
        # a few more sweeps after the hooks (the gathered records and the subset sums taken ahead must have been invalidated)
        s.sweep(7)
        st["Q2"] = s.get(_lib.STATE_Q)
        st["G2"] = s.get(_lib.STATE_G) if mode == 2 else None
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return ch, cv, st


@pytest.mark.parametrize("mode,back_refl,K,N,L", [(2, 1, 4, 777, 150), (2, 0, 3, 300, 120), (1, 1, 5, 500, 90), (2, 1, 8, 1300, 64)])
def test_local_scalar_updates_reproduce_the_plain_chain(mode, back_refl, K, N, L):
    d = make_dataset(N=N, L=L, K=min(K, 4), A=3, miss=0.04, seed=8)
    kw = dict(update=40, burnin=15, thinning=3, ckrep=5, seed=99)
    ref = _chain(d, K, mode, back_refl, {}, **kw)
    # default: sums all-reduced by the peer-memory kernel; then through ncclAllReduce; without taking the next sweep's subset
    # sums ahead; and the round-1 path (records all-gathered, sequential update_S_POP on every rank)
    for env in ({"IG_COMM_SINGLE": "1"}, {"IG_COMM_SINGLE": "1", "IG_NCCL_SCALARS": "1"}, {"IG_COMM_SINGLE": "1", "IG_NO_TREE_AHEAD": "1"},
                {"IG_COMM_SINGLE": "1", "IG_GATHER_RECORDS": "1"},
                # the tally -> P exchange: default over peer memory; through ncclReduceScatter / ncclAllGather; through ncclAllReduce
                {"IG_COMM_SINGLE": "1", "IG_P_NCCL": "1"}, {"IG_COMM_SINGLE": "1", "IG_P_NCCL": "1", "IG_P_ALLREDUCE": "1"}):
        got = _chain(d, K, mode, back_refl, env, **kw)
        for name in FIELDS:
            assert np.array_equal(getattr(got[0], name), getattr(ref[0], name)), (env, name)
        assert got[0].totallkh == ref[0].totallkh and got[0].totallkh2 == ref[0].totallkh2, env
        assert np.array_equal(got[1], ref[1]), env
        for k, v in ref[2].items():
            if v is not None:
                assert np.array_equal(got[2][k], v), (env, k)


def test_fixed_point_totals_match_the_records():
    """totallkh, the column sums of Q and sum log q are fixed-point sums (2^-24, 2^-40, 2^-28 per term): against the
    double-precision sums of the same records they differ by rounding only."""
    d = make_dataset(N=600, L=200, K=3, A=4, miss=0.02, seed=5)
    s = Sampler(SeqData(d.x, d.allelenum, 3, mode=2), seed=4)
    s.chain_init(0, initd=[0.3, 0.5, 0.7])
    s.sweep(12)
    q = s.get(_lib.STATE_Q)
    lk = s.get(_lib.STATE_INDVLKH)
    tot = float(s.get(_lib.STATE_TOTALLKH)[0])
    assert abs(tot - lk.sum()) <= 600 * 2.0 ** -24
    s.close()
