"""Worker for the multi-process tests (launched by torch.distributed.run).

mode "cpu"  (gloo, no GPU): the host-side sharding logic -- shard bounds tile the individuals,
             the per-shard integer tallies all-reduce to the full tally (the exchange the
             library issues over NCCL per sweep), traces gather, the id broadcast works.
mode "groups" (gloo, world 4): chains x shards composed -- two groups of two ranks, one communicator id
             and one tally reduction per group.
mode "gpu"  (nccl, >= 2 GPUs): a chain whose individuals are sharded over the ranks is
             BIT-IDENTICAL to the same chain on one GPU (counter-based RNG keyed on global
             indices + integer all-reduce + fixed-order reductions).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from instruct_b200.shard import (broadcast_unique_id, chains_of_rank, gather_traces, group_layout,  # noqa: E402
                                 make_groups, shard_bounds, shard_genotypes)
from instruct_b200.synth import make_dataset  # noqa: E402


def cpu_mode():
    from oracle.pyoracle import Oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    K = 3
    d = make_dataset(N=101, L=19, K=K, A=4, miss=0.05, seed=3)
    o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(0)
    o.z[...] = rng.integers(0, K, size=o.z.shape)
    b, e = shard_bounds(d.N, world, rank)
    spans = [None] * world
    dist.all_gather_object(spans, (b, e))
    assert spans[0][0] == 0 and spans[-1][1] == d.N and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert shard_genotypes(d.x, world, rank).shape[1] == e - b
    part = torch.from_numpy(o.tally(b, e).astype(np.int32))
    dist.all_reduce(part)
    assert np.array_equal(part.numpy(), o.tally())
    uid = broadcast_unique_id(lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    mine = {c: np.arange(5) + 10.0 * c for c in chains_of_rank(5, world, rank)}
    tr = gather_traces(mine, 5, 5)
    assert np.array_equal(tr, np.arange(5)[None, :] + 10.0 * np.arange(5)[:, None])
    dist.barrier()
    if rank == 0:
        print("CPU_SHARD_OK")
    dist.destroy_process_group()


def groups_mode():
    """world 4 = 2 chains x 2 ranks each (gloo): every group gets ITS OWN communicator id and reduces ITS OWN
    chain's tallies; nothing crosses groups."""
    from oracle.pyoracle import Oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    G = 2
    chain, srank, ranks = group_layout(world, rank, G)
    assert ranks == [chain * G, chain * G + 1] and srank == rank - chain * G
    grp = make_groups(world, rank, G)
    uid = broadcast_unique_id(lambda: bytes([chain]) * 128, rank, src=ranks[0], group=grp)
    assert uid == bytes([chain]) * 128
    K = 3
    d = make_dataset(N=61, L=11, K=K, A=3, miss=0.05, seed=10 + chain)       # one data set per chain
    o = Oracle(d.x, d.allelenum, K)
    o.z[...] = np.random.default_rng(chain).integers(0, K, size=o.z.shape)
    b, e = shard_bounds(d.N, G, srank)
    part = torch.from_numpy(o.tally(b, e).astype(np.int32))
    dist.all_reduce(part, group=grp)
    assert np.array_equal(part.numpy(), o.tally())
    try:
        group_layout(6, 0, 4)
        raise AssertionError("group size must divide the world")
    except ValueError:
        pass
    assert make_groups(2, 0, 2) is None if world == 2 else True
    dist.barrier()
    if rank == 0:
        print("CPU_GROUPS_OK")
    dist.destroy_process_group()


def gpu_mode():
    from instruct_b200 import Sampler, SeqData, _lib
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = 4
    d = make_dataset(N=777, L=150, K=K, A=3, miss=0.04, seed=8)
    kw = dict(update=25, burnin=10, thinning=3, ckrep=4, seed=99)
    xs = shard_genotypes(d.x, world, rank)
    s = Sampler(SeqData(xs, d.allelenum, K), device=local, shard_rank=rank, shard_count=world, totalsize=d.N, **kw)
    s.comm_init(broadcast_unique_id(Sampler.unique_id, rank))
    ch, cv = s.run_chain(0, initd=[0.2, 0.4, 0.6, 0.8])
    zloc = s.get(_lib.STATE_Z)
    # the records live on their own rank during the sweeps (local scalar updates): a state hook gathers them, and the
    # sweeps after it must not use anything computed ahead of the hook
    qall = s.get(_lib.STATE_Q); gall = s.get(_lib.STATE_G)
    s.sweep(6)
    qall2 = s.get(_lib.STATE_Q); sall2 = s.get(_lib.STATE_S)
    s.close()
    ok = True
    if rank == 0:
        s1 = Sampler(SeqData(d.x, d.allelenum, K), device=local, **kw)
        c1, cv1 = s1.run_chain(0, initd=[0.2, 0.4, 0.6, 0.8])
        z1 = s1.get(_lib.STATE_Z)
        q1 = s1.get(_lib.STATE_Q); g1 = s1.get(_lib.STATE_G)
        s1.sweep(6)
        q12 = s1.get(_lib.STATE_Q); s12 = s1.get(_lib.STATE_S)
        s1.close()
        b, e = shard_bounds(d.N, world, 0)
        for name in ("qq", "qq2", "self_rates", "self_rates2", "gen", "gen2", "indvlkh"):
            ok &= bool(np.array_equal(getattr(ch, name), getattr(c1, name)))
        ok &= ch.totallkh == c1.totallkh and ch.totallkh2 == c1.totallkh2 and bool(np.array_equal(cv, cv1))
        ok &= bool(np.array_equal(zloc, z1[:, b:e, :]))
        ok &= bool(np.array_equal(qall, q1)) and bool(np.array_equal(gall, g1))
        ok &= bool(np.array_equal(qall2, q12)) and bool(np.array_equal(sall2, s12))
    ok &= _diploid_variant(rank, world, local, mode=2, back_refl=0, K=3)      # adaptive-independence proposals (-e 0): states double-buffered
    ok &= _diploid_variant(rank, world, local, mode=1, back_refl=1, K=5)      # no selfing: post-sweep sums only
    ok &= _diploid_variant(rank, world, local, mode=3, back_refl=1, K=3)      # individual rates: the all-gather path
    os.environ["IG_NCCL_SCALARS"] = "1"                                       # the same sums through ncclAllReduce instead of peer memory
    ok &= _diploid_variant(rank, world, local, mode=2, back_refl=1, K=4)
    del os.environ["IG_NCCL_SCALARS"]
    os.environ["IG_P_NCCL"] = "1"                                             # tally reduce-scattered / P all-gathered by NCCL instead of the peer kernels
    ok &= _diploid_variant(rank, world, local, mode=2, back_refl=1, K=4)
    del os.environ["IG_P_NCCL"]
    ok &= _tetra_sharded(rank, world, local, 1)
    ok &= _tetra_sharded(rank, world, local, 0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("GPU_SHARD_OK" if int(flag.item()) == 1 else "GPU_SHARD_MISMATCH")
    dist.destroy_process_group()


def _diploid_variant(rank, world, local, mode, back_refl, K):
    from instruct_b200 import Sampler, SeqData
    d = make_dataset(N=401, L=90, K=min(K, 4), A=3, miss=0.03, seed=10 + mode)
    kw = dict(update=24, burnin=9, thinning=2, ckrep=4, seed=5)
    initd = np.linspace(0.2, 0.8, K) if mode == 2 else None
    xs = shard_genotypes(d.x, world, rank)
    s = Sampler(SeqData(xs, d.allelenum, K, mode=mode, back_refl=back_refl), device=local, shard_rank=rank, shard_count=world, totalsize=d.N, **kw)
    s.comm_init(broadcast_unique_id(Sampler.unique_id, rank))
    ch, cv = s.run_chain(0, initd=initd)
    s.close()
    ok = True
    if rank == 0:
        s1 = Sampler(SeqData(d.x, d.allelenum, K, mode=mode, back_refl=back_refl), device=local, **kw)
        c1, cv1 = s1.run_chain(0, initd=initd)
        s1.close()
        for name in ("qq", "qq2", "self_rates", "self_rates2", "gen", "gen2", "indvlkh"):
            ok &= bool(np.array_equal(getattr(ch, name), getattr(c1, name)))
        ok &= ch.totallkh == c1.totallkh and ch.totallkh2 == c1.totallkh2 and bool(np.array_equal(cv, cv1))
        if not ok:
            print(f"variant mode={mode} back_refl={back_refl} differs", flush=True)
    return ok


def _tetra_sharded(rank, world, local, autopoly):
    """Auto- and allotetraploid: the sharded chain (int32 tally all-reduce -- two tallies for the
    allotetraploid model --, per-individual S statistics and records all-gathered, fixed-order
    reductions) is bit-identical to the one-GPU chain."""
    from instruct_b200 import Sampler, SeqData, _lib
    from instruct_b200.synth import make_tetra_dataset
    K = 3
    d = make_tetra_dataset(N=301, L=45, K=K, A=4, miss=0.04, seed=9)
    kw = dict(update=20, burnin=8, thinning=3, ckrep=4, seed=77)
    xs = shard_genotypes(d.x, world, rank)
    s = Sampler(SeqData(xs, d.allelenum, K, ploid=4, autopoly=autopoly), device=local, shard_rank=rank, shard_count=world, totalsize=d.N, **kw)
    s.comm_init(broadcast_unique_id(Sampler.unique_id, rank))
    ch, cv = s.run_chain(0, initd=[0.3, 0.5, 0.7])
    zloc, gloc = s.get(_lib.STATE_Z), s.get(_lib.STATE_GENO)
    s.close()
    ok = True
    if rank == 0:
        s1 = Sampler(SeqData(d.x, d.allelenum, K, ploid=4, autopoly=autopoly), device=local, **kw)
        c1, cv1 = s1.run_chain(0, initd=[0.3, 0.5, 0.7])
        z1, g1 = s1.get(_lib.STATE_Z), s1.get(_lib.STATE_GENO)
        s1.close()
        b, e = shard_bounds(d.N, world, 0)
        for name in ("qq", "qq2", "self_rates", "self_rates2", "indvlkh"):
            ok &= bool(np.array_equal(getattr(ch, name), getattr(c1, name)))
        ok &= ch.totallkh == c1.totallkh and ch.totallkh2 == c1.totallkh2 and bool(np.array_equal(cv, cv1))
        ok &= bool(np.array_equal(zloc, z1[:, b:e, :])) and bool(np.array_equal(gloc, g1[:, b:e, :]))
        if not ok:
            print("TETRA_SHARD_MISMATCH", ch.totallkh, c1.totallkh, ch.self_rates, c1.self_rates)
    return ok


if __name__ == "__main__":
    {"cpu": cpu_mode, "groups": groups_mode, "gpu": gpu_mode}[sys.argv[1]]()
