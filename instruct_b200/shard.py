"""Host-side partitioning for the two ways the sweep shards:

* independent chains (``-c``): chain ``c`` runs on rank ``c % world``; no per-sweep
  communication, the ``ckrep`` log-likelihood values per chain are gathered at the end for
  Gelman-Rubin (check_converg.c:44-91);
* individuals of ONE chain: contiguous, equal-capacity blocks of individuals per rank (the
  individual-minor axis of the [L][N][ploid] store splits cleanly); per sweep one int32
  all-reduce of n[L][A][K] and one all-gather of the per-individual records, both issued by
  the library over NCCL (ig_comm_init) -- torch.distributed only carries the rendezvous.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(N: int, world: int, rank: int):
    """[begin, end) of rank's individuals: equal capacity ceil(N/world), last shard short.
    Must agree with ig_create() (ig_api.cu) -- the all-gather needs equal capacities."""
    cap = -(-N // world)
    b = rank * cap
    e = min(N, b + cap)
    if e <= b:
        raise ValueError(f"shard {rank} of {world} is empty for N={N}")
    return b, e


def shard_genotypes(x: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Slice a packed [L][N][ploid] store along the individual axis."""
    b, e = shard_bounds(x.shape[1], world, rank)
    return np.ascontiguousarray(x[:, b:e, :])


def chains_of_rank(chainnum: int, world: int, rank: int):
    return [c for c in range(chainnum) if c % world == rank]


def group_layout(world: int, rank: int, group_size: int):
    """Chains x shards composed (SURVEY.md section 8e, configs[4] "4 chains x 2 GPUs each"): ``world`` ranks
    form world / group_size groups of consecutive ranks; each group runs ONE chain whose individuals are
    sharded over the group's ranks.  Returns (chain index, shard rank inside the group, the group's ranks)."""
    if group_size < 1 or world % group_size:
        raise ValueError(f"group size {group_size} does not divide {world} ranks")
    c = rank // group_size
    return c, rank % group_size, list(range(c * group_size, (c + 1) * group_size))


def make_groups(world: int, rank: int, group_size: int):
    """torch.distributed sub-groups for group_layout(); every rank creates every group, in the same order
    (a requirement of new_group).  Returns this rank's group (None when one group spans all ranks)."""
    import torch.distributed as dist

    if group_size == world:
        return None
    mine = None
    for c in range(world // group_size):
        g = dist.new_group(list(range(c * group_size, (c + 1) * group_size)))
        if c == rank // group_size:
            mine = g
    return mine


def broadcast_unique_id(make_id, rank: int, src: int = 0, group=None) -> bytes:
    """Rank ``src`` creates the 128-byte NCCL unique id, everybody receives it through
    torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(128, dtype=torch.uint8)
    if rank == src:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def gather_traces(local: dict, chainnum: int, ckrep: int, group=None) -> np.ndarray:
    """All ranks contribute {chain_id: trace[ckrep]}; returns [chainnum][ckrep] on every rank."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(chainnum, ckrep, dtype=torch.float64)
    for c, tr in local.items():
        buf[c] = torch.as_tensor(np.asarray(tr, dtype=np.float64))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            buf = buf.cuda()
        dist.all_reduce(buf, group=group)
        buf = buf.cpu()
    return buf.numpy()
