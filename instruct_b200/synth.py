"""Synthetic genotype generator for the InStruct hot path.

The reference ships no example data, so every test and benchmark
input comes from here.  The generative model is the one InStruct assumes: K clusters with
allele frequencies ``P[k][l] ~ Dirichlet(1_A)``, individual ``i`` drawn mostly from cluster
``i mod K``, a per-cluster selfing rate, a geometric number of selfing generations per
individual, and per locus two allele copies that are forced identical with probability
``1 - 2**-(G-1)``.  Missing genotypes are i.i.d. Bernoulli.

Two products:
  * :func:`make_dataset`   -- dense allele indices in the NEW packed layout
    ``int16[L][N][ploid]`` (negative = missing), optionally on a CUDA device so that
    config-4 sized inputs (4 GB) never exist as Python lists or text;
  * :func:`write_reference_text` -- the same data in the reference's text format
    (two lines per diploid individual, ``label pop a_1 .. a_L``; data_interface.c:133-245)
    for the compiled reference program.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MISSING = -9  # the reference's in-memory code for a missing allele (data_interface.c:494)


@dataclass
class SynthData:
    x: np.ndarray            # int16 [L][N][ploid]
    allelenum: np.ndarray    # int32 [L]
    K: int
    S_true: np.ndarray       # [K] or [N]
    G_true: np.ndarray       # [N]
    Q_true: np.ndarray       # [N][K]
    pop: np.ndarray          # [N] home cluster
    P_true: np.ndarray       # [K][L][A]

    @property
    def L(self):
        return self.x.shape[0]

    @property
    def N(self):
        return self.x.shape[1]

    @property
    def ploid(self):
        return self.x.shape[2]


def make_dataset(N, L, K, A=2, miss=0.0, seed=0, pure=False, own=0.9, s_atoms=None,
                 ploid=2) -> SynthData:
    """Diploid synthetic data set (numpy, for tests and small benchmarks)."""
    rng = np.random.default_rng(seed)
    P = rng.dirichlet(np.ones(A), size=(K, L))                      # [K][L][A]
    pop = np.arange(N) % K
    if pure or K == 1:
        Q = np.eye(K)[pop]
    else:
        Q = np.full((N, K), (1.0 - own) / (K - 1))
        Q[np.arange(N), pop] = own
    if s_atoms is None:
        S_k = np.linspace(0.1, 0.9, K) if K > 1 else np.array([0.5])
        s_i = S_k[pop]
        S_true = S_k
    else:                                                          # individual selfing rates (config 3)
        s_i = rng.choice(np.asarray(s_atoms, dtype=float), size=N)
        S_true = s_i
    G = np.minimum(rng.geometric(1.0 - s_i), 50)                    # generations since outcrossing
    x = np.empty((L, N, ploid), dtype=np.int16)
    cumP = np.cumsum(P, axis=2)                                     # [K][L][A]
    cumQ = np.cumsum(Q, axis=1)
    for l in range(L):
        u = rng.random((N, ploid))
        anc = (rng.random((N, ploid))[:, :, None] > cumQ[:, None, :]).sum(axis=2)   # ancestry per copy
        anc = np.minimum(anc, K - 1)
        cp = cumP[anc, l, :]                                        # [N][ploid][A]
        a = (u[:, :, None] > cp).sum(axis=2)
        a = np.minimum(a, A - 1)
        homo = rng.random(N) < (1.0 - 0.5 ** (G - 1))
        a[homo, 1:] = a[homo, :1]
        x[l] = a.astype(np.int16)
    if miss > 0:
        m = rng.random((L, N)) < miss
        x[m] = MISSING
    x, allelenum = recode_dense(x)
    return SynthData(x=x, allelenum=allelenum, K=K, S_true=S_true, G_true=G, Q_true=Q, pop=pop, P_true=P)


def recode_dense(x: np.ndarray):
    """Renumber alleles per locus in order of first appearance scanning individuals then
    copies, and drop monomorphic loci -- the rule of transform_data (data_interface.c:510-548).
    Returns (x', allelenum)."""
    L, N, ploid = x.shape
    keep, nums = [], []
    out = np.empty_like(x)
    for l in range(L):
        flat = x[l].reshape(-1)
        valid = flat >= 0
        vals = flat[valid]
        if vals.size == 0:
            continue
        uniq, first = np.unique(vals, return_index=True)
        if uniq.size < 2:
            continue
        order = uniq[np.argsort(first)]
        lut = np.full(int(flat.max()) + 1, -1, dtype=np.int16)
        lut[order] = np.arange(order.size, dtype=np.int16)
        row = np.full(flat.shape, MISSING, dtype=np.int16)
        row[valid] = lut[vals]
        out[len(keep)] = row.reshape(N, ploid)
        keep.append(l)
        nums.append(order.size)
    return np.ascontiguousarray(out[: len(keep)]), np.asarray(nums, dtype=np.int32)


def make_dataset_torch(N, L, K, A=2, miss=0.0, seed=0, device="cuda", own=0.9, chunk=2048, i0=0, n_local=None):
    """Config-4 scale generator: same model, torch ops on ``device``, written straight into
    an int16 [L][n_local][2] tensor in locus chunks.  ``i0``/``n_local`` select a block of
    individuals of the SAME global data set (allele frequencies come from a generator seeded
    by ``seed`` alone, individual-level randomness from ``seed`` and ``i0``), so the shards
    that several ranks generate fit together.  Returns (x, allelenum) as torch tensors.
    Loci are assumed polymorphic (true with overwhelming probability for N >= 1000)."""
    import torch

    n_local = N - i0 if n_local is None else n_local
    gP = torch.Generator(device=device)
    gP.manual_seed(seed)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + 17 + i0)
    idx = torch.arange(i0, i0 + n_local, device=device)
    pop = idx % K
    if K > 1:
        Q = torch.full((n_local, K), (1.0 - own) / (K - 1), device=device)
        Q[torch.arange(n_local, device=device), pop] = own
    else:
        Q = torch.ones((n_local, 1), device=device)
    cumQ = torch.cumsum(Q, 1)
    S_k = torch.linspace(0.1, 0.9, K, device=device) if K > 1 else torch.tensor([0.5], device=device)
    s_i = S_k[pop]
    u = torch.rand(n_local, generator=g, device=device).clamp_min(1e-12)
    G = torch.clamp(torch.floor(torch.log(u) / torch.log(s_i)) + 1, 1, 50)
    p_homo = 1.0 - torch.pow(0.5, G - 1)
    x = torch.empty((L, n_local, 2), dtype=torch.int16, device=device)
    for l0 in range(0, L, chunk):
        l1 = min(L, l0 + chunk)
        n = l1 - l0
        e = -torch.log(torch.rand((K, n, A), generator=gP, device=device).clamp_min(1e-12))
        cumP = torch.cumsum(e / e.sum(2, keepdim=True), 2)            # Dirichlet(1_A) [K][n][A]
        anc = (torch.rand((n, n_local, 2, 1), generator=g, device=device) > cumQ[None, :, None, :]).sum(3)
        anc.clamp_(max=K - 1)
        li = torch.arange(n, device=device)[:, None, None].expand(n, n_local, 2)
        cp = cumP[anc, li]                                            # [n][n_local][2][A]
        a = (torch.rand((n, n_local, 2, 1), generator=g, device=device) > cp).sum(3).clamp_(max=A - 1)
        homo = torch.rand((n, n_local), generator=g, device=device) < p_homo[None, :]
        a[:, :, 1] = torch.where(homo, a[:, :, 0], a[:, :, 1])
        if miss > 0:
            mm = torch.rand((n, n_local), generator=g, device=device) < miss
            a[mm] = MISSING
        x[l0:l1] = a.to(torch.int16)
    allelenum = torch.full((L,), A, dtype=torch.int32, device=device)
    return x, allelenum


def write_reference_text(path, x: np.ndarray, labels=True, popdata=True, pop=None,
                         allele_base=100, missing="-9"):
    """Write a diploid data set in the reference's default format (-af 0): one line per
    haploid copy, ``[label] [pop] a_1 ... a_L`` (read_data_from_file, data_interface.c:133)."""
    L, N, ploid = x.shape
    with open(path, "w") as fh:
        for i in range(N):
            for c in range(ploid):
                tok = []
                if labels:
                    tok.append(f"ind{i}")
                if popdata:
                    tok.append(str(int(pop[i]) if pop is not None else 0))
                col = x[:, i, c]
                tok.extend(missing if v < 0 else str(allele_base + 2 * int(v)) for v in col)
                fh.write(" ".join(tok) + "\n")


# ---------------------------------------------------------------------------------------
# autotetraploid data (config 5): the reader keeps, per individual and locus, the SET of
# distinct alleles observed (ascending dense codes) and how many there are ("alleleid"); the
# dosage is hidden (transform_data2, data_interface.c:571-669).
# ---------------------------------------------------------------------------------------
@dataclass
class TetraData:
    x: np.ndarray            # int16 [L][N][4] distinct alleles, ascending, -1 padding (all -1 when missing)
    nd: np.ndarray           # uint8 [L][N] number of distinct alleles, 0 = missing
    allelenum: np.ndarray    # int32 [L]
    K: int
    S_true: np.ndarray
    Q_true: np.ndarray
    pop: np.ndarray
    dosage: np.ndarray       # int16 [L][N][4] the four hidden copies (for write_reference_text_tetra)

    @property
    def L(self):
        return self.x.shape[0]

    @property
    def N(self):
        return self.x.shape[1]


def observed_sets(copies: np.ndarray):
    """[L][N][4] allele copies (negative = missing genotype) -> (x, nd) as the reader stores them."""
    L, N, _ = copies.shape
    srt = np.sort(copies, axis=2)
    first = np.ones_like(srt, dtype=bool)
    first[:, :, 1:] = srt[:, :, 1:] != srt[:, :, :-1]
    missing = (copies < 0).any(axis=2)
    nd = first.sum(axis=2).astype(np.uint8)
    # stable partition: distinct values first, in ascending order
    key = np.where(first, srt, np.int16(32767))
    x = np.sort(key, axis=2).astype(np.int16)
    x[x == 32767] = -1
    x[missing] = -1
    nd[missing] = 0
    return np.ascontiguousarray(x), np.ascontiguousarray(nd)


def make_tetra_dataset(N, L, K, A=4, miss=0.0, seed=0, own=0.9) -> TetraData:
    """Autotetraploid synthetic data: four copies per locus from the individual's admixed
    frequencies; with probability s (the home cluster's selfing rate) a genotype is replaced
    by two copies of each of two of its own alleles, a crude stand-in for the excess
    homozygosity selfing causes (the sampler is tested on conditionals, not on recovering s)."""
    rng = np.random.default_rng(seed)
    P = rng.dirichlet(np.ones(A), size=(K, L))
    pop = np.arange(N) % K
    Q = np.full((N, K), (1.0 - own) / max(K - 1, 1))
    Q[np.arange(N), pop] = own if K > 1 else 1.0
    S_k = np.linspace(0.1, 0.9, K) if K > 1 else np.array([0.5])
    cumP = np.cumsum(P, axis=2)
    cumQ = np.cumsum(Q, axis=1)
    copies = np.empty((L, N, 4), dtype=np.int16)
    for l in range(L):
        anc = np.minimum((rng.random((N, 4))[:, :, None] > cumQ[:, None, :]).sum(axis=2), K - 1)
        a = np.minimum((rng.random((N, 4))[:, :, None] > cumP[anc, l, :]).sum(axis=2), A - 1)
        selfed = rng.random(N) < S_k[pop]
        a[selfed, 2] = a[selfed, 0]
        a[selfed, 3] = a[selfed, 1]
        copies[l] = a
    if miss > 0:
        copies[rng.random((L, N)) < miss] = MISSING
    copies, allelenum = recode_dense(copies)
    x, nd = observed_sets(copies)
    return TetraData(x=x, nd=nd, allelenum=allelenum, K=K, S_true=S_k, Q_true=Q, pop=pop, dosage=copies)


def make_tetra_dataset_torch(N, L, K, A=4, miss=0.0, seed=0, device="cuda", own=0.9, chunk=1024):
    """Config-5 scale generator: the model of :func:`make_tetra_dataset` in torch ops on ``device``,
    written straight into an int16 [L][N][4] tensor of distinct-allele sets (ascending, -1 padding).
    Loci are assumed to show all A alleles (true with overwhelming probability for N >= 1000)."""
    import torch

    gP = torch.Generator(device=device)
    gP.manual_seed(seed)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + 29)
    pop = torch.arange(N, device=device) % K
    Q = torch.full((N, K), (1.0 - own) / max(K - 1, 1), device=device)
    Q[torch.arange(N, device=device), pop] = own if K > 1 else 1.0
    cumQ = torch.cumsum(Q, 1)
    S_k = torch.linspace(0.1, 0.9, K, device=device) if K > 1 else torch.tensor([0.5], device=device)
    s_i = S_k[pop]
    x = torch.empty((L, N, 4), dtype=torch.int16, device=device)
    for l0 in range(0, L, chunk):
        n = min(L, l0 + chunk) - l0
        e = -torch.log(torch.rand((K, n, A), generator=gP, device=device).clamp_min(1e-12))
        cumP = torch.cumsum(e / e.sum(2, keepdim=True), 2)
        anc = (torch.rand((n, N, 4, 1), generator=g, device=device) > cumQ[None, :, None, :]).sum(3).clamp_(max=K - 1)
        li = torch.arange(n, device=device)[:, None, None].expand(n, N, 4)
        a = (torch.rand((n, N, 4, 1), generator=g, device=device) > cumP[anc, li]).sum(3).clamp_(max=A - 1)
        selfed = torch.rand((n, N), generator=g, device=device) < s_i[None, :]
        a[:, :, 2] = torch.where(selfed, a[:, :, 0], a[:, :, 2])
        a[:, :, 3] = torch.where(selfed, a[:, :, 1], a[:, :, 3])
        srt, _ = torch.sort(a, dim=2)
        first = torch.ones_like(srt, dtype=torch.bool)
        first[:, :, 1:] = srt[:, :, 1:] != srt[:, :, :-1]
        key, _ = torch.sort(torch.where(first, srt, torch.full_like(srt, 32767)), dim=2)
        key[key == 32767] = -1
        if miss > 0:
            key[torch.rand((n, N), generator=g, device=device) < miss] = -1
        x[l0:l0 + n] = key.to(torch.int16)
    allelenum = torch.full((L,), A, dtype=torch.int32, device=device)
    return x, allelenum


def write_reference_text_tetra(path, copies: np.ndarray, labels=True, popdata=True, pop=None, allele_base=100, missing="-9"):
    """Write a tetraploid data set in the one-line-per-individual format the reference reads for
    ``-p 4`` (read_data_fmt2, data_interface.c:715): ``[label] [pop]`` then four tokens per locus."""
    L, N, _ = copies.shape
    with open(path, "w") as fh:
        for i in range(N):
            tok = []
            if labels:
                tok.append(f"ind{i}")
            if popdata:
                tok.append(str(int(pop[i]) if pop is not None else 0))
            for l in range(L):
                tok.extend(missing if v < 0 else str(allele_base + 2 * int(v)) for v in copies[l, i])
            fh.write(" ".join(tok) + "\n")
