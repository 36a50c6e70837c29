"""Parity at BASELINE.json's full sizes (configs[3]: K=8, N=10,000, L=100,000 diploid), where the
CPU oracle cannot run the whole state: the tally is recomputed bit-exactly by an independent
implementation (torch.bincount over the Z and X the library holds), the size-independent
invariants of the sweep are checked (every usable copy is counted exactly once, per locus and
per individual; the checksum of the tally equals the checksum of the counts; totallkh is the
sum of indvlkh), and the oracle itself is run on slices (a few individuals over all loci, a few
loci over all individuals) of the identical state."""
import numpy as np
import pytest

from instruct_b200 import Sampler, SeqData, _lib
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def test_config4_full_size_invariants_and_sliced_oracle():
    import torch
    from instruct_b200.synth import make_dataset_torch
    N, L, K, A = 10_000, 100_000, 8, 2
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 40e9:
        pytest.skip("needs 40 GB of free device memory")
    x, an = make_dataset_torch(N, L, K, A=A, miss=0.02, seed=11, device=dev)
    torch.cuda.synchronize()
    shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(L, N, 2), strides=(0, 0, 0))
    sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, mode=2)
    s = Sampler(sd, seed=5, device=0, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
    s.chain_init(0, initd=np.linspace(0.2, 0.8, K))
    s.sweep(3)
    z = torch.from_numpy(s.get(_lib.STATE_Z)).to(dev)               # int8 [L][N][2]
    tally = torch.from_numpy(s.get(_lib.STATE_TALLY)).to(dev)       # int32 [K][L][A]
    cnt = torch.from_numpy(s.get(_lib.STATE_CNT)).to(dev)           # int32 [N][K]
    usable = ~(x < 0).any(dim=2)                                    # [L][N]
    # ---- the tally, recomputed independently and compared bit for bit, in locus blocks
    want_cnt = torch.zeros((N, K), dtype=torch.int64, device=dev)
    step = 5000
    for l0 in range(0, L, step):
        xs, zs, us = x[l0:l0 + step].long(), z[l0:l0 + step].long(), usable[l0:l0 + step]
        nl = xs.shape[0]
        m = us[:, :, None].expand(nl, N, 2)
        li = torch.arange(nl, device=dev)[:, None, None].expand(nl, N, 2)
        key = ((zs * nl + li) * A + xs.clamp_min(0))[m]
        want = torch.bincount(key, minlength=K * nl * A).reshape(K, nl, A)
        assert torch.equal(want.to(torch.int32), tally[:, l0:l0 + step, :]), f"tally differs in loci {l0}.."
        ii = torch.arange(N, device=dev)[None, :, None].expand(nl, N, 2)
        want_cnt += torch.bincount((ii * K + zs)[m], minlength=N * K).reshape(N, K)
    assert torch.equal(want_cnt.to(torch.int32), cnt)
    # ---- invariants: every usable copy counted once, per locus and per individual; checksum of checksums
    assert torch.equal(tally.sum(dim=(0, 2)).long(), 2 * usable.sum(dim=1))
    assert torch.equal(cnt.sum(dim=1).long(), 2 * usable.sum(dim=0))
    assert torch.equal(tally.sum(dim=(1, 2)).long(), cnt.sum(dim=0).long())
    lk = s.get(_lib.STATE_INDVLKH)
    tot = s.get(_lib.STATE_TOTALLKH)[0]
    assert abs(tot - lk.sum()) <= 1e-9 * abs(tot) + len(lk) * 2.0 ** -24     # totallkh is a 2^-24 fixed-point sum
    q = s.get(_lib.STATE_Q)
    np.testing.assert_allclose(q.sum(axis=1), 1.0, rtol=1e-12)
    # ---- the oracle on slices of the identical state
    P = s.get(_lib.STATE_P)                                          # [K][L][A]
    G = s.get(_lib.STATE_G)
    ids = np.array([0, 1, 4999, 9999])
    xs = x[:, ids, :].cpu().numpy()
    o = Oracle(xs, an.cpu().numpy(), K)
    o.z[...] = z[:, ids, :].cpu().numpy()
    o.freq[...] = P
    for j, i in enumerate(ids):
        want = o.log_ld_indv(int(G[i]), j)
        assert abs(lk[i] - want) <= 1e-6 * abs(want), (i, lk[i], want)
    ls = np.arange(0, L, L // 50)[:50]
    o2 = Oracle(x[ls].cpu().numpy(), an[ls].cpu().numpy(), K)
    o2.z[...] = z[ls].cpu().numpy()
    assert np.array_equal(o2.tally(), tally[:, ls, :].cpu().numpy())
    assert np.array_equal(o2.missing_mask(), (~usable[ls]).cpu().numpy().astype(np.uint8))
    s.close()


def test_config5_full_size_invariants():
    """configs[4] per chain (autotetraploid K=6, N=L=20,000): each usable genotype contributes its
    four copies to the tally and to its individual's ancestry counts exactly once."""
    import torch
    from instruct_b200.synth import make_tetra_dataset_torch
    N, L, K, A = 20_000, 20_000, 6, 4
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 40e9:
        pytest.skip("needs 40 GB of free device memory")
    x, an = make_tetra_dataset_torch(N, L, K, A=A, miss=0.02, seed=12, device=dev)
    torch.cuda.synchronize()
    usable = x[:, :, 0] >= 0                                         # [L][N]
    shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(L, N, 4), strides=(0, 0, 0))
    sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, ploid=4, mode=2)
    s = Sampler(sd, seed=6, device=0, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
    s.chain_init(0, initd=np.linspace(0.2, 0.8, K))
    s.sweep(2)
    tally = torch.from_numpy(s.get(_lib.STATE_TALLY)).to(dev)
    cnt = torch.from_numpy(s.get(_lib.STATE_CNT)).to(dev)
    assert torch.equal(tally.sum(dim=(0, 2)).long(), 4 * usable.sum(dim=1))
    assert torch.equal(cnt.sum(dim=1).long(), 4 * usable.sum(dim=0))
    assert torch.equal(tally.sum(dim=(1, 2)).long(), cnt.sum(dim=0).long())
    lk = s.get(_lib.STATE_INDVLKH)
    tot = s.get(_lib.STATE_TOTALLKH)[0]
    assert np.isfinite(tot) and abs(tot - lk.sum()) <= 1e-9 * abs(tot) + len(lk) * 2.0 ** -24     # totallkh is a 2^-24 fixed-point sum
    s.close()
