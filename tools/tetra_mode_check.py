"""Is the high-likelihood state a real mode?  Run the chain that lands there, then evaluate the
oracle's likelihood on the identical state (tools, not a test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from instruct_b200 import Sampler, SeqData, _lib
from oracle.pytetra import TetraOracle
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/posterior_tetra.npz'))
K = int(g["K"]); x = g["x"]; an = g["allelenum"]
nd = (x >= 0).sum(axis=2).astype(np.int32)
sd = SeqData(x, an, K, ploid=4, autopoly=1)
for rep, seed in ((9, 2009), (9, 3009)):
    s = Sampler(sd, seed=seed)
    s.chain_init(rep, initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
    s.set(_lib.STATE_ALPHA, [float(g["alpha"][rep])])
    for chunk in range(6):
        s.sweep(250)
        o = TetraOracle(x, nd, an, K)
        o.z[...] = s.get(_lib.STATE_Z); o.geno[...] = s.get(_lib.STATE_GENO)
        o.qq[...] = s.get(_lib.STATE_Q); o.freq[...] = s.get(_lib.STATE_P); o.self_rates[...] = s.get(_lib.STATE_S)
        o.tables()
        tot = o.cal_lkd()
        got = float(s.get(_lib.STATE_TOTALLKH)[0])
        z = s.get(_lib.STATE_Z); gn = s.get(_lib.STATE_GENO)
        same = (z == z[:, :, :1]).all(axis=2)
        hom = (gn == gn[:, :, :1]).all(axis=2)
        print(f"seed {seed} sweep {250 * (chunk + 1)}: gpu {got:.1f} oracle {tot:.1f} S {np.round(s.get(_lib.STATE_S), 3)} same-z frac {same.mean():.3f} homoz geno frac {hom.mean():.3f} tally ok {np.array_equal(s.get(_lib.STATE_TALLY), o.tally())} qcol {np.round(s.get(_lib.STATE_Q).sum(axis=0), 1)}")
    s.close()
