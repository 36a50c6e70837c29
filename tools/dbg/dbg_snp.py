import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset
from oracle.pyoracle import Oracle
import test_gpu_parity as T
for (N,L,K,A,miss) in [(70,1300,8,2,0.05),(70,1024,8,2,0.0),(70,600,8,2,0.0),(16,1024,8,2,0.0)]:
    d, sd = T._mk(N, L, K, A, miss, seed=3)
    s = Sampler(sd); o = Oracle(d.x, d.allelenum, K)
    rng = np.random.default_rng(5); T._inject(s, o, rng)
    g_old = o.gen.copy(); gprop = rng.integers(1, 12, size=o.N).astype(np.int32)
    s.set(_lib.STATE_GPROP, gprop)
    s.run_phase(_lib.PHASE_ZQ)
    parts = s.get(_lib.STATE_LLPARTS)
    ll_old_g = np.array([o.log_ld_indv(g_old[i], i) for i in range(o.N)])
    ll_old_p = np.array([o.log_ld_indv(gprop[i], i) for i in range(o.N)])
    err = np.abs(parts[:,0] - (ll_old_p - ll_old_g))
    bad = np.where(~(err < 1e-3))[0]
    print((N,L,K,A,miss), s.geometry(), "bad", len(bad), bad[:40].tolist())
    print("  gold", g_old[:20].tolist(), "gprop", gprop[:20].tolist())
    print("  err", np.round(err[:20],3).tolist())
    s.close()
