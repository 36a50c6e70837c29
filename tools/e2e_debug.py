import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from instruct_b200 import Sampler, SeqData, _lib
from instruct_b200.synth import make_dataset_torch
N, L, K = 10_000, 100_000, 8
dev = torch.device("cuda", 0)
x, an = make_dataset_torch(N, L, K, A=2, seed=4, device=dev)
torch.cuda.synchronize()
shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(L, N, 2), strides=(0, 0, 0))
sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, mode=2, nstep_check_empty_cluster=10 ** 9)
for seed, initd in ((2024, None), (2024, np.linspace(0.2, 0.8, K)), (1, None)):
    s = Sampler(sd, seed=seed, update=55, burnin=5, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ch, cv = s.run_chain(0, initd=initd)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"seed {seed} initd {'given' if initd is not None else 'random'}: run_chain {dt:.1f} ms, step {ch.step}/{ch.steps}, totallkh {ch.totallkh:.1f}, S {np.round(ch.self_rates, 3)}, flag {ch.flag_empty_cluster}")
    ms = s.time_sweeps(10)
    print("   next 10 sweeps:", ms / 10, "ms each; alpha", s.get(_lib.STATE_ALPHA), "qcol", np.round(s.get(_lib.STATE_Q).sum(axis=0), 1))
    s.close()
