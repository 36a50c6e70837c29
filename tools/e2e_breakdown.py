"""Where the end-to-end time of one ig_mcmc_updating() call goes at config 4 (tools, not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from instruct_b200 import Sampler, SeqData, Init, mcmc_updating, _lib
from instruct_b200.synth import make_dataset_torch
N, L, K = 10_000, 100_000, 8
dev = torch.device("cuda", 0)
x, an = make_dataset_torch(N, L, K, A=2, seed=4, device=dev)
xh = torch.empty(x.shape, dtype=torch.int16, pin_memory=True); xh.copy_(x); torch.cuda.synchronize()
anh = an.cpu().numpy()
del x
def T(f, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
sd = SeqData(xh.numpy(), anh, K, mode=2, nstep_check_empty_cluster=10 ** 9)
for rep in range(2):
    s, t_load = T(Sampler, sd, update=25, burnin=5, seed=1)
    _, t_init = T(s.chain_init, 0, np.linspace(0.2, 0.8, K))
    _, t_sw = T(s.sweep, 25)
    _, t_close = T(s.close)
    print(f"rep{rep}: create+load {t_load:.1f} ms, chain_init {t_init:.1f}, 25 sweeps {t_sw:.1f}, destroy {t_close:.1f}")
    for upd in (2, 27):
        _, t = T(mcmc_updating, sd, Init(update=upd, burnin=1, thinning=1), 0, None, seed=1, device=0)
        print(f"   ig_mcmc_updating update={upd}: {t:.1f} ms")
