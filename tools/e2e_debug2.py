import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from instruct_b200 import Sampler, SeqData, Init, mcmc_updating, _lib
from instruct_b200.synth import make_dataset_torch
N, L, K = 10_000, 100_000, 8
dev = torch.device("cuda", 0)
x, an = make_dataset_torch(N, L, K, A=2, seed=4, device=dev)
torch.cuda.synchronize()
xh = torch.empty(x.shape, dtype=torch.int16, pin_memory=True); xh.copy_(x); torch.cuda.synchronize()
print("xh equals x:", bool((xh.to(dev) == x).all().item()), "neg frac", float((xh < 0).float().mean()))
for seed, burn in ((2024, 5), (2024, 1), (1, 5)):
    sd_h = SeqData(xh.numpy(), an.cpu().numpy(), K, ploid=2, mode=2, nstep_check_empty_cluster=10 ** 9)
    t0 = time.perf_counter()
    ch = mcmc_updating(sd_h, Init(update=55, burnin=burn, thinning=1), 0, None, seed=seed, device=0)
    print(f"seed {seed} burnin {burn}: {1e3 * (time.perf_counter() - t0):.1f} ms, steps {ch.step}, totallkh {ch.totallkh:.4g}, S {np.round(ch.self_rates, 3)}, qcol {np.round(ch.qq.sum(axis=0), 1)}", flush=True)
