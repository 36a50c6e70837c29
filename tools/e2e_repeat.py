"""Repeatability of the end-to-end call at config 4 (tools, not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from instruct_b200 import SeqData, Init, mcmc_updating
from instruct_b200.synth import make_dataset_torch
N, L, K = 10_000, 100_000, 8
dev = torch.device("cuda", 0)
x, an = make_dataset_torch(N, L, K, A=2, seed=4, device=dev)
xh = torch.empty(x.shape, dtype=torch.int16, pin_memory=True); xh.copy_(x); torch.cuda.synchronize()
anh = an.cpu().numpy()
keep = os.environ.get("KEEP_X", "0") == "1"
if not keep:
    del x
sd = SeqData(xh.numpy(), anh, K, mode=2, nstep_check_empty_cluster=10 ** 9)
for upd in (2, 27, 27, 27, 55, 55, 2):
    t = time.perf_counter()
    mcmc_updating(sd, Init(update=upd, burnin=1, thinning=1), 0, None, seed=1, device=0)
    print(f"update={upd}: {(time.perf_counter() - t) * 1e3:.1f} ms", flush=True)
