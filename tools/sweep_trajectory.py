"""Per-sweep device time along a chain (CUDA events on the library's stream), config 4."""
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
sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, mode=2)
for ug in (0, 2):
    s = Sampler(sd, seed=1, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr(), use_graph=ug)
    s.chain_init(0, np.linspace(0.2, 0.8, K))
    out = []
    for r in range(16):
        t0 = time.perf_counter()
        ms = s.time_sweeps(5)
        wall = (time.perf_counter() - t0) * 1e3
        out.append((round(ms / 5, 3), round(wall / 5, 3)))
    print("use_graph", ug, out)
    print("  alpha", s.get(_lib.STATE_ALPHA), "S", s.get(_lib.STATE_S))
    s.close()
