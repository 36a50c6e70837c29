#!/usr/bin/env python
"""One rank's view of an individual-sharded config-4 chain WITHOUT the communicator (timing only): rank 0 of `--count`
shards runs its kernels on its 1/count of the individuals, the O(N K) scalar kernels on all N records.  Used under ncu's
launch list to see what each kernel of a sharded sweep costs on one GPU; the NCCL calls are the part it leaves out."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from instruct_b200 import Sampler, SeqData
from instruct_b200.synth import make_dataset_torch

ap = argparse.ArgumentParser()
ap.add_argument("--count", type=int, default=8)
ap.add_argument("--N", type=int, default=10000)
ap.add_argument("--L", type=int, default=100000)
ap.add_argument("--K", type=int, default=8)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cap = -(-a.N // a.count)
x, an = make_dataset_torch(a.N, a.L, a.K, A=2, miss=0.0, seed=4, device=dev, i0=0, n_local=cap)
torch.cuda.synchronize()
shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(a.L, cap, 2), strides=(0, 0, 0))
sd = SeqData(shape_only, np.zeros(a.L, dtype=np.int32), a.K, mode=2)
s = Sampler(sd, seed=1, device=0, shard_rank=0, shard_count=a.count, totalsize=a.N, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
s.chain_init(0, initd=np.linspace(0.2, 0.8, a.K))
s.sweep(5); s.sync()
ms = s.time_sweeps(a.steps)
print(f"shard 0 of {a.count}: {ms / a.steps * 1e3:.1f} us per sweep without the collectives, geometry {s.geometry()}")
s.close()
