#!/usr/bin/env python
"""One rank's kernels of a sharded config-4 chain with LOCAL scalar updates, on one GPU: a one-rank NCCL communicator
(IG_COMM_SINGLE=1) over N = the shard's individuals, so every kernel runs on the shape it has at `10000 / N` GPUs and the
collectives cost their launch only.  Run under ncu's launch list to see what each kernel of such a sweep costs."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["IG_COMM_SINGLE"] = "1"
import numpy as np
import torch
from instruct_b200 import Sampler, SeqData
from instruct_b200.synth import make_dataset_torch

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1250)
ap.add_argument("--L", type=int, default=100000)
ap.add_argument("--K", type=int, default=8)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
x, an = make_dataset_torch(a.N, a.L, a.K, A=2, miss=0.0, seed=4, device=dev)
torch.cuda.synchronize()
shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(a.L, a.N, 2), strides=(0, 0, 0))
sd = SeqData(shape_only, np.zeros(a.L, dtype=np.int32), a.K, mode=2)
s = Sampler(sd, seed=1, device=0, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
s.comm_init(Sampler.unique_id())
s.chain_init(0, initd=np.linspace(0.2, 0.8, a.K))
s.sweep(5); s.sync()
ms = s.time_sweeps(a.steps)
print(f"N={a.N}: {ms / a.steps * 1e3:.1f} us per sweep, one-rank communicator")
s.close()
