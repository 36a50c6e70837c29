#!/bin/bash
# 8-GPU measurement of the sharded config-4 chain: the tally -> P exchange over peer memory (default) against the NCCL
# reduce-scatter / all-gather path, short runs; then the full bench line (N-GPU e2e, shard_parity, chains) with the faster one.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
mkdir -p gpurun_out
IG_PHASE_TRACE=1 timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 5 --no-e2e --no-chains 2>gpurun_out/x8_peerp_err.log | tail -1 > gpurun_out/x8_peerp.json
IG_P_NCCL=1 IG_PHASE_TRACE=1 timeout 300 $TR --master-port 29532 bench.py --gpus 8 --steps 100 --warmup 5 --no-e2e --no-chains 2>gpurun_out/x8_ncclp_err.log | tail -1 > gpurun_out/x8_ncclp.json
A=$(python -c "import json;print(json.load(open('gpurun_out/x8_peerp.json'))['ms_per_step'])")
B=$(python -c "import json;print(json.load(open('gpurun_out/x8_ncclp.json'))['ms_per_step'])")
echo "peer P: $A ms   NCCL P: $B ms"
if python -c "import sys; sys.exit(0 if $A <= $B else 1)"; then unset IG_P_NCCL; echo "full line: peer P"; else export IG_P_NCCL=1; echo "full line: NCCL P"; fi
timeout 500 $TR --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 5 2>gpurun_out/x8_full_err.log | tail -1 > gpurun_out/x8_full.json
cut -c1-260 gpurun_out/x8_full.json
