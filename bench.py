#!/usr/bin/env python
"""bench.py -- throughput of the InStruct MCMC sweep on B200 (BASELINE.json metric).

One "step" is one MCMC sweep (one pass of the for(step...) body, mcmc.c:208-235) over the
resident genotype store.  The default workload is BASELINE.json configs[3]: the SNP-scale
single chain K=8, N=10,000, L=100,000, diploid, mode 2 -- the configuration the roofline
target is quoted on, and it fits one GPU (4 GB genotype store + 2 GB Z).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference ...                           the reference's CPU sweep

N > 1 (torchrun, one rank per GPU): ``--shard individuals`` (default) shards the individuals
of the ONE chain over the ranks -- strong scaling, one int32 NCCL all-reduce of n[L][A][K] and
one all-gather of the per-individual records per sweep; ``--shard chains`` runs one
independent chain per rank (weak scaling, no communication); ``--shard groups --group-size G`` runs
N/G independent chains, each sharded over G GPUs with its own NCCL communicator (BASELINE.json
configs[4]: ``--workload c5 --gpus 8 --shard groups --group-size 2`` = 4 chains x 2 GPUs each).

Output: ONE JSON line on rank 0 (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, L, K, A, miss, mode)
    "c4": (10_000, 100_000, 8, 2, 0.0, 2),      # BASELINE.json configs[3]
    "c2": (2_000, 200, 5, 6, 0.05, 2),          # configs[1] (per chain)
    "c1": (200, 10, 2, 8, 0.0, 2),              # configs[0]
    "c3": (5_000, 1_000, 4, 4, 0.0, 3),         # configs[2] (DP prior)
    "tiny": (600, 256, 8, 2, 0.0, 2),
    "c5": (20_000, 20_000, 6, 4, 0.0, 2),       # configs[4]: autotetraploid (-p 4 -ap 1), per chain
    "tiny5": (500, 128, 6, 4, 0.0, 2),
    "c5a": (20_000, 20_000, 6, 4, 0.0, 2),      # the configs[4] shape under the allotetraploid model (-p 4 -ap 0)
    "tiny5a": (500, 128, 6, 4, 0.0, 2),
}
TETRA = {"c5", "tiny5", "c5a", "tiny5a"}
ALLO = {"c5a", "tiny5a"}


def measured_traffic(workload, N, L):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the current build
    (profiles/r2_traffic.json; round 1's file if that is missing), valid for the workload's default shape only."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            if workload in t and (N, L) == WORKLOADS[workload][:2]:
                return float(t[workload]["bytes"])
        except Exception:
            pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            # let the tool exit and the driver finish tearing its client down: while that is going on, cudaMalloc /
            # cudaFree of other processes stall for hundreds of milliseconds (seen in the end-to-end call that follows)
            self.proc.wait(timeout=10)
            time.sleep(1.0)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# the reference's CPU sweep (oracle/_ref harness around the unmodified mcmc.c), bounded sample
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, steps, warmup, budget_s=20.0):
    """Times refh_sweeps() -- the reference's own update_P / update_S_POP / update_G /
    update_ZQ / update_alpha / cal_lkh in its own order -- on a bounded sample of the
    workload (same K, A, missing rate and generative model; fewer individuals and loci so a
    step takes ~1-2 s).  Single-threaded: one chain of the reference cannot use more cores."""
    import numpy as np
    from instruct_b200.synth import make_dataset
    from oracle import pyoracle

    N, L, K, A, miss, mode = WORKLOADS[workload]
    if workload in TETRA:
        return cpu_reference_sample_tetra(workload, steps, warmup, budget_s)
    # ~2.5 M allele copies per sweep at ~2.5 M copy-updates/s ~= 1 s per step
    target = 1_250_000
    n = min(N, 500)
    l = max(8, min(L, target // n))
    d = make_dataset(N=n, L=l, K=K, A=A, miss=miss, seed=4)
    kind = "reference" if pyoracle.have_ref() else "port"
    if kind == "reference":
        eng = pyoracle.Reference(d.x, d.allelenum, K, mode=mode)
        eng.setseeds(13, 4, 1972)
        eng.set_self(np.linspace(0.2, 0.8, K if mode == 2 else n))
        eng.set_gen(np.ones(n, dtype=np.int32))
        eng.update_ZQ(1)
    else:
        eng = pyoracle.Oracle(d.x, d.allelenum, K, mode=mode)
        eng.self_rates[...] = np.linspace(0.2, 0.8, eng.self_rates.size)
        eng.update_ZQ(1)
    copies = float((~(d.x < 0).any(axis=2)).sum() * 2)
    eng.sweeps(max(warmup, 1))
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        eng.sweeps(1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": copies * done / dt, "unit": "copy-updates/s", "cores": 1, "kind": kind,
            "sample": f"N={d.N} L={d.L} K={K} A={A} miss={miss} mode={mode}: {done} sweeps in {dt:.2f} s "
                      f"({done / dt:.3f} sweeps/s on the sample; one chain of the reference is single-threaded)",
            "sweeps_per_sec_sample": done / dt, "ms_per_step": 1e3 * dt / done, "steps": done}


def cpu_reference_sample_tetra(workload, steps, warmup, budget_s):
    """Autotetraploid: whole short chains through the reference's own mcmc_POP_tetra_selfing (poly_geno.c:75)
    at two lengths, differenced so that setup cancels -- the reference's sweep makes 2K+2 passes over the data."""
    import numpy as np
    from instruct_b200.synth import make_tetra_dataset
    from oracle import pyoracle, pytetra

    N, L, K, A, miss, mode = WORKLOADS[workload]
    n, l = min(N, 300), min(L, 100)
    d = make_tetra_dataset(N=n, L=l, K=K, A=A, miss=miss, seed=4)
    copies = float((d.nd > 0).sum() * 4)
    kind = "reference" if pyoracle.have_ref() else "port"
    initd = np.linspace(0.2, 0.8, K)

    def run(u):
        if kind == "reference":
            e = pytetra.RefTetra(d.x, d.nd, d.allelenum, K, autopoly=0 if workload in ALLO else 1)
            e.setseeds(13, 4, 1972)
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            sys.stdout.flush()
            os.dup2(devnull, 1)                     # the reference prints a banner per chain
            try:
                t0 = time.perf_counter()
                e.run_chain(u, 1, 1, ckrep=1, initd=initd)
                return time.perf_counter() - t0
            finally:
                os.dup2(saved, 1); os.close(devnull); os.close(saved)
        e = pytetra.TetraOracle(d.x, d.nd, d.allelenum, K, autopoly=0 if workload in ALLO else 1)
        t0 = time.perf_counter()
        e.run_chain(u, 1, 1, ckrep=1, initd=initd)
        return time.perf_counter() - t0
    u1 = 2
    t1 = run(u1)
    per = max(t1 / u1, 1e-3)
    u2 = u1 + max(2, min(steps, int(budget_s / per)))
    t2 = run(u2)
    done, dt = u2 - u1, max(t2 - t1, 1e-9)
    return {"value": copies * done / dt, "unit": "copy-updates/s", "cores": 1, "kind": kind,
            "sample": f"{'allo' if workload in ALLO else 'auto'}tetraploid N={d.N} L={d.L} K={K} A={A}: {done} sweeps in {dt:.2f} s (chains of {u1} and {u2} sweeps "
                      f"differenced; one chain of the reference is single-threaded)",
            "sweeps_per_sec_sample": done / dt, "ms_per_step": 1e3 * dt / done, "steps": done}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_sample(args.workload, args.steps, args.warmup, budget_s=120.0)
    N, L, K, A, miss, mode = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "genotype_copy_updates_per_sec", "value": cb["value"], "unit": "copy-updates/s",
        "n_gpus": args.gpus, "steps": cb["steps"], "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if args.shard == "individuals" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: K={K} N={N} L={L} A={A} " + (("allotetraploid" if args.workload in ALLO else "autotetraploid") if args.workload in TETRA else f"diploid mode {mode}") + " (bounded sample, see cpu_baseline.sample)"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "copy-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def _barrier_time(dist, dev):
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    return time.perf_counter()


def measure_e2e(args, x, an, K, N, mode, ploid, tetra, nloc, srank, count, chain, gsz, grp, rank, world, local, dist, copies_total):
    """One chain through the public API from HOST buffers, at N GPUs.  N = 1: ``mcmc_updating()`` (the drop-in call).
    N > 1: what the sharded drop-in does on every rank -- context for its shard, H2D of the shard, communicator, the chain,
    CHAIN back on every rank.  Timed between barriers; the slowest rank counts."""
    import numpy as np
    import torch
    from instruct_b200 import Init, Sampler, SeqData, mcmc_updating
    from instruct_b200.shard import broadcast_unique_id

    dev = torch.device("cuda", local)
    xh = torch.empty(x.shape, dtype=torch.int16, pin_memory=True)
    xh.copy_(x)
    anh = an.cpu().numpy()
    torch.cuda.synchronize()
    upd = args.e2e_sweeps or 4 * (args.steps + args.warmup)
    burn = max(1, args.warmup)
    sd_h = SeqData(xh.numpy(), anh, K, ploid=ploid, mode=mode, prior_flag=1 if args.workload == "c3" else 0, alpha_dpm=2.0,
                   nstep_check_empty_cluster=10 ** 9, autopoly=0 if args.workload in ALLO else 1)
    times, ch, comm_s = [], None, []
    for _ in range(3):
        t0 = _barrier_time(dist, dev)
        if world == 1:
            ch = mcmc_updating(sd_h, Init(update=upd, burnin=burn, thinning=1), 0, None, seed=args.seed, device=local)
        else:
            sm = Sampler(sd_h, update=upd, burnin=burn, thinning=1, seed=args.seed, device=local, shard_rank=srank, shard_count=count,
                         totalsize=N, rng_rounds=args.rng_rounds)
            tc = time.perf_counter()
            uid = broadcast_unique_id(Sampler.unique_id, rank, src=chain * gsz, group=grp)
            sm.comm_init(uid)
            comm_s.append(time.perf_counter() - tc)
            ch, _ = sm.run_chain(chain, initd=np.linspace(0.2, 0.8, K) if mode == 2 else None)
            sm.close()
        t1 = time.perf_counter()
        dt = t1 - t0
        if dist is not None:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0].item())
        times.append(dt)
    med = sorted(times)[1]
    h2d = xh.numel() * 2 + anh.size * 4
    d2h = 8 * (N * (2 * K + 3) + 2 * K + 2)
    del xh
    return {"value": copies_total * upd / med, "unit": "copy-updates/s", "h2d_bytes_per_step": h2d / upd, "d2h_bytes_per_step": d2h / upd,
            "sweeps": upd, "seconds": med, "seconds_each_call": times, "statistic": "median of three calls",
            "retained_samples": int(ch.step), "posterior_mean_loglik": float(ch.totallkh), "n_gpus": world,
            "comm_init_seconds_each_call": comm_s if comm_s else None,
            "note": ("one ig_mcmc_updating() call: create + H2D of the pinned genotype store + all sweeps + D2H of CHAIN" if world == 1 else
                     "per rank: create + H2D of its pinned shard + NCCL communicator (a new one per call: comm_init_seconds_each_call, this "
                     "rank's share of the time; a multi-chain run creates it once) + all sweeps of the sharded chain + D2H of CHAIN; "
                     "between barriers, slowest rank; h2d_bytes_per_step is per rank")}


def check_shard_parity(args, K, mode, srank, count, chain, gsz, grp, rank, local, dist):
    """A 30-sweep chain on a small slice, sharded over this run's ranks and unsharded on one GPU, compared bit for bit
    (posterior moments of Q, S, G, the log-likelihood trace).  True / False on rank 0."""
    import numpy as np
    import torch
    from instruct_b200 import Sampler, SeqData
    from instruct_b200.shard import broadcast_unique_id, shard_bounds
    from instruct_b200.synth import make_dataset

    Ns, Ls = 64 * count + 7, 777                       # ragged on purpose
    d = make_dataset(N=Ns, L=Ls, K=K, A=2, miss=0.02, seed=99)
    b, e = shard_bounds(Ns, count, srank)
    kw = dict(update=30, burnin=10, thinning=2, ckrep=5, seed=4242, device=local)
    sd_s = SeqData(np.ascontiguousarray(d.x[:, b:e, :]), d.allelenum, K, mode=mode if mode in (1, 2, 3) else 2)
    sm = Sampler(sd_s, shard_rank=srank, shard_count=count, totalsize=Ns, **kw)
    sm.comm_init(broadcast_unique_id(Sampler.unique_id, rank, src=chain * gsz, group=grp))
    initd = np.linspace(0.2, 0.8, K)
    ch_s, cv_s = sm.run_chain(0, initd=initd)
    sm.close()
    ok = True
    if srank == 0:
        s1 = Sampler(SeqData(d.x, d.allelenum, K, mode=sd_s.mode), **kw)
        ch_1, cv_1 = s1.run_chain(0, initd=initd)
        s1.close()
        ok = (ch_s.totallkh == ch_1.totallkh and np.array_equal(ch_s.qq, ch_1.qq) and np.array_equal(ch_s.self_rates, ch_1.self_rates)
              and np.array_equal(ch_s.gen, ch_1.gen) and np.array_equal(ch_s.indvlkh, ch_1.indvlkh) and np.array_equal(cv_s, cv_1))
    t = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t[0].item()))


def measure_concurrent_chains(args, x, an, K, N, L, mode, ploid, local, copies, C=8):
    """Small data sets (configs 1-2) are launch- and latency-bound: one chain's sweep leaves most SMs idle.  BASELINE's
    config 2 is 8 chains; here C chains run CONCURRENTLY on the one GPU, each on its own context and stream from its own
    host thread (what `inbreed --chains-per-gpu` does), and the aggregate rate is reported."""
    import threading
    import numpy as np
    from instruct_b200 import Sampler, SeqData

    shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(L, N, ploid), strides=(0, 0, 0))
    sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, ploid=ploid, mode=mode, prior_flag=1 if args.workload == "c3" else 0, alpha_dpm=2.0)
    ss = []
    for c in range(C):
        s = Sampler(sd, seed=args.seed, device=local, rng_rounds=args.rng_rounds, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
        s.chain_init(c, initd=np.linspace(0.2, 0.8, K) if mode == 2 else None)
        s.sweep(args.warmup)
        s.sync()
        ss.append(s)
    bar = threading.Barrier(C + 1)
    def work(s):
        bar.wait()
        s.sweep(args.steps)
        s.sync()
        bar.wait()
    th = [threading.Thread(target=work, args=(s,)) for s in ss]
    for t in th:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    for s in ss:
        s.close()
    return {"chains": C, "sweeps_per_chain": args.steps, "seconds": dt, "us_per_sweep_per_chain": 1e6 * dt / (args.steps * C),
            "value": copies * C * args.steps / dt, "unit": "copy-updates/s",
            "note": "C independent chains in flight on ONE GPU (own context, stream and host thread each); wall clock between barriers"}


def measure_chains(args, N, L, K, A, miss, mode, tetra, ploid, rank, world, local, dist):
    """The other way the path shards: one independent chain per GPU on the whole data set, no communication.  Same
    workload, same step count, timed between barriers on the slowest rank."""
    import numpy as np
    import torch
    from instruct_b200 import Sampler, SeqData
    from instruct_b200.synth import make_dataset_torch, make_tetra_dataset_torch

    dev = torch.device("cuda", local)
    if tetra:
        x, an = make_tetra_dataset_torch(N, L, K, A=A, miss=miss, seed=4 + rank, device=dev)
        usable = float((x[:, :, 0] >= 0).sum().item())
    else:
        x, an = make_dataset_torch(N, L, K, A=A, miss=miss, seed=4 + rank, device=dev)
        usable = float((~(x < 0).any(dim=2)).sum().item())
    torch.cuda.synchronize()
    shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(L, N, ploid), strides=(0, 0, 0))
    sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, ploid=ploid, mode=mode, autopoly=0 if args.workload in ALLO else 1)
    s = Sampler(sd, seed=args.seed, device=local, rng_rounds=args.rng_rounds, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
    s.chain_init(rank, initd=np.linspace(0.2, 0.8, K) if mode == 2 else None)
    s.sweep(args.warmup)
    s.sync()
    dist.barrier()
    torch.cuda.synchronize()
    ms = s.time_sweeps(args.steps)
    t = torch.tensor([ms, ploid * usable], device=dev, dtype=torch.float64)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s.close()
    ms = float(mx[0].item())
    return {"parallelism": "chains x%d (one independent chain per GPU, no communication)" % world, "ms_per_step": ms / args.steps,
            "value": float(t[1].item()) * args.steps / (ms * 1e-3), "unit": "copy-updates/s", "scaling": "weak"}


def run_ours(args):
    import numpy as np
    import torch

    from instruct_b200 import Sampler, SeqData, _lib
    from instruct_b200.shard import shard_bounds, broadcast_unique_id, group_layout, make_groups
    from instruct_b200.synth import make_dataset_torch, make_tetra_dataset_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    N, L, K, A, miss, mode = WORKLOADS[args.workload]
    if args.N:
        N = args.N
    if args.L:
        L = args.L
    tetra = args.workload in TETRA
    ploid = 4 if tetra else 2
    shard_ind = world > 1 and args.shard in ("individuals", "groups")
    # --shard groups: world / G independent chains, each with its individuals sharded over G GPUs and its own
    # NCCL communicator (the composition SURVEY.md section 8e names for configs[4]: 4 chains x 2 GPUs)
    gsz = world if args.shard == "individuals" else max(1, min(args.group_size, world))
    if shard_ind and world % gsz:
        raise SystemExit(f"--group-size {gsz} does not divide the {world} ranks")
    chain_of_rank, grp = rank, None
    if shard_ind:
        chain_of_rank, srank, _ = group_layout(world, rank, gsz)
        b, e = shard_bounds(N, gsz, srank)
        nloc, i0, count, seed_data = e - b, b, gsz, 4 + chain_of_rank
        grp = make_groups(world, rank, gsz)
    else:
        nloc, i0, count, srank, seed_data = N, 0, 1, 0, 4 + rank
    if tetra:
        x, an = make_tetra_dataset_torch(N, L, K, A=A, miss=miss, seed=seed_data, device=dev)
        if shard_ind:
            x = x[:, i0:i0 + nloc, :].contiguous()     # every rank generates the same data set and keeps its block
        torch.cuda.synchronize()
        usable = float((x[:, :, 0] >= 0).sum().item())
    else:
        x, an = make_dataset_torch(N, L, K, A=A, miss=miss, seed=seed_data, device=dev, i0=i0, n_local=nloc)
        torch.cuda.synchronize()
        usable = float((~(x < 0).any(dim=2)).sum().item())
    copies_local = float(ploid) * usable
    store_gb = (x.numel() * 2 + x.numel() * (2 if tetra else 1)) / 1e9
    # the genotype store is passed by device pointer (inputs resident in HBM); this SeqData only
    # carries the flags and the (L, Nloc, ploid) shape, through a zero-strided placeholder
    shape_only = np.lib.stride_tricks.as_strided(np.zeros(1, dtype=np.int16), shape=(L, nloc, ploid), strides=(0, 0, 0))
    sd = SeqData(shape_only, np.zeros(L, dtype=np.int32), K, ploid=ploid, mode=mode, prior_flag=1 if args.workload == "c3" else 0,
                 alpha_dpm=2.0, autopoly=0 if args.workload in ALLO else 1)
    s = Sampler(sd, seed=args.seed, device=local, shard_rank=srank, shard_count=count, totalsize=N,
                rng_rounds=args.rng_rounds, x_device_ptr=x.data_ptr(), allelenum_device_ptr=an.data_ptr())
    if shard_ind:
        uid = broadcast_unique_id(Sampler.unique_id, rank, src=chain_of_rank * gsz, group=grp)
        s.comm_init(uid)
    s.chain_init(chain_of_rank, initd=np.linspace(0.2, 0.8, K) if mode == 2 else None)
    s.sweep(args.warmup)
    s.sync()
    # per-launch events around the dominant kernel ride inside the timed region for the HBM-sized
    # workloads; the small ones are launch-bound and replayed as a CUDA graph (no events inside a
    # graph), so their kernel share is measured on a second, directly launched run of the same length
    inline_profile = L * N >= 50_000_000 or world > 1
    if inline_profile:
        s.profile(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                            # before the barrier: starting the sampler must not delay rank 0 into the timed region
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    n0, _, k0 = s.profile_read()
    ms = s.time_sweeps(args.steps)               # CUDA events on the library's stream, sync both sides
    torch.cuda.synchronize()
    nz, zq_ms, k1 = s.profile_read()
    ms_direct = ms
    if not inline_profile:
        s.profile(True)
        ms_direct = s.time_sweeps(args.steps)
        torch.cuda.synchronize()
        nz, zq_ms, _ = s.profile_read()
    if dist is not None:
        t = torch.tensor([ms, 0.0], device=dev, dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        tot = torch.tensor([copies_local], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms = float(mx[0].item())
        copies_total = float(tot[0].item())
        dist.barrier()
    else:
        copies_total = copies_local
    clk = clocks.stop() if rank == 0 else None
    # 4 B per allele copy (6 B for ploid 4: + latent genotype read and write) x copies one launch processes
    algo_bytes_launch = (6.0 if tetra else 4.0) * copies_local
    zq_avg_ms = zq_ms / max(nz, 1)
    peak, peak_src = peaks()
    achieved = algo_bytes_launch / (zq_avg_ms * 1e-3) / 1e9 if zq_avg_ms > 0 else 0.0
    sweeps_per_s = args.steps / (ms * 1e-3)
    value = copies_total * sweeps_per_s

    geo = s.geometry()
    s.close()                                      # the timed context is done: release its buffers before the end-to-end call
    # ---- e2e: the drop-in call with HOST buffers.  Every rank holds its shard of the genotype store in pinned host memory;
    # the timed region is create + H2D of the store + (N > 1: communicator) + the whole chain + D2H of CHAIN, between
    # barriers, max over ranks.  Three calls, the MEDIAN is reported and all three are kept in the line.
    e2e = None
    if not args.no_e2e and (world == 1 or shard_ind):
        e2e = measure_e2e(args, x, an, K, N, mode, ploid, tetra, nloc, srank, count, chain_of_rank, gsz, grp, rank, world, local, dist,
                          copies_total)
    conc = None
    if world == 1 and not tetra and N * L <= 2_000_000 and not args.no_concurrent:
        conc = measure_concurrent_chains(args, x, an, K, N, L, mode, ploid, local, copies_local)
    shard_parity = None
    if world > 1 and shard_ind and not tetra:
        shard_parity = check_shard_parity(args, K, mode, srank, count, chain_of_rank, gsz, grp, rank, local, dist)
    chains_rec = None
    if world > 1 and shard_ind and gsz == world and not args.no_chains:
        del x
        torch.cuda.empty_cache()
        chains_rec = measure_chains(args, N, L, K, A, miss, mode, tetra, ploid, rank, world, local, dist)

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu:
            cb = cpu_reference_sample(args.workload, steps=12, warmup=1, budget_s=15.0)
        line = {
            "metric": "genotype_copy_updates_per_sec", "value": value, "unit": "copy-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (world == 1 or (shard_ind and gsz == world)) else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "sweeps_per_sec": sweeps_per_s,
            "config": {"workload": f"{args.workload}: K={K} N={N} L={L} A={A} " + (("allotetraploid" if args.workload in ALLO else "autotetraploid") if tetra else f"diploid mode {mode}") + f" miss={miss}",
                       "parallelism": ("individual-sharded x%d" % world) if (shard_ind and gsz == world) else
                                      (("%d chains x %d GPUs each (individual-sharded)" % (world // gsz, gsz)) if shard_ind else ("chains x%d" % world)),
                       "l2": "inputs (%.2f GB per GPU) larger than L2" % store_gb,
                       "geometry": geo, "rng": f"philox4x32-{args.rng_rounds or 7} (Z draw), philox4x32-10 (all other draws)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, N, L) if world == 1 else None, "kernel": "tetra_zs + tetra_geno (the two passes of one sweep)" if tetra else "zq_sweep", "launches_timed": nz, "avg_launch_ms": zq_avg_ms,
                         "algorithmic_bytes_per_launch": algo_bytes_launch, "peak_source": peak_src,
                         "share_of_step": zq_ms / ms_direct if ms_direct > 0 else None,
                         "graph_replay": (not inline_profile) and world == 1 and not tetra and args.workload != "c3", "ms_per_step_direct_launch": ms_direct / args.steps},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None,
            "e2e": e2e, "gpu_launches": int(k1 - k0), "clocks": clk,
            "shard_parity": shard_parity, "chains": chains_rec, "concurrent_chains": conc,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="individuals", choices=["individuals", "chains", "groups"])
    ap.add_argument("--group-size", type=int, default=2, help="--shard groups: GPUs per chain (BASELINE configs[4]: 4 chains x 2 GPUs)")
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--rng-rounds", type=int, default=0)
    ap.add_argument("--N", type=int, default=0)
    ap.add_argument("--L", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-concurrent", action="store_true", help="small workloads: skip the concurrent-chains sub-record")
    ap.add_argument("--no-chains", action="store_true", help="N > 1: skip the chain-partitioned sub-record")
    ap.add_argument("--e2e-sweeps", type=int, default=0, help="sweeps of the end-to-end chain (default 4 x (steps + warmup))")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
