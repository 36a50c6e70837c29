#!/bin/bash
# round-end measurement batch on one B200 (run under gpurun): the bench lines kept under profiles/
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --workload c5 --steps 30 --warmup 5 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
python bench.py --workload c5a --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_c5a.json 2> gpurun_out/r2_bench_c5a.err
for w in c1 c2 c3; do python bench.py --workload $w --steps 2000 --warmup 50 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; done
for f in c4 reference_arm c5 c5a c1 c2 c3; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().split("\n")[-1])
    e = d.get("e2e") or {}
    print("$f", "ms/step", round(d.get("ms_per_step", 0), 5), "value %.3e" % d.get("value", 0), "frac", (d.get("roofline") or {}).get("frac"), "e2e %.3e" % (e.get("value") or 0), "cpu", (d.get("cpu_baseline") or {}).get("value"), "conc", (d.get("concurrent_chains") or {}).get("us_per_sweep_per_chain"))
except Exception as ex:
    print("$f FAILED", ex)
PY
done
